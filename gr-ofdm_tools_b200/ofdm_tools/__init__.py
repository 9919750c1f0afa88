"""ofdm_tools -- B200-native drop-in for the OFDM sample-path blocks of gr-ofdm_tools.

Mirrors the export names of the reference package for the hot path only
(python/__init__.py:49-84 of the reference): ofdm_tx_rx_hier, ofdm_radio_hier, sync_radio_hier,
ofdm_txrx_modules.{ofdm_tx, ofdm_rx}, payload_source, payload_sink, and the next rows of SURVEY.md 8(f):
clipper, papr_sink, payload_source_pdu, payload_sink_pdu, ofdm_cr_tools.{make_packet, unmake_packet}, crc (the MAC-level
CRC-32 of gnuradio.digital.crc).
"""
from .phy import OfdmPhy, RxResult, FRAME_DTYPE  # noqa: F401
from . import ofdm_txrx_modules  # noqa: F401
from .ofdm_txrx_modules import ofdm_tx, ofdm_rx  # noqa: F401
from .ofdm_tx_rx_hier import ofdm_tx_rx_hier  # noqa: F401
from .ofdm_radio_hier import ofdm_radio_hier  # noqa: F401
from .sync_radio_hier import sync_radio_hier  # noqa: F401
from .payload_source import payload_source  # noqa: F401
from .payload_sink import payload_sink  # noqa: F401
from . import ofdm_cr_tools  # noqa: F401
from .payload_source_pdu import payload_source_pdu  # noqa: F401
from .payload_sink_pdu import payload_sink_pdu  # noqa: F401
from .clipper import clipper  # noqa: F401
from .papr_sink import papr_sink  # noqa: F401
from . import crc  # noqa: F401

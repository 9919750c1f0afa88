"""B200 drop-in for python/ofdm_txrx_modules.py of the reference: `ofdm_tx` and `ofdm_rx`.

Same constructor parameters, defaults and ValueError conditions
(python/ofdm_txrx_modules.py:143-187,278-322).  Instead of wiring ~12 GNU Radio blocks each, both
classes drive one batched call into libofdmx.so (`work()`); see INTEGRATION.md for the
gr.basic_block adaptor that binds `work()` into a flowgraph.
"""
import numpy as np

from .phy import OfdmPhy, _make_sync_word1, _make_sync_word2, _get_active_carriers  # noqa: F401

_def_fft_len = 64
_def_cp_len = 16
_def_frame_length_tag_key = "frame_length"
_def_packet_length_tag_key = "packet_length"
_def_packet_num_tag_key = "packet_num"
# Data and pilot carriers are the same as in 802.11a (python/ofdm_txrx_modules.py:54-62)
_def_occupied_carriers = (list(range(-26, -21)) + list(range(-20, -7)) + list(range(-6, 0))
                          + list(range(1, 7)) + list(range(8, 21)) + list(range(22, 27)),)
_def_pilot_carriers = ((-21, -7, 7, 21,),)
_pilot_sym_scramble_seq = (
    1, 1, 1, 1, -1, -1, -1, 1, -1, -1, -1, -1, 1, 1, -1, 1, -1, -1, 1, 1, -1, 1, 1, -1, 1, 1, 1, 1, 1, 1, -1, 1,
    1, 1, -1, 1, 1, -1, -1, 1, 1, 1, -1, 1, -1, -1, -1, 1, -1, 1, -1, -1, 1, -1, -1, 1, 1, 1, 1, 1, -1, -1, 1, 1,
    -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1, 1, -1, -1, -1, -1, 1, -1, -1, 1, -1, 1, 1, 1, 1, -1, 1, -1, 1, -1, 1,
    -1, -1, -1, -1, -1, 1, -1, 1, 1, -1, 1, -1, 1, 1, 1, -1, -1, 1, -1, -1, -1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1
)
_def_pilot_symbols = tuple([(x, x, x, -x) for x in _pilot_sym_scramble_seq])
_seq_seed = 42

_SUPPORTED_BPS = {1: "bpsk", 2: "qpsk", 3: "8psk", 4: "qam16", 6: "qam64"}


class _constellation(object):
    """Stand-in for the digital.constellation_* object `_get_constellation` returns
    (python/ofdm_txrx_modules.py:106-118): carries bits_per_symbol() and points()."""

    def __init__(self, bps):
        self._bps = bps

    def bits_per_symbol(self):
        return self._bps

    def base(self):
        return self

    def points(self):
        b = self._bps
        if b == 1:
            return [-1 + 0j, 1 + 0j]
        if b == 2:
            a = 0.707107
            return [complex(-a, -a), complex(a, -a), complex(-a, a), complex(a, a)]
        if b == 3:
            ang = np.float32(np.pi / 8.0)
            return [complex(np.cos(m * ang), np.sin(m * ang)) for m in (1, 7, 15, 9, 3, 5, 13, 11)]
        side = 2 if b == 4 else 4
        step = 1.0 / (side - 0.5)
        pts = []
        for i in range(1 << b):
            y, x, quad = i % side, (i // side) % side, i // (side * side)
            gx, gy = (x + 0.5) * step, (y + 0.5) * step
            pts.append([complex(gx, gy), complex(-gy, gx), complex(-gx, -gy), complex(gy, -gx)][quad])
        return pts


def _get_constellation(bps):
    """Returns a modulator description for a given number of bits per symbol.  The reference prints
    'Modulation not supported.' and exits (python/ofdm_txrx_modules.py:114-118); here: ValueError."""
    if bps not in _SUPPORTED_BPS:
        raise ValueError("Modulation not supported.")
    return _constellation(bps)


class _ofdm_base(object):
    def _setup(self, fft_len, cp_len, occupied_carriers, pilot_carriers, pilot_symbols, bps_header,
               bps_payload, sync_word1, sync_word2, scramble_bits, **extra):
        self.fft_len = fft_len
        self.cp_len = cp_len
        self.occupied_carriers = occupied_carriers
        self.pilot_carriers = pilot_carriers
        self.pilot_symbols = pilot_symbols
        self.bps_header = bps_header
        self.bps_payload = bps_payload
        if sync_word1 is None:
            self.sync_word1 = _make_sync_word1(fft_len, occupied_carriers, pilot_carriers)
        else:
            if len(sync_word1) != self.fft_len:
                raise ValueError("Length of sync sequence(s) must be FFT length.")
            self.sync_word1 = sync_word1
        if sync_word2 is None:
            self.sync_word2 = _make_sync_word2(fft_len, occupied_carriers, pilot_carriers)
        else:
            # sync_word2=() selects the one-sync-word mode, as in the reference (:174-183, :311-321)
            if len(sync_word2) and len(sync_word2) != fft_len:
                raise ValueError("Length of sync sequence(s) must be FFT length.")
            self.sync_word2 = list(sync_word2)
        self.scramble_seed = 0x7f if scramble_bits else 0x00
        _get_constellation(bps_header)
        _get_constellation(bps_payload)
        self.phy = OfdmPhy(fft_len=fft_len, cp_len=cp_len, occupied_carriers=occupied_carriers,
                           pilot_carriers=pilot_carriers, pilot_symbols=pilot_symbols,
                           sync_word1=self.sync_word1, sync_word2=self.sync_word2, bps_header=bps_header,
                           bps_payload=bps_payload, scramble_bits=scramble_bits, **extra)


class ofdm_tx(_ofdm_base):
    """OFDM modulation: byte packets in, complex baseband out.

    Args: as python/ofdm_txrx_modules.py:126-142 of the reference.  `rolloff` is the
    rolloff_len of digital.ofdm_cyclic_prefixer (:247-253): raised-cosine flanks, each burst rolloff-1 samples
    longer; `debug_log` is accepted and ignored as in the reference (:153)."""

    def __init__(self, fft_len=_def_fft_len, cp_len=_def_cp_len,
                 packet_length_tag_key=_def_packet_length_tag_key,
                 occupied_carriers=_def_occupied_carriers,
                 pilot_carriers=_def_pilot_carriers,
                 pilot_symbols=_def_pilot_symbols,
                 bps_header=1, bps_payload=1, sync_word1=None, sync_word2=None,
                 rolloff=0, debug_log=False, scramble_bits=False, **phy_kwargs):
        self.packet_length_tag_key = packet_length_tag_key
        self.rolloff = rolloff
        self._setup(fft_len, cp_len, occupied_carriers, pilot_carriers, pilot_symbols, bps_header,
                    bps_payload, sync_word1, sync_word2, scramble_bits, rolloff=int(rolloff), **phy_kwargs)
        self.sync_words = [self.sync_word1] + ([self.sync_word2] if len(self.sync_word2) else [])
        self._pkt_num = 0

    def work(self, packets):
        """One tagged-stream packet per list entry -> (samples cuda complex64, frame offsets)."""
        out = self.phy.tx(packets, first_pkt_num=self._pkt_num)
        self._pkt_num = (self._pkt_num + len(packets)) & 0xFFF      # packet_header_default counter
        return out


class ofdm_rx(_ofdm_base):
    """OFDM demodulation: complex baseband in, detected packets out
    (args: python/ofdm_txrx_modules.py:263-277)."""

    def __init__(self, fft_len=_def_fft_len, cp_len=_def_cp_len,
                 frame_length_tag_key=_def_frame_length_tag_key,
                 packet_length_tag_key=_def_packet_length_tag_key,
                 packet_num_tag_key=_def_packet_num_tag_key,
                 occupied_carriers=_def_occupied_carriers,
                 pilot_carriers=_def_pilot_carriers,
                 pilot_symbols=_def_pilot_symbols,
                 bps_header=1, bps_payload=1, sync_word1=None, sync_word2=None,
                 debug_log=False, scramble_bits=False, **phy_kwargs):
        self.frame_length_tag_key = frame_length_tag_key
        self.packet_length_tag_key = packet_length_tag_key
        self.packet_num_tag_key = packet_num_tag_key
        self._setup(fft_len, cp_len, occupied_carriers, pilot_carriers, pilot_symbols, bps_header,
                    bps_payload, sync_word1, sync_word2, scramble_bits, **phy_kwargs)

    def work(self, samples, **kw):
        """samples: complex64 cuda tensor [n] or [n_streams, n] -> RxResult."""
        return self.phy.rx(samples, **kw)

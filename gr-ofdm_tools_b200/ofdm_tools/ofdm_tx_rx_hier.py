"""B200 drop-in for python/ofdm_tx_rx_hier.py: `ofdm_tx_rx_hier(fft_len=64, payload_bps=2)`.

TX path: ofdm_tx -> tag_gate -> x0.01 (python/ofdm_tx_rx_hier.py:55-63,74,85-87); the x0.01 is
fused into the IFFT store.  RX path: analog.agc2_cc(1e-1, 1e-2, 1.0, 1.0) with max gain 65536 (:75-76,
:82-83) -> ofdm_rx (:64-73).  The AGC is a per-sample non-linear recurrence: streams run in parallel, the
samples of one stream sequentially; its loop gain is carried from one rx() call to the next (pass
agc=False to skip it for long, already normalised single streams).
"""
from . import ofdm_txrx_modules


class ofdm_tx_rx_hier(object):
    def __init__(self, fft_len=64, payload_bps=2, **phy_kwargs):
        self.fft_len = fft_len
        self.payload_bps = payload_bps
        self.len_tag_key = len_tag_key = "packet_len"
        self.ofdm_tx = ofdm_txrx_modules.ofdm_tx(
            fft_len=fft_len, cp_len=fft_len // 4, packet_length_tag_key=len_tag_key, bps_header=1,
            bps_payload=payload_bps, rolloff=0, debug_log=False, scramble_bits=False, tx_scale=0.01,
            **phy_kwargs)
        self.ofdm_rx = ofdm_txrx_modules.ofdm_rx(
            fft_len=fft_len, cp_len=fft_len // 4, frame_length_tag_key='frame_' + "rx_len",
            packet_length_tag_key=len_tag_key, bps_header=1, bps_payload=payload_bps, debug_log=False,
            scramble_bits=False, **phy_kwargs)

    # port 0 in -> port 1 out
    def tx(self, packets):
        return self.ofdm_tx.work(packets)

    # port 1 in -> port 0 out
    def rx(self, samples, agc=True, **kw):
        if agc:
            samples, self._agc_gain = self.ofdm_rx.phy.agc2(samples, getattr(self, "_agc_gain", None),
                                                            1e-1, 1e-2, 1.0, 65536.0)
        return self.ofdm_rx.work(samples, **kw)

    def get_fft_len(self):
        return self.fft_len

    def set_fft_len(self, fft_len):
        self.fft_len = fft_len

    def get_payload_bps(self):
        return self.payload_bps

    def set_payload_bps(self, payload_bps):
        self.payload_bps = payload_bps

    def get_len_tag_key(self):
        return self.len_tag_key

    def set_len_tag_key(self, len_tag_key):
        self.len_tag_key = len_tag_key

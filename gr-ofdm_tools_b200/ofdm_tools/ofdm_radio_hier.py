"""B200 drop-in for python/ofdm_radio_hier.py: `ofdm_radio_hier`.

Keeps the constructor parameters and defaults of the reference (python/ofdm_radio_hier.py:34-39):
the carrier plan and 128-point sync words below are those defaults (they equal
spectrum_enforcer(128, [], 10), python/ofdm_cr_tools.py:348-378).  Derived values follow
:73-86: fft_len = (len(sync_word1)+len(sync_word2))/2, cp_len = fft_len/4, BPSK header with
scramble_header=True, chanest max_carr_offset=3, x0.01 TX scaling.

Selectors (:137-178) become plain flags: scramble_mode (payload scrambler 0x7f), crc_mode (in-graph
crc32_bb on both paths), clipper_mode (ofdm_tools.clipper(clipping_factor) after the x0.01 scaling, :92,
:229,:239 -- fused into the TX kernel as tx_clip).  The RX AGC (analog.agc2_cc(1e-1, 1e-2, 1.0, 1.0),
max gain 65536, :180-181) runs in front of the receiver as in the reference; its loop gain is carried from
one rx() call to the next.  It is a per-sample non-linear recurrence: streams are processed in parallel,
the samples of one stream sequentially, so pass agc=False to rx() for long single streams whose level is
already normalised.  filter_mode=1 (the default) runs the TX burst stream through the 8th-order
iir_filter_ccd(forward_OOB, feedback_OOB) out-of-band filter (:83-84,:93,:232-237) on the GPU
(OfdmPhy.iir_ccd -> ofdmx_iir_ccd), its history carried across tx() calls; filter_mode=0 bypasses it.
"""
from .phy import OfdmPhy

_BPS = {'bpsk': 1, 'qpsk': 2, '8psk': 3, 'qam16': 4, 'qam64': 6}

_DEF_PILOT_CARRIERS = ((-40, -14, 13, 39),)
_DEF_PILOT_SYMBOLS = ((1, 1, 1, -1),)
_DEF_OCCUPIED_CARRIERS = ([-54, -53, -52, -51, -50, -49, -48, -47, -46, -45, -44, -43, -42, -41, -39, -38, -37, -36, -35, -34, -33, -32, -31, -30, -29, -28, -27, -26, -25, -24, -23, -22, -21, -20, -19, -18, -17, -16, -15, -13, -12, -11, -10, -9, -8, -7, -6, -5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53],)
_DEF_SYNC_WORD1 = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, -1.42, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
_DEF_SYNC_WORD2 = [0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, (-1+0j), (1+0j), (-1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), 0j, (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j]


class ofdm_radio_hier(object):

    def __init__(self, pilot_carriers=_DEF_PILOT_CARRIERS, pilot_symbols=_DEF_PILOT_SYMBOLS,
                 occupied_carriers=_DEF_OCCUPIED_CARRIERS, samp_rate=10000, payload_mod='qpsk',
                 sync_word1=_DEF_SYNC_WORD1, sync_word2=_DEF_SYNC_WORD2,
                 scramble_mode=0, crc_mode=0, clipper_mode=0, filter_mode=1, clipping_factor=10,
                 **phy_kwargs):
        self.pilot_carriers = pilot_carriers
        self.pilot_symbols = pilot_symbols
        self.occupied_carriers = occupied_carriers
        self.samp_rate = samp_rate
        self.sync_word1 = sync_word1
        self.sync_word2 = sync_word2
        self.scramble_mode = scramble_mode
        self.crc_mode = crc_mode
        self.clipping_factor = clipping_factor
        self.clipper_mode = clipper_mode
        self.filter_mode = filter_mode
        if isinstance(payload_mod, str):
            if payload_mod not in _BPS:
                raise ValueError("Modulation not supported.")
            bps = _BPS[payload_mod]
        else:   # callers also pass a constellation object (examples/benchmarks.py:167)
            bps = int(payload_mod.bits_per_symbol())
        self.payload_mod = payload_mod
        self.packet_length_tag_key = "packet_len"
        self.length_tag_key = "frame_len"
        if len(sync_word1) != len(sync_word2):
            raise ValueError("Length of sync sequence(s) must be FFT length.")
        self.fft_len = fft_len = (len(sync_word1) + len(sync_word2)) // 2
        self.scramble_seed = 0x7f
        self.rolloff = 0
        self.len_ocup_carr = len(occupied_carriers[0])
        self.cp_len = cp_len = fft_len // 4
        self.active_carriers = len(occupied_carriers[0]) + 4
        self.phy = OfdmPhy(fft_len=fft_len, cp_len=cp_len, occupied_carriers=occupied_carriers,
                           pilot_carriers=pilot_carriers, pilot_symbols=pilot_symbols,
                           sync_word1=sync_word1, sync_word2=sync_word2, bps_header=1, bps_payload=bps,
                           scramble_bits=bool(scramble_mode), scramble_header=True, crc_mode=int(crc_mode),
                           max_carr_offset=3, tx_scale=0.01,
                           tx_clip=float(clipping_factor) if int(clipper_mode) else 0.0, **phy_kwargs)
        # python/ofdm_radio_hier.py:83-84: taps of the 8th-order out-of-band filter (iir_filter_ccd, :93)
        self.forward_OOB = [0.40789374966665903, 3.2351160543115207, 11.253435139165413, 22.423991613997735,
                            27.99555756436666, 22.423991613997735, 11.253435139165425, 3.235116054311531,
                            0.40789374966666014]
        self.feedback_OOB = [1.0, 6.170110168740749, 16.888669609673336, 26.73762881119027, 26.75444043101795,
                             17.322358010203928, 7.091659316015212, 1.682084643429639, 0.17795354282083842]
        self._pkt_num = 0
        self._agc_gain = None
        self._iir_state = None

    def reconfigure(self, occupied_carriers, pilot_carriers, pilot_symbols, sync_word1, sync_word2):
        """New carrier plan and sync words on the running block -- the tuple ofdm_cr_tools.spectrum_enforcer
        returns (python/ofdm_cr_tools.py:348-378), which cognitive_engine_mac forwards to the radios
        (python/cognitive_engine_mac.py:278-285).  fft_len follows the sync words as in the constructor (:76)."""
        fft_len = (len(sync_word1) + len(sync_word2)) // 2
        self.phy.reconfigure(fft_len=fft_len, cp_len=fft_len // 4, occupied_carriers=occupied_carriers,
                             pilot_carriers=pilot_carriers, pilot_symbols=pilot_symbols,
                             sync_word1=sync_word1, sync_word2=sync_word2)
        self.occupied_carriers, self.pilot_carriers, self.pilot_symbols = occupied_carriers, pilot_carriers, pilot_symbols
        self.sync_word1, self.sync_word2 = sync_word1, sync_word2
        self.fft_len, self.cp_len = fft_len, fft_len // 4
        self.len_ocup_carr = len(occupied_carriers[0])
        self.active_carriers = len(occupied_carriers[0]) + 4

    # port 0 (bytes) in -> port 1 (samples) out
    def tx(self, packets):
        """-> (samples, sample offsets).  With filter_mode=1 (the reference default, :39) the burst stream runs
        through iir_filter_ccd(forward_OOB, feedback_OOB) (:93,:232-237); the filter history is carried from
        call to call like the GNU Radio block carries it from burst to burst."""
        out, soff = self.phy.tx(packets, first_pkt_num=self._pkt_num)
        self._pkt_num = (self._pkt_num + len(packets)) & 0xFFF
        if int(self.filter_mode) and out.numel():
            out, self._iir_state = self.phy.iir_ccd(out, self.forward_OOB, self.feedback_OOB, self._iir_state)
        return out, soff

    # port 1 (samples) in -> port 0 (bytes) out
    def rx(self, samples, agc=True, **kw):
        if agc:
            samples, self._agc_gain = self.phy.agc2(samples, self._agc_gain, 1e-1, 1e-2, 1.0, 65536.0)
        return self.phy.rx(samples, **kw)

    # accessors as generated by GRC in the reference (python/ofdm_radio_hier.py:247-371); as there,
    # they only update attributes -- build a new block to change the PHY.
    def get_pilot_carriers(self):
        return self.pilot_carriers

    def set_pilot_carriers(self, pilot_carriers):
        self.pilot_carriers = pilot_carriers

    def get_pilot_symbols(self):
        return self.pilot_symbols

    def set_pilot_symbols(self, pilot_symbols):
        self.pilot_symbols = pilot_symbols

    def get_occupied_carriers(self):
        return self.occupied_carriers

    def set_occupied_carriers(self, occupied_carriers):
        self.occupied_carriers = occupied_carriers

    def get_samp_rate(self):
        return self.samp_rate

    def set_samp_rate(self, samp_rate):
        self.samp_rate = samp_rate

    def get_payload_mod(self):
        return self.payload_mod

    def set_payload_mod(self, payload_mod):
        self.payload_mod = payload_mod

    def get_sync_word1(self):
        return self.sync_word1

    def set_sync_word1(self, sync_word1):
        self.sync_word1 = sync_word1

    def get_sync_word2(self):
        return self.sync_word2

    def set_sync_word2(self, sync_word2):
        self.sync_word2 = sync_word2

    def get_scramble_mode(self):
        return self.scramble_mode

    def get_crc_mode(self):
        return self.crc_mode

    def get_clipper_mode(self):
        return self.clipper_mode

    def set_clipper_mode(self, clipper_mode):
        self.clipper_mode = clipper_mode

    def get_filter_mode(self):
        return self.filter_mode

    def set_filter_mode(self, filter_mode):
        self.filter_mode = filter_mode

    def get_clipping_factor(self):
        return self.clipping_factor

    def set_clipping_factor(self, clipping_factor):
        self.clipping_factor = clipping_factor

    def get_fft_len(self):
        return self.fft_len

    def get_cp_len(self):
        return self.cp_len

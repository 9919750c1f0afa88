"""B200 drop-in for python/sync_radio_hier.py: `sync_radio_hier(samp_rate=10000)`.

The narrow-band control-channel radio of the reference: the same flat TX/RX chain as ofdm_radio_hier with fixed
parameters (:50-68) -- fft_len 64 (from the sync words, :59), cp_len 16, 29 data carriers, 2 pilots, BPSK header
with scramble_header=True, QPSK payload, no in-graph CRC, no scrambler, no clipper, rolloff 0, chanest
max_carr_offset 3 (:90), x0.01 TX scaling (:114) ALWAYS followed by the 12th-order iir_filter_ccd(forward_OOB,
feedback_OOB) out-of-band filter (:66-67,:73,:150,:165), agc2_cc(1e-1, 1e-2, 1.0, 1.0) with max gain 65536 in
front of the receiver (:117-118).  It is a parameterisation of the kernels behind ofdm_radio_hier, so this class
only fixes the arguments.
"""
from .ofdm_radio_hier import ofdm_radio_hier

# python/sync_radio_hier.py:50-51 (equal to python/ofdm_cr_tools.py:49-54; tests/golden pins both)
_SYNC_WORD2 = [0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, (1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), (-1+0j), (1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (-1+0j), (1+0j), 0j, (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), (1+0j), (1+0j), (1+0j), (1+0j), (-1+0j), 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j, 0j]
_SYNC_WORD1 = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 1.42, 0.0, -1.42, 0.0, 1.42, 0.0, 1.42, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
_PILOT_SYMBOLS = ((1, -1,),)
_PILOT_CARRIERS = ((-13, 12,),)
_OCCUPIED_CARRIERS = ([-16, -15, -14, -12, -11, -10, -9, -8, -7, -6, -5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15],)
# python/sync_radio_hier.py:66-67: 12th-order out-of-band filter
_FORWARD_OOB = [0.005277622700213007, 0.03443705907448985, 0.1214101788557494, 0.29179662246081545, 0.52428014905364, 0.7350677973792328, 0.8210395030022875, 0.7350677973792348, 0.5242801490536404, 0.291796622460816, 0.1214101788557501, 0.03443705907448997, 0.005277622700213012]
_FEEDBACK_OOB = [1.0, -1.0455317889337852, 3.9201525346250072, -3.9114761684448958, 6.54266144224035, -5.737287389902878, 5.820328302284336, -4.134700802700442, 2.7949972248757664, -1.4584448495689168, 0.6358650797085171, -0.19847981428665007, 0.04200458351675313]


class sync_radio_hier(ofdm_radio_hier):

    def __init__(self, samp_rate=10000, **phy_kwargs):
        ofdm_radio_hier.__init__(self, pilot_carriers=_PILOT_CARRIERS, pilot_symbols=_PILOT_SYMBOLS,
                                 occupied_carriers=_OCCUPIED_CARRIERS, samp_rate=samp_rate, payload_mod='qpsk',
                                 sync_word1=_SYNC_WORD1, sync_word2=_SYNC_WORD2, scramble_mode=0, crc_mode=0,
                                 clipper_mode=0, filter_mode=1, **phy_kwargs)
        self.forward_OOB = list(_FORWARD_OOB)
        self.feedback_OOB = list(_FEEDBACK_OOB)

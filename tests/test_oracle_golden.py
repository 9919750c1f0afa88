"""CPU tests of the oracle against the reference-supplied golden vectors (tests/golden/, extracted
from /root/reference by tests/golden/make_golden.py), standard check values and the derived vectors
of SURVEY.md Appendix B, plus oracle-internal round trips."""
import numpy as np
import pytest

import common as cm
import oracle as O


def _c(v):
    return np.array([complex(a, b) for a, b in v])


def test_sync_words_match_grc_literals():
    # apps/ofdm_rx_hier.grc / ofdm_tx_hier.grc parameter blocks sync_word1 / sync_word2 (8-digit literals)
    for key in ("grc_ofdm_rx_hier", "grc_ofdm_tx_hier"):
        g = cm.GOLD[key]
        assert np.allclose(O.make_sync_word1(64, cm.OCC64, cm.PIL64), _c(g["sync_word1"]), atol=1e-8)
        assert np.array_equal(O.make_sync_word2(64, cm.OCC64, cm.PIL64), _c(g["sync_word2"]))


def test_radio_hier_defaults_match_generators():
    # python/ofdm_radio_hier.py:34-38: carriers == spectrum_enforcer(128, [], 10); sync words from the
    # 1.42-amplitude generator (python/ofdm_cr_tools.py:262-293)
    d = cm.GOLD["ofdm_radio_hier_defaults"]
    occ, pil, pls, s1, s2 = O.spectrum_enforcer(128, [], 10)
    assert list(occ[0]) == d["occupied_carriers"][0]
    assert [list(x) for x in pil] == [list(x) for x in d["pilot_carriers"]]
    assert [list(x) for x in pls] == [list(x) for x in d["pilot_symbols"]]
    assert np.array_equal(np.array(s1), _c(d["sync_word1"]))
    assert np.array_equal(np.array(s2), _c(d["sync_word2"]))


def test_narrowband_sync_words():
    # python/sync_radio_hier.py:50-56 and the identical literals python/ofdm_cr_tools.py:49-54
    for key in ("sync_radio_hier", "ofdm_cr_tools_sync"):
        s = cm.GOLD[key]
        ref = cm.GOLD["sync_radio_hier"]
        assert np.array_equal(O.make_sync_word1(64, ref["occupied_carriers"], ref["pilot_carriers"], 1.42), _c(s["sync_word1"]))
        assert np.array_equal(O.make_sync_word2(64, ref["occupied_carriers"], ref["pilot_carriers"]), _c(s["sync_word2"]))


def test_sync_word1_structure():
    sw1 = O.make_sync_word1(64, cm.OCC64, cm.PIL64)
    t = np.fft.ifft(np.fft.ifftshift(sw1))
    assert np.allclose(t[:32], -t[32:])           # odd bins only -> second half = -first half
    assert sw1[32] == 0


def test_crc_check_values():
    assert O.crc32(b"123456789") == 0xCBF43926     # zlib / crc32_bb
    assert O.crc8(b"123456789") == 0xFB            # poly 0x07 init 0xFF (packet_header_default)
    import zlib
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 1500, 4096):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert O.crc32(b) == zlib.crc32(b)


def test_scrambler_keystream_and_header_mask():
    # SURVEY.md Appendix B (derived)
    assert O.scramble(bytes(20), 0x7F).tobytes().hex() == "7f10467971d26e563f08a3bc386937ab1f84515e"
    assert O.scramble(b"abc", 0x00).tobytes() == b"abc"
    assert "".join(map(str, O.lfsr_bits(0x8A, 0x6F, 7, 48))) == "111101100110101011111100000100001100010100111101"


def test_header_format_parse():
    orc = cm.make_oracle(cm.cfg_c1(2))
    h = orc.header_format(96, 0)
    assert len(h) == 48
    assert "".join(map(str, h)) == "000001100000" + "000000000000" + "00100001" + "0" * 16   # CRC-8 0x84
    ok, plen, pnum, psyms, fsyms = orc.header_parse(h)
    assert (ok, plen, pnum, psyms, fsyms) == (True, 96, 0, 384, 8)
    h2 = h.copy()
    h2[3] ^= 1
    assert orc.header_parse(h2)[0] is False
    scr = cm.make_oracle(cm.cfg_c1(2, scramble=True))
    hs = scr.header_format(1500, 4095)
    assert not np.array_equal(hs, orc.header_format(1500, 4095))
    assert scr.header_parse(hs)[:3] == (True, 1500, 4095)


def test_constellations_and_decisions():
    q16 = O.constellation(4)
    exp = {0: (1 / 3, 1 / 3), 3: (1, 1), 5: (-1, 1 / 3), 9: (-1 / 3, -1), 15: (1, -1)}
    for i, (re, im) in exp.items():
        assert abs(q16[i] - complex(re, im)) < 1e-6
    assert abs(np.mean(np.abs(q16) ** 2) - 10 / 9) < 1e-6
    for bps in (1, 2, 3, 4, 6):
        pts = O.constellation(bps)
        assert len(pts) == 1 << bps
        for i, p in enumerate(pts):
            assert O.decide(bps, p) == i
            assert O.decide(bps, p + 0.03 * (1 + 1j)) == i
    assert abs(O.constellation(2)[3] - complex(0.707107, 0.707107)) < 1e-7


def test_repack_bits():
    assert list(O.repack([0b10110100], 8, 2, False)) == [0, 1, 3, 2]
    assert list(O.repack([0, 1, 3, 2], 2, 8, True)) == [0b10110100]
    b = np.arange(7, dtype=np.uint8)
    s3 = O.repack(b, 8, 3, False)
    assert len(s3) == 19                                  # ceil(56/3)
    assert list(O.repack(s3, 3, 8, True)) == list(b)      # floor(57/8) = 7


@pytest.mark.parametrize("n", [16, 64, 1024])
def test_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.allclose(O.fft(x, True), np.fft.fft(x))
    assert np.allclose(O.fft(x, False), np.fft.ifft(x) * n)


def test_tx_frame_structure():
    orc = cm.make_oracle(cm.cfg_c1(2))
    pk = [bytes(range(96))]
    s, off = orc.tx(pk)
    assert list(off) == [0, 880]                           # 2 sync + 1 header + 8 payload symbols x 80
    sym = s[:80]
    assert np.allclose(sym[:16], sym[64:80])               # cyclic prefix
    assert np.allclose(sym[16:48], -sym[48:80], atol=1e-5) # sync word 1: two opposite halves
    f2 = np.fft.fftshift(np.fft.fft(s[80 + 16:160]))
    assert np.allclose(f2, 64 * orc.sw2, atol=1e-3)        # unnormalised IFFT of sync word 2
    crc = cm.make_oracle(cm.cfg_c1(2, crc=1))
    assert crc.frame_samples(96) == 960                    # 100 bytes -> 9 payload symbols


@pytest.mark.parametrize("bps,scr,crc", [(1, False, 0), (2, True, 1), (3, True, 1), (4, True, 1), (6, True, 1)])
def test_oracle_roundtrip(bps, scr, crc):
    cfg = cm.cfg_c1(bps, scr, crc)
    orc = cm.make_oracle(cfg)
    rng = np.random.default_rng(bps)
    pk = cm.rand_packets(rng, 6, 77)
    s, off = orc.tx(pk)
    x = cm.channel(cm.split_frames(s, off), rng, snr_db=45.0, cfo=0.21, taps=cm.MULTIPATH[:2])
    res = orc.rx(x, byte_stride=128)
    assert orc.payloads(res) == pk
    assert np.allclose(res["cfo"], np.pi * 0.21, atol=0.02)
    assert np.array_equal(res["frames"]["pkt_num"], np.arange(6))


def test_plateau_semantics_and_trigger_position():
    cfg = cm.cfg_c1(2)
    orc = cm.make_oracle(cfg)
    s, off = orc.tx([bytes(96)])
    x = np.concatenate([np.zeros(300, np.complex64), s, np.zeros(300, np.complex64)])
    trig, cfo, det = orc.sync(x, want_detect=True)
    assert list(trig) == [373]                             # start + N + cp/2 + 1 (SURVEY.md A.2 probe)
    run = np.flatnonzero(det)
    assert run[0] == 363 and run[-1] == 382
    # fewer than 2*cp items after the flank: no trigger
    trig2, _ = orc.sync(x[:363 + 20])
    assert len(trig2) == 0


def test_demux_rules():
    cfg = cm.cfg_c1(2, True, 1)
    orc = cm.make_oracle(cfg)
    rng = np.random.default_rng(3)
    pk = cm.rand_packets(rng, 4, 50)
    s, off = orc.tx(pk)
    fr = cm.split_frames(s, off)
    fr[1] = fr[1].copy()
    fr[1][2 * 80 + 16:3 * 80] *= -1                        # corrupt header of frame 1
    x = cm.channel(fr, rng, snr_db=40.0, tail=5)
    res = orc.rx(x[:len(x) - 100], byte_stride=64)         # last frame truncated
    assert orc.payloads(res) == [pk[0], pk[2]]
    assert len(res["triggers"]) >= 4
    # f32 FIR port finds the same triggers here
    t32, _ = orc.sync(x, f32=True)
    t64, _ = orc.sync(x)
    assert np.array_equal(t32, t64)

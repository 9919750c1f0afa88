"""AGC front end (rank 1), TX clipper (rank 2), MAC framing + PDU adaptors (rank 3).
CPU part: oracle restatements against known answers and properties, host-side framing.  GPU part
(-m gpu): CUDA kernels through the C ABI, bit-exact against the oracle."""
import struct
import zlib

import numpy as np
import pytest

import common as cm


# python/ofdm_radio_hier.py:83-84
FORWARD_OOB = [0.40789374966665903, 3.2351160543115207, 11.253435139165413, 22.423991613997735, 27.99555756436666,
               22.423991613997735, 11.253435139165425, 3.235116054311531, 0.40789374966666014]
FEEDBACK_OOB = [1.0, 6.170110168740749, 16.888669609673336, 26.73762881119027, 26.75444043101795,
                17.322358010203928, 7.091659316015212, 1.682084643429639, 0.17795354282083842]


def _cnoise(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_iir_ccd_against_direct_form():
    """orc_iir_ccd (iir_filter_ccd, oldstyle=False) against scipy's direct-form filter in complex128 and against
    a pure-Python loop of the GNU Radio recurrence; history carried across calls; FIR-only taps."""
    import oracle as O
    from scipy.signal import lfilter
    rng = np.random.default_rng(11)
    x = _cnoise(rng, 3000)
    y, st = O.iir_ccd(x, FORWARD_OOB, FEEDBACK_OOB)
    ref = lfilter(FORWARD_OOB, FEEDBACK_OOB, x.astype(np.complex128))
    assert np.abs(y - ref).max() <= 2e-7 * np.abs(ref).max()
    # GNU Radio's own order of operations, 40 samples, by hand
    xs, ys, want = [0j] * 8, [0j] * 8, []
    for v in x[:40]:
        v = complex(v)
        acc = complex(FORWARD_OOB[0] * v.real, FORWARD_OOB[0] * v.imag)
        for i in range(1, 9):
            acc = complex(acc.real + FORWARD_OOB[i] * xs[i - 1].real, acc.imag + FORWARD_OOB[i] * xs[i - 1].imag)
        for i in range(1, 9):
            acc = complex(acc.real + (-FEEDBACK_OOB[i]) * ys[i - 1].real, acc.imag + (-FEEDBACK_OOB[i]) * ys[i - 1].imag)
        xs = [v] + xs[:-1]
        ys = [acc] + ys[:-1]
        want.append(np.complex64(acc))
    assert np.array_equal(y[:40], np.array(want, np.complex64))
    y1, s1 = O.iir_ccd(x[:777], FORWARD_OOB, FEEDBACK_OOB)
    y2, s2 = O.iir_ccd(x[777:], FORWARD_OOB, FEEDBACK_OOB, s1)
    assert np.array_equal(np.concatenate([y1, y2]), y) and np.array_equal(s2, st)
    taps = [0.25, 0.5, 0.25]
    yf, _ = O.iir_ccd(x, taps, [1.0])
    assert np.allclose(yf, lfilter(taps, [1.0], x.astype(np.complex128)), rtol=0, atol=1e-6)


def _prefixer_numpy(td_syms, cp, roll):
    """Independent numpy restatement of ofdm_cyclic_prefixer in packet mode: td_syms [n_sym, N] -> burst."""
    n_sym, N = td_syms.shape
    nfl = roll - 1 if roll > 1 else 0
    i = np.arange(1, nfl + 1)
    up = (0.5 * (1 + np.cos(np.pi * i / max(roll, 1) - np.pi))).astype(np.float32)
    down = (0.5 * (1 + np.cos(np.pi * (max(roll, 1) - i) / max(roll, 1) - np.pi))).astype(np.float32)
    out = np.zeros(n_sym * (N + cp) + nfl, np.complex128)
    for s_ in range(n_sym):
        sym = np.concatenate([td_syms[s_, N - cp:], td_syms[s_]])
        sym[:nfl] *= up
        out[s_ * (N + cp): (s_ + 1) * (N + cp)] += sym
        out[(s_ + 1) * (N + cp): (s_ + 1) * (N + cp) + nfl] += td_syms[s_, :nfl] * down
    return out


def test_oracle_tx_rolloff():
    """orc_tx with rolloff > 0 against the rolloff-0 burst re-windowed by an independent numpy restatement of
    ofdm_cyclic_prefixer (flanks sum to one; each burst rolloff-1 samples longer; rolloff 1 == rolloff 0)."""
    import oracle as O
    rng = np.random.default_rng(3)
    pk = cm.rand_packets(rng, 3, 60) + [b""]
    cfg = cm.cfg_c1()
    base = O.Oracle(**cfg)
    s0, off0 = base.tx(pk)
    N, cp = cfg["fft_len"], cfg["cp_len"]
    for roll in (1, 2, 4, 7, 16):
        orc = O.Oracle(rolloff=roll, **cfg)
        s, off = orc.tx(pk)
        nfl = roll - 1 if roll > 1 else 0
        assert np.array_equal(off, off0 + nfl * np.arange(len(pk) + 1))
        assert orc.frame_samples(60) == base.frame_samples(60) + nfl
        for k in range(len(pk)):
            burst0 = s0[off0[k]:off0[k + 1]].astype(np.complex128).reshape(-1, N + cp)
            want = _prefixer_numpy(burst0[:, cp:], cp, roll)
            got = s[off[k]:off[k + 1]]
            assert len(got) == len(want)
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    with pytest.raises(Exception):
        O.Oracle(rolloff=cp + 1, **cfg).tx(pk)


def test_oracle_papr():
    """python/papr_sink.py:46-50 on a constant-envelope block (PAPR 1) and on a single peak."""
    import oracle as O
    x = np.exp(1j * np.arange(512)).astype(np.complex64)
    assert abs(O.papr(x) - 1.0) < 1e-5
    x = np.ones(100, np.complex64)
    x[7] = 3 + 4j
    assert abs(O.papr(x) - 25.0 / ((99 + 25) / 100.0)) < 1e-4


def test_mac_crc32_check_values():
    """MSB-first CRC-32 of digital.crc (SURVEY.md A.13 check value) in the oracle and in the host module."""
    import oracle as O
    from ofdm_tools import crc
    assert O.crc32_mac(b"123456789") == 0xFC891918
    assert crc.crc32(b"123456789") == 0xFC891918
    assert O.crc32(b"123456789") == 0xCBF43926          # the in-graph one is a different CRC
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 3, 4, 5, 63, 64, 1000):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert crc.crc32(b) == O.crc32_mac(b)
        framed = crc.gen_and_append_crc32(b)
        assert framed[:-4] == b and struct.unpack(">I", framed[-4:])[0] == O.crc32_mac(b)
        assert crc.check_crc32(framed) == (True, b)
        if n:
            bad = bytes([framed[0] ^ 1]) + framed[1:]
            assert crc.check_crc32(bad)[0] is False
    assert crc.check_crc32(b"abc") == (False, b"")


def test_make_unmake_packet():
    """python/ofdm_cr_tools.py:1741-1773 and the way examples/benchmarks.py:343-349 uses them."""
    from ofdm_tools import crc
    from ofdm_tools.ofdm_cr_tools import make_packet, unmake_packet
    pkt = make_packet("0007" + "ab" * 15, 96, 'A')
    assert len(pkt) == 96 and pkt[:4] == b"0034" and pkt[4:5] == b"A" and pkt[5:39] == b"0007" + b"ab" * 15
    assert set(pkt[39:]) == {0x55}
    assert unmake_packet(pkt, True) == (b"0007" + b"ab" * 15, b"A", True)
    framed = crc.gen_and_append_crc32(make_packet(b"xyz", 96 - 4, b'B'))
    assert len(framed) == 96
    assert unmake_packet(framed, False) == (b"xyz", b"B", True)
    corrupt = framed[:10] + bytes([framed[10] ^ 0x40]) + framed[11:]
    assert unmake_packet(corrupt, False) == ('BAD', 'BAD', False)


def test_pdu_adaptors_host():
    """payload_source_pdu appends the in-graph CRC-32 (crc32_async_bb(False)), payload_sink_pdu checks and
    strips it and calls callback(addr, tpe, nr, payload) (python/ofdm_cr_tools.py:2124-2151)."""
    from ofdm_tools import payload_source_pdu, payload_sink_pdu
    src = payload_source_pdu()
    src.post_message(1, 2, 3, "hello")
    src.post_message(9, 8, 7, b"\x00\xff" * 40)
    pk = src.pop_packets()
    assert pk[0][:3] == b"\x01\x02\x03" and pk[0][3:-4] == b"hello"
    assert struct.unpack("<I", pk[0][-4:])[0] == zlib.crc32(pk[0][:-4]) & 0xFFFFFFFF
    got = []
    snk = payload_sink_pdu(lambda a, t, n, p: got.append((a, t, n, p)))
    bad = pk[1][:5] + bytes([pk[1][5] ^ 1]) + pk[1][6:]
    out = snk.deliver([pk[0], bad, pk[1], b"ab"])
    assert got == [(1, 2, 3, b"hello"), (9, 8, 7, b"\x00\xff" * 40)]
    assert out == [pk[0][:-4], pk[1][:-4]] and (snk.n_rcvd, snk.n_right) == (4, 2)


def test_oracle_agc2_properties():
    """agc2_cc restatement: hand-computed first steps, attack vs decay branch, clamps, and convergence of
    the output level to the reference on a constant-envelope input."""
    import oracle as O
    f = np.float32
    x = np.array([3 + 4j, 0.1 + 0j, 0, 1e-9], np.complex64)
    y, g = O.agc2(x, gain=1.0)
    # sample 0: out = 3+4j, |out| = 5, tmp = 4 > gain 1 -> attack 0.1: gain = 1 - 0.4 = 0.6
    assert y[0] == 3 + 4j
    g0 = f(1.0) - f(4.0) * f(0.1)
    assert y[1] == np.complex64(complex(f(0.1) * g0, 0.0))
    # sample 1: |out| = 0.06, tmp = -0.94 -> decay: gain += 0.0094
    tmp1 = f(-1.0) + np.sqrt(f(f(0.1) * g0) * f(f(0.1) * g0), dtype=f)
    g1 = f(g0 - f(tmp1 * f(0.01)))
    assert y[2] == 0 and y[3] == np.complex64(complex(f(1e-9) * f(g1 - f(f(-1.0) * f(0.01))), 0.0))
    # negative gain is replaced by 10e-5, the ceiling by max_gain
    _, g = O.agc2(np.array([1000 + 0j], np.complex64), gain=1.0)
    assert g == pytest.approx(10e-5)
    _, g = O.agc2(np.zeros(10, np.complex64), gain=65535.995, max_gain=65536.0)
    assert g == 65536.0
    # convergence: constant envelope 7.3 -> output magnitude -> 1.0
    t = np.arange(5000)
    y, g = O.agc2((7.3 * np.exp(0.1j * t)).astype(np.complex64))
    assert abs(np.abs(y[-1]) - 1.0) < 1e-3 and abs(g - 1 / 7.3) < 1e-3
    # chunked == whole (state carried through the gain)
    xr = (np.random.default_rng(1).standard_normal(2000) * 3).astype(np.complex64)
    ya, ga = O.agc2(xr)
    yb1, gb = O.agc2(xr[:777])
    yb2, gb = O.agc2(xr[777:], gain=gb)
    assert np.array_equal(ya, np.concatenate([yb1, yb2])) and ga == gb


def test_oracle_tx_clip():
    import oracle as O
    cfg = cm.cfg_radio128(2, 1, 1)
    cfg["tx_scale"] = 0.01
    rng = np.random.default_rng(3)
    pk = cm.rand_packets(rng, 3, 100)
    s0, off0 = O.Oracle(**cfg).tx(pk)
    s1, off1 = O.Oracle(tx_clip=0.05, **cfg).tx(pk)
    assert np.array_equal(off0, off1)
    ref = np.clip(s0.real, -0.05, 0.05) + 1j * np.clip(s0.imag, -0.05, 0.05)
    assert np.array_equal(s1, ref.astype(np.complex64))
    assert np.abs(s0.real).max() > 0.05                      # the rail is active on this signal


# ------------------------------------------------------------------------------------------- GPU
def _dev():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _agc_input(rng, n_streams, n):
    """Per-stream level steps over 9 decades, silences and bursts."""
    x = (rng.standard_normal((n_streams, n)) + 1j * rng.standard_normal((n_streams, n))).astype(np.complex64)
    for s in range(n_streams):
        cuts = np.sort(rng.integers(0, n, 4))
        lev = 10.0 ** rng.uniform(-6, 3, 5)
        seg = np.split(np.arange(n), cuts)
        for k, idx in enumerate(seg):
            x[s, idx] *= np.float32(lev[k]) if (s + k) % 7 else np.float32(0.0)
    return x


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams,n", [(1, 5000), (3, 1), (32, 31), (33, 1000), (70, 3001), (257, 640)])
def test_agc2_parity(n_streams, n):
    """ofdmx_agc2 against the oracle: output samples and final loop gains bit-exact; two chunked calls equal
    one call (gain state carried); in place."""
    import torch
    import oracle as O
    rng = np.random.default_rng(100 + n_streams)
    x = _agc_input(rng, n_streams, n)
    ref, gref = O.agc2(x)
    phy = cm.make_phy(cm.cfg_c1())
    xd = torch.from_numpy(x).to(_dev())
    y, g = phy.agc2(xd)
    assert np.array_equal(y.cpu().numpy().view(np.float32), ref.view(np.float32))
    assert np.array_equal(g.cpu().numpy(), gref)
    if n >= 2:
        k = n // 3 + 1
        buf = xd.clone()
        _, g2 = phy.agc2(buf[:, :k].contiguous(), out=None)
        ya, g2 = phy.agc2(xd[:, :k].contiguous())
        yb, g2 = phy.agc2(xd[:, k:].contiguous(), gain=g2)
        assert np.array_equal(torch.cat([ya, yb], 1).cpu().numpy().view(np.float32), ref.view(np.float32))
        assert np.array_equal(g2.cpu().numpy(), gref)
    z = xd.clone()
    phy.agc2(z, out=z)
    assert np.array_equal(z.cpu().numpy().view(np.float32), ref.view(np.float32))
    # custom rates / reference / no ceiling
    ref2, gref2 = O.agc2(x, gain=np.full(n_streams, 0.25, np.float32), attack=0.5, decay=1e-3, reference=0.3, max_gain=0.0)
    y2, g2 = phy.agc2(xd, gain=torch.full((n_streams,), 0.25, device=_dev()), attack=0.5, decay=1e-3, reference=0.3, max_gain=0.0)
    assert np.array_equal(y2.cpu().numpy().view(np.float32), ref2.view(np.float32)) and np.array_equal(g2.cpu().numpy(), gref2)


@pytest.mark.gpu
def test_tx_clip_parity():
    """clipper fused into the TX kernel (ofdm_radio_hier clipper_mode=1) against the oracle."""
    import oracle as O
    cfg = cm.cfg_radio128(2, 1, 1)
    cfg["tx_scale"] = 0.01
    rng = np.random.default_rng(8)
    pk = cm.rand_packets(rng, 5, 350)
    ref, off = O.Oracle(tx_clip=0.05, **cfg).tx(pk)
    s, soff = cm.make_phy(cfg, tx_clip=0.05).tx(pk)
    s = s.cpu().numpy()
    assert np.array_equal(soff.cpu().numpy(), off)
    assert np.abs(s.real).max() <= np.float32(0.05) and np.abs(s.imag).max() <= np.float32(0.05)
    assert (np.abs(s.real) == np.float32(0.05)).sum() > 10
    assert np.linalg.norm(s - ref) / np.linalg.norm(ref) < 1e-5
    # stand-alone block
    import torch
    from ofdm_tools import clipper
    raw, _ = cm.make_phy(cfg).tx(pk)
    c = clipper(0.05).work(raw).cpu().numpy()
    assert np.array_equal(c, s)


@pytest.mark.gpu
def test_hier_facades_with_agc_and_clipper():
    """ofdm_radio_hier / ofdm_tx_rx_hier with the RX AGC in front (reference wiring) on a nominal signal, one
    1000x too weak and one 300x too strong: identical frames and payloads to oracle AGC -> oracle RX; the
    AGC state is carried across rx() calls.  (On the 300x signal the fast attack of agc2_cc(1e-1, 1e-2)
    modulates the envelope inside the OFDM symbols and every 16-QAM packet fails its CRC -- in the oracle
    too: that case pins the parity of the failure path.)"""
    import torch
    import oracle as O
    from ofdm_tools import ofdm_radio_hier, ofdm_tx_rx_hier
    rng = np.random.default_rng(77)
    radio = ofdm_radio_hier(payload_mod='qam16', scramble_mode=1, crc_mode=1, clipper_mode=1, clipping_factor=0.3,
                            filter_mode=0)
    pk = cm.rand_packets(rng, 6, 200)
    s, off = radio.tx(pk)
    s = s.cpu().numpy()
    assert np.abs(s.real).max() <= np.float32(0.3)
    cfg = cm.cfg_radio128(4, 1, 1)
    cfg.update(tx_scale=0.01, max_carr_offset=3, scramble_header=True)
    orc = O.Oracle(**cfg)
    for scale in (1.0, 1e-3, 300.0):
        x = cm.channel(cm.split_frames(s, off.cpu().numpy()), rng, gaps=(300, 900), tail=2000, snr_db=45.0, cfo=0.1,
                       fft_len=128, scale=scale)
        radio._agc_gain = None
        res = radio.rx(torch.from_numpy(x).to(_dev()))
        xa, g = O.agc2(x)
        ref = orc.rx(xa, byte_stride=radio.phy.byte_stride, want_z=False)
        assert np.array_equal(res.frames["trigger"], ref["frames"]["trigger"])
        assert np.array_equal(res.frames["flags"] & 7, ref["frames"]["flags"] & 7)
        assert res.payloads() == orc.payloads(ref)
        assert (res.payloads() == pk) == (scale <= 1.0)
        assert float(radio._agc_gain[0]) == g
        # second call continues the loop gain
        res2 = radio.rx(torch.from_numpy(x).to(_dev()))
        xb, g2 = O.agc2(x, gain=g)
        assert float(radio._agc_gain[0]) == g2
        assert res2.payloads() == orc.payloads(orc.rx(xb, byte_stride=radio.phy.byte_stride, want_z=False))
    trx = ofdm_tx_rx_hier(fft_len=64, payload_bps=2)
    pk = cm.rand_packets(rng, 4, 96)
    s, off = trx.tx(pk)
    x = cm.channel(cm.split_frames(s.cpu().numpy(), off.cpu().numpy()), rng, gaps=(200, 500), tail=1500, snr_db=35.0,
                   fft_len=64, scale=0.02)
    c1 = cm.cfg_c1()
    c1.update(bps_payload=2)
    o1 = O.Oracle(**c1)
    ref = o1.rx(O.agc2(x)[0], want_z=False)
    assert trx.rx(torch.from_numpy(x).to(_dev())).payloads() == o1.payloads(ref) == pk


@pytest.mark.gpu
def test_pdu_adaptors_gpu_crc():
    """payload_source_pdu / payload_sink_pdu with the CRC-32 batch computed by the CUDA crc32 kernel."""
    from ofdm_tools import payload_source_pdu, payload_sink_pdu
    phy = cm.make_phy(cm.cfg_c1())
    rng = np.random.default_rng(2)
    src, host = payload_source_pdu(phy=phy), payload_source_pdu()
    for n in (0, 1, 5, 100, 1500):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        src.post_message(1, 2, n & 0xFF, b)
        host.post_message(1, 2, n & 0xFF, b)
    pk = src.pop_packets()
    assert pk == host.pop_packets()
    got = []
    snk = payload_sink_pdu(lambda a, t, n, p: got.append(p), phy=phy)
    snk.deliver(pk[:2] + [pk[2][:-1] + bytes([pk[2][-1] ^ 0x80])] + pk[3:])
    assert [len(p) for p in got] == [0, 1, 100, 1500]


@pytest.mark.gpu
def test_iir_ccd_parity():
    """ofdmx_iir_ccd against the oracle: one span = the sequential recurrence, bit for bit, history included;
    many parallel spans agree to the round-off noise floor of the direct-form recurrence itself (the reference
    taps put eight poles at radius <= 0.985 next to z = -1, so the double accumulator carries ~1e-10 of relative
    round-off noise whatever the order of evaluation): every float within one ulp of the peak, > 98 % identical;
    strided multi-stream input; FIR-only taps; error paths."""
    import torch
    import oracle as O
    phy = cm.make_phy(cm.cfg_c1())
    rng = np.random.default_rng(21)
    x = _cnoise(rng, 5000)
    xd = torch.from_numpy(x).to(_dev())
    y, st = phy.iir_ccd(xd, FORWARD_OOB, FEEDBACK_OOB, span=1 << 20)
    ref, rst = O.iir_ccd(x, FORWARD_OOB, FEEDBACK_OOB)
    assert np.array_equal(y.cpu().numpy(), ref)
    assert np.array_equal(st.cpu().numpy().ravel()[:32], rst)
    y2, st = phy.iir_ccd(xd, FORWARD_OOB, FEEDBACK_OOB, state=st, span=1 << 20)     # continues the stream
    ref2, rst2 = O.iir_ccd(x, FORWARD_OOB, FEEDBACK_OOB, rst)
    assert np.array_equal(y2.cpu().numpy(), ref2) and np.array_equal(st.cpu().numpy().ravel()[:32], rst2)
    # parallel spans (explicit and library-chosen), ragged tail, carried state
    x = _cnoise(rng, 300001)
    xd = torch.from_numpy(x).to(_dev())
    ref, rst = O.iir_ccd(x, FORWARD_OOB, FEEDBACK_OOB)
    refb, rstb = O.iir_ccd(x, FORWARD_OOB, FEEDBACK_OOB, rst)
    for span in (1024, 4096, 0):
        y, st = phy.iir_ccd(xd, FORWARD_OOB, FEEDBACK_OOB, span=span)
        y = y.cpu().numpy()
        assert np.abs(y - ref).max() <= 2.4e-7 * np.abs(ref).max()
        assert np.mean(y == ref) > 0.98
        assert np.array_equal(y[:1024], ref[:1024])                 # the first span is the sequential filter
        assert np.abs(st.cpu().numpy().ravel()[:32] - rst).max() <= 1e-8 * np.abs(rst).max()
        yb, st = phy.iir_ccd(xd, FORWARD_OOB, FEEDBACK_OOB, state=st, span=span)
        assert np.abs(yb.cpu().numpy() - refb).max() <= 2.4e-7 * np.abs(refb).max()
    # streams = rows of a wider matrix
    m = _cnoise(rng, 37, 6000)
    md = torch.from_numpy(m).to(_dev())
    y, st = phy.iir_ccd(md[:, :5555], FORWARD_OOB, FEEDBACK_OOB, span=1 << 20)
    for r in range(37):
        rr, rs = O.iir_ccd(m[r, :5555], FORWARD_OOB, FEEDBACK_OOB)
        assert np.array_equal(y[r].cpu().numpy(), rr) and np.array_equal(st[r].cpu().numpy()[:32], rs)
    y, _ = phy.iir_ccd(md[:, :5555], FORWARD_OOB, FEEDBACK_OOB, span=2048)
    assert np.abs(y.cpu().numpy() - np.stack([O.iir_ccd(m[r, :5555], FORWARD_OOB, FEEDBACK_OOB)[0] for r in range(37)])).max() < 1e-5
    # FIR only
    yf, _ = phy.iir_ccd(xd, [0.25, 0.5, 0.25], [1.0], span=0)
    assert np.array_equal(yf.cpu().numpy(), O.iir_ccd(x, [0.25, 0.5, 0.25], [1.0])[0])
    with pytest.raises(Exception):
        phy.iir_ccd(xd, [1.0] * 18, FEEDBACK_OOB)
    with pytest.raises(Exception):
        phy.iir_ccd(xd, FORWARD_OOB, FEEDBACK_OOB, span=1024, out=xd)
    with pytest.raises(Exception):
        phy.iir_ccd(xd, [1.0], [1.0, -1.5])                        # pole outside the unit circle


@pytest.mark.gpu
def test_papr_sink_gpu():
    """papr_sink.level() on the GPU against python/papr_sink.py:46-50 restated in the oracle."""
    import torch
    import oracle as O
    from ofdm_tools import papr_sink
    phy = cm.make_phy(cm.cfg_c1())
    rng = np.random.default_rng(4)
    for n, bl in ((1, 512), (511, 512), (100000, 4096), (300000, 300000)):
        x = _cnoise(rng, n)
        snk = papr_sink(bl, phy=phy)
        assert snk.work(torch.from_numpy(x).to(_dev())) == min(n, bl)
        assert abs(snk.level() - O.papr(x[:bl])) <= 1e-5 * O.papr(x[:bl])


@pytest.mark.gpu
def test_radio_hier_filter_mode():
    """ofdm_radio_hier(filter_mode=1), the reference default: the TX burst stream equals oracle TX -> oracle
    iir_filter_ccd (history carried across tx() calls), and the filtered bursts still decode."""
    import torch
    import oracle as O
    from ofdm_tools import ofdm_radio_hier
    rng = np.random.default_rng(9)
    radio = ofdm_radio_hier(payload_mod='qpsk', scramble_mode=1, crc_mode=1)
    assert radio.filter_mode == 1
    cfg = cm.cfg_radio128(2, 1, 1)
    cfg.update(tx_scale=0.01, max_carr_offset=3, scramble_header=True)
    orc = O.Oracle(**cfg)
    st, n0 = None, 0
    for call in range(2):
        pk = cm.rand_packets(rng, 5, 180)
        s, off = radio.tx(pk)
        so, oo = orc.tx(pk, first_pkt_num=n0)
        n0 += len(pk)
        assert np.array_equal(off.cpu().numpy(), oo)
        ref, st = O.iir_ccd(so, radio.forward_OOB, radio.feedback_OOB, st)
        s = s.cpu().numpy()
        assert np.abs(s - ref).max() <= 1e-5 * np.abs(ref).max()
        x = cm.channel(cm.split_frames(s, oo), rng, gaps=(400, 900), tail=2000, snr_db=40.0, fft_len=128, scale=100.0)
        assert radio.rx(torch.from_numpy(x).to(_dev()), agc=False).payloads() == pk


@pytest.mark.gpu
def test_tx_rolloff_parity_and_loopback():
    """ofdm_tx(rolloff=r): CUDA TX against the oracle (offsets exact, samples within 1e-5) for the
    sync_transmit_path setting rolloff = cp_len/4 (python/ofdm_cr_tools.py:1093) and others, with the clipper
    behind it; the windowed bursts decode on the RX chain; rolloff > cp_len is refused as in GNU Radio."""
    import torch
    import oracle as O
    from ofdm_tools import ofdm_tx
    rng = np.random.default_rng(13)
    for cfg, rolls, plen in ((cm.cfg_c1(), (4, 2, 16, 1), 96), (cm.cfg_radio128(4, 1, 1), (8,), 200)):
        for roll in rolls:
            kw = dict(cfg, rolloff=roll, tx_scale=0.01, tx_clip=0.3)
            phy = cm.make_phy(kw)
            orc = O.Oracle(**kw)
            pk = cm.rand_packets(rng, 7, plen) + [b"", b"x"]
            s, off = phy.tx(pk)
            so, oo = orc.tx(pk)
            assert np.array_equal(off.cpu().numpy(), oo)
            s = s.cpu().numpy()
            assert s.shape == so.shape and np.abs(s - so).max() <= 1e-5 * np.abs(so).max()
            assert phy.frame_samples(plen) == orc.frame_samples(plen)
            x = cm.channel(cm.split_frames(s, oo), rng, gaps=(300, 700), tail=1500, snr_db=40.0,
                           fft_len=cfg["fft_len"], scale=100.0)
            got = phy.rx(torch.from_numpy(x).to(_dev())).payloads()
            assert got == orc.payloads(orc.rx(x, byte_stride=phy.byte_stride, want_z=False))
            assert [g for g in got if len(g) > 1] == [p_ for p_ in pk if len(p_) > 1]
    tx = ofdm_tx(fft_len=64, cp_len=16, bps_header=1, bps_payload=2, rolloff=4)
    assert tx.rolloff == 4 and tx.phy.frame_samples(96) == cm.make_phy(cm.cfg_c1()).frame_samples(96) + 3
    with pytest.raises(Exception):
        ofdm_tx(fft_len=64, cp_len=16, rolloff=17).work([b"abc"])


@pytest.mark.gpu
def test_runtime_reconfiguration():
    """ofdm_radio_hier.reconfigure / OfdmPhy.reconfigure (ofdmx_reconfigure): the carrier plans spectrum_enforcer
    derives for three spectrum masks (and a change of fft_len) applied to ONE live context; after every change
    TX and RX equal a fresh oracle of that plan, the packet counter and the launch counter run on, the
    workspace is kept, and a refused plan leaves the old one working."""
    import torch
    import oracle as O
    from ofdm_tools import ofdm_radio_hier, ofdm_cr_tools as T
    rng = np.random.default_rng(31)
    radio = ofdm_radio_hier(payload_mod='qpsk', scramble_mode=1, crc_mode=1, filter_mode=0)
    n0, launches = 0, 0
    for fft_len, mask in ((128, []), (128, list(range(-30, -10))), (256, list(range(40, 90))), (128, [5, 6, 7, -50])):
        occ, pil, pls, sw1, sw2 = T.spectrum_enforcer(fft_len, mask, 10)
        radio.reconfigure(occ, pil, pls, sw1, sw2)
        assert radio.phy.fft_len == fft_len and radio.phy.launch_count() >= launches
        orc = O.Oracle(fft_len=fft_len, cp_len=fft_len // 4, occupied_carriers=occ, pilot_carriers=pil,
                       pilot_symbols=pls, sync_word1=sw1, sync_word2=sw2, bps_header=1, bps_payload=2,
                       scramble_bits=True, scramble_header=True, crc_mode=1, tx_scale=0.01, max_carr_offset=3)
        pk = cm.rand_packets(rng, 5, 150)
        s, off = radio.tx(pk)
        so, oo = orc.tx(pk, first_pkt_num=n0)
        n0 += len(pk)
        s = s.cpu().numpy()
        assert np.array_equal(off.cpu().numpy(), oo) and np.abs(s - so).max() <= 1e-5 * np.abs(so).max()
        x = cm.channel(cm.split_frames(s, oo), rng, gaps=(400, 900), tail=2500, snr_db=40.0, cfo=0.2,
                       fft_len=fft_len, scale=100.0)
        res = radio.rx(torch.from_numpy(x).to(_dev()), agc=False)
        ref = orc.rx(x, byte_stride=radio.phy.byte_stride, want_z=False)
        assert np.array_equal(res.frames["trigger"], ref["frames"]["trigger"])
        assert res.payloads() == orc.payloads(ref) == pk
        launches = radio.phy.launch_count()
        assert launches > 0
    # a plan the library refuses: error raised, the live context keeps working with the old plan
    with pytest.raises(Exception):
        radio.phy.reconfigure(rolloff=1000)
    s2, _ = radio.tx(pk)
    so2, _ = orc.tx(pk, first_pkt_num=n0)
    assert np.abs(s2.cpu().numpy() - so2).max() <= 1e-5 * np.abs(so2).max()


@pytest.mark.gpu
def test_sync_radio_hier_facade():
    """sync_radio_hier (python/sync_radio_hier.py:50-68): fft_len 64 narrow-band plan, QPSK, the 12th-order
    iir_filter_ccd always behind the TX chain (13 + 13 taps: the 13-tap instantiation of the kernel, bit-exact in
    its first span), AGC in front of the receiver.  TX equals oracle TX -> oracle IIR, RX equals oracle AGC -> RX,
    and the filtered bursts decode."""
    import torch
    import oracle as O
    from ofdm_tools import sync_radio_hier
    rng = np.random.default_rng(17)
    radio = sync_radio_hier()
    g = cm.GOLD["sync_radio_hier"]
    assert radio.fft_len == 64 and radio.cp_len == 16 and len(radio.forward_OOB) == len(radio.feedback_OOB) == 13
    assert np.allclose(np.asarray(radio.sync_word1, np.complex64), np.array([complex(a, b) for a, b in g["sync_word1"]]))
    orc = O.Oracle(fft_len=64, cp_len=16, occupied_carriers=radio.occupied_carriers, pilot_carriers=radio.pilot_carriers,
                   pilot_symbols=radio.pilot_symbols, sync_word1=radio.sync_word1, sync_word2=radio.sync_word2,
                   bps_header=1, bps_payload=2, scramble_bits=False, scramble_header=True, crc_mode=0, tx_scale=0.01,
                   max_carr_offset=3)
    x1 = (rng.standard_normal(4000) + 1j * rng.standard_normal(4000)).astype(np.complex64)
    y, st = radio.phy.iir_ccd(torch.from_numpy(x1).to(_dev()), radio.forward_OOB, radio.feedback_OOB, span=1 << 20)
    ref, rst = O.iir_ccd(x1, radio.forward_OOB, radio.feedback_OOB)
    assert np.array_equal(y.cpu().numpy(), ref) and np.array_equal(st.cpu().numpy().ravel()[:48], rst)
    st, n0 = None, 0
    for call in range(2):
        pk = cm.rand_packets(rng, 6, 40)
        s, off = radio.tx(pk)
        so, oo = orc.tx(pk, first_pkt_num=n0)
        n0 += len(pk)
        ref, st = O.iir_ccd(so, radio.forward_OOB, radio.feedback_OOB, st)
        s = s.cpu().numpy()
        assert np.array_equal(off.cpu().numpy(), oo) and np.abs(s - ref).max() <= 1e-5 * np.abs(ref).max()
        x = cm.channel(cm.split_frames(s, oo), rng, gaps=(300, 700), tail=1500, snr_db=40.0, fft_len=64, scale=1.0)
        radio._agc_gain = None
        res = radio.rx(torch.from_numpy(x).to(_dev()))
        want = orc.rx(O.agc2(x)[0], byte_stride=radio.phy.byte_stride, want_z=False)
        assert np.array_equal(res.frames["trigger"], want["frames"]["trigger"])
        assert res.payloads() == orc.payloads(want)
        xs = cm.channel(cm.split_frames(s, oo), rng, gaps=(300, 700), tail=1500, snr_db=40.0, fft_len=64, scale=100.0)
        assert radio.rx(torch.from_numpy(xs).to(_dev()), agc=False).payloads() == pk

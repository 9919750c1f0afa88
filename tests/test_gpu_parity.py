"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bar: trigger indices, header fields, flags and payload bytes bit-exact; pre-decision
equalised symbols within 1e-4 relative EVM; FFT / TX samples within 1e-5 relative L2."""
import numpy as np
import pytest

import common as cm

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _to_dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(_dev())


def _compare_rx(cfg, stream, byte_stride=None, check_z=True):
    orc = cm.make_oracle(cfg)
    phy = cm.make_phy(cfg)
    ref = orc.rx(stream, byte_stride=phy.byte_stride, want_z=check_z, max_pkt_syms=phy.max_pkt_bytes * 8 // phy.bps_payload + 1)
    res = phy.rx(_to_dev(stream), want_z=check_z)
    trig, cfo, st = phy.sync(_to_dev(stream))
    assert np.array_equal(trig, ref["triggers"]), "trigger indices differ"
    np.testing.assert_allclose(cfo, ref["cfo"], rtol=0, atol=2e-6)
    assert res.n_triggers == len(ref["triggers"])
    f, g = res.frames, ref["frames"]
    assert len(f) == len(g), "frame count %d vs oracle %d" % (len(f), len(g))
    assert np.array_equal(f["trigger"], g["trigger"])
    assert np.array_equal(f["pkt_len"], g["pkt_len"])
    assert np.array_equal(f["pkt_num"], g["pkt_num"])
    assert np.array_equal(f["frame_syms"], g["frame_syms"].astype(np.uint16))
    assert np.array_equal(f["carr_offset"], g["carr_offset"].astype(np.int16))
    assert np.array_equal(f["flags"] & 7, g["flags"] & 7)
    slots = res.slots.cpu().numpy()
    for i in range(len(f)):
        n = int(f["pkt_len"][i])
        assert np.array_equal(slots[int(f["slot"][i]), :n], ref["bytes"][i, :n]), "payload bytes differ in frame %d" % i
    if check_z and len(f):
        z = res.z.cpu().numpy()
        hl = phy.header_len()
        for i in range(len(f)):
            ns = hl + (int(f["pkt_len"][i]) * 8 + phy.bps_payload - 1) // phy.bps_payload
            e = cm.rel_evm(z[int(f["slot"][i]), :ns], ref["z"][i, :ns])
            assert e <= 1e-4, "relative EVM %.3g in frame %d" % (e, i)
    assert res.payloads() == orc.payloads(ref)
    return res, ref


def _frames(cfg, rng, n, length):
    orc = cm.make_oracle(cfg)
    pk = cm.rand_packets(rng, n, length)
    s, off = orc.tx(pk)
    return pk, cm.split_frames(s, off)


@pytest.mark.parametrize("n", [64, 128, 1024, 2048])
def test_fft_parity(n):
    import oracle as O
    cfg = cm.cfg_c1() if n == 64 else (cm.cfg_radio128() if n == 128 else (cm.cfg_c3() if n == 1024 else cm.cfg_c4()))
    phy = cm.make_phy(cfg)
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((37, n)) + 1j * rng.standard_normal((37, n))).astype(np.complex64)
    y = phy.fft(_to_dev(x), True).cpu().numpy()
    yi = phy.fft(_to_dev(x), False).cpu().numpy()
    for r in range(x.shape[0]):
        ref = np.fft.fftshift(O.fft(x[r], True))
        assert cm.rel_evm(y[r], ref) <= 1e-5
        refi = O.fft(np.fft.ifftshift(x[r]), False)
        assert cm.rel_evm(yi[r], refi) <= 1e-5


def test_crc32_parity():
    import oracle as O
    phy = cm.make_phy(cm.cfg_c1())
    rng = np.random.default_rng(3)
    pk = [b"", b"1", b"123", b"1234", b"123456789"] + [rng.integers(0, 256, int(l), dtype=np.uint8).tobytes()
                                                       for l in [5, 15, 16, 17, 100, 1500, 4095, 4096, 4097, 8192, 10000]]
    got = phy.crc32(pk)
    assert [int(v) for v in got] == [O.crc32(b) for b in pk]
    assert int(got[4]) == 0xCBF43926


@pytest.mark.parametrize("cfg,length", [
    (cm.cfg_c1(2), 96), (cm.cfg_c1(1, True, 1), 50), (cm.cfg_c1(3, True, 1), 100), (cm.cfg_c1(4, True, 0), 7),
    (cm.cfg_radio128(2, 1, 1), 350), (cm.cfg_c3(), 1500), (cm.cfg_c4(), 1500),
])
def test_tx_parity(cfg, length):
    orc = cm.make_oracle(dict(cfg, tx_scale=0.01))
    phy = cm.make_phy(dict(cfg, tx_scale=0.01))
    rng = np.random.default_rng(11)
    pk = cm.rand_packets(rng, 5, length) + cm.rand_packets(rng, 2, max(1, length // 3))
    ref, roff = orc.tx(pk, first_pkt_num=4093)
    got, goff = phy.tx(pk, first_pkt_num=4093)
    assert np.array_equal(goff.cpu().numpy(), roff)
    assert cm.rel_evm(got.cpu().numpy(), ref) <= 1e-5


@pytest.mark.parametrize("bps,scr,crc", [(2, False, 0), (1, True, 1), (3, True, 1), (4, True, 1), (6, False, 1)])
def test_rx_parity_c1(bps, scr, crc):
    cfg = cm.cfg_c1(bps, scr, crc)
    rng = np.random.default_rng(100 + bps)
    pk, fr = _frames(cfg, rng, 24, 96)
    stream = cm.channel(fr, rng, snr_db=25.0 if bps < 6 else 45.0, cfo=0.17, fft_len=64, scale=0.01)
    res, ref = _compare_rx(cfg, stream)
    assert len(res.frames) == 24
    assert res.payloads() == pk


def test_rx_parity_radio128():
    cfg = cm.cfg_radio128()
    rng = np.random.default_rng(7)
    pk, fr = _frames(cfg, rng, 12, 350)
    stream = cm.channel(fr, rng, snr_db=22.0, cfo=-0.3, fft_len=128, taps=cm.MULTIPATH[:3])
    res, ref = _compare_rx(cfg, stream)
    assert res.payloads() == pk


@pytest.mark.parametrize("int_off,snr", [(0, 40.0), (2, 40.0), (-2, 40.0), (0, 25.0)])
def test_rx_parity_c3(int_off, snr):
    """config 3 (back-to-back frames, CFO 0.3 + integer offset, 4-tap multipath).  At 25 dB the
    reference's alpha=0.1 decision-directed equaliser makes uncoded 16-QAM packets fail their
    CRC-32: that run pins the failure path (flags and wrong bytes identical to the oracle)."""
    cfg = cm.cfg_c3()
    rng = np.random.default_rng(33 + int_off)
    pk, fr = _frames(cfg, rng, 6, 1500)
    stream = cm.channel(fr, rng, gaps=(0, 0), lead=777, tail=3000, snr_db=snr, cfo=0.3 + int_off, fft_len=1024,
                        taps=cm.MULTIPATH)
    res, ref = _compare_rx(cfg, stream)
    assert len(res.frames) == 6
    assert np.all(res.frames["carr_offset"] == int_off)
    if snr >= 40.0:
        assert res.payloads() == pk
    else:
        assert not np.all(res.frames["flags"] & 2)      # some CRC failures, same ones as the oracle


def test_rx_parity_c4_64qam():
    cfg = cm.cfg_c4()
    rng = np.random.default_rng(44)
    pk, fr = _frames(cfg, rng, 4, 1500)
    stream = cm.channel(fr, rng, gaps=(100, 900), tail=3000, snr_db=50.0, cfo=0.1, fft_len=2048)
    res, ref = _compare_rx(cfg, stream)
    assert res.payloads() == pk


def test_rx_noise_only_and_empty():
    cfg = cm.cfg_c1(2)
    rng = np.random.default_rng(5)
    noise = (rng.standard_normal(50000) + 1j * rng.standard_normal(50000)).astype(np.complex64)
    res, ref = _compare_rx(cfg, noise, check_z=False)
    assert len(res.frames) == 0
    res, ref = _compare_rx(cfg, np.zeros(5000, np.complex64), check_z=False)
    assert len(res.frames) == 0 and res.n_triggers == 0
    res, ref = _compare_rx(cfg, np.zeros(17, np.complex64), check_z=False)     # shorter than one window


def test_rx_truncated_and_corrupt():
    """Frame cut by the buffer end is not emitted; a corrupted header makes the demux resume its
    search right after the failed trigger; ragged packet lengths."""
    cfg = cm.cfg_c1(2, True, 1)
    rng = np.random.default_rng(9)
    orc = cm.make_oracle(cfg)
    pk = [rng.integers(0, 256, int(l), dtype=np.uint8).tobytes() for l in [1, 17, 96, 255, 300, 96, 40]]
    s, off = orc.tx(pk)
    fr = cm.split_frames(s, off)
    fr[2] = fr[2].copy()
    fr[2][2 * 80 + 16:3 * 80] *= -1.0          # flip the header symbol -> CRC-8 failure
    stream = cm.channel(fr, rng, snr_db=30.0, cfo=0.05, fft_len=64, tail=5)
    cut = len(stream) - 200                      # cuts into the last frame (560 samples long)
    res, ref = _compare_rx(cfg, stream[:cut])
    got = res.payloads()
    assert got == [pk[0], pk[1], pk[3], pk[4], pk[5]]
    assert res.n_triggers >= 7                   # corrupted and truncated frames still trigger


def test_rx_back_to_back_and_multistream():
    cfg = cm.cfg_c1(2, False, 0)
    rng = np.random.default_rng(21)
    n = 30000
    streams = []
    for s in range(5):
        pk, fr = _frames(cfg, rng, 10 + s, 96)
        x = cm.channel(fr, rng, gaps=(0, 0) if s % 2 else (100, 400), lead=50 * s + 1, tail=10, snr_db=28.0, cfo=0.02 * s)
        y = np.zeros(n, np.complex64)
        y[: len(x)] = x[:n]
        y[len(x):] = (1e-3 * (rng.standard_normal(n - len(x)) + 1j * rng.standard_normal(n - len(x)))).astype(np.complex64) if len(x) < n else 0
        streams.append(y)
    batch = np.stack(streams)
    phy = cm.make_phy(cfg)
    orc = cm.make_oracle(cfg)
    res = phy.rx(_to_dev(batch))
    k = 0
    slots = res.slots.cpu().numpy()
    for s in range(5):
        ref = orc.rx(batch[s], byte_stride=phy.byte_stride, want_z=False)
        for i in range(len(ref["frames"])):
            f = res.frames[k]
            assert f["stream"] == s and f["trigger"] == ref["frames"]["trigger"][i]
            assert np.array_equal(slots[int(f["slot"]), :96], ref["bytes"][i, :96])
            k += 1
    assert k == len(res.frames)


def test_rx_host_path_matches_device_path():
    cfg = cm.cfg_c1(2, True, 1)
    rng = np.random.default_rng(77)
    pk, fr = _frames(cfg, rng, 9, 120)
    stream = cm.channel(fr, rng, snr_db=25.0, cfo=0.1)
    phy = cm.make_phy(cfg)
    a = phy.rx(_to_dev(stream))
    b = phy.rx_host(stream)
    assert np.array_equal(a.frames, b.frames)
    assert a.payloads() == b.payloads() == pk


def test_sync_stress_sparse_frames():
    """config-5 style: long noise with sparse embedded frames; detection indices exact."""
    cfg = cm.cfg_c3()
    rng = np.random.default_rng(5)
    orc = cm.make_oracle(cfg)
    pk, fr = _frames(cfg, rng, 3, 1500)
    n = 400000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)
    amp = np.sqrt(100.0) / np.sqrt(np.mean(np.abs(fr[0]) ** 2))      # 20 dB SNR (0.9 threshold needs > 12.7 dB)
    for i, f in enumerate(fr):
        o = 30000 + i * 120000
        x[o:o + len(f)] += amp * f
    x = x.astype(np.complex64)
    phy = cm.make_phy(cfg)
    trig, cfo, st = phy.sync(_to_dev(x))
    rt, rc = orc.sync(x)
    assert np.array_equal(trig, rt) and len(trig) >= 3
    np.testing.assert_allclose(cfo, rc, atol=2e-6, rtol=0)


def test_large_roundtrip_properties():
    """Size-independent properties at a size the oracle cannot finish: TX -> channel -> RX on the
    GPU returns every packet, CRC ok, in order."""
    cfg = cm.cfg_c3()
    phy = cm.make_phy(cfg, tx_scale=0.01)
    rng = np.random.default_rng(1)
    n_pkts = 1024
    pk = cm.rand_packets(rng, n_pkts, 1500)
    s, off = phy.tx(pk)
    assert int(off[-1]) == n_pkts * 9864
    g = torch.Generator(device=_dev()).manual_seed(2)
    pw = float((s.abs() ** 2).mean())
    noise = torch.randn(s.shape[0], 2, device=_dev(), generator=g) * np.sqrt(pw / 10 ** 4.0 / 2)
    t = torch.arange(s.shape[0], device=_dev(), dtype=torch.float64)
    rot = torch.polar(torch.ones_like(t), 2 * np.pi * 0.3 / 1024 * t).to(torch.complex64)
    y = (s * rot + torch.view_as_complex(noise)).contiguous()
    y = torch.cat([torch.zeros(500, dtype=torch.complex64, device=_dev()), y, torch.zeros(3000, dtype=torch.complex64, device=_dev())])
    res = phy.rx(y)
    assert len(res.frames) == n_pkts
    assert np.all(res.frames["flags"] & 2)
    assert np.array_equal(res.frames["pkt_num"], np.arange(n_pkts) & 0xFFF)
    assert res.payloads() == pk
    d = np.diff(res.frames["trigger"])
    assert np.all(np.abs(d - 9864) <= 72)


@pytest.mark.parametrize("n_seg", [2, 5])
def test_segmented_stream_equals_unsplit(n_seg):
    """one long stream cut into segments (one per GPU) with halos, every trigger's record
    returned (emit-all), demux chain re-run over the owned triggers on the host == unsplit result."""
    from ofdm_tools import dist
    cfg = cm.cfg_c1(2, True, 1)
    rng = np.random.default_rng(123)
    pk, fr = _frames(cfg, rng, 60, 96)
    x = cm.channel(fr, rng, gaps=(0, 300), snr_db=28.0, cfo=0.11, tail=400)
    phy = cm.make_phy(cfg, max_pkt_bytes=128)
    whole = phy.rx(_to_dev(x))
    recs, payloads = dist.rx_segmented(phy, _to_dev(x), n_seg)
    assert np.array_equal(recs["trigger"], whole.frames["trigger"])
    assert np.array_equal(recs["pkt_num"], whole.frames["pkt_num"])
    assert np.array_equal(recs["flags"], whole.frames["flags"])
    assert payloads == whole.payloads() == pk


def test_emit_all_returns_every_trigger():
    cfg = cm.cfg_c1(2, True, 1)
    rng = np.random.default_rng(8)
    pk, fr = _frames(cfg, rng, 8, 40)
    fr[3] = fr[3].copy()
    fr[3][2 * 80 + 16:3 * 80] *= -1.0
    x = cm.channel(fr, rng, snr_db=30.0)
    phy = cm.make_phy(cfg)
    normal = phy.rx(_to_dev(x))
    phy.set_emit_all(True)
    every = phy.rx(_to_dev(x))
    phy.set_emit_all(False)
    assert len(every.frames) == every.n_triggers >= 8 and len(normal.frames) == 7
    acc = every.frames[(every.frames["flags"] & 8) != 0]
    ok = acc[(acc["flags"] & 5) == 5]
    assert np.array_equal(ok["trigger"], normal.frames["trigger"])


def test_generic_kernel_at_1024_matches_fast_path(monkeypatch):
    """The any-fft_len frame kernel and the fft_len-1024 fast path give identical records and bytes."""
    cfg = cm.cfg_c3()
    rng = np.random.default_rng(55)
    pk, fr = _frames(cfg, rng, 5, 700)
    x = cm.channel(fr, rng, gaps=(0, 50), lead=300, tail=3000, snr_db=40.0, cfo=-0.2, fft_len=1024, taps=cm.MULTIPATH)
    fast = cm.make_phy(cfg).rx(_to_dev(x), want_z=True)
    monkeypatch.setenv("OFDMX_FORCE_GENERIC", "1")
    gen = cm.make_phy(cfg).rx(_to_dev(x), want_z=True)
    monkeypatch.setenv("OFDMX_FORCE_GENERIC", "0")
    monkeypatch.setenv("OFDMX_NO_TMA", "1")
    tma = cm.make_phy(cfg).rx(_to_dev(x))          # plain-load sync kernel
    monkeypatch.setenv("OFDMX_NO_TMA", "0")
    monkeypatch.setenv("OFDMX_NO_WARP_SYNC", "1")
    ring = cm.make_phy(cfg).rx(_to_dev(x))         # TMA ring sync kernel (fft_len 1024 defaults to the warp-autonomous one)
    assert np.array_equal(fast.frames, gen.frames) and fast.payloads() == gen.payloads() == pk
    assert np.array_equal(fast.frames, tma.frames) and tma.payloads() == pk
    assert np.array_equal(fast.frames, ring.frames) and ring.payloads() == pk
    assert cm.rel_evm(fast.z.cpu().numpy()[:5, :3000], gen.z.cpu().numpy()[:5, :3000]) < 1e-5


def test_facade_loopbacks():
    """The reference's manual loopback (examples/radioA.grc: hier port-1 output fed back into its
    port-1 input) through the drop-in classes, with payload_source / payload_sink at the ends."""
    import threading
    from ofdm_tools import ofdm_radio_hier, ofdm_tx_rx_hier, payload_source, payload_sink
    rng = np.random.default_rng(4)
    for radio, plen in ((ofdm_radio_hier(payload_mod='qam16', scramble_mode=1, crc_mode=1), 350),
                        (ofdm_radio_hier(), 350), (ofdm_tx_rx_hier(fft_len=64, payload_bps=2), 96)):
        src = payload_source(packet_len=plen)
        data = rng.integers(0, 256, plen * 7, dtype=np.uint8).tobytes()
        src.send_pkt_s(data[: plen * 3 + 5])
        src.send_pkt_s(data[plen * 3 + 5:])
        pk = src.pop_packets()
        assert len(pk) == 7 and b"".join(pk) == data
        samples, offsets = radio.tx(pk)
        assert samples.abs().max() < 1.0                       # x0.01 scaling applied
        pad = torch.zeros(700, dtype=torch.complex64, device=samples.device)
        res = radio.rx(torch.cat([pad, samples, pad]))
        got, done = [], threading.Event()
        snk = payload_sink(lambda p: (got.append(p), done.set() if len(got) == 7 else None))
        snk.deliver(res.payloads())
        assert done.wait(5.0) and got == pk
        assert np.array_equal(res.frames["pkt_num"], np.arange(7))
    # second TX call continues the header packet counter (packet_header_default::d_header_number)
    s2, _ = radio.tx(pk[:2])
    r2 = radio.rx(torch.cat([pad, s2, pad]))
    assert list(r2.frames["pkt_num"]) == [7, 8]


def test_frame_kernel_variants_agree(monkeypatch):
    """fft_len 1024: warp-per-frame kernel (default when eligible) vs CTA-per-frame kernel vs oracle."""
    cfg = cm.cfg_c3()
    rng = np.random.default_rng(66)
    pk, fr = _frames(cfg, rng, 7, 1500)
    x = cm.channel(fr, rng, gaps=(0, 0), lead=100, tail=3000, snr_db=40.0, cfo=0.3 - 2, fft_len=1024, taps=cm.MULTIPATH)
    warp = cm.make_phy(cfg, max_pkt_bytes=1504).rx(_to_dev(x), want_z=True)
    monkeypatch.setenv("OFDMX_NO_WARP_FRAME", "1")
    cta = cm.make_phy(cfg, max_pkt_bytes=1504).rx(_to_dev(x), want_z=True)
    assert np.array_equal(warp.frames, cta.frames) and warp.payloads() == cta.payloads() == pk
    assert np.all(warp.frames["carr_offset"] == -2)
    assert cm.rel_evm(warp.z.cpu().numpy()[:7, :3608], cta.z.cpu().numpy()[:7, :3608]) < 1e-5
    ref = cm.make_oracle(cfg).rx(x, byte_stride=1504, want_z=True, max_pkt_syms=1504 * 2 + 1)
    assert np.array_equal(warp.frames["trigger"], ref["frames"]["trigger"])
    for i in range(7):
        assert cm.rel_evm(warp.z.cpu().numpy()[i, :3608], ref["z"][i, :3608]) <= 1e-4


@pytest.mark.parametrize("bps", [1, 2, 3, 4, 6])
def test_warp_frame_kernel_ragged_lengths(bps):
    """fft_len 1024 warp-per-frame kernel: every payload modulation, packet lengths around the edges of the
    word-wise pack and CRC code (1..5 bytes, 64-byte lane chunks, OFDM-symbol multiples, > 2048 bytes, the
    12-bit maximum), default max_pkt_bytes (4095).  Bytes and flags bit-exact against the oracle."""
    cfg = cm.cfg_c3(bps_payload=bps)
    rng = np.random.default_rng(900 + bps)
    sym = 600 * bps // 8
    lens = [1, 2, 3, 4, 5, 59, 60, 61, 63, 64, 65, 124, 127, 128, sym - 4, sym - 3, sym, sym + 1, 1000, 1499, 2043,
            2044, 2045, 2047, 2052, 2053, 3001, 4091]
    if bps == 1:
        lens = [n for n in lens if n <= 1499]          # keep the BPSK stream short
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    s, off = cm.make_oracle(cfg).tx(pk)
    stream = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 300), lead=500, tail=3000, snr_db=60.0, cfo=0.2,
                        fft_len=1024)
    res, ref = _compare_rx(cfg, stream, check_z=False)
    assert res.payloads() == pk
    assert np.all(res.frames["flags"] & 2)


@pytest.mark.parametrize("n_streams,n,odd", [(1, 100000, False), (1, 77777, True), (3, 40001, False), (5, 3000, False),
                                              (2, 600, False), (1, 16, False)])
def test_sync_kernel_variants_detect_bits(monkeypatch, n_streams, n, odd):
    """fft_len 1024: warp-autonomous sync kernel (default) vs TMA ring kernel vs plain-load kernel vs oracle on
    streams with frames near both ends, odd lengths and several streams: identical triggers and CFO."""
    cfg = cm.cfg_c3()
    rng = np.random.default_rng(n + n_streams)
    orc = cm.make_oracle(cfg)
    pk, fr = _frames(cfg, rng, 2, 300)
    xs = []
    for s in range(n_streams):
        x = 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        for o in (0, n // 2 + 37 * s, n - len(fr[0]) // 2):          # at the very start, mid, cut off at the end
            if 0 <= o < n:
                m = min(len(fr[0]), n - o)
                x[o:o + m] += fr[s % 2][:m]
        xs.append(x.astype(np.complex64))
    x = np.stack(xs)
    res = {}
    for name, env in (("warp", {}), ("ring", {"OFDMX_NO_WARP_SYNC": "1"}), ("plain", {"OFDMX_NO_TMA": "1"})):
        for k in ("OFDMX_NO_WARP_SYNC", "OFDMX_NO_TMA"):
            monkeypatch.setenv(k, env.get(k, "0"))
        phy = cm.make_phy(cfg)
        res[name] = phy.sync(_to_dev(x if n_streams > 1 else x[0]))
    ref_t, ref_s, ref_c = [], [], []
    for s in range(n_streams):
        t, c = orc.sync(x[s])
        ref_t += list(t); ref_c += list(c); ref_s += [s] * len(t)
    for name in ("warp", "ring", "plain"):
        trig, cfo, st = res[name]
        assert np.array_equal(trig, np.array(ref_t, np.int64)), name
        assert np.array_equal(st, np.array(ref_s)), name
        np.testing.assert_allclose(cfo, np.array(ref_c, np.float32), atol=2e-6, rtol=0)


@pytest.mark.parametrize("bps,crc,scr,clip", [(1, 1, True, 0.0), (2, 0, False, 0.0), (3, 1, True, 0.0), (4, 1, True, 0.004),
                                              (6, 1, True, 0.0), (4, 0, True, 0.0)])
def test_tx_warp_kernel_ragged(monkeypatch, bps, crc, scr, clip):
    """fft_len 1024 warp-per-packet TX kernel (default when eligible): every payload modulation, with and
    without in-graph CRC / scrambler / clipper, packet lengths around the byte-window and symbol edges and packets
    starting at unaligned payload offsets -- against the oracle and against the generic TX kernel."""
    cfg = cm.cfg_c3(bps_payload=bps, crc_mode=crc, scramble_bits=scr)
    cfg["tx_scale"] = 0.01
    rng = np.random.default_rng(700 + bps + crc)
    sym = 600 * bps // 8
    lens = [1, 2, 3, 4, 5, 15, 16, 17, 63, 64, 65, sym - 4, sym - 1, sym, sym + 1, 2 * sym, 999, 1500, 2047, 2048, 2049]
    if bps > 1:
        lens += [3001, 4091 if crc else 4095]
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    ref, roff = cm.make_oracle(dict(cfg, tx_clip=clip)).tx(pk, first_pkt_num=4090)     # the counter wraps at 4096
    phy = cm.make_phy(cfg, tx_clip=clip)
    s, soff = phy.tx(pk, first_pkt_num=4090)
    s = s.cpu().numpy()
    assert np.array_equal(soff.cpu().numpy(), roff)
    for i in range(len(pk)):
        a, b = s[roff[i]:roff[i + 1]], ref[roff[i]:roff[i + 1]]
        e = np.linalg.norm(a - b) / np.linalg.norm(b)
        assert e < 1e-5, "frame %d (len %d): relative error %.3g" % (i, lens[i], e)
    if clip:
        assert np.abs(s.real).max() <= np.float32(clip) and (np.abs(s.real) == np.float32(clip)).sum() > 100
    assert "tx_framew_kernel" in _kernels_used(phy, lambda: phy.tx(pk))
    monkeypatch.setenv("OFDMX_NO_WARP_TX", "1")
    gen = cm.make_phy(cfg, tx_clip=clip)
    g, goff = gen.tx(pk, first_pkt_num=4090)
    assert "tx_frame_kernel" in _kernels_used(gen, lambda: gen.tx(pk))
    assert np.array_equal(goff.cpu().numpy(), roff)
    assert np.linalg.norm(g.cpu().numpy() - s) / np.linalg.norm(s) < 1e-5
    if clip:
        return          # the rail at 0.004 distorts the signal on purpose
    # and the receiver decodes what the new transmitter sends
    x = cm.channel(cm.split_frames(s, roff), rng, gaps=(0, 200), lead=400, tail=3000, snr_db=60.0, fft_len=1024,
                   scale=100.0)
    rx = cm.make_phy(cfg).rx(_to_dev(x), want_z=False)
    assert rx.payloads() == pk


def _kernels_used(phy, fn):
    phy.profile(True)
    fn()
    used = set(phy.profile_read())
    phy.profile(False)
    return used


@pytest.mark.parametrize("which,bps,int_off", [("c1", 1, 0), ("c1", 2, 2), ("c1", 3, -2), ("c1", 4, 0), ("c1", 6, 0),
                                               ("radio128", 1, 0), ("radio128", 2, -2), ("radio128", 3, 2),
                                               ("radio128", 4, 0), ("radio128", 6, 0)])
def test_warp_frame_kernel_small_fft(which, bps, int_off):
    """Warp-per-frame receiver at fft_len 64 (ofdm_tx_rx_hier plan, unlimited carrier-offset search) and fft_len 128
    (ofdm_radio_hier plan: 103 data carriers, so a symbol is not a whole number of bytes for most modulations and
    the packet is packed at the end): every payload modulation, ragged packet lengths, integer carrier offsets;
    records, bytes and equalised symbols against the oracle."""
    if which == "c1":
        cfg, n_fft = cm.cfg_c1(bps, True, 1), 64
    else:
        cfg, n_fft = cm.cfg_radio128(bps, 1, 1), 128
    rng = np.random.default_rng(1000 + 7 * bps + int_off + n_fft)
    lens = [1, 2, 3, 4, 5, 11, 12, 13, 47, 48, 49, 96, 100, 255, 256, 350, 351]
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    s, off = cm.make_oracle(cfg).tx(pk)
    stream = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 400), lead=300, tail=1500, snr_db=45.0, cfo=0.15 + int_off,
                        fft_len=n_fft)
    phy = cm.make_phy(cfg)
    phy.profile(True)
    res, ref = _compare_rx(cfg, stream)
    used = set(phy.profile_read())      # profile of this phy is empty: _compare_rx builds its own; check a fresh call
    phy.profile(True)
    r2 = phy.rx(_to_dev(stream))
    assert "rx_framew_kernel" in set(phy.profile_read())
    assert np.array_equal(r2.frames, res.frames)
    assert res.payloads() == pk
    assert np.all(res.frames["carr_offset"] == int_off)


@pytest.mark.parametrize("which,bps", [("c1", 1), ("c1", 2), ("c1", 3), ("c1", 4), ("c1", 6),
                                       ("radio128", 1), ("radio128", 2), ("radio128", 3), ("radio128", 4), ("radio128", 6)])
def test_tx_warp_kernel_small_fft(monkeypatch, which, bps):
    """Warp-per-packet TX kernel at fft_len 64 and 128 (register FFT + lane-shuffle FFT on conjugated data): every
    payload modulation, ragged packet lengths, against the oracle and the generic TX kernel; then decoded by the
    receiver."""
    cfg = cm.cfg_c1(bps, True, 1) if which == "c1" else cm.cfg_radio128(bps, 1, 1)
    cfg["tx_scale"] = 0.01
    rng = np.random.default_rng(300 + bps + len(which))
    lens = [1, 2, 3, 4, 5, 11, 12, 13, 47, 48, 49, 96, 100, 255, 256, 350, 351, 1000]
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    ref, roff = cm.make_oracle(cfg).tx(pk, first_pkt_num=7)
    phy = cm.make_phy(cfg)
    s, soff = phy.tx(pk, first_pkt_num=7)
    s = s.cpu().numpy()
    assert "tx_framew_kernel" in _kernels_used(phy, lambda: phy.tx(pk))
    assert np.array_equal(soff.cpu().numpy(), roff)
    for i in range(len(pk)):
        a, b = s[roff[i]:roff[i + 1]], ref[roff[i]:roff[i + 1]]
        e = np.linalg.norm(a - b) / np.linalg.norm(b)
        assert e < 1e-5, "frame %d (len %d): relative error %.3g" % (i, lens[i], e)
    monkeypatch.setenv("OFDMX_NO_WARP_TX", "1")
    gen = cm.make_phy(cfg)
    g, goff = gen.tx(pk, first_pkt_num=7)
    assert "tx_frame_kernel" in _kernels_used(gen, lambda: gen.tx(pk))
    assert np.linalg.norm(g.cpu().numpy() - s) / np.linalg.norm(s) < 1e-5
    x = cm.channel(cm.split_frames(s, roff), rng, gaps=(0, 200), lead=400, tail=1500, snr_db=60.0, fft_len=cfg["fft_len"],
                   scale=100.0)
    assert cm.make_phy(cfg).rx(_to_dev(x), want_z=False).payloads() == pk


@pytest.mark.parametrize("bps,int_off", [(2, 0), (4, 2), (6, -2)])
def test_warp_frame_kernel_fft2048(monkeypatch, bps, int_off):
    """Warp-per-frame receiver at fft_len 2048 (two interleaved 1024-point register transforms recombined in the bin
    accessor): ragged lengths, integer carrier offsets, multipath; records, bytes and equalised symbols against the
    oracle."""
    cfg = cm.cfg_c4(bps_payload=bps)
    rng = np.random.default_rng(2048 + bps)
    lens = [1, 4, 5, 149, 150, 151, 899, 900, 901, 1500, 2999]
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    s, off = cm.make_oracle(cfg).tx(pk)
    stream = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 700), lead=900, tail=5000, snr_db=50.0, cfo=0.25 + int_off,
                        fft_len=2048, taps=cm.MULTIPATH)
    res, ref = _compare_rx(cfg, stream)
    phy = cm.make_phy(cfg)
    # default at fft_len 2048: the pair-of-warps-per-frame kernel (byte-multiple symbols), else one warp per frame
    used = _kernels_used(phy, lambda: phy.rx(_to_dev(stream)))
    assert ("rx_framep_kernel" in used) if (1200 * bps) % 8 == 0 else ("rx_framew_kernel" in used), used
    assert res.payloads() == pk
    assert np.all(res.frames["carr_offset"] == int_off)
    # and the one-warp-per-frame kernel on the same input
    monkeypatch.setenv("OFDMX_NO_PAIR_FRAME", "1")
    res1, _ = _compare_rx(cfg, stream)
    phy1 = cm.make_phy(cfg)
    assert "rx_framew_kernel" in _kernels_used(phy1, lambda: phy1.rx(_to_dev(stream)))
    assert res1.payloads() == pk


@pytest.mark.parametrize("which,n_streams,n", [("c1", 1, 50000), ("c1", 4, 3000), ("c1", 1, 33333), ("radio128", 1, 60000),
                                               ("radio128", 3, 5000), ("c1", 2, 40)])
def test_sync_kernel_variants_small_fft(monkeypatch, which, n_streams, n):
    """fft_len 64 / 128: short-window warp-autonomous sync kernel (default) vs TMA ring kernel vs plain-load kernel vs
    oracle, frames at the very start, in the middle and cut off at the end, several streams: identical triggers and
    CFO."""
    cfg = cm.cfg_c1(2, False, 0) if which == "c1" else cm.cfg_radio128(2, 1, 1)
    rng = np.random.default_rng(n + n_streams)
    orc = cm.make_oracle(cfg)
    pk, fr = _frames(cfg, rng, 2, 60)
    xs = []
    for s in range(n_streams):
        x = 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        for o in (0, n // 3 + 11 * s, 2 * n // 3, n - len(fr[0]) // 2):
            if 0 <= o < n:
                m = min(len(fr[0]), n - o)
                x[o:o + m] += fr[s % 2][:m]
        xs.append(x.astype(np.complex64))
    x = np.stack(xs)
    res = {}
    for name, env in (("warp", {}), ("ring", {"OFDMX_NO_WARP_SYNC": "1"}), ("plain", {"OFDMX_NO_TMA": "1"})):
        for k in ("OFDMX_NO_WARP_SYNC", "OFDMX_NO_TMA"):
            monkeypatch.setenv(k, env.get(k, "0"))
        phy = cm.make_phy(cfg)
        res[name] = phy.sync(_to_dev(x if n_streams > 1 else x[0]))
    ref_t, ref_s, ref_c = [], [], []
    for s in range(n_streams):
        t, c = orc.sync(x[s])
        ref_t += list(t); ref_c += list(c); ref_s += [s] * len(t)
    assert len(ref_t) >= (2 if n > 1000 else 0)
    for name in ("warp", "ring", "plain"):
        trig, cfo, st = res[name]
        assert np.array_equal(trig, np.array(ref_t, np.int64)), name
        assert np.array_equal(st, np.array(ref_s)), name
        np.testing.assert_allclose(cfo, np.array(ref_c, np.float32), atol=2e-6, rtol=0)


@pytest.mark.parametrize("n_fft,bps,int_off", [(256, 2, 0), (256, 4, 2), (512, 3, -2), (512, 6, 0)])
def test_warp_kernels_fft256_512(monkeypatch, n_fft, bps, int_off):
    """fft_len 256 and 512 on the carrier plans `spectrum_enforcer` derives (python/ofdm_cr_tools.py:348-378): the
    warp-per-packet TX kernel against the oracle and the generic kernel, then the warp-per-frame receiver against the
    oracle (records, bytes, equalised symbols), ragged lengths and integer carrier offsets."""
    from ofdm_tools import ofdm_cr_tools as T
    occ, pil, pls, sw1, sw2 = T.spectrum_enforcer(n_fft, [], 10)
    cfg = dict(fft_len=n_fft, cp_len=n_fft // 4, occupied_carriers=[list(occ[0])], pilot_carriers=[list(pil[0])],
               pilot_symbols=[list(pls[0])], sync_word1=[complex(v) for v in sw1], sync_word2=[complex(v) for v in sw2],
               bps_header=1, bps_payload=bps, scramble_bits=True, scramble_header=True, crc_mode=1, max_carr_offset=3,
               tx_scale=0.01)
    rng = np.random.default_rng(n_fft + bps)
    lens = [1, 3, 4, 5, 63, 64, 65, 200, 511, 512, 513, 1000]
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    orc = cm.make_oracle(cfg)
    ref, roff = orc.tx(pk)
    phy = cm.make_phy(cfg)
    s, soff = phy.tx(pk)
    s = s.cpu().numpy()
    assert "tx_framew_kernel" in _kernels_used(phy, lambda: phy.tx(pk))
    assert np.array_equal(soff.cpu().numpy(), roff)
    assert np.linalg.norm(s - ref) / np.linalg.norm(ref) < 1e-5
    stream = cm.channel(cm.split_frames(ref, roff), rng, gaps=(0, 500), lead=400, tail=3000, snr_db=45.0, cfo=0.2 + int_off,
                        fft_len=n_fft, scale=100.0)
    res, rr = _compare_rx(cfg, stream)
    assert "rx_framew_kernel" in _kernels_used(phy, lambda: phy.rx(_to_dev(stream)))
    assert res.payloads() == pk
    assert np.all(res.frames["carr_offset"] == int_off)

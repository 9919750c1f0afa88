"""Round-2 GPU parity tests: branches the first round left untested (several occupied-carrier / pilot sets,
pilots inside the occupied set, fft_len 32 and 4096 on the generic kernels, unaligned payload slots), the
oversize-frame rule of the demux chain, the GNU Radio version switches, and reconfiguration of the slot size."""
import numpy as np
import pytest

import common as cm

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from test_gpu_parity import _compare_rx, _dev, _frames, _to_dev  # noqa: E402


def _multiset(base, fft_len):
    """Two occupied-carrier sets and two pilot sets; the pilots of set 1 lie inside occupied set 0, so the
    equaliser's pilot branch is reachable (ofdm_equalizer_simpledfe visits the union of the occupied sets)."""
    occ0 = list(base["occupied_carriers"][0])
    p0 = list(base["pilot_carriers"][0])
    step = len(occ0) // 5
    p1 = [occ0[step], occ0[2 * step], occ0[3 * step], occ0[4 * step]]
    occ1 = [c for c in occ0 if c not in p1]
    d = dict(base, occupied_carriers=[occ0, occ1], pilot_carriers=[p0, p1], pilot_symbols=[[1, 1, 1, -1], [1, -1, 1, 1]])
    assert d["fft_len"] == fft_len
    return d


@pytest.mark.parametrize("which,bps", [("c1", 2), ("c1", 4), ("c3", 4), ("c3", 2)])
def test_several_carrier_sets_and_pilot_inside_occupied(which, bps):
    """rx_frame_kernel (fft_len 64) and rx_frame1024_kernel<.., SIMPLE=false, ..> (fft_len 1024) with cycling
    occupied / pilot sets, against the oracle; TX through the generic allocator."""
    rng = np.random.default_rng(21)
    if which == "c1":
        cfg = _multiset(cm.cfg_c1(bps, True, 1), 64)
        plen, kw = 150, {}
    else:
        cfg = _multiset(cm.cfg_c3(bps_payload=bps), 1024)
        plen, kw = 1500, dict(fft_len=1024, taps=cm.MULTIPATH)
    orc, phy = cm.make_oracle(cfg), cm.make_phy(cfg)
    pk = cm.rand_packets(rng, 5, plen)
    s_ref, off_ref = orc.tx(pk)
    s_gpu, off_gpu = phy.tx(pk)
    assert np.array_equal(off_gpu.cpu().numpy(), off_ref)
    assert cm.rel_evm(s_gpu.cpu().numpy(), s_ref) < 1e-5
    stream = cm.channel(cm.split_frames(s_ref, off_ref), rng, gaps=(0, 500), snr_db=35.0, cfo=0.25, lead=400, tail=3000, **kw)
    phy.profile(True)
    res, ref = _compare_rx(cfg, stream)
    assert res.payloads() == pk
    p2 = cm.make_phy(cfg)
    p2.profile(True)
    p2.rx(_to_dev(stream))
    names = set(p2.profile_read())
    assert ("rx_frame1024_kernel" in names) if which == "c3" else ("rx_frame_kernel" in names), names


def _plan(fft_len, n_data, bps_header):
    half = n_data // 2 + 2
    pil = [-(half - half // 3), -(half // 3), half // 3, half - half // 3]
    occ = [k for k in range(-half, half + 1) if k != 0 and k not in pil][:n_data]
    return dict(fft_len=fft_len, cp_len=fft_len // 4, occupied_carriers=[occ], pilot_carriers=[pil],
                pilot_symbols=[[1, 1, 1, -1]], bps_header=bps_header, bps_payload=2, scramble_bits=True, crc_mode=1)


@pytest.mark.parametrize("fft_len,n_data,bps_h,plen", [(32, 20, 2, 40), (4096, 2400, 1, 1500)])
def test_fft_len_32_and_4096(fft_len, n_data, bps_h, plen):
    """The ends of the advertised fft_len range (include/ofdmx.h) on the any-fft_len kernels, TX and RX."""
    rng = np.random.default_rng(fft_len)
    cfg = _plan(fft_len, n_data, bps_h)
    orc, phy = cm.make_oracle(cfg), cm.make_phy(cfg)
    pk = cm.rand_packets(rng, 4, plen)
    s_ref, off_ref = orc.tx(pk)
    s_gpu, off_gpu = phy.tx(pk)
    assert np.array_equal(off_gpu.cpu().numpy(), off_ref)
    assert cm.rel_evm(s_gpu.cpu().numpy(), s_ref) < 1e-5
    stream = cm.channel(cm.split_frames(s_ref, off_ref), rng, gaps=(100, 900), snr_db=35.0, cfo=0.2, fft_len=fft_len,
                        lead=300, tail=3 * fft_len)
    res, ref = _compare_rx(cfg, stream)
    assert res.payloads() == pk


def test_unaligned_payload_slots_take_the_byte_path():
    """bytes_out with a stride that is not a multiple of 16: the warp kernels fall back from word-wise packing /
    CRC to the byte path (crc32_warp); results equal the aligned run and the oracle."""
    rng = np.random.default_rng(33)
    cfg = cm.cfg_c3()
    orc = cm.make_oracle(cfg)
    pk, fr = _frames(cfg, rng, 4, 1500)
    stream = cm.channel(fr, rng, gaps=(0, 0), snr_db=40.0, cfo=0.3, fft_len=1024, taps=cm.MULTIPATH, lead=300, tail=3000)
    ref = orc.rx(stream, byte_stride=1520, want_z=False)
    phy = cm.make_phy(cfg, max_pkt_bytes=1504)
    x = _to_dev(stream)
    bufs = phy.rx_buffers(64, _dev())
    stride = 1509                                            # odd stride: slots start at every alignment
    raw = torch.zeros(64 * stride + 16, dtype=torch.uint8, device=_dev())
    from ofdm_tools import _lib
    import ctypes as C
    _lib.check(_lib.load().ofdmx_rx(phy.ctx, x.data_ptr(), 1, x.numel(), x.numel(), bufs["frames"].data_ptr(), 64,
                                    raw.data_ptr() + 1, stride, None, 0, bufs["counts"].data_ptr(),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), phy.ctx)
    c = bufs["counts"].cpu().numpy()
    from ofdm_tools.phy import FRAME_DTYPE
    rec = np.frombuffer(bufs["frames"][: int(c[1]) * 32].cpu().numpy().tobytes(), FRAME_DTYPE)
    assert np.array_equal(rec["trigger"], ref["frames"]["trigger"]) and np.array_equal(rec["flags"] & 7, ref["frames"]["flags"] & 7)
    host = raw.cpu().numpy()
    for i, f in enumerate(rec):
        a = 1 + int(f["slot"]) * stride
        assert np.array_equal(host[a:a + 1504], ref["bytes"][i, :1504])
    assert np.all(rec["flags"] & 2)


def test_oversize_frame_is_consumed_and_the_stream_goes_on():
    """A header whose length field exceeds max_pkt_bytes: the demux consumes the declared payload and the frames
    behind it are still delivered (the chain used to drop the rest of the stream)."""
    rng = np.random.default_rng(44)
    cfg = cm.cfg_c1(2, True, 1)
    orc = cm.make_oracle(cfg)
    pk = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (96, 300, 96, 40, 300, 96)]
    s, off = orc.tx(pk)
    stream = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 600), snr_db=30.0, cfo=0.1, lead=300, tail=900)
    ref = orc.rx(stream, want_z=False)
    assert len(ref["frames"]) == 6
    from ofdm_tools import _lib
    for force in (False, True):
        import os
        os.environ["OFDMX_FORCE_GENERIC"] = "1" if force else "0"
        try:
            phy = cm.make_phy(cfg, max_pkt_bytes=112)        # 96 + 4 CRC fits, 300 + 4 does not
            res = phy.rx(_to_dev(stream))
        finally:
            os.environ.pop("OFDMX_FORCE_GENERIC", None)
        f = res.frames
        assert np.array_equal(f["trigger"], ref["frames"]["trigger"])          # every frame examined and emitted
        assert np.array_equal(f["pkt_len"], ref["frames"]["pkt_len"])
        over = (f["flags"] & _lib.F_OVERSIZE) != 0
        assert list(over) == [False, True, False, False, True, False]
        assert np.all(f["flags"][over] & _lib.F_COMPLETE) and not np.any(f["flags"][over] & _lib.F_CRC_OK)
        assert res.payloads() == [pk[0], pk[2], pk[3], pk[5]]
        # and the host stitching rule agrees (dist.demux_chain on emit-all records)
        from ofdm_tools import dist
        phy.set_emit_all(True)
        allrec = phy.rx(_to_dev(stream)).frames
        phy.set_emit_all(False)
        emit = dist.demux_chain(allrec, 64, 16, phy.params.demux_holdoff, len(stream))
        assert np.array_equal(allrec["trigger"][emit], f["trigger"])


def test_reconfigure_slot_size_then_rx_host():
    """after reconfigure(max_pkt_bytes=...) the pinned host staging of rx_host follows the new stride."""
    rng = np.random.default_rng(45)
    cfg = cm.cfg_c1(2, True, 1)
    orc = cm.make_oracle(cfg)
    pk = cm.rand_packets(rng, 5, 96)
    s, off = orc.tx(pk)
    stream = cm.channel(cm.split_frames(s, off), rng, gaps=(100, 400), snr_db=30.0, lead=300, tail=900)
    phy = cm.make_phy(cfg, max_pkt_bytes=112)
    assert phy.rx_host(stream).payloads() == pk
    phy.reconfigure(max_pkt_bytes=2000)
    assert phy.byte_stride == 2000 and phy.rx_host(stream).payloads() == pk
    phy.reconfigure(max_pkt_bytes=104)
    assert phy.rx_host(stream).payloads() == pk


def test_version_switches_against_the_oracle():
    """The GNU Radio version switches: 16-QAM amplitude normalisation (>= 3.8), the agc2 rate
    rule with fabsf (>= 3.8), iir_filter_ccd oldstyle=True -- each against the oracle with the same switch."""
    import oracle as O
    rng = np.random.default_rng(46)
    cfg = cm.cfg_c3(qam_normalization=1)
    orc, phy = cm.make_oracle(cfg), cm.make_phy(cfg)
    pk = cm.rand_packets(rng, 3, 1500)
    s_ref, off_ref = orc.tx(pk)
    s_gpu, _ = phy.tx(pk)
    assert cm.rel_evm(s_gpu.cpu().numpy(), s_ref) < 1e-5
    s0, _ = cm.make_oracle(cm.cfg_c3()).tx(pk)
    assert cm.rel_evm(s0, s_ref) > 1e-3                                       # the switch does change the waveform
    stream = cm.channel(cm.split_frames(s_ref, off_ref), rng, gaps=(0, 0), snr_db=40.0, cfo=0.3, fft_len=1024,
                        taps=cm.MULTIPATH, lead=300, tail=3000)
    res, ref = _compare_rx(cfg, stream)
    assert res.payloads() == pk
    # agc2 rate rule
    x = ((rng.standard_normal((3, 5000)) + 1j * rng.standard_normal((3, 5000))) * np.array([[0.01], [1.0], [30.0]])).astype(np.complex64)
    x[:, 2000:2600] = 0
    for ar in (False, True):
        r, g = O.agc2(x, abs_rate=ar)
        y, gg = phy.agc2(_to_dev(x), abs_rate=ar)
        assert np.array_equal(y.cpu().numpy(), r) and np.array_equal(gg.cpu().numpy(), g)
    assert not np.array_equal(O.agc2(x, abs_rate=True)[0], O.agc2(x, abs_rate=False)[0])
    # iir oldstyle
    ff, fb = [0.2, 0.3, 0.1], [1.0, 0.5, -0.2]
    xi = (rng.standard_normal(3000) + 1j * rng.standard_normal(3000)).astype(np.complex64)
    for old in (False, True):
        r, _ = O.iir_ccd(xi, ff, fb, oldstyle=old)
        y, _ = phy.iir_ccd(_to_dev(xi), ff, fb, span=1 << 20, oldstyle=old)
        assert np.array_equal(y.cpu().numpy(), r)


def _guarded(shape, dtype, fill=0xA5, guard=4096):
    """A tensor carved out of the middle of a larger allocation whose borders hold a byte pattern."""
    n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
    n16 = (n + 15) // 16 * 16
    raw = torch.full((guard + n16 + guard,), fill, dtype=torch.uint8, device=_dev())
    view = raw[guard:guard + n].view(dtype).view(*shape)
    return raw, view, guard, n


def _guards_intact(raw, guard, n, fill=0xA5):
    h = raw.cpu().numpy()
    return bool(np.all(h[:guard] == fill) and np.all(h[guard + (n + 15) // 16 * 16:] == fill)
                and np.all(h[guard + n: guard + (n + 15) // 16 * 16] == fill))


def test_guard_bands_and_repeatability():
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer_closed_on_pool.log), so out-of-bounds
    writes and races are looked for directly: every output buffer of RX / TX / sync / AGC / IIR sits between guard
    bands that must come back untouched (buffers sized exactly), and five runs over the same input must agree bit
    for bit (records, payload bytes, pre-decision symbols) on each kernel family."""
    from ofdm_tools import _lib
    import ctypes as C
    L = _lib.load()
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(77)
    for cfg, plen, kw in ((cm.cfg_c1(2, True, 1), 96, {}), (cm.cfg_c3(), 1500, dict(fft_len=1024, taps=cm.MULTIPATH)),
                          (cm.cfg_c4(), 1500, dict(fft_len=2048)), (cm.cfg_radio128(4), 200, dict(fft_len=128))):
        orc = cm.make_oracle(cfg)
        pk = cm.rand_packets(rng, 6, plen)
        s, off = orc.tx(pk)
        x = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 300), snr_db=35.0, cfo=0.2, lead=400, tail=5000, **kw)
        for force in ("0", "1"):
            import os
            os.environ["OFDMX_FORCE_GENERIC"] = force
            try:
                phy = cm.make_phy(cfg, max_pkt_bytes=plen + 4)
                # ---- TX: samples + offsets sized exactly
                need = sum(int(phy.frame_samples(len(b))) for b in pk)
                rs, vs, g, n = _guarded((need,), torch.complex64)
                ro, vo, g2, n2 = _guarded((len(pk) + 1,), torch.int64)
                flat = torch.from_numpy(np.frombuffer(b"".join(pk), np.uint8).copy()).to(_dev())
                po = torch.arange(len(pk) + 1, dtype=torch.int64, device=_dev()) * plen
                _lib.check(L.ofdmx_tx(phy.ctx, flat.data_ptr(), po.data_ptr(), len(pk), 0, vs.data_ptr(), need, vo.data_ptr(), st()), phy.ctx)
                torch.cuda.synchronize()
                assert _guards_intact(rs, g, n) and _guards_intact(ro, g2, n2), "TX wrote outside its buffers"
                assert cm.rel_evm(vs.cpu().numpy(), s) < 1e-5
                # ---- RX: records, slots, counts, z tap sized exactly to max_frames
                xd = _to_dev(x)
                mf = 16
                zs = phy.header_len() + (plen + 4) * 8 // phy.bps_payload + 1
                outs = []
                for rep in range(5):
                    rf, vf, gf, nf = _guarded((mf * 32,), torch.uint8)
                    rb, vb, gb, nb = _guarded((mf, phy.byte_stride), torch.uint8)
                    rz, vz, gz, nz = _guarded((mf, zs), torch.complex64)
                    rc, vc, gc, nc = _guarded((4,), torch.int32)
                    vz.zero_(); vb.zero_(); vf.zero_()
                    _lib.check(L.ofdmx_rx(phy.ctx, xd.data_ptr(), 1, xd.numel(), xd.numel(), vf.data_ptr(), mf, vb.data_ptr(),
                                          phy.byte_stride, vz.data_ptr(), zs, vc.data_ptr(), st()), phy.ctx)
                    torch.cuda.synchronize()
                    for r_, g_, n_ in ((rf, gf, nf), (rb, gb, nb), (rz, gz, nz), (rc, gc, nc)):
                        assert _guards_intact(r_, g_, n_), "RX wrote outside its buffers"
                    c = vc.cpu().numpy()
                    assert c[1] == len(pk) and not c[2]
                    outs.append((vf.cpu().numpy().tobytes(), vb.cpu().numpy().tobytes(), vz.cpu().numpy().tobytes()))
                assert all(o == outs[0] for o in outs[1:]), "RX results differ between identical runs"
                # ---- sync only
                rt, vt, gt, nt_ = _guarded((32,), torch.int64)
                rcf, vcf, gcf, ncf = _guarded((32,), torch.float32)
                rst, vst, gst, nst = _guarded((32,), torch.int32)
                cnt = torch.zeros(4, dtype=torch.int32, device=_dev())
                _lib.check(L.ofdmx_sync(phy.ctx, xd.data_ptr(), 1, xd.numel(), xd.numel(), vt.data_ptr(), vcf.data_ptr(), vst.data_ptr(),
                                        32, cnt.data_ptr(), st()), phy.ctx)
                torch.cuda.synchronize()
                assert _guards_intact(rt, gt, nt_) and _guards_intact(rcf, gcf, ncf) and _guards_intact(rst, gst, nst)
            finally:
                os.environ.pop("OFDMX_FORCE_GENERIC", None)
    # ---- conditioning kernels: odd sizes, output exactly sized
    phy = cm.make_phy(cm.cfg_c1())
    for n_streams, n in ((3, 1001), (37, 333), (1, 70001)):
        x = (rng.standard_normal((n_streams, n)) + 1j * rng.standard_normal((n_streams, n))).astype(np.complex64)
        xd = _to_dev(x)
        ry, vy, gy, ny = _guarded((n_streams, n), torch.complex64)
        gain = torch.ones(n_streams, dtype=torch.float32, device=_dev())
        phy.agc2(xd, gain=gain, out=vy)
        torch.cuda.synchronize()
        assert _guards_intact(ry, gy, ny), "agc2 wrote outside its buffer"
        ry, vy, gy, ny = _guarded((n_streams, n), torch.complex64)
        phy.iir_ccd(xd, [0.2, 0.3, 0.1], [1.0, 0.5, -0.2], out=vy)
        torch.cuda.synchronize()
        assert _guards_intact(ry, gy, ny), "iir_ccd wrote outside its buffer"


def test_reconfigure_at_a_call_boundary_without_draining():
    """ofdmx_reconfigure swaps the plan without cudaDeviceSynchronize / cudaMalloc: RX calls of plan A are enqueued,
    the plan is swapped while they are still in flight, RX calls of plan B are enqueued behind them on the same stream,
    and only then does the host synchronise.  Every call must have seen its own plan's tables (results equal fresh
    contexts / the oracle), the allocation and host-sync counters must not move across the swaps, and the call
    returns in well under a millisecond.  The carrier plans come from spectrum_translator -> spectrum_enforcer
    (python/cognitive_engine_mac.py:278-285)."""
    import time
    from ofdm_tools import _lib, OfdmPhy, ofdm_cr_tools as T
    rng = np.random.default_rng(91)
    # plans: spectrum constraints in Hz -> bins -> carrier plan + sync words
    plans = []
    for hz in ([], [2.40e9 + 120e3, 2.40e9 - 310e3], [2.40e9 + 5e3]):
        bins = T.spectrum_translator(hz, 2.40e9, 1.0e6, 128, 8)
        occ, pil, pls, sw1, sw2 = T.spectrum_enforcer(128, bins, 10)
        plans.append(dict(fft_len=128, cp_len=32, occupied_carriers=occ, pilot_carriers=pil, pilot_symbols=pls,
                          sync_word1=sw1, sync_word2=sw2, bps_header=1, bps_payload=2, scramble_bits=True,
                          scramble_header=True, crc_mode=1, max_carr_offset=3))
    assert len(plans[1]["occupied_carriers"][0]) < len(plans[0]["occupied_carriers"][0])
    streams, packets = [], []
    for pl in plans:
        orc = cm.make_oracle(pl)
        pk = cm.rand_packets(rng, 12, 150)
        s, off = orc.tx(pk)
        streams.append(_to_dev(cm.channel(cm.split_frames(s, off), rng, gaps=(300, 900), snr_db=35.0, cfo=0.2, fft_len=128,
                                          lead=400, tail=2500)))
        packets.append(pk)
    # a long filler stream keeps the device busy while the host swaps plans (so the swap really overlaps running work)
    big_pl = plans[0]
    o0 = cm.make_oracle(big_pl)
    s, off = o0.tx(cm.rand_packets(rng, 40, 150))
    filler = _to_dev(np.tile(cm.channel(cm.split_frames(s, off), rng, gaps=(0, 0), snr_db=35.0, fft_len=128, lead=400, tail=2500), 200))
    phy = OfdmPhy(max_pkt_bytes=160, **plans[0])
    bufs = [phy.rx_buffers(64, _dev()) for _ in range(8)]
    fb = phy.rx_buffers(filler.numel() // 400, _dev())
    phy.rx_enqueue(filler, fb)                      # sizes the workspace for the largest call
    phy.rx_enqueue(streams[0], bufs[0])
    torch.cuda.synchronize()
    phy.reconfigure(**{k: plans[1][k] for k in ("occupied_carriers", "pilot_carriers", "pilot_symbols", "sync_word1", "sync_word2")})
    phy.reconfigure(**{k: plans[0][k] for k in ("occupied_carriers", "pilot_carriers", "pilot_symbols", "sync_word1", "sync_word2")})
    torch.cuda.synchronize()                        # both halves of the table arena exist now
    a0, h0 = phy.counter(_lib.CNT_DEVICE_ALLOCS), phy.counter(_lib.CNT_HOST_SYNCS)
    order, lat = [0, 1, 2, 1, 0, 2, 0], []
    cur = 0
    for i, pi in enumerate(order):
        if pi != cur:
            kw = {k: plans[pi][k] for k in ("occupied_carriers", "pilot_carriers", "pilot_symbols", "sync_word1", "sync_word2")}
            fresh = OfdmPhy(max_pkt_bytes=160, **dict(plans[pi]))       # host-side parameter block only
            t0 = time.perf_counter()
            _lib.check(_lib.load().ofdmx_reconfigure(__import__("ctypes").byref(phy._ctx), __import__("ctypes").byref(fresh.params)), phy._ctx)
            lat.append(time.perf_counter() - t0)
            cur = pi
        phy.rx_enqueue(filler, fb)                  # ~ms of device work in front of ...
        phy.rx_enqueue(streams[pi], bufs[i])        # ... the call whose result is checked
    assert phy.counter(_lib.CNT_DEVICE_ALLOCS) == a0, "a reconfiguration or a steady-state call allocated"
    assert phy.counter(_lib.CNT_HOST_SYNCS) == h0, "a reconfiguration or a steady-state call blocked the host"
    assert phy.counter(_lib.CNT_RECONFIGS) == 2 + len(lat)
    torch.cuda.synchronize()
    for i, pi in enumerate(order):
        c = bufs[i]["counts"].cpu().numpy()
        rec = np.frombuffer(bufs[i]["frames"][: int(c[1]) * 32].cpu().numpy().tobytes(), phy_frame_dtype())
        ref = cm.make_oracle(plans[pi]).rx(streams[pi].cpu().numpy(), byte_stride=160, want_z=False)
        assert np.array_equal(rec["trigger"], ref["frames"]["trigger"]), "call %d ran on the wrong plan's tables" % i
        sl = bufs[i]["slots"].cpu().numpy()
        got = [bytes(sl[int(f["slot"]), : int(f["pkt_len"]) - 4]) for f in rec if f["flags"] & 2]
        assert got == packets[pi]
    assert max(lat) < 2e-3, "ofdmx_reconfigure took %.0f us" % (max(lat) * 1e6)
    print("reconfigure call latency (host): %s us" % [round(v * 1e6) for v in lat])


def phy_frame_dtype():
    from ofdm_tools.phy import FRAME_DTYPE
    return FRAME_DTYPE


@pytest.mark.parametrize("which,cfo", [("c1", 0.2), ("c3", 0.3), ("c3", 2.2), ("c3", -1.8), ("radio", 0.1)])
def test_single_sync_word_mode(which, cfo):
    """sync_word2=(): TX and RX through the library against the oracle (frame sizes, samples, triggers, carrier
    offsets, header fields, bytes, pre-decision symbols), with fractional and integer carrier offsets.  (Carrier
    plans end on carriers sync word 1 covers, and fft_len 64 / 128 pin max_carr_offset to 0: see
    tests/test_host.py::test_single_sync_word_mode_oracle_roundtrip.)"""
    rng = np.random.default_rng(55)
    if which == "c1":
        occ = [[k for k in cm.OCC64[0] if abs(k) != 26]]
        cfg, plen, kw = cm.cfg_c1(2, True, 1, sync_word2=(), occupied_carriers=occ, max_carr_offset=0), 96, {}
    elif which == "c3":
        occ = [[k for k in cm.cfg_c3()["occupied_carriers"][0] if abs(k) != 302]]
        cfg, plen, kw = cm.cfg_c3(sync_word2=(), occupied_carriers=occ), 1500, dict(fft_len=1024, taps=cm.MULTIPATH)
    else:
        base = cm.cfg_radio128(4)
        occ = [[k for k in base["occupied_carriers"][0] if k != -54]]
        cfg, plen, kw = dict(base, sync_word2=(), occupied_carriers=occ, max_carr_offset=0), 200, dict(fft_len=128)
    orc, phy = cm.make_oracle(cfg), cm.make_phy(cfg)
    assert phy.n_sync_words == 1 and phy.frame_samples(plen) == orc.frame_samples(plen)
    pk = cm.rand_packets(rng, 5, plen)
    s_ref, off_ref = orc.tx(pk)
    s_gpu, off_gpu = phy.tx(pk)
    assert np.array_equal(off_gpu.cpu().numpy(), off_ref)
    assert cm.rel_evm(s_gpu.cpu().numpy(), s_ref) < 1e-5
    stream = cm.channel(cm.split_frames(s_ref, off_ref), rng, gaps=(0, 700), snr_db=35.0, cfo=cfo, lead=400, tail=4000, **kw)
    res, ref = _compare_rx(cfg, stream)
    assert res.payloads() == pk
    assert np.all(res.frames["carr_offset"] == int(round(cfo / 2.0)) * 2)
    # the segment stitcher and the stream adaptor know about the shorter preamble
    from ofdm_tools import dist
    recs, pay = dist.rx_segmented(cm.make_phy(cfg), _to_dev(stream), 3, max_pkt_bytes=plen + 4)
    assert pay == pk and np.array_equal(recs["trigger"], res.frames["trigger"])


@pytest.mark.parametrize("n_streams,n,kind", [(1, 3000000, "ofdm"), (3, 1500000, "levels"), (1, 2000000, "gaps"), (2, 1200000, "weak")])
def test_agc2_time_parallel_spans_are_bit_exact(n_streams, n, kind):
    """ofdmx_agc2 on few long streams: spans run in parallel from a warm-up and are accepted only when their entry gain
    equals the predecessor's exit gain bit for bit, failing spans are re-run -- the output and the final gains must
    equal the sequential recurrence (the oracle) exactly, whatever the signal does: steady OFDM-like input, level
    jumps, zero gaps (no contraction while the gain ramps), and a weak signal whose loop barely contracts (the
    warm-ups do not converge: everything is repaired / walked sequentially, still exact)."""
    import time
    import oracle as O
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((n_streams, n)) + 1j * rng.standard_normal((n_streams, n))).astype(np.complex64) * 0.2
    if kind == "levels":
        for i, a in enumerate((1.0, 30.0, 0.03, 4.0, 0.5, 200.0)):
            x[:, i * n // 6:(i + 1) * n // 6] *= a
    elif kind == "gaps":
        x[:, n // 5: n // 5 + 300000] = 0
        x[:, n // 2: n // 2 + 5000] = 0
        x[:, 3 * n // 4:] *= 50.0
    elif kind == "weak":
        x *= 1e-3
    phy = cm.make_phy(cm.cfg_c1())
    xd = _to_dev(x)
    ref, gref = O.agc2(x)
    y, g = phy.agc2(xd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y, g = phy.agc2(xd)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert np.array_equal(y.cpu().numpy(), ref), "span-parallel agc2 differs from the sequential recurrence"
    assert np.array_equal(g.cpu().numpy(), np.atleast_1d(gref).astype(np.float32))
    # continuing the stream in a second call (gain carried) stays exact, also with the fabsf rate rule
    k = n // 2 + 77
    ya, ga = phy.agc2(xd[:, :k].contiguous(), abs_rate=True)
    yb, ga = phy.agc2(xd[:, k:].contiguous(), gain=ga, abs_rate=True)
    r2, g2 = O.agc2(x, abs_rate=True)
    assert np.array_equal(torch.cat([ya, yb], 1).cpu().numpy(), r2) and np.array_equal(ga.cpu().numpy(), np.atleast_1d(g2).astype(np.float32))
    print("agc2 %s: %d x %d samples in %.2f ms = %.2f Gsamples/s" % (kind, n_streams, n, dt * 1e3, n_streams * n / dt / 1e9))


@pytest.mark.parametrize("which", ["c1", "c3", "radio"])
def test_channel_estimate_debug_tap(which):
    """ofdmx_set_debug_taps: the taps ofdm_chanest_vcvc hands to the equaliser (ofdm_sync_chan_taps) against the
    oracle's, from the warp-per-frame kernel (occupied carriers) and from the any-fft_len kernel (every bin of the
    reference tag), on a multipath channel with an integer carrier offset."""
    import os
    rng = np.random.default_rng(61)
    cfg, plen, kw = {"c1": (cm.cfg_c1(2, True, 1), 96, {}), "c3": (cm.cfg_c3(), 1500, dict(fft_len=1024)),
                     "radio": (cm.cfg_radio128(4), 200, dict(fft_len=128))}[which]
    orc = cm.make_oracle(cfg)
    pk, fr = _frames(cfg, rng, 4, plen)
    x = cm.channel(fr, rng, gaps=(0, 500), snr_db=35.0, cfo=2.2 if which != "c1" else 0.2, taps=cm.MULTIPATH, lead=400, tail=4000, **kw)
    ref = orc.rx(x, want_z=True, want_taps=True, byte_stride=4096)
    assert len(ref["frames"]) == 4
    occ = [(c + cfg["fft_len"]) % cfg["fft_len"] for c in cfg["occupied_carriers"][0]]
    occ_sh = [(b + cfg["fft_len"] // 2) % cfg["fft_len"] for b in occ]
    for force in ("0", "1"):
        os.environ["OFDMX_FORCE_GENERIC"] = force
        try:
            phy = cm.make_phy(cfg, max_pkt_bytes=plen + 4)
            taps = torch.zeros((16, cfg["fft_len"]), dtype=torch.complex64, device=_dev())
            phy.set_debug_taps(taps)
            res = phy.rx(_to_dev(x), max_frames=16, want_z=True)
            phy.set_debug_taps(None)
        finally:
            os.environ.pop("OFDMX_FORCE_GENERIC", None)
        assert np.array_equal(res.frames["trigger"], ref["frames"]["trigger"])
        th = taps.cpu().numpy()
        for i, f in enumerate(res.frames):
            got, want = th[int(f["slot"])], ref["taps"][i]
            # the taps carry the frame's NCO phase reference (the oracle accumulates the oscillator phase over the
            # whole stream, the kernels restart it at every trigger; it cancels in y / H): compare up to that phasor
            ph = np.vdot(got[occ_sh], want[occ_sh])
            got = got * (ph / abs(ph))
            assert cm.rel_evm(got[occ_sh], want[occ_sh]) < 1e-4
            if force == "1":
                assert cm.rel_evm(got, want) < 1e-4          # pilots and every other bin of the tag as well
        assert np.abs(ref["taps"][0][occ_sh]).min() > 0


@pytest.mark.parametrize("which,roll", [("c3", 18), ("c3", 72), ("c1", 4), ("radio", 8), ("c3", 2)])
def test_rolloff_on_the_warp_tx_kernel(monkeypatch, which, roll):
    """ofdm_cyclic_prefixer with rolloff_len > 0 on the warp-per-packet TX kernel (sync_transmit_path uses cp_len/4,
    python/ofdm_cr_tools.py:1093): offsets exact and samples within 1e-5 of the peak against the oracle and against the
    CTA-per-packet kernel, with scaling and clipper behind it, ragged packet lengths; the bursts decode."""
    rng = np.random.default_rng(roll)
    cfg, plen = {"c3": (cm.cfg_c3(), 1500), "c1": (cm.cfg_c1(2, True, 1), 96), "radio": (cm.cfg_radio128(4), 200)}[which]
    kw = dict(cfg, rolloff=roll, tx_scale=0.01, tx_clip=0.6 if which == "c3" else 0.25)   # (clips the 16-QAM peaks mildly)
    orc = cm.make_oracle(kw)
    pk = cm.rand_packets(rng, 6, plen) + [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (1, plen // 3, plen - 1)]
    so, oo = orc.tx(pk)
    outs = {}
    for name, env in (("warp", "0"), ("cta", "1")):
        monkeypatch.setenv("OFDMX_NO_WARP_TX", env)
        phy = cm.make_phy(kw)
        phy.profile(True)
        s, off = phy.tx(pk)
        used = set(phy.profile_read())
        assert ("tx_framew_kernel" in used) == (name == "warp"), used
        assert np.array_equal(off.cpu().numpy(), oo)
        outs[name] = s.cpu().numpy()
        assert outs[name].shape == so.shape and np.abs(outs[name] - so).max() <= 1e-5 * np.abs(so).max(), name
    assert np.abs(outs["warp"] - outs["cta"]).max() <= 2e-6 * np.abs(so).max()
    monkeypatch.setenv("OFDMX_NO_WARP_TX", "0")
    x = cm.channel(cm.split_frames(outs["warp"], oo), rng, gaps=(300, 700), tail=4000, snr_db=40.0, fft_len=cfg["fft_len"], scale=100.0)
    got = cm.make_phy(kw).rx(_to_dev(x)).payloads()
    assert got == pk


@pytest.mark.gpu
@pytest.mark.parametrize("fft_len,n_streams,n", [(32, 1, 70001), (64, 2, 90000), (128, 1, 50000), (256, 3, 60002), (512, 1, 150000),
                                                 (512, 2, 700), (64, 1, 511), (256, 1, 16 * 8 * 512 + 5), (2048, 1, 300001), (2048, 2, 70000),
                                                 (2048, 1, 1023), (1024, 2, 50002)])
def test_short_window_warp_sync_kernel(monkeypatch, fft_len, n_streams, n):
    """fft_len 32 .. 512: the short-window warp-autonomous Schmidl & Cox kernel (default), fft_len 1024 / 2048: the
    chunk-size template of the long-window one; each vs the TMA ring kernel vs the oracle: identical triggers and CFO.  Streams with frames at both ends, a 40 dB louder burst directly in front of a
    quiet frame (every partial window sum must stay local), stretches of exact zeros, odd lengths, spans that end
    inside the last tile."""
    rng = np.random.default_rng(fft_len + n)
    cfg = _plan(fft_len, {32: 20, 64: 48, 128: 96, 256: 200, 512: 400, 1024: 600, 2048: 1200}[fft_len], 1)
    orc = cm.make_oracle(cfg)
    pk = cm.rand_packets(rng, 2, 24)
    s_ref, off_ref = orc.tx(pk)
    fr = cm.split_frames(s_ref, off_ref)
    xs = []
    for s in range(n_streams):
        x = 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        L = len(fr[0])
        for o in (0, n // 3 + 7 * s, 2 * n // 3 + 1, n - L // 2):
            if 0 <= o < n:
                m = min(L, n - o)
                x[o:o + m] += fr[s % 2][:m]
        o = n // 2
        if o + 3 * L < 2 * n // 3:
            x[o:o + L] += 100.0 * fr[0][:L]                  # loud burst ...
            x[o + L:o + 2 * L] += fr[1][:L]                  # ... a normal frame right behind it
            x[o + 2 * L + 40:o + 2 * L + 40 + 3 * fft_len] = 0.0   # exact zeros: R == 0 must not detect
        xs.append(x.astype(np.complex64))
    x = np.stack(xs)
    res = {}
    for name, env in (("warp", {}), ("ring", {"OFDMX_NO_WARP_SYNC": "1"})):
        monkeypatch.setenv("OFDMX_NO_WARP_SYNC", env.get("OFDMX_NO_WARP_SYNC", "0"))
        phy = cm.make_phy(cfg)
        res[name] = phy.sync(_to_dev(x if n_streams > 1 else x[0]))
        if n >= fft_len:
            from test_gpu_parity import _kernels_used
            used = _kernels_used(phy, lambda: phy.sync(_to_dev(x if n_streams > 1 else x[0])))
            kname = "sync_metric_warpn_kernel" if fft_len <= 512 else "sync_metric_warp_kernel"
            assert (kname in used) == (name == "warp"), used
    ref_t, ref_s, ref_c = [], [], []
    for s in range(n_streams):
        t, c = orc.sync(x[s])
        ref_t += list(t); ref_c += list(c); ref_s += [s] * len(t)
    assert len(ref_t) >= (2 if n > 20 * fft_len else 0)
    for name in ("warp", "ring"):
        trig, cfo, st = res[name]
        assert np.array_equal(trig, np.array(ref_t, np.int64)), name
        assert np.array_equal(st, np.array(ref_s)), name
        np.testing.assert_allclose(cfo, np.array(ref_c, np.float32), atol=2e-6, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("bps,snr,cfo", [(6, 28.0, 0.2), (4, 22.0, -0.35), (2, 15.0, 0.45)])
def test_pair_kernel_next_trigger_inside_the_last_symbol(bps, snr, cfo):
    """fft_len 2048, back-to-back frames: the next frame's trigger often arrives a few samples early and so falls inside
    this frame's last symbol, where the sample-and-hold NCO frequency changes.  The pair-of-warps kernel patches the
    few affected samples after its uniform-frequency path (and takes the per-sample path when the trigger lies
    deeper in the symbol): records, bytes and equalised symbols against the oracle, and the case must occur."""
    cfg = cm.cfg_c4(bps_payload=bps)
    rng = np.random.default_rng(77 + bps)
    sym_bytes = 1200 * bps // 8
    pk = [rng.integers(0, 256, 2 * sym_bytes - 4, dtype=np.uint8).tobytes() for _ in range(40)]   # whole symbols
    s, off = cm.make_oracle(cfg).tx(pk)
    fr = cm.split_frames(s, off)
    # a shortened frame now and then puts the next trigger deep inside the symbol (the per-sample path)
    fr = [f if i % 9 != 4 else f[:len(f) - 300] for i, f in enumerate(fr)]
    stream = cm.channel(fr, rng, gaps=(0, 0), lead=700, tail=6000, snr_db=snr, cfo=cfo, fft_len=2048)
    res, ref = _compare_rx(cfg, stream)
    t = ref["triggers"]
    D = 2048 + cfg["cp_len"]
    early = 0
    for i, f in enumerate(ref["frames"]):
        nx = t[np.searchsorted(t, f["trigger"], side="right"):][:1]
        last_end = int(f["trigger"]) + (3 + int(f["frame_syms"]) - 1) * D + cfg["cp_len"] + 2047
        early += int(len(nx) and nx[0] <= last_end)
    assert early >= 5, early
    from test_gpu_parity import _kernels_used
    phy = cm.make_phy(cfg)
    assert "rx_framep_kernel" in _kernels_used(phy, lambda: phy.rx(_to_dev(stream)))
    assert len(res.frames) >= 30

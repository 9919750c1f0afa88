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

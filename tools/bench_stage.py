#!/usr/bin/env python3
"""Stage-level measurement for the SURVEY.md 8(f) rows built after the main path (bench.py keeps the
headline RX metric).  Prints one JSON line per stage.

  python tools/bench_stage.py --stage agc2 [--streams 4096] [--samples 56320] [--steps 10]
  python tools/bench_stage.py --stage tx [--frames 65536] [--steps 10]
  python tools/bench_stage.py --stage rx_small [--steps 5]

tx: the TX chain (CRC-32 append, header, scrambler, mapping, carrier allocation, IFFT, cyclic prefix, x0.01) on
BASELINE config[2] packets (fft_len 1024, 16-QAM, 1500 bytes) resident in HBM; value = Msamples/s produced,
roofline = (payload bytes read + 8 B per sample written) / time against the measured HBM peak, cpu_baseline = the
oracle TX on all host cores over a bounded sample, parity gate against the oracle on the first packets.

rx_small: the full RX chain on the other configurations of BASELINE.json (multi-stream, fft_len 64 / 128 / 2048): configs[1] (fft_len 64, 802.11a carrier
plan of ofdm_tx_rx_hier, QPSK, 96-byte packets, 4096 independent streams x 64 frames) and the ofdm_radio_hier
default plan (fft_len 128, 103 data carriers, 16-QAM, 350-byte packets + CRC-32, 1024 streams x 64 frames); one JSON
line per configuration, parity gate = two streams against the oracle (records and bytes bit-exact).

agc2: analog.agc2_cc over `streams` independent streams of `samples` complex samples resident in HBM
(BASELINE config[1] shape: 4096 streams x 64 frames x 880 samples).  value = Msamples/s (CUDA events on the
launching stream), roofline = (8 B read + 8 B written per sample) / kernel time against the measured HBM
peak, cpu_baseline = the oracle on one host core over a bounded sample of the same streams.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "gr-ofdm_tools_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402


def stage_tx(args, torch, cm, dev, peak):
    import oracle as O
    plen = {"c3": 1500, "c1": 96, "radio128": 350}[args.plan]
    cfg = {"c3": cm.cfg_c3, "c1": lambda: cm.cfg_c1(2, False, 0), "radio128": lambda: cm.cfg_radio128(4, 1, 1)}[args.plan]()
    cfg["tx_scale"] = 0.01
    if args.rolloff:
        cfg["rolloff"] = args.rolloff
    phy = cm.make_phy(cfg, max_pkt_bytes=plen + 4)
    rng = np.random.default_rng(3)
    nf = args.frames
    payload = torch.from_numpy(rng.integers(0, 256, nf * plen, dtype=np.uint8)).to(dev)
    off = torch.arange(nf + 1, dtype=torch.int64, device=dev) * plen
    out = torch.empty(nf * int(phy.frame_samples(plen)), dtype=torch.complex64, device=dev)
    soff = torch.zeros(nf + 1, dtype=torch.int64, device=dev)
    for _ in range(args.warmup):
        s, soff = phy.tx((payload, off), out=out, soff=soff)
    torch.cuda.synchronize()
    phy.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s, soff = phy.tx((payload, off), out=out, soff=soff)     # enqueue only: pre-allocated output
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = phy.profile_read()
    n = int(soff[-1])
    # parity gate on the first packets + CPU baseline on a bounded sample
    k = min(256, nf)
    pk = [bytes(payload[i * plen:(i + 1) * plen].cpu().numpy()) for i in range(k)]
    orc = O.Oracle(**cfg)
    t0 = time.perf_counter()
    ref, roff = orc.tx(pk)
    cpu_s = time.perf_counter() - t0
    got = s[: int(roff[-1])].cpu().numpy()
    err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    assert err < 1e-5, "TX differs from the oracle: %g" % err
    kern_ms = {kk: v[0] / v[1] * (v[1] / args.steps) for kk, v in prof.items()}
    dom = max(kern_ms, key=kern_ms.get)
    ach = (nf * float(plen) + 8.0 * n) / (kern_ms[dom] * 1e-3) / 1e9
    print(json.dumps({
        "stage": "tx", "metric": "OFDM TX Msamples/s (fft_len=%d)" % phy.fft_len, "value": n / (ms * 1e-3) / 1e6,
        "unit": "Msamples/s", "ms_per_step": ms, "steps": args.steps,
        "config": {"workload": "plan %s: %d packets of %d bytes -> %d samples, resident in HBM, output pre-allocated (enqueue-only calls)" % (args.plan, nf, plen, n)},
        "kernels_ms_per_step": kern_ms,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak},
        "parity": {"rel_l2_vs_oracle": err, "packets": k},
        "cpu_baseline": {"value": int(roff[-1]) / cpu_s / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "%d packets" % k},
    }))


def stage_rx_small(args, torch, cm, dev, peak):
    import oracle as O
    cases = (("configs[1]: fft_len 64, QPSK, 96 B, 4096 streams x 64 frames", cm.cfg_c1(2, False, 0), 96, 4096, 64),
             ("ofdm_radio_hier defaults: fft_len 128, 16-QAM, 350 B + CRC-32, 1024 streams x 64 frames", cm.cfg_radio128(4, 1, 1), 350, 1024, 64),
             ("configs[3], one GPU's share: fft_len 2048, 64-QAM, 1500 B + CRC-32, 128 streams x 256 frames", cm.cfg_c4(), 1500, 128, 256))
    steps = min(args.steps, 5)
    for name, cfg, plen, n_streams, n_frames in cases:
        phy = cm.make_phy(cfg)
        rng = np.random.default_rng(1)
        n = n_streams * n_frames
        payload = torch.from_numpy(rng.integers(0, 256, n * plen, dtype=np.uint8)).to(dev)
        off = torch.arange(n + 1, dtype=torch.int64, device=dev) * plen
        s, soff = phy.tx((payload, off))
        fs = int(soff[1])
        pad = torch.zeros(n_streams, 4 * phy.fft_len, dtype=torch.complex64, device=dev)
        x = torch.cat([pad, s.view(n_streams, n_frames * fs), pad], 1).contiguous()
        g = torch.Generator(device=dev).manual_seed(2)
        x += 0.001 * torch.view_as_complex(torch.randn(*x.shape, 2, device=dev, generator=g))
        bufs = phy.rx_buffers(n + 64, dev)
        for _ in range(args.warmup):
            phy.rx_enqueue(x, bufs)
        torch.cuda.synchronize()
        phy.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            phy.rx_enqueue(x, bufs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = phy.rx_collect(bufs)
        prof = {k: v[0] / steps for k, v in phy.profile_read().items()}
        phy.profile(False)
        assert len(res.frames) == n, "decoded %d of %d frames" % (len(res.frames), n)
        # parity gate + CPU baseline: two streams through the oracle
        orc = O.Oracle(**cfg)
        xs = x[:2].cpu().numpy()
        t0 = time.perf_counter()
        refs = [orc.rx(xs[i], byte_stride=phy.byte_stride, want_z=False) for i in range(2)]
        cpu_s = time.perf_counter() - t0
        r2 = phy.rx(x[:2].contiguous(), want_z=False)
        sl = r2.slots.cpu().numpy()
        k = 0
        for i in range(2):
            f = refs[i]["frames"]
            gsel = r2.frames[r2.frames["stream"] == i]
            assert np.array_equal(gsel["trigger"], f["trigger"]) and np.array_equal(gsel["flags"] & 7, f["flags"] & 7)
            for q in range(len(f)):
                nb = int(f["pkt_len"][q])
                assert np.array_equal(sl[int(gsel["slot"][q]), :nb], refs[i]["bytes"][q, :nb])
                k += 1
        dom = max(prof, key=prof.get)
        algo = 8.0 * x.numel() + float(res.frames["pkt_len"].astype(np.int64).sum()) + 32.0 * n
        print(json.dumps({
            "stage": "rx_small", "metric": "OFDM RX Msamples/s", "value": x.numel() / (ms * 1e-3) / 1e6, "unit": "Msamples/s",
            "ms_per_step": ms, "steps": steps, "config": {"workload": name + ", resident in HBM"},
            "kernels_ms_per_step": prof,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / peak, "note": "whole chain: algorithmic bytes of the call / step time",
                         "dominant_kernel": dom},
            "parity": {"frames_checked_bit_exact": k},
            "cpu_baseline": {"value": xs.size / cpu_s / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port (exact float64 sync, not the FIR port)",
                             "sample": "2 streams"},
        }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="agc2", choices=["agc2", "tx", "rx_small"])
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--rolloff", type=int, default=0, help="tx stage: ofdm_cyclic_prefixer rolloff_len")
    ap.add_argument("--plan", default="c3", choices=["c3", "c1", "radio128"], help="tx stage: carrier plan (c3 = BASELINE config[2])")
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=64 * 880)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import common as cm
    dev = torch.device("cuda", 0)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    if args.stage == "tx":
        return stage_tx(args, torch, cm, dev, peak)
    if args.stage == "rx_small":
        return stage_rx_small(args, torch, cm, dev, peak)
    phy = cm.make_phy(cm.cfg_c1())
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.view_as_complex(torch.randn(args.streams, args.samples, 2, device=dev, generator=g))
    x *= (10.0 ** torch.empty(args.streams, 1, device=dev).uniform_(-3, 2, generator=g)).to(torch.complex64)
    out = torch.empty_like(x)
    for _ in range(args.warmup):
        phy.agc2(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        phy.agc2(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    n = args.streams * args.samples
    # parity gate + CPU baseline on a bounded sample
    import oracle as O
    ns = min(8, args.streams)
    xs = x[:ns].cpu().numpy()
    t0 = time.perf_counter()
    ref, gref = O.agc2(xs)
    cpu_s = time.perf_counter() - t0
    y, gg = phy.agc2(x[:ns].contiguous())
    assert np.array_equal(y.cpu().numpy().view(np.float32), ref.view(np.float32)), "agc2 differs from the oracle"
    ach = 16.0 * n / (ms * 1e-3) / 1e9
    print(json.dumps({
        "stage": "agc2", "metric": "AGC Msamples/s (analog.agc2_cc, %d streams)" % args.streams,
        "value": n / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms, "steps": args.steps,
        "config": {"workload": "%d streams x %d samples, levels 1e-3..1e2, resident in HBM (larger than L2)" % (args.streams, args.samples)},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "note": "16 algorithmic bytes per sample; the kernel is bound by the per-stream recurrence latency, not by bandwidth"},
        "cpu_baseline": {"value": ns * args.samples / cpu_s / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                         "sample": "%d streams x %d samples" % (ns, args.samples)},
    }))


if __name__ == "__main__":
    main()

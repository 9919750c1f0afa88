#!/usr/bin/env python3
"""Stage-level measurement for the SURVEY.md 8(f) rows built after the main path (bench.py keeps the
headline RX metric).  Prints one JSON line per stage.

  python tools/bench_stage.py --stage agc2 [--streams 4096] [--samples 56320] [--steps 10]

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="agc2", choices=["agc2"])
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=64 * 880)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import common as cm
    dev = torch.device("cuda", 0)
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
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
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

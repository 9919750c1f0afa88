#!/usr/bin/env python3
"""bench.py -- OFDM RX Msamples/s on BASELINE.json config[2] (the configuration the metric is quoted
on): fft_len 1024, cp 72, 600 data carriers, 16-QAM, 1500-byte packets (+CRC-32, scrambler), CFO 0.3
subcarriers + 4-tap multipath + AWGN, back-to-back frames, one stream per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the GNU Radio chain

A step = one pass of the full RX chain (sync -> ... -> CRC) over the rank's resident sample buffer
(inputs far larger than L2).  `value` = samples of all ranks / max-over-ranks device time.  `e2e` = the
same through ofdmx_rx_host with pinned HOST buffers (H2D + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "gr-ofdm_tools_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "OFDM RX Msamples/s (fft_len=1024, 16-QAM)"
FRAME_SAMPLES = 9864            # (2 sync + 1 header + 6 payload symbols) x 1096
FRAME_ALGO_BYTES = 8 * FRAME_SAMPLES + 1500 + 32   # samples once + payload + record
SNR_DB = 40.0
CFO = 0.3
WORKLOAD = ("config[2]: fft_len=1024 cp=72 600 carriers 16-QAM 1500B+CRC32 scrambled, back-to-back frames, "
            "CFO 0.3 + 4-tap multipath + AWGN %.0f dB, 1 stream/GPU" % SNR_DB)


def phy_cfg():
    import common as cm
    return cm.cfg_c3()


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region.  The region lasts tens of milliseconds,
    so the samples come from NVML directly (one query per ~millisecond from a thread); `nvidia-smi -lms` is
    the fallback when the NVML binding is missing (it may return no sample for so short a region)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
            ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.h, self.sm, self.mask, self.run = None, None, [], 0, False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nv, self.h = pynvml, h
        except Exception:
            self.nv = None

    def _poll(self):
        nv, h = self.nv, self.h
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(reasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nv is not None:
            self.run = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.run = False
            self.th.join(timeout=1.0)
            try:
                mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx,
                    "reasons": [n for n, b in self.BITS if self.mask & b], "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def make_stream(phy, n_frames, seed, dev):
    """Synthetic input resident in HBM: GPU TX of random packets, then the channel (torch is only the
    data generator here)."""
    import torch
    import common as cm
    rng = np.random.default_rng(seed)
    payload = torch.from_numpy(rng.integers(0, 256, n_frames * 1500, dtype=np.uint8)).to(dev)
    off = torch.arange(n_frames + 1, dtype=torch.int64, device=dev) * 1500
    s, soff = phy.tx((payload, off))
    assert int(soff[-1]) == n_frames * FRAME_SAMPLES
    lead, tail = 512, 4096
    x = torch.zeros(lead + s.numel() + tail, dtype=torch.complex64, device=dev)
    for d, v in cm.MULTIPATH:                      # 4-tap multipath
        x[lead + d: lead + d + s.numel()] += s * complex(v)
    del s
    chunk = 1 << 24
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    pw = 0.0
    for a in range(0, x.numel(), chunk):           # CFO + AWGN in chunks (bounded temporaries)
        b = min(a + chunk, x.numel())
        t = torch.arange(a, b, device=dev, dtype=torch.float64)
        x[a:b] *= torch.polar(torch.ones_like(t), 2 * np.pi * CFO / 1024 * t).to(torch.complex64)
        if a == 0:
            pw = float((x[lead:b].abs() ** 2).mean())
        sig = float(np.sqrt(pw / 10 ** (SNR_DB / 10) / 2))
        x[a:b] += torch.view_as_complex(torch.randn(b - a, 2, device=dev, generator=g) * sig)
    return x, payload


def run_reference(args):
    """CPU arm: the oracle port of the reference's GNU Radio chain (float32 FIR sync as GNU Radio
    evaluates it + the demod chain), all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import common as cm
    cfg = phy_cfg()
    orc = cm.make_oracle(cfg)
    rng = np.random.default_rng(1)
    n_frames = args.ref_frames
    pk = cm.rand_packets(rng, n_frames, 1500)
    s, off = orc.tx(pk)
    x = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 0), lead=512, tail=4096, snr_db=SNR_DB, cfo=CFO,
                   fft_len=1024, taps=cm.MULTIPATH)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        orc.rx_baseline(x, byte_stride=1520)
    t0 = time.perf_counter()
    nf = 0
    for _ in range(args.steps):
        nf += len(orc.rx_baseline(x, byte_stride=1520))
    dt = time.perf_counter() - t0
    val = len(x) * args.steps / dt / 1e6
    sample = "%d back-to-back frames (%d samples) per step" % (n_frames, len(x))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "what": "CPU restatement of the GNU Radio chain (oracle port; GNU Radio itself is not installable here)"},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_decoded_per_step": nf // max(1, args.steps),
    }))


def cpu_baseline(ref_frames):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import common as cm
    orc = cm.make_oracle(phy_cfg())
    rng = np.random.default_rng(1)
    pk = cm.rand_packets(rng, ref_frames, 1500)
    s, off = orc.tx(pk)
    x = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 0), lead=512, tail=4096, snr_db=SNR_DB, cfo=CFO,
                   fft_len=1024, taps=cm.MULTIPATH)
    orc.rx_baseline(x[: 20 * FRAME_SAMPLES], byte_stride=1520)     # warm
    t0 = time.perf_counter()
    r = orc.rx_baseline(x, byte_stride=1520)
    dt = time.perf_counter() - t0
    return {"value": len(x) / dt / 1e6, "unit": "Msamples/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d back-to-back frames (%d samples), %d decoded, %.1f s" % (ref_frames, len(x), len(r), dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=65536, help="frames per GPU (65536 = 646 M samples = 5.2 GB)")
    ap.add_argument("--ref-frames", type=int, default=1024, help="frames in the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ofdm_tools import OfdmPhy
    from ofdm_tools import dist as odist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    phy = OfdmPhy(device=local, tx_scale=0.01, max_pkt_bytes=1504, **phy_cfg())
    x, payload = make_stream(phy, args.frames, seed=1000 + rank, dev=dev)
    n = x.numel()
    max_frames = args.frames + 64
    torch.cuda.synchronize()

    bufs = phy.rx_buffers(max_frames, dev)
    gathered = torch.zeros((world, 4), dtype=torch.int32, device=dev)

    def enqueue():
        """One step, launch only: the RX chain, then the per-rank frame counters all-gathered over
        NCCL (the only collective on this path).  No host synchronisation inside the timed region."""
        phy.rx_enqueue(x, bufs)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), bufs["counts"])

    def step():
        enqueue()
        res = phy.rx_collect(bufs)
        summ = odist.summarize(res, n)
        if world > 1:
            summ = odist.gather_stats(summ, dev)
        return res, summ

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res, summ = step()
    # correctness gate: every frame decoded, CRC ok, payload identical to what was sent
    assert len(res.frames) == args.frames, "decoded %d of %d frames" % (len(res.frames), args.frames)
    assert bool(np.all(res.frames["flags"] & 2)), "CRC failures in the bench stream"
    slots = res.slots[res.frames["slot"][:64].astype(np.int64)].cpu().numpy()[:, :1500]
    assert np.array_equal(slots.reshape(-1), payload[: 64 * 1500].cpu().numpy()), "payload mismatch"

    sampler = ClockSampler(local)
    launches0 = phy.launch_count()
    phy.profile(True)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        enqueue()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    res = phy.rx_collect(bufs)
    summ = odist.summarize(res, n)
    if world > 1:
        summ = odist.gather_stats(summ, dev)
    assert len(res.frames) == args.frames and bool(np.all(res.frames["flags"] & 2))
    prof = phy.profile_read()
    phy.profile(False)
    launches = phy.launch_count() - launches0
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(t[1])
    total_samples = n * world * args.steps
    value = total_samples / (ms * 1e-3) / 1e6

    # ---- e2e: host buffers through ofdmx_rx_host (pinned input), H2D + D2H inside the timed region
    xh = torch.empty(n, dtype=torch.complex64, pin_memory=True)
    xh.copy_(x)
    xh_np = xh.numpy()
    r = phy.rx_host(xh_np, max_frames=max_frames)       # warm (allocates the staging buffers)
    assert len(r.frames) == args.frames
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        r = phy.rx_host(xh_np, max_frames=max_frames)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = n * world / float(te[0]) / 1e6
    d2h = 16 + 32 * len(r.frames) + r.n_triggers * phy.byte_stride

    if rank == 0:
        peak, peak_src = peaks()
        tot_ms = sum(v[0] for v in prof.values())
        dom = max(prof, key=lambda k: prof[k][0])
        dom_ms, dom_calls = prof[dom]
        # algorithmic bytes one launch of the dominant kernel is responsible for:
        # the frame kernel: 80 444 B per frame; the sync kernel: 8 B per sample.
        algo = FRAME_ALGO_BYTES * args.frames if dom.startswith("rx_frame") else 8 * n
        achieved = algo / (dom_ms / dom_calls * 1e-3) / 1e9
        traffic = None
        try:   # dram bytes per algorithmic byte from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr[dom]["dram_bytes_per_algorithmic_byte"] * algo
        except Exception:
            pass
        chain_gbs = (8 * n + args.frames * (1500 + 32)) * world * args.steps / (ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": args.frames, "samples_per_gpu": n,
                       "l2": "inputs (%.1f GB/GPU) larger than L2" % (8 * n / 1e9),
                       "parallelism": "independent streams sharded 1/GPU; per-frame stats all-gathered over NCCL"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_share_of_step": dom_ms / tot_ms,
                         "chain_frac": chain_gbs / world / peak,
                         "note": "achieved = algorithmic bytes of the kernel's launch / its mean CUDA-event duration; "
                                 "chain_frac = whole-RX algorithmic GB/s per GPU / peak"},
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": int(d2h),
                    "api": "ofdmx_rx_host (C ABI, pinned host input)"},
            "gpu_launches": launches, "clocks": clocks,
            "stats": {k: int(v) for k, v in summ.items() if k != "per_rank"},
        }
        # the CPU port is timed on rank 0 at N=1 only (at N>1 the other ranks' host threads would compete with it)
        out["cpu_baseline"] = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(args.ref_frames)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

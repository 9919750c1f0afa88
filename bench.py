#!/usr/bin/env python3
"""bench.py -- OFDM RX Msamples/s.  The headline is BASELINE.json configs[2] (the configuration the metric is
quoted on): fft_len 1024, cp 72, 600 data carriers, 16-QAM, 1500-byte packets (+CRC-32, scrambler), CFO 0.3
subcarriers + 4-tap multipath + AWGN 25 dB, 65 536 back-to-back frames, one stream per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
  python bench.py --config {0,1,2,3,4} ...                 one BASELINE.json configuration at its stated size
  python bench.py --impl reference ...                     CPU arm: the oracle port of the GNU Radio chain

A step = one pass of the hot path (full RX chain sync -> ... -> CRC; configs[4]: the Schmidl & Cox stage alone)
over the rank's resident sample buffer.  `value` = samples of all ranks / max-over-ranks device time of the K
steps.  `e2e` = the same through the host-buffer call (H2D + D2H inside the timed region).  Every configuration
carries a gate: the GPU result on a sub-sample must equal the oracle's on the same samples (trigger indices,
flags, header fields, payload bytes), and size-independent properties hold at the full size.

The default run (config 2, the line's top level) also measures the other four configurations at N=1 (key
"configs"; skip with --headline-only) and, at N>1, configs[3] sharded over the ranks (key "config3").
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
L2_BYTES = 126 * 1024 * 1024
CPU_SECONDS = 10.0          # bounded CPU sample: passes over it are repeated for about this long (headline; 3 s for the side configs)


# ------------------------------------------------------------------------------------------------------------
# The five configurations of BASELINE.json at the sizes SURVEY.md 8(d) states.
def config_table():
    import common as cm
    return {
        0: dict(name="configs[0] ofdm_hier loopback", cfg=cm.cfg_c1(2, False, 0), fft_len=64, plen=96, streams=1,
                frames=2000, gaps=(200, 2000), snr=20.0, cfo=0.0, taps=None, lead=600, tail=2400, scaling="weak",
                gate_streams=1, ref_frames=2000,
                workload="config[0]: fft_len=64 cp=16 802.11a carriers, BPSK header / QPSK payload, 96-byte packets, "
                         "AWGN 20 dB, 2000 frames separated by 200-2000 zero samples, 1 stream"),
        1: dict(name="configs[1] 4096 streams", cfg=cm.cfg_c1(2, False, 0), fft_len=64, plen=96, streams=4096,
                frames=64, gaps=(200, 2000), snr=20.0, cfo=0.0, taps=None, lead=600, tail=2400, scaling="strong",
                gate_streams=2, ref_frames=64 * 32,
                workload="config[1]: fft_len=64 cp=16, QPSK, 96-byte packets, 4096 independent streams x 64 frames "
                         "(gaps 200-2000), AWGN 20 dB"),
        2: dict(name="configs[2] headline", cfg=cm.cfg_c3(), fft_len=1024, plen=1500, streams=1, frames=65536,
                gaps=(0, 0), snr=25.0, cfo=0.3, taps=cm.MULTIPATH, lead=512, tail=4096, scaling="weak",
                gate_frames=256, ref_frames=4096,
                workload="config[2]: fft_len=1024 cp=72 600 carriers 16-QAM 1500B+CRC32 scrambled, back-to-back "
                         "frames, CFO 0.3 + 4-tap multipath + AWGN 25 dB, 1 stream/GPU"),
        3: dict(name="configs[3] 1024 streams 64-QAM", cfg=cm.cfg_c4(), fft_len=2048, plen=1500, streams=1024,
                frames=256, gaps=(0, 0), snr=40.0, cfo=0.2, taps=None, lead=512, tail=4608, scaling="strong",
                gate_streams=2, ref_frames=256,
                workload="config[3]: fft_len=2048 cp=144 1200 carriers 64-QAM 1500B+CRC32, 1024 streams x 256 "
                         "back-to-back frames sharded over the GPUs, CFO 0.2 + AWGN 40 dB"),
        4: dict(name="configs[4] preamble search", cfg=cm.cfg_c3(), fft_len=1024, plen=1500, streams=1, frames=1000,
                n_samples=1000000000, snr=10.0, snr2=16.0, scaling="weak", gate_samples=1 << 22, ref_samples=1 << 24,
                workload="config[4]: Schmidl&Cox sync only over 1e9 samples of unit-variance complex noise with 1000 "
                         "embedded config[2]-format frames at 10 dB (second pass: the same frames at 16 dB)"),
    }


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """Host threads the CPU arm may use: the cores this process is allowed on (torchrun exports OMP_NUM_THREADS=1
    to its workers; the reference arm sets the OpenMP team size explicitly instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


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


# ------------------------------------------------------------------------------------------------------------
# Synthetic inputs, generated on the GPU (torch is only the data generator here).
def make_streams(phy, C, n_streams, seed, dev, stream0=0):
    """n_streams rows of C['frames'] frames each: GPU TX of random packets, gaps, multipath, CFO, AWGN.
    Returns (x [n_streams, L] complex64, payload uint8 [n_streams*frames*plen], frame start offsets int64
    [n_streams, frames] relative to the row, frame length in samples)."""
    import torch
    F, plen, fft_len = C["frames"], C["plen"], C["fft_len"]
    g = torch.Generator(device=dev).manual_seed(seed)
    npk = n_streams * F
    payload = torch.randint(0, 256, (npk * plen,), dtype=torch.uint8, device=dev, generator=g)
    off = torch.arange(npk + 1, dtype=torch.int64, device=dev) * plen
    FS = int(phy.frame_samples(plen))
    s = torch.empty(npk * FS, dtype=torch.complex64, device=dev)
    phy.tx((payload, off), out=s)
    torch.cuda.synchronize()
    s = s.view(n_streams, F, FS)
    lo, hi = C["gaps"]
    if hi > 0:
        gaps = torch.randint(lo, hi + 1, (n_streams, F), device=dev, generator=g, dtype=torch.int64)
        gaps[:, 0] = 0
    else:
        gaps = torch.zeros((n_streams, F), dtype=torch.int64, device=dev)
    starts = C["lead"] + torch.cumsum(gaps, 1) + torch.arange(F, device=dev, dtype=torch.int64)[None, :] * FS
    L = int(starts[:, -1].max()) + FS + C["tail"]
    L = (L + 15) // 16 * 16                                  # 128-byte rows: every stream qualifies for the TMA / cp.async loads
    x = torch.zeros((n_streams, L), dtype=torch.complex64, device=dev)
    if hi == 0:
        x[:, C["lead"]: C["lead"] + F * FS] = s.reshape(n_streams, F * FS)
    else:
        ar = torch.arange(FS, device=dev, dtype=torch.int64)
        rows = torch.arange(n_streams, device=dev, dtype=torch.int64)[:, None]
        for f in range(F):
            x[rows, starts[:, f, None] + ar[None, :]] = s[:, f, :]
    pw = float((s[: min(n_streams, 8)].abs() ** 2).mean())
    del s
    if C["taps"] is not None:
        y = torch.zeros_like(x)
        for d, v in C["taps"]:
            y[:, d:] += x[:, : L - d] * complex(v)
        x = y
        del y
    sig = float(np.sqrt(pw / 10 ** (C["snr"] / 10) / 2))
    CW = 1 << 24                                             # CFO + AWGN in blocks of <= 16 M samples (bounded temporaries)
    for c0 in range(0, L, CW):
        c1 = min(c0 + CW, L)
        rot = None
        if C["cfo"]:
            t = torch.arange(c0, c1, device=dev, dtype=torch.float64)
            rot = torch.polar(torch.ones_like(t), 2 * np.pi * C["cfo"] / fft_len * t).to(torch.complex64)
            del t
        rows = max(1, CW // (c1 - c0))
        for a in range(0, n_streams, rows):
            b = min(a + rows, n_streams)
            if rot is not None:
                x[a:b, c0:c1] *= rot
            x[a:b, c0:c1] += torch.view_as_complex(torch.randn(b - a, c1 - c0, 2, device=dev, generator=g) * sig)
    return x, payload, starts, FS


def make_search_stream(phy, C, seed, dev):
    """config[4]: unit-variance complex Gaussian noise with C['frames'] config[2]-format frames added at drawn,
    non-overlapping offsets.  Returns (x [n] complex64, offsets int64 numpy, frame waveform scaled to 0 dB SNR)."""
    import torch
    n, F = C["n_samples"], C["frames"]
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    for a in range(0, n, 1 << 25):
        b = min(a + (1 << 25), n)
        x[a:b] = torch.view_as_complex(torch.randn(b - a, 2, device=dev, generator=g) * float(np.sqrt(0.5)))
    rng = np.random.default_rng(6)
    payload = torch.from_numpy(rng.integers(0, 256, C["plen"], dtype=np.uint8)).to(dev)
    off = torch.tensor([0, C["plen"]], dtype=torch.int64, device=dev)
    fr, _ = phy.tx((payload, off), out=torch.empty(int(phy.frame_samples(C["plen"])), dtype=torch.complex64, device=dev))
    torch.cuda.synchronize()
    fr = fr / float(torch.sqrt((fr.abs() ** 2).mean()))       # unit power = 0 dB against the noise
    FS = fr.numel()
    # non-overlapping offsets: one frame per slot of n/F samples, at a drawn position inside it; the first four in the
    # first gate_samples so that the oracle gate sees frames
    slot = n // F
    offs = np.arange(F, dtype=np.int64) * slot + rng.integers(2048, slot - FS - 4096, F)
    k = min(4, F)
    offs[:k] = np.sort(rng.choice(np.arange(4096, min(C["gate_samples"], k * slot) - 2 * FS, 2 * FS), k, replace=False))
    return x, np.sort(offs), fr


def add_frames(x, offs, fr, amp):
    for o in offs:
        x[int(o): int(o) + fr.numel()] += fr * float(amp)


# ------------------------------------------------------------------------------------------------------------
def oracle_gate_rx(C, phy, x_rows, byte_stride):
    """GPU == oracle on the given rows (each a whole stream or a prefix): triggers, flags, header fields, bytes."""
    import torch
    import common as cm
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    orc = cm.make_oracle(C["cfg"])
    n_frames = n_crc = 0
    for row in x_rows:
        xr = row.contiguous()
        res = phy.rx(xr)
        ref = orc.rx(xr.cpu().numpy(), byte_stride=byte_stride, want_z=False)
        rf, gf = ref["frames"], res.frames
        assert len(gf) == len(rf), "gate: GPU %d frames, oracle %d" % (len(gf), len(rf))
        for k in ("trigger", "pkt_len", "pkt_num", "frame_syms", "carr_offset"):
            assert np.array_equal(gf[k].astype(np.int64), rf[k].astype(np.int64)), "gate: %s differs" % k
        assert np.array_equal(gf["flags"] & 7, rf["flags"] & 7), "gate: flags differ"
        sl = res.slots[torch.from_numpy(gf["slot"].astype(np.int64)).to(res.slots.device)].cpu().numpy()
        for i in range(len(gf)):
            nb = int(gf["pkt_len"][i])
            assert np.array_equal(sl[i, :nb], ref["bytes"][i, :nb]), "gate: payload bytes of frame %d differ" % i
        n_frames += len(gf)
        n_crc += int(np.count_nonzero(gf["flags"] & 2))
    return {"frames_compared": n_frames, "crc_ok": n_crc, "against": "oracle (orc_rx), records and payload bytes equal"}


def time_steps(enqueue, steps, flush, barrier):
    """K timed steps on the current stream.  Returns (ms of the whole bracket with the flushes taken out, per-step
    ms list).  With `flush` (inputs smaller than L2) a buffer larger than L2 is rewritten between the steps."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for a, b in ev:
        if flush is not None:
            flush.add_(1)
        a.record()
        enqueue()
        b.record()
    barrier()
    per = [a.elapsed_time(b) for a, b in ev]
    if flush is None:
        tot = ev[0][0].elapsed_time(ev[-1][1])
    else:
        tot = float(sum(per))
    return tot, per


def run_rx_config(cid, C, args, dev, world, rank, dist, clock_sampler=None, with_cpu=True):
    """One RX configuration on this rank's share.  Returns the result dict (rank 0) or None."""
    import torch
    from ofdm_tools import OfdmPhy
    from ofdm_tools import dist as odist
    local = dev.index
    n_streams_all = args.streams if (args.streams and C["streams"] > 1) else C["streams"]
    if C["scaling"] == "strong":
        mine = odist.shard_streams(n_streams_all, rank, world)
        n_streams = len(mine)
        seed = 1000 * cid + 17 + mine.start if n_streams else 0
    else:
        n_streams, seed = n_streams_all, 1000 * cid + 17 + rank
    frames = args.frames if (cid == 2 and args.frames) else C["frames"]
    C = dict(C, frames=frames)
    max_pkt = C["plen"] + 4
    phy = OfdmPhy(device=local, tx_scale=0.01, max_pkt_bytes=max_pkt, **C["cfg"])
    x, payload, starts, FS = make_streams(phy, C, n_streams, seed, dev)
    n = x.numel()
    L = x.shape[1]
    torch.cuda.synchronize()
    max_frames = n_streams * frames + 64 * n_streams
    bufs = phy.rx_buffers(max_frames, dev)
    gathered = torch.zeros((world, 4), dtype=torch.int32, device=dev)
    flush = torch.zeros(L2_BYTES * 2 // 4, dtype=torch.int32, device=dev) if 8 * n < 2 * L2_BYTES else None

    def enqueue():
        phy.rx_enqueue(x, bufs)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), bufs["counts"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        enqueue()
    res = phy.rx_collect(bufs)
    # ---- gates.  (1) full size, size-independent: every transmitted frame is found at its position, and every
    # frame whose CRC-32 passes carries exactly the bytes that were sent (compared on the GPU over ALL frames)
    fr = res.frames
    assert len(fr) > 0, "no frame decoded"
    st_idx = torch.from_numpy(fr["stream"].astype(np.int64)).to(dev)
    trig = torch.from_numpy(fr["trigger"].astype(np.int64)).to(dev)
    skey = (torch.arange(n_streams, device=dev, dtype=torch.int64)[:, None] * L + starts).reshape(-1)   # ascending
    gidx = (torch.searchsorted(skey, st_idx * L + trig, right=True) - 1).clamp_(0, n_streams * frames - 1)
    fidx = gidx % frames
    dpos = st_idx * L + trig - skey[gidx]
    want_lo, want_hi = C["fft_len"], C["fft_len"] + 2 * (phy.cp_len + 24)
    located = (dpos >= want_lo) & (dpos <= want_hi)           # trigger ~ frame start + fft_len + cp/2 (+ channel delay)
    n_found = int(torch.unique(gidx[located]).numel())
    crc = torch.from_numpy((fr["flags"] & 2).astype(np.bool_)).to(dev) & located
    slots = res.slots[torch.from_numpy(fr["slot"].astype(np.int64)).to(dev)][:, : C["plen"]]
    sent = payload.view(n_streams * frames, C["plen"])[gidx]
    same = (slots == sent).all(1)
    n_crc_ok = int(crc.sum())
    n_bad = int((crc & ~same).sum())
    n_same = int((same & located).sum())
    if phy.crc_mode:
        assert n_bad == 0, "gate: %d frames pass the CRC-32 with bytes that differ from what was sent" % n_bad
    else:
        n_crc_ok = n_bad = None                               # no in-graph CRC in this configuration
    min_found = 0.999 if C["snr"] >= 25.0 else 0.98
    assert n_found >= min_found * n_streams * frames, "gate: found %d of %d frames" % (n_found, n_streams * frames)
    gate = {"full_size": {"frames_sent": n_streams * frames, "frames_found": n_found, "records": int(len(fr)),
                          "crc_ok": n_crc_ok, "crc_ok_with_wrong_bytes": n_bad, "payload_equal_to_sent": n_same}}
    # (2) GPU == oracle on a sub-sample (rank 0)
    if rank == 0:
        if "gate_frames" in C:
            k = min(C["gate_frames"], frames)
            rows = [x[0, : C["lead"] + k * FS + 2 * (phy.fft_len + phy.cp_len)]]
        else:
            rows = [x[i] for i in range(min(C["gate_streams"], n_streams))]
        gate["oracle"] = oracle_gate_rx(C, phy, rows, phy.byte_stride)

    if clock_sampler is not None and rank == 0:
        clock_sampler.start()
    launches0 = phy.launch_count()
    phy.profile(True)
    ms, per = time_steps(enqueue, args.steps, flush, barrier)
    clocks = clock_sampler.stop() if (clock_sampler is not None and rank == 0) else None
    res = phy.rx_collect(bufs)
    assert len(res.frames) == len(fr)
    prof = phy.profile_read()
    phy.profile(False)
    launches = phy.launch_count() - launches0
    t = torch.tensor([ms, float(launches), float(n), float(n_found), float(n_crc_ok or 0), float(n_streams * frames)]
                     + [float(np.min(per)), float(np.median(per))], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, best, med = float(tmax[0]), float(tmax[6]), float(tmax[7])
        launches, n_all = int(t[1]), int(t[2])
        gate["full_size"].update(frames_sent=int(t[5]), frames_found=int(t[3]), crc_ok=int(t[4]) if phy.crc_mode else None,
                                 records="%d on rank 0" % len(fr), payload_equal_to_sent="%d on rank 0" % n_same)
    else:
        best, med, n_all = float(np.min(per)), float(np.median(per)), n
    value = n_all * args.steps / (ms * 1e-3) / 1e6

    # ---- e2e: host buffers through ofdmx_rx_host (pinned input), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(x.shape, dtype=torch.complex64, pin_memory=True)
        xh.copy_(x)
        xh_np = xh.numpy()
        r = phy.rx_host(xh_np, max_frames=max_frames)       # warm (allocates the staging buffers)
        assert len(r.frames) == len(fr)
        # the box's host-to-device ceiling for this shape: the bare pinned copy, all ranks at once (what bounds e2e)
        xd = torch.empty_like(x)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            xd.copy_(xh, non_blocking=True)
        torch.cuda.synchronize()
        tc = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        h2d_ceiling = 8 * n / float(tc[0]) / 1e9
        del xd
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r = phy.rx_host(xh_np, max_frames=max_frames)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        d2h = 16 + 32 * len(r.frames) + r.n_triggers * phy.byte_stride
        e2e = {"value": n_all / float(te[0]) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * n,
               "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "h2d_gbs_per_rank": 8 * n / float(te[0]) / 1e9, "api": "ofdmx_rx_host (C ABI, pinned host input)",
               "bare_pinned_h2d_copy_gbs_per_rank": h2d_ceiling,
               "note": "bare_pinned_h2d_copy = cudaMemcpyAsync of the same pinned buffer alone, all ranks concurrently "
                       "(max over ranks): the box's host-to-device ceiling for this shape; e2e cannot exceed "
                       "8 B/sample / that"}
        del xh, xh_np
    # ---- the hier-block surface: analog.agc2_cc in front of the receiver (ofdm_radio_hier.rx / ofdm_tx_rx_hier.rx with
    # their default agc=True, python/ofdm_radio_hier.py:180-181), headline configuration only
    facade = None
    if cid == 2 and not args.no_agc:
        y = torch.empty_like(x)
        gain = torch.ones(n_streams, dtype=torch.float32, device=dev)

        def enqueue_agc():
            gain.fill_(1.0)
            phy.agc2(x, gain=gain, out=y)
            phy.rx_enqueue(y, bufs)

        for _ in range(2):
            enqueue_agc()
        ra = phy.rx_collect(bufs)
        ms_a, per_a = time_steps(enqueue_agc, max(3, args.steps // 2), None, barrier)
        agc_only = time_steps(lambda: phy.agc2(x, gain=gain, out=y), 3, None, barrier)[1]
        facade = {"api": "agc2_cc(1e-1, 1e-2, 1.0, 1.0) + ofdm_rx on one stream (what ofdm_radio_hier.rx() runs by default)",
                  "value": n * len(per_a) / (ms_a * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_a / len(per_a),
                  "agc2_ms": float(np.median(agc_only)), "agc2_msamples_per_s": n / (float(np.median(agc_only)) * 1e-3) / 1e6,
                  "frames": int(len(ra.frames))}
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as O
            P = min(n, 1 << 25)
            gain.fill_(1.0)
            yp, gp = phy.agc2(x[:, :P].contiguous(), gain=gain)
            rp, gr = O.agc2(x[0, :P].cpu().numpy())
            facade["gate"] = {"samples": P, "output_equal_to_oracle": bool(np.array_equal(yp[0].cpu().numpy(), rp)),
                              "gain_equal": bool(np.float32(gr) == gp.cpu().numpy()[0])}
            assert facade["gate"]["output_equal_to_oracle"] and facade["gate"]["gain_equal"], "agc2 differs from the oracle"
            # the receiver behind the AGC, GPU vs oracle on the first 64 frames of the AGC output
            import common as cm
            Q = C["lead"] + 64 * FS
            og = cm.make_oracle(C["cfg"]).rx(rp[:Q], want_z=False, byte_stride=phy.byte_stride)
            gg = phy.rx(yp[0, :Q].contiguous())
            facade["gate"]["rx_behind_agc_equal_to_oracle"] = bool(np.array_equal(gg.frames["trigger"], og["frames"]["trigger"]))
            facade["gate"]["frames_in_first_64"] = int(len(og["frames"]))
            assert facade["gate"]["rx_behind_agc_equal_to_oracle"]
            facade["note"] = ("agc2_cc(attack 0.1, decay 0.01) has a time constant of 1 / (0.01 |x|) ~ 400 samples on this signal, "
                              "shorter than the fft_len/2 = 512 lag of the Schmidl & Cox correlation: the gain differs between "
                              "the two preamble halves, the metric stays below 0.9 and the reference chain itself (oracle) "
                              "detects almost nothing at fft_len 1024 behind its AGC -- throughput is what this leg reports")
        del y
    if rank != 0:
        return None
    peak, peak_src = peaks()
    tot_ms = sum(v[0] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms, dom_calls = prof[dom]
    algo_frame = 8 * FS + C["plen"] + 32                     # samples once + payload + record
    algo_chain = 8 * n + n_streams * frames * (C["plen"] + 32)
    algo = algo_frame * n_streams * frames if dom.startswith("rx_frame") else 8 * n
    achieved = algo / (dom_ms / dom_calls * 1e-3) / 1e9
    traffic = None
    try:   # dram bytes per algorithmic byte from the committed ncu --set full capture (headline kernels)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if cid == 2:
            traffic = tr[dom]["dram_bytes_per_algorithmic_byte"] * algo
        elif dom in tr.get("configs", {}).get(str(cid), {}):
            traffic = tr["configs"][str(cid)][dom]["dram_bytes_per_algorithmic_byte"] * algo
    except Exception:
        pass
    out = {
        "value": value, "unit": "Msamples/s", "ms_per_step": ms / args.steps, "best_ms": best, "median_ms": med,
        "steps": args.steps, "scaling": C["scaling"],
        "config": {"workload": C["workload"], "streams_this_rank": n_streams, "frames_per_stream": frames,
                   "samples_this_rank": n,
                   "l2": ("inputs (%.2f GB on this rank) larger than L2" % (8 * n / 1e9)) if flush is None
                   else "inputs (%.0f MB) fit L2: a %d MB buffer is rewritten between the timed steps" % (8 * n / 1e6, 2 * L2_BYTES >> 20),
                   "parallelism": "independent streams sharded over the ranks; per-rank frame counters all-gathered over NCCL"},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel_share_of_step": dom_ms / tot_ms,
                     "chain_frac": algo_chain * args.steps / (ms * 1e-3) / 1e9 / peak,
                     "note": "achieved = algorithmic bytes of the kernel's launch / its mean CUDA-event duration; "
                             "chain_frac = whole-RX algorithmic GB/s of this rank / peak"},
        "kernels_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "gate": gate,
    }
    if facade is not None:
        out["hier_block_rx_with_agc"] = facade
    if with_cpu and world == 1:
        out["cpu_baseline"] = cpu_baseline_rx(C, C["ref_frames"])
    return out


def run_sync_config(cid, C, args, dev, world, rank, dist):
    """configs[4]: the Schmidl & Cox stage alone (ofdmx_sync) over 1e9 samples per GPU."""
    import torch
    import common as cm
    from ofdm_tools import OfdmPhy
    local = dev.index
    n = args.search_samples or C["n_samples"]
    C = dict(C, n_samples=n, frames=max(8, int(C["frames"] * n / C["n_samples"])))
    phy = OfdmPhy(device=local, tx_scale=0.01, max_pkt_bytes=C["plen"] + 4, **C["cfg"])
    x, offs, fr = make_search_stream(phy, C, 5 + rank, dev)
    FS = fr.numel()
    max_trig = 1 << 16
    trig = torch.zeros(max_trig, dtype=torch.int64, device=dev)
    cfo = torch.zeros(max_trig, dtype=torch.float32, device=dev)
    st = torch.zeros(max_trig, dtype=torch.int32, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    from ofdm_tools import _lib
    L = _lib.load()
    import ctypes as Cc

    def enqueue(xx=None):
        xx = x if xx is None else xx
        _lib.check(L.ofdmx_sync(phy.ctx, xx.data_ptr(), 1, xx.numel(), xx.numel(), trig.data_ptr(), cfo.data_ptr(),
                                st.data_ptr(), max_trig, counts.data_ptr(),
                                Cc.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), phy.ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def detections():
        c = counts.cpu().numpy()
        assert not c[2], "trigger overflow"
        return trig[: int(c[0])].cpu().numpy()

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    orc = cm.make_oracle(C["cfg"])
    passes = []
    amp_prev = 0.0
    for snr in (C["snr"], C["snr2"]):
        amp = float(np.sqrt(10 ** (snr / 10)))
        add_frames(x, offs, fr, amp - amp_prev)
        amp_prev = amp
        torch.cuda.synchronize()
        for _ in range(args.warmup):
            enqueue()
        det = detections()
        # property at full size: which embedded frames have a trigger at their preamble (start + fft_len + cp/2 +- cp)
        exp = offs + C["fft_len"] + phy.cp_len // 2
        j = np.searchsorted(det, exp - phy.cp_len)
        hit = (j < len(det)) & (np.abs(det[np.minimum(j, max(len(det) - 1, 0))] - exp) <= phy.cp_len) if len(det) else np.zeros(len(exp), bool)
        # oracle gate on the first gate_samples (they hold four of the frames)
        g = min(C["gate_samples"], n)
        enqueue(x[:g])
        dg = detections()
        rt, _ = orc.sync(x[:g].cpu().numpy())
        assert np.array_equal(dg, rt), "gate: detection indices differ from the oracle on the first %d samples" % g
        enqueue()
        passes.append({"snr_db": snr, "detections": int(len(det)), "embedded_frames": int(len(offs)),
                       "embedded_frames_detected": int(hit.sum()), "other_detections": int(len(det) - hit.sum()),
                       "oracle_gate": {"samples": int(g), "detections_compared": int(len(rt)), "equal": True}})
    phy.profile(True)
    launches0 = phy.launch_count()
    ms, per = time_steps(enqueue, args.steps, None, barrier)
    prof = phy.profile_read()
    phy.profile(False)
    launches = phy.launch_count() - launches0
    t = torch.tensor([ms, float(launches), float(np.min(per)), float(np.median(per))], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(t[1])
        t = tmax
    value = n * world * args.steps / (ms * 1e-3) / 1e6
    # e2e: pinned host samples -> device -> ofdmx_sync -> detections back on the host
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(n, dtype=torch.complex64, pin_memory=True)
        xh.copy_(x)
        xd = torch.empty_like(x)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, args.e2e_steps // 2)):
            xd.copy_(xh, non_blocking=True)
            enqueue(xd)
            d = detections()
        e2e_s = (time.perf_counter() - t0) / max(1, args.e2e_steps // 2)
        e2e = {"value": n * world / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * n,
               "d2h_bytes_per_step": 16 + 8 * len(d), "api": "OfdmPhy.sync path: pinned host -> device copy + ofdmx_sync + triggers to host"}
        del xh, xd
    if rank != 0:
        return None
    peak, peak_src = peaks()
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms, dom_calls = prof[dom]
    achieved = 8 * n / (dom_ms / dom_calls * 1e-3) / 1e9
    out = {
        "value": value, "unit": "Msamples/s", "ms_per_step": ms / args.steps, "best_ms": float(t[2]), "median_ms": float(t[3]),
        "steps": args.steps, "scaling": "weak",
        "config": {"workload": C["workload"], "samples_per_gpu": n, "l2": "inputs (%.1f GB/GPU) larger than L2" % (8 * n / 1e9),
                   "note": "S&C needs (S/(S+N))^2 >= 0.9, i.e. > 12.8 dB SNR: at the 10 dB the configuration states the "
                           "reference's detector (threshold 0.9) finds almost none of the frames, as the oracle confirms; the "
                           "second pass embeds the same frames at 16 dB"},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel_share_of_step": dom_ms / sum(v[0] for v in prof.values()),
                     "chain_frac": 8 * n * args.steps / (ms * 1e-3) / 1e9 / peak},
        "kernels_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        "e2e": e2e, "gpu_launches": launches, "gate": {"passes": passes},
    }
    if world == 1:
        out["cpu_baseline"] = cpu_baseline_sync(C, x[: min(n, C["ref_samples"])].cpu().numpy())
    return out


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's GNU Radio chain, on all host threads
def ref_stream(C, n_frames, seed=1):
    import common as cm
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    orc = cm.make_oracle(C["cfg"])
    rng = np.random.default_rng(seed)
    pk = cm.rand_packets(rng, n_frames, C["plen"])
    s, off = orc.tx(pk)
    x = cm.channel(cm.split_frames(s, off), rng, gaps=C["gaps"], lead=C["lead"], tail=C["tail"], snr_db=C["snr"],
                   cfo=C["cfo"], fft_len=C["fft_len"], taps=C["taps"])
    return orc, x


def cpu_baseline_rx(C, n_frames):
    import oracle as O
    orc, x = ref_stream(C, n_frames)
    threads = O.set_threads(host_threads())
    bs = (C["plen"] + 4 + 15) // 16 * 16
    orc.rx_baseline(x[: max(len(x) // 16, 4096)], byte_stride=bs)     # warm
    t0 = time.perf_counter()
    passes = 0
    while passes < 1 or (time.perf_counter() - t0 < CPU_SECONDS and passes < 1000):
        r = orc.rx_baseline(x, byte_stride=bs)
        passes += 1
    dt = time.perf_counter() - t0
    return {"value": passes * len(x) / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": "%d passes over %d frames (%d samples) of the same workload, %d decoded per pass, %.1f s; float32 FIR "
                      "sync as GNU Radio evaluates it + per-trigger demodulation, OpenMP over %d threads"
                      % (passes, n_frames, len(x), len(r), dt, threads)}


def cpu_baseline_sync(C, x):
    import oracle as O
    import common as cm
    orc = cm.make_oracle(C["cfg"])
    threads = O.set_threads(host_threads())
    orc.sync(x[: 1 << 18], f32=True)
    t0 = time.perf_counter()
    passes = 0
    while passes < 1 or (time.perf_counter() - t0 < CPU_SECONDS and passes < 1000):
        tr, _ = orc.sync(x, f32=True)
        passes += 1
    dt = time.perf_counter() - t0
    return {"value": passes * len(x) / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": "%d passes over the first %d samples of the same stream, %d detections, %.1f s; float32 FIR port of "
                      "ofdm_sync_sc_cfb, OpenMP over %d threads" % (passes, len(x), len(tr), dt, threads)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    C = config_table()[args.config]
    threads = O.set_threads(host_threads())
    if args.config == 4:
        import common as cm
        orc = cm.make_oracle(C["cfg"])
        rng = np.random.default_rng(5)
        n = args.ref_samples
        x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)).astype(np.complex64)
        pk = cm.rand_packets(rng, 1, C["plen"])
        s, _ = orc.tx(pk)
        s = s / np.sqrt(np.mean(np.abs(s) ** 2)) * np.sqrt(10 ** (C["snr2"] / 10))
        for o in range(100000, n - len(s), n // 4):
            x[o:o + len(s)] += s
        run = lambda: len(orc.sync(x, f32=True)[0])
        sample = "%d samples of noise with embedded frames per step" % n
    else:
        n_frames = args.ref_frames or C["ref_frames"]
        orc, x = ref_stream(C, n_frames)
        bs = (C["plen"] + 4 + 15) // 16 * 16
        run = lambda: len(orc.rx_baseline(x, byte_stride=bs))
        sample = "%d frames (%d samples) of the workload per step" % (n_frames, len(x))
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    nf = 0
    for _ in range(args.steps):
        nf += run()
    dt = time.perf_counter() - t0
    val = len(x) * args.steps / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC if args.config == 2 else "OFDM RX Msamples/s (%s)" % C["name"],
        "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": C["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C["workload"],
                   "what": "CPU restatement of the GNU Radio chain (oracle port; GNU Radio itself is not installable here): "
                           "float32 FIR Schmidl&Cox + per-trigger demodulation, OpenMP over all host threads"},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "results_per_step": nf // max(1, args.steps),
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[0, 1, 2, 3, 4], help="BASELINE.json configs[i]; 2 = the headline")
    ap.add_argument("--frames", type=int, default=0, help="config 2 only: frames per GPU (default 65536 = 646 M samples = 5.2 GB)")
    ap.add_argument("--streams", type=int, default=0, help="configs 1 and 3 only: number of streams (default: as stated)")
    ap.add_argument("--search-samples", type=int, default=0, help="config 4 only: samples per GPU (default 1e9)")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames in the CPU sample (default: per configuration)")
    ap.add_argument("--ref-samples", type=int, default=1 << 23, help="--impl reference --config 4: samples per step")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-agc", action="store_true", help="skip the agc2 + rx leg of the headline configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="do not measure the other configurations next to the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    table = config_table()
    C = table[args.config]
    sampler = ClockSampler(local)
    with_cpu = not args.no_cpu_baseline

    def one(cid, sampler=None, with_cpu=True):
        torch.cuda.empty_cache()
        if cid == 4:
            return run_sync_config(cid, table[cid], args, dev, world, rank, dist)
        return run_rx_config(cid, table[cid], args, dev, world, rank, dist, clock_sampler=sampler, with_cpu=with_cpu)

    head = one(args.config, sampler, with_cpu)
    extra = {}
    if args.config == 2 and not args.headline_only:
        # the other configurations next to the headline: all of them on one GPU, the sharded one on N GPUs
        global CPU_SECONDS
        saved = (args.steps, args.e2e_steps)
        args.steps, args.e2e_steps = min(args.steps, 10), min(args.e2e_steps, 3)
        CPU_SECONDS = 3.0
        for cid in ((0, 1, 3, 4) if world == 1 else (3,)):
            try:
                r = one(cid, None, with_cpu)
            except Exception as e:           # a side measurement must not take the headline down
                r = {"error": "%s: %s" % (type(e).__name__, e)} if rank == 0 else None
            if rank == 0:
                extra[str(cid)] = r
        args.steps, args.e2e_steps = saved
    if rank == 0:
        out = {"metric": METRIC if args.config == 2 else "OFDM RX Msamples/s (%s)" % C["name"],
               "value": head["value"], "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": head["ms_per_step"], "best_ms": head["best_ms"], "median_ms": head["median_ms"],
               "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        for k in ("config", "roofline", "kernels_ms_per_step", "e2e", "gpu_launches", "clocks", "gate", "cpu_baseline",
                  "hier_block_rx_with_agc"):
            out[k] = head.get(k)
        if "clocks" not in head or head.get("clocks") is None:
            out["clocks"] = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if extra:
            out["configs"] = extra
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

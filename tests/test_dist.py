"""CPU tests of the multi-GPU host logic: segment planning / stitching against the unsplit demux
rule, and the statistics all-gather under torch.distributed (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest

import common as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _synthetic_records(rng, n_samples, fft_len, cp_len, n_trig):
    """Random per-trigger records as the frame kernel would emit them (emit-all mode)."""
    from ofdm_tools import FRAME_DTYPE, _lib
    D = fft_len + cp_len
    trig = np.sort(rng.choice(np.arange(fft_len, n_samples - 1), size=n_trig, replace=False))
    rec = np.zeros(n_trig, FRAME_DTYPE)
    rec["trigger"] = trig
    rec["slot"] = np.arange(n_trig)
    for i in range(n_trig):
        fl = 0
        if trig[i] + 3 * D <= n_samples:
            fl |= _lib.F_HDR_SEEN
            if rng.random() < 0.6:
                fl |= _lib.F_HDR_OK
                L = int(rng.integers(0, 7))
                rec["frame_syms"][i] = L
                rec["pkt_len"][i] = 10 * L
                if trig[i] + (3 + L) * D <= n_samples:
                    fl |= _lib.F_COMPLETE | _lib.F_CRC_OK
        rec["flags"][i] = fl
    return rec


@pytest.mark.parametrize("seed,world", [(0, 2), (1, 3), (2, 8), (3, 5)])
def test_segment_stitching_equals_unsplit_chain(seed, world):
    from ofdm_tools import dist
    rng = np.random.default_rng(seed)
    fft_len, cp_len, n = 64, 16, 200000
    D = fft_len + cp_len
    rec = _synthetic_records(rng, n, fft_len, cp_len, 1500)
    # keep every frame away from the stream end so that completeness does not depend on the split
    rec = rec[rec["trigger"] + 9 * D + D < n]
    ref = rec[dist.demux_chain(rec, fft_len, cp_len, D, n)]
    plan = dist.plan_segments(n, world, fft_len, cp_len, 9 * D)
    segs = []
    for (l0, a, b, l1) in plan:
        s = rec[(rec["trigger"] >= l0) & (rec["trigger"] < l1)].copy()
        s["trigger"] -= l0
        segs.append(s)
    got, own = dist.merge_segments(segs, plan, fft_len, cp_len, D, n)
    assert np.array_equal(got["trigger"], ref["trigger"])
    assert np.array_equal(got["slot"], ref["slot"])
    assert len(own) == len(got) and own.min() >= 0 and own.max() < world


def test_demux_chain_rules():
    from ofdm_tools import dist, FRAME_DTYPE, _lib
    D = 80
    ok = _lib.F_HDR_SEEN | _lib.F_HDR_OK | _lib.F_COMPLETE
    rec = np.zeros(6, FRAME_DTYPE)
    rec["trigger"] = [100, 150, 101 + 4 * D, 100 + 5 * D, 2000, 2001]
    rec["frame_syms"] = [3, 3, 3, 3, 0, 2]
    rec["flags"] = [ok, ok, ok, ok, _lib.F_HDR_SEEN, ok]
    emit = dist.demux_chain(rec, 64, 16, D, 100000)
    # frame 0 covers [100, 100 + 6D - D): trigger 150 ignored, 101+4D ignored (< 100+5D), 100+5D accepted;
    # failed header at 2000 -> resume at 2001
    assert list(emit) == [True, False, False, True, False, True]
    emit1 = dist.demux_chain(rec, 64, 16, 1, 100000)        # GNU Radio >= 3.7.10: holdoff of one item
    assert list(emit1) == [True, False, False, False, False, True]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (os.path.join(ROOT, "gr-ofdm_tools_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist_t
    from ofdm_tools import dist
    dist_t.init_process_group("gloo", rank=rank, world_size=world)
    streams = list(dist.shard_streams(5, rank, world))
    summ = {"n_samples": 1000 * len(streams), "n_triggers": 7 + rank, "n_frames": 5 + rank, "n_crc_ok": 4 + rank,
            "n_payload_bytes": 100 * (rank + 1), "n_ranks": 1}
    tot = dist.gather_stats(summ)
    # segment records travel as Python objects; every rank stitches the same global list
    rng = np.random.default_rng(5)
    n, fft_len, cp_len = 120000, 64, 16
    D = fft_len + cp_len
    rec = _synthetic_records(rng, n, fft_len, cp_len, 700)
    rec = rec[rec["trigger"] + 10 * D < n]
    plan = dist.plan_segments(n, world, fft_len, cp_len, 9 * D)
    l0, a, b, l1 = plan[rank]
    mine = rec[(rec["trigger"] >= l0) & (rec["trigger"] < l1)].copy()
    mine["trigger"] -= l0
    gathered = [None] * world
    dist_t.all_gather_object(gathered, mine)
    got, own = dist.merge_segments(gathered, plan, fft_len, cp_len, D, n)
    ref = rec[dist.demux_chain(rec, fft_len, cp_len, D, n)]
    q.put((rank, {k: v for k, v in tot.items() if k != "per_rank"}, bool(np.array_equal(got["trigger"], ref["trigger"])),
           streams))
    dist_t.destroy_process_group()


def test_gloo_two_ranks_stats_and_stitching():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tot, same, streams in res:
        assert tot == {"n_samples": 5000, "n_triggers": 15, "n_frames": 11, "n_crc_ok": 9, "n_payload_bytes": 300,
                       "n_ranks": 2}
        assert same
    assert res[0][3] == [0, 1, 2] and res[1][3] == [3, 4]

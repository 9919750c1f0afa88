"""Multi-GPU plumbing: one process per GPU, independent streams / stream segments per rank, and the
only collective on this path -- an all-gather of the fixed-size per-rank frame statistics.

The reference has no distributed layer (single-process flowgraph); the sample path
needs no exchange because frames and streams are independent.
"""
import numpy as np

from . import _lib

STAT_KEYS = ("n_samples", "n_triggers", "n_frames", "n_crc_ok", "n_payload_bytes", "n_ranks")


def shard_streams(n_streams, rank, world):
    """Contiguous block of stream indices owned by `rank` (stream s -> rank s // ceil(n/world))."""
    per = (n_streams + world - 1) // world
    lo = min(rank * per, n_streams)
    return range(lo, min(lo + per, n_streams))


def plan_segments(n_samples, world, fft_len, cp_len, max_frame_samples):
    """Split one long stream into `world` contiguous segments.

    Returns [(load_start, own_start, own_stop, load_stop)] per rank.  A trigger belongs to the rank
    whose [own_start, own_stop) contains it.  The leading halo (fft_len + 2*cp_len + 2 samples) makes
    the S&C window sums and the plateau detector state identical to the unsplit stream at own_start;
    the trailing halo (max_frame_samples + fft_len + cp_len) lets every owned frame complete."""
    per = (n_samples + world - 1) // world
    lead = fft_len + 2 * cp_len + 2
    trail = max_frame_samples + fft_len + cp_len
    out = []
    for r in range(world):
        a = min(r * per, n_samples)
        b = min(a + per, n_samples)
        out.append((max(0, a - lead), a, b, min(n_samples, b + trail)))
    return out


def summarize(res, n_samples):
    """Fixed-size per-rank summary of one RX call (RxResult)."""
    f = res.frames
    crc_ok = (f["flags"] & _lib.F_CRC_OK) != 0
    return {
        "n_samples": int(n_samples), "n_triggers": int(res.n_triggers), "n_frames": int(len(f)),
        "n_crc_ok": int(crc_ok.sum()), "n_payload_bytes": int(f["pkt_len"][crc_ok].astype(np.int64).sum()),
        "n_ranks": 1,
    }


def gather_stats(summary, device=None):
    """All-gather the per-rank summaries (NCCL over NVLink on GPUs, gloo on CPU) and return the sums.
    Also returns the per-rank table under key 'per_rank' when called with torch.distributed up."""
    import torch
    import torch.distributed as dist
    vec = torch.tensor([summary[k] for k in STAT_KEYS], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(summary)
    table = [torch.empty_like(vec) for _ in range(dist.get_world_size())]
    dist.all_gather(table, vec)
    tab = torch.stack(table).cpu().numpy()
    out = {k: int(tab[:, i].sum()) for i, k in enumerate(STAT_KEYS)}
    out["per_rank"] = tab
    return out

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


def demux_chain(records, fft_len, cp_len, holdoff, n_samples, n_sync_words=2):
    """The header_payload_demux acceptance rule over per-trigger records of ONE stream,
    sorted by trigger: returns the boolean mask of the frames the flowgraph emits.  Same rule as the
    device chain kernels (csrc/ofdmx_chain.cuh); used on the host to stitch segments."""
    D = fft_len + cp_len
    pre = n_sync_words + 1                          # OFDM symbols in front of the payload
    emit = np.zeros(len(records), bool)
    pos = 0
    for i, f in enumerate(records):
        t = int(f["trigger"])
        if t < pos:
            continue
        fl = int(f["flags"])
        if t + pre * D > n_samples or not (fl & _lib.F_HDR_SEEN):
            break                                   # demux waits for header samples that never come
        if not (fl & _lib.F_HDR_OK):
            pos = t + 1
            continue
        L = int(f["frame_syms"])
        if t + (pre + L) * D > n_samples or not (fl & _lib.F_COMPLETE):
            break                                   # demux waits for payload samples that never come
        emit[i] = True
        pos = t + (pre + L) * D - holdoff if L > 0 else t + pre * D
    return emit


def merge_segments(seg_records, plan, fft_len, cp_len, holdoff, n_samples, n_sync_words=2):
    """Stitch per-segment emit-all records (triggers relative to each segment's load_start) into the
    frame list of the unsplit stream.  seg_records[r]: FRAME_DTYPE array from rank r (every trigger);
    plan: plan_segments(...).  Returns (records with absolute triggers, owner rank per record)."""
    owned, owner = [], []
    for r, (rec, (l0, a, b, l1)) in enumerate(zip(seg_records, plan)):
        rec = rec.copy()
        rec["trigger"] += l0
        keep = (rec["trigger"] >= a) & (rec["trigger"] < b)
        owned.append(rec[keep])
        owner.append(np.full(int(keep.sum()), r, np.int32))
    allrec = np.concatenate(owned) if owned else np.zeros(0)
    allown = np.concatenate(owner) if owner else np.zeros(0, np.int32)
    emit = demux_chain(allrec, fft_len, cp_len, holdoff, n_samples, n_sync_words)
    out = allrec[emit].copy()
    out["flags"] |= _lib.F_ACCEPTED
    return out, allown[emit]


def rx_segmented(phy, samples, n_segments, max_pkt_bytes=None):
    """Run one long stream as n_segments independent RX calls (what n_segments GPUs would each do) and
    stitch the results; returns (records, list of payload bytes).  `samples`: 1-D cuda complex64."""
    n = samples.numel()
    D = phy.fft_len + phy.cp_len
    max_frame = int(phy.frame_samples((max_pkt_bytes or phy.max_pkt_bytes) - (4 if phy.crc_mode else 0)))
    plan = plan_segments(n, n_segments, phy.fft_len, phy.cp_len, max_frame)
    phy.set_emit_all(True)
    try:
        segs = [phy.rx(samples[l0:l1]) for (l0, a, b, l1) in plan]
    finally:
        phy.set_emit_all(False)
    recs, own = merge_segments([s.frames for s in segs], plan, phy.fft_len, phy.cp_len,
                               phy.params.demux_holdoff, n, phy.n_sync_words)
    payloads = []
    for f, r in zip(recs, own):
        nb = int(f["pkt_len"])
        row = segs[r].slots[int(f["slot"])].cpu().numpy()
        if phy.crc_mode:
            if not (int(f["flags"]) & _lib.F_CRC_OK):
                continue
            nb -= 4
        if int(f["flags"]) & _lib.F_OVERSIZE:
            continue
        payloads.append(bytes(row[:nb]))
    return recs, payloads


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

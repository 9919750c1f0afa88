// ofdmx_chain.cuh -- header_payload_demux acceptance chain over the speculative per-trigger records.
//
// The demux is a sequential state machine: after examining trigger i it resumes its
// search at an item index that depends on i's header (failed: t+1; decoded: after the payload), and
// ignores every trigger before that.  With next[i] = first trigger at or after the resume point
// (or the first trigger of the next stream when the demux stalls / the stream ends), the examined
// triggers of ALL streams are the orbit of trigger 0 under next.  next[i] > i, so the orbit is found
// block-wise:
//   chain_next_kernel   per block of 4096 triggers: next[] and, by in-shared pointer jumping,
//                       exit[i] = first node >= block end on i's path
//   chain_entry_kernel  walk exit[] from trigger 0: <= one entry node per block (short serial walk)
//   chain_mark_kernel   per block: exact pointer doubling in shared memory from the entry node ->
//                       examined marks; emit flag = examined && header ok && frame complete
//   chain_scan_kernel / chain_emit_kernel   ordered compaction of the emitted records
#pragma once
#include "ofdmx_dev.cuh"
#include "ofdmx.h"

#define CH_B 4096
#define CH_T 1024

__global__ void __launch_bounds__(CH_T)
chain_next_kernel(const KP p, const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                  const ofdmx_frame *__restrict__ spec, const int *__restrict__ stream_start,
                  const int *__restrict__ n_trig_dev, int *__restrict__ nxt, int *__restrict__ exitp)
{
    __shared__ int e[CH_B];
    const int nt = *n_trig_dev;
    const int base = blockIdx.x * CH_B;
    if (base >= nt) return;
    const int end = min(base + CH_B, nt), cnt = end - base;
    for (int li = threadIdx.x; li < cnt; li += CH_T) {
        const int i = base + li;
        const ofdmx_frame f = spec[i];
        const int b = stream_start[trig_stream[i] + 1];          // first trigger of the next stream
        long long resume = 0;
        int nx = -1;
        if (!(f.flags & OFDMX_F_HDR_SEEN)) nx = b;                // demux stalls waiting for the header
        else if (!(f.flags & OFDMX_F_HDR_OK)) resume = f.trigger + 1;   // header CRC failed
        else if (!(f.flags & OFDMX_F_COMPLETE)) nx = b;           // demux stalls waiting for the payload
        else
            resume = (f.frame_syms > 0) ? f.trigger + (long long)(p.nsw + 1 + f.frame_syms) * p.D - p.holdoff
                                        : f.trigger + (long long)(p.nsw + 1) * p.D;
        if (nx < 0) {
            int lo = i + 1, hi = b;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (trig[mid] < resume) lo = mid + 1; else hi = mid;
            }
            nx = lo;
        }
        nxt[i] = nx;
        e[li] = nx;
    }
    __syncthreads();
    for (int round = 0; round < 12; round++) {                    // 2^12 = CH_B
        int moved = 0;
        for (int li = threadIdx.x; li < cnt; li += CH_T) {
            const int v = e[li];
            if (v < end) { e[li] = e[v - base]; moved = 1; }
        }
        if (!__syncthreads_or(moved)) break;
    }
    for (int li = threadIdx.x; li < cnt; li += CH_T) exitp[base + li] = e[li];
}

__global__ void chain_entry_kernel(const int *__restrict__ n_trig_dev, const int *__restrict__ exitp,
                                   int *__restrict__ entry, int nblk)
{
    for (int g = threadIdx.x; g < nblk; g += blockDim.x) entry[g] = -1;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nt = *n_trig_dev;
        int cur = 0;
        while (cur < nt) {
            entry[cur / CH_B] = cur;
            cur = exitp[cur];
        }
    }
}

__global__ void __launch_bounds__(CH_T)
chain_mark_kernel(const ofdmx_frame *__restrict__ spec, const int *__restrict__ n_trig_dev,
                  const int *__restrict__ nxt, const int *__restrict__ entry, uint8_t *__restrict__ emitflag,
                  int *__restrict__ blockcount, int emit_all)
{
    __shared__ int jA[CH_B], jB[CH_B];
    __shared__ uint8_t mA[CH_B], mB[CH_B];
    __shared__ int wt[33];
    const int nt = *n_trig_dev;
    const int base = blockIdx.x * CH_B;
    if (base >= nt) {
        if (threadIdx.x == 0) blockcount[blockIdx.x] = 0;
        return;
    }
    const int end = min(base + CH_B, nt), cnt = end - base;
    const int ent = entry[blockIdx.x];
    for (int li = threadIdx.x; li < cnt; li += CH_T) {
        const int nx = nxt[base + li];
        jA[li] = (nx < end) ? nx - base : CH_B;
        mA[li] = (base + li == ent) ? 1 : 0;
    }
    __syncthreads();
    int *jc = jA, *jn = jB;
    uint8_t *mc = mA, *mn = mB;
    if (ent >= 0) {
        for (int span = 1; span < cnt; span <<= 1) {
            for (int li = threadIdx.x; li < cnt; li += CH_T) mn[li] = mc[li];
            __syncthreads();
            for (int li = threadIdx.x; li < cnt; li += CH_T) {
                const int jx = jc[li];
                if (jx < cnt) {
                    if (mc[li]) mn[jx] = 1;
                    jn[li] = jc[jx];
                } else {
                    jn[li] = CH_B;
                }
            }
            __syncthreads();
            int *tj = jc; jc = jn; jn = tj;
            uint8_t *tm = mc; mc = mn; mn = tm;
        }
    }
    int local = 0;
    for (int li = threadIdx.x; li < cnt; li += CH_T) {
        const unsigned fl = spec[base + li].flags;
        // bit 0: record is emitted; bit 1: the demux examined this trigger
        const uint8_t em = (emit_all || (mc[li] && (fl & OFDMX_F_HDR_OK) && (fl & OFDMX_F_COMPLETE))) ? 1 : 0;
        emitflag[base + li] = em | (mc[li] ? 2 : 0);
        local += em;
    }
    int total;
    block_excl_scan(local, wt, total);
    if (threadIdx.x == 0) blockcount[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
chain_scan_kernel(int *__restrict__ blockcount, int nblk, ofdmx_counts *__restrict__ counts)
{
    __shared__ int wt[33];
    int carry = 0;
    for (int b0 = 0; b0 < nblk; b0 += 1024) {
        const int idx = b0 + threadIdx.x;
        const int v = (idx < nblk) ? blockcount[idx] : 0;
        int total;
        const int ex = block_excl_scan(v, wt, total);
        if (idx < nblk) blockcount[idx] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts->n_frames = carry;
}

__global__ void __launch_bounds__(CH_T)
chain_emit_kernel(const ofdmx_frame *__restrict__ spec, const int *__restrict__ n_trig_dev,
                  const uint8_t *__restrict__ emitflag, const int *__restrict__ blockbase,
                  ofdmx_frame *__restrict__ frames_out)
{
    __shared__ int wt[33];
    const int nt = *n_trig_dev;
    const int base = blockIdx.x * CH_B;
    if (base >= nt) return;
    const int end = min(base + CH_B, nt);
    int carry = blockbase[blockIdx.x];
    for (int i0 = base; i0 < end; i0 += CH_T) {
        const int i = i0 + threadIdx.x;
        const int ef = (i < end) ? emitflag[i] : 0;
        const int em = ef & 1;
        int total;
        const int ex = block_excl_scan(em, wt, total);
        if (em) {
            ofdmx_frame f = spec[i];
            if (ef & 2) f.flags |= OFDMX_F_ACCEPTED;
            frames_out[carry + ex] = f;
        }
        carry += total;
        __syncthreads();
    }
}

// All five steps in ONE single-CTA kernel for calls whose trigger capacity fits one block (<= CH_B triggers: small
// receive calls, streaming chunks): next[] -> examined marks by pointer doubling from trigger 0 -> ordered emit.  The
// five launches above cost ~50 us on such a call -- a quarter of a configs[0] step -- for microseconds of work.
__global__ void __launch_bounds__(CH_T)
chain_small_kernel(const KP p, const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                   const ofdmx_frame *__restrict__ spec, const int *__restrict__ stream_start,
                   const int *__restrict__ n_trig_dev, int emit_all, ofdmx_counts *__restrict__ counts,
                   ofdmx_frame *__restrict__ frames_out)
{
    __shared__ int jA[CH_B], jB[CH_B];
    __shared__ uint8_t mA[CH_B], mB[CH_B];
    __shared__ int wt[33];
    const int nt = min(*n_trig_dev, CH_B), cnt = nt;
    if (cnt <= 0) {
        if (threadIdx.x == 0) counts->n_frames = 0;
        return;
    }
    for (int i = threadIdx.x; i < cnt; i += CH_T) {
        const ofdmx_frame f = spec[i];
        const int b = stream_start[trig_stream[i] + 1];          // first trigger of the next stream
        long long resume = 0;
        int nx = -1;
        if (!(f.flags & OFDMX_F_HDR_SEEN)) nx = b;                // demux stalls waiting for the header
        else if (!(f.flags & OFDMX_F_HDR_OK)) resume = f.trigger + 1;   // header CRC failed
        else if (!(f.flags & OFDMX_F_COMPLETE)) nx = b;           // demux stalls waiting for the payload
        else
            resume = (f.frame_syms > 0) ? f.trigger + (long long)(p.nsw + 1 + f.frame_syms) * p.D - p.holdoff
                                        : f.trigger + (long long)(p.nsw + 1) * p.D;
        if (nx < 0) {
            int lo = i + 1, hi = b;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (trig[mid] < resume) lo = mid + 1; else hi = mid;
            }
            nx = lo;
        }
        jA[i] = (nx < cnt) ? nx : CH_B;
        mA[i] = (i == 0) ? 1 : 0;                                 // the orbit starts at trigger 0
    }
    __syncthreads();
    int *jc = jA, *jn = jB;
    uint8_t *mc = mA, *mn = mB;
    for (int span = 1; span < cnt; span <<= 1) {
        for (int li = threadIdx.x; li < cnt; li += CH_T) mn[li] = mc[li];
        __syncthreads();
        for (int li = threadIdx.x; li < cnt; li += CH_T) {
            const int jx = jc[li];
            if (jx < cnt) {
                if (mc[li]) mn[jx] = 1;
                jn[li] = jc[jx];
            } else {
                jn[li] = CH_B;
            }
        }
        __syncthreads();
        int *tj = jc; jc = jn; jn = tj;
        uint8_t *tm = mc; mc = mn; mn = tm;
    }
    int carry = 0;
    for (int i0 = 0; i0 < cnt; i0 += CH_T) {
        const int i = i0 + threadIdx.x;
        ofdmx_frame f;
        int em = 0;
        bool examined = false;
        if (i < cnt) {
            f = spec[i];
            examined = mc[i] != 0;
            em = (emit_all || (examined && (f.flags & OFDMX_F_HDR_OK) && (f.flags & OFDMX_F_COMPLETE))) ? 1 : 0;
        }
        int total;
        const int ex = block_excl_scan(em, wt, total);
        if (em) {
            if (examined) f.flags |= OFDMX_F_ACCEPTED;
            frames_out[carry + ex] = f;
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts->n_frames = carry;
}

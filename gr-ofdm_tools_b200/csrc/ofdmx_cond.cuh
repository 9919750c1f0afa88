// ofdmx_cond.cuh -- signal conditioning around the OFDM chain.
//
// analog.agc2_cc in front of ofdm_rx (python/ofdm_tx_rx_hier.py:75-76, python/ofdm_radio_hier.py:180-181) is a
// per-sample NON-LINEAR recurrence on the loop gain:
//     out = in * g;  e = |out| - reference;  g -= e * (e > g ? attack : decay);  clamp g
// so a stream cannot be split without changing its results.  The parallel axis is the stream: one lane per
// stream, 32 streams per warp.  Streams are rows of the sample matrix, so a warp moves tiles of 32 streams x
// 32 samples through shared memory: coalesced 256-byte row segments on the global side, a conflict-free
// transposed walk (row stride 33 float2) on the recurrence side.  Arithmetic: float32 with one rounding per
// operation in GNU Radio's order (no FMA contraction), IEEE square root -- bit-exact against the oracle.
#pragma once
#include "ofdmx_dev.cuh"

#define AGC_WARPS 4

__global__ void __launch_bounds__(AGC_WARPS * 32)
agc2_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, long long stride, int n_streams,
            float attack, float decay, float reference, float max_gain, float *__restrict__ gain_io, int abs_rate)
{
    __shared__ float2 tile[AGC_WARPS][32][33];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s0 = (blockIdx.x * AGC_WARPS + w) * 32;
    if (s0 >= n_streams) return;
    const int ns = min(32, n_streams - s0);
    float g = (lane < ns) ? gain_io[s0 + lane] : 1.0f;
    const float2 *src = in + (long long)s0 * stride + lane;
    float2 *dst = out + (long long)s0 * stride + lane;
    // the next tile travels in registers while the recurrence walks the current one (the walk is a chain of
    // dependent float operations ~180 cycles per sample long: nothing else hides the global latency)
    float2 nxt[32];
#pragma unroll
    for (int r = 0; r < 32; r++)
        nxt[r] = (r < ns && lane < n) ? __ldcs(src + (long long)r * stride) : make_float2(0.f, 0.f);
    for (long long i0 = 0; i0 < n; i0 += 32) {
        const int nc = (int)min((long long)32, n - i0);
#pragma unroll
        for (int r = 0; r < 32; r++) tile[w][r][lane] = nxt[r];
        __syncwarp();
        const long long i1 = i0 + 32;
#pragma unroll
        for (int r = 0; r < 32; r++)
            if (r < ns && i1 + lane < n) nxt[r] = __ldcs(src + (long long)r * stride + i1);
        if (lane < ns) {
#pragma unroll 4
            for (int k = 0; k < nc; k++) {
                const float2 x = tile[w][lane][k];
                const float re = __fmul_rn(x.x, g), im = __fmul_rn(x.y, g);
                tile[w][lane][k] = make_float2(re, im);
                const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
                const float tmp = __fadd_rn(-reference, mag);
                const float rate = ((abs_rate ? fabsf(tmp) : tmp) > g) ? attack : decay;
                g = __fsub_rn(g, __fmul_rn(tmp, rate));
                if (g < 0.0f) g = 10e-5f;
                if (max_gain > 0.0f && g > max_gain) g = max_gain;
            }
        }
        __syncwarp();
        if (lane < nc) {
#pragma unroll 8
            for (int r = 0; r < ns; r++) __stcs(dst + (long long)r * stride + i0, tile[w][r][lane]);
        }
        __syncwarp();
    }
    if (lane < ns) gain_io[s0 + lane] = g;
}

// ---------------------------------------------------------------------------------------------------------
// agc2_cc on FEW streams: time-parallel spans with a bit-exact acceptance test.
//
// The gain recurrence is contractive while the loop tracks a signal (two gain trajectories fed the same samples
// approach each other by a factor 1 - rate |x| per sample), so in float32 a trajectory started `warm` samples early
// from a guessed gain normally becomes IDENTICAL, bit for bit, to the true one before its span begins.  "Normally" is
// not a proof, so it is checked: span k is accepted only if the gain it had when it entered its span equals the exit
// gain of span k-1 bit for bit -- identical state and identical input give an identical future, so an accepted chain
// from the exact first span IS the sequential result.  Spans that fail are re-run from their predecessor's exit gain
// (agc2_span_kernel with repair = 1), a few rounds of that are enqueued, and what is still open afterwards is walked
// sequentially (agc2_mopup_kernel) -- so the output always equals the sequential recurrence; only the time varies.
//   row r = (stream s, span k): samples [k*span - (k ? warm : 0), min(n, (k+1)*span)), output written from k*span on.
//   entry[r] / exitg[r]: gain before the first / after the last sample of the span; need[r]: row has to be (re)run.
// Same 32 x 32 tile transposes as agc2_kernel (rows are spans of one stream, `span` samples apart).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float agc2_step(float2 &x, float g, float attack, float decay, float reference, float max_gain,
                                           int abs_rate)
{
    const float re = __fmul_rn(x.x, g), im = __fmul_rn(x.y, g);
    x = make_float2(re, im);
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
    const float tmp = __fadd_rn(-reference, mag);
    const float rate = ((abs_rate ? fabsf(tmp) : tmp) > g) ? attack : decay;
    g = __fsub_rn(g, __fmul_rn(tmp, rate));
    if (g < 0.0f) g = 10e-5f;
    if (max_gain > 0.0f && g > max_gain) g = max_gain;
    return g;
}

// A warp takes 32 CONSECUTIVE spans of ONE stream (warps per stream = ceil(spans / 32)) and gives them all the same
// warm-up W (the largest any of its rows asks for, a multiple of 32): row rr of a tile is then the fixed distance
// rr * span from row 0, so loads and stores are plain pointer arithmetic, every lane walks the same iteration range
// [0, W + span), the entry gain is captured and the output starts at iteration W for everybody.  Lanes whose warm-up
// would begin before the stream (k * span < W; the first span in particular) simply start walking later, at sample
// 0 with the gain the call was entered with -- which is exact, not a guess.  The walk itself is the warp's serial
// time (few streams = few warps per SM), so the next tile's loads are in flight while the current tile is walked.
__global__ void __launch_bounds__(AGC_WARPS * 32)
agc2_span_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, long long stride, int n_streams,
                 int spans, long long span, int warm, float attack, float decay, float reference, float max_gain,
                 const float *__restrict__ gain_io, float guess, float *__restrict__ entry, float *__restrict__ exitg,
                 int *__restrict__ need, int repair, int abs_rate)
{
    __shared__ float2 tile[AGC_WARPS][32][33];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wps = (spans + 31) >> 5;                              // warps per stream
    const long long gwarp = (long long)blockIdx.x * AGC_WARPS + w;
    if (gwarp >= (long long)n_streams * wps) return;
    const int s = (int)(gwarp / wps), kbase = (int)(gwarp - (long long)s * wps) * 32;
    const int k = kbase + lane;
    const bool live = k < spans;
    const long long row = (long long)s * spans + k;
    bool run = live;
    if (repair) run = live && k > 0 && need[row] != 0;
    const unsigned runmask = __ballot_sync(0xffffffffu, run);
    if (!runmask) return;
    const float2 *ins = in + (long long)s * stride;
    float2 *outs = out + (long long)s * stride;
    const long long a = (long long)k * span;                        // first sample of this lane's span
    // warm-up: the loop forgets its state at 1 - decay |x| per sample, so ~24 time constants 1 / (decay E|x|) take a
    // gain error of order one below float resolution (plus a margin for the last bits to coincide); E|x| is sampled
    // in front of the span.  Capped at `warm`: a row whose loop barely contracts (weak signal,
    // zero gap) will fail the acceptance test and be repaired instead.
    int W = 0;
    if (!repair) {
        float want = 0.f;
        if (run && k > 0) {
            // 64 samples spread over the (up to) 8192 samples in front of the span: what contracts the warm-up is the
            // mean level over its whole length, and a short window that happens to sit on a quiet stretch of a frame made
            // one row in most warps ask for three times the warm-up of the others (the warp walks the maximum)
            const float2 *q = ins + a;
            float m = 0.f;
            const int st = (int)max(4LL, min(a, 8192LL) / 64);
            const int cnt = (int)min(64LL, a / st);
            for (int i = 1; i <= cnt; i++) { const float2 v = q[-(long long)st * i]; m += fabsf(v.x) + fabsf(v.y); }
            m = (cnt > 0) ? m * 0.75f / (float)cnt : 0.f;           // |re| + |im| ~ 1.3 |x|
            const float tau = 1.0f / fmaxf(decay * m, 1e-9f);
            want = fminf(24.0f * tau + 1536.0f, (float)warm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) want = fmaxf(want, __shfl_xor_sync(0xffffffffu, want, o));
        W = ((int)want + 31) & ~31;
    }
    // this lane walks iterations [start, stop): iteration it is sample a - W + it of the stream
    const int start = (int)max(0LL, (long long)W - a);
    const int stop = run ? W + (int)min(span, n - a) : 0;
    const int total = W + (int)span;                                // warp-uniform iteration count (multiple-of-32 tiles)
    float g = 1.0f;
    if (run) g = repair ? exitg[row - 1] : gain_io[s];
    if (run && repair) entry[row] = g;
    (void)guess;
    const long long base0 = (long long)kbase * span - W;            // stream index of row 0, iteration 0
    float2 v[32];
    auto load_tile = [&](int it0) {
        const long long g0 = base0 + it0 + lane;
#pragma unroll
        for (int rr = 0; rr < 32; rr++) {
            const long long gi = g0 + (long long)rr * span;
            v[rr] = make_float2(0.f, 0.f);
            if (((runmask >> rr) & 1u) && gi >= 0 && gi < n) v[rr] = __ldcs(ins + gi);
        }
    };
    load_tile(0);
    for (int it0 = 0; it0 < total; it0 += 32) {
#pragma unroll
        for (int rr = 0; rr < 32; rr++) tile[w][rr][lane] = v[rr];
        __syncwarp();
        if (it0 + 32 < total) load_tile(it0 + 32);
        if (it0 + 32 > start && it0 < stop) {
            const int q0 = max(0, start - it0), q1 = min(32, stop - it0);
            if (!repair && it0 <= W && W < it0 + 32 && W >= start && W < stop) {
                // the tile in which the span proper begins: capture the entry gain in front of sample W
                for (int q = q0; q < q1; q++) {
                    if (it0 + q == W) entry[row] = g;
                    float2 x = tile[w][lane][q];
                    g = agc2_step(x, g, attack, decay, reference, max_gain, abs_rate);
                    tile[w][lane][q] = x;
                }
            } else {
#pragma unroll 4
                for (int q = q0; q < q1; q++) {
                    float2 x = tile[w][lane][q];
                    g = agc2_step(x, g, attack, decay, reference, max_gain, abs_rate);
                    tile[w][lane][q] = x;
                }
            }
        }
        __syncwarp();
        if (it0 + 32 > W) {                                          // output starts at iteration W
            const long long g0 = base0 + it0 + lane;
            const bool col = it0 + lane >= W;
#pragma unroll 4
            for (int rr = 0; rr < 32; rr++) {
                const long long gi = g0 + (long long)rr * span;
                if (col && ((runmask >> rr) & 1u) && gi < n && kbase + rr < spans) __stcs(outs + gi, tile[w][rr][lane]);
            }
        }
        __syncwarp();
    }
    if (run) {
        if (!repair && W >= stop) entry[row] = g;                   // empty span (cannot happen for k < spans; kept for safety)
        exitg[row] = g;
    }
}

// need[r] = the span's entry gain differs (bitwise) from its predecessor's exit gain; counts them per stream.
__global__ void agc2_verify_kernel(int n_streams, int spans, const float *__restrict__ entry, const float *__restrict__ exitg,
                                   int *__restrict__ need, int *__restrict__ n_open)
{
    const long long rows = (long long)n_streams * spans;
    int cnt = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(r % spans);
        int nd = 0;
        if (k > 0) nd = __float_as_uint(entry[r]) != __float_as_uint(exitg[r - 1]);
        need[r] = nd;
        cnt += nd;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_open, cnt);
}

// What is still open after the repair rounds: one lane per stream walks from the first open span, span by span, from
// its predecessor's exit gain, and stops as soon as a span's new exit gain equals the stored one with the rest of the
// chain consistent.  Plain sequential agc2 from there: slow, exact, and normally not needed (n_open == 0 -> return).
__global__ void agc2_mopup_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, long long stride,
                                  int n_streams, int spans, long long span, float attack, float decay, float reference,
                                  float max_gain, float *__restrict__ entry, float *__restrict__ exitg,
                                  const int *__restrict__ n_open, int abs_rate)
{
    if (*n_open == 0) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const long long r0 = (long long)s * spans;
    for (int k = 1; k < spans; k++) {
        if (__float_as_uint(entry[r0 + k]) == __float_as_uint(exitg[r0 + k - 1])) continue;
        float g = exitg[r0 + k - 1];
        entry[r0 + k] = g;
        const long long a = (long long)k * span, e = min(n, a + span);
        const float2 *src = in + (long long)s * stride;
        float2 *dst = out + (long long)s * stride;
        for (long long i = a; i < e; i++) {
            float2 x = src[i];
            g = agc2_step(x, g, attack, decay, reference, max_gain, abs_rate);
            dst[i] = x;
        }
        exitg[r0 + k] = g;
    }
}

// final gains back to gain_io
__global__ void agc2_final_kernel(int n_streams, int spans, const float *__restrict__ exitg, float *__restrict__ gain_io)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_streams) gain_io[s] = exitg[(long long)s * spans + spans - 1];
}

// ---------------------------------------------------------------------------------------------------------
// filter.iir_filter_ccd(fftaps, fbtaps, oldstyle=False): the out-of-band TX filter of ofdm_radio_hier
// (python/ofdm_radio_hier.py:83-84,93,232-237; python/sync_radio_hier.py:73,165) -- SURVEY.md 8(f) rank 2.
// [UPSTREAM gr-filter iir_filter<gr_complex, gr_complex, double, gr_complexd>::filter]:
//     acc  = ff[0]*x[n];  acc += ff[i]*x[n-i] (i = 1..);  acc += (-fb[j])*y[n-j] (j = 1..);  y[n] = acc
// accumulated in complex double (real taps: the two components are independent real recurrences), output
// rounded to complex float.  A linear recurrence, so unlike the AGC it CAN be cut in time: a lane that starts
// from a zero state `warm` samples before its span reproduces the sequential result up to the impulse
// response left after `warm` samples; the host picks `warm` where that tail is below 1e-18 of its peak, i.e.
// below the rounding of the double accumulator.  Lane (s, c) owns span c of `span` samples of stream s, runs
// the recurrence from max(0, c*span - warm) (from the carried state when that is sample 0) and stores only its
// own span.  The first span of a stream is therefore bit-exact against the sequential filter; the others
// agree to the round-off noise floor of the direct-form recurrence itself (the reference taps cluster eight
// poles at radius <= 0.985 next to z = -1: ~1e-10 relative in the double accumulator, measured: every float
// within one ulp, 99.4 % identical).  Same 32 x 32 shared-memory
// transpose as the AGC: rows are lanes, coalesced 256-byte row segments on the global side.
// The kernel is a template of the tap capacity IIR_MAXT (9: the 8th-order filter of ofdm_radio_hier; 13: the
// 12th-order one of sync_radio_hier, python/sync_radio_hier.py:66-67; 17) because the histories live in registers.
// state_io[n_streams][IIR_STATE_DOUBLES] doubles: x[n-1..] (re,im pairs, IIR_MAXT-1 of them) then y[n-1..].
#define IIR_WARPS 2
#define IIR_STATE_DOUBLES 64
template <int IIR_MAXT>
struct iir_taps { double ff[IIR_MAXT]; double fb[IIR_MAXT]; };  // fb already negated, fb[0] unused; zero padded

// Tiles travel global -> shared with 8-byte cp.async (no staging registers: the 32 history doubles and the
// accumulators fill the register file), double buffered so the next tile is in flight while the recurrence
// walks the current one.
template <int IIR_MAXT>
__global__ void __launch_bounds__(IIR_WARPS * 32)
iir_ccd_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, long long stride, int n_streams,
               int spans_per_stream, long long span, long long warm, const iir_taps<IIR_MAXT> taps,
               double *__restrict__ state_io)
{
    __shared__ float2 tile[IIR_WARPS][2][32][33];
    __shared__ long long row_off[IIR_WARPS][32];      // element offset of the first sample a row filters
    __shared__ int row_len[IIR_WARPS][32], row_skip[IIR_WARPS][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)n_streams * spans_per_stream;
    const long long l0 = ((long long)blockIdx.x * IIR_WARPS + w) * 32;
    if (l0 >= total) return;
    const int nl = (int)min((long long)32, total - l0);
    const long long me = l0 + min(lane, nl - 1);
    const int ms = (int)(me / spans_per_stream);
    const int mc = (int)(me - (long long)ms * spans_per_stream);
    const long long own0 = (long long)mc * span;                 // first sample I store
    const long long beg = max((long long)0, own0 - warm);        // first sample I filter
    const long long end = min(n, own0 + span);                   // one past my last sample
    const int len = (int)(end - beg);
    row_off[w][lane] = (long long)ms * stride + beg;
    row_len[w][lane] = lane < nl ? len : 0;
    row_skip[w][lane] = (int)(own0 - beg);
    double xr[IIR_MAXT - 1], xi[IIR_MAXT - 1], yr[IIR_MAXT - 1], yi[IIR_MAXT - 1];
#pragma unroll
    for (int k = 0; k < IIR_MAXT - 1; k++) { xr[k] = xi[k] = yr[k] = yi[k] = 0.0; }
    if (beg == 0) {
        const double *st = state_io + (long long)ms * IIR_STATE_DOUBLES;
#pragma unroll
        for (int k = 0; k < IIR_MAXT - 1; k++) {
            xr[k] = st[2 * k]; xi[k] = st[2 * k + 1];
            yr[k] = st[2 * (IIR_MAXT - 1) + 2 * k]; yi[k] = st[2 * (IIR_MAXT - 1) + 2 * k + 1];
        }
    }
    int maxlen = lane < nl ? len : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    __syncwarp();
    const uint32_t tile0 = (uint32_t)__cvta_generic_to_shared(&tile[w][0][0][0]);
    auto fill = [&](int buf, int i0) {
        if (i0 < maxlen) {
#pragma unroll
            for (int r = 0; r < 32; r++)
                if (i0 + lane < row_len[w][r])
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile0 + (uint32_t)(((buf * 32 + r) * 33 + lane) * 8)),
                                 "l"(in + row_off[w][r] + i0 + lane) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fill(0, 0);
    for (int i0 = 0, buf = 0; i0 < maxlen; i0 += 32, buf ^= 1) {
        fill(buf ^ 1, i0 + 32);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const int nc = max(0, min(32, row_len[w][lane] - i0));
#pragma unroll 8
        for (int k = 0; k < nc; k++) {
            const float2 x = tile[w][buf][lane][k];
            const double dx = (double)x.x, dy = (double)x.y;
            double ar = __dmul_rn(taps.ff[0], dx), ai = __dmul_rn(taps.ff[0], dy);
#pragma unroll
            for (int t = 1; t < IIR_MAXT; t++) {
                ar = __dadd_rn(ar, __dmul_rn(taps.ff[t], xr[t - 1]));
                ai = __dadd_rn(ai, __dmul_rn(taps.ff[t], xi[t - 1]));
            }
#pragma unroll
            for (int t = 1; t < IIR_MAXT; t++) {
                ar = __dadd_rn(ar, __dmul_rn(taps.fb[t], yr[t - 1]));
                ai = __dadd_rn(ai, __dmul_rn(taps.fb[t], yi[t - 1]));
            }
#pragma unroll
            for (int t = IIR_MAXT - 2; t > 0; t--) {
                xr[t] = xr[t - 1]; xi[t] = xi[t - 1]; yr[t] = yr[t - 1]; yi[t] = yi[t - 1];
            }
            xr[0] = dx; xi[0] = dy; yr[0] = ar; yi[0] = ai;
            tile[w][buf][lane][k] = make_float2((float)ar, (float)ai);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r++)
            if (i0 + lane < row_len[w][r] && i0 + lane >= row_skip[w][r])
                __stcs(out + row_off[w][r] + i0 + lane, tile[w][buf][r][lane]);
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // the lane that filtered the last sample of a stream leaves the history for the next call
    if (lane < nl && end == n) {
        double *st = state_io + (long long)ms * IIR_STATE_DOUBLES;
#pragma unroll
        for (int k = 0; k < IIR_MAXT - 1; k++) {
            st[2 * k] = xr[k]; st[2 * k + 1] = xi[k];
            st[2 * (IIR_MAXT - 1) + 2 * k] = yr[k]; st[2 * (IIR_MAXT - 1) + 2 * k + 1] = yi[k];
        }
    }
}

// PAPR probe of python/papr_sink.py:46-50: peak |x|^2 and sum |x|^2 of a block.
// One partial {sum (double), peak (float)} per CTA; the host wrapper divides.  Grid-stride float4 loads.
__global__ void __launch_bounds__(256)
papr_kernel(const float2 *__restrict__ in, long long n, double *__restrict__ part_sum, float *__restrict__ part_peak)
{
    double s = 0.0;
    float pk = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = __ldcs(in + i);
        const float p = v.x * v.x + v.y * v.y;
        s += (double)p;
        pk = fmaxf(pk, p);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    }
    __shared__ double ss[8];
    __shared__ float sp[8];
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sp[threadIdx.x >> 5] = pk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) { s += ss[k]; pk = fmaxf(pk, sp[k]); }
        part_sum[blockIdx.x] = s;
        part_peak[blockIdx.x] = pk;
    }
}

// out[0] = peak / (sum / n) (the value papr_sink.level() returns), out[1] = peak, out[2] = mean square
__global__ void __launch_bounds__(32)
papr_final_kernel(const double *__restrict__ part_sum, const float *__restrict__ part_peak, int n_parts, long long n,
                  float *__restrict__ out)
{
    double s = 0.0;
    float pk = 0.f;
    for (int i = threadIdx.x; i < n_parts; i += 32) { s += part_sum[i]; pk = fmaxf(pk, part_peak[i]); }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    }
    if (threadIdx.x == 0) {
        const double ms = s / (double)n;
        out[0] = (float)((double)pk / ms);
        out[1] = pk;
        out[2] = (float)ms;
    }
}

// ofdmx_symbol_small.cuh -- one OFDM symbol for the warp-per-frame receiver, fft_len 64 .. 512.
//
// A lane owns R = fft_len/32 samples of the symbol (m = 32 a + lane).  With k = k1 + R k2:
//     X[k1 + R k2] = sum_b [ (sum_a x[32 a + b] W_R^(a k1)) W_N^(b k1) ] W_32^(b k2)
// pass 1 is an R-point FFT in registers (generated, fft_small_gen.cuh), pass 2 a 32-point FFT ACROSS THE LANES by
// butterfly shuffles (five stages, decimation in frequency, one per register slot k1).  The result is written to
// shared memory in natural bin order, like f1k_symbol does for fft_len 1024, so the equaliser code is shared.
#pragma once
#include "ofdmx_frame1024.cuh"
#include "fft_small_gen.cuh"

template <int R> __device__ __forceinline__ void fftR_fwd(float2 (&v)[R]);
template <> __device__ __forceinline__ void fftR_fwd<2>(float2 (&v)[2]) { fft2_fwd(v); }
template <> __device__ __forceinline__ void fftR_fwd<4>(float2 (&v)[4]) { fft4_fwd(v); }
template <> __device__ __forceinline__ void fftR_fwd<8>(float2 (&v)[8]) { fft8_fwd(v); }
template <> __device__ __forceinline__ void fftR_fwd<16>(float2 (&v)[16]) { fft16_fwd(v); }

template <int R> __device__ __forceinline__ constexpr int brevR(int x)
{
    int r = 0;
    for (int b = 1, o = R >> 1; b < R; b <<= 1, o >>= 1)
        if (x & b) r |= o;
    return r;
}

// per-lane factors of the lane-FFT stages with span 16, 8, 4: the upper lane of a butterfly multiplies its difference
// by W_{2 span}^(lane mod span), the lower lane keeps its sum -- stored as a multiplier per lane (1 for lower lanes),
// so that every stage is "shuffle, one FMA per component with a per-lane sign, one complex multiply" without selects
// (spans 2 and 1 have trivial twiddles)
struct LaneTw { float2 w16, w8, w4; };

__device__ __forceinline__ LaneTw lane_twiddles(int lane)
{
    LaneTw t;
    float sn, cs;
    sincospif(-(float)(lane & 15) * (1.0f / 16.0f), &sn, &cs); t.w16 = (lane & 16) ? make_float2(cs, sn) : make_float2(1.f, 0.f);
    sincospif(-(float)(lane & 7) * (1.0f / 8.0f), &sn, &cs); t.w8 = (lane & 8) ? make_float2(cs, sn) : make_float2(1.f, 0.f);
    sincospif(-(float)(lane & 3) * (1.0f / 4.0f), &sn, &cs); t.w4 = (lane & 4) ? make_float2(cs, sn) : make_float2(1.f, 0.f);
    return t;
}

// 32-point forward DFT across the lanes: in: z (lane b), out: lane l holds X[brev5(l)]
__device__ __forceinline__ float2 lane_fft32(float2 z, int lane, const LaneTw &tw)
{
#pragma unroll
    for (int span = 16; span >= 1; span >>= 1) {
        const float tx = __shfl_xor_sync(0xffffffffu, z.x, span), ty = __shfl_xor_sync(0xffffffffu, z.y, span);
        // lower lane keeps a + b, upper lane gets (a - b) w, where a is the lower lane's value: other + sg * own
        const float sg = (lane & span) ? -1.0f : 1.0f;
        float2 s = make_float2(fmaf(sg, z.x, tx), fmaf(sg, z.y, ty));
        if (span == 16) s = cmul(s, tw.w16);
        else if (span == 8) s = cmul(s, tw.w8);
        else if (span == 4) s = cmul(s, tw.w4);
        else if (span == 2) s = ((lane & 3) == 3) ? make_float2(s.y, -s.x) : s;          // W_4^1 = -j
        z = s;
    }
    return z;
}

// One symbol: load + derotate + NFFT-point FFT.  Result: Y[k] = X[k], natural order, k < NFFT.
// tws[k1 * 32 + b] = W_NFFT^(b k1), k1 < R.
// phs: the derotation phasor of this lane's first sample, carried from symbol to symbol -- recomputed exactly when
// `fresh` (first symbol of a frame and every 16th after it), else advanced by sD = one symbol (D samples) of NCO
// rotation; at 2 .. 16 samples per lane the per-symbol sincospi was a tenth of the symbol's instructions.
template <int NFFT>
__device__ __forceinline__ void fsmall_symbol(const KP &p, const float2 *__restrict__ r, long long n, long long i0,
                                              long long t, double kappa, float2 st, bool slow, int j, int jend,
                                              const long long *__restrict__ trig, const float *__restrict__ cfo,
                                              float2 *__restrict__ Y, const float2 *__restrict__ tws, int lane,
                                              const LaneTw &ltw, float2 &phs, float2 sD, bool fresh)
{
    constexpr int R = NFFT / 32;
    float2 v[R];
    const long long s0 = i0 - p.D;
    if (s0 >= 0 && s0 + NFFT <= n) {                  // whole symbol inside the stream (warp-uniform)
        const float2 *src = r + s0 + lane;
#pragma unroll
        for (int a = 0; a < R; a++) v[a] = f1k_ld_stream(src + 32 * a);   // no L1 allocation: the small tables stay resident
    } else {
#pragma unroll
        for (int a = 0; a < R; a++) {
            const long long s = s0 + lane + 32 * a;
            v[a] = (s >= 0 && s < n) ? __ldg(&r[s]) : make_float2(0.f, 0.f);
        }
    }
    {
        if (fresh) {
            double tb = kappa * (double)(i0 + lane - t + 1);
            tb -= rint(tb);
            float sn, cs;
            sincospif(2.0f * (float)tb, &sn, &cs);
            phs = make_float2(cs, sn);
        } else {
            phs = cmul(phs, sD);
        }
        float2 ph = phs;
        const long long tnx = (slow && j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
#pragma unroll
        for (int a = 0; a < R; a++) {
            const long long i = i0 + lane + 32 * a;
            float2 pa = ph;
            if (i >= tnx) {      // sample-and-hold value changed inside the frame: exact piecewise phase
                double turns = nco_turns(i, j, jend, trig, cfo, NFFT);
                turns -= rint(turns);
                float s2, c2;
                sincospif(2.0f * (float)turns, &s2, &c2);
                pa = make_float2(c2, s2);
            }
            v[a] = cmul(v[a], pa);
            if (a + 1 < R) ph = cmul(ph, st);
        }
    }
    fftR_fwd<R>(v);                                   // v[brevR(k1)] = y_b[k1]
    const int k2 = brev5(lane);
#pragma unroll
    for (int q = 0; q < R; q++) {
        const int k1 = brevR<R>(q);
        const float2 z = (k1 == 0) ? v[q] : cmul(v[q], tws[k1 * 32 + lane]);
        Y[k1 + R * k2] = lane_fft32(z, lane, ltw);
    }
}

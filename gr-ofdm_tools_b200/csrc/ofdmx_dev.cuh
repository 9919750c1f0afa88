// ofdmx_dev.cuh -- device-side building blocks shared by the OFDM PHY kernels (sm_100a).
//
// Block semantics follow GNU Radio 3.7 as wired by the reference;
// each helper names the block it replaces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define OFDMX_THREADS 256
#define OFDMX_MAX_PTS 64

// Kernel parameter block, passed by value (lives in the constant bank).
struct KP {
    int N, logN, cp, D, hl;
    int n_occ_sets, n_pil_sets, n_pil_sym_sets, n_occ_u, n_cv;
    int bps_h, bps_p, crc_mode, gneg, gpos, holdoff;
    int max_pkt_bytes, max_pkt_syms;
    double thr;
    float alpha, tx_scale, tx_clip;
    int roll;                    // ofdm_cyclic_prefixer rolloff_len (0 = rectangular, else >= 2)
    const float *roll_flank;     // [2 * (roll - 1)]: up flank, then down flank
    const float2 *tw;            // [N] exp(-2 pi i k / N)
    const int *occ_bins;         // flat, set-major, list order, shifted bins
    const int *occ_base;         // [n_occ_sets]
    const int *occ_size;         // [n_occ_sets]
    const int *occ_u;            // [n_occ_u] union of occupied bins (what the equaliser visits)
    const uint8_t *pil_flag;     // [max(1,n_pil_sets)][N]
    const float2 *pil_val;       // [max(1,n_pil_sets)][N]
    const int *pil_bins;         // TX: flat pilot bins, set-major
    const int *pil_base;
    const int *pil_size;
    const float2 *pil_sym;       // TX: flat pilot symbols, set-major
    const int *pil_sym_base;
    const float2 *sw1;           // [N] shifted
    const float2 *sw2;
    const int *cv_k;             // [n_cv] bins where sw2/sw1 is defined and nonzero
    const float2 *cv_conj;       // [n_cv] conj(sw2[k]/sw1[k])
    const float2 *inv_sw2;       // [N] 1/sw2[k] or 0
    const uint8_t *hdr_mask;     // [hl]
    const uint8_t *keystream;    // [max_pkt_bytes + 8]
    const uint32_t *crc_tab;     // [256] reflected CRC-32 table
    const uint32_t *crc_pow;     // [256] x^(128*(255-j)) mod P
    const float2 *hpts;          // [2^bps_h] header constellation points
    const float2 *ppts;          // [2^bps_p] payload constellation points
    const uint8_t *lut_h;        // constellation_rect sector LUTs (QAM only)
    const uint8_t *lut_p;
    const float2 *inv_hpts;      // 1 / constellation point
    const float2 *inv_ppts;
    const int *pos_su;           // [n_occ_sets][n_occ_u] position of union carrier u in set's list, or -1
    int max_frame_syms;
    const uint8_t *crc8_bit;     // [24] CRC-8 contribution of header bit l (len bits 0..11, counter bits 12..23)
    unsigned crc8_zero;          // CRC-8 of the all-zero header fields
    const uint32_t *crc_pow64;   // [32] x^(512*(31-l)) mod P (warp CRC)
    const uint32_t *crc_pow8;    // [65] x^(8*t) mod P (warp CRC tail shift)
    int y1_lo, y1_span;          // shifted-bin range of sync symbol 1 that the offset search reads
    int pil_in_occ;              // some pilot carrier is also in occupied_carriers (equaliser pilot branch reachable)
    int nsw;                     // sync words in front of the header symbol: 2, or 1 (sync_word2=(): ofdm_chanest_vcvc
                                 // estimates the carrier offset by correlating |Y1[k]-Y1[k+2]|^2 with the known
                                 // differences (cv_k / cv_conj.x) and takes the taps from sync word 1, interpolated)
    int interp, first_act, last_act;   // single-word mode: d_interpolate and the active range of ofdm_chanest_vcvc
    float2 *h_taps;              // debug tap (ofdmx_set_debug_taps): the channel taps ofdm_chanest_vcvc hands to the
    long long h_stride;          // equaliser (tag ofdm_sync_chan_taps), fft_len complex per trigger slot, shifted order
    float qiw_h, qiw_p;          // constellation_rect: 1 / sector width of the header / payload QAM table
                                 // (0.5 (side - 1) without normalisation; divided by the scale factor otherwise)
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)   // a * conj(b)
{
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cdivf(float2 a, float2 b)
{
    float d = b.x * b.x + b.y * b.y;
    float2 n = cmul_conj(a, b);
    return make_float2(n.x / d, n.y / d);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// decision_maker of constellation_bpsk / _qpsk / _8psk / constellation_rect (16-, 64-QAM)
__device__ __forceinline__ int ofdm_decide(int bps, float re, float im, const uint8_t *lut, float inv_w)
{
    if (bps == 1) return re > 0.f;
    if (bps == 2) return 2 * (im > 0.f) + (re > 0.f);
    if (bps == 3) {
        int r = (fabsf(re) <= fabsf(im)) ? 4 : 0;
        if (re <= 0.f) r |= 1;
        if (im <= 0.f) r |= 2;
        return r;
    }
    const int side = (bps == 4) ? 4 : 8;
    int rs = __float2int_rz(fmaf(re, inv_w, 0.5f * side));
    int is = __float2int_rz(fmaf(im, inv_w, 0.5f * side));
    rs = min(max(rs, 0), side - 1);
    is = min(max(is, 0), side - 1);
    return lut[rs * side + is];
}

// ---------------------------------------------------------------------------------------------
// fft_vcc core: in-place shared-memory FFT, input in bit-reversed order, output natural order.
// Radix-4 passes (two merged radix-2 DIT stages), one radix-2 pass first when log2 N is odd.
// All threads of the block must call it; buf holds N float2.
template <bool INVERSE>
__device__ __forceinline__ void fft_smem(float2 *buf, int N, int logN, const float2 *__restrict__ tw)
{
    int s = 0;
    if (logN & 1) {
        for (int b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
            float2 a = buf[2 * b], c = buf[2 * b + 1];
            buf[2 * b] = cadd(a, c);
            buf[2 * b + 1] = csub(a, c);
        }
        __syncthreads();
        s = 1;
    }
    for (; s < logN; s += 2) {
        const int h = 1 << s;
        for (int b = threadIdx.x; b < (N >> 2); b += blockDim.x) {
            const int j = b & (h - 1);
            const int base = ((b >> s) << (s + 2)) + j;
            float2 x0 = buf[base], x1 = buf[base + h], x2 = buf[base + 2 * h], x3 = buf[base + 3 * h];
            float2 w1 = __ldg(&tw[j * (N >> (s + 1))]);
            float2 w2 = __ldg(&tw[j * (N >> (s + 2))]);
            if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; }
            x1 = cmul(x1, w1);
            x3 = cmul(x3, w1);
            float2 a0 = cadd(x0, x1), a1 = csub(x0, x1), a2 = cadd(x2, x3), a3 = csub(x2, x3);
            a2 = cmul(a2, w2);
            a3 = cmul(a3, w2);
            a3 = INVERSE ? make_float2(-a3.y, a3.x) : make_float2(a3.y, -a3.x);   // * (+/- i)
            buf[base] = cadd(a0, a2);
            buf[base + 2 * h] = csub(a0, a2);
            buf[base + h] = cadd(a1, a3);
            buf[base + 3 * h] = csub(a1, a3);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int bitrev(int v, int logN) { return (int)(__brev((unsigned)v) >> (32 - logN)); }

// ---------------------------------------------------------------------------------------------
// CRC-32 (digital.crc32_bb arithmetic = zlib CRC-32), block-parallel.
// Reflected GF(2) representation: bit 31 is x^0.
__device__ __forceinline__ uint32_t gf2_mul_x(uint32_t b) { return (b & 1u) ? ((b >> 1) ^ 0xEDB88320u) : (b >> 1); }
__device__ __forceinline__ uint32_t gf2_mul(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = gf2_mul_x(b);
    }
    return p;
}

// CRC-32 of msg[0..len) held in shared (or global) memory; all threads of a 256-thread block call
// it (256..480 threads), the result is returned to every thread.  scratch: 16 uint32 in shared memory.
// Method: the message is viewed as the tail of a 4096-byte zero-padded buffer cut into 256
// 16-byte chunks; each thread computes the raw (zero-init) CRC register of its chunk and shifts it
// by x^(8*bytes that follow); the XOR of all terms is the raw CRC.  The 0xFFFFFFFF initial value is
// folded in by XOR-ing the first four message bytes with 0xFF.  Longer messages are processed in
// 4096-byte super-chunks.
static __device__ uint32_t crc32_block(const uint8_t *msg, int len, const uint32_t *__restrict__ tab,
                                const uint32_t *__restrict__ powtab, uint32_t *scratch)
{
    const int tid = threadIdx.x;
    if (len < 4) {
        // init folding needs >= 4 message bytes; tiny messages are done serially
        uint32_t c = 0xFFFFFFFFu;
        for (int i = 0; i < len; i++) c = tab[(c ^ msg[i]) & 0xFF] ^ (c >> 8);
        return c ^ 0xFFFFFFFFu;
    }
    uint32_t reg_total = 0;   // raw register so far (same value on every thread)
    int done = 0;
    int first_chunk_len = len % 4096;
    if (first_chunk_len == 0) first_chunk_len = 4096;
    while (done < len) {
        const int clen = (done == 0) ? first_chunk_len : 4096;
        const int pad = 4096 - clen;                 // leading virtual zeros
        // thread's 16 bytes: virtual positions [tid*16, tid*16+16)
        uint32_t reg = 0;
        for (int q = 0; q < 16 && tid < 256; q++) {
            int vp = tid * 16 + q - pad;
            if (vp < 0) continue;
            int gi = done + vp;
            uint8_t byte = msg[gi];
            if (gi < 4) byte ^= 0xFF;                // init = 0xFFFFFFFF
            reg = tab[(reg ^ byte) & 0xFF] ^ (reg >> 8);
        }
        uint32_t term = reg ? gf2_mul(reg, powtab[tid & 255]) : 0u;
        // XOR-reduce over the block
        for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
        __syncthreads();
        if ((tid & 31) == 0) scratch[tid >> 5] = term;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) t ^= scratch[w];
            // previous register shifted by the 4096 bytes just processed: x^(8*4096) = powtab[0]*x^128
            uint32_t prev = reg_total;
            if (prev) {
                uint32_t sh = powtab[0];
                for (int i = 0; i < 128; i++) sh = gf2_mul_x(sh);
                prev = gf2_mul(prev, sh);
            }
            scratch[15] = prev ^ t;
        }
        __syncthreads();
        reg_total = scratch[15];
        done += clen;
    }
    return reg_total ^ 0xFFFFFFFFu;
}

// CRC-8 of the OFDM header fields (poly 0x07, init 0xFF; packet_header_default)
__device__ __forceinline__ uint8_t crc8_hdr(unsigned len, unsigned num)
{
    uint8_t b[4] = { (uint8_t)(len & 0xFF), (uint8_t)(len >> 8), (uint8_t)(num & 0xFF), (uint8_t)(num >> 8) };
    uint8_t crc = 0xFF;
    for (int i = 0; i < 4; i++) {
        crc ^= b[i];
        for (int k = 0; k < 8; k++) crc = (crc & 0x80) ? (uint8_t)((crc << 1) ^ 0x07) : (uint8_t)(crc << 1);
    }
    return crc;
}

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one int per thread (blockDim.x <= 1024); warp_tot: 33 ints shared
__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < nw) ? warp_tot[lane] : 0;
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) warp_tot[lane] = winc - w;
        if (lane == 31) warp_tot[32] = winc;
    }
    __syncthreads();
    total = warp_tot[32];
    return warp_tot[wid] + inc - v;
}

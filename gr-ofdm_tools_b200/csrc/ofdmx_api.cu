// ofdmx_api.cu -- C ABI (include/ofdmx.h) over the sm_100a kernels.  No CPU fallback: every
// entry point needs a CUDA device and reports OFDMX_ERR_CUDA otherwise.
#define OFDMX_GENERIC_KERNELS 1
#include "ofdmx_kernels.cuh"
#include "ofdmx_sync.cuh"
#include "ofdmx_sync_tma.cuh"
#include "ofdmx_frame1024.cuh"
#include "ofdmx_frame1024w.cuh"
#include "ofdmx_frame2048p.cuh"
#include "ofdmx_cond.cuh"
#include "ofdmx_sync_warp.cuh"
#include "ofdmx_sync_warpn.cuh"
#include "ofdmx_tx1024w.cuh"
#include "ofdmx_chain.cuh"
#include "ofdmx_launch.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

enum KSlot { K_SYNC = 0, K_PLATEAU, K_TRIG_COUNT, K_TRIG_SCAN, K_TRIG_SCATTER, K_CFO, K_FRAME, K_CHAIN_NEXT, K_CHAIN_ENTRY,
             K_CHAIN_MARK, K_CHAIN_SCAN, K_CHAIN_EMIT, K_TX_OFF, K_TX, K_FFT, K_CRC, K_FRAME1K, K_FRAME1KW, K_SYNC_FAST,
             K_SYNC_TMA, K_AGC2, K_SYNC_WARP, K_TX1KW, K_IIR, K_PAPR, K_AGC2_AUX, K_FRAMEP, K_SYNC_WARPN, K_CHAIN_SMALL, K_NSLOTS };
static const char *const kSlotNames[K_NSLOTS] = {
    "(unused)", "plateau_kernel", "(unused)", "trig_scan_kernel", "trig_scatter_kernel",
    "cfo_kernel", "rx_frame_kernel", "chain_next_kernel", "chain_entry_kernel", "chain_mark_kernel",
    "chain_scan_kernel", "chain_emit_kernel", "tx_offsets_kernel", "tx_frame_kernel", "fft_vcc_kernel", "crc32_kernel",
    "rx_frame1024_kernel", "rx_framew_kernel", "sync_metric_fast_kernel", "sync_metric_tma_kernel", "agc2_kernel", "sync_metric_warp_kernel", "tx_framew_kernel", "iir_ccd_kernel", "papr_kernel", "agc2_verify/mopup/final_kernel", "rx_framep_kernel", "sync_metric_warpn_kernel", "chain_small_kernel" };

struct ProfRec { int slot; cudaEvent_t a, b; };

// Everything that is a function of the PHY parameters (ofdmx_params): the kernel parameter block with its table
// pointers and the launch configuration derived from it.  ofdmx_create builds the first one, ofdmx_reconfigure
// builds the next one next to it and swaps (the tables of two consecutive plans live in the two halves of one
// device arena, so work enqueued under the old plan keeps reading valid tables).
struct PlanFields {
    ofdmx_params prm{};             // scalar fields only are used after creation (the pointers are the caller's)
    KP kp{};
    std::vector<int> occ_sizes;
    int hl = 0;
    size_t frame_smem = 0, tx_smem = 0, sync_fast_smem = 0, frame1k_smem = 0, sync_tma_smem = 0;
    int sync_tma_occ = 3;           // resident CTAs per SM of the TMA sync kernel (persistent grid size)
    bool sync_warp_ok = false;
    bool frame1kw = false;          // warp-per-frame kernel eligible
    size_t frame1kw_smem = 0;
    int frame1kw_warps = FW_WARPS;
    int frame1kw_dec_all = 0;       // > 0: bits per OFDM symbol not a byte multiple: decisions of the whole packet kept (capacity)
    int frame1kw_ctas = 1;          // resident CTAs per SM of the warp-per-frame kernel (fft_len < 1024: several)
    bool tx1kw = false;             // warp-per-packet TX kernel usable
    size_t tx1kw_smem = 0;
    const uint16_t *tx_map = nullptr;
    const float2 *sync_td = nullptr;
    int frame1kw_dec_off = -1;      // float2 index inside the symbol buffer where the decisions live (-1: own array)
    uint32_t x_2048 = 0;            // x^(8*2048) mod P
    int frame1k_warps = 0;          // > 0: fft_len 1024 fast path with this many warps per CTA
    bool framep = false;            // fft_len 2048: pair-of-warps-per-frame kernel eligible
    size_t framep_smem = 0;
    int framep_hsz = 0;
    const uint16_t *pair_tab = nullptr;
};

struct ofdmx_ctx : PlanFields {
    bool profiling = false;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[K_NSLOTS] = {0};
    int64_t prof_calls[K_NSLOTS] = {0};
    int64_t prof_errors = 0;        // event create / record failures since ofdmx_profile(ctx, 1)
    int device = 0;
    int sm_count = 148;
    std::string err;
    // table arena: two halves, the plans alternate between them
    DevBuf arena[2];
    int arena_cur = 0;
    static const int STAGE_SLOTS = 8;
    void *stage_pinned = nullptr;   // pinned host staging of the table images: a ring of STAGE_SLOTS slots, so that the
    size_t stage_cap = 0;           // host only ever waits for the upload issued STAGE_SLOTS reconfigurations ago
    int stage_next = 0;
    cudaEvent_t ev_stage[STAGE_SLOTS] = { nullptr };   // upload out of slot i finished
    bool stage_used[STAGE_SLOTS] = { false };
    cudaStream_t cfg_stream = nullptr;   // uploads of table images
    cudaEvent_t ev_tables = nullptr;     // recorded behind the upload of the current plan's tables
    cudaEvent_t ev_done[2] = { nullptr, nullptr };   // end of the work that reads arena half i (taken when the plan is left)
    bool done_valid[2] = { false, false };
    bool tables_pending = false;         // calls make their stream wait for ev_tables until it has completed
    cudaStream_t last_stream = nullptr;  // stream of the most recent call (the context is single-owner)
    bool last_stream_valid = false;
    // workspace
    DevBuf ws;
    DevBuf ws_host;                 // workspace of ofdmx_rx_host (runs on own_stream, concurrently with the caller's stream)
    DevBuf ws_papr;                 // partial sums of ofdmx_papr (may run on another stream than an RX call in flight)
    DevBuf ws_agc;                  // span bookkeeping of ofdmx_agc2 (entry / exit gains, re-run flags)
    bool no_agc_spans = false;      // OFDMX_NO_AGC_SPANS=1: always one lane per stream
    bool no_pair_frame = false;     // OFDMX_NO_PAIR_FRAME=1: fft_len 2048 on the one-warp-per-frame kernel
    float2 *h_taps = nullptr;       // ofdmx_set_debug_taps
    long long h_stride = 0;
    int64_t launches = 0;
    int64_t n_dev_allocs = 0;       // cudaMalloc / cudaHostAlloc calls made by this context
    int64_t n_host_syncs = 0;       // host-blocking synchronisations made by this context
    int64_t n_reconfigs = 0;
    // host-buffer path
    DevBuf h_samples, h_frames, h_bytes, h_counts;
    cudaStream_t own_stream = nullptr;
    bool no_tma = false;            // OFDMX_NO_TMA=1: use the plain-load sync kernel
    bool no_warp_sync = false;      // OFDMX_NO_WARP_SYNC=1: the TMA ring kernel instead of the warp-autonomous ones
    bool emit_all = false;          // ofdmx_set_emit_all: frames_out receives every trigger's record
    bool no_warp_frame = false;     // OFDMX_NO_WARP_FRAME=1: use the CTA-per-frame fft_len-1024 kernel
    bool force_generic = false;     // OFDMX_FORCE_GENERIC=1: always use the any-fft_len frame kernel
    bool no_warp_tx = false;        // OFDMX_NO_WARP_TX=1: generic TX kernel
};

namespace {

int fail(ofdmx_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(ctx, OFDMX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

// Table image under construction: bytes + the pointer fields that have to point into it once it has a device
// address (offsets are 256-byte aligned).
struct Stage {
    std::vector<unsigned char> bytes;
    std::vector<std::pair<const void **, size_t>> fixups;
};

template <typename T>
int upload(Stage &st, const std::vector<T> &v, const T **out)
{
    const size_t off = (st.bytes.size() + 255) / 256 * 256;
    const size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    st.bytes.resize(off + n, 0);
    if (!v.empty()) std::memcpy(st.bytes.data() + off, v.data(), v.size() * sizeof(T));
    st.fixups.emplace_back(reinterpret_cast<const void **>(out), off);
    return 0;
}

int grow(ofdmx_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CUDA_TRY(ctx, cudaMalloc(&b.p, want));
    ctx->n_dev_allocs++;
    b.cap = want;
    return 0;
}

// records a CUDA event pair around one kernel launch when profiling is on
struct KTimer {
    ofdmx_ctx *c;
    cudaStream_t st;
    ProfRec r;
    bool on;
    KTimer(ofdmx_ctx *ctx, int slot, cudaStream_t s) : c(ctx), st(s), on(ctx->profiling)
    {
        ctx->launches++;
        if (!on) return;
        auto get = [&]() {
            cudaEvent_t e = nullptr;
            if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
            return e;
        };
        r.slot = slot; r.a = get(); r.b = get();
        // a failed event creation / record drops this sample (and is reported by ofdmx_profile_read) rather than
        // timing garbage
        if (!r.a || !r.b || cudaEventRecord(r.a, st) != cudaSuccess) { drop(); on = false; }
    }
    void drop()
    {
        if (r.a) c->prof_pool.push_back(r.a);
        if (r.b) c->prof_pool.push_back(r.b);
        c->prof_errors++;
        (void)cudaGetLastError();
    }
    ~KTimer()
    {
        if (!on) return;
        if (cudaEventRecord(r.b, st) != cudaSuccess) { drop(); return; }
        c->prof_pending.push_back(r);
    }
};
#define KT(slot) KTimer kt_(ctx_, slot, st)

inline int shifted_bin(int c, int n)
{
    if (c < 0) c += n;
    return (c + n / 2) % n;
}

// _get_constellation(bps) (python/ofdm_txrx_modules.py:106-118); 6 = 64-QAM by the same qam.py rule
bool make_constellation(int bps, int qam_norm, std::vector<float2> &pts, std::vector<uint8_t> &lut, float *inv_w)
{
    *inv_w = 1.0f;
    pts.clear();
    lut.assign(64, 0);
    if (bps == 1) {
        pts = { make_float2(-1.f, 0.f), make_float2(1.f, 0.f) };
    } else if (bps == 2) {
        const float a = 0.707107f;
        pts = { make_float2(-a, -a), make_float2(a, -a), make_float2(-a, a), make_float2(a, a) };
    } else if (bps == 3) {
        static const int mult[8] = { 1, 7, 15, 9, 3, 5, 13, 11 };
        const float ang = (float)(M_PI / 8.0);
        for (int i = 0; i < 8; i++) pts.push_back(make_float2((float)std::cos(mult[i] * ang), (float)std::sin(mult[i] * ang)));
    } else if (bps == 4 || bps == 6) {
        const int m = 1 << bps, q = (bps == 4) ? 2 : 4;
        const double step = 1.0 / (q - 0.5);
        for (int i = 0; i < m; i++) {
            const int y = i % q, x = (i / q) % q, quad = i / (q * q);
            const double gx = (x + 0.5) * step, gy = (y + 0.5) * step;
            double re, im;
            if (quad == 0) { re = gx; im = gy; }
            else if (quad == 1) { re = -gy; im = gx; }
            else if (quad == 2) { re = -gx; im = -gy; }
            else { re = gy; im = -gx; }
            pts.push_back(make_float2((float)re, (float)im));
        }
        // constellation_rect: sector centre -> closest point
        const int side = 2 * q;
        const double w = 2.0 / (side - 1);
        for (int rs = 0; rs < side; rs++)
            for (int is = 0; is < side; is++) {
                const double cr = (rs + 0.5 - side / 2.0) * w, ci = (is + 0.5 - side / 2.0) * w;
                int best = 0;
                double bd = 1e300;
                for (int i = 0; i < m; i++) {
                    const double dr = cr - pts[i].x, di = ci - pts[i].y, d = dr * dr + di * di;
                    if (d < bd) { bd = d; best = i; }
                }
                lut[rs * side + is] = (uint8_t)best;
            }
        // [UPSTREAM constellation.cc, GNU Radio >= 3.8] constellation_rect(..., AMPLITUDE_NORMALIZATION): the points
        // and the sector widths are scaled by n / sum |p| (3.7, the reference's generation, has no such step)
        double scale = 1.0;
        if (qam_norm == 1) {
            double sum = 0.0;
            for (auto &v : pts) sum += std::sqrt((double)v.x * v.x + (double)v.y * v.y);
            scale = (double)m / sum;
            for (auto &v : pts) v = make_float2((float)(v.x * scale), (float)(v.y * scale));
        }
        *inv_w = (float)(1.0 / (w * scale));
    } else {
        return false;
    }
    return true;
}

// gnuradio/digital/lfsr.h
struct Lfsr {
    uint32_t sr, mask, len;
    Lfsr(uint32_t m, uint32_t seed, uint32_t l) : sr(seed), mask(m), len(l) {}
    unsigned next()
    {
        unsigned out = sr & 1u;
        unsigned nb = (unsigned)__builtin_popcount(sr & mask) & 1u;
        sr = (sr >> 1) | (nb << len);
        return out;
    }
};

uint32_t h_gf2_mul_x(uint32_t b) { return (b & 1u) ? ((b >> 1) ^ 0xEDB88320u) : (b >> 1); }

struct CrcTables {
    std::vector<uint32_t> tab, pow, pow64, pow8;
    uint32_t x_2048;
};
const CrcTables &crc_tables()
{
    static const CrcTables t = [] {
        CrcTables r;
        r.tab.resize(256); r.pow.resize(256); r.pow64.resize(32); r.pow8.resize(65);
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t v = i;
            for (int k = 0; k < 8; k++) v = (v & 1u) ? ((v >> 1) ^ 0xEDB88320u) : (v >> 1);
            r.tab[i] = v;
        }
        uint32_t x = 0x80000000u;   // x^0
        for (int j = 255; j >= 0; j--) {
            r.pow[j] = x;           // x^(128*(255-j))
            for (int b = 0; b < 128; b++) x = h_gf2_mul_x(x);
        }
        x = 0x80000000u;
        for (int l = 31; l >= 0; l--) {
            r.pow64[l] = x;         // x^(512*(31-l))
            for (int b = 0; b < 512; b++) x = h_gf2_mul_x(x);
        }
        r.x_2048 = x;               // after 32 steps of 512 bits: x^16384
        x = 0x80000000u;
        for (int t2 = 0; t2 <= 64; t2++) {
            r.pow8[t2] = x;         // x^(8*t)
            for (int b = 0; b < 8; b++) x = h_gf2_mul_x(x);
        }
        return r;
    }();
    return t;
}

int payload_ofdm_syms(const PlanFields *c, int n_syms)
{
    int cnt = 0, acc = 0, s = 1 % c->prm.n_occ_sets;
    while (acc < n_syms) {
        cnt++;
        acc += c->occ_sizes[s];
        s = (s + 1) % c->prm.n_occ_sets;
    }
    return cnt;
}

// workspace carve-up for RX
struct RxWs {
    uint32_t *detmask, *trigmask;
    int *blocksum, *n_trig, *stream_start, *stream_count, *jumpA, *jumpB, *entry, *blockcount;
    int nblk;
    long long *trig;
    int *trig_stream;
    float *cfo;
    ofdmx_frame *spec;
    uint8_t *markA, *markB;
    size_t total;
    long long wps, n_words;
    int nb;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

RxWs carve(void *base, int64_t n_streams, int64_t n_samples, int64_t max_trig)
{
    RxWs w{};
    w.wps = (n_samples + 31) / 32;
    if (w.wps < 1) w.wps = 1;
    w.n_words = w.wps * n_streams;
    w.nb = (int)((w.n_words + OFDMX_THREADS * TRIG_WPT - 1) / (OFDMX_THREADS * TRIG_WPT));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return base ? static_cast<char *>(base) + o : nullptr;
    };
    w.detmask = (uint32_t *)take(sizeof(uint32_t) * (size_t)w.n_words);
    w.trigmask = (uint32_t *)take(sizeof(uint32_t) * (size_t)w.n_words);
    w.blocksum = (int *)take(sizeof(int) * (size_t)(w.nb + 1));
    w.n_trig = (int *)take(sizeof(int) * 4);
    w.stream_start = (int *)take(sizeof(int) * (size_t)(n_streams + 2));
    w.stream_count = (int *)take(sizeof(int) * (size_t)(n_streams + 2));
    w.nblk = (int)((max_trig + CH_B - 1) / CH_B);
    w.entry = (int *)take(sizeof(int) * (size_t)(w.nblk + 1));
    w.blockcount = (int *)take(sizeof(int) * (size_t)(w.nblk + 1));
    w.jumpA = (int *)take(sizeof(int) * (size_t)(max_trig + 1));
    w.jumpB = (int *)take(sizeof(int) * (size_t)(max_trig + 1));
    w.trig = (long long *)take(sizeof(long long) * (size_t)(max_trig + 1));
    w.trig_stream = (int *)take(sizeof(int) * (size_t)(max_trig + 1));
    w.cfo = (float *)take(sizeof(float) * (size_t)(max_trig + 1));
    w.spec = (ofdmx_frame *)take(sizeof(ofdmx_frame) * (size_t)(max_trig + 1));
    w.markA = (uint8_t *)take((size_t)(max_trig + 1));
    w.markB = (uint8_t *)take((size_t)(max_trig + 1));
    w.total = off;
    return w;
}

int check_device(ofdmx_ctx *ctx)
{
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// sample streams as a 3-D tensor {32 floats (16 samples), rows, streams}; false if TMA cannot describe the buffer
bool make_sample_map(CUtensorMap *tmap, const float2 *samples, int64_t n_streams, int64_t n_samples, int64_t stride)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(samples) & 15) != 0) return false;
    if (n_streams > 1 && (stride & 1)) return false;                 // stream stride must be a multiple of 16 bytes
    const int64_t rows = n_samples / 16;
    if (rows < 1 || rows > 0x7fffffffLL - 4096) return false;
    cuuint64_t dims[3] = { 32, (cuuint64_t)rows, (cuuint64_t)n_streams };
    cuuint64_t strides[2] = { 128, (cuuint64_t)stride * 8 };
    if (n_streams == 1) strides[1] = (cuuint64_t)((rows * 128 + 15) / 16 * 16);
    cuuint32_t box[3] = { 32, ST_BOX_ROWS, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    return fn(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2 *>(samples), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Spans of the warp-autonomous sync kernels: the warps take spans round-robin (static), so the spans of a stream are
// sized to come out even -- as many per stream as fit four rounds of all resident warps, all of (almost) the same
// length -- instead of a fixed length that leaves a short last span per stream and a ragged last round.
static void plan_spans(long long tiles, long long n_streams, long long warps, long long min_span, long long &span, long long &spans)
{
    const long long per = std::max<long long>(1, (4 * warps) / std::max<long long>(1, n_streams));
    span = (tiles + per - 1) / per;
    span = std::max<long long>(min_span, std::min<long long>(span, 4096));
    spans = (tiles + span - 1) / span;
    span = (tiles + spans - 1) / spans;                     // equalise: the same number of spans, none much shorter
}

// sync front end shared by ofdmx_rx and ofdmx_sync: detect bits -> triggers -> cfo
int run_sync(ofdmx_ctx *ctx, const RxWs &w, const float2 *samples, int64_t n_streams, int64_t n_samples,
             int64_t stride, int64_t max_trig, ofdmx_counts *counts_dev, cudaStream_t st)
{
    const KP &kp = ctx->kp;
    ofdmx_ctx *ctx_ = ctx;
    CUDA_TRY(ctx, cudaMemsetAsync(w.blocksum, 0, sizeof(int) * (size_t)(w.nb + 1), st));   // per-block trigger counts (plateau_kernel)
    CUtensorMap tmap;
    if ((kp.N == 1024 || kp.N == 2048) && ctx->sync_warp_ok && !ctx->no_warp_sync && !ctx->no_tma && (reinterpret_cast<uintptr_t>(samples) & 15) == 0
        && (n_streams == 1 || (stride & 1) == 0) && (n_samples / SW_TILE + 2) * n_streams < 0x7fffffffLL) {
        // fft_len 1024 / 2048: warp-autonomous streaming (no block barriers), spans of tiles of fft_len/2 samples per warp
        const int C = kp.N / 64, wpc = SW_WARPS * 16 / C;                  // chunk size, warps per CTA
        const long long tile = 32LL * C;
        const long long tiles = (n_samples + tile - 1) / tile;
        const long long warps = (long long)ctx->sm_count * wpc;
        long long span, spans;
        plan_spans(tiles, n_streams, warps, 16, span, spans);
        const long long total = spans * n_streams;
        const unsigned grid = (unsigned)std::min<long long>((total + wpc - 1) / wpc, (long long)ctx->sm_count);
        KT(K_SYNC_WARP);
        if (C == 16)
            sync_metric_warp_kernel<16><<<grid, wpc * 32, SW_SMEM_BYTES, st>>>(samples, n_samples, stride, (float)kp.thr, kp.thr,
                                                                              w.detmask, w.trigmask, w.wps, (int)tiles, (int)span, (int)spans, (int)total);
        else
            sync_metric_warp_kernel<32><<<grid, wpc * 32, SW_SMEM_BYTES, st>>>(samples, n_samples, stride, (float)kp.thr, kp.thr,
                                                                              w.detmask, w.trigmask, w.wps, (int)tiles, (int)span, (int)spans, (int)total);
    } else if (kp.N <= 512 && ctx->sync_warp_ok && !ctx->no_warp_sync && !ctx->no_tma && (reinterpret_cast<uintptr_t>(samples) & 15) == 0
               && (n_streams == 1 || (stride & 1) == 0) && (n_samples / SW_TILE + 2) * n_streams < 0x7fffffffLL) {
        // fft_len 32 .. 512: the short-window warp-autonomous kernel (one warm-up tile per span)
        const long long tiles = (n_samples + SW_TILE - 1) / SW_TILE;
        const long long warps = (long long)ctx->sm_count * SW_WARPS;
        long long span, spans;
        plan_spans(tiles, n_streams, warps, 4, span, spans);     // short spans only when the input is small: latency of small calls
        const long long total = spans * n_streams;
        const unsigned grid = (unsigned)std::min<long long>((total + SW_WARPS - 1) / SW_WARPS, (long long)ctx->sm_count);
        KT(K_SYNC_WARPN);
#define SWN(NN) sync_metric_warpn_kernel<NN><<<grid, SW_WARPS * 32, SW_WARPS * SW_RING_BYTES, st>>>(samples, n_samples, stride, (float)kp.thr, kp.thr, \
                                                                                              w.detmask, w.trigmask, w.wps, (int)tiles, (int)span, (int)spans, (int)total)
        switch (kp.N) {
        case 32: SWN(32); break;
        case 64: SWN(64); break;
        case 128: SWN(128); break;
        case 256: SWN(256); break;
        default: SWN(512); break;
        }
#undef SWN
    } else if (!ctx->no_tma && make_sample_map(&tmap, samples, n_streams, n_samples, stride)) {
        // TMA path: 3-D map {32 floats, rows of 16 samples, streams}; whole rows only (the kernel patches the tail)
        const long long tiles = (n_samples + SV_T - 1) / SV_T;
        const long long spans = (tiles + ST_SPAN_TILES - 1) / ST_SPAN_TILES;
        const long long total = spans * n_streams;
        const unsigned grid = (unsigned)std::min<long long>(total, (long long)ctx->sm_count * ctx->sync_tma_occ);
        KT(K_SYNC_TMA);
        sync_metric_tma_kernel<<<grid, SV_THREADS, ctx->sync_tma_smem, st>>>(tmap, samples, n_samples, stride, kp.N, (float)kp.thr,
                                                                              kp.thr, w.detmask, w.trigmask, w.wps, tiles, spans, total);
    } else {
        const long long tiles = (n_samples + SV_T - 1) / SV_T;
        dim3 grid((unsigned)tiles, (unsigned)n_streams);
        KT(K_SYNC_FAST);
#define SVF(NN) sync_metric_fast_kernel<NN><<<grid, SV_THREADS, ctx->sync_fast_smem, st>>>(samples, n_samples, stride, kp.N, (float)kp.thr, kp.thr, w.detmask, w.trigmask, w.wps)
        switch (kp.N) {
        case 64: SVF(64); break;
        case 128: SVF(128); break;
        case 1024: SVF(1024); break;
        case 2048: SVF(2048); break;
        default: SVF(0); break;
        }
#undef SVF
    }
    const long long pb = (w.n_words + OFDMX_THREADS * PL_WPT - 1) / (OFDMX_THREADS * PL_WPT);
    {
        KT(K_PLATEAU);
        if (n_samples + 2LL * kp.cp + 64 < 0x7fffffffLL)      // per-stream sample indices fit an int
            plateau_kernel<int><<<(unsigned)pb, OFDMX_THREADS, 0, st>>>(w.detmask, w.trigmask, n_samples, w.wps, n_streams, kp.cp, w.blocksum);
        else
            plateau_kernel<long long><<<(unsigned)pb, OFDMX_THREADS, 0, st>>>(w.detmask, w.trigmask, n_samples, w.wps, n_streams, kp.cp, w.blocksum);
    }
    { KT(K_TRIG_SCAN); trig_scan_kernel<<<1, 1024, 0, st>>>(w.blocksum, w.nb, (int)max_trig, counts_dev, w.n_trig, w.stream_start, n_streams); }
    { KT(K_TRIG_SCATTER); trig_scatter_kernel<<<w.nb, OFDMX_THREADS, 0, st>>>(w.trigmask, w.n_words, w.wps, w.blocksum, (int)max_trig,
                                                                             w.trig, w.trig_stream, w.stream_start); }
    {
        KT(K_CFO);
        const unsigned cg = ctx->sm_count * 8;
        if (kp.N <= 64)
            cfo_small_kernel<8><<<cg, OFDMX_THREADS, 0, st>>>(samples, n_samples, stride, kp.N, w.trig, w.trig_stream, w.n_trig, w.cfo);
        else if (kp.N == 128)
            cfo_small_kernel<16><<<cg, OFDMX_THREADS, 0, st>>>(samples, n_samples, stride, kp.N, w.trig, w.trig_stream, w.n_trig, w.cfo);
        else
            cfo_kernel<<<cg, OFDMX_THREADS, 0, st>>>(samples, n_samples, stride, kp.N, w.trig, w.trig_stream, w.n_trig, w.cfo);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

}  // namespace

// dispatch over fft_len to the per-fft_len objects (ofdmx_k_framew.cu, ofdmx_k_txw.cu)
cudaError_t ofdmx_fw_configure(int nfft, int bps, size_t smem, int threads, int *occ)
{
    switch (nfft) {
    case 64: return ofdmx_fw_configure_64(bps, smem, threads, occ);
    case 128: return ofdmx_fw_configure_128(bps, smem, threads, occ);
    case 256: return ofdmx_fw_configure_256(bps, smem, threads, occ);
    case 512: return ofdmx_fw_configure_512(bps, smem, threads, occ);
    case 1024: return ofdmx_fw_configure_1024(bps, smem, threads, occ);
    case 2048: return ofdmx_fw_configure_2048(bps, smem, threads, occ);
    default: return cudaErrorInvalidValue;
    }
}
bool ofdmx_fw_launch(int nfft, int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const FwArgs &a)
{
    switch (nfft) {
    case 64: return ofdmx_fw_launch_64(bps, grid, threads, smem, st, a);
    case 128: return ofdmx_fw_launch_128(bps, grid, threads, smem, st, a);
    case 256: return ofdmx_fw_launch_256(bps, grid, threads, smem, st, a);
    case 512: return ofdmx_fw_launch_512(bps, grid, threads, smem, st, a);
    case 1024: return ofdmx_fw_launch_1024(bps, grid, threads, smem, st, a);
    case 2048: return ofdmx_fw_launch_2048(bps, grid, threads, smem, st, a);
    default: return false;
    }
}
cudaError_t ofdmx_txw_configure(int nfft, int bps, size_t smem)
{
    switch (nfft) {
    case 64: return ofdmx_txw_configure_64(bps, smem);
    case 128: return ofdmx_txw_configure_128(bps, smem);
    case 256: return ofdmx_txw_configure_256(bps, smem);
    case 512: return ofdmx_txw_configure_512(bps, smem);
    case 1024: return ofdmx_txw_configure_1024(bps, smem);
    default: return cudaErrorInvalidValue;
    }
}
bool ofdmx_txw_launch(int nfft, int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const TxwArgs &a)
{
    switch (nfft) {
    case 64: return ofdmx_txw_launch_64(bps, grid, threads, smem, st, a);
    case 128: return ofdmx_txw_launch_128(bps, grid, threads, smem, st, a);
    case 256: return ofdmx_txw_launch_256(bps, grid, threads, smem, st, a);
    case 512: return ofdmx_txw_launch_512(bps, grid, threads, smem, st, a);
    case 1024: return ofdmx_txw_launch_1024(bps, grid, threads, smem, st, a);
    default: return false;
    }
}

// =============================================================================================
extern "C" {

int ofdmx_abi_version(void) { return OFDMX_ABI_VERSION; }
int ofdmx_params_size(void) { return (int)sizeof(ofdmx_params); }
int ofdmx_frame_size(void) { return (int)sizeof(ofdmx_frame); }

const char *ofdmx_last_error(const ofdmx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C++" {
// Validates the parameters and builds the plan: derived constants, the table image (stg) and the launch
// configuration.  Host work only (plus cudaFuncSetAttribute / occupancy queries, which do not touch the device
// queues): this is what both ofdmx_create and ofdmx_reconfigure run.
static int build_plan(const ofdmx_params *prm, PlanFields *c, Stage &stg)
{
    const int N = prm->fft_len;
    if (N < 32 || N > 4096 || (N & (N - 1))) return fail(nullptr, OFDMX_ERR_PARAM, "fft_len must be a power of two in 32..4096");
    if (prm->cp_len < 0 || prm->cp_len > N) return fail(nullptr, OFDMX_ERR_PARAM, "cp_len out of range");
    if (prm->rolloff < 0 || prm->rolloff > prm->cp_len)
        return fail(nullptr, OFDMX_ERR_PARAM, "cyclic prefixer: rolloff len must smaller than the cyclic prefix.");
    if (prm->n_occ_sets < 1 || !prm->occ_sizes || !prm->occ_carriers) return fail(nullptr, OFDMX_ERR_PARAM, "occupied_carriers missing");
    if (!prm->sync_word1) return fail(nullptr, OFDMX_ERR_PARAM, "Length of sync sequence(s) must be FFT length.");
    const int nsw = prm->sync_word2 ? 2 : 1;      // sync_word2 == NULL: the single-sync-word mode (sync_word2=())
    std::vector<float2> hpts, ppts;
    std::vector<uint8_t> lut_h, lut_p;
    float qiw_h = 1.f, qiw_p = 1.f;
    if (prm->qam_normalization < 0 || prm->qam_normalization > 1)
        return fail(nullptr, OFDMX_ERR_PARAM, "unknown version switch value");
    for (int i = 0; i < 7; i++)
        if (prm->reserved[i] != 0) return fail(nullptr, OFDMX_ERR_PARAM, "ofdmx_params.reserved must be zero (ABI mismatch?)");
    if (!make_constellation(prm->bps_header, prm->qam_normalization, hpts, lut_h, &qiw_h)
        || !make_constellation(prm->bps_payload, prm->qam_normalization, ppts, lut_p, &qiw_p))
        return fail(nullptr, OFDMX_ERR_PARAM, "Modulation not supported.");
    if (prm->max_pkt_bytes < 1 || prm->max_pkt_bytes > 4095) return fail(nullptr, OFDMX_ERR_PARAM, "max_pkt_bytes must be 1..4095");

    c->prm = *prm;

    KP &kp = c->kp;
    kp.N = N;
    kp.logN = 0;
    while ((1 << kp.logN) < N) kp.logN++;
    kp.cp = prm->cp_len;
    kp.D = N + prm->cp_len;
    kp.n_occ_sets = prm->n_occ_sets;
    kp.n_pil_sets = prm->n_pilot_sets;
    kp.n_pil_sym_sets = prm->n_pilot_sym_sets > 0 ? prm->n_pilot_sym_sets : 1;
    kp.bps_h = prm->bps_header;
    kp.bps_p = prm->bps_payload;
    kp.crc_mode = prm->crc_mode ? 1 : 0;
    kp.holdoff = prm->demux_holdoff;
    kp.max_pkt_bytes = prm->max_pkt_bytes;
    kp.max_pkt_syms = (prm->max_pkt_bytes * 8 + prm->bps_payload - 1) / prm->bps_payload;
    kp.thr = (double)prm->threshold;
    kp.alpha = prm->alpha;
    kp.tx_scale = prm->tx_scale;
    kp.tx_clip = prm->tx_clip;
    kp.roll = prm->rolloff > 1 ? prm->rolloff : 0;      // a flank of length 1 would just be rectangular
    kp.roll_flank = nullptr;
    kp.qiw_h = qiw_h;
    kp.qiw_p = qiw_p;
    kp.h_taps = nullptr;
    kp.h_stride = 0;

    int rc = 0;
    auto bail = [&](int code) { return code; };

    // carrier plan
    std::vector<int> occ_bins, occ_base, occ_size, occ_u;
    std::vector<uint8_t> occ_mask(N, 0);
    for (int s = 0, a = 0; s < prm->n_occ_sets; s++) {
        occ_base.push_back(a);
        occ_size.push_back(prm->occ_sizes[s]);
        if (prm->occ_sizes[s] < 1) return bail(fail(nullptr, OFDMX_ERR_PARAM, "empty occupied carrier set"));
        for (int k = 0; k < prm->occ_sizes[s]; k++) {
            const int cn = prm->occ_carriers[a + k];
            if (cn < -N / 2 || cn >= N) return bail(fail(nullptr, OFDMX_ERR_PARAM, "carrier index out of range"));
            const int bin = shifted_bin(cn, N);
            occ_bins.push_back(bin);
            occ_mask[bin] = 1;
        }
        a += prm->occ_sizes[s];
    }
    for (int k = 0; k < N; k++) if (occ_mask[k]) occ_u.push_back(k);
    c->occ_sizes = occ_size;
    c->hl = occ_size[0];
    kp.hl = c->hl;
    kp.n_occ_u = (int)occ_u.size();

    const int nps = std::max(1, prm->n_pilot_sets);
    std::vector<uint8_t> pil_flag((size_t)nps * N, 0);
    std::vector<float2> pil_val((size_t)nps * N, make_float2(0.f, 0.f));
    std::vector<int> pil_bins, pil_base, pil_size, pil_sym_base;
    std::vector<float2> pil_sym;
    for (int s = 0, a = 0; s < prm->n_pilot_sym_sets; s++) {
        pil_sym_base.push_back(a);
        for (int k = 0; k < prm->pilot_sym_sizes[s]; k++)
            pil_sym.push_back(make_float2(prm->pilot_symbols[2 * (a + k)], prm->pilot_symbols[2 * (a + k) + 1]));
        a += prm->pilot_sym_sizes[s];
    }
    if (pil_sym_base.empty()) pil_sym_base.push_back(0);
    for (int s = 0, a = 0; s < prm->n_pilot_sets; s++) {
        pil_base.push_back(a);
        pil_size.push_back(prm->pilot_sizes[s]);
        // ofdm_equalizer_1d_pilots: "pilot carriers and -symbols do not match" -> ValueError
        if (s >= prm->n_pilot_sym_sets || prm->pilot_sym_sizes[s] != prm->pilot_sizes[s])
            return bail(fail(nullptr, OFDMX_ERR_PARAM, "pilot carriers and -symbols do not match."));
        for (int k = 0; k < prm->pilot_sizes[s]; k++) {
            const int cn = prm->pilot_carriers[a + k];
            if (cn < -N / 2 || cn >= N) return bail(fail(nullptr, OFDMX_ERR_PARAM, "pilot carrier index out of range"));
            const int bin = shifted_bin(cn, N);
            pil_bins.push_back(bin);
            pil_flag[(size_t)s * N + bin] = 1;
            pil_val[(size_t)s * N + bin] = pil_sym[pil_sym_base[s] + k];
        }
        a += prm->pilot_sizes[s];
    }
    // every pilot-symbol set used on TX (i % n_sym_sets) must cover the carrier set it lands on
    for (int s = 0; s < prm->n_pilot_sym_sets && prm->n_pilot_sets > 0; s++)
        if (prm->pilot_sym_sizes[s] < *std::max_element(pil_size.begin(), pil_size.end()))
            return bail(fail(nullptr, OFDMX_ERR_PARAM, "pilot carriers and -symbols do not match."));

    // sync words, chanest tables
    std::vector<float2> sw1(N), sw2(N, make_float2(0.f, 0.f)), inv_sw2(N), cv_conj;
    std::vector<int> cv_k;
    int first = 0, last = N - 1;
    for (int k = 0; k < N; k++) {
        sw1[k] = make_float2(prm->sync_word1[2 * k], prm->sync_word1[2 * k + 1]);
        if (nsw == 2) sw2[k] = make_float2(prm->sync_word2[2 * k], prm->sync_word2[2 * k + 1]);
    }
    // ofdm_chanest_vcvc: d_ref_sym = sync word 2, or sync word 1 when there is only one
    const std::vector<float2> &ref = (nsw == 2) ? sw2 : sw1;
    for (int k = 0; k < N; k++) if (ref[k].x != 0.f || ref[k].y != 0.f) { first = k; break; }
    for (int k = N - 1; k >= 0; k--) if (ref[k].x != 0.f || ref[k].y != 0.f) { last = k; break; }
    kp.nsw = nsw;
    kp.interp = 0;
    if (nsw == 1 && first + 1 < N && sw1[first + 1].x == 0.f && sw1[first + 1].y == 0.f) {
        // [UPSTREAM ofdm_chanest_vcvc_impl ctor] a sync word on every second carrier: taps are interpolated
        if (last + 1 < N) last++;
        kp.interp = 1;
    }
    kp.first_act = first;
    kp.last_act = last;
    for (int k = 0; k < N; k++) {
        std::complex<double> a(sw1[k].x, sw1[k].y), b(sw2[k].x, sw2[k].y), r(ref[k].x, ref[k].y);
        inv_sw2[k] = make_float2(0.f, 0.f);
        if (r != 0.0) {
            std::complex<double> iv = 1.0 / r;
            inv_sw2[k] = make_float2((float)iv.real(), (float)iv.imag());
        }
        if (nsw == 2 && a != 0.0) {
            std::complex<double> cv = b / a;
            if (cv != 0.0) {
                cv_k.push_back(k);
                cv_conj.push_back(make_float2((float)cv.real(), (float)-cv.imag()));
            }
        }
    }
    if (nsw == 1) {
        // d_known_symbol_diffs[i] = |sw1[i] - sw1[i+2]|^2 for i = first, first+2, ... < last-2 (and < N-2); the non-zero
        // ones go to the (cv_k, cv_conj.x) tables the offset search walks
        for (int i = first; i < last - 2 && i < N - 2; i += 2) {
            const double dr = (double)sw1[i].x - sw1[i + 2].x, di = (double)sw1[i].y - sw1[i + 2].y;
            const float v = (float)(dr * dr + di * di);
            if (v != 0.f) { cv_k.push_back(i); cv_conj.push_back(make_float2(v, 0.f)); }
        }
    }
    kp.n_cv = (int)cv_k.size();
    int gneg = -first, gpos = N - last - 1;
    if (prm->max_carr_offset != -1) {
        gneg = std::max(-prm->max_carr_offset, gneg);
        gpos = std::min(prm->max_carr_offset, gpos);
    }
    if (gneg % 2) gneg++;
    if (gpos % 2) gpos--;
    kp.gneg = gneg;
    kp.gpos = gpos;
    for (int k : cv_k)
        if (nsw == 2 && (k + gneg < 0 || k + gpos >= N)) return bail(fail(nullptr, OFDMX_ERR_PARAM, "sync words inconsistent with carrier offset range"));

    // twiddles
    std::vector<float2> tw(N);
    for (int k = 0; k < N; k++) {
        const double a = -2.0 * M_PI * k / N;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    // header scramble mask (packet_header_ofdm) and payload keystream (additive_scrambler_bb)
    std::vector<uint8_t> hdr_mask(c->hl, 0), keystream((size_t)prm->max_pkt_bytes + 8, 0);
    if (prm->scramble_header) {
        Lfsr l(0x8a, 0x6f, 7);
        for (int i = 0; i < c->hl; i++)
            for (int k = 0; k < prm->bps_header; k++) hdr_mask[i] ^= (uint8_t)(l.next() << k);
    }
    {
        Lfsr l(0x8a, (uint32_t)prm->scramble_seed, 7);
        for (size_t i = 0; i < keystream.size(); i++)
            for (int k = 0; k < 8; k++) keystream[i] ^= (uint8_t)(l.next() << k);
    }
    // CRC tables: independent of the parameters, computed once per process (a reconfiguration only copies them)
    const CrcTables &ct = crc_tables();
    const std::vector<uint32_t> &crc_tab = ct.tab, &crc_pow = ct.pow, &crc_pow64 = ct.pow64, &crc_pow8 = ct.pow8;
    c->x_2048 = ct.x_2048;

    // reciprocals of the constellation points, position of each union carrier in each set's list
    std::vector<float2> inv_hpts, inv_ppts;
    for (auto &v : hpts) { std::complex<double> z = 1.0 / std::complex<double>(v.x, v.y); inv_hpts.push_back(make_float2((float)z.real(), (float)z.imag())); }
    for (auto &v : ppts) { std::complex<double> z = 1.0 / std::complex<double>(v.x, v.y); inv_ppts.push_back(make_float2((float)z.real(), (float)z.imag())); }
    std::vector<int> pos_su((size_t)prm->n_occ_sets * occ_u.size(), -1);
    for (int s = 0; s < prm->n_occ_sets; s++)
        for (int q = 0; q < occ_size[s]; q++) {
            const int bin = occ_bins[occ_base[s] + q];
            const int u = (int)(std::lower_bound(occ_u.begin(), occ_u.end(), bin) - occ_u.begin());
            if (pos_su[(size_t)s * occ_u.size() + u] < 0) pos_su[(size_t)s * occ_u.size() + u] = q;
        }
    kp.max_frame_syms = payload_ofdm_syms(c, kp.max_pkt_syms);
    kp.pil_in_occ = 0;
    for (int ps = 0; ps < prm->n_pilot_sets; ps++)
        for (int k = 0; k < N; k++)
            if (pil_flag[(size_t)ps * N + k] && occ_mask[k]) kp.pil_in_occ = 1;

    // CRC-8 (poly 0x07, init 0xFF) is affine in the message bits: crc(m) = crc(0) ^ XOR of per-bit terms
    auto crc8_bytes = [](unsigned len, unsigned num) {
        uint8_t b[4] = { (uint8_t)(len & 0xFF), (uint8_t)(len >> 8), (uint8_t)(num & 0xFF), (uint8_t)(num >> 8) };
        uint8_t crc = 0xFF;
        for (int i = 0; i < 4; i++) {
            crc ^= b[i];
            for (int k = 0; k < 8; k++) crc = (crc & 0x80) ? (uint8_t)((crc << 1) ^ 0x07) : (uint8_t)(crc << 1);
        }
        return crc;
    };
    std::vector<uint8_t> crc8_bit(24);
    kp.crc8_zero = crc8_bytes(0, 0);
    for (int l = 0; l < 24; l++)
        crc8_bit[l] = (uint8_t)(crc8_bytes(l < 12 ? (1u << l) : 0u, l < 12 ? 0u : (1u << (l - 12))) ^ kp.crc8_zero);

    kp.y1_lo = 0;
    kp.y1_span = 1;
    if (!cv_k.empty()) {
        kp.y1_lo = *std::min_element(cv_k.begin(), cv_k.end()) + kp.gneg;
        kp.y1_span = *std::max_element(cv_k.begin(), cv_k.end()) + kp.gpos - kp.y1_lo + 1;
    }

#define UP(vec, field)                                          \
    if ((rc = upload(stg, vec, &kp.field)) != 0) return bail(rc);
    UP(tw, tw) UP(occ_bins, occ_bins) UP(occ_base, occ_base) UP(occ_size, occ_size) UP(occ_u, occ_u)
    UP(pil_flag, pil_flag) UP(pil_val, pil_val) UP(pil_bins, pil_bins) UP(pil_base, pil_base)
    UP(pil_size, pil_size) UP(pil_sym, pil_sym) UP(pil_sym_base, pil_sym_base) UP(sw1, sw1) UP(sw2, sw2)
    UP(cv_k, cv_k) UP(cv_conj, cv_conj) UP(inv_sw2, inv_sw2) UP(hdr_mask, hdr_mask) UP(keystream, keystream)
    UP(crc_tab, crc_tab) UP(crc_pow, crc_pow) UP(hpts, hpts) UP(ppts, ppts) UP(lut_h, lut_h) UP(lut_p, lut_p)
    UP(inv_hpts, inv_hpts) UP(inv_ppts, inv_ppts) UP(pos_su, pos_su) UP(crc8_bit, crc8_bit) UP(crc_pow64, crc_pow64) UP(crc_pow8, crc_pow8)
#undef UP
    if (kp.roll) {
        // [UPSTREAM ofdm_cyclic_prefixer_impl.cc ctor]: flanks are one sample shorter than rolloff_len
        // (the first sample of the up / down flank is always zero / one), float vectors
        std::vector<float> fl((size_t)2 * (kp.roll - 1));
        for (int i = 1; i < kp.roll; i++) {
            fl[i - 1] = (float)(0.5 * (1 + cos(M_PI * i / kp.roll - M_PI)));
            fl[kp.roll - 1 + i - 1] = (float)(0.5 * (1 + cos(M_PI * (kp.roll - i) / kp.roll - M_PI)));
        }
        if ((rc = upload(stg, fl, &kp.roll_flank)) != 0) return bail(rc);
    }
    // ---- fft_len 1024 warp-per-packet TX kernel: per-bin allocation map and the constant sync symbols
    if (nsw == 2 && (N == 1024 || N == 512 || N == 256 || N == 128 || N == 64) && prm->n_occ_sets == 1 && prm->n_pilot_sets <= 1 && !kp.pil_in_occ
        && kp.bps_h == 1 && c->hl >= 32 && kp.roll <= 512) {   // (the flank samples are parked in half of the transpose buffer)
        std::vector<uint16_t> tx_map((size_t)N, (uint16_t)TXW_EMPTY);
        for (int q = 0; q < occ_size[0]; q++) tx_map[occ_bins[occ_base[0] + q] ^ (N / 2)] = (uint16_t)q;   // later entries win, as in the scatter
        for (int q = 0; q < (int)pil_bins.size(); q++) tx_map[pil_bins[q] ^ (N / 2)] = (uint16_t)(TXW_PILOT | q);
        const int nfl = kp.roll ? kp.roll - 1 : 0;
        std::vector<float2> sync_td((size_t)2 * kp.D + nfl);
        std::vector<double> tlr((size_t)std::max(nfl, 1), 0.0), tli((size_t)std::max(nfl, 1), 0.0);   // prefixer delay line (double)
        for (int o = 0; o < 2; o++) {
            const float *sw = o ? prm->sync_word2 : prm->sync_word1;       // shifted order, (re, im)
            // unnormalised inverse DFT of the word (natural bin order = shifted index ^ N/2), radix-2 in double
            std::vector<std::complex<double>> X((size_t)N);
            for (int nn = 0; nn < N; nn++) X[nn] = std::complex<double>(sw[2 * (nn ^ (N / 2))], sw[2 * (nn ^ (N / 2)) + 1]);
            for (int i = 1, jrev = 0; i < N; i++) {
                int bit = N >> 1;
                for (; jrev & bit; bit >>= 1) jrev ^= bit;
                jrev ^= bit;
                if (i < jrev) std::swap(X[i], X[jrev]);
            }
            for (int len = 2; len <= N; len <<= 1) {
                const double ang = 2.0 * M_PI / len;
                for (int i = 0; i < N; i += len)
                    for (int k = 0; k < len / 2; k++) {
                        const std::complex<double> w(cos(ang * k), sin(ang * k));
                        const std::complex<double> u = X[i + k], v = X[i + k + len / 2] * w;
                        X[i + k] = u + v;
                        X[i + k + len / 2] = u - v;
                    }
            }
            std::vector<double> xr((size_t)N), xi((size_t)N);
            for (int t = 0; t < N; t++) { xr[t] = X[t].real(); xi[t] = X[t].imag(); }
            for (int m = 0; m < kp.D; m++) {
                const int t = (m - kp.cp + N) & (N - 1);
                double sr = xr[t], si = xi[t];
                if (m < nfl) {       // ofdm_cyclic_prefixer flanks: out = x * up + delay; float32 flanks as in the kernels
                    const double up = (double)(float)(0.5 * (1 + cos(M_PI * (m + 1) / kp.roll - M_PI)));
                    sr = sr * up + tlr[m];
                    si = si * up + tli[m];
                }
                float vr = (float)(sr * (double)kp.tx_scale), vi = (float)(si * (double)kp.tx_scale);
                if (kp.tx_clip > 0.f) {
                    vr = vr < -kp.tx_clip ? -kp.tx_clip : (vr > kp.tx_clip ? kp.tx_clip : vr);
                    vi = vi < -kp.tx_clip ? -kp.tx_clip : (vi > kp.tx_clip ? kp.tx_clip : vi);
                }
                sync_td[(size_t)o * kp.D + m] = make_float2(vr, vi);
            }
            for (int m = 0; m < nfl; m++) {      // delay line left by this symbol: first body samples x down flank
                const double dn = (double)(float)(0.5 * (1 + cos(M_PI * (kp.roll - (m + 1)) / kp.roll - M_PI)));
                tlr[m] = xr[m] * dn;
                tli[m] = xi[m] * dn;
            }
        }
        for (int m = 0; m < nfl; m++) sync_td[(size_t)2 * kp.D + m] = make_float2((float)tlr[m], (float)tli[m]);
        if ((rc = upload(stg, tx_map, &c->tx_map)) != 0) return bail(rc);
        if ((rc = upload(stg, sync_td, &c->sync_td)) != 0) return bail(rc);
        c->tx1kw = true;
    }

    // shared-memory budgets
    {
        c->frame_smem = (size_t)N * 8 * 4 + 64 + N + align_up(c->hl, 16) + align_up(kp.max_pkt_syms, 16)
                        + align_up(kp.max_pkt_bytes, 16) + 16;
        c->tx_smem = (size_t)N * 8 + 64 + align_up(kp.max_pkt_bytes + 8, 16) + align_up(c->hl, 16) + 16 + (size_t)kp.roll * 8;
        if (N == 1024 && nsw == 2) {
            c->frame1k_warps = std::max(4, std::min(F1K_MAXW, 3 + kp.max_frame_syms));
            c->frame1k_smem = frame1024_smem_bytes(c->frame1k_warps, kp.n_occ_u, c->hl, kp.max_pkt_syms, kp.max_pkt_bytes);
            const cudaError_t e1 = ofdmx_f1k_configure(kp.bps_p, c->frame1k_smem);
            if (e1 != cudaSuccess)
                return bail(fail(nullptr, OFDMX_ERR_CUDA, "shared memory configuration failed: %s", cudaGetErrorString(e1)));
        }
        {
            const int ngc = (kp.gpos - kp.gneg) / 2 + 1;
            const bool simple = (kp.n_occ_sets == 1 && kp.n_pil_sets <= 1 && !kp.pil_in_occ);
            // guard band of the symbol buffer (natural bin order): the longest run of bins that no equaliser read
            // (occupied carrier + any candidate offset) touches; the per-symbol decisions go there if they fit
            c->frame1kw_dec_off = -1;
            if (N == 1024) {
                std::vector<char> used(1024, 0);
                for (int u = 0; u < kp.n_occ_u; u++)
                    for (int g = kp.gneg; g <= kp.gpos; g++) {
                        const int b = occ_u[u] + g;
                        if (b >= 0 && b < 1024) used[b ^ 512] = 1;
                    }
                int best = 0, best_at = -1, run = 0;
                for (int i = 0; i < 1024; i++) {
                    run = used[i] ? 0 : run + 1;
                    if (run > best) { best = run; best_at = i - run + 1; }
                }
                const int at = (best_at + 1) & ~1;                       // 16-byte aligned
                const int need = (kp.n_occ_u + 7) / 8 + 2;               // float2 slots for n_occ_u bytes
                if (best_at >= 0 && at + need <= best_at + best) c->frame1kw_dec_off = at;
            }
            c->frame1kw_dec_all = ((occ_size[0] * kp.bps_p) % 8 == 0) ? 0 : (kp.max_frame_syms + 1) * occ_size[0];
            if (c->frame1kw_dec_all > 0) c->frame1kw_dec_off = -1;
            c->frame1kw_warps = FW_WARPS;
            while (c->frame1kw_warps > 4 && framew_smem_bytes(N, kp.n_occ_u, kp.y1_span, c->frame1kw_warps, c->frame1kw_dec_off >= 0, c->frame1kw_dec_all) > 227 * 1024)
                c->frame1kw_warps--;
            c->frame1kw_smem = framew_smem_bytes(N, kp.n_occ_u, kp.y1_span, c->frame1kw_warps, c->frame1kw_dec_off >= 0, c->frame1kw_dec_all);
            // warp-per-frame kernel: fft_len 1024 (register FFT 32x32, <= 4 carrier-offset candidates), fft_len 2048 (two
            // interleaved 1024-point transforms) and fft_len 64 / 128 (register FFT + lane-shuffle FFT)
            // (carriers are walked in list order: a list that names a carrier twice goes to the CTA-per-frame kernels)
            c->frame1kw = ((N == 1024 && ngc <= 4 && c->frame1kw_dec_all == 0) || N == 64 || N == 128 || N == 256 || N == 512 || N == 2048) && simple && kp.bps_h == 1
                          && c->hl >= 32 && c->hl <= 2048 && kp.n_occ_u == occ_size[0] && nsw == 2
                          && c->frame1kw_smem <= 227 * 1024;
            // fft_len 2048: the same plans on the pair-of-warps kernel (bits per OFDM symbol a byte multiple)
            if (N == 2048 && nsw == 2 && simple && kp.bps_h == 1 && c->hl >= 32 && kp.n_occ_u == occ_size[0] && ngc <= 4
                && c->frame1kw_dec_all == 0) {
                std::vector<uint16_t> lists[2][3];
                for (int q = 0; q < occ_size[0]; q++) {
                    const int ks = occ_bins[occ_base[0] + q], kn = ks ^ (N / 2), h = kn & 1;
                    lists[h][0].push_back((uint16_t)(kn >> 1));
                    lists[h][1].push_back((uint16_t)q);
                    lists[h][2].push_back((uint16_t)ks);
                }
                const int nh = (int)std::max(lists[0][0].size(), lists[1][0].size());
                std::vector<uint16_t> tab((size_t)4 + 6 * (size_t)nh, 0);
                tab[0] = (uint16_t)lists[0][0].size(); tab[1] = (uint16_t)lists[1][0].size(); tab[2] = (uint16_t)nh;
                for (int h = 0; h < 2; h++)
                    for (int a = 0; a < 3; a++)
                        std::copy(lists[h][a].begin(), lists[h][a].end(), tab.begin() + 4 + (size_t)(3 * h + a) * nh);
                c->framep_hsz = (std::max(nh, kp.y1_span / 2 + 2) + 1) & ~1;
                c->framep_smem = framep_smem_bytes(kp.n_occ_u, nh, c->framep_hsz);
                if (c->framep_smem <= 227 * 1024 && nh > 0 && ofdmx_fp_configure(kp.bps_p, c->framep_smem) == cudaSuccess) {
                    if ((rc = upload(stg, tab, &c->pair_tab)) != 0) return bail(rc);
                    c->framep = true;
                }
            }
            if (c->frame1kw) {
                int occ = 1;
                const cudaError_t e1 = ofdmx_fw_configure(N, kp.bps_p, c->frame1kw_smem, c->frame1kw_warps * 32, &occ);
                c->frame1kw_ctas = std::max(1, occ);
                if (e1 != cudaSuccess) c->frame1kw = false;
            }
        }
        if (c->tx1kw) {
            c->tx1kw_smem = txw_smem_bytes(N, kp.max_pkt_bytes, TXW_WARPS, kp.roll);
            const cudaError_t e1 = (c->tx1kw_smem <= 227 * 1024) ? ofdmx_txw_configure(N, kp.bps_p, c->tx1kw_smem) : cudaErrorInvalidValue;
            if (e1 != cudaSuccess || c->tx1kw_smem > 227 * 1024) c->tx1kw = false;
        }
        c->sync_tma_smem = sync_tma_smem_bytes(N);
        if (ofdmx_raise_smem_limit(sync_metric_tma_kernel, c->sync_tma_smem) != cudaSuccess)
            return bail(fail(nullptr, OFDMX_ERR_CUDA, "shared memory configuration failed: %s", cudaGetErrorString(cudaGetLastError())));
        {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sync_metric_tma_kernel, SV_THREADS, c->sync_tma_smem) == cudaSuccess && occ > 0)
                c->sync_tma_occ = occ;
        }
        c->sync_warp_ok = false;
        {
            const size_t wb = SW_WARPS * SW_RING_BYTES;
            switch (N) {
            case 32: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warpn_kernel<32>, wb) == cudaSuccess; break;
            case 64: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warpn_kernel<64>, wb) == cudaSuccess; break;
            case 128: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warpn_kernel<128>, wb) == cudaSuccess; break;
            case 256: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warpn_kernel<256>, wb) == cudaSuccess; break;
            case 512: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warpn_kernel<512>, wb) == cudaSuccess; break;
            case 1024: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warp_kernel<16>, wb) == cudaSuccess; break;
            case 2048: c->sync_warp_ok = ofdmx_raise_smem_limit(sync_metric_warp_kernel<32>, wb) == cudaSuccess; break;
            default: break;
            }
        }
        c->sync_fast_smem = sync_fast_smem_bytes(N);
        if (ofdmx_raise_smem_limit(sync_metric_fast_kernel<0>, c->sync_fast_smem) != cudaSuccess
            || ofdmx_raise_smem_limit(sync_metric_fast_kernel<64>, c->sync_fast_smem) != cudaSuccess
            || ofdmx_raise_smem_limit(sync_metric_fast_kernel<128>, c->sync_fast_smem) != cudaSuccess
            || ofdmx_raise_smem_limit(sync_metric_fast_kernel<1024>, c->sync_fast_smem) != cudaSuccess
            || ofdmx_raise_smem_limit(sync_metric_fast_kernel<2048>, c->sync_fast_smem) != cudaSuccess
            || ofdmx_raise_smem_limit(rx_frame_kernel, c->frame_smem) != cudaSuccess
            || ofdmx_raise_smem_limit(tx_frame_kernel, c->tx_smem) != cudaSuccess)
            return bail(fail(nullptr, OFDMX_ERR_CUDA, "shared memory configuration failed: %s (is this an sm_100 device?)",
                             cudaGetErrorString(cudaGetLastError())));
    }
    return OFDMX_OK;
}

// Gives the table image of `np` a device address in the arena half that the running plan does not use, starts its
// upload on the configuration stream and makes `np` the context's plan.  No device-wide synchronisation and, once
// the arena is large enough, no allocation: work enqueued under the previous plan keeps its (by-value) parameter
// block and its half of the arena; the next call makes its stream wait for the upload event.
static int commit_plan(ofdmx_ctx *c, PlanFields &np, Stage &stg, bool first)
{
    const int h = first ? 0 : 1 - c->arena_cur;
    const size_t need = std::max<size_t>((stg.bytes.size() + 255) / 256 * 256, 256);
    if (!first && c->last_stream_valid) {
        // the plan being left: everything enqueued so far is the last work that reads its half of the arena
        CUDA_TRY(c, cudaEventRecord(c->ev_done[c->arena_cur], c->last_stream));
        c->done_valid[c->arena_cur] = true;
    }
    // the half being overwritten was read by the plan before the running one: the upload is ordered (on the device)
    // behind the end of that plan's work; the host goes on
    if (c->done_valid[h]) CUDA_TRY(c, cudaStreamWaitEvent(c->cfg_stream, c->ev_done[h], 0));
    if (c->arena[h].cap < need) {
        // first use of this half, or a plan larger than anything before (capacity grows with head-room)
        if (c->arena[h].p) {
            CUDA_TRY(c, cudaStreamSynchronize(c->cfg_stream));
            if (c->done_valid[h]) CUDA_TRY(c, cudaEventSynchronize(c->ev_done[h]));
            c->n_host_syncs++;
            CUDA_TRY(c, cudaFree(c->arena[h].p));
            c->arena[h] = DevBuf();
        }
        const size_t want = std::max<size_t>(2 * need, 1 << 20);
        CUDA_TRY(c, cudaMalloc(&c->arena[h].p, want));
        c->n_dev_allocs++;
        c->arena[h].cap = want;
    }
    if (c->stage_cap < need) {
        if (c->stage_pinned) {
            CUDA_TRY(c, cudaStreamSynchronize(c->cfg_stream));    // earlier images may still be on their way
            c->n_host_syncs++;
            CUDA_TRY(c, cudaFreeHost(c->stage_pinned));
            c->stage_pinned = nullptr;
            for (bool &u : c->stage_used) u = false;
        }
        const size_t want = std::max<size_t>(2 * need, 1 << 19);
        CUDA_TRY(c, cudaHostAlloc(&c->stage_pinned, want * ofdmx_ctx::STAGE_SLOTS, cudaHostAllocDefault));
        c->n_dev_allocs++;
        c->stage_cap = want;
    }
    const int slot = c->stage_next;
    c->stage_next = (slot + 1) % ofdmx_ctx::STAGE_SLOTS;
    if (c->stage_used[slot] && cudaEventQuery(c->ev_stage[slot]) != cudaSuccess) {
        CUDA_TRY(c, cudaEventSynchronize(c->ev_stage[slot]));     // STAGE_SLOTS reconfigurations in flight: back-pressure
        c->n_host_syncs++;
    }
    unsigned char *host = static_cast<unsigned char *>(c->stage_pinned) + (size_t)slot * c->stage_cap;
    std::memcpy(host, stg.bytes.data(), stg.bytes.size());
    unsigned char *base = static_cast<unsigned char *>(c->arena[h].p);
    for (auto &f : stg.fixups) *f.first = base + f.second;
    CUDA_TRY(c, cudaMemcpyAsync(base, host, stg.bytes.size(), cudaMemcpyHostToDevice, c->cfg_stream));
    CUDA_TRY(c, cudaEventRecord(c->ev_stage[slot], c->cfg_stream));
    c->stage_used[slot] = true;
    CUDA_TRY(c, cudaEventRecord(c->ev_tables, c->cfg_stream));
    static_cast<PlanFields &>(*c) = np;
    c->arena_cur = h;
    c->tables_pending = true;
    return OFDMX_OK;
}

// every entry point that launches work: the stream waits (device-side) for the current plan's tables
static int use_stream(ofdmx_ctx *c, cudaStream_t st, bool remember = true)
{
    if (c->tables_pending) {
        if (cudaEventQuery(c->ev_tables) == cudaSuccess) c->tables_pending = false;
        else CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_tables, 0));
    }
    if (remember) {         // (the private stream of ofdmx_rx_host is drained before that call returns)
        c->last_stream = st;
        c->last_stream_valid = true;
    }
    return 0;
}
}  // extern "C++"

int ofdmx_create(const ofdmx_params *prm, int device, ofdmx_ctx **out)
{
    if (!prm || !out) return fail(nullptr, OFDMX_ERR_PARAM, "null argument");
    *out = nullptr;
    int ndev = 0;
    // parameter errors are reported before any CUDA call (they surface on a box without a GPU too)
    PlanFields probe;
    Stage pstg;
    const bool have_dev = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > device && cudaSetDevice(device) == cudaSuccess;
    if (!have_dev) {
        (void)cudaGetLastError();
        // parameter validation comes first in build_plan, before its first CUDA call
        if (build_plan(prm, &probe, pstg) == OFDMX_ERR_PARAM) return OFDMX_ERR_PARAM;
        return fail(nullptr, OFDMX_ERR_CUDA, "no CUDA device %d (this library has no CPU fallback)", device);
    }
    ofdmx_ctx *c = new ofdmx_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    auto bail = [&](int code) { ofdmx_destroy(c); return code; };
    if (int rc = build_plan(prm, &probe, pstg)) return bail(rc);
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&c->cfg_stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&c->ev_tables, cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&c->ev_done[0], cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&c->ev_done[1], cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(nullptr, OFDMX_ERR_CUDA, "stream / event creation failed"));
    for (cudaEvent_t &e : c->ev_stage)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(nullptr, OFDMX_ERR_CUDA, "stream / event creation failed"));
    if (int rc = commit_plan(c, probe, pstg, true)) { g_err = c->err; return bail(rc); }
    if (const char *fg = getenv("OFDMX_FORCE_GENERIC")) c->force_generic = (fg[0] == '1');
    if (const char *nw = getenv("OFDMX_NO_WARP_FRAME")) c->no_warp_frame = (nw[0] == '1');
    if (const char *nt = getenv("OFDMX_NO_TMA")) c->no_tma = (nt[0] == '1');   // plain-load sync kernel instead of the TMA ring
    if (const char *ns = getenv("OFDMX_NO_WARP_SYNC")) c->no_warp_sync = (ns[0] == '1');
    if (const char *nx = getenv("OFDMX_NO_WARP_TX")) c->no_warp_tx = (nx[0] == '1');
    if (const char *na = getenv("OFDMX_NO_AGC_SPANS")) c->no_agc_spans = (na[0] == '1');
    if (const char *np2 = getenv("OFDMX_NO_PAIR_FRAME")) c->no_pair_frame = (np2[0] == '1');
    *out = c;
    return OFDMX_OK;
}

void ofdmx_destroy(ofdmx_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &r : c->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    cudaDeviceSynchronize();        // nothing of this context may still be running when its tables go away
    for (DevBuf *b : { &c->arena[0], &c->arena[1], &c->ws, &c->ws_host, &c->ws_papr, &c->ws_agc, &c->h_samples, &c->h_frames, &c->h_bytes, &c->h_counts })
        if (b->p) cudaFree(b->p);
    if (c->stage_pinned) cudaFreeHost(c->stage_pinned);
    if (c->ev_tables) cudaEventDestroy(c->ev_tables);
    for (cudaEvent_t e : c->ev_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_stage) if (e) cudaEventDestroy(e);
    if (c->cfg_stream) cudaStreamDestroy(c->cfg_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int ofdmx_profile(ofdmx_ctx *c, int enable)
{
    if (!c) return OFDMX_ERR_PARAM;
    c->profiling = enable != 0;
    for (auto &r : c->prof_pending) { c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b); }
    c->prof_pending.clear();
    for (int i = 0; i < K_NSLOTS; i++) { c->prof_ms[i] = 0; c->prof_calls[i] = 0; }
    c->prof_errors = 0;
    return OFDMX_OK;
}
int ofdmx_profile_slots(void) { return K_NSLOTS; }
const char *ofdmx_profile_name(int slot) { return (slot >= 0 && slot < K_NSLOTS) ? kSlotNames[slot] : ""; }
int ofdmx_profile_read(ofdmx_ctx *c, float *ms_total, int64_t *calls)
{
    if (!c || !ms_total || !calls) return OFDMX_ERR_PARAM;
    if (int rc = check_device(c)) return rc;
    for (auto &r : c->prof_pending) {
        CUDA_TRY(c, cudaEventSynchronize(r.b));
        float ms = 0.f;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, r.a, r.b));
        c->prof_ms[r.slot] += ms;
        c->prof_calls[r.slot]++;
        c->prof_pool.push_back(r.a);
        c->prof_pool.push_back(r.b);
    }
    c->prof_pending.clear();
    for (int i = 0; i < K_NSLOTS; i++) { ms_total[i] = (float)c->prof_ms[i]; calls[i] = c->prof_calls[i]; }
    if (c->prof_errors) return fail(c, OFDMX_ERR_CUDA, "%lld kernel timings lost (CUDA event creation / record failed)", (long long)c->prof_errors);
    return OFDMX_OK;
}

int ofdmx_set_emit_all(ofdmx_ctx *c, int enable)
{
    if (!c) return OFDMX_ERR_PARAM;
    c->emit_all = enable != 0;
    return OFDMX_OK;
}

int ofdmx_set_debug_taps(ofdmx_ctx *c, float *h_taps_dev, int64_t h_stride)
{
    if (!c || (h_taps_dev && h_stride < 1)) return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_set_debug_taps arguments");
    c->h_taps = (float2 *)h_taps_dev;
    c->h_stride = h_taps_dev ? h_stride : 0;
    return OFDMX_OK;
}

int ofdmx_header_len(const ofdmx_ctx *c) { return c ? c->hl : 0; }
int64_t ofdmx_launch_count(const ofdmx_ctx *c) { return c ? c->launches : 0; }

int64_t ofdmx_tx_frame_samples(const ofdmx_ctx *c, int64_t payload_bytes)
{
    if (!c || payload_bytes < 0) return -1;
    const int64_t lp = payload_bytes + (c->kp.crc_mode ? 4 : 0);
    const int64_t ns = (lp * 8 + c->kp.bps_p - 1) / c->kp.bps_p;
    return (int64_t)(c->kp.nsw + 1 + payload_ofdm_syms(c, (int)ns)) * c->kp.D + (c->kp.roll ? c->kp.roll - 1 : 0);
}

int ofdmx_reserve(ofdmx_ctx *c, int64_t n_streams, int64_t n_samples, int64_t max_frames)
{
    if (!c || n_streams < 1 || n_samples < 0 || max_frames < 1) return fail(c, OFDMX_ERR_PARAM, "bad reserve shape");
    if (int rc = check_device(c)) return rc;
    RxWs w = carve(nullptr, n_streams, n_samples, max_frames);
    return grow(c, c->ws, w.total);
}

int ofdmx_sync(ofdmx_ctx *c, const float *samples_dev, int64_t n_streams, int64_t n_samples, int64_t stride,
               int64_t *trig_out, float *cfo_out, int32_t *stream_out, int64_t max_trig,
               ofdmx_counts *counts_dev, void *stream)
{
    if (!c || !samples_dev || !counts_dev || n_streams < 1 || n_samples < 1 || max_trig < 1 || stride < n_samples)
        return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_sync arguments");
    if (max_trig > 0x7ffffff0LL || n_samples / 32 * n_streams > 0x7fffffffLL * 512)
        return fail(c, OFDMX_ERR_PARAM, "problem too large");
    if (int rc = check_device(c)) return rc;
    if (int rc = ofdmx_reserve(c, n_streams, n_samples, max_trig)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = use_stream(c, st)) return rc;
    RxWs w = carve(c->ws.p, n_streams, n_samples, max_trig);
    if (int rc = run_sync(c, w, (const float2 *)samples_dev, n_streams, n_samples, stride, max_trig, counts_dev, st)) return rc;
    // results are bounded by max_trig entries; copy whole arrays (stale tail entries are ignored by n_triggers)
    if (trig_out) CUDA_TRY(c, cudaMemcpyAsync(trig_out, w.trig, sizeof(long long) * (size_t)max_trig, cudaMemcpyDeviceToDevice, st));
    if (cfo_out) CUDA_TRY(c, cudaMemcpyAsync(cfo_out, w.cfo, sizeof(float) * (size_t)max_trig, cudaMemcpyDeviceToDevice, st));
    if (stream_out) CUDA_TRY(c, cudaMemcpyAsync(stream_out, w.trig_stream, sizeof(int) * (size_t)max_trig, cudaMemcpyDeviceToDevice, st));
    return OFDMX_OK;
}

extern "C++" {
static int rx_core(ofdmx_ctx *c, DevBuf &wsbuf, const float *samples_dev, int64_t n_streams, int64_t n_samples, int64_t stride,
                   ofdmx_frame *frames_out, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
                   float *z_out, int64_t z_stride, ofdmx_counts *counts_dev, cudaStream_t st);
}

int ofdmx_rx(ofdmx_ctx *c, const float *samples_dev, int64_t n_streams, int64_t n_samples, int64_t stride,
             ofdmx_frame *frames_out, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
             float *z_out, int64_t z_stride, ofdmx_counts *counts_dev, void *stream)
{
    if (!c) return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_rx arguments");
    return rx_core(c, c->ws, samples_dev, n_streams, n_samples, stride, frames_out, max_frames, bytes_out, byte_stride, z_out,
                   z_stride, counts_dev, (cudaStream_t)stream);
}

extern "C++" {
static int rx_core(ofdmx_ctx *c, DevBuf &wsbuf, const float *samples_dev, int64_t n_streams, int64_t n_samples, int64_t stride,
                   ofdmx_frame *frames_out, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
                   float *z_out, int64_t z_stride, ofdmx_counts *counts_dev, cudaStream_t st)
{
    if (!c || !samples_dev || !frames_out || !bytes_out || !counts_dev || n_streams < 1 || n_samples < 1
        || max_frames < 1 || stride < n_samples)
        return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_rx arguments");
    if (byte_stride < c->kp.max_pkt_bytes) return fail(c, OFDMX_ERR_CAPACITY, "byte_stride < max_pkt_bytes");
    if (z_out && z_stride < c->hl) return fail(c, OFDMX_ERR_CAPACITY, "z_stride < header_len");
    if (max_frames > 0x7ffffff0LL) return fail(c, OFDMX_ERR_PARAM, "max_frames too large");
    if (int rc = check_device(c)) return rc;
    if (int rc = grow(c, wsbuf, carve(nullptr, n_streams, n_samples, max_frames).total)) return rc;
    if (int rc = use_stream(c, st, &wsbuf != &c->ws_host)) return rc;
    RxWs w = carve(wsbuf.p, n_streams, n_samples, max_frames);
    const float2 *smp = (const float2 *)samples_dev;
    if (int rc = run_sync(c, w, smp, n_streams, n_samples, stride, max_frames, counts_dev, st)) return rc;
    ofdmx_ctx *ctx_ = c;
    KP kp_call = c->kp;
    kp_call.h_taps = c->h_taps;
    kp_call.h_stride = c->h_stride;
    if (c->h_taps && c->h_stride < c->kp.N) return fail(c, OFDMX_ERR_CAPACITY, "debug taps: stride < fft_len");
    // (the channel-tap debug output exists in the warp-per-frame kernel's tapped variant and in the any-fft_len kernel)
    const bool taps_generic = c->h_taps && !(c->frame1kw && z_out);
    if (c->framep && !c->force_generic && !c->no_warp_frame && !c->no_pair_frame && !(c->h_taps && !z_out)) {
        KT(K_FRAMEP);
        FwArgs fa{ kp_call, smp, n_samples, stride, w.trig, w.trig_stream, w.cfo, w.stream_start, w.n_trig, w.spec, bytes_out, byte_stride,
                   (float2 *)z_out, z_stride, c->x_2048, -1, 0 };
        if (!ofdmx_fp_launch(c->kp.bps_p, (unsigned)c->sm_count, c->framep_smem, st, fa, c->pair_tab, c->framep_hsz))
            return fail(c, OFDMX_ERR_PARAM, "no pair-of-warps frame kernel for this configuration");
    } else if (c->frame1kw && !c->force_generic && !c->no_warp_frame && !taps_generic) {
        KT(K_FRAME1KW);
        FwArgs fa{ kp_call, smp, n_samples, stride, w.trig, w.trig_stream, w.cfo, w.stream_start, w.n_trig, w.spec, bytes_out, byte_stride,
                   (float2 *)z_out, z_stride, c->x_2048, c->frame1kw_dec_off, c->frame1kw_dec_all };
        if (!ofdmx_fw_launch(c->kp.N, c->kp.bps_p, (unsigned)(c->sm_count * c->frame1kw_ctas), c->frame1kw_warps * 32, c->frame1kw_smem, st, fa))
            return fail(c, OFDMX_ERR_PARAM, "no warp-per-frame kernel for this configuration");
    } else if (c->frame1k_warps > 0 && !c->force_generic && !taps_generic) {
        KT(K_FRAME1K);
        const bool simple = (c->kp.n_occ_sets == 1 && c->kp.n_pil_sets <= 1 && !c->kp.pil_in_occ);
        F1kArgs fa{ c->kp, c->frame1k_warps, smp, n_samples, stride, w.trig, w.trig_stream, w.cfo, w.stream_start, w.n_trig, w.spec,
                    bytes_out, byte_stride, (float2 *)z_out, z_stride };
        ofdmx_f1k_launch(c->kp.bps_p, simple, (unsigned)(c->sm_count * 2), c->frame1k_smem, st, fa);
    } else {
        KT(K_FRAME);
        rx_frame_kernel<<<c->sm_count * 2, OFDMX_THREADS, c->frame_smem, st>>>(
            kp_call, smp, n_samples, stride, w.trig, w.trig_stream, w.cfo, w.stream_start, w.n_trig, w.spec, bytes_out,
            byte_stride, (float2 *)z_out, z_stride);
    }
    if (w.nblk == 1) {      // trigger capacity of the call fits one block: one launch instead of five
        KT(K_CHAIN_SMALL);
        chain_small_kernel<<<1, CH_T, 0, st>>>(c->kp, w.trig, w.trig_stream, w.spec, w.stream_start, w.n_trig, c->emit_all ? 1 : 0,
                                               counts_dev, frames_out);
        CUDA_TRY(c, cudaGetLastError());
        return OFDMX_OK;
    }
    { KT(K_CHAIN_NEXT); chain_next_kernel<<<w.nblk, CH_T, 0, st>>>(c->kp, w.trig, w.trig_stream, w.spec, w.stream_start, w.n_trig,
                                                                 w.jumpA, w.jumpB); }
    { KT(K_CHAIN_ENTRY); chain_entry_kernel<<<1, 256, 0, st>>>(w.n_trig, w.jumpB, w.entry, w.nblk); }
    { KT(K_CHAIN_MARK); chain_mark_kernel<<<w.nblk, CH_T, 0, st>>>(w.spec, w.n_trig, w.jumpA, w.entry, w.markA, w.blockcount, c->emit_all ? 1 : 0); }
    { KT(K_CHAIN_SCAN); chain_scan_kernel<<<1, 1024, 0, st>>>(w.blockcount, w.nblk, counts_dev); }
    { KT(K_CHAIN_EMIT); chain_emit_kernel<<<w.nblk, CH_T, 0, st>>>(w.spec, w.n_trig, w.markA, w.blockcount, frames_out); }
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}
}  // extern "C++"

int ofdmx_rx_host(ofdmx_ctx *c, const float *samples_host, int64_t n_streams, int64_t n_samples,
                  ofdmx_frame *frames_host, int64_t max_frames, uint8_t *bytes_host, int64_t byte_stride,
                  ofdmx_counts *counts_host)
{
    if (!c || !samples_host || !frames_host || !bytes_host || !counts_host || n_streams < 1 || n_samples < 1 || max_frames < 1
        || max_frames > 0x7ffffff0LL || n_streams > (1LL << 40) / n_samples)
        return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_rx_host arguments");
    if (byte_stride < c->kp.max_pkt_bytes) return fail(c, OFDMX_ERR_CAPACITY, "byte_stride < max_pkt_bytes");
    if (int rc = check_device(c)) return rc;
    const size_t sbytes = sizeof(float2) * (size_t)n_streams * (size_t)n_samples;
    if (int rc = grow(c, c->h_samples, sbytes)) return rc;
    if (int rc = grow(c, c->h_frames, sizeof(ofdmx_frame) * (size_t)max_frames)) return rc;
    if (int rc = grow(c, c->h_bytes, (size_t)max_frames * (size_t)byte_stride)) return rc;
    if (int rc = grow(c, c->h_counts, sizeof(ofdmx_counts))) return rc;
    cudaStream_t st = c->own_stream;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_samples.p, samples_host, sbytes, cudaMemcpyHostToDevice, st));
    // own workspace: this call runs on the context's private stream, next to whatever the caller has enqueued
    if (int rc = rx_core(c, c->ws_host, (const float *)c->h_samples.p, n_streams, n_samples, n_samples,
                         (ofdmx_frame *)c->h_frames.p, max_frames, (uint8_t *)c->h_bytes.p, byte_stride, nullptr, 0,
                         (ofdmx_counts *)c->h_counts.p, st))
        return rc;
    CUDA_TRY(c, cudaMemcpyAsync(counts_host, c->h_counts.p, sizeof(ofdmx_counts), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    c->n_host_syncs += 2;
    const int64_t nf = std::min<int64_t>(counts_host->n_frames, max_frames);
    const int64_t ntr = std::min<int64_t>(counts_host->n_triggers, max_frames);
    if (nf > 0)
        CUDA_TRY(c, cudaMemcpyAsync(frames_host, c->h_frames.p, sizeof(ofdmx_frame) * (size_t)nf, cudaMemcpyDeviceToHost, st));
    if (ntr > 0)
        CUDA_TRY(c, cudaMemcpyAsync(bytes_host, c->h_bytes.p, (size_t)ntr * (size_t)byte_stride, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return OFDMX_OK;
}

int ofdmx_tx(ofdmx_ctx *c, const uint8_t *payload_dev, const int64_t *pkt_off_dev, int64_t n_pkts,
             int32_t first_pkt_num, float *samples_out, int64_t cap_samples, int64_t *sample_off_dev, void *stream)
{
    if (!c || !payload_dev || !pkt_off_dev || !samples_out || !sample_off_dev || n_pkts < 0)
        return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_tx arguments");
    if (n_pkts == 0) return OFDMX_OK;
    if (int rc = check_device(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = use_stream(c, st)) return rc;
    ofdmx_ctx *ctx_ = c;
    { KT(K_TX_OFF); tx_offsets_kernel<<<1, 1024, 0, st>>>(c->kp, (const long long *)pkt_off_dev, n_pkts, (long long *)sample_off_dev); }
    if (c->tx1kw && !c->no_warp_tx && !c->force_generic) {
        const unsigned grid = (unsigned)std::min<int64_t>((n_pkts + TXW_WARPS - 1) / TXW_WARPS, (int64_t)c->sm_count);
        const int pbb = (int)tx1024w_pb_bytes(c->kp.max_pkt_bytes);
        KT(K_TX1KW);
        TxwArgs ta{ c->kp, payload_dev, (const long long *)pkt_off_dev, n_pkts, first_pkt_num, (float2 *)samples_out, cap_samples,
                    (const long long *)sample_off_dev, c->tx_map, c->sync_td, c->x_2048, pbb };
        if (!ofdmx_txw_launch(c->kp.N, c->kp.bps_p, grid, TXW_WARPS * 32, c->tx1kw_smem, st, ta))
            return fail(c, OFDMX_ERR_PARAM, "no warp-per-packet TX kernel for this configuration");
    } else {
        const unsigned grid = (unsigned)std::min<int64_t>(n_pkts, (int64_t)c->sm_count * 8);
        KT(K_TX);
        tx_frame_kernel<<<grid, OFDMX_THREADS, c->tx_smem, st>>>(c->kp, payload_dev, (const long long *)pkt_off_dev, n_pkts,
                                                              first_pkt_num, (float2 *)samples_out, cap_samples,
                                                              (const long long *)sample_off_dev);
    }
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}

int ofdmx_fft(ofdmx_ctx *c, const float *in_dev, float *out_dev, int64_t n_syms, int forward, void *stream)
{
    if (!c || !in_dev || !out_dev || n_syms < 0) return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_fft arguments");
    if (n_syms == 0) return OFDMX_OK;
    if (int rc = check_device(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = use_stream(c, st)) return rc;
    const unsigned grid = (unsigned)std::min<int64_t>(n_syms, (int64_t)c->sm_count * 8);
    const size_t sm = (size_t)c->kp.N * 8;
    ofdmx_ctx *ctx_ = c;
    KT(K_FFT);
    if (forward)
        fft_vcc_kernel<false><<<grid, OFDMX_THREADS, sm, st>>>((const float2 *)in_dev, (float2 *)out_dev, n_syms, c->kp.N, c->kp.logN, c->kp.tw);
    else
        fft_vcc_kernel<true><<<grid, OFDMX_THREADS, sm, st>>>((const float2 *)in_dev, (float2 *)out_dev, n_syms, c->kp.N, c->kp.logN, c->kp.tw);
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}

int ofdmx_crc32(ofdmx_ctx *c, const uint8_t *bytes_dev, const int64_t *pkt_off_dev, int64_t n_pkts,
                uint32_t *crc_out_dev, void *stream)
{
    if (!c || !bytes_dev || !pkt_off_dev || !crc_out_dev || n_pkts < 0) return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_crc32 arguments");
    if (n_pkts == 0) return OFDMX_OK;
    if (int rc = check_device(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = use_stream(c, st)) return rc;
    const unsigned grid = (unsigned)std::min<int64_t>(n_pkts, (int64_t)c->sm_count * 8);
    ofdmx_ctx *ctx_ = c;
    { KT(K_CRC); crc32_kernel<<<grid, OFDMX_THREADS, 0, st>>>(bytes_dev, (const long long *)pkt_off_dev, n_pkts, crc_out_dev,
                                                 c->kp.crc_tab, c->kp.crc_pow); }
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}

int ofdmx_agc2(ofdmx_ctx *c, const float *in_dev, float *out_dev, int64_t n_streams, int64_t n, int64_t stride,
               float attack, float decay, float reference, float max_gain, float *gain_io_dev, int32_t flags, void *stream)
{
    if (!c || !in_dev || !out_dev || !gain_io_dev || n_streams < 0 || n < 0 || stride < n || n_streams > (1 << 24))
        return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_agc2 arguments");
    if (n_streams == 0 || n == 0) return OFDMX_OK;
    if (int rc = check_device(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ofdmx_ctx *ctx_ = c;
    const int abs_rate = (flags & OFDMX_AGC2_ABS_RATE) ? 1 : 0;
    // Few streams: cut every stream into spans that run in parallel from a warm-up, accepted only when their entry
    // gain equals the predecessor's exit gain bit for bit (ofdmx_cond.cuh).  Needs separate in / out buffers (a
    // re-run reads the input again) and spans several warm-up lengths long.
    {
        // spans of >= 8 K samples, one per resident lane at most (the kernel is bound by the serial walk of a warp, so
        // one wave of short spans beats fewer long ones although the warm-up -- 10-15 K samples for a tracked signal --
        // is then longer than the span); the warm-up of a warp's rows is chosen on the device from the level in front of
        // them and capped at 4 spans
        long long min_span = 8192;
        if (const char *e = getenv("OFDMX_AGC_SPAN")) min_span = std::max(1024LL, atoll(e));
        const long long lanes = (long long)c->sm_count * 16 * 32;
        const bool disjoint = (out_dev + 2 * ((n_streams - 1) * stride + n) <= in_dev) || (in_dev + 2 * ((n_streams - 1) * stride + n) <= out_dev);
        long long spans = std::min<long long>(lanes / n_streams, n / min_span);
        if (disjoint && spans >= 8 && !c->no_agc_spans && 5 * ((n + spans - 1) / spans + 32) < 0x7fffffffLL) {   // span + warm-up as int
            const long long span = ((n + spans - 1) / spans + 31) / 32 * 32;
            const long long warm = 4 * span;
            spans = (n + span - 1) / span;
            const long long rows = n_streams * spans;
            const int rounds = 6;
            if (int rc = grow(c, c->ws_agc, (size_t)rows * 12 + 256)) return rc;
            float *entry = (float *)c->ws_agc.p, *exitg = entry + rows;
            int *need = (int *)(exitg + rows), *n_open = need + rows;
            // a warp takes 32 consecutive spans of one stream
            const unsigned grid = (unsigned)((n_streams * ((spans + 31) / 32) + AGC_WARPS - 1) / AGC_WARPS);
            const unsigned vgrid = (unsigned)std::min<long long>((rows + 255) / 256, (long long)c->sm_count * 4);
            CUDA_TRY(c, cudaMemsetAsync(n_open, 0, sizeof(int) * 16, st));
            { KT(K_AGC2); agc2_span_kernel<<<grid, AGC_WARPS * 32, 0, st>>>((const float2 *)in_dev, (float2 *)out_dev, n, stride, (int)n_streams,
                (int)spans, span, (int)warm, attack, decay, reference, max_gain, gain_io_dev, 1.0f, entry, exitg, need, 0, abs_rate); }
            for (int r = 0; r < rounds; r++) {
                { KT(K_AGC2_AUX); agc2_verify_kernel<<<vgrid, 256, 0, st>>>((int)n_streams, (int)spans, entry, exitg, need, n_open + r); }
                { KT(K_AGC2); agc2_span_kernel<<<grid, AGC_WARPS * 32, 0, st>>>((const float2 *)in_dev, (float2 *)out_dev, n, stride, (int)n_streams,
                    (int)spans, span, (int)warm, attack, decay, reference, max_gain, gain_io_dev, 1.0f, entry, exitg, need, 1, abs_rate); }
            }
            { KT(K_AGC2_AUX); agc2_verify_kernel<<<vgrid, 256, 0, st>>>((int)n_streams, (int)spans, entry, exitg, need, n_open + rounds); }
            { KT(K_AGC2_AUX); agc2_mopup_kernel<<<(unsigned)((n_streams + 63) / 64), 64, 0, st>>>((const float2 *)in_dev, (float2 *)out_dev, n, stride,
                (int)n_streams, (int)spans, span, attack, decay, reference, max_gain, entry, exitg, n_open + rounds, abs_rate); }
            { KT(K_AGC2_AUX); agc2_final_kernel<<<(unsigned)((n_streams + 255) / 256), 256, 0, st>>>((int)n_streams, (int)spans, exitg, gain_io_dev); }
            CUDA_TRY(c, cudaGetLastError());
            return OFDMX_OK;
        }
    }
    const unsigned grid = (unsigned)((n_streams + 32 * AGC_WARPS - 1) / (32 * AGC_WARPS));
    { KT(K_AGC2); agc2_kernel<<<grid, AGC_WARPS * 32, 0, st>>>((const float2 *)in_dev, (float2 *)out_dev, n, stride, (int)n_streams,
                                                      attack, decay, reference, max_gain, gain_io_dev,
                                                      (flags & OFDMX_AGC2_ABS_RATE) ? 1 : 0); }
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}

int64_t ofdmx_iir_state_doubles(void) { return IIR_STATE_DOUBLES; }

extern "C++" {
template <int MAXT>
static int iir_launch(ofdmx_ctx *c, const float *in_dev, float *out_dev, int64_t n_streams, int64_t n, int64_t stride,
                      const double *fftaps, int32_t n_ff, const double *fbtaps, int32_t n_fb, int64_t span,
                      double *state_io_dev, bool oldstyle, cudaStream_t st)
{
    iir_taps<MAXT> t;
    for (int i = 0; i < MAXT; i++) {
        t.ff[i] = i < n_ff ? fftaps[i] : 0.0;
        t.fb[i] = (i >= 1 && i < n_fb) ? (oldstyle ? fbtaps[i] : -fbtaps[i]) : 0.0;   // oldstyle=False: a[k] enter with a minus sign
    }
    // warm-up length: where the impulse response has decayed below 1e-18 of its peak (host, double)
    int64_t warm = 0;
    {
        double x[MAXT - 1] = { 0 }, y[MAXT - 1] = { 0 }, peak = 0.0;
        const int64_t cap = 1 << 20;
        int64_t last = 0;
        for (int64_t k = 0; k < cap; k++) {
            const double in = k == 0 ? 1.0 : 0.0;
            double acc = t.ff[0] * in;
            for (int i = 1; i < MAXT; i++) acc += t.ff[i] * x[i - 1];
            for (int i = 1; i < MAXT; i++) acc += t.fb[i] * y[i - 1];
            for (int i = MAXT - 2; i > 0; i--) { x[i] = x[i - 1]; y[i] = y[i - 1]; }
            x[0] = in; y[0] = acc;
            const double a = std::fabs(acc);
            if (!(a < 1e300)) return fail(c, OFDMX_ERR_PARAM, "ofdmx_iir_ccd: unstable filter");
            if (a > peak) peak = a;
            if (a > 1e-18 * peak) last = k;
            if (k - last > 64) break;
        }
        if (last >= cap - 65) return fail(c, OFDMX_ERR_PARAM, "ofdmx_iir_ccd: impulse response does not decay");
        warm = last + 1 + (MAXT - 1);
    }
    // span: caller's choice, else one full wave of lanes (12 resident warps per SM; the recurrence is a chain of
    // dependent double additions ~350 cycles per sample long, so more lanes is the only way to go faster until
    // the FP64 pipe is full), in several waves once a span would exceed 16 warm-up lengths
    if (span <= 0) {
        const int64_t capacity = (int64_t)c->sm_count * 12 * 32;
        int64_t sps = std::max<int64_t>(1, capacity / n_streams);
        span = (n + sps - 1) / sps;
        if (span > 16 * warm) {
            sps *= (span + 16 * warm - 1) / (16 * warm);
            span = (n + sps - 1) / sps;
        }
        span = std::max<int64_t>((span + 31) / 32 * 32, 1024);
    }
    if (out_dev == in_dev && span < n) return fail(c, OFDMX_ERR_PARAM, "ofdmx_iir_ccd: in-place needs span >= n");
    span = std::min<int64_t>(span, std::max<int64_t>(n, 1));
    if (span + warm >= (1ll << 31)) return fail(c, OFDMX_ERR_PARAM, "ofdmx_iir_ccd: span too long (2^31 samples with the warm-up)");
    const int64_t sps = (n + span - 1) / span;
    if (sps > (1 << 30)) return fail(c, OFDMX_ERR_PARAM, "ofdmx_iir_ccd: too many spans");
    const int64_t lanes = n_streams * sps;
    const unsigned grid = (unsigned)((lanes + 32 * IIR_WARPS - 1) / (32 * IIR_WARPS));
    ofdmx_ctx *ctx_ = c;
    { KT(K_IIR); iir_ccd_kernel<MAXT><<<grid, IIR_WARPS * 32, 0, st>>>((const float2 *)in_dev, (float2 *)out_dev, n, stride, (int)n_streams,
                                                              (int)sps, span, warm, t, state_io_dev); }
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}
}  // extern "C++"

int ofdmx_iir_ccd(ofdmx_ctx *c, const float *in_dev, float *out_dev, int64_t n_streams, int64_t n, int64_t stride,
                  const double *fftaps, int32_t n_ff, const double *fbtaps, int32_t n_fb, int64_t span,
                  double *state_io_dev, int32_t flags, void *stream)
{
    const bool oldstyle = (flags & OFDMX_IIR_OLDSTYLE) != 0;
    if (!c || !in_dev || !out_dev || !state_io_dev || !fftaps || n_streams < 0 || n < 0 || stride < n ||
        n_streams > (1 << 24) || n_ff < 1 || n_ff > 17 || n_fb < 0 || n_fb > 17 || (n_fb > 0 && !fbtaps))
        return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_iir_ccd arguments (at most 17 feed-forward and 17 feedback taps)");
    if (n_streams == 0 || n == 0) return OFDMX_OK;
    if (int rc = check_device(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int m = std::max(n_ff, n_fb);
    if (m <= 9) return iir_launch<9>(c, in_dev, out_dev, n_streams, n, stride, fftaps, n_ff, fbtaps, n_fb, span, state_io_dev, oldstyle, st);
    if (m <= 13) return iir_launch<13>(c, in_dev, out_dev, n_streams, n, stride, fftaps, n_ff, fbtaps, n_fb, span, state_io_dev, oldstyle, st);
    return iir_launch<17>(c, in_dev, out_dev, n_streams, n, stride, fftaps, n_ff, fbtaps, n_fb, span, state_io_dev, oldstyle, st);
}

int ofdmx_papr(ofdmx_ctx *c, const float *in_dev, int64_t n, float *out3_dev, void *stream)
{
    if (!c || !in_dev || !out3_dev || n <= 0) return fail(c, OFDMX_ERR_PARAM, "bad ofdmx_papr arguments");
    if (int rc = check_device(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int parts = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
    if (int rc = grow(c, c->ws_papr, (size_t)c->sm_count * 8 * 16)) return rc;
    double *ps = (double *)c->ws_papr.p;
    float *pp = (float *)(ps + parts);
    ofdmx_ctx *ctx_ = c;
    { KT(K_PAPR); papr_kernel<<<parts, 256, 0, st>>>((const float2 *)in_dev, n, ps, pp); }
    papr_final_kernel<<<1, 32, 0, st>>>(ps, pp, parts, n, out3_dev);
    CUDA_TRY(c, cudaGetLastError());
    return OFDMX_OK;
}

int ofdmx_reconfigure(ofdmx_ctx **ctx_io, const ofdmx_params *prm)
{
    if (!ctx_io || !*ctx_io || !prm) return fail(ctx_io ? *ctx_io : nullptr, OFDMX_ERR_PARAM, "bad ofdmx_reconfigure arguments");
    ofdmx_ctx *c = *ctx_io;
    if (int rc = check_device(c)) return rc;
    PlanFields np;
    Stage stg;
    if (int rc = build_plan(prm, &np, stg)) { c->err = g_err; return rc; }     // the running plan is untouched
    if (int rc = commit_plan(c, np, stg, false)) return rc;
    c->n_reconfigs++;
    return OFDMX_OK;
}

int64_t ofdmx_counter(const ofdmx_ctx *c, int which)
{
    if (!c) return -1;
    switch (which) {
    case OFDMX_CNT_LAUNCHES: return c->launches;
    case OFDMX_CNT_DEVICE_ALLOCS: return c->n_dev_allocs;
    case OFDMX_CNT_HOST_SYNCS: return c->n_host_syncs;
    case OFDMX_CNT_RECONFIGS: return c->n_reconfigs;
    default: return -1;
    }
}

}  // extern "C"

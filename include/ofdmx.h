/*
 * ofdmx.h -- C ABI of the B200-native OFDM physical layer (libofdmx.so).
 *
 * This is the drop-in boundary for the sample-path blocks that gr-ofdm_tools' hier blocks wire
 * together (citations relative to the reference tree):
 *   python/ofdm_txrx_modules.py:120-254  ofdm_tx   (header gen, scrambler, repack, mapper, mux,
 *                                                   carrier allocator, IFFT, cyclic prefixer)
 *   python/ofdm_txrx_modules.py:257-426  ofdm_rx   (S&C sync, delay+NCO+mixer, header/payload
 *                                                   demux, FFT, chanest, DFE equaliser, serializer,
 *                                                   decoder, repack, descrambler)
 *   python/ofdm_radio_hier.py:92-244     the same chain flattened, plus crc32_bb and selectors
 *   python/ofdm_tx_rx_hier.py:55-87      ofdm_tx -> x0.01 ; ofdm_rx
 * The reference has no native/FFI layer of its own (lib/CMakeLists.txt:27-34 is empty): its
 * "FFI" for this path is the GNU Radio block API (io signatures of uint8 / gr_complex items plus
 * "packet_len" stream tags).  The entry points below are what a gr.basic_block adaptor binds;
 * INTEGRATION.md shows that adaptor.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (a CUDA stream is passed as void*).
 *   - every *_dev pointer is device memory owned by the caller; the library owns only the opaque
 *     context and its internal workspace (grown on demand, never on a steady-state call).
 *   - complex samples are interleaved float32 (re,im) == gr_complex; frequency-domain vectors are
 *     in shifted order (DC at fft_len/2) as in the reference (python/ofdm_txrx_modules.py:28-29).
 *   - packet delimitation: CSR-style int64 offsets replace "packet_len" tags.
 *   - all calls return 0 on success or a negative ofdmx_status; ofdmx_last_error() gives text.
 *   - a context is single-owner (not thread-safe); different contexts may run concurrently.
 *   - there is no CPU fallback: every entry point fails with OFDMX_ERR_CUDA if no sm_100 device.
 */
#ifndef OFDMX_H
#define OFDMX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDMX_ABI_VERSION 4

typedef enum {
    OFDMX_OK = 0,
    OFDMX_ERR_PARAM = -1,     /* invalid parameter (e.g. sync word length != fft_len) */
    OFDMX_ERR_CUDA = -2,      /* CUDA runtime error / no device */
    OFDMX_ERR_CAPACITY = -3,  /* caller buffer too small */
    OFDMX_ERR_NOMEM = -4
} ofdmx_status;

typedef struct ofdmx_ctx ofdmx_ctx;

/* Constructor parameters of ofdm_tx / ofdm_rx (python/ofdm_txrx_modules.py:143-155,278-291) and of
 * ofdm_radio_hier (python/ofdm_radio_hier.py:34-39), flattened.  All pointers are HOST memory and
 * are copied by ofdmx_create. */
typedef struct {
    int32_t fft_len;                /* power of two, 32..4096 (every power of two in that range has a parity test) */
    int32_t cp_len;
    int32_t n_occ_sets;             /* occupied_carriers: set-major flat list of carrier numbers */
    const int32_t *occ_sizes;
    const int32_t *occ_carriers;
    int32_t n_pilot_sets;           /* pilot_carriers */
    const int32_t *pilot_sizes;
    const int32_t *pilot_carriers;
    int32_t n_pilot_sym_sets;       /* pilot_symbols (re,im interleaved) */
    const int32_t *pilot_sym_sizes;
    const float *pilot_symbols;
    const float *sync_word1;        /* fft_len x (re,im), shifted order */
    const float *sync_word2;        /* NULL: the single-sync-word mode, sync_word2=() in the reference constructors
                                       (python/ofdm_txrx_modules.py:174-183,311-321): one preamble symbol, two OFDM
                                       symbols before the payload, channel taps from sync word 1 (interpolated) */
    int32_t bps_header;             /* 1 BPSK, 2 QPSK, 3 8PSK, 4 16-QAM, 6 64-QAM (extension) */
    int32_t bps_payload;
    int32_t scramble_header;        /* packet_header_ofdm(scramble_header=...) */
    int32_t scramble_seed;          /* additive_scrambler seed: 0x7f = on, 0x00 = off */
    int32_t crc_mode;               /* 1: in-graph digital.crc32_bb on TX and RX */
    float   threshold;              /* ofdm_sync_sc_cfb plateau threshold (0.9) */
    int32_t max_carr_offset;        /* ofdm_chanest_vcvc max_carr_offset; -1 = unlimited */
    float   alpha;                  /* ofdm_equalizer_simpledfe alpha (0.1) */
    float   tx_scale;               /* multiply_const_vcc after the prefixer (0.01 in the hier blocks) */
    int32_t demux_holdoff;          /* header_payload_demux: items left unconsumed after a payload:
                                       fft_len+cp_len (GNU Radio < 3.7.10) or 1 (>= 3.7.10) */
    int32_t max_pkt_bytes;          /* largest header length field the caller sizes slots for (<=4095) */
    float   tx_clip;                /* ofdm_tools.clipper(clipping_factor) fused after the scaling: re and im
                                       railed to +-tx_clip (python/clipper.py:45-58,
                                       python/ofdm_radio_hier.py:92,229,239); 0 = no clipper */
    int32_t rolloff;                /* ofdm_cyclic_prefixer(fft_len, fft_len+cp_len, rolloff, key)
                                       (python/ofdm_txrx_modules.py:247-253; cp_len/4 in python/ofdm_cr_tools.py:1093):
                                       raised-cosine flanks of rolloff-1 samples; every burst grows by rolloff-1
                                       samples (the flushed down flank of its last symbol).  0 or 1 = rectangular;
                                       must not exceed cp_len */
    /* ---- GNU Radio version switches (every semantic the survey marked "(?)") ---- */
    int32_t qam_normalization;      /* constellation_rect normalisation of the 16-/64-QAM tables: 0 = none (GNU Radio 3.7,
                                       the reference's generation: mean power 10/9), 1 = AMPLITUDE_NORMALIZATION (the
                                       default from 3.8 on: points and sector widths scaled by n / sum|p|) */
    int32_t reserved[7];            /* must be zero */
} ofdmx_params;

/* Per-frame record (replaces the stream tags / PMT header dict of the reference). 32 bytes. */
typedef struct {
    int64_t  trigger;       /* index of the trigger item in the stream (S&C output coordinates) */
    float    cfo;           /* fine frequency estimate arg(P) at the trigger [rad] */
    int32_t  stream;        /* stream index */
    uint32_t flags;         /* OFDMX_F_* */
    uint16_t pkt_len;       /* header length field: bytes incl. in-graph CRC */
    uint16_t pkt_num;       /* header packet counter */
    uint16_t frame_syms;    /* payload OFDM symbols */
    int16_t  carr_offset;   /* integer carrier offset (ofdm_sync_carr_offset tag) */
    uint32_t slot;          /* payload bytes are at bytes_out_dev + slot*byte_stride */
} ofdmx_frame;

#define OFDMX_F_HDR_OK   1u   /* header CRC-8 matched */
#define OFDMX_F_CRC_OK   2u   /* payload CRC-32 matched (always set when crc_mode == 0) */
#define OFDMX_F_COMPLETE 4u   /* whole frame lies inside the buffer */
#define OFDMX_F_ACCEPTED 8u   /* the header/payload demux would have examined this trigger */
#define OFDMX_F_HDR_SEEN 16u  /* the 3 header-side symbols lie inside the buffer */
#define OFDMX_F_OVERSIZE 32u  /* header ok, payload samples present, but pkt_len > max_pkt_bytes: the demux
                                 consumes the frame (the search resumes behind it) and the record is emitted,
                                 but the payload is not decoded (no slot content, OFDMX_F_CRC_OK never set) */

typedef struct {
    int32_t n_triggers;     /* plateau-detector triggers found (all streams) */
    int32_t n_frames;       /* frame records written */
    int32_t overflow;       /* nonzero: n_triggers exceeded max_frames, output truncated */
    int32_t reserved;
} ofdmx_counts;

/* ---- lifetime ---- */
int  ofdmx_abi_version(void);
int  ofdmx_params_size(void);      /* sizeof(ofdmx_params) / sizeof(ofdmx_frame) as this library was compiled: an adaptor */
int  ofdmx_frame_size(void);       /* checks its own struct declarations against them before the first call */
int  ofdmx_create(const ofdmx_params *params, int device, ofdmx_ctx **ctx_out);
void ofdmx_destroy(ofdmx_ctx *ctx);
const char *ofdmx_last_error(const ofdmx_ctx *ctx);   /* ctx may be NULL (creation errors) */
/* pre-size the internal workspace so that later calls of this shape do not allocate */
int  ofdmx_reserve(ofdmx_ctx *ctx, int64_t n_streams, int64_t n_samples, int64_t max_frames);
int  ofdmx_header_len(const ofdmx_ctx *ctx);          /* items in the header = len(occupied_carriers[0]) */
int64_t ofdmx_tx_frame_samples(const ofdmx_ctx *ctx, int64_t payload_bytes);
int64_t ofdmx_launch_count(const ofdmx_ctx *ctx);     /* kernels launched by this context so far */

/* ---- TX: replaces ofdm_tx (+ crc32_bb + x tx_scale) ----
 * payload_dev[pkt_offsets[i] .. pkt_offsets[i+1]) is packet i.  Packet i is written to
 * samples_out_dev[sample_offsets_dev[i] .. sample_offsets_dev[i+1]) (offsets computed on device). */
int ofdmx_tx(ofdmx_ctx *ctx, const uint8_t *payload_dev, const int64_t *pkt_offsets_dev,
             int64_t n_pkts, int32_t first_pkt_num, float *samples_out_dev, int64_t cap_samples,
             int64_t *sample_offsets_dev /* n_pkts+1 */, void *cuda_stream);

/* ---- RX: replaces ofdm_rx (+ crc32_bb check) ----
 * samples_dev: n_streams rows of n_samples complex items, row stride stream_stride (items).
 * frames_out_dev: records of the frames the demux accepts with a valid header, ordered by
 * (stream, trigger).  bytes_out_dev: max_frames slots of byte_stride bytes.  z_out_dev (optional
 * debug tap, may be NULL): pre-decision equalised symbols, header_len + payload symbols per slot.
 * counts_dev: one ofdmx_counts. */
int ofdmx_rx(ofdmx_ctx *ctx, const float *samples_dev, int64_t n_streams, int64_t n_samples,
             int64_t stream_stride, ofdmx_frame *frames_out_dev, int64_t max_frames,
             uint8_t *bytes_out_dev, int64_t byte_stride, float *z_out_dev, int64_t z_stride,
             ofdmx_counts *counts_dev, void *cuda_stream);

/* enable != 0: ofdmx_rx writes a record for EVERY plateau trigger (header fields and flags as decoded
 * speculatively; OFDMX_F_ACCEPTED marks the triggers the demux examined) instead of only the accepted
 * frames.  Used to resolve the demux state exactly across the seams of a stream split over several GPUs. */
int ofdmx_set_emit_all(ofdmx_ctx *ctx, int enable);

/* Per-stage debug taps (the reference taps its chain with file sinks behind debug_log, python/ofdm_cr_tools.py:1297-1298,
 * 1346-1348,1414-1416,1461-1467,1513-1520).  Stage by stage, what this library offers:
 *   tx signal                      -> the output of ofdmx_tx;   post-allocator symbols -> ofdmx_fft(forward) of it
 *   sync detect / frequency offset -> ofdmx_sync (trigger indices, arg(P) per trigger)
 *   channel estimate               -> ofdmx_set_debug_taps: ofdmx_rx then writes the taps ofdm_chanest_vcvc hands to the
 *                                     equaliser (tag ofdm_sync_chan_taps: fft_len complex, shifted order) for every
 *                                     trigger whose header was seen, at h_taps_dev[slot * h_stride ..), up to one unit
 *                                     phasor per frame (the NCO phase reference restarts at every trigger; it cancels
 *                                     in y / H); zero the buffer
 *                                     first -- the warp-per-frame kernel fills the occupied carriers only (it runs when
 *                                     z_out is given as well; otherwise the any-fft_len kernel serves the call)
 *   pre-decision symbols (post-eq) -> z_out of ofdmx_rx;   integer offset / header fields -> ofdmx_set_emit_all records
 *   post-demod bytes               -> bytes_out of ofdmx_rx
 * h_taps_dev == NULL switches the tap off.  h_stride in complex items, >= fft_len. */
int ofdmx_set_debug_taps(ofdmx_ctx *ctx, float *h_taps_dev, int64_t h_stride);

/* Same call with HOST buffers (pinned or pageable): copies in, runs, copies records, counts and
 * the used payload slots back, and synchronises.  This is the call a GNU Radio work() makes. */
int ofdmx_rx_host(ofdmx_ctx *ctx, const float *samples_host, int64_t n_streams, int64_t n_samples,
                  ofdmx_frame *frames_host, int64_t max_frames, uint8_t *bytes_host,
                  int64_t byte_stride, ofdmx_counts *counts_host);

/* ---- Schmidl & Cox only: replaces digital.ofdm_sync_sc_cfb + plateau detector ----
 * trig_out_dev / cfo_out_dev / stream_out_dev: up to max_trig triggers ordered by (stream, index). */
int ofdmx_sync(ofdmx_ctx *ctx, const float *samples_dev, int64_t n_streams, int64_t n_samples,
               int64_t stream_stride, int64_t *trig_out_dev, float *cfo_out_dev,
               int32_t *stream_out_dev, int64_t max_trig, ofdmx_counts *counts_dev,
               void *cuda_stream);

/* ---- per-kernel timing (CUDA events recorded on the caller's stream around every kernel) ----
 * ofdmx_profile(ctx, 1) starts recording and clears the totals; ofdmx_profile_read synchronises the
 * recorded events and adds them to the totals: ms_total[i] / calls[i] for kernel slot i <
 * ofdmx_profile_slots(); ofdmx_profile_name(i) names the slot. */
int ofdmx_profile(ofdmx_ctx *ctx, int enable);
int ofdmx_profile_slots(void);
const char *ofdmx_profile_name(int slot);
int ofdmx_profile_read(ofdmx_ctx *ctx, float *ms_total, int64_t *calls);

/* ---- single blocks, exported for block-level parity tests and reuse ---- */
/* fft.fft_vcc(fft_len, forward, (), shift=True): n_syms vectors of fft_len items */
int ofdmx_fft(ofdmx_ctx *ctx, const float *in_dev, float *out_dev, int64_t n_syms, int forward,
              void *cuda_stream);
/* digital.crc32_bb arithmetic: zlib CRC-32 of each packet */
int ofdmx_crc32(ofdmx_ctx *ctx, const uint8_t *bytes_dev, const int64_t *pkt_offsets_dev,
                int64_t n_pkts, uint32_t *crc_out_dev, void *cuda_stream);

/* ---- conditioning around the chain ---- */
/* analog.agc2_cc(attack, decay, reference, gain) + set_max_gain(max_gain), the block both hier surfaces put in
 * front of ofdm_rx (python/ofdm_tx_rx_hier.py:75-76,82-83; python/ofdm_radio_hier.py:180-181).  The gain
 * recurrence is sequential per stream; streams run in parallel.  gain_io_dev[n_streams]: loop gain before the
 * first sample on entry, after the last sample on return, so consecutive calls continue the stream exactly.
 * in_dev == out_dev is allowed.  stride in complex samples. */
int ofdmx_agc2(ofdmx_ctx *ctx, const float *in_dev, float *out_dev, int64_t n_streams, int64_t n,
               int64_t stride, float attack, float decay, float reference, float max_gain,
               float *gain_io_dev, int32_t flags, void *cuda_stream);
#define OFDMX_AGC2_ABS_RATE 1   /* rate = attack when fabsf(tmp) > gain (agc2.h from GNU Radio 3.8 on); default: tmp > gain (3.7) */

/* ---- TX conditioning and the PAPR probe (rank 4) ---- */
/* filter.iir_filter_ccd(fftaps, fbtaps, oldstyle=False), the out-of-band filter ofdm_radio_hier puts behind the
 * TX chain when filter_mode=1 (python/ofdm_radio_hier.py:83-84,93,232-237; python/sync_radio_hier.py:73,165):
 * complex float in/out, double taps, complex double accumulator, y[n] = sum ff[i] x[n-i] - sum_{j>=1} fb[j] y[n-j]
 * (fb[0] is not used, as in GNU Radio).  At most 17 + 17 taps (the reference filters: 9 + 9 taps in
 * ofdm_radio_hier, 13 + 13 in sync_radio_hier).  Each stream is
 * cut into spans of `span` samples (<= 0: chosen by the library) that run in parallel, each warmed up from a
 * zero state over the length the impulse response needs to fall below 1e-18 of its peak; the first span starts
 * from state_io and is bit-exact against the sequential filter, later spans agree to the round-off noise of the
 * direct-form recurrence (~1e-10 relative in the accumulator for the reference taps).  state_io_dev[n_streams][ofdmx_iir_state_doubles()] doubles (zeros = a fresh block) holds the
 * filter history before the first sample on entry and after the last one on return.  In place only when
 * span >= n.  stride in complex samples. */
int64_t ofdmx_iir_state_doubles(void);
int ofdmx_iir_ccd(ofdmx_ctx *ctx, const float *in_dev, float *out_dev, int64_t n_streams, int64_t n,
                  int64_t stride, const double *fftaps, int32_t n_ff, const double *fbtaps, int32_t n_fb,
                  int64_t span, double *state_io_dev, int32_t flags, void *cuda_stream);
#define OFDMX_IIR_OLDSTYLE 1    /* iir_filter_ccd(..., oldstyle=True): feedback taps enter with a plus sign as given */

/* papr_sink.level() (python/papr_sink.py:46-54) of one block of n complex samples:
 * out3_dev[0] = max|x|^2 / mean|x|^2, out3_dev[1] = max|x|^2, out3_dev[2] = mean|x|^2. */
int ofdmx_papr(ofdmx_ctx *ctx, const float *in_dev, int64_t n, float *out3_dev, void *cuda_stream);

/* Run-time reconfiguration: what cognitive_engine_mac does to the radios when
 * spectrum_enforcer hands it a new carrier plan and sync words (python/cognitive_engine_mac.py:278-285,
 * python/ofdm_cr_tools.py:348-378).  Swaps the PHY tables of the context at a call boundary WITHOUT draining the
 * device: the new table image is built on the host, uploaded on a side stream into the half of the context's table
 * arena that the running plan does not use, and the next call makes its stream wait for that upload (an event).
 * Work already enqueued keeps running on the old tables.  No cudaDeviceSynchronize; no allocation once the arena
 * (sized with head-room at creation) holds the plan; everything that does not depend on the parameters (workspace,
 * staging buffers, streams, profiling state, counters) is kept.  The handle stays the same (*ctx_io is kept for
 * source compatibility).  On failure the running plan is untouched. */
int ofdmx_reconfigure(ofdmx_ctx **ctx_io, const ofdmx_params *prm);

/* Life-time counters of a context (ofdmx_launch_count is OFDMX_CNT_LAUNCHES): lets a caller -- and the tests --
 * assert that a steady-state call or a reconfiguration neither allocated nor blocked the host. */
#define OFDMX_CNT_LAUNCHES      0   /* kernels launched */
#define OFDMX_CNT_DEVICE_ALLOCS 1   /* cudaMalloc / cudaHostAlloc calls */
#define OFDMX_CNT_HOST_SYNCS    2   /* host-blocking synchronisations (ofdmx_rx_host: 2 per call by design) */
#define OFDMX_CNT_RECONFIGS     3   /* successful ofdmx_reconfigure calls */
int64_t ofdmx_counter(const ofdmx_ctx *ctx, int which);

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/ofdm_oracle.h -- CPU restatement of the gr-ofdm_tools OFDM PHY hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (gr-ofdm_tools_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY STATUS: "parity partially pinned".  The reference (/root/reference) contains no
 * arithmetic of its own on this path: it wires stock GNU Radio 3.7 blocks
 * (python/ofdm_txrx_modules.py:189-254,324-426; python/ofdm_radio_hier.py:92-244), and GNU
 * Radio (un-vendored dependency, >= 3.7.2 per CMakeLists.txt:112-113) is absent from this
 * image.  The restatement follows the published GNU Radio 3.7 block semantics (SURVEY.md
 * Appendix A) and is pinned against the only known-answer data the reference ships: the
 * sync-word literals in apps/ofdm_rx_hier.grc, python/ofdm_radio_hier.py:37-38,
 * python/sync_radio_hier.py:50-51, the default carrier plan python/ofdm_radio_hier.py:34-35,
 * and the standard CRC check values.  Everything else is pinned only by derived vectors
 *  and round-trip properties.
 *
 * Arithmetic convention: every floating-point stage is evaluated in float64 from the float32
 * input samples (the order-independent mathematical value of GNU Radio's float32 chain, whose
 * own result depends on the VOLK kernel's summation order).  Integer/bit stages are exact.
 */
#ifndef OFDM_ORACLE_H
#define OFDM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t fft_len;            /* N, power of two >= 16 */
    int32_t cp_len;
    /* carrier plan: raw carrier numbers (negative allowed), set-major flat arrays
     * (python/ofdm_txrx_modules.py:54-62) */
    int32_t n_occ_sets;
    const int32_t *occ_sizes;
    const int32_t *occ_carriers;
    int32_t n_pilot_sets;
    const int32_t *pilot_sizes;
    const int32_t *pilot_carriers;
    int32_t n_pilot_sym_sets;
    const int32_t *pilot_sym_sizes;
    const float *pilot_symbols;     /* re,im interleaved */
    const float *sync_word1;        /* N x (re,im), shifted order (DC at N/2) */
    const float *sync_word2;
    int32_t bps_header;             /* 1,2,3,4 (6 = 64-QAM extension) */
    int32_t bps_payload;
    int32_t scramble_header;        /* packet_header_ofdm scramble_header flag */
    int32_t scramble_seed;          /* additive_scrambler seed: 0x7f on, 0x00 off */
    int32_t crc_mode;               /* 1: digital.crc32_bb in graph (ofdm_radio_hier.py:121-122) */
    float   threshold;              /* plateau threshold, 0.9 */
    int32_t max_carr_offset;        /* ofdm_chanest_vcvc max_carr_offset, -1 = unlimited */
    float   alpha;                  /* simpledfe alpha, 0.1 */
    float   tx_scale;               /* multiply_const after the cyclic prefixer; 1.0 for bare ofdm_tx */
    int32_t demux_holdoff;          /* items left unconsumed after a payload:
                                       fft_len+cp_len (GNU Radio < 3.7.10) or 1 (>= 3.7.10) */
    float   tx_clip;                /* ofdm_tools.clipper(clipping_factor) after the scaling: re and im railed to
                                       +-tx_clip (python/clipper.py:45-58, ofdm_radio_hier.py:92,229); 0 = off */
    int32_t rolloff;                /* ofdm_cyclic_prefixer rolloff_len (python/ofdm_txrx_modules.py:247-253);
                                       0 or 1 = rectangular */
    int32_t qam_normalization;      /* 16-/64-QAM constellation_rect: 0 = no normalisation (GNU Radio 3.7), 1 =
                                       AMPLITUDE_NORMALIZATION (default from 3.8 on) */
} orc_params;

typedef struct {
    int64_t  trigger;       /* index of the trigger item (S&C output / delayed-stream coordinates) */
    float    cfo;           /* fine CFO estimate arg(P) at the trigger [rad] */
    int32_t  carr_offset;   /* integer carrier offset from ofdm_chanest_vcvc */
    uint32_t flags;         /* bit0 header CRC-8 ok, bit1 CRC-32 ok (set when crc_mode==0 too),
                               bit2 frame complete in buffer, bit3 accepted by the demux */
    uint16_t pkt_len;       /* header length field: payload bytes incl. in-graph CRC */
    uint16_t pkt_num;       /* header counter field */
    uint32_t frame_syms;    /* payload OFDM symbols */
    uint32_t slot;          /* payload slot index (ordinal of this frame in the output) */
} orc_frame;

#define ORC_F_HDR_OK   1u
#define ORC_F_CRC_OK   2u
#define ORC_F_COMPLETE 4u
#define ORC_F_ACCEPTED 8u
#define ORC_F_HDR_SEEN 16u   /* orc_rx_all only: the 3 header-side symbols lie inside the buffer */

/* ---- integer primitives ---- */
uint32_t orc_crc32(const uint8_t *buf, int64_t len);                 /* zlib CRC-32 (crc32_bb) */
uint8_t  orc_crc8(const uint8_t *buf, int64_t len);                  /* poly 0x07 init 0xFF */
void     orc_lfsr_bits(uint32_t mask, uint32_t seed, uint32_t reg_len, uint8_t *bits, int64_t n);
void     orc_scramble(uint8_t *buf, int64_t len, uint32_t seed);     /* additive_scrambler 0x8a/7 */
int64_t  orc_repack(const uint8_t *in, int64_t n_in, int k, int l, int align_output, uint8_t *out);
int      orc_header_len(const orc_params *p);
void     orc_header_format(const orc_params *p, int pkt_len, int pkt_num, uint8_t *out);
int      orc_header_parse(const orc_params *p, const uint8_t *in, int *pkt_len_bytes, int *pkt_num,
                          int *pkt_syms, int *frame_syms);
int      orc_constellation(int bps, float *points /* 2^bps x (re,im) */);
int      orc_decide(int bps, double re, double im);
/* the same with the constellation_rect normalisation switch (0 none, 1 amplitude) */
int      orc_constellation_n(int bps, int norm, float *points);
int      orc_decide_n(int bps, int norm, double re, double im);
/* debug tap: orc_rx (with z_out given) also writes the channel taps ofdm_chanest_vcvc hands to the equaliser, fft_len
 * complex (shifted order) per emitted frame ordinal; NULL switches it off.  Not thread-safe (test use). */
void     orc_set_taps_out(float *taps);
int      orc_rx_all(const orc_params *p, const float *r, int64_t n_samp, orc_frame *recs, int64_t max_recs,
                    uint8_t *bytes_out, int64_t byte_stride, int64_t *n_recs);
/* OpenMP team size used by the parallel loops (0: leave as is); returns the team size in effect.  The
 * reference arm sets it explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers. */
int      orc_set_threads(int n);

/* ---- float primitives ---- */
void orc_fft(int n, int forward, const double *in, double *out);   /* unnormalised DFT, no shift */

/* TX: bytes -> samples.  Returns 0 or <0. sample_off has n_pkts+1 entries. */
int orc_tx(const orc_params *p, const uint8_t *payload, const int64_t *pkt_off, int64_t n_pkts,
           int32_t first_pkt_num, float *samples_out, int64_t cap_samples, int64_t *sample_off);
int64_t orc_tx_frame_samples(const orc_params *p, int64_t payload_bytes);

/* Schmidl & Cox: exact (float64) metric, plateau detector, fine CFO at each trigger. */
int orc_sync(const orc_params *p, const float *samples, int64_t n, uint8_t *detect /* n or NULL */,
             int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig);
/* float32 FIR port of ofdm_sync_sc_cfb as GNU Radio evaluates it (for CPU-baseline timing only) */
int orc_sync_f32(const orc_params *p, const float *samples, int64_t n,
                 int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig);

/* Full RX chain.  recs: one per frame the demux accepted with a valid header and a complete
 * payload (in stream order).  bytes_out slot i at i*byte_stride.  z_out (optional): equalised
 * pre-decision symbols, header_len + pkt_syms complex per frame at i*z_stride (complex units). */
int orc_rx(const orc_params *p, const float *samples, int64_t n,
           orc_frame *recs, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
           float *z_out, int64_t z_stride, int64_t *n_frames,
           int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig);

/* same chain with the float32 FIR sync port in front: used only to time the CPU baseline */
int orc_rx_baseline(const orc_params *p, const float *samples, int64_t n,
                    orc_frame *recs, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
                    int64_t *n_frames, int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig);

/* ---- next rows of SURVEY.md 8(f) ---- */
/* analog.agc2_cc(attack, decay, reference, gain) with set_max_gain (python/ofdm_tx_rx_hier.py:75-76,
 * python/ofdm_radio_hier.py:180-181): one stream, float32 arithmetic in GNU Radio's order.
 * *gain: loop gain before the first sample on entry, after the last sample on return. */
void orc_agc2(const float *in, float *out, int64_t n, float attack, float decay, float reference,
              float max_gain, float *gain);
void orc_agc2_v(const float *in, float *out, int64_t n, float attack, float decay, float reference,
                float max_gain, float *gain, int abs_rate);
/* digital.crc32() of gr-digital/lib/crc32.cc as used by digital.crc.gen_and_append_crc32
 * (examples/benchmarks.py:347, python/ofdm_cr_tools.py:1760): MSB-first, poly 0x04C11DB7,
 * init and final XOR 0xFFFFFFFF; check value 0xFC891918. */
void orc_iir_ccd(const float *in, float *out, int64_t n, const double *ff, int n_ff, const double *fb, int n_fb,
                 double *state);
void orc_iir_ccd_v(const float *in, float *out, int64_t n, const double *ff, int n_ff, const double *fb, int n_fb,
                   double *state, int oldstyle);
uint32_t orc_crc32_mac(const uint8_t *buf, int64_t len);

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/ofdm_oracle.c -- CPU restatement of the gr-ofdm_tools OFDM PHY hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ofdm_oracle.h).  PARITY: partially pinned -- the reference has
 * no tests or golden outputs for this path; the only reference-supplied known answers are the
 * sync-word literals (checked in tests/test_oracle_golden.py).
 *
 * Each function cites the reference call site (relative to /root/reference) whose behaviour it
 * restates and the GNU Radio 3.7 block that call site instantiates ([UPSTREAM], SURVEY.md
 * Appendix A -- GNU Radio itself is an un-vendored dependency and is not in this image).
 */
#include "ofdm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cd;
#define TWO_PI 6.283185307179586476925286766559

/* ------------------------------------------------------------------------------------------ */
/* CRC-32: digital.crc32_bb (python/ofdm_radio_hier.py:121-122, python/ofdm_cr_tools.py:1275,1501)
 * [UPSTREAM crc32_bb_impl.cc: boost::crc_optimal<32,0x04C11DB7,0xFFFFFFFF,0xFFFFFFFF,true,true>] */
uint32_t orc_crc32(const uint8_t *buf, int64_t len)
{
    uint32_t crc = 0xFFFFFFFFu;
    for (int64_t i = 0; i < len; i++) {
        crc ^= buf[i];
        for (int b = 0; b < 8; b++)
            crc = (crc >> 1) ^ (0xEDB88320u & (0u - (crc & 1u)));
    }
    return crc ^ 0xFFFFFFFFu;
}

/* CRC-8 of the header: [UPSTREAM packet_header_default.cc: crc_optimal<8,0x07,0xFF,0x00,false,false>]
 * instantiated by digital.packet_header_ofdm (python/ofdm_txrx_modules.py:191-196). */
uint8_t orc_crc8(const uint8_t *buf, int64_t len)
{
    uint8_t crc = 0xFF;
    for (int64_t i = 0; i < len; i++) {
        crc ^= buf[i];
        for (int b = 0; b < 8; b++)
            crc = (uint8_t)((crc & 0x80) ? ((crc << 1) ^ 0x07) : (crc << 1));
    }
    return crc;
}

/* [UPSTREAM gnuradio/digital/lfsr.h] Fibonacci LFSR used by additive_scrambler_bb
 * (python/ofdm_txrx_modules.py:212-219) and by the header scramble mask. */
void orc_lfsr_bits(uint32_t mask, uint32_t seed, uint32_t reg_len, uint8_t *bits, int64_t n)
{
    uint32_t sr = seed;
    for (int64_t i = 0; i < n; i++) {
        bits[i] = (uint8_t)(sr & 1u);
        uint32_t nb = (uint32_t)__builtin_popcount(sr & mask) & 1u;
        sr = (sr >> 1) | (nb << reg_len);
    }
}

/* digital.additive_scrambler_bb(0x8a, seed, 7, 0, bits_per_byte=8, reset_tag_key=packet_len):
 * python/ofdm_txrx_modules.py:212-219 (TX), :408-415 (RX).  The LFSR is reset at every packet. */
void orc_scramble(uint8_t *buf, int64_t len, uint32_t seed)
{
    uint32_t sr = seed;
    for (int64_t i = 0; i < len; i++) {
        uint8_t sb = 0;
        for (int k = 0; k < 8; k++) {
            uint8_t o = (uint8_t)(sr & 1u);
            uint32_t nb = (uint32_t)__builtin_popcount(sr & 0x8au) & 1u;
            sr = (sr >> 1) | (nb << 7);
            sb ^= (uint8_t)(o << k);
        }
        buf[i] ^= sb;
    }
}

/* blocks.repack_bits_bb(k, l, key, align_output) in packet mode, LSB first:
 * python/ofdm_txrx_modules.py:220-224 (8 -> bps, align_output False) and :416 (bps -> 8, True).
 * Returns number of output items. */
int64_t orc_repack(const uint8_t *in, int64_t n_in, int k, int l, int align_output, uint8_t *out)
{
    int64_t n_out = n_in * k / l;
    if (!align_output && ((n_in * k) % l) != 0) n_out++;
    int in_idx = 0, out_idx = 0;
    int64_t n_read = 0, n_written = 0;
    while (n_written < n_out && n_read < n_in) {
        if (out_idx == 0) out[n_written] = 0;
        out[n_written] |= (uint8_t)(((in[n_read] >> in_idx) & 1) << out_idx);
        in_idx = (in_idx + 1) % k;
        out_idx = (out_idx + 1) % l;
        if (in_idx == 0) n_read++;
        if (out_idx == 0) n_written++;
    }
    if (out_idx) n_written++;
    return n_written;
}

/* ------------------------------------------------------------------------------------------ */
/* Constellations: _get_constellation(bps) python/ofdm_txrx_modules.py:106-118 and the string
 * mapping python/ofdm_radio_hier.py:61-68.  [UPSTREAM constellation.cc, digital/qam.py]
 * bps 6 (64-QAM by the same qam.py rule) is an extension (BASELINE.json config 4). */
int orc_constellation(int bps, float *pts)
{
    if (bps == 1) {
        pts[0] = -1.f; pts[1] = 0.f; pts[2] = 1.f; pts[3] = 0.f;
        return 2;
    }
    if (bps == 2) {
        const float a = 0.707107f; /* SQRT_TWO literal in constellation.cc */
        const float q[8] = { -a, -a, a, -a, -a, a, a, a };
        memcpy(pts, q, sizeof q);
        return 4;
    }
    if (bps == 3) {
        static const int mult[8] = { 1, 7, 15, 9, 3, 5, 13, 11 };
        const float angle = (float)(M_PI / 8.0);
        for (int i = 0; i < 8; i++) {
            pts[2 * i] = (float)cos(mult[i] * angle);
            pts[2 * i + 1] = (float)sin(mult[i] * angle);
        }
        return 8;
    }
    if (bps == 4 || bps == 6) {
        /* qam_constellation(m, differential=True, mod_code='none') ->
         * make_differential_constellation(m, gray_coded=False); not power-normalised */
        int m = 1 << bps;
        int side = (bps == 4) ? 2 : 4; /* points per quadrant side */
        double step = 1.0 / (side - 0.5);
        for (int i = 0; i < m; i++) {
            int y = i % side, x = (i / side) % side, quad = i / (side * side);
            double gx = (x + 0.5) * step, gy = (y + 0.5) * step, re, im;
            switch (quad) {
            case 0: re = gx; im = gy; break;
            case 1: re = -gy; im = gx; break;
            case 2: re = -gx; im = -gy; break;
            default: re = gy; im = -gx; break;
            }
            pts[2 * i] = (float)re;
            pts[2 * i + 1] = (float)im;
        }
        return m;
    }
    return -1;
}

/* constellation_rect sector -> value LUT (find_sector_values / get_closest_point) */
static void rect_lut(int bps, int *lut)
{
    float pts[128];
    int m = orc_constellation(bps, pts);
    int side = (bps == 4) ? 4 : 8;
    double w = 2.0 / (side - 1);
    for (int rs = 0; rs < side; rs++)
        for (int is = 0; is < side; is++) {
            double cr = (rs + 0.5 - side / 2.0) * w, ci = (is + 0.5 - side / 2.0) * w;
            int best = 0;
            double bd = 1e300;
            for (int i = 0; i < m; i++) {
                double dr = cr - pts[2 * i], di = ci - pts[2 * i + 1];
                double d = dr * dr + di * di;
                if (d < bd) { bd = d; best = i; }
            }
            lut[rs * side + is] = best;
        }
}

/* decision_maker of the constellation object handed to ofdm_equalizer_simpledfe and
 * constellation_decoder_cb (python/ofdm_txrx_modules.py:342-350,362,386-395,407). */
/* [UPSTREAM constellation.cc, GNU Radio >= 3.8] AMPLITUDE_NORMALIZATION: points (and the sector widths of
 * constellation_rect) scaled by n / sum |p|; 3.7 has no normalisation step.  Restated from memory. */
static double qam_scale(int bps, int norm)
{
    if (norm != 1 || (bps != 4 && bps != 6)) return 1.0;
    static double cache[2] = { 0.0, 0.0 };
    double cv = cache[bps == 6];
    if (cv != 0.0) return cv;
    float pts[128];
    int m = orc_constellation(bps, pts);
    double sum = 0;
    for (int i = 0; i < m; i++) sum += sqrt((double)pts[2 * i] * pts[2 * i] + (double)pts[2 * i + 1] * pts[2 * i + 1]);
    cache[bps == 6] = (double)m / sum;      /* idempotent: a race writes the same value */
    return (double)m / sum;
}

int orc_constellation_n(int bps, int norm, float *pts)
{
    int m = orc_constellation(bps, pts);
    double sc = qam_scale(bps, norm);
    if (sc != 1.0)
        for (int i = 0; i < 2 * m; i++) pts[i] = (float)((double)pts[i] * sc);
    return m;
}

int orc_decide_n(int bps, int norm, double re, double im)
{
    if (bps == 1) return re > 0;
    if (bps == 2) return 2 * (im > 0) + (re > 0);
    if (bps == 3) {
        int r = 0;
        if (fabs(re) <= fabs(im)) r = 4;
        if (re <= 0) r |= 1;
        if (im <= 0) r |= 2;
        return r;
    }
    static int lut16[16], lut64[64], init = 0;
    if (!init) {
#pragma omp critical(orc_lut)
        {
            if (!init) { rect_lut(4, lut16); rect_lut(6, lut64); init = 1; }
        }
    }
    int side = (bps == 4) ? 4 : 8;
    /* the sector LUT is invariant under the common scaling of points and widths */
    float w = (float)(2.0 / (side - 1) * qam_scale(bps, norm));
    int rsec = (int)(re / w + side / 2.0);
    int isec = (int)(im / w + side / 2.0);
    if (rsec < 0) rsec = 0;
    if (rsec >= side) rsec = side - 1;
    if (isec < 0) isec = 0;
    if (isec >= side) isec = side - 1;
    return (bps == 4 ? lut16 : lut64)[rsec * side + isec];
}

int orc_decide(int bps, double re, double im) { return orc_decide_n(bps, 0, re, im); }

int orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Header: digital.packet_header_ofdm(occupied_carriers, n_syms=1, ..., bits_per_header_sym,
 * bits_per_payload_sym, scramble_header) python/ofdm_txrx_modules.py:191-196,363-371;
 * python/ofdm_radio_hier.py:81.  [UPSTREAM packet_header_default.cc, packet_header_ofdm.cc] */
int orc_header_len(const orc_params *p) { return p->occ_sizes[0]; }

static void header_mask(const orc_params *p, uint8_t *mask)
{
    int hl = orc_header_len(p);
    memset(mask, 0, (size_t)hl);
    if (!p->scramble_header) return;
    int nb = hl * p->bps_header;
    uint8_t *bits = (uint8_t *)malloc((size_t)nb);
    orc_lfsr_bits(0x8a, 0x6f, 7, bits, nb);
    for (int i = 0; i < hl; i++)
        for (int k = 0; k < p->bps_header; k++)
            mask[i] ^= (uint8_t)(bits[i * p->bps_header + k] << k);
    free(bits);
}

void orc_header_format(const orc_params *p, int pkt_len, int pkt_num, uint8_t *out)
{
    int hl = orc_header_len(p), bpb = p->bps_header, msk = (1 << bpb) - 1;
    pkt_len &= 0x0FFF;
    pkt_num &= 0x0FFF;
    uint8_t cb[4] = { (uint8_t)(pkt_len & 0xFF), (uint8_t)(pkt_len >> 8),
                      (uint8_t)(pkt_num & 0xFF), (uint8_t)(pkt_num >> 8) };
    uint8_t crc = orc_crc8(cb, 4);
    memset(out, 0, (size_t)hl);
    int k = 0;
    for (int i = 0; i < 12 && k < hl; i += bpb, k++) out[k] = (uint8_t)((pkt_len >> i) & msk);
    for (int i = 0; i < 12 && k < hl; i += bpb, k++) out[k] = (uint8_t)((pkt_num >> i) & msk);
    for (int i = 0; i < 8 && k < hl; i += bpb, k++) out[k] = (uint8_t)((crc >> i) & msk);
    uint8_t *mask = (uint8_t *)malloc((size_t)hl);
    header_mask(p, mask);
    for (int i = 0; i < hl; i++) out[i] ^= mask[i];
    free(mask);
}

/* returns 1 if the header CRC-8 matches */
int orc_header_parse(const orc_params *p, const uint8_t *in, int *pkt_len_bytes, int *pkt_num,
                     int *pkt_syms, int *frame_syms)
{
    int hl = orc_header_len(p), bpb = p->bps_header, msk = (1 << bpb) - 1;
    uint8_t *d = (uint8_t *)malloc((size_t)hl), *mask = (uint8_t *)malloc((size_t)hl);
    header_mask(p, mask);
    for (int i = 0; i < hl; i++) d[i] = in[i] ^ mask[i];
    unsigned len = 0, num = 0;
    int k = 0, ok = 1;
    for (int i = 0; i < 12 && k < hl; i += bpb, k++) len |= ((unsigned)(d[k] & msk)) << i;
    if (k < hl) {
        for (int i = 0; i < 12 && k < hl; i += bpb, k++) num |= ((unsigned)(d[k] & msk)) << i;
        if (k < hl) {
            uint8_t cb[4] = { (uint8_t)(len & 0xFF), (uint8_t)(len >> 8),
                              (uint8_t)(num & 0xFF), (uint8_t)(num >> 8) };
            uint8_t crc = orc_crc8(cb, 4);
            for (int i = 0; i < 8 && k < hl; i += bpb, k++)
                if ((d[k] & msk) != ((crc >> i) & msk)) ok = 0;
        }
    }
    free(d);
    free(mask);
    *pkt_len_bytes = (int)len;
    *pkt_num = (int)num;
    int ps = (int)len * 8 / p->bps_payload;
    if (((int)len * 8) % p->bps_payload) ps++;
    *pkt_syms = ps;
    /* frame_len walks the carrier sets from set 0 (packet_header_ofdm::header_parser) */
    int fl = 0, acc = 0, s = 0;
    while (acc < ps) {
        fl++;
        acc += p->occ_sizes[s];
        s = (s + 1) % p->n_occ_sets;
    }
    *frame_syms = fl;
    return ok;
}

/* ------------------------------------------------------------------------------------------ */
/* fft.fft_vcc core: unnormalised DFT in either direction (python/ofdm_txrx_modules.py:241-246,
 * 340,385).  [UPSTREAM fft_vcc_fftw.cc]  The half-swap ("shift") is applied by the callers. */
typedef struct { int n; cd *tw; int *rev; } fft_plan;
static fft_plan g_plans[8];
static int g_nplans = 0;

static const fft_plan *plan_get(int n)
{
    const fft_plan *r = NULL;
#pragma omp critical(orc_fft_plan)
    {
        for (int i = 0; i < g_nplans; i++)
            if (g_plans[i].n == n) r = &g_plans[i];
        if (!r && g_nplans < 8) {
            fft_plan *pl = &g_plans[g_nplans];
            pl->n = n;
            pl->tw = (cd *)malloc(sizeof(cd) * (size_t)n);
            pl->rev = (int *)malloc(sizeof(int) * (size_t)n);
            int lg = 0;
            while ((1 << lg) < n) lg++;
            for (int i = 0; i < n; i++) {
                pl->tw[i] = cos(TWO_PI * i / n) - I * sin(TWO_PI * i / n);
                int rv = 0;
                for (int b = 0; b < lg; b++)
                    if (i & (1 << b)) rv |= 1 << (lg - 1 - b);
                pl->rev[i] = rv;
            }
            g_nplans++;
            r = pl;
        }
    }
    return r;
}

static void fft_cd(int n, int forward, const cd *in, cd *out)
{
    const fft_plan *pl = plan_get(n);
    for (int i = 0; i < n; i++) out[pl->rev[i]] = in[i];
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, stride = n / len;
        for (int s = 0; s < n; s += len)
            for (int j = 0; j < half; j++) {
                cd w = pl->tw[j * stride];
                if (!forward) w = conj(w);
                cd a = out[s + j], b = out[s + j + half] * w;
                out[s + j] = a + b;
                out[s + j + half] = a - b;
            }
    }
}

void orc_fft(int n, int forward, const double *in, double *out)
{
    fft_cd(n, forward, (const cd *)in, (cd *)out);
}

/* ------------------------------------------------------------------------------------------ */
static inline int shifted_bin(int c, int n)
{
    if (c < 0) c += n;
    return (c + n / 2) % n;
}

/* sync words in front of the header symbol: sync_word2 == NULL is sync_word2=() of the reference constructors
 * (python/ofdm_txrx_modules.py:174-183,311-321): one sync word, header_payload_demux(n_sync_words + 1 = 2, ...) */
static int n_sync_words(const orc_params *p) { return p->sync_word2 ? 2 : 1; }

static int check_params(const orc_params *p)
{
    int n = p->fft_len;
    if (n < 16 || (n & (n - 1))) return -1;
    if (p->cp_len < 0 || p->cp_len > n) return -1;
    if (p->n_occ_sets < 1) return -1;
    int b = p->bps_payload, h = p->bps_header;
    if (!(b == 1 || b == 2 || b == 3 || b == 4 || b == 6)) return -1;
    if (!(h == 1 || h == 2 || h == 3 || h == 4 || h == 6)) return -1;
    return 0;
}

/* number of payload OFDM symbols the carrier allocator emits for n_syms payload symbols that
 * follow a one-OFDM-symbol header (allocator walk continues with set 1 % n_sets) */
static int alloc_payload_ofdm_syms(const orc_params *p, int n_syms)
{
    int cnt = 0, acc = 0, s = 1 % p->n_occ_sets;
    while (acc < n_syms) {
        cnt++;
        acc += p->occ_sizes[s];
        s = (s + 1) % p->n_occ_sets;
    }
    return cnt;
}

int64_t orc_tx_frame_samples(const orc_params *p, int64_t payload_bytes)
{
    int64_t lp = payload_bytes + (p->crc_mode ? 4 : 0);
    int64_t ns = (lp * 8 + p->bps_payload - 1) / p->bps_payload;
    return (int64_t)(n_sync_words(p) + 1 + alloc_payload_ofdm_syms(p, (int)ns)) * (p->fft_len + p->cp_len)
           + (p->rolloff > 1 ? p->rolloff - 1 : 0);
}

/* TX chain: python/ofdm_txrx_modules.py:189-254 (ofdm_tx), python/ofdm_radio_hier.py:212-231
 * (crc32_bb, scrambler selectors), python/ofdm_tx_rx_hier.py:74,85-87 (x0.01).
 * [UPSTREAM packet_headergenerator_bb, chunks_to_symbols_bc, tagged_stream_mux,
 *  ofdm_carrier_allocator_cvc_impl.cc, fft_vcc (inverse, shift), ofdm_cyclic_prefixer (rolloff 0)] */
int orc_tx(const orc_params *p, const uint8_t *payload, const int64_t *pkt_off, int64_t n_pkts,
           int32_t first_pkt_num, float *samples_out, int64_t cap_samples, int64_t *sample_off)
{
    if (check_params(p)) return -1;
    const int n = p->fft_len, cp = p->cp_len, hl = orc_header_len(p);
    float hpts[128], ppts[128];
    orc_constellation_n(p->bps_header, p->qam_normalization, hpts);
    orc_constellation_n(p->bps_payload, p->qam_normalization, ppts);
    /* offsets of each set in the flat arrays */
    int *occ_base = (int *)malloc(sizeof(int) * (size_t)p->n_occ_sets);
    for (int s = 0, a = 0; s < p->n_occ_sets; s++) { occ_base[s] = a; a += p->occ_sizes[s]; }
    int *pil_base = (int *)malloc(sizeof(int) * (size_t)(p->n_pilot_sets + 1));
    for (int s = 0, a = 0; s < p->n_pilot_sets; s++) { pil_base[s] = a; a += p->pilot_sizes[s]; }
    int *pls_base = (int *)malloc(sizeof(int) * (size_t)(p->n_pilot_sym_sets + 1));
    for (int s = 0, a = 0; s < p->n_pilot_sym_sets; s++) { pls_base[s] = a; a += p->pilot_sym_sizes[s]; }

    /* [UPSTREAM ofdm_cyclic_prefixer_impl.cc ctor]: rolloff_len 1 is rectangular; the flanks (float vectors)
     * are rolloff_len-1 long because the first sample of the up / down flank is always zero / one */
    const int roll = p->rolloff > 1 ? p->rolloff : 0, nfl = roll ? roll - 1 : 0;
    if (roll > cp) return -1;
    float up_flank[4096], down_flank[4096];
    cd delay_line[4096];
    for (int i = 1; i < roll; i++) {
        up_flank[i - 1] = (float)(0.5 * (1 + cos(M_PI * i / roll - M_PI)));
        down_flank[i - 1] = (float)(0.5 * (1 + cos(M_PI * (roll - i) / roll - M_PI)));
        delay_line[i - 1] = 0.0;
    }

    int64_t pos = 0;
    int rc = 0;
    cd *fd = (cd *)malloc(sizeof(cd) * (size_t)n), *sw = (cd *)malloc(sizeof(cd) * (size_t)n),
       *td = (cd *)malloc(sizeof(cd) * (size_t)n);
    for (int64_t pk = 0; pk < n_pkts && !rc; pk++) {
        int64_t len = pkt_off[pk + 1] - pkt_off[pk];
        int64_t lp = len + (p->crc_mode ? 4 : 0);
        uint8_t *buf = (uint8_t *)malloc((size_t)lp + 8);
        memcpy(buf, payload + pkt_off[pk], (size_t)len);
        if (p->crc_mode) { /* crc32_bb(False): append CRC little-endian */
            uint32_t c = orc_crc32(buf, len);
            buf[len] = (uint8_t)c; buf[len + 1] = (uint8_t)(c >> 8);
            buf[len + 2] = (uint8_t)(c >> 16); buf[len + 3] = (uint8_t)(c >> 24);
        }
        uint8_t *hdr = (uint8_t *)malloc((size_t)hl);
        orc_header_format(p, (int)lp, (first_pkt_num + (int)pk) & 0xFFF, hdr);
        orc_scramble(buf, lp, (uint32_t)p->scramble_seed);
        int64_t ns_max = lp * 8 / p->bps_payload + 2;
        uint8_t *chunks = (uint8_t *)malloc((size_t)ns_max);
        int64_t ns = orc_repack(buf, lp, 8, p->bps_payload, 0, chunks);
        int n_pay = alloc_payload_ofdm_syms(p, (int)ns);
        const int nsw = n_sync_words(p);
        int n_ofdm = nsw + 1 + n_pay;
        sample_off[pk] = pos;
        if (pos + (int64_t)n_ofdm * (n + cp) + nfl > cap_samples) { rc = -2; }
        int64_t sym_idx = 0; /* index into concatenated header+payload symbols */
        int set = 0;
        for (int o = 0; o < n_ofdm && !rc; o++) {
            for (int k = 0; k < n; k++) fd[k] = 0;
            if (o == 0) {
                for (int k = 0; k < n; k++) fd[k] = p->sync_word1[2 * k] + I * p->sync_word1[2 * k + 1];
            } else if (o == 1 && nsw == 2) {
                for (int k = 0; k < n; k++) fd[k] = p->sync_word2[2 * k] + I * p->sync_word2[2 * k + 1];
            } else {
                int di = o - nsw; /* data OFDM symbol index, header = 0 */
                for (int k = 0; k < p->occ_sizes[set]; k++) {
                    int64_t tot = hl + ns;
                    if (sym_idx >= tot) break;
                    int bin = shifted_bin(p->occ_carriers[occ_base[set] + k], n);
                    if (sym_idx < hl) {
                        int v = hdr[sym_idx];
                        fd[bin] = hpts[2 * v] + I * hpts[2 * v + 1];
                    } else {
                        int v = chunks[sym_idx - hl];
                        fd[bin] = ppts[2 * v] + I * ppts[2 * v + 1];
                    }
                    sym_idx++;
                }
                set = (set + 1) % p->n_occ_sets;
                if (p->n_pilot_sets > 0) {
                    int ps = di % p->n_pilot_sets, pss = di % p->n_pilot_sym_sets;
                    for (int k = 0; k < p->pilot_sizes[ps]; k++) {
                        int bin = shifted_bin(p->pilot_carriers[pil_base[ps] + k], n);
                        fd[bin] = p->pilot_symbols[2 * (pls_base[pss] + k)]
                                  + I * p->pilot_symbols[2 * (pls_base[pss] + k) + 1];
                    }
                }
            }
            /* fft_vcc(inverse, shift=True): swap input halves, then backward DFT, no 1/N */
            for (int k = 0; k < n; k++) sw[k] = fd[(k + n / 2) % n];
            fft_cd(n, 0, sw, td);
            float *o_ = samples_out + 2 * (pos + (int64_t)o * (n + cp));
            for (int m = 0; m < n + cp; m++) {
                cd v = td[(m - cp + n) % n];
                if (m < nfl) {
                    /* [UPSTREAM ofdm_cyclic_prefixer_impl.cc work(), restated from memory -- parity unpinned]:
                     *   out[i] = out[i] * d_up_flank[i] + d_delay_line[i];  d_delay_line[i] = in[i] * d_down_flank[i]
                     * for i < rolloff_len-1, flanks 0.5*(1+cos(pi*i/rolloff_len - pi)) and its mirror, i = 1.. */
                    v = v * up_flank[m] + delay_line[m];
                    delay_line[m] = td[m] * down_flank[m];
                }
                v *= (double)p->tx_scale;
                float vr = (float)creal(v), vi = (float)cimag(v);
                if (p->tx_clip > 0.0f) {   /* analog.rail_ff(-c, c) on re and im (python/clipper.py:45-58) */
                    vr = vr < -p->tx_clip ? -p->tx_clip : (vr > p->tx_clip ? p->tx_clip : vr);
                    vi = vi < -p->tx_clip ? -p->tx_clip : (vi > p->tx_clip ? p->tx_clip : vi);
                }
                o_[2 * m] = vr;
                o_[2 * m + 1] = vi;
            }
        }
        /* tagged-stream mode: the delay line is flushed behind the last symbol and cleared */
        for (int m = 0; m < nfl; m++) {
            cd v = delay_line[m] * (double)p->tx_scale;
            float vr = (float)creal(v), vi = (float)cimag(v);
            if (p->tx_clip > 0.0f) {
                vr = vr < -p->tx_clip ? -p->tx_clip : (vr > p->tx_clip ? p->tx_clip : vr);
                vi = vi < -p->tx_clip ? -p->tx_clip : (vi > p->tx_clip ? p->tx_clip : vi);
            }
            samples_out[2 * (pos + (int64_t)n_ofdm * (n + cp) + m)] = vr;
            samples_out[2 * (pos + (int64_t)n_ofdm * (n + cp) + m) + 1] = vi;
            delay_line[m] = 0.0;
        }
        pos += (int64_t)n_ofdm * (n + cp) + nfl;
        free(buf); free(hdr); free(chunks);
    }
    sample_off[n_pkts] = pos;
    free(fd); free(sw); free(td); free(occ_base); free(pil_base); free(pls_base);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Schmidl & Cox: digital.ofdm_sync_sc_cfb(fft_len, cp_len) python/ofdm_txrx_modules.py:324,
 * python/ofdm_radio_hier.py:99.  [UPSTREAM ofdm_sync_sc_cfb_impl.cc: delay(N/2), conj, multiply,
 * fir_filter_ccf(N/2 taps of -1), mag^2, fir_filter_fff(N taps of 0.5), square, divide,
 * plateau_detector_fb(cp_len, 0.9), complex_to_arg, sample_and_hold]
 * Exact semantics: P[n] = -sum_{k<N/2} r[n-k] conj(r[n-k-N/2]); R[n] = 0.5 sum_{k<N} |r[n-k]|^2;
 * detect[n] = R^2 > 0 && |P|^2 >= thr * R^2 (== |P|^2/R^2 >= thr, NaN-safe), all in float64. */
static void plateau(const uint8_t *det, int64_t n, int max_len, int64_t *trig, int64_t max_trig,
                    int64_t *n_trig)
{
    /* [UPSTREAM plateau_detector_fb_impl.cc general_work], evaluated over the whole stream */
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; i++) {
        if (det[i]) {
            if (n - i < 2 * (int64_t)max_len) break;
            int64_t start = i;
            while (i < n && det[i]) i++;
            if (i - start > 1) {
                if (cnt < max_trig) trig[cnt] = start + (i - start) / 2;
                cnt++;
                i = (i + max_len < n - 1) ? i + max_len : n - 1;
            }
        }
    }
    *n_trig = cnt;
}

static void sc_pr_at(const float *r, int64_t n_, int n, int64_t idx, double *pre, double *pim, double *e)
{
    (void)n_;
    double sr = 0, si = 0, se = 0;
    int h = n / 2;
    for (int k = 0; k < h; k++) {
        int64_t a = idx - k, b = idx - k - h;
        if (b < 0) break;
        double ar = r[2 * a], ai = r[2 * a + 1], br = r[2 * b], bi = r[2 * b + 1];
        sr += ar * br + ai * bi;
        si += ai * br - ar * bi;
    }
    for (int k = 0; k < n; k++) {
        int64_t a = idx - k;
        if (a < 0) break;
        double ar = r[2 * a], ai = r[2 * a + 1];
        se += ar * ar + ai * ai;
    }
    *pre = -sr; *pim = -si; *e = se;
}

int orc_sync(const orc_params *p, const float *r, int64_t n_samp, uint8_t *detect_out,
             int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig)
{
    if (check_params(p)) return -1;
    const int n = p->fft_len;
    const double thr = (double)p->threshold;
    uint8_t *det = detect_out ? detect_out : (uint8_t *)malloc((size_t)(n_samp > 0 ? n_samp : 1));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_samp; i++) {
        double pr, pi, e;
        sc_pr_at(r, n_samp, n, i, &pr, &pi, &e);
        double R = 0.5 * e, R2 = R * R, pm2 = pr * pr + pi * pi;
        det[i] = (uint8_t)(R2 > 0.0 && pm2 >= thr * R2);
    }
    plateau(det, n_samp, p->cp_len, trig, max_trig, n_trig);
    int64_t nt = *n_trig < max_trig ? *n_trig : max_trig;
    for (int64_t j = 0; j < nt; j++) {
        double pr, pi, e;
        sc_pr_at(r, n_samp, n, trig[j], &pr, &pi, &e);
        cfo[j] = (float)atan2(pi, pr);
    }
    if (!detect_out) free(det);
    return 0;
}

/* float32 port evaluated the way GNU Radio does it (fresh FIR dot product per output item, float
 * accumulators, divide, compare) -- used only to time a CPU baseline that costs what the
 * reference's sync block costs (O(fft_len) MAC per input sample). */
int orc_sync_f32(const orc_params *p, const float *r, int64_t n_samp,
                 int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig)
{
    if (check_params(p)) return -1;
    const int n = p->fft_len, h = n / 2;
    uint8_t *det = (uint8_t *)malloc((size_t)(n_samp > 0 ? n_samp : 1));
    float *xr = (float *)malloc(sizeof(float) * (size_t)(n_samp + n));
    float *xi = (float *)malloc(sizeof(float) * (size_t)(n_samp + n));
    float *en = (float *)malloc(sizeof(float) * (size_t)(n_samp + n));
    float *pa = (float *)malloc(sizeof(float) * (size_t)(n_samp > 0 ? n_samp : 1));
    /* history of n zeros in front, as the FIR blocks see it */
    for (int i = 0; i < n; i++) xr[i] = xi[i] = en[i] = 0.f;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_samp; i++) {
        float ar = r[2 * i], ai = r[2 * i + 1], br = 0.f, bi = 0.f;
        if (i >= h) { br = r[2 * (i - h)]; bi = r[2 * (i - h) + 1]; }
        xr[n + i] = ar * br + ai * bi;
        xi[n + i] = ai * br - ar * bi;
        en[n + i] = ar * ar + ai * ai;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_samp; i++) {
        float sr = 0.f, si = 0.f, se = 0.f;
        const float *a = xr + n + i - h + 1, *b = xi + n + i - h + 1, *c = en + i + 1;
        /* VOLK evaluates these dot products with SIMD partial sums; allow the same here */
#pragma omp simd reduction(+ : sr, si)
        for (int k = 0; k < h; k++) { sr += -1.0f * a[k]; si += -1.0f * b[k]; }
#pragma omp simd reduction(+ : se)
        for (int k = 0; k < n; k++) se += 0.5f * c[k];
        float m = (sr * sr + si * si) / (se * se);
        det[i] = (uint8_t)(m >= p->threshold);
        pa[i] = atan2f(si, sr);
    }
    plateau(det, n_samp, p->cp_len, trig, max_trig, n_trig);
    int64_t nt = *n_trig < max_trig ? *n_trig : max_trig;
    for (int64_t j = 0; j < nt; j++) cfo[j] = pa[trig[j]];
    free(det); free(xr); free(xi); free(en); free(pa);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* RX chain after sync: python/ofdm_txrx_modules.py:325-426, python/ofdm_radio_hier.py:186-244. */
typedef struct {
    const orc_params *p;
    const float *r;
    int64_t n_samp;
    const int64_t *trig;    /* raw triggers (sorted) */
    const float *cfo;
    const double *base;     /* NCO phase just before each trigger */
    int64_t n_trig;
    uint8_t *occ_mask;      /* [N] shifted */
    uint8_t *pil_mask;      /* [n_pilot_sets][N] */
    cd *pil_val;            /* [n_pilot_sets][N] */
    int *occ_base;
    cd sw1[4096], sw2[4096];
    const float *z_base;    /* z_out of the call (locates the frame ordinal for the taps tap) */
} rx_ctx;

/* NCO phase at delayed-stream item i: analog.frequency_modulator_fc(-2.0/fft_len) driven by the
 * sample-and-held arg(P) (python/ofdm_txrx_modules.py:326,337).  [UPSTREAM
 * frequency_modulator_fc_impl.cc: phase += k*in[i] then sincos; sample_and_hold_ff updates on the
 * trigger item].  Closed form of that accumulation in float64. */
static double nco_phase(const rx_ctx *c, int64_t i)
{
    int64_t lo = 0, hi = c->n_trig; /* last trigger <= i */
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (c->trig[mid] <= i) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0.0;
    int64_t j = lo - 1;
    double k = -2.0 / c->p->fft_len;
    return c->base[j] + k * (double)c->cfo[j] * (double)(i - c->trig[j] + 1);
}

/* one OFDM symbol: blocks.delay(N+cp) * NCO -> header_payload_demux CP strip -> fft_vcc(forward,
 * shift) (python/ofdm_txrx_modules.py:325-340,385).  i0 = delayed-stream index of the first FFT
 * sample. out: N bins, shifted order. */
static void rx_symbol(const rx_ctx *c, int64_t i0, cd *out)
{
    const int n = c->p->fft_len, D = n + c->p->cp_len;
    cd *td = (cd *)malloc(sizeof(cd) * (size_t)n), *fd = (cd *)malloc(sizeof(cd) * (size_t)n);
    for (int m = 0; m < n; m++) {
        int64_t i = i0 + m, s = i - D;
        cd v = 0;
        if (s >= 0 && s < c->n_samp) v = c->r[2 * s] + I * c->r[2 * s + 1];
        double ph = nco_phase(c, i);
        td[m] = v * (cos(ph) + I * sin(ph));
    }
    fft_cd(n, 1, td, fd);
    for (int k = 0; k < n; k++) out[k] = fd[(k + n / 2) % n];
    free(td); free(fd);
}

/* digital.ofdm_chanest_vcvc(sw1, sw2, 1[, 0, max_carr_offset]) python/ofdm_txrx_modules.py:341,
 * python/ofdm_radio_hier.py:106.  [UPSTREAM ofdm_chanest_vcvc_impl.cc get_carr_offset/get_chan_taps] */
/* One sync symbol: [UPSTREAM ofdm_chanest_vcvc_impl.cc, restated from memory -- parity unpinned]
 *   ctor:  d_ref_sym = sync_symbol1; first/last active = first/last non-zero entry of it; if sync_symbol1[first+1] == 0
 *          { last++; d_interpolate = true; };  d_known_symbol_diffs[i] = |s1[i] - s1[i+2]|^2 for
 *          i = first, first+2, ... while i < last-2 && i < N-2
 *   get_carr_offset ("Correlate"): new_diffs[i] = |Y1[i] - Y1[i+2]|^2 (i < N-2); for g (even, in range):
 *          sum = sum_j known[j] * new_diffs[j+g] over j with known[j] != 0; first strict maximum wins
 *          (indices j+g outside [0, N) are read out of bounds upstream; taken as 0 here)
 *   get_chan_taps: taps[i-off] = Y1[i] / s1[i-off] where s1 != 0; with d_interpolate taps[i] = taps[i-1] for
 *          i = first+1, first+3, ... < last and taps[last] = taps[last-1] */
static int chanest_single(const rx_ctx *c, const cd *y1, cd *taps)
{
    const int n = c->p->fft_len;
    int first = 0, last = n - 1, interp = 0;
    for (int i = 0; i < n; i++) if (c->sw1[i] != 0) { first = i; break; }
    for (int i = n - 1; i >= 0; i--) if (c->sw1[i] != 0) { last = i; break; }
    if (first + 1 < n && c->sw1[first + 1] == 0) { if (last + 1 < n) last++; interp = 1; }
    int gneg = -first, gpos = n - last - 1;
    if (c->p->max_carr_offset != -1) {
        if (-c->p->max_carr_offset > gneg) gneg = -c->p->max_carr_offset;
        if (c->p->max_carr_offset < gpos) gpos = c->p->max_carr_offset;
    }
    if (gneg % 2) gneg++;
    if (gpos % 2) gpos--;
    double *nd = (double *)calloc((size_t)n, sizeof(double));
    for (int i = 0; i < n - 2; i++) { cd d = y1[i] - y1[i + 2]; nd[i] = creal(d) * creal(d) + cimag(d) * cimag(d); }
    double best = 0;
    int off = 0;
    for (int g = gneg; g <= gpos; g += 2) {
        double sum = 0;
        for (int i = first; i < last - 2 && i < n - 2; i += 2) {
            cd d = c->sw1[i] - c->sw1[i + 2];
            float kd = (float)(creal(d) * creal(d) + cimag(d) * cimag(d));
            if (kd != 0.f && i + g >= 0 && i + g < n) sum += (double)kd * nd[i + g];
        }
        if (sum > best) { best = sum; off = g; }
    }
    free(nd);
    for (int k = 0; k < n; k++) taps[k] = 0;
    int ls = 0, le = n;
    if (off > 0) ls = off; else if (off < 0) le = n + off;
    for (int i = ls; i < le; i++)
        if (c->sw1[i - off] != 0) taps[i - off] = y1[i] / c->sw1[i - off];
    if (interp) {
        for (int i = first + 1; i < last; i += 2) taps[i] = taps[i - 1];
        taps[last] = taps[last - 1];
    }
    return off;
}

static int chanest(const rx_ctx *c, const cd *y1, const cd *y2, cd *taps)
{
    const int n = c->p->fft_len;
    int first = 0, last = n - 1;
    for (int i = 0; i < n; i++) if (c->sw2[i] != 0) { first = i; break; }
    for (int i = n - 1; i >= 0; i--) if (c->sw2[i] != 0) { last = i; break; }
    int gneg = -first, gpos = n - last - 1;
    if (c->p->max_carr_offset != -1) {
        if (-c->p->max_carr_offset > gneg) gneg = -c->p->max_carr_offset;
        if (c->p->max_carr_offset < gpos) gpos = c->p->max_carr_offset;
    }
    if (gneg % 2) gneg++;
    if (gpos % 2) gpos--;
    double best = 0;
    int off = 0;
    for (int g = gneg; g <= gpos; g += 2) {
        cd acc = 0;
        for (int k = 0; k < n; k++) {
            if (c->sw1[k] == 0) continue;
            cd cv = c->sw2[k] / c->sw1[k];
            if (cv == 0) continue;
            acc += conj(y1[k + g]) * conj(cv) * y2[k + g];
        }
        double a = cabs(acc);
        if (a > best) { best = a; off = g; }
    }
    for (int k = 0; k < n; k++) taps[k] = 0;
    int ls = 0, le = n;
    if (off > 0) ls = off; else if (off < 0) le = n + off;
    for (int i = ls; i < le; i++)
        if (c->sw2[i - off] != 0) taps[i - off] = y2[i] / c->sw2[i - off];
    return off;
}

/* digital.ofdm_frame_equalizer_vcvc(simpledfe(...), cp_len, key, propagate, fixed_len) with
 * digital.ofdm_equalizer_simpledfe(fft_len, const, occupied, pilots, pilot_symbols,
 * symbols_skipped, alpha) python/ofdm_txrx_modules.py:343-357,387-400.
 * [UPSTREAM ofdm_frame_equalizer_vcvc_impl.cc work, ofdm_equalizer_simpledfe.cc equalize]
 * frame: n_sym x N (shifted), overwritten with the decided points; z: pre-decision y/H. */
static void frame_equalize(const rx_ctx *c, cd *frame, int n_sym, int off, cd *H, int bps,
                           int symbols_skipped, cd *z /* n_sym x N or NULL */)
{
    const orc_params *p = c->p;
    const int n = p->fft_len;
    const int64_t tot = (int64_t)n * n_sym;
    float pts[128];
    orc_constellation_n(bps, p->qam_normalization, pts);
    /* shift the whole frame buffer by the integer carrier offset (flat memcpy semantics) */
    cd *tmp = (cd *)malloc(sizeof(cd) * (size_t)tot);
    for (int64_t q = 0; q < tot; q++) {
        int64_t s = q + off;
        tmp[q] = (s >= 0 && s < tot) ? frame[s] : 0;
    }
    for (int i = 0; i < n_sym; i++) {
        float arg = (float)(-TWO_PI * off * p->cp_len / n * (i + 1));
        cd pc = cos((double)arg) + I * sin((double)arg);
        for (int k = 0; k < n; k++) frame[(int64_t)i * n + k] = tmp[(int64_t)i * n + k] * pc;
    }
    free(tmp);
    const double alpha = (double)p->alpha;
    int pset = p->n_pilot_sets ? symbols_skipped % p->n_pilot_sets : 0;
    for (int i = 0; i < n_sym; i++) {
        for (int k = 0; k < n; k++) {
            cd *y = &frame[(int64_t)i * n + k];
            if (z) z[(int64_t)i * n + k] = 0;
            if (!c->occ_mask[k]) continue;
            if (p->n_pilot_sets && c->pil_mask[(int64_t)pset * n + k]) {
                cd pv = c->pil_val[(int64_t)pset * n + k];
                H[k] = alpha * H[k] + (1 - alpha) * (*y) / pv;
                *y = pv;
            } else {
                cd ze = *y / H[k];
                if (z) z[(int64_t)i * n + k] = ze;
                int d = orc_decide_n(bps, p->qam_normalization, creal(ze), cimag(ze));
                cd se = pts[2 * d] + I * pts[2 * d + 1];
                H[k] = alpha * H[k] + (1 - alpha) * (*y) / se;
                *y = se;
            }
        }
        if (p->n_pilot_sets) pset = (pset + 1) % p->n_pilot_sets;
    }
    float arg = (float)(TWO_PI * off * p->cp_len / n * n_sym);
    cd pc = cos((double)arg) + I * sin((double)arg);
    for (int k = 0; k < n; k++) H[k] *= pc;
}

static void rx_ctx_init(rx_ctx *cp, const orc_params *p, const float *r, int64_t n_samp, const int64_t *trig,
                        const float *cfo, int64_t n_trig_v)
{
    const int n = p->fft_len;
    const int64_t *n_trig = &n_trig_v;
#define c (*cp)
    memset(&c, 0, sizeof c);
    c.p = p; c.r = r; c.n_samp = n_samp; c.trig = trig; c.cfo = cfo; c.n_trig = *n_trig;
    double *base = (double *)malloc(sizeof(double) * (size_t)(*n_trig + 1));
    {
        double k = -2.0 / n, ph = 0;
        for (int64_t j = 0; j < *n_trig; j++) {
            base[j] = ph;
            int64_t nxt = (j + 1 < *n_trig) ? trig[j + 1] : trig[j];
            ph += k * (double)cfo[j] * (double)(nxt - trig[j]);
        }
    }
    c.base = base;
    c.occ_mask = (uint8_t *)calloc((size_t)n, 1);
    c.occ_base = (int *)malloc(sizeof(int) * (size_t)p->n_occ_sets);
    for (int s = 0, a = 0; s < p->n_occ_sets; s++) {
        c.occ_base[s] = a;
        for (int k = 0; k < p->occ_sizes[s]; k++) c.occ_mask[shifted_bin(p->occ_carriers[a + k], n)] = 1;
        a += p->occ_sizes[s];
    }
    int nps = p->n_pilot_sets > 0 ? p->n_pilot_sets : 1;
    c.pil_mask = (uint8_t *)calloc((size_t)nps * n, 1);
    c.pil_val = (cd *)calloc((size_t)nps * n, sizeof(cd));
    for (int s = 0, a = 0; s < p->n_pilot_sets; s++) {
        for (int k = 0; k < p->pilot_sizes[s]; k++) {
            int bin = shifted_bin(p->pilot_carriers[a + k], n);
            c.pil_mask[(int64_t)s * n + bin] = 1;
            /* ofdm_equalizer_1d_pilots indexes pilot_symbols[set] with the carrier-set index */
            int off = 0;
            for (int q = 0; q < s && q < p->n_pilot_sym_sets; q++) off += p->pilot_sym_sizes[q];
            c.pil_val[(int64_t)s * n + bin] = p->pilot_symbols[2 * (off + k)] + I * p->pilot_symbols[2 * (off + k) + 1];
        }
        a += p->pilot_sizes[s];
    }
    for (int k = 0; k < n; k++) {
        c.sw1[k] = p->sync_word1[2 * k] + I * p->sync_word1[2 * k + 1];
        c.sw2[k] = p->sync_word2 ? p->sync_word2[2 * k] + I * p->sync_word2[2 * k + 1] : 0;
    }

#undef c
}

static void rx_ctx_free(rx_ctx *c)
{
    free((void *)c->base); free(c->occ_mask); free(c->occ_base); free(c->pil_mask); free(c->pil_val);
}

/* Everything the chain computes for ONE trigger: 3 header-side symbols, chanest, header equaliser + parser and,
 * when the header parses and the samples are there, the payload (python/ofdm_txrx_modules.py:340-426).
 * status: 1 = header CRC-8 failed, 2 = payload samples missing, 3 = decoded.  The caller has checked that the
 * 3 header-side symbols lie inside the buffer. */
typedef struct { int status, off, plen, pnum, psyms, fsyms, crc_ok; int64_t nbytes; } trig_dec;

static float *g_taps_out = NULL;      /* debug tap of orc_rx (set by orc_set_taps_out): N complex per emitted frame */

void orc_set_taps_out(float *taps) { g_taps_out = taps; }

static void decode_trigger(const rx_ctx *c, int64_t ti, int hl, uint8_t *dst, int64_t byte_stride,
                           float *zo, int64_t z_stride, trig_dec *o)
{
    const orc_params *p = c->p;
    const int n = p->fft_len, D = n + p->cp_len;
    const int64_t t = c->trig[ti];
    memset(o, 0, sizeof *o);
    const int nsw = n_sync_words(p), pre = nsw + 1;      /* symbols in front of the payload */
    if (t + pre * (int64_t)D > c->n_samp) { o->status = 2; return; }
    cd *y = (cd *)malloc(sizeof(cd) * (size_t)n * 3), *H = (cd *)malloc(sizeof(cd) * (size_t)n);
    cd *zh = (cd *)malloc(sizeof(cd) * (size_t)n);
    uint8_t *hbits = (uint8_t *)malloc((size_t)hl);
    /* y[0], y[1]: sync symbols (y[1] unused with one sync word); y[2]: header symbol */
    for (int j = 0; j < nsw; j++) rx_symbol(c, t + (int64_t)j * D + p->cp_len, y + (size_t)j * n);
    rx_symbol(c, t + (int64_t)nsw * D + p->cp_len, y + (size_t)2 * n);
    int off = (nsw == 2) ? chanest(c, y, y + n, H) : chanest_single(c, y, H);
    if (g_taps_out && zo) { /* (zo identifies the output slot: taps go to the same frame ordinal) */
        float *tp = g_taps_out + ((zo - c->z_base) / (2 * z_stride)) * 2 * (int64_t)n;
        for (int k = 0; k < n; k++) { tp[2 * k] = (float)creal(H[k]); tp[2 * k + 1] = (float)cimag(H[k]); }
    }
    frame_equalize(c, y + 2 * n, 1, off, H, p->bps_header, 0, zh);
    /* header serializer: set 0, all carriers; constellation_decoder_cb(header const) */
    for (int k = 0; k < hl; k++) {
        int bin = shifted_bin(p->occ_carriers[c->occ_base[0] + k], n);
        cd s = y[2 * n + bin];
        hbits[k] = (uint8_t)orc_decide_n(p->bps_header, p->qam_normalization, creal(s), cimag(s));
    }
    int plen, pnum, psyms, fsyms;
    int ok = orc_header_parse(p, hbits, &plen, &pnum, &psyms, &fsyms);
    o->off = off;
    if (!ok) { o->status = 1; goto done; }
    o->plen = plen; o->pnum = pnum; o->psyms = psyms; o->fsyms = fsyms;
    if (t + (int64_t)(pre + fsyms) * D > c->n_samp) { o->status = 2; goto done; }
    o->status = 3;
    o->crc_ok = 1;
    if (zo)
        for (int k = 0; k < hl; k++) {
            int bin = shifted_bin(p->occ_carriers[c->occ_base[0] + k], n);
            zo[2 * k] = (float)creal(zh[bin]); zo[2 * k + 1] = (float)cimag(zh[bin]);
        }
    if (fsyms > 0) {
        cd *pf = (cd *)malloc(sizeof(cd) * (size_t)n * (size_t)fsyms);
        cd *pz = (cd *)malloc(sizeof(cd) * (size_t)n * (size_t)fsyms);
        for (int i = 0; i < fsyms; i++)
            rx_symbol(c, t + (int64_t)(pre + i) * D + p->cp_len, pf + (size_t)i * n);
        frame_equalize(c, pf, fsyms, off, H, p->bps_payload, 1, pz);
        /* ofdm_serializer_vcc(fft_len, occupied, frame_key, packet_len_key, 1) +
         * constellation_decoder_cb + repack_bits_bb(bps, 8, key, True) + descrambler */
        uint8_t *syms = (uint8_t *)malloc((size_t)psyms + 1);
        int64_t cnt = 0;
        int set = 1 % p->n_occ_sets;
        for (int i = 0; i < fsyms && cnt < psyms; i++) {
            for (int k = 0; k < p->occ_sizes[set] && cnt < psyms; k++) {
                int bin = shifted_bin(p->occ_carriers[c->occ_base[set] + k], n);
                cd s = pf[(size_t)i * n + bin];
                syms[cnt] = (uint8_t)orc_decide_n(p->bps_payload, p->qam_normalization, creal(s), cimag(s));
                if (zo && hl + cnt < z_stride) {
                    zo[2 * (hl + cnt)] = (float)creal(pz[(size_t)i * n + bin]);
                    zo[2 * (hl + cnt) + 1] = (float)cimag(pz[(size_t)i * n + bin]);
                }
                cnt++;
            }
            set = (set + 1) % p->n_occ_sets;
        }
        uint8_t *pb = (uint8_t *)malloc((size_t)(cnt * p->bps_payload / 8 + 2));
        int64_t nbytes = orc_repack(syms, cnt, p->bps_payload, 8, 1, pb);
        orc_scramble(pb, nbytes, (uint32_t)p->scramble_seed);
        if (p->crc_mode) { /* crc32_bb(True): compare with trailing 4 bytes (LE) */
            if (nbytes < 4) o->crc_ok = 0;
            else {
                uint32_t cc = orc_crc32(pb, nbytes - 4);
                uint32_t got = (uint32_t)pb[nbytes - 4] | ((uint32_t)pb[nbytes - 3] << 8)
                               | ((uint32_t)pb[nbytes - 2] << 16) | ((uint32_t)pb[nbytes - 1] << 24);
                o->crc_ok = (cc == got);
            }
        }
        if (nbytes > byte_stride) nbytes = byte_stride;
        memcpy(dst, pb, (size_t)nbytes);
        o->nbytes = nbytes;
        free(pf); free(pz); free(syms); free(pb);
    }
done:
    free(y); free(H); free(zh); free(hbits);
}

static int rx_impl(const orc_params *p, const float *r, int64_t n_samp,
                   orc_frame *recs, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
                   float *z_out, int64_t z_stride, int64_t *n_frames,
                   int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig, int f32sync)
{
    if (check_params(p)) return -1;
    const int n = p->fft_len, D = n + p->cp_len, hl = orc_header_len(p);
    if (n > 4096) return -1;
    int rc = f32sync ? orc_sync_f32(p, r, n_samp, trig, cfo, max_trig, n_trig)
                     : orc_sync(p, r, n_samp, NULL, trig, cfo, max_trig, n_trig);
    if (rc) return rc;
    if (*n_trig > max_trig) return -3;

    rx_ctx c;
    rx_ctx_init(&c, p, r, n_samp, trig, cfo, *n_trig);
    c.z_base = z_out;

    /* Per-trigger decode (header, then payload) is a pure function of the trigger: the float32-sync baseline
     * decodes every trigger speculatively on all host threads and then walks the demux state machine over the
     * results -- what the CUDA path does; the exact path decodes on demand inside the walk. */
    trig_dec *spec = NULL;
    uint8_t *spec_bytes = NULL;
    if (f32sync && *n_trig > 0) {
        spec = (trig_dec *)calloc((size_t)*n_trig, sizeof(trig_dec));
        spec_bytes = (uint8_t *)malloc((size_t)*n_trig * (size_t)byte_stride);
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t ti = 0; ti < *n_trig; ti++)
            decode_trigger(&c, ti, hl, spec_bytes + ti * byte_stride, byte_stride, NULL, 0, &spec[ti]);
    }
    int64_t nf = 0, pos = 0, ti = 0;
    const int pre = n_sync_words(p) + 1;
    uint8_t *scratch = (uint8_t *)malloc((size_t)(byte_stride > 0 ? byte_stride : 1));
    /* digital.header_payload_demux(3, fft_len, cp_len, key, "", True) state machine
     * (python/ofdm_txrx_modules.py:328-334,382) [UPSTREAM header_payload_demux_impl.cc] */
    while (ti < *n_trig) {
        int64_t t = trig[ti];
        if (t < pos) { ti++; continue; }
        if (t + pre * (int64_t)D > n_samp) break; /* header never completes */
        trig_dec dd, *d = &dd;
        const int full = nf >= max_frames;                     /* no room: decode into the scratch slot, then fail */
        uint8_t *dst = full ? scratch : bytes_out + nf * byte_stride;
        float *zo = (z_out && !full) ? z_out + 2 * nf * z_stride : NULL;
        if (spec) {
            d = &spec[ti];
            if (d->status == 3) memcpy(dst, spec_bytes + ti * byte_stride, (size_t)d->nbytes);
        } else {
            decode_trigger(&c, ti, hl, dst, byte_stride, zo, z_stride, &dd);
        }
        if (d->status == 1) { pos = t + 1; ti++; continue; }   /* header CRC-8 failed */
        if (d->status == 2) break;                             /* payload never completes */
        if (full) { rc = -4; break; }
        orc_frame *f = &recs[nf];
        memset(f, 0, sizeof *f);
        f->trigger = t; f->cfo = cfo[ti]; f->carr_offset = d->off;
        f->pkt_len = (uint16_t)d->plen; f->pkt_num = (uint16_t)d->pnum; f->frame_syms = (uint32_t)d->fsyms;
        f->flags = ORC_F_HDR_OK | ORC_F_COMPLETE | ORC_F_ACCEPTED;
        f->slot = (uint32_t)nf;
        if (d->crc_ok) f->flags |= ORC_F_CRC_OK;
        nf++;
        if (d->fsyms > 0) pos = t + (int64_t)(pre + d->fsyms) * D - p->demux_holdoff;
        else pos = t + pre * (int64_t)D;
        ti++;
    }
    free(spec); free(spec_bytes); free(scratch);
    *n_frames = nf;
    rx_ctx_free(&c);
    return rc;
}

/* One record per raw plateau trigger, each decoded on its own (what ofdmx_set_emit_all returns): flags carry
 * ORC_F_HDR_SEEN (the 3 header-side symbols lie inside the buffer), ORC_F_HDR_OK, ORC_F_COMPLETE, ORC_F_CRC_OK;
 * slot = trigger ordinal.  The demux acceptance rule is NOT applied (ORC_F_ACCEPTED never set). */
int orc_rx_all(const orc_params *p, const float *r, int64_t n_samp, orc_frame *recs, int64_t max_recs,
               uint8_t *bytes_out, int64_t byte_stride, int64_t *n_recs)
{
    if (check_params(p)) return -1;
    int64_t max_trig = n_samp / (p->cp_len > 0 ? p->cp_len : 1) + 16, nt = 0;
    int64_t *trig = (int64_t *)malloc(sizeof(int64_t) * (size_t)max_trig);
    float *cfo = (float *)malloc(sizeof(float) * (size_t)max_trig);
    int rc = orc_sync(p, r, n_samp, NULL, trig, cfo, max_trig, &nt);
    if (rc == 0 && nt > max_recs) rc = -4;
    if (rc) { free(trig); free(cfo); return rc; }
    rx_ctx c;
    rx_ctx_init(&c, p, r, n_samp, trig, cfo, nt);
    const int hl = orc_header_len(p), D = p->fft_len + p->cp_len;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ti = 0; ti < nt; ti++) {
        orc_frame *f = &recs[ti];
        memset(f, 0, sizeof *f);
        f->trigger = trig[ti]; f->cfo = cfo[ti]; f->slot = (uint32_t)ti;
        if (trig[ti] + (n_sync_words(p) + 1) * (int64_t)D > n_samp) continue;
        trig_dec d;
        decode_trigger(&c, ti, hl, bytes_out + ti * byte_stride, byte_stride, NULL, 0, &d);
        f->flags = ORC_F_HDR_SEEN;
        f->carr_offset = d.off;
        if (d.status == 1) continue;
        f->flags |= ORC_F_HDR_OK;
        f->pkt_len = (uint16_t)d.plen; f->pkt_num = (uint16_t)d.pnum; f->frame_syms = (uint32_t)d.fsyms;
        if (d.status == 3) f->flags |= ORC_F_COMPLETE | (d.crc_ok ? ORC_F_CRC_OK : 0);
    }
    *n_recs = nt;
    rx_ctx_free(&c);
    free(trig); free(cfo);
    return 0;
}

int orc_rx(const orc_params *p, const float *r, int64_t n_samp,
           orc_frame *recs, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
           float *z_out, int64_t z_stride, int64_t *n_frames,
           int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig)
{
    return rx_impl(p, r, n_samp, recs, max_frames, bytes_out, byte_stride, z_out, z_stride, n_frames,
                   trig, cfo, max_trig, n_trig, 0);
}

/* CPU-baseline variant: the sync stage is the float32 FIR port (what GNU Radio's blocks cost);
 * everything after the trigger list is the same code as orc_rx. */
int orc_rx_baseline(const orc_params *p, const float *r, int64_t n_samp,
                    orc_frame *recs, int64_t max_frames, uint8_t *bytes_out, int64_t byte_stride,
                    int64_t *n_frames, int64_t *trig, float *cfo, int64_t max_trig, int64_t *n_trig)
{
    return rx_impl(p, r, n_samp, recs, max_frames, bytes_out, byte_stride, NULL, 0, n_frames,
                   trig, cfo, max_trig, n_trig, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* Next rows.
 *
 * analog.agc2_cc: python/ofdm_tx_rx_hier.py:75-76, python/ofdm_radio_hier.py:180-181 construct
 * agc2_cc(1e-1, 1e-2, 1.0, 1.0) and set_max_gain(65536).  [UPSTREAM gr-analog
 * include/gnuradio/analog/agc2.h, kernel::agc2_cc::scale(), restated from memory -- parity unpinned:]
 *     output = input * gain
 *     tmp    = -reference + sqrt(output.re^2 + output.im^2)
 *     rate   = (tmp > gain) ? attack : decay
 *     gain  -= tmp * rate
 *     if (gain < 0) gain = 10e-5;  if (max_gain > 0 && gain > max_gain) gain = max_gain
 * All in float32, one rounding per operation (this file is compiled with -ffp-contract=off). */
void orc_agc2(const float *in, float *out, int64_t n, float attack, float decay, float reference,
              float max_gain, float *gain)
{
    orc_agc2_v(in, out, n, attack, decay, reference, max_gain, gain, 0);
}

/* abs_rate != 0: the rule of agc2.h from GNU Radio 3.8 on, rate = (fabsf(tmp) > gain) ? attack : decay */
void orc_agc2_v(const float *in, float *out, int64_t n, float attack, float decay, float reference,
                float max_gain, float *gain, int abs_rate)
{
    float g = *gain;
    for (int64_t i = 0; i < n; i++) {
        const float re = in[2 * i] * g, im = in[2 * i + 1] * g;
        out[2 * i] = re;
        out[2 * i + 1] = im;
        const float rr = re * re, ii = im * im;
        const float tmp = -reference + sqrtf(rr + ii);
        const float rate = ((abs_rate ? fabsf(tmp) : tmp) > g) ? attack : decay;
        g -= tmp * rate;
        if (g < 0.0f) g = 10e-5f;
        if (max_gain > 0.0f && g > max_gain) g = max_gain;
    }
    *gain = g;
}

/* MAC-level CRC-32: [UPSTREAM gr-digital/lib/crc32.cc update_crc32: table-driven,
 * crc = (crc << 8) ^ table[((crc >> 24) ^ byte) & 0xFF], init 0xFFFFFFFF, result ^ 0xFFFFFFFF]. */
uint32_t orc_crc32_mac(const uint8_t *buf, int64_t len)
{
    uint32_t crc = 0xFFFFFFFFu;
    for (int64_t i = 0; i < len; i++) {
        crc ^= (uint32_t)buf[i] << 24;
        for (int k = 0; k < 8; k++) crc = (crc & 0x80000000u) ? (crc << 1) ^ 0x04C11DB7u : (crc << 1);
    }
    return crc ^ 0xFFFFFFFFu;
}

/* filter.iir_filter_ccd(fftaps, fbtaps, oldstyle=False) -- python/ofdm_radio_hier.py:83-84,93 (8th-order
 * out-of-band TX filter), python/sync_radio_hier.py:73.  [UPSTREAM gr-filter include/gnuradio/filter/iir_filter.h,
 * iir_filter<gr_complex, gr_complex, double, gr_complexd>::filter and set_taps(), restated from memory --
 * parity unpinned:]
 *     set_taps (oldstyle == false): d_fbtaps[i] = -fbtaps[i] for i >= 1 (fbtaps[0] is never used)
 *     acc = d_fftaps[0] * (gr_complexd)in
 *     for i = 1 .. n-1:  acc += d_fftaps[i] * (gr_complexd)prev_input[i]      (prev_input[1] = x[n-1], ...)
 *     for i = 1 .. m-1:  acc += d_fbtaps[i] * prev_output[i]                  (complex double history)
 *     out = (gr_complex)acc
 * real tap x complex double = component-wise products; one rounding per operation (-ffp-contract=off).
 * state[2*(n_ff-1) + 2*(n_fb-1)] doubles: x history (re,im pairs, newest first) then y history. */
void orc_iir_ccd(const float *in, float *out, int64_t n, const double *ff, int n_ff, const double *fb, int n_fb,
                 double *state)
{
    orc_iir_ccd_v(in, out, n, ff, n_ff, fb, n_fb, state, 0);
}

/* oldstyle != 0: iir_filter_ccd(..., oldstyle=True), set_taps() keeps the feedback taps as given (plus sign) */
void orc_iir_ccd_v(const float *in, float *out, int64_t n, const double *ff, int n_ff, const double *fb, int n_fb,
                   double *state, int oldstyle)
{
    double *xh = state, *yh = state + 2 * (n_ff - 1);
    const int nx = n_ff - 1, ny = n_fb > 0 ? n_fb - 1 : 0;
    for (int64_t k = 0; k < n; k++) {
        const double xr = (double)in[2 * k], xi = (double)in[2 * k + 1];
        double ar = ff[0] * xr, ai = ff[0] * xi;
        for (int i = 1; i < n_ff; i++) {
            ar += ff[i] * xh[2 * (i - 1)];
            ai += ff[i] * xh[2 * (i - 1) + 1];
        }
        for (int i = 1; i < n_fb; i++) {
            const double t = oldstyle ? fb[i] : -fb[i];
            ar += t * yh[2 * (i - 1)];
            ai += t * yh[2 * (i - 1) + 1];
        }
        for (int i = nx - 1; i > 0; i--) { xh[2 * i] = xh[2 * i - 2]; xh[2 * i + 1] = xh[2 * i - 1]; }
        for (int i = ny - 1; i > 0; i--) { yh[2 * i] = yh[2 * i - 2]; yh[2 * i + 1] = yh[2 * i - 1]; }
        if (nx > 0) { xh[0] = xr; xh[1] = xi; }
        if (ny > 0) { yh[0] = ar; yh[1] = ai; }
        out[2 * k] = (float)ar;
        out[2 * k + 1] = (float)ai;
    }
}

/*
 * lteo_sync.c -- CPU restatement of the cell-search primitives (TEST INFRASTRUCTURE, see lte_oracle.h): what srsUE
 * reaches through srslte_ue_cellsearch_scan (/root/reference/ue/src/phy/phch_recv.cc:146-177) -- PSS correlation,
 * CFO estimate, SSS detection -- plus the PSS/SSS generators for the synthetic subframes.  SPEC.md 13.
 * 3GPP TS 36.211 6.11.1 (PSS: Zadoff-Chu roots 25/29/34), 6.11.2 (SSS: interleaved scrambled m-sequences).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lte_oracle.h"

/* frequency-domain PSS d_u(n), n = 0..61, N_id_2 = 0, 1, 2 */
void lteo_pss_seq(int n_id_2, lteo_cd_t *d62) {
  static const int roots[3] = {25, 29, 34};
  int u = roots[n_id_2];
  for (int n = 0; n < 62; n++) {
    int m = (n < 31) ? n * (n + 1) : (n + 1) * (n + 2);
    double a = -M_PI * (double)u * (double)(m % 126) / 63.0;        /* exp(-j pi u m / 63) has period 126 in m */
    d62[n].re = cos(a); d62[n].im = sin(a);
  }
}

static void mseq(int taps, int8_t *s31) {     /* x(i+5) = sum of the tapped x(i+j) mod 2, x(0..4) = 0,0,0,0,1; s = 1 - 2x */
  int x[31] = {0, 0, 0, 0, 1};
  for (int i = 0; i < 26; i++) {
    int v = 0;
    for (int j = 0; j < 5; j++) if (taps & (1 << j)) v ^= x[i + j];
    x[i + 5] = v;
  }
  for (int i = 0; i < 31; i++) s31[i] = (int8_t)(1 - 2 * x[i]);
}

/* SSS d(0..61) of cell group N_id_1 (0..167) with N_id_2, in subframe 0 (sf5 = 0) or 5 (sf5 = 1) */
void lteo_sss_seq(int n_id_1, int n_id_2, int sf5, int8_t *d62) {
  int8_t s[31], c[31], z[31];
  mseq(0x05, s);       /* x(i+5) = x(i+2) + x(i) */
  mseq(0x09, c);       /* x(i+5) = x(i+3) + x(i) */
  mseq(0x17, z);       /* x(i+5) = x(i+4) + x(i+2) + x(i+1) + x(i) */
  int qp = n_id_1 / 30;
  int q = (n_id_1 + qp * (qp + 1) / 2) / 30;
  int mp = n_id_1 + q * (q + 1) / 2;
  int m0 = mp % 31, m1 = (m0 + mp / 31 + 1) % 31;
  for (int n = 0; n < 31; n++) {
    int s0 = s[(n + m0) % 31], s1 = s[(n + m1) % 31];
    int c0 = c[(n + n_id_2) % 31], c1 = c[(n + n_id_2 + 3) % 31];
    int z0 = z[(n + (m0 % 8)) % 31], z1 = z[(n + (m1 % 8)) % 31];
    if (!sf5) { d62[2 * n] = (int8_t)(s0 * c0); d62[2 * n + 1] = (int8_t)(s1 * c1 * z0); }
    else      { d62[2 * n] = (int8_t)(s1 * c0); d62[2 * n + 1] = (int8_t)(s0 * c1 * z1); }
  }
}

/* adds PSS (last symbol of slot 0) and SSS (the symbol before) to the grid of a subframe 0 or 5, all ports' grids get
 * the same signal on port 0 only (single-antenna transmission of the synchronisation signals) */
void lteo_sync_tx(const lteo_cell_t *cell, int sf_idx, lteo_cd_t *grid) {
  int nsc = 12 * cell->nof_prb, k0 = nsc / 2 - 31;
  lteo_cd_t p[62];
  int8_t s[62];
  lteo_pss_seq(cell->cell_id % 3, p);
  lteo_sss_seq(cell->cell_id / 3, cell->cell_id % 3, sf_idx == 5, s);
  const int lp = LTEO_NSLOT(cell->cp) - 1;          /* PSS: last symbol of slot 0 (6 or 5), SSS: the one before */
  for (int n = 0; n < 62; n++) {
    grid[lp * nsc + k0 + n] = p[n];
    grid[(lp - 1) * nsc + k0 + n].re = s[n]; grid[(lp - 1) * nsc + k0 + n].im = 0.0;
  }
}

/* time-domain PSS replica at 1.92 Msps: 128-point IDFT of the sequence on bins -31..-1, 1..31, scaled 1/sqrt(128)
 * (the OFDM modulator's convention), evaluated in double and rounded once */
void lteo_pss_time_n(int n_id_2, int nfft, lteo_cf_t *t) {
  lteo_cd_t d[62];
  lteo_pss_seq(n_id_2, d);
  for (int n = 0; n < nfft; n++) {
    double re = 0, im = 0;
    for (int i = 0; i < 62; i++) {
      int bin = (i < 31) ? i - 31 : i - 30;                       /* -31..-1, 1..31 */
      double a = 2.0 * M_PI * (double)(bin * n) / (double)nfft;
      re += d[i].re * cos(a) - d[i].im * sin(a);
      im += d[i].re * sin(a) + d[i].im * cos(a);
    }
    t[n].re = (float)(re / sqrt((double)nfft)); t[n].im = (float)(im / sqrt((double)nfft));
  }
}
void lteo_pss_time(int n_id_2, lteo_cf_t *t128) { lteo_pss_time_n(n_id_2, 128, t128); }

/* c(p) = sum_{n<128} x[p+n] conj(t[n]), n ascending, every product and sum rounded once; first and second half kept
 * apart for the CFO estimate */
static void pss_corr_at(const lteo_cf_t *x, const lteo_cf_t *t, int nfft, lteo_cf_t *c1, lteo_cf_t *c2) {
  for (int h = 0; h < 2; h++) {
    float re = 0.0f, im = 0.0f;
    for (int n = nfft / 2 * h; n < nfft / 2 * h + nfft / 2; n++) {
      float pr = x[n].re * t[n].re + x[n].im * t[n].im;
      float pi = x[n].im * t[n].re - x[n].re * t[n].im;
      re = re + pr; im = im + pi;
    }
    if (h == 0) { c1->re = re; c1->im = im; } else { c2->re = re; c2->im = im; }
  }
}

/*
 * PSS search over one buffer of n_samples at 1.92 Msps: for every position p < n_samples - 127 and every root the
 * correlation c = c1 + c2 and its power |c|^2; the peak is the largest power, lowest (root, position) index on ties
 * with the key root * n_pos + p.  Returns the peak power; outputs position (first sample of the PSS symbol body),
 * N_id_2 and the CFO in units of the subcarrier spacing: angle(conj(c1) c2) / pi.
 */
float lteo_pss_search(const lteo_cf_t *x, int n_samples, int *peak_pos, int *n_id_2, float *cfo, float *mean_power) {
  return lteo_pss_search_n(x, n_samples, 128, -1, 0, peak_pos, n_id_2, cfo, mean_power);
}

/* the same at any LTE sampling rate (replica of nfft samples), optionally for one root only and from first_pos on */
float lteo_pss_search_n(const lteo_cf_t *x, int n_samples, int nfft, int force_n_id_2, int first_pos, int *peak_pos, int *n_id_2,
                        float *cfo, float *mean_power) {
  lteo_cf_t *t[3];
  for (int u = 0; u < 3; u++) { t[u] = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * nfft); lteo_pss_time_n(u, nfft, t[u]); }
  int n_pos = n_samples - (nfft - 1);
  float best = -1.0f; int bu = 0, bp = 0;
  lteo_cf_t b1 = {0, 0}, b2 = {0, 0};
  double acc = 0.0;
  for (int u = 0; u < 3; u++) {
    if (force_n_id_2 >= 0 && u != force_n_id_2) continue;
    for (int p = first_pos; p < n_pos; p++) {
      lteo_cf_t c1, c2;
      pss_corr_at(x + p, t[u], nfft, &c1, &c2);
      float cr = c1.re + c2.re, ci = c1.im + c2.im;
      float pw = cr * cr + ci * ci;
      acc += pw;
      if (pw > best) { best = pw; bu = u; bp = p; b1 = c1; b2 = c2; }
    }
  }
  for (int u = 0; u < 3; u++) free(t[u]);
  if (peak_pos) *peak_pos = bp;
  if (n_id_2) *n_id_2 = bu;
  if (cfo) {
    float re = b1.re * b2.re + b1.im * b2.im, im = b1.re * b2.im - b1.im * b2.re;      /* conj(c1) c2 */
    *cfo = (float)(atan2((double)im, (double)re) / M_PI);
  }
  if (mean_power) *mean_power = (float)(acc / ((force_n_id_2 >= 0 ? 1.0 : 3.0) * (n_pos - first_pos)));
  return best;
}

/*
 * SSS detection given the PSS position (needs peak_pos >= 137): 128-point FFTs (SPEC.md 2) of the PSS symbol and of
 * the symbol before it (normal CP: 9 samples at this rate), channel H[k] = Y_pss[k] conj(d_u[k]) with d_u rounded to
 * float, Z[k] = Y_sss[k] conj(H[k]), complex correlation with all 168 x 2 sequences summed over k ascending, metric
 * |sum|^2 (a carrier offset turns Z by a common phase between the two symbols, which the magnitude ignores).  Returns
 * the best N_id_1; outputs whether the half-frame starts with subframe 5 and the metric.
 */
int lteo_sss_detect(const lteo_cf_t *x, int peak_pos, int n_id_2, int *sf5, float *corr_out) {
  return lteo_sss_detect_n(x, peak_pos, n_id_2, 128, sf5, corr_out);
}

/* the same at any LTE sampling rate: nfft-point transforms, the SSS symbol nfft + 9 nfft / 128 samples before the PSS */
static int sss_detect_gap(const lteo_cf_t *x, int peak_pos, int n_id_2, int nfft, int gap, int *sf5, float *corr_out);
int lteo_sss_detect_n(const lteo_cf_t *x, int peak_pos, int n_id_2, int nfft, int *sf5, float *corr_out) {
  return sss_detect_gap(x, peak_pos, n_id_2, nfft, nfft + 9 * nfft / 128, sf5, corr_out);
}

/* Cyclic-prefix detection (SPEC.md 15b.7; srsLTE's srslte_sync_detect_cp, reported at phch_recv.cc:189): the SSS symbol
 * lies nfft + 9 nfft / 128 samples before the PSS with the normal prefix and nfft + nfft / 4 with the extended one.
 * cp_mode 0 / 1 look at that one place; 2 tries both (each only if it lies inside the buffer) and keeps the larger
 * metric, the normal prefix on a tie.  Returns N_id_1, or -1 if no hypothesis could be tested. */
int lteo_sss_detect_cp(const lteo_cf_t *x, int peak_pos, int n_id_2, int nfft, int cp_mode, int *sf5, float *corr_out, int *cp_out) {
  int bn = -1, b5 = 0, bcp = 0;
  float best = 0.0f;
  for (int hyp = 0; hyp < 2; hyp++) {
    if (cp_mode != 2 && hyp != cp_mode) continue;
    int gap = nfft + (hyp ? nfft / 4 : 9 * nfft / 128), s5 = 0;
    float corr = 0.0f;
    if (peak_pos < gap) continue;
    int n1 = sss_detect_gap(x, peak_pos, n_id_2, nfft, gap, &s5, &corr);
    if (bn < 0 || corr > best) { bn = n1; b5 = s5; best = corr; bcp = hyp; }
  }
  if (sf5) *sf5 = b5;
  if (corr_out) *corr_out = best;
  if (cp_out) *cp_out = bcp;
  return bn;
}

static int sss_detect_gap(const lteo_cf_t *x, int peak_pos, int n_id_2, int nfft, int gap, int *sf5, float *corr_out) {
  lteo_cf_t *yp = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * 2 * nfft), *ys = yp + nfft;
  lteo_cd_t d[62];
  lteo_fft(x + peak_pos, yp, nfft);
  lteo_fft(x + peak_pos - gap, ys, nfft);
  lteo_pss_seq(n_id_2, d);
  float zr[62], zi[62];
  for (int i = 0; i < 62; i++) {
    int bin = (i < 31) ? nfft + (i - 31) : i - 30;
    float dr = (float)d[i].re, di = (float)d[i].im;
    float hr = yp[bin].re * dr + yp[bin].im * di, hi = yp[bin].im * dr - yp[bin].re * di;   /* Y conj(d) */
    zr[i] = ys[bin].re * hr + ys[bin].im * hi;                                              /* Ys conj(H) */
    zi[i] = ys[bin].im * hr - ys[bin].re * hi;
  }
  float best = 0.0f; int bn = 0, b5 = 0, first = 1;
  for (int s5 = 0; s5 < 2; s5++)
    for (int n1 = 0; n1 < 168; n1++) {
      int8_t sq[62];
      lteo_sss_seq(n1, n_id_2, s5, sq);
      float ar = 0.0f, ai = 0.0f;
      for (int i = 0; i < 62; i++) { ar = ar + (sq[i] > 0 ? zr[i] : -zr[i]); ai = ai + (sq[i] > 0 ? zi[i] : -zi[i]); }
      float acc = ar * ar + ai * ai;
      if (first || acc > best) { best = acc; bn = n1; b5 = s5; first = 0; }
    }
  free(yp);
  if (sf5) *sf5 = b5;
  if (corr_out) *corr_out = best;
  return bn;
}

/* ------------------------------------------------------------------------------------------------
 * Carrier-frequency-offset correction (SPEC.md 14).  Follows what srsLTE's subframe synchroniser does to the
 * samples before handing them to the worker (srslte_cfo_correct inside srslte_ue_sync_zerocopy, called at
 * /root/reference/ue/src/phy/phch_recv.cc:322; the estimate it uses is reported at :328).  srsLTE is an
 * un-vendored dependency: parity unpinned, arithmetic frozen here.
 * ---------------------------------------------------------------------------------------------- */
#include <pthread.h>
static lteo_cf_t g_cexp[LTEO_CFO_TABLE];
static pthread_once_t g_cexp_once = PTHREAD_ONCE_INIT;
static void cexp_init(void) {
  for (int j = 0; j < LTEO_CFO_TABLE; j++) {
    double a = 2.0 * M_PI * (double)j / (double)LTEO_CFO_TABLE;
    g_cexp[j].re = (float)cos(a);
    g_cexp[j].im = (float)sin(a);
  }
}

void lteo_cfo_table(lteo_cf_t *tab) {
  pthread_once(&g_cexp_once, cexp_init);
  memcpy(tab, g_cexp, sizeof(g_cexp));
}

/* phase step per sample in units of 2^-32 turns: removes an offset of `cfo` subcarrier spacings at nfft samples per symbol */
int32_t lteo_cfo_step(float cfo, int nfft) {
  double s = -(double)cfo / (double)nfft * 4294967296.0;
  return (int32_t)(uint32_t)(uint64_t)llrint(s);
}

void lteo_cfo_correct(const lteo_cf_t *in, lteo_cf_t *out, int n_samples, int32_t step) {
  pthread_once(&g_cexp_once, cexp_init);
  for (int n = 0; n < n_samples; n++) {
    uint32_t ph = (uint32_t)n * (uint32_t)step;
    lteo_cf_t w = g_cexp[ph >> (32 - LTEO_CFO_TABLE_LOG2)], x = in[n];
    float a = x.re * w.re, b = x.im * w.im, c = x.re * w.im, d = x.im * w.re;
    out[n].re = a - b;
    out[n].im = c + d;
  }
}

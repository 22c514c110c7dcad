/*
 * lteo_rx.c -- receive side of the CPU oracle: the restated hot path (TEST INFRASTRUCTURE, see
 * lte_oracle.h; arithmetic frozen in oracle/SPEC.md).
 *
 * Stage order follows what the reference drives through srsLTE:
 *   srslte_ue_dl_decode_fft_estimate   /root/reference/ue/src/phy/phch_worker.cc:254   -> lteo_ofdm_rx, lteo_chest
 *   srslte_ue_dl_cfg_grant             /root/reference/ue/src/phy/phch_worker.cc:337   -> lteo_cbsegm, lteo_pdsch_re_list
 *   srslte_pdsch_decode_rnti           /root/reference/ue/src/phy/phch_worker.cc:347-348 -> lteo_pdsch_decode
 *   soft-buffer ownership / reset      /root/reference/ue/src/mac/dl_harq.cc:191-259
 * All float arithmetic is single precision with one rounding per operation (compile with
 * -ffp-contract=off); the CUDA kernels perform the same operations in the same order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lte_oracle.h"

/* Small process-wide caches of derived tables (scrambling sequences, rate-matching orders, QPP permutations): a
 * receiver computes them once per configuration, not once per subframe (srsLTE precomputes its sequences in
 * srslte_pdsch_set_rnti and its interleavers at init).  Entries are immutable once published and never freed. */
#include <pthread.h>
#define CACHE_N 32
typedef struct { uint32_t a, b, c, d; int n; void *p; } cache_ent_t;
static pthread_mutex_t g_cache_mu = PTHREAD_MUTEX_INITIALIZER;
static void *cache_get(cache_ent_t *tab, int *cnt, uint32_t a, uint32_t b, uint32_t c, uint32_t d, int *n) {
  void *p = 0;
  pthread_mutex_lock(&g_cache_mu);
  for (int i = 0; i < *cnt; i++)
    if (tab[i].a == a && tab[i].b == b && tab[i].c == c && tab[i].d == d) { p = tab[i].p; if (n) *n = tab[i].n; break; }
  pthread_mutex_unlock(&g_cache_mu);
  return p;
}
/* returns the pointer to use: the published one (p is freed when another thread won the race or the table is full
 * and keep_if_full is 0; when full the caller owns p and *owned is set) */
static void *cache_put(cache_ent_t *tab, int *cnt, uint32_t a, uint32_t b, uint32_t c, uint32_t d, int n, void *p, int *owned) {
  *owned = 0;
  pthread_mutex_lock(&g_cache_mu);
  for (int i = 0; i < *cnt; i++)
    if (tab[i].a == a && tab[i].b == b && tab[i].c == c && tab[i].d == d) {
      void *q = tab[i].p;
      pthread_mutex_unlock(&g_cache_mu);
      free(p);
      return q;
    }
  if (*cnt < CACHE_N) {
    cache_ent_t e = {a, b, c, d, n, p};
    tab[(*cnt)++] = e;
  } else {
    *owned = 1;
  }
  pthread_mutex_unlock(&g_cache_mu);
  return p;
}

/* ------------------------------------------------------------------------------------------------
 * OFDM demodulation (SPEC.md 2)
 * ---------------------------------------------------------------------------------------------- */
static cache_ent_t g_tw_cache[CACHE_N], g_rev_cache[CACHE_N];
static int g_tw_cnt = 0, g_rev_cnt = 0;

/* twiddle table of lteo_fft_twiddles(n) and the bit-reversal permutation, computed once per size */
static const lteo_cf_t *cached_twiddles(int n, int *owned) {
  lteo_cf_t *tw = (lteo_cf_t *)cache_get(g_tw_cache, &g_tw_cnt, (uint32_t)n, 0, 0, 0, 0);
  *owned = 0;
  if (!tw) {
    tw = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * (n / 2 + 1));
    lteo_fft_twiddles(n, tw);
    tw = (lteo_cf_t *)cache_put(g_tw_cache, &g_tw_cnt, (uint32_t)n, 0, 0, 0, n, tw, owned);
  }
  return tw;
}
static const uint16_t *cached_bitrev(int n, int *owned) {
  uint16_t *rv = (uint16_t *)cache_get(g_rev_cache, &g_rev_cnt, (uint32_t)n, 0, 0, 0, 0);
  *owned = 0;
  if (!rv) {
    int bits = 0;
    while ((1 << bits) < n) bits++;
    rv = (uint16_t *)malloc(sizeof(uint16_t) * n);
    for (int i = 0; i < n; i++) {
      int r = 0;
      for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
      rv[i] = (uint16_t)r;
    }
    rv = (uint16_t *)cache_put(g_rev_cache, &g_rev_cnt, (uint32_t)n, 0, 0, 0, n, rv, owned);
  }
  return rv;
}
static void fft_pow2(const lteo_cf_t *in, lteo_cf_t *out, int n, int in_stride, const lteo_cf_t *tw, int ntw) {
  int rv_owned = 0;
  const uint16_t *rv = cached_bitrev(n, &rv_owned);
  for (int i = 0; i < n; i++) out[rv[i]] = in[(size_t)i * in_stride];
  if (rv_owned) free((void *)rv);
  for (int m = 2; m <= n; m <<= 1) {
    int half = m / 2, step = ntw / m;
    for (int g = 0; g < n; g += m)
      for (int j = 0; j < half; j++) {
        lteo_cf_t w = tw[j * step], a = out[g + j], b = out[g + j + half], t;
        t.re = w.re * b.re - w.im * b.im;
        t.im = w.re * b.im + w.im * b.re;
        out[g + j].re = a.re + t.re;        out[g + j].im = a.im + t.im;
        out[g + j + half].re = a.re - t.re; out[g + j + half].im = a.im - t.im;
      }
  }
}

/* forward DFT, n = 2^a or 3 * 2^a (1536): radix-2 DIT butterflies, one final radix-3 DIT stage */
void lteo_fft(const lteo_cf_t *in, lteo_cf_t *out, int n) {
  int tw_owned = 0;
  if (n % 3 != 0) {
    const lteo_cf_t *tw = cached_twiddles(n, &tw_owned);
    fft_pow2(in, out, n, 1, tw, n);
    if (tw_owned) free((void *)tw);
    return;
  }
  int m = n / 3;
  const lteo_cf_t *tw = cached_twiddles(m, &tw_owned);
  lteo_cf_t *f = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * n);
  for (int r = 0; r < 3; r++) fft_pow2(in + r, f + r * m, m, 3, tw, m);
  const float c3 = (float)(sqrt(3.0) / 2.0);
  for (int k = 0; k < m; k++) {
    /* full-circle twiddles of the radix-3 stage, evaluated in double and rounded once */
    double a1 = -2.0 * M_PI * (double)k / (double)n, a2 = -2.0 * M_PI * (double)(2 * k) / (double)n;
    lteo_cf_t w1 = {(float)cos(a1), (float)sin(a1)}, w2 = {(float)cos(a2), (float)sin(a2)};
    lteo_cf_t f0 = f[k], f1 = f[m + k], f2 = f[2 * m + k], t1, t2, s, d, mm;
    t1.re = w1.re * f1.re - w1.im * f1.im; t1.im = w1.re * f1.im + w1.im * f1.re;
    t2.re = w2.re * f2.re - w2.im * f2.im; t2.im = w2.re * f2.im + w2.im * f2.re;
    s.re = t1.re + t2.re; s.im = t1.im + t2.im;
    d.re = t1.re - t2.re; d.im = t1.im - t2.im;
    out[k].re = f0.re + s.re; out[k].im = f0.im + s.im;
    mm.re = f0.re - 0.5f * s.re; mm.im = f0.im - 0.5f * s.im;
    out[m + k].re = mm.re + c3 * d.im;     out[m + k].im = mm.im - c3 * d.re;
    out[2 * m + k].re = mm.re - c3 * d.im; out[2 * m + k].im = mm.im + c3 * d.re;
  }
  if (tw_owned) free((void *)tw);
  free(f);
}

void lteo_ofdm_rx(int nof_prb, const lteo_cf_t *iq, lteo_cf_t *sf_symbols) { lteo_ofdm_rx_cp(nof_prb, 0, iq, sf_symbols); }

/* cp = 1: 12 symbols behind 512-sample (at 2048) prefixes; rows 12 and 13 of the grid are left untouched */
void lteo_ofdm_rx_cp(int nof_prb, int cp, const lteo_cf_t *iq, lteo_cf_t *sf_symbols) {
  int n = lteo_symbol_sz(nof_prb), nsc = 12 * nof_prb, pos = 0;
  lteo_cf_t *x = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * n);
  const float sc = (float)(1.0 / sqrt((double)n));
  for (int l = 0; l < LTEO_NSYMB(cp); l++) {
    pos += lteo_cp_len_x(n, l, cp);
    lteo_fft(iq + pos, x, n);
    pos += n;
    for (int k = 0; k < nsc; k++) {
      int bin = (k < nsc / 2) ? (n - nsc / 2 + k) : (k - nsc / 2 + 1);
      sf_symbols[l * nsc + k].re = x[bin].re * sc;
      sf_symbols[l * nsc + k].im = x[bin].im * sc;
    }
  }
  free(x);
}

/* ------------------------------------------------------------------------------------------------
 * Channel estimation (SPEC.md 3)
 * ---------------------------------------------------------------------------------------------- */
/* the reduction order every measurement uses: 32 strided partial sums, then an xor-butterfly */
static float lane_reduce(const float *v, int n) {
  float p[32], q[32];
  for (int l = 0; l < 32; l++) p[l] = 0.0f;
  for (int i = 0; i < n; i++) p[i & 31] = p[i & 31] + v[i];
  for (int off = 16; off >= 1; off >>= 1) {
    for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ off];
    memcpy(p, q, sizeof(p));
  }
  return p[0];
}

static const int crs_syms_cp[2][4] = {{0, 4, 7, 11}, {0, 3, 6, 9}};

void lteo_chest(const lteo_cell_t *cell, int sf_idx, const lteo_cf_t *sf, lteo_cf_t *ce, float *meas) {
  const int nsc = 12 * cell->nof_prb, M = 2 * cell->nof_prb, np = cell->nof_ports;
  const float isq2 = (float)(1.0 / sqrt(2.0)), w = 0.1f, c = 0.8f;
  const int *crs_syms = crs_syms_cp[cell->cp ? 1 : 0], nslot = LTEO_NSLOT(cell->cp);
  float ftab[17];                                   /* ftab[t + 5] = (float)(t / 6), t = -5..11 */
  for (int t = -5; t <= 11; t++) ftab[t + 5] = (float)((double)t / 6.0);
  lteo_cf_t *ls = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * M);
  lteo_cf_t *sm = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * M);
  lteo_cf_t *hs = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * 4 * nsc);
  float *v_noise = (float *)malloc(sizeof(float) * np * 4 * M);      /* ports 2 / 3 fill half of theirs */
  float *v_rsrp = (float *)malloc(sizeof(float) * 4 * M);
  float *v_rssi = (float *)malloc(sizeof(float) * 4 * nsc);
  int n_noise = 0, n_rsrp = 0, n_rssi = 0;
  int8_t rs[220], is[220];
  int32_t kk[220];
  for (int p = 0; p < np; p++) {
    /* ports 2 / 3 carry pilots in symbol 1 of each slot only: two pilot symbols, one time segment (SPEC.md 15c) */
    const int psyms[2] = {1, nslot + 1};
    const int nsi = p < 2 ? 4 : 2;
    const int *syms = p < 2 ? crs_syms : psyms;
    for (int si = 0; si < nsi; si++) {
      int l = syms[si];
      lteo_crs_positions(cell, p, l, kk);
      lteo_crs_values(cell, sf_idx, l, rs, is);
      for (int m = 0; m < M; m++) {
        lteo_cf_t y = sf[l * nsc + kk[m]];
        float tre = (rs[m] > 0 ? y.re : -y.re) + (is[m] > 0 ? y.im : -y.im);
        float tim = (rs[m] > 0 ? y.im : -y.im) - (is[m] > 0 ? y.re : -y.re);
        ls[m].re = tre * isq2; ls[m].im = tim * isq2;
        if (p == 0) v_rsrp[n_rsrp++] = ls[m].re * ls[m].re + ls[m].im * ls[m].im;
      }
      sm[0] = ls[0]; sm[M - 1] = ls[M - 1];
      for (int m = 1; m < M - 1; m++) {
        sm[m].re = (w * ls[m - 1].re + c * ls[m].re) + w * ls[m + 1].re;
        sm[m].im = (w * ls[m - 1].im + c * ls[m].im) + w * ls[m + 1].im;
        float dr = ls[m].re - sm[m].re, di = ls[m].im - sm[m].im;
        v_noise[n_noise++] = dr * dr + di * di;
      }
      int off = kk[0];
      for (int k = 0; k < nsc; k++) {
        int m = (k - off >= 0) ? (k - off) / 6 : 0;
        if (m > M - 2) m = M - 2;
        int t = k - (6 * m + off);
        float f = ftab[t + 5];
        lteo_cf_t a = sm[m], b = sm[m + 1];
        hs[si * nsc + k].re = a.re + (b.re - a.re) * f;
        hs[si * nsc + k].im = a.im + (b.im - a.im) * f;
      }
    }
    /* time interpolation between CRS symbols 0,4,7,11; symbols 12,13 extrapolate from (7,11)
     * (extended cyclic prefix: 0,3,6,9; symbols 10,11 extrapolate from (6,9)) */
    for (int l = 0; l < 2 * nslot; l++) {
      int s0 = (p >= 2) ? 0 : (l < crs_syms[1]) ? 0 : (l < nslot) ? 1 : 2;
      int l0 = syms[s0], l1 = syms[s0 + 1];
      float f = (float)((double)(l - l0) / (double)(l1 - l0));
      for (int k = 0; k < nsc; k++) {
        lteo_cf_t a = hs[s0 * nsc + k], b = hs[(s0 + 1) * nsc + k];
        lteo_cf_t *o = &ce[(size_t)p * 14 * nsc + l * nsc + k];
        if (l == l0) { *o = a; continue; }
        if (l == l1) { *o = b; continue; }
        o->re = a.re + (b.re - a.re) * f;
        o->im = a.im + (b.im - a.im) * f;
      }
    }
  }
  for (int si = 0; si < 4; si++)
    for (int k = 0; k < nsc; k++) {
      lteo_cf_t y = sf[crs_syms[si] * nsc + k];
      v_rssi[n_rssi++] = y.re * y.re + y.im * y.im;
    }
  if (meas) {
    float noise = lane_reduce(v_noise, n_noise) / (float)n_noise / 0.06f;
    float rsrp = lane_reduce(v_rsrp, n_rsrp) / (float)n_rsrp;
    float rssi = lane_reduce(v_rssi, n_rssi) / (float)n_rssi;
    meas[0] = noise; meas[1] = rsrp; meas[2] = rssi;
    meas[3] = (float)cell->nof_prb * rsrp / rssi;
    meas[4] = rsrp / noise;
  }
  free(ls); free(sm); free(hs); free(v_noise); free(v_rsrp); free(v_rssi);
}

/* ------------------------------------------------------------------------------------------------
 * Equaliser (SPEC.md 4): MMSE / ZF for one port, Alamouti (SFBC) combiner for two
 * ---------------------------------------------------------------------------------------------- */
void lteo_equalize(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const lteo_cf_t *sf,
                   const lteo_cf_t *ce, float n0, lteo_cf_t *d, int *nof_re) {
  int nsc = 12 * cell->nof_prb;
  int32_t *re = (int32_t *)malloc(sizeof(int32_t) * 14 * nsc);
  int nre = lteo_pdsch_re_list(cell, cfg, re);
  if (cfg->tm == 2 && cell->nof_ports >= 2) {
    const float sq2 = (float)sqrt(2.0);
    for (int i = 0; i + 1 < nre; i += 2) {
      const lteo_cf_t *ce0 = ce + (size_t)LTEO_DIV_PA(cell->nof_ports, i / 2) * 14 * nsc, *ce1 = ce + (size_t)LTEO_DIV_PB(cell->nof_ports, i / 2) * 14 * nsc;
      lteo_cf_t r0 = sf[re[i]], r1 = sf[re[i + 1]], h0 = ce0[re[i]], h1 = ce1[re[i]];
      float den = ((h0.re * h0.re + h0.im * h0.im) + (h1.re * h1.re + h1.im * h1.im)) + n0;
      float a_re = h0.re * r0.re + h0.im * r0.im, a_im = h0.re * r0.im - h0.im * r0.re;   /* conj(h0) r0 */
      float b_re = h1.re * r1.re + h1.im * r1.im, b_im = h1.im * r1.re - h1.re * r1.im;   /* h1 conj(r1) */
      float c_re = h0.re * r1.re + h0.im * r1.im, c_im = h0.re * r1.im - h0.im * r1.re;   /* conj(h0) r1 */
      float e_re = h1.re * r0.re + h1.im * r0.im, e_im = h1.im * r0.re - h1.re * r0.im;   /* h1 conj(r0) */
      d[i].re = ((a_re + b_re) * sq2) / den;     d[i].im = ((a_im + b_im) * sq2) / den;
      d[i + 1].re = ((c_re - e_re) * sq2) / den; d[i + 1].im = ((c_im - e_im) * sq2) / den;
    }
  } else {
    for (int i = 0; i < nre; i++) {
      lteo_cf_t y = sf[re[i]], h = ce[re[i]];
      float den = (h.re * h.re + h.im * h.im) + n0;
      d[i].re = (y.re * h.re + y.im * h.im) / den;
      d[i].im = (y.im * h.re - y.re * h.im) / den;
    }
  }
  if (nof_re) *nof_re = nre;
  free(re);
}

/* ------------------------------------------------------------------------------------------------
 * Soft demapper -> int16 (SPEC.md 5), descrambler, rate de-matcher (SPEC.md 6)
 * ---------------------------------------------------------------------------------------------- */
static inline int16_t q16(float v) {
  if (!(v == v)) return 0;
  float t = truncf(v);
  if (t > 32767.0f) t = 32767.0f;
  if (t < -32767.0f) t = -32767.0f;
  return (int16_t)t;
}

void lteo_demod(const lteo_cf_t *d, int nof_re, int qm, int16_t *llr) {
  if (qm == 2) {
    const float s = (float)(100.0 * sqrt(2.0));
    for (int i = 0; i < nof_re; i++) {
      llr[2 * i] = q16(-(s * d[i].re));
      llr[2 * i + 1] = q16(-(s * d[i].im));
    }
  } else if (qm == 4) {
    const float s = 400.0f, c1 = (float)(2.0 * 400.0 / sqrt(10.0));
    for (int i = 0; i < nof_re; i++) {
      float tr = s * d[i].re, ti = s * d[i].im;
      llr[4 * i] = q16(-tr);
      llr[4 * i + 1] = q16(-ti);
      llr[4 * i + 2] = q16(fabsf(tr) - c1);
      llr[4 * i + 3] = q16(fabsf(ti) - c1);
    }
  } else {
    const float s = 700.0f, c1 = (float)(4.0 * 700.0 / sqrt(42.0)), c2 = (float)(2.0 * 700.0 / sqrt(42.0));
    for (int i = 0; i < nof_re; i++) {
      float tr = s * d[i].re, ti = s * d[i].im;
      float br = fabsf(tr) - c1, bi = fabsf(ti) - c1;
      llr[6 * i] = q16(-tr);
      llr[6 * i + 1] = q16(-ti);
      llr[6 * i + 2] = q16(br);
      llr[6 * i + 3] = q16(bi);
      llr[6 * i + 4] = q16(fabsf(br) - c2);
      llr[6 * i + 5] = q16(fabsf(bi) - c2);
    }
  }
}

static cache_ent_t g_gold_cache[CACHE_N];
static int g_gold_cnt = 0;

void lteo_descramble(int16_t *llr, int n, uint32_t c_init) {
  int owned = 0;
  uint8_t *c = (uint8_t *)cache_get(g_gold_cache, &g_gold_cnt, c_init, (uint32_t)n, 0, 0, 0);
  if (!c) {
    c = (uint8_t *)malloc(n > 0 ? n : 1);
    lteo_gold(c_init, n, c);
    c = (uint8_t *)cache_put(g_gold_cache, &g_gold_cnt, c_init, (uint32_t)n, 0, 0, n, c, &owned);
  }
  /* branch-free two's-complement negation where c[i] == 1 (the sequence is random: a branch mispredicts half the time) */
  for (int i = 0; i < n; i++) llr[i] = (int16_t)((llr[i] ^ -(int16_t)c[i]) + c[i]);
  if (owned) free(c);
}

static inline int16_t sat_add(int a, int b) {
  int s = a + b;
  if (s > LTEO_SB_MAX) s = LTEO_SB_MAX;
  if (s < -LTEO_SB_MAX) s = -LTEO_SB_MAX;
  return (int16_t)s;
}

/* accumulates E received LLRs into the soft buffer w (3K+12 triples), ascending e order; filler
 * positions of d0/d1 are forced to LTEO_FILLER_LLR */
/* ------------------------------------------------------------------------------------------------
 * PCFICH (SPEC.md 9): the CFI that srslte_ue_dl_decode_fft_estimate returns (phch_worker.cc:254)
 * ---------------------------------------------------------------------------------------------- */
int lteo_pcfich_decode(const lteo_cell_t *cell, int sf_idx, const lteo_cf_t *sf, const lteo_cf_t *ce, float n0,
                       int32_t *corr) {
  int nsc = 12 * cell->nof_prb;
  int32_t k[16];
  lteo_cf_t d[16];
  int16_t llr[32];
  lteo_pcfich_re(cell, k);
  if (cell->nof_ports >= 2) {
    const float sq2 = (float)sqrt(2.0);
    for (int i = 0; i < 16; i += 2) {          /* the Alamouti combiner of lteo_equalize */
      const lteo_cf_t *ce0 = ce + (size_t)LTEO_DIV_PA(cell->nof_ports, i / 2) * 14 * nsc, *ce1 = ce + (size_t)LTEO_DIV_PB(cell->nof_ports, i / 2) * 14 * nsc;
      lteo_cf_t r0 = sf[k[i]], r1 = sf[k[i + 1]], h0 = ce0[k[i]], h1 = ce1[k[i]];
      float den = ((h0.re * h0.re + h0.im * h0.im) + (h1.re * h1.re + h1.im * h1.im)) + n0;
      float a_re = h0.re * r0.re + h0.im * r0.im, a_im = h0.re * r0.im - h0.im * r0.re;
      float b_re = h1.re * r1.re + h1.im * r1.im, b_im = h1.im * r1.re - h1.re * r1.im;
      float c_re = h0.re * r1.re + h0.im * r1.im, c_im = h0.re * r1.im - h0.im * r1.re;
      float e_re = h1.re * r0.re + h1.im * r0.im, e_im = h1.im * r0.re - h1.re * r0.im;
      d[i].re = ((a_re + b_re) * sq2) / den;     d[i].im = ((a_im + b_im) * sq2) / den;
      d[i + 1].re = ((c_re - e_re) * sq2) / den; d[i + 1].im = ((c_im - e_im) * sq2) / den;
    }
  } else {
    for (int i = 0; i < 16; i++) {
      lteo_cf_t y = sf[k[i]], h = ce[k[i]];
      float den = (h.re * h.re + h.im * h.im) + n0;
      d[i].re = (y.re * h.re + y.im * h.im) / den;
      d[i].im = (y.im * h.re - y.re * h.im) / den;
    }
  }
  lteo_demod(d, 16, 2, llr);
  uint32_t c_init = ((uint32_t)(sf_idx + 1) * (uint32_t)(2 * cell->cell_id + 1) << 9) + (uint32_t)cell->cell_id;
  lteo_descramble(llr, 32, c_init);
  int best = 0;
  int32_t c3[3];
  for (int c = 0; c < 3; c++) {
    int32_t acc = 0;
    for (int n = 0; n < 32; n++) acc += ((n % 3) != c) ? (int32_t)llr[n] : -(int32_t)llr[n];   /* LLR > 0 <=> bit 1 */
    c3[c] = acc;
    if (acc > c3[best]) best = c;
  }
  if (corr) memcpy(corr, c3, sizeof(c3));
  return best + 1;
}

static cache_ent_t g_rm_cache[CACHE_N];
static int g_rm_cnt = 0;

void lteo_rm_rx(const int16_t *e, int E, int K, int F, int rv, int16_t *w) {
  int n = 0, owned = 0;
  int32_t *seq = (int32_t *)cache_get(g_rm_cache, &g_rm_cnt, (uint32_t)K, (uint32_t)F, (uint32_t)rv, 0, &n);
  if (!seq) {
    seq = (int32_t *)malloc(sizeof(int32_t) * 3 * (K + 4));
    n = lteo_rm_sequence(K, F, rv, seq);
    seq = (int32_t *)cache_put(g_rm_cache, &g_rm_cnt, (uint32_t)K, (uint32_t)F, (uint32_t)rv, 0, n, seq, &owned);
  }
  for (int i = 0, j = 0; i < E; i++) {
    w[seq[j]] = sat_add(w[seq[j]], e[i]);
    if (++j == n) j = 0;
  }
  for (int k = 0; k < F; k++) { w[3 * k] = LTEO_FILLER_LLR; w[3 * k + 1] = LTEO_FILLER_LLR; }
  if (owned) free(seq);
}

/* ------------------------------------------------------------------------------------------------
 * Turbo decoder (SPEC.md 7): int16 max-log-MAP, parallel windows with next-iteration initialisation
 * ---------------------------------------------------------------------------------------------- */
/* two's-complement wrap; every time the exact value does not fit is counted so that tests can check the
 * no-wrap argument of SPEC.md 7.6 (the counter is diagnostic only, not thread safe) */
static long g_wrap_events = 0;
static inline int16_t w16(int v) {
  if (v > 32767 || v < -32768) g_wrap_events++;
  return (int16_t)(uint16_t)v;
}
long lteo_wrap_events(int reset) { long v = g_wrap_events; if (reset) g_wrap_events = 0; return v; }
static inline int16_t clampi(int v, int lim) { return (int16_t)(v > lim ? lim : (v < -lim ? -lim : v)); }
static inline int16_t max16(int16_t a, int16_t b) { return a > b ? a : b; }

/* trellis of the RSC code: state s = (s1 s2 s3); for input u: next state and parity */
static int tr_next[8][2], tr_par[8][2], tr_init = 0;
static void trellis_init(void) {
  if (tr_init) return;
  for (int s = 0; s < 8; s++)
    for (int u = 0; u < 2; u++) {
      int s1 = (s >> 2) & 1, s2 = (s >> 1) & 1, s3 = s & 1;
      int fb = u ^ s2 ^ s3;
      tr_par[s][u] = fb ^ s1 ^ s3;
      tr_next[s][u] = (fb << 2) | (s1 << 1) | s2;
    }
  tr_init = 1;
}

static inline void gammas(int16_t x, int16_t y, int16_t g[2][2]) {
  g[0][0] = 0; g[0][1] = y; g[1][0] = x; g[1][1] = w16(x + y);
}

/* beta_k from beta_{k+1}; normalised when k % 4 == 0 */
static void beta_step(const int16_t *bn, int16_t *b, int16_t x, int16_t y, int k) {
  int16_t g[2][2];
  gammas(x, y, g);
  for (int s = 0; s < 8; s++) {
    int16_t v0 = bn[tr_next[s][0]], v1 = bn[tr_next[s][1]];
    int p0 = tr_par[s][0], p1 = tr_par[s][1];
    if (p0) v0 = w16(v0 + g[0][1]);            /* gamma(0,0) = 0: no addition */
    v1 = w16(v1 + g[1][p1]);
    b[s] = max16(v0, v1);
  }
  if (k % LTEO_TD_NORM == 0) {
    int16_t n = b[0];
    for (int s = 0; s < 8; s++) b[s] = w16(b[s] - n);
  }
}

/* alpha_{k+1} from alpha_k; normalised when (k+1) % 4 == 0 */
static void alpha_step(const int16_t *a, int16_t *an, int16_t x, int16_t y, int k) {
  int16_t g[2][2], t[8];
  gammas(x, y, g);
  for (int s = 0; s < 8; s++) t[s] = -32768;
  int seen[8] = {0};
  for (int s = 0; s < 8; s++)
    for (int u = 0; u < 2; u++) {
      int p = tr_par[s][u], n = tr_next[s][u];
      int16_t v = a[s];
      if (u || p) v = w16(v + g[u][p]);
      t[n] = seen[n] ? max16(t[n], v) : v;
      seen[n] = 1;
    }
  if ((k + 1) % LTEO_TD_NORM == 0) {
    int16_t n = t[0];
    for (int s = 0; s < 8; s++) t[s] = w16(t[s] - n);
  }
  memcpy(an, t, sizeof(t));
}

static int16_t ext_step(const int16_t *a, const int16_t *bn, int16_t y) {
  int16_t A[2][2];
  int seen[2][2] = {{0, 0}, {0, 0}};
  for (int s = 0; s < 8; s++)
    for (int u = 0; u < 2; u++) {
      int p = tr_par[s][u];
      int16_t v = w16(a[s] + bn[tr_next[s][u]]);
      A[u][p] = seen[u][p] ? max16(A[u][p], v) : v;
      seen[u][p] = 1;
    }
  int16_t l1 = max16(A[1][0], w16(A[1][1] + y));
  int16_t l0 = max16(A[0][0], w16(A[0][1] + y));
  return w16(l1 - l0);
}

typedef struct {
  int K, W, P;
  int16_t (*a_nii[2])[8];   /* [parity of iteration][window][state]  alpha at window start      */
  int16_t (*b_nii[2])[8];   /*                                        beta at window end         */
} map_state_t;

static void map_state_alloc(map_state_t *m, int K, int W) {
  m->K = K; m->W = W; m->P = K / W;
  for (int i = 0; i < 2; i++) {
    m->a_nii[i] = (int16_t(*)[8])calloc(m->P, 8 * sizeof(int16_t));
    m->b_nii[i] = (int16_t(*)[8])calloc(m->P, 8 * sizeof(int16_t));
  }
}
static void map_state_free(map_state_t *m) {
  for (int i = 0; i < 2; i++) { free(m->a_nii[i]); free(m->b_nii[i]); }
}

/* one max-log-MAP pass over all windows.  it = iteration index from 0: boundary metrics are read
 * from slot it&1 (written by iteration it-1, zeros for it == 0) and written to slot (it+1)&1. */
static void map_pass(map_state_t *m, int it, const int16_t *x, const int16_t *y, const int16_t *xt,
                     const int16_t *yt, int16_t *ext) {
  const int K = m->K, W = m->W, P = m->P, rd = it & 1, wr = (it + 1) & 1;
  int16_t (*beta)[8] = (int16_t(*)[8])malloc(sizeof(int16_t) * 8 * (W + 1));
  int16_t bt[4][8];
  /* trellis termination: beta_{K+3} = (0, -INF, ...), three regular backward steps over the tail */
  for (int s = 0; s < 8; s++) bt[3][s] = s ? -LTEO_TD_INF : 0;
  for (int t = 2; t >= 0; t--) beta_step(bt[t + 1], bt[t], xt[t], yt[t], K + t);
  for (int j = 0; j < P; j++) {
    int k0 = j * W;
    int16_t a[8], an[8];
    if (j == P - 1) memcpy(beta[W], bt[0], sizeof(a));
    else memcpy(beta[W], m->b_nii[rd][j], sizeof(a));
    for (int i = W - 1; i >= 0; i--) beta_step(beta[i + 1], beta[i], x[k0 + i], y[k0 + i], k0 + i);
    if (j > 0) memcpy(m->b_nii[wr][j - 1], beta[0], sizeof(a));
    if (j == 0) for (int s = 0; s < 8; s++) a[s] = s ? -LTEO_TD_INF : 0;
    else memcpy(a, m->a_nii[rd][j], sizeof(a));
    for (int i = 0; i < W; i++) {
      ext[k0 + i] = ext_step(a, beta[i + 1], y[k0 + i]);
      alpha_step(a, an, x[k0 + i], y[k0 + i], k0 + i);
      memcpy(a, an, sizeof(a));
    }
    if (j < P - 1) memcpy(m->a_nii[wr][j + 1], a, sizeof(a));
  }
  free(beta);
}

static cache_ent_t g_qpp_cache[CACHE_N];
static int g_qpp_cnt = 0;
static const uint16_t *cached_qpp(int K, int *owned) {
  uint16_t *pi = (uint16_t *)cache_get(g_qpp_cache, &g_qpp_cnt, (uint32_t)K, 0, 0, 0, 0);
  *owned = 0;
  if (!pi) {
    pi = (uint16_t *)malloc(sizeof(uint16_t) * K);
    lteo_qpp_perm(K, pi);
    pi = (uint16_t *)cache_put(g_qpp_cache, &g_qpp_cnt, (uint32_t)K, 0, 0, 0, K, pi, owned);
  }
  return pi;
}

int lteo_tdec_dbg(const int16_t *in, int K, int max_iter, int crc_type, uint8_t *bits, int *crc_ok,
                  int16_t *la_out, int window_override) {
  trellis_init();
  int W = window_override > 0 ? window_override : lteo_window_len(K);
  int pi_owned = 0;
  const uint16_t *pi = cached_qpp(K, &pi_owned);
  int16_t *sys = (int16_t *)malloc(sizeof(int16_t) * K * 8);
  int16_t *p1 = sys + K, *p2 = p1 + K, *la = p2 + K, *x = la + K, *ext = x + K, *A = ext + K;
  int16_t xt1[3], yt1[3], xt2[3], yt2[3];
  for (int k = 0; k < K; k++) {
    sys[k] = clampi(in[3 * k], LTEO_TD_C);
    p1[k] = clampi(in[3 * k + 1], LTEO_TD_C);
    p2[k] = clampi(in[3 * k + 2], LTEO_TD_C);
    la[k] = 0;
  }
  for (int t = 0; t < 3; t++) {
    xt1[t] = clampi(in[3 * K + 2 * t], LTEO_TD_C);     yt1[t] = clampi(in[3 * K + 2 * t + 1], LTEO_TD_C);
    xt2[t] = clampi(in[3 * K + 6 + 2 * t], LTEO_TD_C); yt2[t] = clampi(in[3 * K + 6 + 2 * t + 1], LTEO_TD_C);
  }
  map_state_t m1, m2;
  map_state_alloc(&m1, K, W);
  map_state_alloc(&m2, K, W);
  int it, ok = 0;
  for (it = 0; it < max_iter; it++) {
    for (int k = 0; k < K; k++) x[k] = w16(sys[k] + la[k]);
    map_pass(&m1, it, x, p1, xt1, yt1, ext);
    for (int k = 0; k < K; k++) A[k] = w16(sys[k] + clampi(ext[k], LTEO_TD_E));
    for (int k = 0; k < K; k++) x[k] = A[pi[k]];
    map_pass(&m2, it, x, p2, xt2, yt2, ext);
    for (int k = 0; k < K; k++) {
      la[pi[k]] = clampi(ext[k], LTEO_TD_E);
      bits[pi[k]] = (w16(x[k] + ext[k]) > 0) ? 1 : 0;
    }
    if (crc_type) {
      ok = lteo_crc_bits(bits, K, crc_type == 1 ? LTEO_CRC24A : LTEO_CRC24B, 24) == 0;
      if (ok) { it++; break; }
    }
  }
  if (crc_ok) *crc_ok = ok;
  if (la_out) memcpy(la_out, la, sizeof(int16_t) * K);
  map_state_free(&m1); map_state_free(&m2);
  if (pi_owned) free((void *)pi);
  free(sys);
  return it;
}

#ifdef LTEO_SIMD
#include "lteo_simd.inc"
int lteo_simd_build(void) { return 1; }
#else
int lteo_simd_build(void) { return 0; }
#endif

int lteo_tdec(const int16_t *in, int K, int max_iter, int crc_type, uint8_t *bits, int *crc_ok) {
#ifdef LTEO_SIMD
  return lteo_tdec_avx2(in, K, max_iter, crc_type, bits, crc_ok);
#endif
  return lteo_tdec_dbg(in, K, max_iter, crc_type, bits, crc_ok, 0, 0);
}

/* ------------------------------------------------------------------------------------------------
 * PDSCH decode (what srslte_pdsch_decode_rnti does) and the whole-chain wrapper
 * ---------------------------------------------------------------------------------------------- */
#define SB_STRIDE (3 * LTEO_MAX_K + 12)

int lteo_pdsch_decode(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const lteo_cf_t *sf,
                      const lteo_cf_t *ce, float noise_est, int max_iter, int16_t *softbuf, uint8_t *payload,
                      lteo_cf_t *d_out, int16_t *e_out, int *cb_iters, int *cb_crc) {
  lteo_cbsegm_t s;
  if (lteo_cbsegm(cfg->tbs, &s)) return -2;
  int nsc = 12 * cell->nof_prb, nre = 0, nl = (cfg->tm == 2) ? 2 : 1;
  lteo_cf_t *d = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * 14 * nsc);
  lteo_equalize(cell, cfg, sf, ce, noise_est, d, &nre);
  int G = nre * cfg->qm;
  int16_t *e = (int16_t *)malloc(sizeof(int16_t) * (G + 8));
  lteo_demod(d, nre, cfg->qm, e);
  uint32_t c_init = ((uint32_t)cfg->rnti << 14) | ((uint32_t)cfg->sf_idx << 9) | (uint32_t)cell->cell_id;
  lteo_descramble(e, G, c_init);
  if (d_out) memcpy(d_out, d, sizeof(lteo_cf_t) * nre);
  if (e_out) memcpy(e_out, e, sizeof(int16_t) * G);
  uint8_t *tb = (uint8_t *)malloc(s.B + 64), *cbits = (uint8_t *)malloc(LTEO_MAX_K);
  int rp = 0, wp = 0, all_ok = 1;
  for (int r = 0; r < s.C; r++) {
    int K = lteo_cb_len(&s, r), F = (r == 0) ? s.F : 0, L = (s.C > 1) ? 24 : 0;
    int E = lteo_cb_E(&s, G, cfg->qm, nl, r);
    int16_t *w = softbuf + (size_t)r * SB_STRIDE;
    lteo_rm_rx(e + rp, E, K, F, cfg->rv, w);
    rp += E;
    int ok = 0;
    int it = lteo_tdec(w, K, max_iter, s.C > 1 ? 2 : 1, cbits, &ok);
    if (cb_iters) cb_iters[r] = it;
    if (cb_crc) cb_crc[r] = ok;
    if (!ok) all_ok = 0;
    for (int i = F; i < K - L; i++) tb[wp++] = cbits[i];
  }
  /* wp == B = tbs + 24 */
  int tb_ok = all_ok && (lteo_crc_bits(tb, s.B, LTEO_CRC24A, 24) == 0);
  if (s.C == 1) tb_ok = all_ok;
  memset(payload, 0, (cfg->tbs + 7) / 8);
  for (int i = 0; i < cfg->tbs; i++) payload[i >> 3] |= (uint8_t)(tb[i] << (7 - (i & 7)));
  free(d); free(e); free(tb); free(cbits);
  return tb_ok ? 0 : -1;
}

int lteo_ue_dl_decode(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const lteo_cf_t *iq, float noise_est,
                      int noise_mode, int max_iter, int16_t *softbuf, uint8_t *payload, float *meas,
                      int *avg_iters) {
  int nsc = 12 * cell->nof_prb;
  lteo_cf_t *sf = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * 14 * nsc);
  lteo_cf_t *ce = (lteo_cf_t *)malloc(sizeof(lteo_cf_t) * 14 * nsc * cell->nof_ports);
  float mm[5];
  int iters[32];
  lteo_ofdm_rx_cp(cell->nof_prb, cell->cp, iq, sf);
  lteo_chest(cell, cfg->sf_idx, sf, ce, mm);
  if (meas) memcpy(meas, mm, sizeof(mm));
  int rc = lteo_pdsch_decode(cell, cfg, sf, ce, noise_mode ? mm[0] : noise_est, max_iter, softbuf, payload, 0, 0,
                             iters, 0);
  if (avg_iters) {
    lteo_cbsegm_t s;
    lteo_cbsegm(cfg->tbs, &s);
    int sum = 0;
    for (int r = 0; r < s.C; r++) sum += iters[r];
    *avg_iters = sum / s.C;
  }
  free(sf); free(ce);
  return rc;
}

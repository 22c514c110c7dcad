/*
 * lteo_tx.c -- transmit side of the CPU oracle: the synthetic-subframe generator used by tests and
 * bench.py (TEST INFRASTRUCTURE, see lte_oracle.h).  3GPP TS 36.212 5.1.1-5.1.4 / 5.3.2 and
 * TS 36.211 6.3, 6.10.1, 6.12.  Not part of the reference's receive path; it exists because the
 * reference has no IQ fixture (ue/test/phy/ue_itf_test_sib1.cc needs a live cell).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lte_oracle.h"

static const uint8_t perm_cols[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                      1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

/* constituent RSC encoder step: g0 = 1 + D^2 + D^3 (feedback), g1 = 1 + D + D^3 */
static inline int rsc_step(int *s, int u, int terminate, int *x_out) {
  int s1 = (*s >> 2) & 1, s2 = (*s >> 1) & 1, s3 = *s & 1;
  if (terminate) u = s2 ^ s3;
  int fb = u ^ s2 ^ s3;
  int z = fb ^ s1 ^ s3;
  *s = (fb << 2) | (s1 << 1) | s2;
  if (x_out) *x_out = u;
  return z;
}

/* 36.212 5.1.3.2: d is written as interleaved triples d[3k+i] = d^(i)_k, k = 0..K+3 */
void lteo_turbo_encode(const uint8_t *c, int K, uint8_t *d) {
  uint16_t *pi = (uint16_t *)malloc(sizeof(uint16_t) * K);
  lteo_qpp_perm(K, pi);
  int s1 = 0, s2 = 0;
  for (int k = 0; k < K; k++) {
    d[3 * k] = c[k] & 1;
    d[3 * k + 1] = (uint8_t)rsc_step(&s1, c[k] & 1, 0, 0);
    d[3 * k + 2] = (uint8_t)rsc_step(&s2, c[pi[k]] & 1, 0, 0);
  }
  int x[3], z[3], xp[3], zp[3];
  for (int t = 0; t < 3; t++) z[t] = rsc_step(&s1, 0, 1, &x[t]);
  for (int t = 0; t < 3; t++) zp[t] = rsc_step(&s2, 0, 1, &xp[t]);
  /* d0: x_K, z_K+1, x'_K, z'_K+1 ; d1: z_K, x_K+2, z'_K, x'_K+2 ; d2: x_K+1, z_K+2, x'_K+1, z'_K+2 */
  uint8_t *t = d + 3 * K;
  t[0] = x[0];  t[1] = z[0];  t[2] = x[1];
  t[3] = z[1];  t[4] = x[2];  t[5] = z[2];
  t[6] = xp[0]; t[7] = zp[0]; t[8] = xp[1];
  t[9] = zp[1]; t[10] = xp[2]; t[11] = zp[2];
  free(pi);
}

/* 36.212 5.1.4.1: the order in which the circular buffer is read for redundancy version rv, as
 * indices into the triples array d[3k+i] (k < K+4), skipping <NULL> (dummy and filler) positions.
 * Shared by the TX rate matcher and the RX de-matcher; returns the number of entries N_nd. */
int lteo_rm_sequence(int K, int F, int rv, int32_t *seq) {
  int D = K + 4, R = (D + 31) / 32, Kpi = 32 * R, ND = Kpi - D, Kw = 3 * Kpi;
  int k0 = R * (2 * ((Kw + 8 * R - 1) / (8 * R)) * rv + 2);
  int n = 0;
  for (int jj = 0; jj < Kw; jj++) {
    int j = (k0 + jj) % Kw, stream, kk, y;
    if (j < Kpi) { stream = 0; kk = j; }
    else { stream = 1 + ((j - Kpi) & 1); kk = (j - Kpi) >> 1; }
    if (stream < 2) y = (kk % R) * 32 + perm_cols[kk / R];
    else y = (perm_cols[kk / R] + 32 * (kk % R) + 1) % Kpi;
    int di = y - ND;
    if (di < 0) continue;                       /* dummy */
    if (stream < 2 && di < F) continue;         /* filler <NULL> in d0, d1 */
    seq[n++] = 3 * di + stream;
  }
  return n;
}

int lteo_rm_tx(const uint8_t *d, int K, int F, int E, int rv, uint8_t *e) {
  int32_t *seq = (int32_t *)malloc(sizeof(int32_t) * 3 * (K + 4));
  int n = lteo_rm_sequence(K, F, rv, seq);
  for (int i = 0; i < E; i++) e[i] = d[seq[i % n]];
  free(seq);
  return n;
}

static void bytes_to_bits(const uint8_t *bytes, int nbits, uint8_t *bits) {
  for (int i = 0; i < nbits; i++) bits[i] = (bytes[i >> 3] >> (7 - (i & 7))) & 1;
}

/* TB -> CRC24A -> segmentation (+CRC24B) -> turbo -> rate match -> concatenate -> scramble */
int lteo_pdsch_encode_bits(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const uint8_t *tb_bytes,
                           uint8_t *e_bits, int *G_out) {
  lteo_cbsegm_t s;
  if (lteo_cbsegm(cfg->tbs, &s)) return -1;
  int nre = lteo_pdsch_re_list(cell, cfg, 0);
  int G = nre * cfg->qm, nl = (cfg->tm == 2) ? 2 : 1;
  uint8_t *b = (uint8_t *)malloc(s.B);
  bytes_to_bits(tb_bytes, cfg->tbs, b);
  uint32_t crc = lteo_crc_bits(b, cfg->tbs, LTEO_CRC24A, 24);
  for (int i = 0; i < 24; i++) b[cfg->tbs + i] = (crc >> (23 - i)) & 1;
  uint8_t *cb = (uint8_t *)malloc(LTEO_MAX_K), *d = (uint8_t *)malloc(3 * (LTEO_MAX_K + 4));
  int rp = 0, wp = 0;
  for (int r = 0; r < s.C; r++) {
    int K = lteo_cb_len(&s, r), F = (r == 0) ? s.F : 0;
    int L = (s.C > 1) ? 24 : 0, n = 0;
    for (int i = 0; i < F; i++) cb[n++] = 0;
    while (n < K - L) cb[n++] = b[rp++];
    if (L) {
      uint32_t c24 = lteo_crc_bits(cb, K - L, LTEO_CRC24B, 24);
      for (int i = 0; i < 24; i++) cb[n++] = (c24 >> (23 - i)) & 1;
    }
    lteo_turbo_encode(cb, K, d);
    int E = lteo_cb_E(&s, G, cfg->qm, nl, r);
    lteo_rm_tx(d, K, F, E, cfg->rv, e_bits + wp);
    wp += E;
  }
  free(b); free(cb); free(d);
  /* 36.211 6.3.1 scrambling, q = 0 */
  uint32_t c_init = ((uint32_t)cfg->rnti << 14) | ((uint32_t)cfg->sf_idx << 9) | (uint32_t)cell->cell_id;
  uint8_t *c = (uint8_t *)malloc(wp);
  lteo_gold(c_init, wp, c);
  for (int i = 0; i < wp; i++) e_bits[i] ^= c[i];
  free(c);
  if (G_out) *G_out = G;
  return (wp == G) ? 0 : -2;
}

/* ---- uplink: UL-SCH coding of srslte_ue_ul_pusch_encode_rnti_softbuffer (phch_worker.cc:545-590 -> srslte_pusch_encode ->
 * srslte_ulsch_encode), no control information multiplexed (36.212 5.2.2: 5.2.2.1 CRC24A, 5.2.2.2 segmentation + CRC24B,
 * 5.2.2.3 turbo coding, 5.2.2.4 rate matching with the whole circular buffer, 5.2.2.5 concatenation, 5.2.2.8 channel
 * interleaver: Qm-bit groups written row by row into a matrix with n_symb columns, read column by column) followed by the
 * PUSCH scrambling of 36.211 5.3.1 (c_init = rnti 2^14 + sf_idx 2^9 + cell_id).  out_bits: G = 12 nof_prb n_symb qm bits.
 * n_symb = 12 (normal CP) or 11 (last symbol punctured for SRS). */
int lteo_ulsch_encode(int tbs, int qm, int nof_prb, int n_symb, int rv, int rnti, int sf_idx, int cell_id,
                      const uint8_t *tb_bytes, uint8_t *out_bits) {
  lteo_cbsegm_t s;
  if (lteo_cbsegm(tbs, &s)) return -1;
  const int rows = 12 * nof_prb, G = rows * n_symb * qm;
  uint8_t *b = (uint8_t *)malloc(s.B), *g = (uint8_t *)malloc(G);
  bytes_to_bits(tb_bytes, tbs, b);
  uint32_t crc = lteo_crc_bits(b, tbs, LTEO_CRC24A, 24);
  for (int i = 0; i < 24; i++) b[tbs + i] = (crc >> (23 - i)) & 1;
  uint8_t *cb = (uint8_t *)malloc(LTEO_MAX_K), *d = (uint8_t *)malloc(3 * (LTEO_MAX_K + 4));
  int rp = 0, wp = 0;
  for (int r = 0; r < s.C; r++) {
    int K = lteo_cb_len(&s, r), F = (r == 0) ? s.F : 0, L = (s.C > 1) ? 24 : 0, n = 0;
    for (int i = 0; i < F; i++) cb[n++] = 0;
    while (n < K - L) cb[n++] = b[rp++];
    if (L) {
      uint32_t c24 = lteo_crc_bits(cb, K - L, LTEO_CRC24B, 24);
      for (int i = 0; i < 24; i++) cb[n++] = (c24 >> (23 - i)) & 1;
    }
    lteo_turbo_encode(cb, K, d);
    int E = lteo_cb_E(&s, G, qm, 1, r);
    lteo_rm_tx(d, K, F, E, rv, g + wp);
    wp += E;
  }
  free(b); free(cb); free(d);
  if (wp != G) { free(g); return -2; }
  /* channel interleaver: symbol k = row * n_symb + col of the input is symbol col * rows + row of the output */
  for (int col = 0; col < n_symb; col++)
    for (int row = 0; row < rows; row++)
      for (int q = 0; q < qm; q++) out_bits[(col * rows + row) * qm + q] = g[(row * n_symb + col) * qm + q];
  free(g);
  uint32_t c_init = ((uint32_t)rnti << 14) | ((uint32_t)sf_idx << 9) | (uint32_t)cell_id;
  uint8_t *c = (uint8_t *)malloc(G);
  lteo_gold(c_init, G, c);
  for (int i = 0; i < G; i++) out_bits[i] ^= c[i];
  free(c);
  return G;
}

/* 36.211 7.1 modulation mapper */
static lteo_cd_t modulate(const uint8_t *b, int qm) {
  lteo_cd_t s;
  if (qm == 2) {
    double a = 1.0 / sqrt(2.0);
    s.re = (1 - 2 * b[0]) * a; s.im = (1 - 2 * b[1]) * a;
  } else if (qm == 4) {
    double a = 1.0 / sqrt(10.0);
    s.re = (1 - 2 * b[0]) * (1 + 2 * b[2]) * a;
    s.im = (1 - 2 * b[1]) * (1 + 2 * b[3]) * a;
  } else {
    double a = 1.0 / sqrt(42.0);
    s.re = (1 - 2 * b[0]) * (4 - (1 - 2 * b[2]) * (2 - (1 - 2 * b[4]))) * a;
    s.im = (1 - 2 * b[1]) * (4 - (1 - 2 * b[3]) * (2 - (1 - 2 * b[5]))) * a;
  }
  return s;
}

/* Resource grid [port][l][k] with PDSCH (TM1 or TM2 SFBC) and CRS; everything else left zero. */
int lteo_pdsch_tx_grid(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const uint8_t *tb_bytes,
                       lteo_cd_t *grid) {
  int nsc = 12 * cell->nof_prb, np = cell->nof_ports;
  memset(grid, 0, sizeof(lteo_cd_t) * np * 14 * nsc);
  int32_t *re = (int32_t *)malloc(sizeof(int32_t) * 14 * nsc);
  int nre = lteo_pdsch_re_list(cell, cfg, re);
  uint8_t *e = (uint8_t *)malloc(nre * cfg->qm + 8);
  int G, rc = lteo_pdsch_encode_bits(cell, cfg, tb_bytes, e, &G);
  if (rc) { free(re); free(e); return rc; }
  if (cfg->tm == 2 && np >= 2) {
    double a = 1.0 / sqrt(2.0);
    for (int i = 0; i + 1 < nre; i += 2) {
      lteo_cd_t x0 = modulate(e + i * cfg->qm, cfg->qm), x1 = modulate(e + (i + 1) * cfg->qm, cfg->qm);
      lteo_cd_t *g0 = grid + (size_t)LTEO_DIV_PA(np, i / 2) * 14 * nsc, *g1 = grid + (size_t)LTEO_DIV_PB(np, i / 2) * 14 * nsc;
      g0[re[i]].re = x0.re * a;      g0[re[i]].im = x0.im * a;
      g1[re[i]].re = -x1.re * a;     g1[re[i]].im = x1.im * a;      /* -conj(x1) */
      g0[re[i + 1]].re = x1.re * a;  g0[re[i + 1]].im = x1.im * a;
      g1[re[i + 1]].re = x0.re * a;  g1[re[i + 1]].im = -x0.im * a; /*  conj(x0) */
    }
  } else {
    for (int i = 0; i < nre; i++) grid[re[i]] = modulate(e + i * cfg->qm, cfg->qm);
  }
  /* cell-specific reference signals */
  int8_t rs[220], is[220];
  int32_t kk[220];
  double a = 1.0 / sqrt(2.0);
  for (int p = 0; p < np; p++)
    for (int l = 0; l < LTEO_NSYMB(cell->cp); l++) {
      int n = lteo_crs_positions(cell, p, l, kk);
      if (!n) continue;
      lteo_crs_values(cell, cfg->sf_idx, l, rs, is);
      for (int m = 0; m < n; m++) {
        lteo_cd_t *g = grid + (size_t)p * 14 * nsc + l * nsc + kk[m];
        g->re = rs[m] * a; g->im = is[m] * a;
      }
    }
  free(re); free(e);
  return 0;
}

void lteo_pcfich_tx(const lteo_cell_t *cell, int sf_idx, int cfi, lteo_cd_t *grid) {
  int nsc = 12 * cell->nof_prb;
  int32_t k[16];
  uint8_t b[32];
  lteo_pcfich_re(cell, k);
  lteo_pcfich_bits(cell, sf_idx, cfi, b);
  if (cell->nof_ports >= 2) {
    double a = 1.0 / sqrt(2.0);
    for (int i = 0; i < 16; i += 2) {
      lteo_cd_t *g0 = grid + (size_t)LTEO_DIV_PA(cell->nof_ports, i / 2) * 14 * nsc, *g1 = grid + (size_t)LTEO_DIV_PB(cell->nof_ports, i / 2) * 14 * nsc;
      lteo_cd_t x0 = modulate(b + 2 * i, 2), x1 = modulate(b + 2 * i + 2, 2);
      g0[k[i]].re = x0.re * a;      g0[k[i]].im = x0.im * a;
      g1[k[i]].re = -x1.re * a;     g1[k[i]].im = x1.im * a;
      g0[k[i + 1]].re = x1.re * a;  g0[k[i + 1]].im = x1.im * a;
      g1[k[i + 1]].re = x0.re * a;  g1[k[i + 1]].im = -x0.im * a;
    }
  } else {
    for (int i = 0; i < 16; i++) grid[k[i]] = modulate(b + 2 * i, 2);
  }
}

/* PDCCH (36.211 6.8): multiplex the encoded DCIs at their CCE positions (<NIL> elsewhere), scramble, QPSK, quadruplet
 * interleaver + cyclic shift, map to the control-region REGs (TX diversity with 2 ports) */
int lteo_pdcch_tx(const lteo_cell_t *cell, int sf_idx, int cfi, int ng_x6, int n_dci, const lteo_dci_tx_t *list,
                  lteo_cd_t *grid) {
  int nsc = 12 * cell->nof_prb, max_reg = 4 * 3 * cell->nof_prb;
  int32_t *rk = (int32_t *)malloc(sizeof(int32_t) * max_reg), *rl = (int32_t *)malloc(sizeof(int32_t) * max_reg);
  int n_reg = lteo_pdcch_regs(cell, cfi, ng_x6, rk, rl), n_cce = n_reg / 9, nb = 8 * n_reg;
  uint8_t *b = (uint8_t *)malloc(nb), *c = (uint8_t *)malloc(nb);
  memset(b, 2, nb);                                             /* 2 = <NIL>: nothing transmitted */
  for (int i = 0; i < n_dci; i++) {
    if (list[i].ncce < 0 || list[i].ncce + list[i].L > n_cce) { free(rk); free(rl); free(b); free(c); return -1; }
    lteo_dci_encode(list[i].bits, list[i].nof_bits, list[i].rnti, 72 * list[i].L, b + 72 * list[i].ncce);
  }
  lteo_gold(((uint32_t)sf_idx << 9) + (uint32_t)cell->cell_id, nb, c);
  for (int i = 0; i < nb; i++) if (b[i] < 2) b[i] ^= c[i];
  int32_t *src = (int32_t *)malloc(sizeof(int32_t) * n_reg);
  lteo_pdcch_quad_perm(n_reg, cell->cell_id, src);
  double a = 1.0 / sqrt(2.0);
  for (int m = 0; m < n_reg; m++) {
    const uint8_t *q = b + 8 * src[m];
    if (q[0] == 2) continue;                                    /* CCEs are whole: a quadruplet is all data or all NIL */
    int32_t k4[4];
    lteo_reg_res(cell, rk[m], rl[m], k4);
    lteo_cd_t x[4];
    for (int i = 0; i < 4; i++) x[i] = modulate(q + 2 * i, 2);
    lteo_cd_t *g0 = grid + rl[m] * nsc, *g1 = grid + 14 * nsc + rl[m] * nsc;
    if (cell->nof_ports >= 2) {
      for (int i = 0; i < 4; i += 2) {
        g0 = grid + (size_t)LTEO_DIV_PA(cell->nof_ports, i / 2) * 14 * nsc + rl[m] * nsc;
        g1 = grid + (size_t)LTEO_DIV_PB(cell->nof_ports, i / 2) * 14 * nsc + rl[m] * nsc;
        g0[k4[i]].re = x[i].re * a;          g0[k4[i]].im = x[i].im * a;
        g1[k4[i]].re = -x[i + 1].re * a;     g1[k4[i]].im = x[i + 1].im * a;
        g0[k4[i + 1]].re = x[i + 1].re * a;  g0[k4[i + 1]].im = x[i + 1].im * a;
        g1[k4[i + 1]].re = x[i].re * a;      g1[k4[i + 1]].im = -x[i].im * a;
      }
    } else {
      for (int i = 0; i < 4; i++) g0[k4[i]] = x[i];
    }
  }
  free(rk); free(rl); free(b); free(c); free(src);
  return n_cce;
}

static void fft_d(lteo_cd_t *x, int n, int inverse, const lteo_cd_t *tab, int ntab) {
  /* plain recursive double-precision FFT, any n = 2^a * 3^b (generator only); tab[i] = exp(-2 pi i/ntab) */
  if (n == 1) return;
  int r = (n % 2 == 0) ? 2 : 3, m = n / r, stride = ntab / n;
  lteo_cd_t *t = (lteo_cd_t *)malloc(sizeof(lteo_cd_t) * n);
  for (int q = 0; q < r; q++)
    for (int i = 0; i < m; i++) t[q * m + i] = x[i * r + q];
  for (int q = 0; q < r; q++) fft_d(t + q * m, m, inverse, tab, ntab);
  for (int k = 0; k < n; k++) {
    double ar = 0, ai = 0;
    for (int q = 0; q < r; q++) {
      lteo_cd_t w = tab[(int)((int64_t)q * k % n) * stride];
      double c = w.re, s = inverse ? -w.im : w.im;
      lteo_cd_t v = t[q * m + k % m];
      ar += v.re * c - v.im * s;
      ai += v.re * s + v.im * c;
    }
    x[k].re = ar; x[k].im = ai;
  }
  free(t);
}

/* OFDM modulation of one port's grid: unitary IFFT (scale 1/sqrt(N)) */
void lteo_ofdm_tx(int nof_prb, const lteo_cd_t *grid, lteo_cd_t *iq) { lteo_ofdm_tx_cp(nof_prb, 0, grid, iq); }

void lteo_ofdm_tx_cp(int nof_prb, int cpx, const lteo_cd_t *grid, lteo_cd_t *iq) {
  int n = lteo_symbol_sz(nof_prb), nsc = 12 * nof_prb, pos = 0;
  lteo_cd_t *x = (lteo_cd_t *)malloc(sizeof(lteo_cd_t) * n);
  lteo_cd_t *tab = (lteo_cd_t *)malloc(sizeof(lteo_cd_t) * n);
  for (int i = 0; i < n; i++) { tab[i].re = cos(-2.0 * M_PI * i / n); tab[i].im = sin(-2.0 * M_PI * i / n); }
  double sc = 1.0 / sqrt((double)n);
  for (int l = 0; l < LTEO_NSYMB(cpx); l++) {
    memset(x, 0, sizeof(lteo_cd_t) * n);
    for (int k = 0; k < nsc; k++) {
      int bin = (k < nsc / 2) ? (n - nsc / 2 + k) : (k - nsc / 2 + 1);
      x[bin] = grid[l * nsc + k];
    }
    fft_d(x, n, 1, tab, n);
    int cp = lteo_cp_len_x(n, l, cpx);
    for (int i = 0; i < cp; i++) { iq[pos].re = x[n - cp + i].re * sc; iq[pos].im = x[n - cp + i].im * sc; pos++; }
    for (int i = 0; i < n; i++) { iq[pos].re = x[i].re * sc; iq[pos].im = x[i].im * sc; pos++; }
  }
  free(x); free(tab);
}

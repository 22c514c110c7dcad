/*
 * lteo_common.c -- tables and small integer helpers of the CPU oracle (TEST INFRASTRUCTURE, see
 * lte_oracle.h).  Follows 3GPP TS 36.211 / 36.212 Rel-8; the reference (srsUE) reaches these through
 * the external srsLTE calls at /root/reference/ue/src/phy/phch_worker.cc:254,337,347-348.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "lte_oracle.h"

/* TS 36.212 Table 5.1.3-3 (K, f1, f2); every row checked to be a permutation in tests/ */
static const uint16_t qpp_tab[188][3] = {
  {40,3,10},
  {48,7,12},
  {56,19,42},
  {64,7,16},
  {72,7,18},
  {80,11,20},
  {88,5,22},
  {96,11,24},
  {104,7,26},
  {112,41,84},
  {120,103,90},
  {128,15,32},
  {136,9,34},
  {144,17,108},
  {152,9,38},
  {160,21,120},
  {168,101,84},
  {176,21,44},
  {184,57,46},
  {192,23,48},
  {200,13,50},
  {208,27,52},
  {216,11,36},
  {224,27,56},
  {232,85,58},
  {240,29,60},
  {248,33,62},
  {256,15,32},
  {264,17,198},
  {272,33,68},
  {280,103,210},
  {288,19,36},
  {296,19,74},
  {304,37,76},
  {312,19,78},
  {320,21,120},
  {328,21,82},
  {336,115,84},
  {344,193,86},
  {352,21,44},
  {360,133,90},
  {368,81,46},
  {376,45,94},
  {384,23,48},
  {392,243,98},
  {400,151,40},
  {408,155,102},
  {416,25,52},
  {424,51,106},
  {432,47,72},
  {440,91,110},
  {448,29,168},
  {456,29,114},
  {464,247,58},
  {472,29,118},
  {480,89,180},
  {488,91,122},
  {496,157,62},
  {504,55,84},
  {512,31,64},
  {528,17,66},
  {544,35,68},
  {560,227,420},
  {576,65,96},
  {592,19,74},
  {608,37,76},
  {624,41,234},
  {640,39,80},
  {656,185,82},
  {672,43,252},
  {688,21,86},
  {704,155,44},
  {720,79,120},
  {736,139,92},
  {752,23,94},
  {768,217,48},
  {784,25,98},
  {800,17,80},
  {816,127,102},
  {832,25,52},
  {848,239,106},
  {864,17,48},
  {880,137,110},
  {896,215,112},
  {912,29,114},
  {928,15,58},
  {944,147,118},
  {960,29,60},
  {976,59,122},
  {992,65,124},
  {1008,55,84},
  {1024,31,64},
  {1056,17,66},
  {1088,171,204},
  {1120,67,140},
  {1152,35,72},
  {1184,19,74},
  {1216,39,76},
  {1248,19,78},
  {1280,199,240},
  {1312,21,82},
  {1344,211,252},
  {1376,21,86},
  {1408,43,88},
  {1440,149,60},
  {1472,45,92},
  {1504,49,846},
  {1536,71,48},
  {1568,13,28},
  {1600,17,80},
  {1632,25,102},
  {1664,183,104},
  {1696,55,954},
  {1728,127,96},
  {1760,27,110},
  {1792,29,112},
  {1824,29,114},
  {1856,57,116},
  {1888,45,354},
  {1920,31,120},
  {1952,59,610},
  {1984,185,124},
  {2016,113,420},
  {2048,31,64},
  {2112,17,66},
  {2176,171,136},
  {2240,209,420},
  {2304,253,216},
  {2368,367,444},
  {2432,265,456},
  {2496,181,468},
  {2560,39,80},
  {2624,27,164},
  {2688,127,504},
  {2752,143,172},
  {2816,43,88},
  {2880,29,300},
  {2944,45,92},
  {3008,157,188},
  {3072,47,96},
  {3136,13,28},
  {3200,111,240},
  {3264,443,204},
  {3328,51,104},
  {3392,51,212},
  {3456,451,192},
  {3520,257,220},
  {3584,57,336},
  {3648,313,228},
  {3712,271,232},
  {3776,179,236},
  {3840,331,120},
  {3904,363,244},
  {3968,375,248},
  {4032,127,168},
  {4096,31,64},
  {4160,33,130},
  {4224,43,264},
  {4288,33,134},
  {4352,477,408},
  {4416,35,138},
  {4480,233,280},
  {4544,357,142},
  {4608,337,480},
  {4672,37,146},
  {4736,71,444},
  {4800,71,120},
  {4864,37,152},
  {4928,39,462},
  {4992,127,234},
  {5056,39,158},
  {5120,39,80},
  {5184,31,96},
  {5248,113,902},
  {5312,41,166},
  {5376,251,336},
  {5440,43,170},
  {5504,21,86},
  {5568,43,174},
  {5632,45,176},
  {5696,45,178},
  {5760,161,120},
  {5824,89,182},
  {5888,323,184},
  {5952,47,186},
  {6016,23,94},
  {6080,47,190},
  {6144,263,480}
};

int lteo_qpp_table_size(void) { return 188; }
int lteo_qpp_K(int idx) { return (idx >= 0 && idx < 188) ? qpp_tab[idx][0] : -1; }

int lteo_qpp_params(int K, int *f1, int *f2) {
  for (int i = 0; i < 188; i++)
    if (qpp_tab[i][0] == K) { *f1 = qpp_tab[i][1]; *f2 = qpp_tab[i][2]; return 0; }
  return -1;
}

void lteo_qpp_perm(int K, uint16_t *pi) {
  int f1, f2;
  if (lteo_qpp_params(K, &f1, &f2)) return;
  for (int64_t i = 0; i < K; i++) pi[i] = (uint16_t)((f1 * i + (int64_t)f2 * i * i) % K);
}

/* SPEC.md 7.3: window length W(K) of the parallel-window (NII) max-log-MAP decoder; P = K / W windows.
 * Candidates are the divisors of K that are multiples of 8.  Among candidates in [64, 128] the largest
 * one giving an even P wins, else the largest one; if there is none in that range, the smallest
 * candidate >= 64 (possibly K itself, i.e. a single window = exact full-length recursions). */
int lteo_window_len(int K) {
  int best_even = 0, best_any = 0;
  for (int w = 64; w <= 128 && w <= K; w += 8)
    if (K % w == 0) { best_any = w; if (((K / w) & 1) == 0) best_even = w; }
  if (best_even) return best_even;
  if (best_any) return best_any;
  for (int w = 64; w <= K; w += 8) if (K % w == 0) return w;
  return K;
}

int lteo_symbol_sz(int nof_prb) {
  if (nof_prb <= 0) return -1;
  if (nof_prb <= 6) return 128;
  if (nof_prb <= 15) return 256;
  if (nof_prb <= 25) return 512;
  if (nof_prb <= 50) return 1024;
  if (nof_prb <= 75) return 1536;
  if (nof_prb <= 110) return 2048;
  return -1;
}

int lteo_cp_len(int nfft, int l) { return ((l % 7) == 0 ? 160 : 144) * nfft / 2048; }
/* 36.211 6.12: extended cyclic prefix = 512 samples at 30.72 Msps for each of the 6 symbols of a slot */
int lteo_cp_len_x(int nfft, int l, int cp) { return cp ? 512 * nfft / 2048 : lteo_cp_len(nfft, l); }

/* bitwise CRC, zero initial state, no final xor (36.212 5.1.1) */
uint32_t lteo_crc_bits(const uint8_t *bits, int n, uint32_t poly, int order) {
  uint32_t reg = 0, top = 1u << order, mask = top - 1;
  for (int i = 0; i < n; i++) {
    reg = (reg << 1) | (bits[i] & 1);
    if (reg & top) reg ^= poly;
  }
  for (int i = 0; i < order; i++) {
    reg <<= 1;
    if (reg & top) reg ^= poly;
  }
  return reg & mask;
}

/* 36.211 7.2 length-31 Gold sequence, Nc = 1600 */
void lteo_gold(uint32_t c_init, int n, uint8_t *c) {
  int total = n + 1600 + 31;
  uint8_t *x1 = (uint8_t *)calloc(total, 1), *x2 = (uint8_t *)calloc(total, 1);
  x1[0] = 1;
  for (int i = 0; i < 31; i++) x2[i] = (c_init >> i) & 1;
  for (int i = 0; i < total - 31; i++) {
    x1[i + 31] = x1[i + 3] ^ x1[i];
    x2[i + 31] = x2[i + 3] ^ x2[i + 2] ^ x2[i + 1] ^ x2[i];
  }
  for (int i = 0; i < n; i++) c[i] = x1[i + 1600] ^ x2[i + 1600];
  free(x1); free(x2);
}

static int next_K(int b) { /* smallest table K >= b */
  for (int i = 0; i < 188; i++) if (qpp_tab[i][0] >= b) return qpp_tab[i][0];
  return -1;
}
static int prev_K(int k) {
  for (int i = 187; i >= 0; i--) if (qpp_tab[i][0] < k) return qpp_tab[i][0];
  return -1;
}

/* 36.212 5.1.2 */
int lteo_cbsegm(int tbs, lteo_cbsegm_t *s) {
  memset(s, 0, sizeof(*s));
  if (tbs <= 0) return -1;
  int B = tbs + 24, Bp, C;
  if (B <= 6144) { C = 1; Bp = B; }
  else { C = (B + 6119) / 6120; Bp = B + 24 * C; }
  int Kp = next_K((Bp + C - 1) / C);
  if (Kp < 0) return -1;
  int Km = 0, Cm = 0, Cp = C;
  if (C > 1) {
    Km = prev_K(Kp);
    int dK = Kp - Km;
    Cm = (C * Kp - Bp) / dK;
    Cp = C - Cm;
  }
  s->tbs = tbs; s->B = B; s->C = C; s->Kp = Kp; s->Km = Km; s->Cp = Cp; s->Cm = Cm;
  s->F = Cp * Kp + Cm * Km - Bp;
  return 0;
}

int lteo_cb_len(const lteo_cbsegm_t *s, int r) { return r < s->Cm ? s->Km : s->Kp; }

/* 36.212 5.1.4.1.2: rate-matching output size of code block r */
int lteo_cb_E(const lteo_cbsegm_t *s, int G, int qm, int nl, int r) {
  int Gp = G / (nl * qm), gamma = Gp % s->C;
  if (r <= s->C - gamma - 1) return nl * qm * (Gp / s->C);
  return nl * qm * ((Gp + s->C - 1) / s->C);
}

/* CRS positions of antenna port `port` in OFDM symbol l (0..13, or 0..11 with the extended cyclic prefix): 2*nof_prb
 * subcarrier indices, or 0 if the symbol carries no CRS for this port (ports 0/1: symbols 0 and N_symb - 3 of each slot;
 * 36.211 6.10.1.2) */
int lteo_crs_positions(const lteo_cell_t *cell, int port, int l, int32_t *k_out) {
  int nslot = LTEO_NSLOT(cell->cp), ls = l % nslot;
  int v;
  if (port >= 2) {                     /* ports 2 / 3: symbol 1 of each slot, v = 3 (n_s mod 2) and 3 + 3 (n_s mod 2) */
    if (ls != 1) return 0;
    v = (3 * (l / nslot) + (port == 3 ? 3 : 0)) % 6;
  } else {
    if (ls != 0 && ls != nslot - 3) return 0;
    v = (port == 0) ? (ls == 0 ? 0 : 3) : (ls == 0 ? 3 : 0);
  }
  int off = (v + cell->cell_id % 6) % 6;
  for (int m = 0; m < 2 * cell->nof_prb; m++) k_out[m] = 6 * m + off;
  return 2 * cell->nof_prb;
}

/* PCFICH (36.211 6.7.4): subcarriers, in OFDM symbol 0, of the 16 symbols d(0..15).  Quadruplet i goes to the
 * resource-element group that starts at k = kbar + floor(i N_RB / 2) * 6 (mod 12 N_RB), kbar = 6 * (N_ID mod 2 N_RB);
 * inside the group the REs with k mod 3 == N_ID mod 3 are reserved for the CRS of ports 0 AND 1 (36.211 6.2.4: both
 * are assumed present for the mapping even with one port), the four others carry the quadruplet in ascending k. */
void lteo_pcfich_re(const lteo_cell_t *cell, int32_t *k16) {
  int nrb = cell->nof_prb, nsc = 12 * nrb, n = 0;
  int kbar = 6 * (cell->cell_id % (2 * nrb));
  for (int i = 0; i < 4; i++) {
    int k0 = (kbar + ((i * nrb) / 2) * 6) % nsc;
    for (int j = 0; j < 6; j++)
      if ((k0 + j) % 3 != cell->cell_id % 3) k16[n++] = k0 + j;
  }
}

/* the 32 scrambled CFI code bits (36.212 5.3.4 code words, 36.211 6.7.1 scrambling) */
void lteo_pcfich_bits(const lteo_cell_t *cell, int sf_idx, int cfi, uint8_t *b32) {
  uint8_t c[32];
  uint32_t c_init = ((uint32_t)(sf_idx + 1) * (uint32_t)(2 * cell->cell_id + 1) << 9) + (uint32_t)cell->cell_id;
  lteo_gold(c_init, 32, c);
  /* code words <0,1,1,...>, <1,0,1,...>, <1,1,0,...>: bit n is 0 where n mod 3 == cfi - 1 */
  for (int n = 0; n < 32; n++) b32[n] = (uint8_t)(((n % 3) != (cfi - 1)) ^ c[n]);
}

/* CRS symbol values for symbol l of subframe sf_idx: r(m') = (re_sign + j im_sign)/sqrt(2),
 * m' = m + 110 - nof_prb, m = 0..2*nof_prb-1 (same sequence for both ports) */
void lteo_crs_values(const lteo_cell_t *cell, int sf_idx, int l, int8_t *re_sign, int8_t *im_sign) {
  int nslot = LTEO_NSLOT(cell->cp), ns = 2 * sf_idx + l / nslot, ls = l % nslot;
  /* 36.211 6.10.1.1: c_init = 2^10 (7 (n_s + 1) + l + 1)(2 N_ID + 1) + 2 N_ID + N_CP, N_CP = 1 normal / 0 extended */
  uint32_t c_init = 1024u * (7 * (ns + 1) + ls + 1) * (2 * cell->cell_id + 1) + 2 * cell->cell_id + (cell->cp ? 0 : 1);
  uint8_t c[440];
  lteo_gold(c_init, 440, c);
  for (int m = 0; m < 2 * cell->nof_prb; m++) {
    int mp = m + 110 - cell->nof_prb;
    re_sign[m] = c[2 * mp] ? -1 : 1;
    im_sign[m] = c[2 * mp + 1] ? -1 : 1;
  }
}

/* Ordered list of PDSCH resource elements (grid index l*nsc + k): symbols after the control region,
 * allocated PRBs of the symbol's slot ascending, subcarriers ascending, skipping the CRS of every configured port and, in
 * subframes 0/5, the PSS/SSS (and PBCH in subframe 0) REs of the six central PRBs (36.211 6.3.5). */
int lteo_pdsch_re_list(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, int32_t *re_idx) {
  int nsc = 12 * cell->nof_prb, n = 0;
  int lstart = cfg->cfi + (cell->nof_prb <= 10 ? 1 : 0);
  int c_lo = nsc / 2 - 36, c_hi = nsc / 2 + 36;       /* central 72 subcarriers */
  const int nslot = LTEO_NSLOT(cell->cp);
  for (int l = lstart; l < 2 * nslot; l++) {
    int ls = l % nslot, crs = (ls == 0 || ls == nslot - 3 || (cell->nof_ports == 4 && ls == 1));
    int o0 = -1, o1 = -1;
    if (crs) {
      int v0 = (ls == 0) ? 0 : 3;          /* symbol 1 (ports 2 / 3): both offsets v and v + 3 are taken, so v0 does not matter */
      o0 = (v0 + cell->cell_id % 6) % 6;
      if (cell->nof_ports > 1) o1 = (o0 + 3) % 6;
    }
    int sync = 0;
    if ((cfg->sf_idx == 0 || cfg->sf_idx == 5) && (l == nslot - 2 || l == nslot - 1)) sync = 1;   /* SSS, PSS */
    if (cfg->sf_idx == 0 && l >= nslot && l <= nslot + 3) sync = 1;                                /* PBCH    */
    for (int prb = 0; prb < cell->nof_prb; prb++) {
      uint8_t pm = cfg->prb_mask[prb];                 /* bit 0: both slots, bit 1: slot 0 only, bit 2: slot 1 only */
      if (!((pm & 1) || (pm & (l >= nslot ? 4 : 2)))) continue;
      for (int k = 12 * prb; k < 12 * prb + 12; k++) {
        if (crs && (k % 6 == o0 || k % 6 == o1)) continue;
        /* with a single configured port only port-0 CRS REs are reserved */
        if (sync && k >= c_lo && k < c_hi) continue;
        if (re_idx) re_idx[n] = l * nsc + k;
        n++;
      }
    }
  }
  return n;
}

/* twiddle table used by BOTH the oracle FFT and (uploaded once) by the GPU FFT kernel:
 * tw[k] = (float) exp(-2 pi i k / n) evaluated in double; the four axis values are forced exact. */
void lteo_fft_twiddles(int n, lteo_cf_t *tw) {
  for (int k = 0; k < n / 2; k++) {
    double a = -2.0 * M_PI * (double)k / (double)n;
    tw[k].re = (float)cos(a);
    tw[k].im = (float)sin(a);
    if ((4 * k) % n == 0) {
      int q = 4 * k / n;          /* 0 or 1 for k < n/2 */
      tw[k].re = (q == 0) ? 1.0f : 0.0f;
      tw[k].im = (q == 0) ? 0.0f : -1.0f;
    }
  }
}

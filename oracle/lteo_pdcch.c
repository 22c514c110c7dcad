/*
 * lteo_pdcch.c -- CPU restatement of the PDCCH receive path (TEST INFRASTRUCTURE, see lte_oracle.h):
 * what srsUE reaches through srslte_pdcch_extract_llr and srslte_ue_dl_find_dl_dci_type
 * (/root/reference/ue/src/phy/phch_worker.cc:260,293) -- control-region resource-element groups, the PDCCH
 * quadruplet interleaver, soft demodulation of all control-channel elements, the UE-specific / common search
 * spaces, rate de-matching, a tail-biting Viterbi decoder and the RNTI-masked CRC16 -- plus the matching encoder
 * used by the synthetic-subframe generator.  Arithmetic contract: oracle/SPEC.md section 10.
 * 3GPP: TS 36.211 6.2.4, 6.7-6.9; TS 36.212 5.1.3.1, 5.1.4.2, 5.3.3; TS 36.213 9.1.1.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lte_oracle.h"

/* number of OFDM symbols of the control region */
int lteo_ctrl_symbols(int nof_prb, int cfi) { return cfi + (nof_prb <= 10 ? 1 : 0); }

/* number of PHICH groups, normal CP: ceil(Ng * N_RB / 8) with Ng = ng_x6 / 6 (1/6, 1/2, 1, 2).  With the extended cyclic
 * prefix there are twice as many groups but two share one mapping unit of three REGs (36.211 6.9.3), so this is also the
 * number of PHICH mapping units the control region reserves there. */
int lteo_phich_groups(int nof_prb, int ng_x6) { return (ng_x6 * nof_prb + 47) / 48; }

/*
 * Resource-element groups of the control region that carry PDCCH, in the mapping order of 36.211 6.8.5
 * (subcarrier k' ascending, then symbol l'): reg_k/reg_l = first subcarrier and symbol of each.  Symbol 0 holds
 * 2 REGs of 6 REs per PRB (4 data REs, the CRS positions of ports 0/1 skipped), the other control symbols 3 REGs
 * of 4 REs per PRB (1 or 2 antenna ports) -- except symbol 3 under the extended cyclic prefix (a four-symbol control
 * region at <= 10 PRB) and symbol 1 of a four-port cell (CRS of ports 2 / 3), which are laid out like symbol 0.  The 4 PCFICH REGs and the 3 REGs of every PHICH group (normal PHICH
 * duration: all in symbol 0, 36.211 6.9.3) are excluded.  Returns the number of REGs.
 */
int lteo_pdcch_regs(const lteo_cell_t *cell, int cfi, int ng_x6, int32_t *reg_k, int32_t *reg_l) {
  int nrb = cell->nof_prb, nsc = 12 * nrb, nsym = lteo_ctrl_symbols(nrb, cfi);
  int n0 = 2 * nrb;                          /* REGs of symbol 0, index = k / 6 */
  uint8_t *used = (uint8_t *)calloc(n0, 1);
  int32_t k16[16];
  lteo_pcfich_re(cell, k16);
  for (int i = 0; i < 4; i++) used[k16[4 * i] / 6] = 1;
  /* PHICH: REGs of symbol 0 not assigned to PCFICH are numbered 0..n'0-1 from the lowest frequency;
   * group m', i = 0..2 takes number (N_ID + m' + floor(i n'0 / 3)) mod n'0 */
  int np0 = n0 - 4;
  int32_t *free_idx = (int32_t *)malloc(sizeof(int32_t) * n0);
  for (int r = 0, n = 0; r < n0; r++) if (!used[r]) free_idx[n++] = r;
  int ngroups = lteo_phich_groups(nrb, ng_x6);
  for (int m = 0; m < ngroups; m++)
    for (int i = 0; i < 3; i++) used[free_idx[(cell->cell_id + m + (i * np0) / 3) % np0]] = 2;
  int n = 0;
  for (int k = 0; k < nsc; k += 2) {         /* REG starts are multiples of 6 (symbol 0) or 4 (others) */
    for (int l = 0; l < nsym; l++) {
      if (l == 0) { if (k % 6 == 0 && !used[k / 6]) { reg_k[n] = k; reg_l[n] = 0; n++; } }
      else if ((cell->cp && l == 3) || (cell->nof_ports == 4 && l == 1)) { if (k % 6 == 0) { reg_k[n] = k; reg_l[n] = l; n++; } }
      else if (k % 4 == 0) { reg_k[n] = k; reg_l[n] = l; n++; }
    }
  }
  free(used); free(free_idx);
  return n;
}

/* the 4 data subcarriers of a REG */
void lteo_reg_res(const lteo_cell_t *cell, int k0, int l, int32_t *k4) {
  if (l == 0 || (cell->cp && l == 3) || (cell->nof_ports == 4 && l == 1)) { for (int j = 0, n = 0; j < 6; j++) if ((k0 + j) % 3 != cell->cell_id % 3) k4[n++] = k0 + j; }
  else for (int j = 0; j < 4; j++) k4[j] = k0 + j;
}

static const uint8_t cc_perm[32] = {1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31,
                                    0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30};

/* 36.212 5.1.4.2.1 sub-block interleaver for convolutionally coded channels on D elements: out[j] = index of the
 * input element read j-th (dummy elements at the head of the matrix are skipped).  Returns D. */
int lteo_cc_interleaver(int D, int32_t *out) {
  int R = (D + 31) / 32, ND = 32 * R - D, n = 0;
  for (int c = 0; c < 32; c++)
    for (int r = 0; r < R; r++) {
      int y = r * 32 + cc_perm[c];
      if (y >= ND) out[n++] = y - ND;
    }
  return n;
}

/* 36.211 6.8.5: mapping position m' (REG m' of lteo_pdcch_regs) carries quadruplet src[m'] of the multiplexed,
 * scrambled, modulated PDCCH stream: interleave, then cyclic shift by N_ID */
void lteo_pdcch_quad_perm(int n_quad, int cell_id, int32_t *src) {
  int32_t *w = (int32_t *)malloc(sizeof(int32_t) * n_quad);
  lteo_cc_interleaver(n_quad, w);
  for (int m = 0; m < n_quad; m++) src[m] = w[(m + cell_id) % n_quad];
  free(w);
}

/* rate-1/3 tail-biting convolutional code, K = 7, G = (133, 171, 165) octal (36.212 5.1.3.1).  out: 3 D bits as
 * streams d0[D] d1[D] d2[D] */
static const int cc_poly[3] = {0133, 0171, 0165};
void lteo_conv_encode(const uint8_t *c, int D, uint8_t *d) {
  int sr = 0;                                     /* bit 5 = most recent input c(k-1), bit 0 = c(k-6) */
  for (int i = 0; i < 6; i++) sr |= (c[D - 1 - i] & 1) << (5 - i);
  for (int k = 0; k < D; k++) {
    int reg = ((c[k] & 1) << 6) | sr;             /* bit 6 = c(k) */
    for (int j = 0; j < 3; j++) d[j * D + k] = (uint8_t)(__builtin_popcount(reg & cc_poly[j]) & 1);
    sr = reg >> 1;
  }
}

/* circular-buffer order of the rate matcher (36.212 5.1.4.2): seq[j] = index into d (stream-major, 3 D) of the j-th
 * position of w = v0 | v1 | v2 without <NULL>s; returns 3 D */
int lteo_cc_rm_sequence(int D, int32_t *seq) {
  int32_t *p = (int32_t *)malloc(sizeof(int32_t) * D);
  lteo_cc_interleaver(D, p);
  int n = 0;
  for (int s = 0; s < 3; s++)
    for (int j = 0; j < D; j++) seq[n++] = s * D + p[j];
  free(p);
  return n;
}

/* DCI payload -> CRC16 masked with the RNTI -> convolutional code -> rate matching to E bits */
void lteo_dci_encode(const uint8_t *bits, int nof_bits, uint16_t rnti, int E, uint8_t *e) {
  int D = nof_bits + 16;
  uint8_t *c = (uint8_t *)malloc(D), *d = (uint8_t *)malloc(3 * D);
  int32_t *seq = (int32_t *)malloc(sizeof(int32_t) * 3 * D);
  memcpy(c, bits, nof_bits);
  uint32_t crc = lteo_crc_bits(bits, nof_bits, LTEO_CRC16, 16) ^ rnti;
  for (int i = 0; i < 16; i++) c[nof_bits + i] = (uint8_t)((crc >> (15 - i)) & 1);
  lteo_conv_encode(c, D, d);
  int n = lteo_cc_rm_sequence(D, seq);
  for (int k = 0; k < E; k++) e[k] = d[seq[k % n]];
  free(c); free(d); free(seq);
}

/* UE-specific (rnti != 0 path) or common search space of 36.213 9.1.1: candidates as (L, first CCE); returns count */
int lteo_pdcch_search_space(int nof_cce, int sf_idx, uint16_t rnti, int common, int32_t *cand_L, int32_t *cand_ncce) {
  static const int L_ue[4] = {1, 2, 4, 8}, M_ue[4] = {6, 6, 2, 2};
  static const int L_c[2] = {4, 8}, M_c[2] = {4, 2};
  int n = 0;
  if (common) {
    for (int a = 0; a < 2; a++) {
      int L = L_c[a], lim = nof_cce < 16 ? nof_cce : 16;
      for (int m = 0; m < M_c[a]; m++)
        if ((m + 1) * L <= lim) { cand_L[n] = L; cand_ncce[n] = m * L; n++; }
    }
    return n;
  }
  uint32_t Y = rnti;
  for (int k = 0; k <= sf_idx; k++) Y = (39827u * Y) % 65537u;
  for (int a = 0; a < 4; a++) {
    int L = L_ue[a], nl = nof_cce / L;
    if (nl < 1) continue;
    for (int m = 0; m < M_ue[a] && m < nl; m++) {
      cand_L[n] = L; cand_ncce[n] = L * (int)((Y + (uint32_t)m) % (uint32_t)nl); n++;
    }
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * Receiver
 * ---------------------------------------------------------------------------------------------- */
/* LLRs of every PDCCH bit of the subframe, 72 per CCE in CCE order (8 * n_reg values; the REGs beyond the last
 * whole CCE included), equalised with the PDSCH formulas (SPEC.md 4), QPSK demapper (5), descrambled with
 * c_init = sf_idx * 2^9 + N_ID.  Returns the number of CCEs. */
int lteo_pdcch_extract_llr(const lteo_cell_t *cell, int sf_idx, int cfi, int ng_x6, const lteo_cf_t *sf,
                           const lteo_cf_t *ce, float n0, int16_t *llr) {
  int nsc = 12 * cell->nof_prb, max_reg = 4 * 3 * cell->nof_prb;
  int32_t *rk = (int32_t *)malloc(sizeof(int32_t) * max_reg), *rl = (int32_t *)malloc(sizeof(int32_t) * max_reg);
  int n_reg = lteo_pdcch_regs(cell, cfi, ng_x6, rk, rl);
  int32_t *src = (int32_t *)malloc(sizeof(int32_t) * n_reg);
  lteo_pdcch_quad_perm(n_reg, cell->cell_id, src);
  const float sq2 = (float)sqrt(2.0);
  for (int m = 0; m < n_reg; m++) {
    int32_t k4[4];
    lteo_cf_t d[4];
    lteo_reg_res(cell, rk[m], rl[m], k4);
    const lteo_cf_t *y = sf + rl[m] * nsc, *h0p = ce + rl[m] * nsc, *h1p;
    if (cell->nof_ports >= 2) {
      for (int i = 0; i < 4; i += 2) {
        h0p = ce + (size_t)LTEO_DIV_PA(cell->nof_ports, i / 2) * 14 * nsc + rl[m] * nsc;
        h1p = ce + (size_t)LTEO_DIV_PB(cell->nof_ports, i / 2) * 14 * nsc + rl[m] * nsc;
        lteo_cf_t r0 = y[k4[i]], r1 = y[k4[i + 1]], h0 = h0p[k4[i]], h1 = h1p[k4[i]];
        float den = ((h0.re * h0.re + h0.im * h0.im) + (h1.re * h1.re + h1.im * h1.im)) + n0;
        float a_re = h0.re * r0.re + h0.im * r0.im, a_im = h0.re * r0.im - h0.im * r0.re;
        float b_re = h1.re * r1.re + h1.im * r1.im, b_im = h1.im * r1.re - h1.re * r1.im;
        float c_re = h0.re * r1.re + h0.im * r1.im, c_im = h0.re * r1.im - h0.im * r1.re;
        float e_re = h1.re * r0.re + h1.im * r0.im, e_im = h1.im * r0.re - h1.re * r0.im;
        d[i].re = ((a_re + b_re) * sq2) / den;     d[i].im = ((a_im + b_im) * sq2) / den;
        d[i + 1].re = ((c_re - e_re) * sq2) / den; d[i + 1].im = ((c_im - e_im) * sq2) / den;
      }
    } else {
      for (int i = 0; i < 4; i++) {
        lteo_cf_t yy = y[k4[i]], h = h0p[k4[i]];
        float den = (h.re * h.re + h.im * h.im) + n0;
        d[i].re = (yy.re * h.re + yy.im * h.im) / den;
        d[i].im = (yy.im * h.re - yy.re * h.im) / den;
      }
    }
    lteo_demod(d, 4, 2, llr + 8 * src[m]);
  }
  lteo_descramble(llr, 8 * n_reg, ((uint32_t)sf_idx << 9) + (uint32_t)cell->cell_id);
  free(rk); free(rl); free(src);
  return n_reg / 9;
}

/*
 * One blind-decoding attempt: the E = 72 L LLRs of the candidate are de-rate-matched (int32 accumulation of
 * repeated positions, punctured positions 0), decoded by a tail-biting Viterbi decoder and the 16-bit remainder
 * CRC(payload) xor received-CRC is returned: it equals the RNTI the DCI was addressed to when decoding succeeded.
 * Viterbi (SPEC.md 10): the D trellis steps are run three times in a row from all-zero path metrics with int32
 * metrics, branch metric = sum_j (code bit ? +LLR : -LLR), ties keep the predecessor with input-history bit 0;
 * trace back from the best final state (lowest index on ties); the middle D decisions are the output.
 */
/* tail-biting Viterbi over soft[3 D] (D = nof_bits + 16) + CRC16: writes the D decided bits to bits_out (payload first)
 * and returns CRC16(payload) xor received CRC */
uint16_t lteo_viterbi_crc16(const int32_t *soft, int nof_bits, uint8_t *bits_out) {
  int D = nof_bits + 16, T = 3 * D;
  /* state = the 6 previous inputs, bit 5 = most recent.  Next state after input u: (u << 5) | (state >> 1). */
  int32_t pm[64], nm[64];
  uint8_t *surv = (uint8_t *)malloc((size_t)T * 64);      /* surv[t][ns] = dropped oldest bit of the chosen predecessor */
  memset(pm, 0, sizeof(pm));
  for (int t = 0; t < T; t++) {
    int k = t % D;
    int32_t s0 = soft[k], s1 = soft[D + k], s2 = soft[2 * D + k];
    for (int ns = 0; ns < 64; ns++) {
      int u = ns >> 5;
      int32_t best = 0; int bb = 0;
      for (int b = 0; b < 2; b++) {                         /* b = oldest bit of the predecessor state */
        int ps = ((ns & 31) << 1) | b;
        int reg = (u << 6) | ps;
        int32_t bm = (__builtin_popcount(reg & cc_poly[0]) & 1 ? s0 : -s0) + (__builtin_popcount(reg & cc_poly[1]) & 1 ? s1 : -s1) +
                     (__builtin_popcount(reg & cc_poly[2]) & 1 ? s2 : -s2);
        int32_t v = pm[ps] + bm;
        if (b == 0 || v > best) { best = v; bb = b; }
      }
      nm[ns] = best;
      surv[(size_t)t * 64 + ns] = (uint8_t)bb;
    }
    memcpy(pm, nm, sizeof(pm));
  }
  int st = 0;
  for (int s = 1; s < 64; s++) if (pm[s] > pm[st]) st = s;
  uint8_t *dec = (uint8_t *)malloc(T);
  for (int t = T - 1; t >= 0; t--) {
    dec[t] = (uint8_t)(st >> 5);                            /* the input that led into st */
    st = ((st & 31) << 1) | surv[(size_t)t * 64 + st];
  }
  uint8_t *c = dec + D;                                     /* middle repetition */
  memcpy(bits_out, c, D);
  uint32_t crc = lteo_crc_bits(c, nof_bits, LTEO_CRC16, 16), rx = 0;
  for (int i = 0; i < 16; i++) rx = (rx << 1) | c[nof_bits + i];
  free(surv); free(dec);
  return (uint16_t)(crc ^ rx);
}

uint16_t lteo_pdcch_decode_candidate(const int16_t *llr, int L, int nof_bits, uint8_t *bits_out) {
  int D = nof_bits + 16, E = 72 * L;
  int32_t *seq = (int32_t *)malloc(sizeof(int32_t) * 3 * D);
  int32_t *soft = (int32_t *)calloc(3 * D, sizeof(int32_t));
  uint8_t *bits = (uint8_t *)malloc(D);
  int n = lteo_cc_rm_sequence(D, seq);
  for (int k = 0; k < E; k++) soft[seq[k % n]] += llr[k];
  uint16_t rem = lteo_viterbi_crc16(soft, nof_bits, bits);
  memcpy(bits_out, bits, nof_bits);
  free(seq); free(soft); free(bits);
  return rem;
}

/* blind search over the UE-specific (common = 0) or common search space for a DCI of nof_bits addressed to rnti.
 * Returns 1 and fills bits_out / found_L / found_ncce for the first matching candidate in search-space order,
 * 0 if none (the convention of srslte_ue_dl_find_dl_dci_type, phch_worker.cc:293). */
int lteo_pdcch_find_dci(const int16_t *llr, int nof_cce, int sf_idx, uint16_t rnti, int common, int nof_bits,
                        uint8_t *bits_out, int *found_L, int *found_ncce) {
  int32_t cl[32], cn[32];
  int n = lteo_pdcch_search_space(nof_cce, sf_idx, rnti, common, cl, cn);
  for (int i = 0; i < n; i++) {
    if (lteo_pdcch_decode_candidate(llr + 72 * cn[i], cl[i], nof_bits, bits_out) == rnti) {
      if (found_L) *found_L = cl[i];
      if (found_ncce) *found_ncce = cn[i];
      return 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * PHICH (36.211 6.9; srslte_ue_dl_decode_phich, phch_worker.cc:381).  SPEC.md 11.
 * ---------------------------------------------------------------------------------------------- */
/* the 12 subcarriers (OFDM symbol 0) of PHICH group n_group: 3 REGs x 4 data REs, normal PHICH duration.  With the
 * extended cyclic prefix groups 2m' and 2m'+1 share mapping unit m' (36.211 6.9.3). */
void lteo_phich_res(const lteo_cell_t *cell, int ng_x6, int n_group, int32_t *k12) {
  int nrb = cell->nof_prb, n0 = 2 * nrb;
  if (cell->cp) n_group /= 2;
  uint8_t *used = (uint8_t *)calloc(n0, 1);
  int32_t k16[16];
  lteo_pcfich_re(cell, k16);
  for (int i = 0; i < 4; i++) used[k16[4 * i] / 6] = 1;
  int32_t *free_idx = (int32_t *)malloc(sizeof(int32_t) * n0);
  int np0 = 0;
  for (int r = 0; r < n0; r++) if (!used[r]) free_idx[np0++] = r;
  (void)ng_x6;
  for (int i = 0; i < 3; i++) {
    int reg = free_idx[(cell->cell_id + n_group + (i * np0) / 3) % np0];
    lteo_reg_res(cell, 6 * reg, 0, k12 + 4 * i);
  }
  free(used); free(free_idx);
}

/* (n_group, n_seq) of the PHICH that answers an uplink transmission (36.213 9.1.2, FDD, normal CP) */
void lteo_phich_index(int nof_prb, int ng_x6, int I_lowest, int n_dmrs, int *n_group, int *n_seq) {
  int ngroups = lteo_phich_groups(nof_prb, ng_x6);
  *n_group = (I_lowest + n_dmrs) % ngroups;
  *n_seq = (I_lowest / ngroups + n_dmrs) % 8;
}

/* the same for either cyclic prefix: N_group = 2 ceil(Ng N_RB / 8) and n_seq modulo 2 N_SF = 4 with the extended one */
void lteo_phich_index_cp(int nof_prb, int ng_x6, int cp, int I_lowest, int n_dmrs, int *n_group, int *n_seq) {
  int ngroups = (cp ? 2 : 1) * lteo_phich_groups(nof_prb, ng_x6);
  *n_group = (I_lowest + n_dmrs) % ngroups;
  *n_seq = (I_lowest / ngroups + n_dmrs) % (cp ? 4 : 8);
}

/* extended cyclic prefix (36.211 Table 6.9.1-2, N_SF = 2): [+1 +1], [+1 -1], [+j +j], [+j -j] */
static void phich_w2(int n_seq, int i, int *re, int *im) {
  int v = ((n_seq & 1) && i) ? -1 : 1;
  if (n_seq < 2) { *re = v; *im = 0; } else { *re = 0; *im = v; }
}

/* orthogonal sequence element w(i), i < 4, of sequence n_seq < 8 as (re, im) in {0, +-1} */
static void phich_w(int n_seq, int i, int *re, int *im) {
  static const int8_t w4[4][4] = {{1, 1, 1, 1}, {1, -1, 1, -1}, {1, 1, -1, -1}, {1, -1, -1, 1}};
  int v = w4[n_seq & 3][i];
  if (n_seq < 4) { *re = v; *im = 0; } else { *re = 0; *im = v; }
}

/* adds one PHICH (ack = 1: HI 1) to a grid; several PHICHs of a group superpose */
void lteo_phich_tx(const lteo_cell_t *cell, int sf_idx, int ng_x6, int n_group, int n_seq, int ack, lteo_cd_t *grid) {
  int nsc = 12 * cell->nof_prb;
  int32_t k[12];
  uint8_t c[12];
  lteo_phich_res(cell, ng_x6, n_group, k);
  lteo_gold(((uint32_t)(sf_idx + 1) * (uint32_t)(2 * cell->cell_id + 1) << 9) + (uint32_t)cell->cell_id, 12, c);
  double a = 1.0 / sqrt(2.0);
  lteo_cd_t d[12];
  for (int i = 0; i < 12; i++) {
    int wr, wi;
    phich_w(n_seq, i % 4, &wr, &wi);
    double z = (ack ? -1.0 : 1.0) * a * (c[i] ? -1.0 : 1.0);     /* BPSK: z (1 + j), real factor z */
    /* w * z(1+j): (wr + j wi)(1 + j) = (wr - wi) + j (wr + wi) */
    d[i].re = z * (wr - wi); d[i].im = z * (wr + wi);
  }
  if (cell->cp) {
    /* N_SF = 2: six symbols d(i) = w(i mod 2)(1 - 2 c(i)) z; resource-group alignment (36.211 6.9.2): an even group
     * fills the first two elements of each quadruplet, an odd one the last two */
    lteo_cd_t d6[6];
    for (int i = 0; i < 6; i++) {
      int wr, wi;
      phich_w2(n_seq, i % 2, &wr, &wi);
      double z = (ack ? -1.0 : 1.0) * a * (c[i] ? -1.0 : 1.0);
      d6[i].re = z * (wr - wi); d6[i].im = z * (wr + wi);
    }
    memset(d, 0, sizeof(d));
    for (int i = 0; i < 3; i++) { d[4 * i + 2 * (n_group & 1)] = d6[2 * i]; d[4 * i + 2 * (n_group & 1) + 1] = d6[2 * i + 1]; }
  }
  lteo_cd_t *g0 = grid, *g1 = grid + 14 * nsc;
  if (cell->nof_ports >= 2) {
    for (int i = 0; i < 12; i += 2) {
      /* four ports (36.211 6.9.2): quadruplet i / 4 goes out on ports (0, 2) or (1, 3) as a whole, alternating with
       * i / 4 + n_group (normal prefix) or i / 4 + n_group / 2 (extended) */
      const int par = i / 4 + (cell->cp ? n_group / 2 : n_group);
      g0 = grid + (size_t)LTEO_DIV_PA(cell->nof_ports, par) * 14 * nsc; g1 = grid + (size_t)LTEO_DIV_PB(cell->nof_ports, par) * 14 * nsc;
      g0[k[i]].re += d[i].re * a;          g0[k[i]].im += d[i].im * a;
      g1[k[i]].re += -d[i + 1].re * a;     g1[k[i]].im += d[i + 1].im * a;
      g0[k[i + 1]].re += d[i + 1].re * a;  g0[k[i + 1]].im += d[i + 1].im * a;
      g1[k[i + 1]].re += d[i].re * a;      g1[k[i + 1]].im += -d[i].im * a;
    }
  } else {
    for (int i = 0; i < 12; i++) { g0[k[i]].re += d[i].re; g0[k[i]].im += d[i].im; }
  }
}

/* returns the HI (1 = ACK).  The 12 REs are equalised with the SPEC.md 4 formulas, multiplied by conj(w) (exact sign
 * swaps) and the scrambling sign, m_i = re + im, metric = ((m_0 + m_1) + m_2) + ... in float; ACK iff metric < 0 */
int lteo_phich_decode(const lteo_cell_t *cell, int sf_idx, int ng_x6, const lteo_cf_t *sf, const lteo_cf_t *ce, float n0,
                      int n_group, int n_seq, float *metric_out) {
  int nsc = 12 * cell->nof_prb;
  int32_t k[12];
  uint8_t c[12];
  lteo_cf_t d[12];
  lteo_phich_res(cell, ng_x6, n_group, k);
  lteo_gold(((uint32_t)(sf_idx + 1) * (uint32_t)(2 * cell->cell_id + 1) << 9) + (uint32_t)cell->cell_id, 12, c);
  if (cell->nof_ports >= 2) {
    const float sq2 = (float)sqrt(2.0);
    for (int i = 0; i < 12; i += 2) {
      const int par = i / 4 + (cell->cp ? n_group / 2 : n_group);
      const lteo_cf_t *ce0 = ce + (size_t)LTEO_DIV_PA(cell->nof_ports, par) * 14 * nsc, *ce1 = ce + (size_t)LTEO_DIV_PB(cell->nof_ports, par) * 14 * nsc;
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
    for (int i = 0; i < 12; i++) {
      lteo_cf_t y = sf[k[i]], h = ce[k[i]];
      float den = (h.re * h.re + h.im * h.im) + n0;
      d[i].re = (y.re * h.re + y.im * h.im) / den;
      d[i].im = (y.im * h.re - y.re * h.im) / den;
    }
  }
  float metric = 0.0f;
  int first = 1;
  for (int i = 0; i < 12; i++) {
    int wr, wi, ci = i;
    if (cell->cp) {
      /* the group's half of every quadruplet: element q = 2 (i / 4) + i % 2 of the six spread symbols */
      if (((i >> 1) & 1) != (n_group & 1)) continue;
      ci = 2 * (i / 4) + (i & 1);
      phich_w2(n_seq, i & 1, &wr, &wi);
    } else
    phich_w(n_seq, i % 4, &wr, &wi);
    /* d * conj(w): w = +-1 -> +-d;  w = +-j -> d * (-+j) = (+-d.im, -+d.re) */
    float tr, ti;
    if (wi == 0) { tr = wr > 0 ? d[i].re : -d[i].re; ti = wr > 0 ? d[i].im : -d[i].im; }
    else { tr = wi > 0 ? d[i].im : -d[i].im; ti = wi > 0 ? -d[i].re : d[i].re; }
    if (c[ci]) { tr = -tr; ti = -ti; }
    float m = tr + ti;
    metric = first ? m : metric + m;
    first = 0;
  }
  if (metric_out) *metric_out = metric;
  return metric < 0.0f;
}

/* ------------------------------------------------------------------------------------------------
 * PBCH / MIB (36.211 6.6, 36.212 5.3.1; srslte_ue_mib_decode, phch_recv.cc:247).  SPEC.md 12.
 * ---------------------------------------------------------------------------------------------- */
/* grid indices (l * nsc + k) of the 240 PBCH resource elements of subframe 0: slot 1, symbols 0..3, the 72 central
 * subcarriers, k first then l; the CRS positions of antenna ports 0..3 are always left out (symbols 0 and 1) */
void lteo_pbch_res(const lteo_cell_t *cell, int32_t *g240) { (void)lteo_pbch_res_n(cell, g240); }

/* the same for either cyclic prefix; returns the number of resource elements: 240, or 216 with the extended prefix, where
 * symbol 3 of the slot carries CRS as well (36.211 6.6.4: the reference signals of ports 0..3 are always left out) and
 * the coded block is E = 1728 bits, 432 per radio frame */
int lteo_pbch_res_n(const lteo_cell_t *cell, int32_t *g240) {
  int nsc = 12 * cell->nof_prb, k0 = nsc / 2 - 36, n = 0, nslot = LTEO_NSLOT(cell->cp);
  for (int l = 0; l < 4; l++)
    for (int k = 0; k < 72; k++) {
      if ((l < 2 || (cell->cp && l == 3)) && (k0 + k) % 3 == cell->cell_id % 3) continue;
      g240[n++] = (nslot + l) * nsc + k0 + k;
    }
  return n;
}

/* the 16-bit CRC mask that signals the number of transmit antenna ports */
static uint32_t pbch_crc_mask(int nof_ports) { return nof_ports == 1 ? 0x0000u : nof_ports == 2 ? 0xFFFFu : 0x5555u; }

/* adds the PBCH of radio frame (sfn mod 4) = frame_idx to the grid of a subframe 0 */
void lteo_pbch_tx(const lteo_cell_t *cell, const uint8_t *mib24, int frame_idx, lteo_cd_t *grid) {
  int nsc = 12 * cell->nof_prb;
  uint8_t c[40], d[120], e[1920], scr[1920];
  int32_t seq[120], g[240];
  memcpy(c, mib24, 24);
  uint32_t crc = lteo_crc_bits(mib24, 24, LTEO_CRC16, 16) ^ pbch_crc_mask(cell->nof_ports);
  for (int i = 0; i < 16; i++) c[24 + i] = (uint8_t)((crc >> (15 - i)) & 1);
  lteo_conv_encode(c, 40, d);
  lteo_cc_rm_sequence(40, seq);
  const int nre = lteo_pbch_res_n(cell, g), nb = 2 * nre;      /* 480 bits per radio frame, 432 with the extended prefix */
  for (int k = 0; k < 4 * nb; k++) e[k] = d[seq[k % 120]];
  lteo_gold((uint32_t)cell->cell_id, 4 * nb, scr);
  double a = 1.0 / sqrt(2.0);
  lteo_cd_t *g0 = grid, *g1 = grid + 14 * nsc;
  for (int i = 0; i < nre; i += 2) {
    lteo_cd_t x[2];
    for (int j = 0; j < 2; j++) {
      int b0 = e[nb * frame_idx + 2 * (i + j)] ^ scr[nb * frame_idx + 2 * (i + j)];
      int b1 = e[nb * frame_idx + 2 * (i + j) + 1] ^ scr[nb * frame_idx + 2 * (i + j) + 1];
      x[j].re = (b0 ? -a : a); x[j].im = (b1 ? -a : a);
    }
    if (cell->nof_ports >= 2) {
      g0 = grid + (size_t)LTEO_DIV_PA(cell->nof_ports, i / 2) * 14 * nsc; g1 = grid + (size_t)LTEO_DIV_PB(cell->nof_ports, i / 2) * 14 * nsc;
      g0[g[i]].re = x[0].re * a;      g0[g[i]].im = x[0].im * a;
      g1[g[i]].re = -x[1].re * a;     g1[g[i]].im = x[1].im * a;
      g0[g[i + 1]].re = x[1].re * a;  g0[g[i + 1]].im = x[1].im * a;
      g1[g[i + 1]].re = x[0].re * a;  g1[g[i + 1]].im = -x[0].im * a;
    } else {
      g0[g[i]] = x[0]; g0[g[i + 1]] = x[1];
    }
  }
}

/* LLRs of the 480 PBCH bits of one subframe 0 under the hypothesis of `hyp_ports` transmit ports (1: single-port
 * equaliser with port 0's estimate; 2: Alamouti combiner), SPEC.md 4-5 arithmetic, not yet descrambled */
void lteo_pbch_llr(const lteo_cell_t *cell, int hyp_ports, const lteo_cf_t *sf, const lteo_cf_t *ce, float n0, int16_t *llr480) {
  int nsc = 12 * cell->nof_prb;
  int32_t g[240];
  lteo_cf_t d[240];
  const int nre = lteo_pbch_res_n(cell, g);
  if (hyp_ports >= 2) {
    const float sq2 = (float)sqrt(2.0);
    for (int i = 0; i < nre; i += 2) {
      const lteo_cf_t *ce0 = ce + (size_t)LTEO_DIV_PA(hyp_ports, i / 2) * 14 * nsc, *ce1 = ce + (size_t)LTEO_DIV_PB(hyp_ports, i / 2) * 14 * nsc;
      lteo_cf_t r0 = sf[g[i]], r1 = sf[g[i + 1]], h0 = ce0[g[i]], h1 = ce1[g[i]];
      float den = ((h0.re * h0.re + h0.im * h0.im) + (h1.re * h1.re + h1.im * h1.im)) + n0;
      float a_re = h0.re * r0.re + h0.im * r0.im, a_im = h0.re * r0.im - h0.im * r0.re;
      float b_re = h1.re * r1.re + h1.im * r1.im, b_im = h1.im * r1.re - h1.re * r1.im;
      float c_re = h0.re * r1.re + h0.im * r1.im, c_im = h0.re * r1.im - h0.im * r1.re;
      float e_re = h1.re * r0.re + h1.im * r0.im, e_im = h1.im * r0.re - h1.re * r0.im;
      d[i].re = ((a_re + b_re) * sq2) / den;     d[i].im = ((a_im + b_im) * sq2) / den;
      d[i + 1].re = ((c_re - e_re) * sq2) / den; d[i + 1].im = ((c_im - e_im) * sq2) / den;
    }
  } else {
    for (int i = 0; i < nre; i++) {
      lteo_cf_t y = sf[g[i]], h = ce[g[i]];
      float den = (h.re * h.re + h.im * h.im) + n0;
      d[i].re = (y.re * h.re + y.im * h.im) / den;
      d[i].im = (y.im * h.re - y.re * h.im) / den;
    }
  }
  lteo_demod(d, nre, 2, llr480);
}

/* Blind MIB decode from one subframe 0: port hypotheses 1 then 2 (2 only when the estimator ran with two ports), for
 * each the four positions of the frame inside the 40 ms BCH period: descramble with that quarter of the sequence,
 * exact int32 soft combining of the 480 LLRs onto the 120 coded bits, the tail-biting Viterbi decoder of the PDCCH
 * (D = 40), CRC16 xor received CRC == antenna mask.  Returns 1 and fills mib24 / nof_ports / sfn_offset for the first
 * match, else 0. */
int lteo_pbch_decode(const lteo_cell_t *cell, const lteo_cf_t *sf, const lteo_cf_t *ce, float n0, uint8_t *mib24,
                     int *nof_ports, int *sfn_offset) {
  uint8_t scr[1920];
  const int nb = cell->cp ? 432 : 480;
  lteo_gold((uint32_t)cell->cell_id, 4 * nb, scr);
  for (int hyp = 1; hyp <= cell->nof_ports; hyp *= 2) {      /* 1, 2, 4 transmit ports, as far as the estimator ran */
    int16_t llr[480], des[480];
    lteo_pbch_llr(cell, hyp, sf, ce, n0, llr);
    for (int q = 0; q < 4; q++) {
      for (int k = 0; k < nb; k++) des[k] = scr[nb * q + k] ? (int16_t)-llr[k] : llr[k];
      /* the PDCCH candidate decoder with E = 480 = 4 x 120: L such that 72 L = 480 does not exist, so call the pieces */
      int32_t seq[120], soft[120];
      lteo_cc_rm_sequence(40, seq);
      memset(soft, 0, sizeof(soft));
      for (int k = 0; k < nb; k++) soft[seq[(nb * q + k) % 120]] += des[k];     /* the frame's place in the 4 nb-bit block */
      uint8_t bits[40];
      uint16_t rem = lteo_viterbi_crc16(soft, 24, bits);
      if (rem == pbch_crc_mask(hyp)) {
        memcpy(mib24, bits, 24);
        if (nof_ports) *nof_ports = hyp;
        if (sfn_offset) *sfn_offset = q;
        return 1;
      }
    }
  }
  return 0;
}

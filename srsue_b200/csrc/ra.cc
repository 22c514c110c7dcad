// Downlink resource-allocation host logic: DCI format 1A / 1 payload -> srslte_ra_dl_dci_t -> srslte_ra_dl_grant_t.
//
// Replaces, for ue/src/phy/phch_worker.cc:297 (decode_pdcch_dl), srsLTE's srslte_dci_msg_to_dl_grant and the helpers it
// calls.  Written from 36.212 5.3.3.1.2/5.3.3.1.3 (field order), 36.213 7.1.6.1-7.1.6.3 (allocation types 0/1/2),
// 7.1.7.1 (MCS -> modulation, I_TBS) and 7.1.7.2.1 (transport block size).  srsLTE itself is an un-vendored
// dependency of the reference, so parity is anchored on the reference's call site and on the packer in tests/.
//
// The 27 x 110 transport-block-size table (36.213 Table 7.1.7.2.1-1) is published data; this tree carries the thirteen
// columns it could write down and check structurally (tbs_table.inc), the caller installs the whole table once with
// srsue_gpu_ra_set_tbs_table() (srsLTE keeps it as tbs_table[27][110] in lib/phch/tbs_tables.h), which takes
// precedence.  A size that neither source has fails loudly.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "srsue_gpu/srslte_compat.h"

namespace {

#include "tbs_table.inc"

std::mutex g_tbs_mu;
std::vector<int32_t> g_tbs;                 // 27 rows (I_TBS) x 110 columns (N_PRB - 1)
std::atomic<bool> g_tbs_set{false};

uint32_t take(uint8_t** y, int n) { uint32_t v = 0; for (int i = 0; i < n; i++) v = (v << 1) | (*(*y)++ & 1u); return v; }
int ceil_log2(uint32_t v) { int b = 0; while ((1u << b) < v) b++; return b; }
uint32_t rbg_size(uint32_t nof_prb) { return nof_prb <= 10 ? 1 : nof_prb <= 26 ? 2 : nof_prb <= 63 ? 3 : 4; }
bool is_crnti(uint16_t r) { return r >= SRSLTE_CRNTI_START && r <= SRSLTE_CRNTI_END; }

// ---- distributed virtual resource blocks (36.211 6.2.3.2) -------------------------------------------------------------
uint32_t n_gap(uint32_t nof_prb, bool gap2) {            // Table 6.2.3.2-1
  if (nof_prb <= 10) return (nof_prb + 1) / 2;
  if (nof_prb == 11) return 4;
  if (nof_prb <= 19) return 8;
  if (nof_prb <= 26) return 12;
  if (nof_prb <= 44) return 18;
  if (nof_prb <= 49) return 27;
  if (nof_prb <= 63) return gap2 ? 9 : 27;
  if (nof_prb <= 79) return gap2 ? 16 : 32;
  return gap2 ? 16 : 48;
}
uint32_t n_vrb_dl(uint32_t nof_prb, bool gap2) {
  const uint32_t g = n_gap(nof_prb, gap2);
  return gap2 ? (nof_prb / (2 * g)) * 2 * g : 2 * std::min(g, nof_prb - g);
}
// physical resource block of distributed virtual resource block n_vrb in slot 0 / 1
uint32_t dvrb_to_prb(uint32_t nof_prb, bool gap2, uint32_t n_vrb, int slot) {
  const uint32_t g = n_gap(nof_prb, gap2), P = rbg_size(nof_prb);
  const uint32_t nt = gap2 ? 2 * g : n_vrb_dl(nof_prb, false);            // size of one interleaving unit
  const uint32_t n_row = (nt + 4 * P - 1) / (4 * P) * P, n_null = 4 * n_row - nt;
  const uint32_t unit = n_vrb / nt, v = n_vrb % nt;
  const uint32_t p1 = 2 * n_row * (v % 2) + v / 2, p2 = n_row * (v % 4) + v / 4;
  uint32_t t;
  if (n_null && v >= nt - n_null && v % 2 == 1) t = p1 - n_row;
  else if (n_null && v >= nt - n_null && v % 2 == 0) t = p1 - n_row + n_null / 2;
  else if (n_null && v < nt - n_null && v % 4 >= 2) t = p2 - n_null / 2;
  else t = p2;
  if (slot) t = (t + nt / 2) % nt;
  t += nt * unit;
  return t < nt / 2 ? t : t + g - nt / 2;
}

}  // namespace

extern "C" {

uint32_t srsue_gpu_host_dvrb_to_prb(uint32_t nof_prb, int gap2, uint32_t n_vrb, int slot) {
  if (nof_prb < 6 || nof_prb > SRSLTE_MAX_PRB || (gap2 && nof_prb < 50) || n_vrb >= n_vrb_dl(nof_prb, gap2 != 0)) return 0xFFFFFFFFu;
  return dvrb_to_prb(nof_prb, gap2 != 0, n_vrb, slot ? 1 : 0);
}
uint32_t srsue_gpu_host_n_vrb_dl(uint32_t nof_prb, int gap2) {
  return (nof_prb < 6 || nof_prb > SRSLTE_MAX_PRB || (gap2 && nof_prb < 50)) ? 0 : n_vrb_dl(nof_prb, gap2 != 0);
}


int srsue_gpu_ra_set_tbs_table(const int32_t* table, uint32_t nof_rows, uint32_t nof_cols) {
  if (!table || nof_rows != 27 || nof_cols != 110) return SRSLTE_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(g_tbs_mu);
  g_tbs.assign(table, table + 27 * 110);
  g_tbs_set.store(true);
  return SRSLTE_SUCCESS;
}

int srsue_gpu_ra_have_tbs_table(void) { return g_tbs_set.load() ? 1 : 0; }

int srsue_gpu_ra_builtin_tbs_columns(int32_t* n_prb, int cap) {
  const int n = (int)(sizeof(kTbsBuiltinCols) / sizeof(kTbsBuiltinCols[0]));
  for (int i = 0; i < n && i < cap && n_prb; i++) n_prb[i] = kTbsBuiltinCols[i];
  return n;
}

uint32_t srslte_ra_type0_P(uint32_t nof_prb) { return rbg_size(nof_prb); }

uint32_t srslte_ra_type2_n_rb(uint32_t nof_prb) { return (uint32_t)ceil_log2(nof_prb * (nof_prb + 1) / 2); }

uint32_t srslte_ra_type2_to_riv(uint32_t L_crb, uint32_t RB_start, uint32_t nof_prb) {
  return (L_crb - 1 <= nof_prb / 2) ? nof_prb * (L_crb - 1) + RB_start : nof_prb * (nof_prb - L_crb + 1) + nof_prb - 1 - RB_start;
}

void srslte_ra_type2_from_riv(uint32_t riv, uint32_t* L_crb, uint32_t* RB_start, uint32_t nof_prb, uint32_t nof_vrb) {
  uint32_t L = riv / nof_prb + 1, s = riv % nof_prb;
  if (s >= nof_vrb || L > nof_vrb - s) { L = nof_prb - L + 2; s = nof_prb - 1 - s; }   // the folded half of 36.213 7.1.6.3
  if (L_crb) *L_crb = L;
  if (RB_start) *RB_start = s;
}

// 36.213 Table 7.1.7.1-1
int srslte_ra_tbs_idx_from_mcs(uint32_t mcs_idx) { return mcs_idx < 10 ? (int)mcs_idx : mcs_idx < 17 ? (int)mcs_idx - 1 : mcs_idx < 29 ? (int)mcs_idx - 2 : SRSLTE_ERROR; }

srslte_mod_t srslte_ra_mod_from_mcs(uint32_t mcs_idx) {
  return (mcs_idx < 10 || mcs_idx == 29) ? SRSLTE_MOD_QPSK : (mcs_idx < 17 || mcs_idx == 30) ? SRSLTE_MOD_16QAM : SRSLTE_MOD_64QAM;
}

int srslte_ra_tbs_from_idx(uint32_t tbs_idx, uint32_t n_prb) {
  if (tbs_idx >= 27 || n_prb < 1 || n_prb > 110) return SRSLTE_ERROR;
  if (!g_tbs_set.load()) {
    for (size_t c = 0; c < sizeof(kTbsBuiltinCols) / sizeof(kTbsBuiltinCols[0]); c++)
      if (kTbsBuiltinCols[c] == (int)n_prb) return kTbsBuiltin[c][tbs_idx];
    fprintf(stderr, "[srsue_gpu] srslte_ra_tbs_from_idx: N_PRB = %u is not among the built-in columns and no transport-block-size table "
                    "is installed (srsue_gpu_ra_set_tbs_table)\n", n_prb);
    return SRSLTE_ERROR;
  }
  std::lock_guard<std::mutex> lk(g_tbs_mu);
  return g_tbs[tbs_idx * 110 + (n_prb - 1)];
}

// 36.212 5.3.3.1.3 (format 1A) and 5.3.3.1.2 (format 1), FDD field widths
int srslte_dci_msg_unpack_pdsch(srslte_dci_msg_t* msg, srslte_ra_dl_dci_t* d, uint32_t nof_prb, bool crc_is_crnti) {
  if (!msg || !d || nof_prb < 6 || nof_prb > SRSLTE_MAX_PRB) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(d, 0, sizeof(*d));
  uint8_t* y = msg->data;
  if (msg->format == SRSLTE_DCI_FORMAT1A) {
    const uint32_t need = 15 + srslte_ra_type2_n_rb(nof_prb);
    if (msg->nof_bits < need) return SRSLTE_ERROR;
    if (take(&y, 1) != 1) return SRSLTE_ERROR;                          // 0: this is a format 0 (uplink) message
    d->dci_is_1a = true;
    d->alloc_type = SRSLTE_RA_ALLOC_TYPE2;
    d->type2_alloc.mode = take(&y, 1) ? SRSLTE_RA_TYPE2_DIST : SRSLTE_RA_TYPE2_LOC;
    int riv_bits = (int)srslte_ra_type2_n_rb(nof_prb);
    if (d->type2_alloc.mode == SRSLTE_RA_TYPE2_DIST && crc_is_crnti && nof_prb >= 50) {     // MSB of the field: the gap
      d->type2_alloc.n_gap = take(&y, 1) ? SRSLTE_RA_TYPE2_NG2 : SRSLTE_RA_TYPE2_NG1;
      riv_bits--;
    }
    d->type2_alloc.riv = take(&y, riv_bits);
    d->mcs_idx = take(&y, 5);
    d->harq_process = take(&y, 3);
    const uint32_t ndi = take(&y, 1);
    if (crc_is_crnti) d->ndi = ndi != 0;
    else if (nof_prb >= 50 && d->type2_alloc.mode == SRSLTE_RA_TYPE2_DIST) d->type2_alloc.n_gap = ndi ? SRSLTE_RA_TYPE2_NG2 : SRSLTE_RA_TYPE2_NG1;
    d->rv_idx = (int)take(&y, 2);
    const uint32_t tpc = take(&y, 2);
    d->tpc = tpc;
    if (!crc_is_crnti) d->type2_alloc.n_prb1a = (tpc & 1) ? SRSLTE_RA_TYPE2_NPRB1A_3 : SRSLTE_RA_TYPE2_NPRB1A_2;
    const uint32_t nof_vrb = d->type2_alloc.mode == SRSLTE_RA_TYPE2_LOC ? nof_prb : n_vrb_dl(nof_prb, d->type2_alloc.n_gap == SRSLTE_RA_TYPE2_NG2);
    srslte_ra_type2_from_riv(d->type2_alloc.riv, &d->type2_alloc.L_crb, &d->type2_alloc.RB_start, nof_prb, nof_vrb);
    if (d->type2_alloc.L_crb < 1 || d->type2_alloc.RB_start + d->type2_alloc.L_crb > nof_vrb) return SRSLTE_ERROR;
    return SRSLTE_SUCCESS;
  }
  if (msg->format == SRSLTE_DCI_FORMAT1) {
    const uint32_t P = rbg_size(nof_prb), nbm = (nof_prb + P - 1) / P;
    if (msg->nof_bits < (nof_prb > 10 ? 1u : 0u) + nbm + 13) return SRSLTE_ERROR;
    d->alloc_type = (nof_prb > 10 && take(&y, 1)) ? SRSLTE_RA_ALLOC_TYPE1 : SRSLTE_RA_ALLOC_TYPE0;
    if (d->alloc_type == SRSLTE_RA_ALLOC_TYPE0) {
      d->type0_alloc.rbg_bitmask = take(&y, (int)nbm);
    } else {
      const int nsub = ceil_log2(P);
      d->type1_alloc.rbg_subset = take(&y, nsub);
      d->type1_alloc.shift = take(&y, 1) != 0;
      d->type1_alloc.vrb_bitmask = take(&y, (int)nbm - nsub - 1);
    }
    d->mcs_idx = take(&y, 5);
    d->harq_process = take(&y, 3);
    d->ndi = take(&y, 1) != 0;
    d->rv_idx = (int)take(&y, 2);
    d->tpc = take(&y, 2);
    return SRSLTE_SUCCESS;
  }
  return SRSLTE_ERROR;
}

// 36.213 7.1.6: the physical resource blocks of both slots
int srslte_ra_dl_dci_to_grant_prb_allocation(srslte_ra_dl_dci_t* d, srslte_ra_dl_grant_t* grant, uint32_t nof_prb) {
  if (!d || !grant || nof_prb < 6 || nof_prb > SRSLTE_MAX_PRB) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(grant->prb_idx, 0, sizeof(grant->prb_idx));
  grant->nof_prb = 0;
  const uint32_t P = rbg_size(nof_prb);
  switch (d->alloc_type) {
    case SRSLTE_RA_ALLOC_TYPE0: {
      const uint32_t nbm = (nof_prb + P - 1) / P;
      for (uint32_t i = 0; i < nbm; i++) {
        if (!((d->type0_alloc.rbg_bitmask >> (nbm - 1 - i)) & 1u)) continue;
        for (uint32_t j = 0; j < P && i * P + j < nof_prb; j++) { grant->prb_idx[0][i * P + j] = true; grant->nof_prb++; }
      }
      break;
    }
    case SRSLTE_RA_ALLOC_TYPE1: {
      const uint32_t nbm = (nof_prb + P - 1) / P, n_type1 = nbm - (uint32_t)ceil_log2(P) - 1, p = d->type1_alloc.rbg_subset;
      if (p >= P) return SRSLTE_ERROR;
      // size of RBG subset p (36.213 7.1.6.2)
      const uint32_t q = (nof_prb - 1) / P, full = (nof_prb - 1) / (P * P) * P;
      const uint32_t n_subset = p < q % P ? full + P : p == q % P ? full + (nof_prb - 1) % P + 1 : full;
      if (n_subset < n_type1) return SRSLTE_ERROR;
      const uint32_t shift = d->type1_alloc.shift ? n_subset - n_type1 : 0;
      for (uint32_t i = 0; i < n_type1; i++) {
        if (!((d->type1_alloc.vrb_bitmask >> (n_type1 - 1 - i)) & 1u)) continue;
        const uint32_t v = i + shift, n = (v / P) * P * P + p * P + v % P;
        if (n >= nof_prb) return SRSLTE_ERROR;
        grant->prb_idx[0][n] = true;
        grant->nof_prb++;
      }
      break;
    }
    case SRSLTE_RA_ALLOC_TYPE2: {
      if (d->type2_alloc.mode == SRSLTE_RA_TYPE2_DIST) {
        const bool gap2 = d->type2_alloc.n_gap == SRSLTE_RA_TYPE2_NG2;
        if ((gap2 && nof_prb < 50) || d->type2_alloc.RB_start + d->type2_alloc.L_crb > n_vrb_dl(nof_prb, gap2)) return SRSLTE_ERROR;
        for (uint32_t i = 0; i < d->type2_alloc.L_crb; i++)
          for (int slot = 0; slot < 2; slot++) {
            const uint32_t n = dvrb_to_prb(nof_prb, gap2, d->type2_alloc.RB_start + i, slot);
            if (n >= nof_prb) return SRSLTE_ERROR;
            grant->prb_idx[slot][n] = true;
          }
        grant->nof_prb = d->type2_alloc.L_crb;
        return SRSLTE_SUCCESS;                            // the two slots differ: no copy below
      }
      if (d->type2_alloc.RB_start + d->type2_alloc.L_crb > nof_prb) return SRSLTE_ERROR;
      for (uint32_t i = 0; i < d->type2_alloc.L_crb; i++) grant->prb_idx[0][d->type2_alloc.RB_start + i] = true;
      grant->nof_prb = d->type2_alloc.L_crb;
      break;
    }
    default: return SRSLTE_ERROR;
  }
  std::memcpy(grant->prb_idx[1], grant->prb_idx[0], sizeof(grant->prb_idx[0]));
  return grant->nof_prb ? SRSLTE_SUCCESS : SRSLTE_ERROR;
}

int srslte_ra_dl_dci_to_grant(srslte_ra_dl_dci_t* d, uint32_t nof_prb, bool crc_is_crnti, srslte_ra_dl_grant_t* grant) {
  int rc = srslte_ra_dl_dci_to_grant_prb_allocation(d, grant, nof_prb);
  if (rc) return rc;
  grant->mcs.idx = d->mcs_idx;
  if (!crc_is_crnti) {
    // SI-/P-/RA-RNTI on format 1A: QPSK, I_TBS = MCS, column N_PRB^1A in {2, 3} (36.213 7.1.7, 7.1.7.2.1)
    if (!d->dci_is_1a || d->mcs_idx > 26) return SRSLTE_ERROR;
    grant->mcs.mod = SRSLTE_MOD_QPSK;
    grant->mcs.tbs = srslte_ra_tbs_from_idx(d->mcs_idx, d->type2_alloc.n_prb1a == SRSLTE_RA_TYPE2_NPRB1A_2 ? 2 : 3);
  } else {
    grant->mcs.mod = srslte_ra_mod_from_mcs(d->mcs_idx);
    if (d->mcs_idx >= 29) grant->mcs.tbs = 0;                                   // retransmission: size of the first transmission (MAC keeps it)
    else grant->mcs.tbs = srslte_ra_tbs_from_idx((uint32_t)srslte_ra_tbs_idx_from_mcs(d->mcs_idx), grant->nof_prb);
  }
  grant->Qm = grant->mcs.mod == SRSLTE_MOD_QPSK ? 2 : grant->mcs.mod == SRSLTE_MOD_16QAM ? 4 : 6;
  return grant->mcs.tbs < 0 ? SRSLTE_ERROR : SRSLTE_SUCCESS;
}

int srslte_dci_msg_to_dl_grant(srslte_dci_msg_t* msg, uint16_t msg_rnti, uint32_t nof_prb, srslte_ra_dl_dci_t* dl_dci, srslte_ra_dl_grant_t* grant) {
  if (!msg || !dl_dci || !grant) return SRSLTE_ERROR_INVALID_INPUTS;
  const bool crnti = is_crnti(msg_rnti);
  int rc = srslte_dci_msg_unpack_pdsch(msg, dl_dci, nof_prb, crnti);
  if (rc) return rc;
  return srslte_ra_dl_dci_to_grant(dl_dci, nof_prb, crnti, grant);
}

// inverse of the unpacker (what an eNodeB-side test bench needs); returns the number of bits written or < 0
int srslte_dci_msg_pack_pdsch(srslte_ra_dl_dci_t* d, srslte_dci_format_t format, srslte_dci_msg_t* msg, uint32_t nof_prb, bool crc_is_crnti) {
  if (!d || !msg || nof_prb < 6 || nof_prb > SRSLTE_MAX_PRB) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(msg, 0, sizeof(*msg));
  uint8_t* y = msg->data;
  auto put = [&](uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) *y++ = (uint8_t)((v >> i) & 1u); };
  if (format == SRSLTE_DCI_FORMAT1A) {
    if (d->alloc_type != SRSLTE_RA_ALLOC_TYPE2) return SRSLTE_ERROR;
    put(1, 1);
    const bool dist = d->type2_alloc.mode == SRSLTE_RA_TYPE2_DIST;
    put(dist, 1);
    int riv_bits = (int)srslte_ra_type2_n_rb(nof_prb);
    if (dist && crc_is_crnti && nof_prb >= 50) { put(d->type2_alloc.n_gap == SRSLTE_RA_TYPE2_NG2, 1); riv_bits--; }
    const uint32_t riv = srslte_ra_type2_to_riv(d->type2_alloc.L_crb, d->type2_alloc.RB_start, nof_prb);
    if (riv >> riv_bits) return SRSLTE_ERROR;
    put(riv, riv_bits);
    put(d->mcs_idx, 5);
    put(d->harq_process, 3);
    put(crc_is_crnti ? d->ndi : (d->type2_alloc.n_gap == SRSLTE_RA_TYPE2_NG2), 1);
    put((uint32_t)d->rv_idx, 2);
    put(crc_is_crnti ? d->tpc : (d->type2_alloc.n_prb1a == SRSLTE_RA_TYPE2_NPRB1A_3 ? 1u : 0u), 2);
  } else if (format == SRSLTE_DCI_FORMAT1) {
    const uint32_t P = rbg_size(nof_prb), nbm = (nof_prb + P - 1) / P;
    if (d->alloc_type == SRSLTE_RA_ALLOC_TYPE2 || (nof_prb <= 10 && d->alloc_type != SRSLTE_RA_ALLOC_TYPE0)) return SRSLTE_ERROR;
    if (nof_prb > 10) put(d->alloc_type == SRSLTE_RA_ALLOC_TYPE1, 1);
    if (d->alloc_type == SRSLTE_RA_ALLOC_TYPE0) put(d->type0_alloc.rbg_bitmask, (int)nbm);
    else {
      const int nsub = ceil_log2(P);
      put(d->type1_alloc.rbg_subset, nsub);
      put(d->type1_alloc.shift, 1);
      put(d->type1_alloc.vrb_bitmask, (int)nbm - nsub - 1);
    }
    put(d->mcs_idx, 5);
    put(d->harq_process, 3);
    put(d->ndi, 1);
    put((uint32_t)d->rv_idx, 2);
    put(d->tpc, 2);
  } else return SRSLTE_ERROR;
  const int nbits = srsue_gpu_host_dci_format_sizeof(format == SRSLTE_DCI_FORMAT1A ? 0 : 1, (int)nof_prb);   // padding (zeros) included
  if (nbits <= 0 || (y - msg->data) > nbits) return SRSLTE_ERROR;
  msg->nof_bits = (uint32_t)nbits;
  msg->format = format;
  return nbits;
}

// ---- CQI reporting helpers (phch_worker.cc:504-527) ------------------------------------------------------------------
// Wideband CQI from the channel estimator's SNR in dB.  [UPSTREAM-RECALL] srsLTE's lib/phch/cqi.c keeps the SNR at which
// each CQI index becomes usable (1.95 dB for CQI 1 ... 29 dB for CQI 15) and reports the number of entries below the
// measured SNR; the table is recalled, not readable offline (srsLTE is not vendored).  At the reference's call site
// (phch_worker.cc:504-527) this is the value the eNodeB's link adaptation sees, so the thresholds must not be optimistic:
// an earlier ladder here (-6.5 ... 21 dB) reported several CQI steps more than srsLTE for the same SNR.
uint8_t srslte_cqi_from_snr(float snr_db) {
  static const float cqi_to_snr[15] = {1.95f, 4.f, 6.f, 8.f, 10.f, 11.95f, 14.05f, 16.f, 17.9f, 19.9f, 21.5f, 23.45f, 25.0f, 27.30f, 29.f};
  int idx = 0;
  while (idx < 15 && cqi_to_snr[idx] < snr_db) idx++;
  return (uint8_t)idx;
}

// 36.213 Table 7.2.2-1A (FDD): does a periodic CQI/PMI report with configuration index I_cqi_pmi fall on this tti?
bool srslte_cqi_send(uint32_t I_cqi_pmi, uint32_t tti) {
  static const struct { uint32_t first, last, period; } row[] = {
      {0, 1, 2}, {2, 6, 5}, {7, 16, 10}, {17, 36, 20}, {37, 76, 40}, {77, 156, 80}, {157, 316, 160},
      {318, 349, 32}, {350, 413, 64}, {414, 541, 128}};
  for (const auto& r : row)
    if (I_cqi_pmi >= r.first && I_cqi_pmi <= r.last) return (tti + r.period - (I_cqi_pmi - r.first) % r.period) % r.period == 0;
  return false;                                                            // 317 and 542..1023 are reserved
}

// wideband: 4 bits; UE-selected subband (PUCCH format 2, one bandwidth part): 4 bits + 1 label bit.  MSB first.
int srslte_cqi_value_pack(srslte_cqi_value_t* value, uint8_t* buff) {
  if (!value || !buff) return SRSLTE_ERROR_INVALID_INPUTS;
  uint8_t* y = buff;
  auto put = [&](uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) *y++ = (uint8_t)((v >> i) & 1u); };
  switch (value->type) {
    case SRSLTE_CQI_TYPE_WIDEBAND: put(value->wideband.wideband_cqi, 4); break;
    case SRSLTE_CQI_TYPE_SUBBAND: put(value->subband.subband_cqi, 4); put(value->subband.subband_label, 1); break;
    default: return SRSLTE_ERROR;
  }
  return (int)(y - buff);
}

char* srslte_ra_dl_dci_string(srslte_ra_dl_dci_t* d) {
  static thread_local char s[160];
  if (!d) { s[0] = 0; return s; }
  const char* t = d->alloc_type == SRSLTE_RA_ALLOC_TYPE0 ? "type0" : d->alloc_type == SRSLTE_RA_ALLOC_TYPE1 ? "type1" : "type2";
  if (d->alloc_type == SRSLTE_RA_ALLOC_TYPE2)
    snprintf(s, sizeof(s), "%s rb_start=%u l_crb=%u mcs=%u harq_pid=%u rv=%d ndi=%d", t, d->type2_alloc.RB_start, d->type2_alloc.L_crb,
             d->mcs_idx, d->harq_process, d->rv_idx, (int)d->ndi);
  else
    snprintf(s, sizeof(s), "%s mask=0x%x mcs=%u harq_pid=%u rv=%d ndi=%d", t,
             d->alloc_type == SRSLTE_RA_ALLOC_TYPE0 ? d->type0_alloc.rbg_bitmask : d->type1_alloc.vrb_bitmask, d->mcs_idx, d->harq_process,
             d->rv_idx, (int)d->ndi);
  return s;
}

}  // extern "C"

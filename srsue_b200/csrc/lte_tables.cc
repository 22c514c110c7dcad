// lte_tables.cc -- host-side LTE tables for libsrsue_gpu (see lte_tables.h).
// 3GPP TS 36.211 6.3.5 / 6.10.1 / 7.2 and TS 36.212 5.1.1-5.1.4; window rule per oracle/SPEC.md 7.3.
#include "lte_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace srsue {

static const uint16_t kQpp[188][3] = {
#include "qpp_table.inc"
};

int qpp_index(int K) {
  int lo = 0, hi = 187;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    if (kQpp[mid][0] == K) return mid;
    if (kQpp[mid][0] < K) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

int qpp_K(int idx) { return (idx >= 0 && idx < 188) ? kQpp[idx][0] : -1; }

bool qpp_params(int K, int* f1, int* f2) {
  int i = qpp_index(K);
  if (i < 0) return false;
  *f1 = kQpp[i][1]; *f2 = kQpp[i][2];
  return true;
}

int window_len(int K) {
  int even = 0, any = 0;
  for (int w = 64; w <= 128 && w <= K; w += 8) {
    if (K % w) continue;
    any = w;
    if ((K / w) % 2 == 0) even = w;
  }
  if (even) return even;
  if (any) return any;
  for (int w = 64; w <= K; w += 8)
    if (K % w == 0) return w;
  return K;
}

TurboGeom turbo_geom(int K) {
  TurboGeom g{};
  g.K = K;
  g.W = window_len(K);
  g.P = K / g.W;
  g.Ppad = (g.P + 1) & ~1;
  g.T = g.Ppad / 2;
  g.plane = g.W * g.Ppad;
  g.cb_elems = 3 * g.plane + 16;
  return g;
}

int symbol_sz(int nof_prb) {
  static const int lim[6] = {6, 15, 25, 50, 75, 110}, sz[6] = {128, 256, 512, 1024, 1536, 2048};
  if (nof_prb <= 0) return -1;
  for (int i = 0; i < 6; i++)
    if (nof_prb <= lim[i]) return sz[i];
  return -1;
}

uint32_t crc_bits(const uint8_t* bits, int n, uint32_t poly, int order) {
  const uint32_t top = 1u << order;
  uint32_t r = 0;
  for (int i = 0; i < n + order; i++) {
    r = (r << 1) | (i < n ? (bits[i] & 1u) : 0u);
    if (r & top) r ^= poly;
  }
  return r & (top - 1);
}

static uint32_t gf_mul24(uint32_t a, uint32_t b, uint32_t poly) {
  uint32_t r = 0;
  for (int i = 23; i >= 0; i--) {
    r <<= 1;
    if (r & 0x1000000u) r ^= poly;
    if ((b >> i) & 1u) r ^= a;
  }
  return r & 0xFFFFFFu;
}

uint32_t crc_xpow(uint32_t poly, uint64_t e) {
  uint32_t result = 1, base = 2;    // polynomials "1" and "x"
  while (e) {
    if (e & 1) result = gf_mul24(result, base, poly);
    base = gf_mul24(base, base, poly);
    e >>= 1;
  }
  return result;
}

void gold_bits(uint32_t c_init, int n, uint8_t* c) {
  // two 31-bit LFSRs advanced 1600 steps before the first output (36.211 7.2)
  uint32_t x1 = 1, x2 = c_init & 0x7FFFFFFFu;
  auto step = [&]() {
    uint32_t n1 = ((x1 >> 3) ^ x1) & 1u;
    uint32_t n2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
    x1 = (x1 >> 1) | (n1 << 30);
    x2 = (x2 >> 1) | (n2 << 30);
  };
  for (int i = 0; i < 1600; i++) step();
  for (int i = 0; i < n; i++) {
    c[i] = (uint8_t)((x1 ^ x2) & 1u);
    step();
  }
}

void gold_packed(uint32_t c_init, int n, std::vector<uint32_t>& w) {
  std::vector<uint8_t> c(n);
  gold_bits(c_init, n, c.data());
  w.assign((n + 31) / 32, 0u);
  for (int i = 0; i < n; i++) w[i >> 5] |= (uint32_t)c[i] << (i & 31);
}

bool cbsegm(int tbs, CbSegm* s) {
  std::memset(s, 0, sizeof(*s));
  if (tbs <= 0) return false;
  const int B = tbs + 24;
  int C = 1, Bp = B;
  if (B > kMaxK) { C = (B + 6119) / 6120; Bp = B + 24 * C; }
  const int need = (Bp + C - 1) / C;
  int ip = 0;
  while (ip < 188 && kQpp[ip][0] < need) ip++;
  if (ip == 188) return false;
  s->tbs = tbs; s->B = B; s->C = C; s->Kp = kQpp[ip][0]; s->Cp = C;
  if (C > 1) {
    if (ip == 0) return false;
    s->Km = kQpp[ip - 1][0];
    s->Cm = (C * s->Kp - Bp) / (s->Kp - s->Km);
    s->Cp = C - s->Cm;
  }
  s->F = s->Cp * s->Kp + s->Cm * s->Km - Bp;
  return true;
}

int cb_E(const CbSegm& s, int G, int qm, int nl, int r) {
  const int Gp = G / (nl * qm), gamma = Gp % s.C;
  return nl * qm * (r <= s.C - gamma - 1 ? Gp / s.C : (Gp + s.C - 1) / s.C);
}

int crs_offset(const CellCfg& cell, int port, int l) {
  const int nslot = slot_symb(cell.cp), ls = l % nslot;
  if (port >= 2) {                                      // ports 2/3: symbol 1 of each slot, v = 3 (n_s mod 2) [+ 3 for port 3]
    if (ls != 1) return -1;
    return (3 * (l / nslot) + (port == 3 ? 3 : 0) + cell.cell_id % 6) % 6;
  }
  if (ls != 0 && ls != nslot - 3) return -1;            // ports 0/1: symbols 0 and N_symb - 3 of a slot (36.211 6.10.1.2)
  const int v = ((ls == 0) == (port == 0)) ? 0 : 3;
  return (v + cell.cell_id % 6) % 6;
}

void crs_signs(const CellCfg& cell, int sf_idx, int l, std::vector<int8_t>& re_sign, std::vector<int8_t>& im_sign) {
  const int nslot = slot_symb(cell.cp), ns = 2 * sf_idx + l / nslot, ls = l % nslot, M = 2 * cell.nof_prb;
  // 36.211 6.10.1.1: ... + 2 N_ID + N_CP with N_CP = 1 (normal) or 0 (extended cyclic prefix)
  const uint32_t c_init = 1024u * (7u * (ns + 1) + ls + 1) * (2u * cell.cell_id + 1) + 2u * cell.cell_id + (cell.cp ? 0u : 1u);
  std::vector<uint8_t> c(440);
  gold_bits(c_init, 440, c.data());
  re_sign.resize(M); im_sign.resize(M);
  for (int m = 0; m < M; m++) {
    const int mp = m + 110 - cell.nof_prb;
    re_sign[m] = c[2 * mp] ? -1 : 1;
    im_sign[m] = c[2 * mp + 1] ? -1 : 1;
  }
}

void pdsch_re_list(const CellCfg& cell, const PdschCfg& cfg, std::vector<int32_t>& re) {
  re.clear();
  const int nsc = 12 * cell.nof_prb;
  const int first = cfg.cfi + (cell.nof_prb <= 10 ? 1 : 0);
  const int mid_lo = nsc / 2 - 36, mid_hi = nsc / 2 + 36;
  const int nslot = slot_symb(cell.cp);
  for (int l = first; l < 2 * nslot; l++) {
    int o0 = crs_offset(cell, 0, l);
    int o1 = (cell.nof_ports > 1) ? crs_offset(cell, 1, l) : -1;
    if (cell.nof_ports == 4 && o0 < 0) { o0 = crs_offset(cell, 2, l); o1 = crs_offset(cell, 3, l); }     // symbol 1 of a slot
    // SSS and PSS close slot 0 of subframes 0 and 5, the PBCH opens slot 1 of subframe 0
    bool central_reserved = ((cfg.sf_idx == 0 || cfg.sf_idx == 5) && (l == nslot - 2 || l == nslot - 1)) ||
                            (cfg.sf_idx == 0 && l >= nslot && l <= nslot + 3);
    for (int prb = 0; prb < cell.nof_prb; prb++) {
      if (!prb_in_slot(cfg.prb_mask[prb], l / nslot)) continue;
      for (int k = 12 * prb; k < 12 * prb + 12; k++) {
        if (o0 >= 0 && (k % 6 == o0 || k % 6 == o1)) continue;
        if (central_reserved && k >= mid_lo && k < mid_hi) continue;
        re.push_back(l * nsc + k);
      }
    }
  }
}

void turbo_perm_table(const TurboGeom& g, std::vector<uint16_t>& tab) {
  int f1 = 0, f2 = 0;
  qpp_params(g.K, &f1, &f2);
  tab.assign((size_t)g.W * 2 * g.T, 0);
  for (int j = 0; j < g.Ppad; j++)
    for (int i = 0; i < g.W; i++) {
      const size_t e = ((size_t)i * 2 + (j & 1)) * g.T + j / 2;
      if (j >= g.P) { tab[e] = (uint16_t)(2 * (i * g.Ppad + j)); continue; }     // padding column maps onto itself
      const int64_t k = (int64_t)j * g.W + i;
      const int n = (int)((f1 * k + (int64_t)f2 * k * k) % g.K);
      tab[e] = (uint16_t)(2 * ((n % g.W) * g.Ppad + n / g.W));
    }
}

// For natural bit n of a code block: where its hard decision lies in the decoder's DEC2-order output row
// (turbo.cu: [W/8][T] x u16, low byte window 2t, high byte window 2t+1, step i of a group at bit 7 - i%8), as
// 8 * byte + bit with the bit counted from the LSB.
void turbo_deint_table(const TurboGeom& g, std::vector<uint16_t>& tab) {
  int f1 = 0, f2 = 0;
  qpp_params(g.K, &f1, &f2);
  tab.assign((size_t)g.K, 0);
  for (int j = 0; j < g.P; j++)
    for (int i = 0; i < g.W; i++) {
      const int64_t k = (int64_t)j * g.W + i;
      const int n = (int)((f1 * k + (int64_t)f2 * k * k) % g.K);
      const int byte = ((i / 8) * g.T + j / 2) * 2 + (j & 1);
      tab[n] = (uint16_t)(byte * 8 + (7 - i % 8));
    }
}
void turbo_crc_table(const TurboGeom& g, uint32_t poly, std::vector<uint32_t>& tlin) {
  int f1 = 0, f2 = 0;
  qpp_params(g.K, &f1, &f2);
  std::vector<uint32_t> nat(g.K);                    // contribution of natural position n
  uint32_t v = crc_xpow(poly, 24);                   // last bit, n = K-1
  for (int n = g.K - 1; n >= 0; n--) {
    nat[n] = v;
    v <<= 1;                                         // times x, reduced
    if (v & 0x1000000u) v ^= poly;
  }
  tlin.assign((size_t)g.plane, 0u);
  for (int j = 0; j < g.P; j++)
    for (int i = 0; i < g.W; i++) {
      const int64_t k = (int64_t)j * g.W + i;
      const int n = (int)((f1 * k + (int64_t)f2 * k * k) % g.K);
      tlin[((((size_t)(i / 2) * g.T + j / 2) * 2 + i % 2) * 2) + (j & 1)] = nat[n];
    }
}

int tcb_offset(const TurboGeom& g, int triple_index) {
  const int k = triple_index / 3, stream = triple_index % 3;
  if (k < g.K) {
    const int j = k / g.W, i = k % g.W;       // window, step
    return stream * g.plane + ((((i / 4) * g.T + j / 2) * 4 + i % 4) * 2) + (j & 1);
  }
  return 3 * g.plane + (triple_index - 3 * g.K);      // 12 tail values in srsLTE order
}

int rm_gather_table(const TurboGeom& g, int F, int rv, std::vector<uint16_t>& tab) {
  static const uint8_t colperm[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                      1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
  const int D = g.K + 4, R = (D + 31) / 32, Kpi = 32 * R, ND = Kpi - D, Kw = 3 * Kpi;
  const int k0 = R * (2 * ((Kw + 8 * R - 1) / (8 * R)) * rv + 2);
  tab.assign(g.cb_elems, 0xFFFF);
  for (int k = 0; k < F; k++) { tab[tcb_offset(g, 3 * k)] = 0xFFFE; tab[tcb_offset(g, 3 * k + 1)] = 0xFFFE; }
  int n = 0;
  for (int step = 0; step < Kw; step++) {
    const int j = (k0 + step) % Kw;
    int stream, col_pos;
    if (j < Kpi) { stream = 0; col_pos = j; } else { stream = 1 + ((j - Kpi) & 1); col_pos = (j - Kpi) >> 1; }
    const int row = col_pos % R, col = col_pos / R;
    const int y = (stream < 2) ? row * 32 + colperm[col] : (colperm[col] + 32 * row + 1) % Kpi;
    const int d = y - ND;
    if (d < 0 || (stream < 2 && d < F)) continue;
    tab[tcb_offset(g, 3 * d + stream)] = (uint16_t)n++;
  }
  return n;
}

// Transmit side of the same circular buffer (uplink encoder, ulsch.cu): the order in which the non-<NULL> elements are read
// for redundancy version rv, as (k << 2) | stream with k the index of the triple d(0..2)_k, k < K + 4.
void rm_tx_sequence(int K, int F, int rv, std::vector<uint16_t>& seq) {
  static const uint8_t colperm[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                      1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
  const int D = K + 4, R = (D + 31) / 32, Kpi = 32 * R, ND = Kpi - D, Kw = 3 * Kpi;
  const int k0 = R * (2 * ((Kw + 8 * R - 1) / (8 * R)) * rv + 2);
  seq.clear();
  for (int step = 0; step < Kw; step++) {
    const int j = (k0 + step) % Kw;
    int stream, col_pos;
    if (j < Kpi) { stream = 0; col_pos = j; } else { stream = 1 + ((j - Kpi) & 1); col_pos = (j - Kpi) >> 1; }
    const int row = col_pos % R, col = col_pos / R;
    const int y = (stream < 2) ? row * 32 + colperm[col] : (colperm[col] + 32 * row + 1) % Kpi;
    const int d = y - ND;
    if (d < 0 || (stream < 2 && d < F)) continue;
    seq.push_back((uint16_t)((d << 2) | stream));
  }
}

void fft_twiddles(int n, std::vector<float>& tw) {
  if (n % 3 == 0) {
    // n = 3 m (1536): the table of the three m-point transforms, then the full-circle twiddles w1 = w_n^k and
    // w2 = w_n^2k of the final radix-3 stage, evaluated in double and rounded once (SPEC.md 2)
    const int m = n / 3;
    fft_twiddles(m, tw);
    tw.resize(m + 4 * m);
    for (int k = 0; k < m; k++) {
      const double a1 = -2.0 * M_PI * (double)k / (double)n, a2 = -2.0 * M_PI * (double)(2 * k) / (double)n;
      tw[m + 2 * k] = (float)std::cos(a1); tw[m + 2 * k + 1] = (float)std::sin(a1);
      tw[3 * m + 2 * k] = (float)std::cos(a2); tw[3 * m + 2 * k + 1] = (float)std::sin(a2);
    }
    return;
  }
  tw.resize(n);
  for (int k = 0; k < n / 2; k++) {
    const double a = -2.0 * M_PI * (double)k / (double)n;
    float re = (float)std::cos(a), im = (float)std::sin(a);
    if ((4 * k) % n == 0) { re = (4 * k / n == 0) ? 1.0f : 0.0f; im = (4 * k / n == 0) ? 0.0f : -1.0f; }
    tw[2 * k] = re; tw[2 * k + 1] = im;
  }
  if (n == 2048) {
    // Per-pass copies for the radix 16 x 16 x 8 kernel (ofdm.cu, fft2048_r16), laid out [twiddle of the pass][k] so that
    // consecutive k are consecutive entries.  Same values as above: twiddle j of a pass that combines sub-transforms of
    // length L at index k is w_n^((k + c_j L) n / (2^s L)) with s the stage (1..4) and c_j the offset of its butterfly group.
    auto entry = [&](int idx) { return std::pair<float, float>(tw[2 * idx], tw[2 * idx + 1]); };
    auto append = [&](int L, int stages, int nk) {
      for (int j = 0; j < (1 << stages) - 1; j++) {
        int s = 1, c = j;                                        // j = 2^(s-1) - 1 + c, c < 2^(s-1)
        while (c >= (1 << (s - 1))) { c -= 1 << (s - 1); s++; }
        for (int k = 0; k < nk; k++) {
          const auto e = entry((k + c * L) * (n / ((1 << s) * L)));
          tw.push_back(e.first); tw.push_back(e.second);
        }
      }
    };
    append(16, 4, 16);     // pass 2: 15 x 16
    append(256, 3, 256);   // pass 3:  7 x 256
  }
}

// Quadruplet i of the PCFICH sits in the resource-element group starting at kbar + floor(i N_RB / 2) * 6 with
// kbar = 6 (N_ID mod 2 N_RB); the REs with k mod 3 == N_ID mod 3 belong to the CRS of ports 0/1 (both always assumed
// for this mapping, 36.211 6.2.4).  Replaces the PCFICH part of srslte_ue_dl_decode_fft_estimate
// (/root/reference/ue/src/phy/phch_worker.cc:254).
void pcfich_re(const CellCfg& cell, int32_t* k16) {
  const int nrb = cell.nof_prb, nsc = 12 * nrb;
  const int kbar = 6 * (cell.cell_id % (2 * nrb));
  int n = 0;
  for (int i = 0; i < 4; i++) {
    const int k0 = (kbar + ((i * nrb) / 2) * 6) % nsc;
    for (int j = 0; j < 6; j++)
      if ((k0 + j) % 3 != cell.cell_id % 3) k16[n++] = k0 + j;
  }
}

uint32_t pcfich_scramble(const CellCfg& cell, int sf_idx) {
  const uint32_t c_init = (((uint32_t)(sf_idx + 1) * (uint32_t)(2 * cell.cell_id + 1)) << 9) + (uint32_t)cell.cell_id;
  std::vector<uint32_t> w;
  gold_packed(c_init, 32, w);
  return w[0];
}

// ---- PDCCH tables.  Replace the bookkeeping inside srslte_pdcch_extract_llr / srslte_ue_dl_find_dl_dci_type
// (/root/reference/ue/src/phy/phch_worker.cc:260,293): srsLTE's regs.c, pdcch.c and dci.c.
int ctrl_symbols(int nof_prb, int cfi) { return cfi + (nof_prb <= 10 ? 1 : 0); }

int pdcch_regs(const CellCfg& cell, int cfi, int ng_x6, std::vector<int32_t>& re4) {
  const int nrb = cell.nof_prb, nsc = 12 * nrb, nsym = ctrl_symbols(nrb, cfi), n0 = 2 * nrb;
  std::vector<uint8_t> taken(n0, 0);
  int32_t pc[16];
  pcfich_re(cell, pc);
  for (int i = 0; i < 4; i++) taken[pc[4 * i] / 6] = 1;
  // PHICH groups (normal duration: symbol 0 only, 36.211 6.9.3) over the REGs PCFICH left free
  std::vector<int> free0;
  for (int r = 0; r < n0; r++) if (!taken[r]) free0.push_back(r);
  const int np0 = (int)free0.size(), groups = (ng_x6 * nrb + 47) / 48;
  for (int g = 0; g < groups; g++)
    for (int i = 0; i < 3; i++) taken[free0[(cell.cell_id + g + (i * np0) / 3) % np0]] = 1;
  re4.clear();
  for (int k = 0; k < nsc; k += 2)
    for (int l = 0; l < nsym; l++) {
      if (l == 0) {
        if (k % 6 || taken[k / 6]) continue;
        for (int j = 0; j < 6; j++) if ((k + j) % 3 != cell.cell_id % 3) re4.push_back(k + j);
      } else if ((cell.cp && l == 3) || (cell.nof_ports == 4 && l == 1)) {
        // extended cyclic prefix: the fourth control symbol (<= 10 PRB) carries CRS; four ports: symbol 1 does (ports 2 / 3)
        if (k % 6) continue;
        for (int j = 0; j < 6; j++) if ((k + j) % 3 != cell.cell_id % 3) re4.push_back(l * nsc + k + j);
      } else {
        if (k % 4) continue;
        for (int j = 0; j < 4; j++) re4.push_back(l * nsc + k + j);
      }
    }
  return (int)re4.size() / 4;
}

namespace {
const uint8_t kCcCols[32] = {1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30};
// 36.212 5.1.4.2.1: read order of D elements written row-wise (dummies first) into 32 permuted columns
void cc_interleaver(int D, std::vector<int32_t>& out) {
  const int R = (D + 31) / 32, nd = 32 * R - D;
  out.clear();
  for (int c = 0; c < 32; c++)
    for (int r = 0; r < R; r++) {
      const int y = r * 32 + kCcCols[c];
      if (y >= nd) out.push_back(y - nd);
    }
}
}  // namespace

void pdcch_quad_perm(int n_quad, int cell_id, std::vector<int32_t>& src) {
  std::vector<int32_t> w;
  cc_interleaver(n_quad, w);
  src.resize(n_quad);
  for (int m = 0; m < n_quad; m++) src[m] = w[(m + cell_id) % n_quad];
}

void cc_rm_sequence(int D, std::vector<int32_t>& seq) {
  std::vector<int32_t> p;
  cc_interleaver(D, p);
  seq.clear();
  for (int s = 0; s < 3; s++)
    for (int j = 0; j < D; j++) seq.push_back(s * D + p[j]);
}

int pdcch_search_space(int nof_cce, int sf_idx, uint16_t rnti, bool common, int32_t* cand_L, int32_t* cand_ncce) {
  int n = 0;
  if (common) {
    const int lim = std::min(nof_cce, 16);
    for (int L = 4, M = 4; L <= 8; L *= 2, M /= 2)
      for (int m = 0; m < M; m++)
        if ((m + 1) * L <= lim) { cand_L[n] = L; cand_ncce[n] = m * L; n++; }
    return n;
  }
  uint32_t Y = rnti;
  for (int k = 0; k <= sf_idx; k++) Y = (39827u * Y) % 65537u;
  const int Ms[4] = {6, 6, 2, 2};
  for (int a = 0; a < 4; a++) {
    const int L = 1 << a, nl = nof_cce / L;
    for (int m = 0; m < Ms[a] && m < nl; m++) { cand_L[n] = L; cand_ncce[n] = L * (int)((Y + (uint32_t)m) % (uint32_t)nl); n++; }
  }
  return n;
}

void cfo_table(std::vector<float>& tab) {
  const int n = 1 << kCfoTableLog2;
  tab.resize(2 * (size_t)n);
  for (int j = 0; j < n; j++) {
    const double a = 2.0 * 3.14159265358979323846 * (double)j / (double)n;
    tab[2 * j] = (float)std::cos(a);
    tab[2 * j + 1] = (float)std::sin(a);
  }
}

int32_t cfo_step(float cfo, int nfft) {
  const double s = -(double)cfo / (double)nfft * 4294967296.0;
  return (int32_t)(uint32_t)(uint64_t)std::llrint(s);
}

int dci_format_sizeof(int fmt, int nof_prb) {
  auto ambiguous = [](int n) { return n == 12 || n == 14 || n == 16 || n == 20 || n == 24 || n == 26 || n == 32 || n == 40 || n == 44 || n == 56; };
  int riv = 0;
  while ((1 << riv) < nof_prb * (nof_prb + 1) / 2) riv++;
  int n1a = 15 + riv;                         // flag, local/distributed, RIV, MCS 5, HARQ 3, NDI, RV 2, TPC 2 (>= format 0)
  if (ambiguous(n1a)) n1a++;
  if (fmt == 0) return n1a;
  const int P = nof_prb <= 10 ? 1 : nof_prb <= 26 ? 2 : nof_prb <= 63 ? 3 : 4;
  int n1 = (nof_prb > 10 ? 1 : 0) + (nof_prb + P - 1) / P + 13;   // RA header, bitmap, MCS 5, HARQ 3, NDI, RV 2, TPC 2
  while (n1 == n1a || ambiguous(n1)) n1++;
  return n1;
}

int phich_groups(int nof_prb, int ng_x6) { return (ng_x6 * nof_prb + 47) / 48; }

void phich_res(const CellCfg& cell, int n_group, int32_t* k12) {
  const int n0 = 2 * cell.nof_prb;
  if (cell.cp) n_group /= 2;       // extended cyclic prefix: groups 2m' and 2m'+1 share mapping unit m' (36.211 6.9.3)
  std::vector<uint8_t> taken(n0, 0);
  int32_t pc[16];
  pcfich_re(cell, pc);
  for (int i = 0; i < 4; i++) taken[pc[4 * i] / 6] = 1;
  std::vector<int> free0;
  for (int r = 0; r < n0; r++) if (!taken[r]) free0.push_back(r);
  const int np0 = (int)free0.size();
  for (int i = 0, n = 0; i < 3; i++) {
    const int k0 = 6 * free0[(cell.cell_id + n_group + (i * np0) / 3) % np0];
    for (int j = 0; j < 6; j++) if ((k0 + j) % 3 != cell.cell_id % 3) k12[n++] = k0 + j;
  }
}

void phich_index(int nof_prb, int ng_x6, int I_lowest, int n_dmrs, int* n_group, int* n_seq, int cp) {
  // 36.213 9.1.2: N_group doubles and the sequence index is taken modulo 2 N_SF = 4 with the extended cyclic prefix
  const int g = (cp ? 2 : 1) * phich_groups(nof_prb, ng_x6);
  *n_group = (I_lowest + n_dmrs) % g;
  *n_seq = (I_lowest / g + n_dmrs) % (cp ? 4 : 8);
}

// slot 1, symbols 0..3, the 72 central subcarriers, k first then l; CRS positions of ports 0..3 left out (symbols 0, 1)
// Returns the number of elements: 240, or 216 with the extended cyclic prefix (symbol 3 of the slot carries CRS too).
int pbch_res(const CellCfg& cell, int32_t* g240) {
  const int nsc = 12 * cell.nof_prb, k0 = nsc / 2 - 36, first = slot_symb(cell.cp);
  int n = 0;
  for (int l = 0; l < 4; l++)
    for (int k = 0; k < 72; k++) {
      if ((l < 2 || (cell.cp && l == 3)) && (k0 + k) % 3 == cell.cell_id % 3) continue;
      g240[n++] = (first + l) * nsc + k0 + k;
    }
  return n;
}

// ---- synchronisation signals.  Replace the sequence generators behind srslte_ue_cellsearch_scan
// (/root/reference/ue/src/phy/phch_recv.cc:146-177): srsLTE's pss.c / sss.c.
namespace {
void pss_freq_d(int n_id_2, double* re, double* im) {
  static const int roots[3] = {25, 29, 34};
  const int u = roots[n_id_2];
  for (int n = 0; n < 62; n++) {
    const int m = (n < 31) ? n * (n + 1) : (n + 1) * (n + 2);
    const double a = -M_PI * (double)u * (double)(m % 126) / 63.0;
    re[n] = std::cos(a); im[n] = std::sin(a);
  }
}
void mseq31(int taps, int8_t* s31) {
  int x[31] = {0, 0, 0, 0, 1};
  for (int i = 0; i < 26; i++) {
    int v = 0;
    for (int j = 0; j < 5; j++) if (taps & (1 << j)) v ^= x[i + j];
    x[i + 5] = v;
  }
  for (int i = 0; i < 31; i++) s31[i] = (int8_t)(1 - 2 * x[i]);
}
}  // namespace

void pss_freq(int n_id_2, float* d62x2) {
  double re[62], im[62];
  pss_freq_d(n_id_2, re, im);
  for (int n = 0; n < 62; n++) { d62x2[2 * n] = (float)re[n]; d62x2[2 * n + 1] = (float)im[n]; }
}

void pss_time_n(int n_id_2, int nfft, float* tx2) {
  double dr[62], di[62];
  pss_freq_d(n_id_2, dr, di);
  for (int n = 0; n < nfft; n++) {
    double re = 0, im = 0;
    for (int i = 0; i < 62; i++) {
      const int bin = (i < 31) ? i - 31 : i - 30;
      const double a = 2.0 * M_PI * (double)(bin * n) / (double)nfft;
      re += dr[i] * std::cos(a) - di[i] * std::sin(a);
      im += dr[i] * std::sin(a) + di[i] * std::cos(a);
    }
    tx2[2 * n] = (float)(re / std::sqrt((double)nfft)); tx2[2 * n + 1] = (float)(im / std::sqrt((double)nfft));
  }
}
void pss_time(int n_id_2, float* t128x2) { pss_time_n(n_id_2, 128, t128x2); }

void sss_seq(int n_id_1, int n_id_2, int sf5, int8_t* d62) {
  int8_t s[31], c[31], z[31];
  mseq31(0x05, s); mseq31(0x09, c); mseq31(0x17, z);
  const int qp = n_id_1 / 30, q = (n_id_1 + qp * (qp + 1) / 2) / 30, mp = n_id_1 + q * (q + 1) / 2;
  const int m0 = mp % 31, m1 = (m0 + mp / 31 + 1) % 31;
  for (int n = 0; n < 31; n++) {
    const int s0 = s[(n + m0) % 31], s1 = s[(n + m1) % 31];
    const int c0 = c[(n + n_id_2) % 31], c1 = c[(n + n_id_2 + 3) % 31];
    const int z0 = z[(n + (m0 % 8)) % 31], z1 = z[(n + (m1 % 8)) % 31];
    if (!sf5) { d62[2 * n] = (int8_t)(s0 * c0); d62[2 * n + 1] = (int8_t)(s1 * c1 * z0); }
    else { d62[2 * n] = (int8_t)(s1 * c0); d62[2 * n + 1] = (int8_t)(s0 * c1 * z1); }
  }
}

}  // namespace srsue

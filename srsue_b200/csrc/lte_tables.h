// lte_tables.h -- host-side LTE tables and derived look-up tables for libsrsue_gpu.
//
// Everything here is integer bookkeeping that srsLTE performs inside srslte_ue_dl_cfg_grant /
// srslte_pdsch_init / srslte_tdec_init (called from /root/reference/ue/src/phy/phch_worker.cc:74,337);
// the results are uploaded once per (cell, grant shape) or per code-block size and cached on the device.
#pragma once
#include <cstdint>
#include <vector>

namespace srsue {

constexpr int kMaxK = 6144;
constexpr int kTdC = 511;       // clamp of channel LLRs at the decoder input / soft-buffer saturation
constexpr int kTdE = 2047;      // clamp of extrinsic LLRs
constexpr int kTdInf = 10000;   // finite minus-infinity of known trellis states
constexpr uint32_t kCrc24A = 0x1864CFBu, kCrc24B = 0x1800063u;

// cp: 0 = normal cyclic prefix (7 symbols per slot), 1 = extended (6).  Subframe grids keep a stride of 14 symbols per port
// in either case; with the extended prefix rows 12 and 13 are never touched.
struct CellCfg { int nof_prb, nof_ports, cell_id, cp = 0; };
inline int nof_symb(int cp) { return cp ? 12 : 14; }
inline int slot_symb(int cp) { return cp ? 6 : 7; }

struct PdschCfg {
  int sf_idx, cfi, rnti, qm, tbs, rv, tm, nof_prb_alloc;
  uint8_t prb_mask[110];     // bit 0: PRB allocated in both slots; bit 1: in slot 0 only; bit 2: in slot 1 only
};
inline bool prb_in_slot(uint8_t m, int slot) { return (m & 1) || (m & (slot ? 4 : 2)); }


struct CbSegm { int tbs, B, C, Kp, Km, Cp, Cm, F; };

// Geometry of the windowed decoder for one code-block size (oracle/SPEC.md 7.3) and of the
// device-native decoder-input layout ("tcb"): three planes sys/par1/par2 of W x Ppad int16 followed by
// 16 int16 holding the 12 tail values.  Inside a plane the element of (window j, step i) sits at
// (((i/4)*T + j/2)*4 + i%4)*2 + j%2: the 4 steps x 2 windows of one decoder thread are one 16-byte chunk and the
// chunks of consecutive threads are contiguous, so the 16-byte asynchronous copy a warp issues covers whole sectors.
struct TurboGeom {
  int K, W, P, Ppad, T;   // T = Ppad / 2 threads per code block
  int plane;              // W * Ppad
  int cb_elems;           // 3 * plane + 16
};

bool qpp_params(int K, int* f1, int* f2);
int qpp_index(int K);                       // row in the table, -1 if K is not a valid size
int qpp_K(int idx);
int window_len(int K);
TurboGeom turbo_geom(int K);
int symbol_sz(int nof_prb);
inline int cp_len(int nfft, int l) { return ((l % 7) == 0 ? 160 : 144) * nfft / 2048; }
inline int cp_len_x(int nfft, int l, int cp) { return cp ? 512 * nfft / 2048 : cp_len(nfft, l); }

uint32_t crc_bits(const uint8_t* bits, int n, uint32_t poly, int order);
// x^e mod g(x) for the 24-bit CRC polynomials
uint32_t crc_xpow(uint32_t poly, uint64_t e);

void gold_bits(uint32_t c_init, int n, uint8_t* c);
// packed LSB-first: bit i of the sequence is (w[i/32] >> (i%32)) & 1
void gold_packed(uint32_t c_init, int n, std::vector<uint32_t>& w);

// PCFICH (36.211 6.7): subcarriers of d(0..15) in OFDM symbol 0, and the 32 scrambling bits of subframe sf_idx packed
// LSB first
void pcfich_re(const CellCfg& cell, int32_t* k16);
uint32_t pcfich_scramble(const CellCfg& cell, int sf_idx);

// ---- PDCCH (36.211 6.8, 36.212 5.1.3.1 / 5.1.4.2, 36.213 9.1.1) ----
int ctrl_symbols(int nof_prb, int cfi);
// REGs of the control region that carry PDCCH in mapping order (k' ascending, then l'); re4[4*m + i] = grid index
// (l * nsc + k) of the i-th data RE of REG m.  ng_x6 = 6 * Ng (1, 3, 6, 12), normal PHICH duration.
int pdcch_regs(const CellCfg& cell, int cfi, int ng_x6, std::vector<int32_t>& re4);
// src[m'] = quadruplet of the multiplexed PDCCH stream carried by REG m' (interleaver + cyclic shift by N_ID)
void pdcch_quad_perm(int n_quad, int cell_id, std::vector<int32_t>& src);
// circular-buffer order of the convolutional-code rate matcher for D = payload + 16 bits: seq[j] = stream * D + k
void cc_rm_sequence(int D, std::vector<int32_t>& seq);
// search-space candidates (L, first CCE) of subframe sf_idx for rnti (UE-specific) or the common space
int pdcch_search_space(int nof_cce, int sf_idx, uint16_t rnti, bool common, int32_t* cand_L, int32_t* cand_ncce);
// Cell search (36.211 6.11): PSS in frequency (62 values) and as a 128-sample time replica at 1.92 Msps (both evaluated in
// double and rounded once, interleaved re/im), SSS as +-1 for (N_id_1, N_id_2, subframe 0 or 5)
void pss_freq(int n_id_2, float* d62x2);
void pss_time(int n_id_2, float* t128x2);
void pss_time_n(int n_id_2, int nfft, float* tx2);   // the same replica at any LTE sampling rate
void sss_seq(int n_id_1, int n_id_2, int sf5, int8_t* d62);
// PBCH (36.211 6.6.4): grid indices of the 240 (extended cyclic prefix: 216) resource elements in a subframe 0
int pbch_res(const CellCfg& cell, int32_t* g240);
// PHICH (36.211 6.9, 36.213 9.1.2; normal duration): number of groups (of mapping units with the extended prefix), the 12 subcarriers of a group in symbol 0,
// and (group, sequence) of the indicator that answers an uplink transmission
int phich_groups(int nof_prb, int ng_x6);
void phich_res(const CellCfg& cell, int n_group, int32_t* k12);
void phich_index(int nof_prb, int ng_x6, int I_lowest, int n_dmrs, int* n_group, int* n_seq, int cp = 0);
// payload size of DCI format 1A (fmt = 0) or 1 (fmt = 1) for FDD, including the padding rules of 36.212 5.3.3.1
int dci_format_sizeof(int fmt, int nof_prb);

bool cbsegm(int tbs, CbSegm* s);
inline int cb_len(const CbSegm& s, int r) { return r < s.Cm ? s.Km : s.Kp; }
int cb_E(const CbSegm& s, int G, int qm, int nl, int r);

// ordered PDSCH resource elements as indices l*nsc + k into the subframe grid
void pdsch_re_list(const CellCfg& cell, const PdschCfg& cfg, std::vector<int32_t>& re);
// CRS of one OFDM symbol: first pilot offset for `port` (-1 if none) and the +-1 signs of re/im
int crs_offset(const CellCfg& cell, int port, int l);
void crs_signs(const CellCfg& cell, int sf_idx, int l, std::vector<int8_t>& re_sign, std::vector<int8_t>& im_sign);

// DEC2 access table, [W][2][T] entries: for trellis step i of window 2t + h the BYTE offset of pi((2t + h) * W + i) in the
// shared exchange array A, which is laid out [W][Ppad] int16 (consecutive threads read consecutive 16-bit entries)
void turbo_perm_table(const TurboGeom& g, std::vector<uint16_t>& tab);
// CRC contribution of the hard bit decided at every DEC2 trellis step, [W/2][T][2][2] entries (one 16-byte chunk per
// thread and pair of steps, chunks of consecutive threads contiguous): the CRC of the K decoded bits is the XOR over the set bits n of x^(K-1-n+24) mod g (zero
// initial state, no final xor); entry (q, t, i, h) belongs to n = pi((2t + h) * W + 2 q + i), 0 in padding columns
void turbo_deint_table(const TurboGeom& g, std::vector<uint16_t>& tab);   // [K]: natural bit -> source bit of the decoder's DEC2-order row
void turbo_crc_table(const TurboGeom& g, uint32_t poly, std::vector<uint32_t>& tlin);

// Rate de-matching gather table for one (K, F, rv): for every element m of the tcb buffer
// (cb_elems entries) the first circular-buffer read index n in [0, N) that lands on it, 0xFFFF if
// none (padding), 0xFFFE for filler positions.  Returns N (number of non-null positions).
int rm_gather_table(const TurboGeom& g, int F, int rv, std::vector<uint16_t>& tab);

// offset of decoder-input element (srsLTE triples index 3k+stream, k < K+4) inside the tcb buffer
int tcb_offset(const TurboGeom& g, int triple_index);

void rm_tx_sequence(int K, int F, int rv, std::vector<uint16_t>& seq);   // (k << 2) | stream in circular-buffer read order
void fft_twiddles(int n, std::vector<float>& tw /* re,im pairs, n/2 entries */);
// CFO correction (SPEC.md 14): 4096-entry unit circle and the 2^-32-turn phase step per sample
constexpr int kCfoTableLog2 = 12;
void cfo_table(std::vector<float>& tab /* re,im pairs, 4096 entries */);
int32_t cfo_step(float cfo, int nfft);

}  // namespace srsue

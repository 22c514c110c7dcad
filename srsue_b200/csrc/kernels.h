// kernels.h -- device-side argument structs and kernel declarations of libsrsue_gpu (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lte_tables.h"

namespace srsue {

constexpr int kTurboMaxThreads = 384;

struct TurboGeomDev { int K, W, P, Ppad, T, plane, cb_elems; };

struct TurboArgs {
  const int16_t* in;         // code blocks in tcb layout
  long long in_stride;       // int16 elements between consecutive code blocks
  const int32_t* cb_list;    // optional indirection (nullptr: code block i is entry i)
  int n_cb;
  uint8_t* dbits;            // [launch-local cb][dbits_stride] hard decisions in DEC2 order (tdec_deinterleave_kernel reads them)
  int dbits_stride;          // bytes, even
  int32_t* out_status;       // [cb] iterations | crc_ok << 8
  int max_iter, crc_type;    // crc_type: 0 none, 1 CRC24A, 2 CRC24B
  int min_iter;              // a passing CRC stops the block only from this iteration on (1: as srsLTE; = max_iter: fixed count)
  int K, W, P, Ppad, T, plane;
  const uint16_t* perm_tab;  // [W][2][T]: byte offset in the exchange array A of pi((2t + h) * W + i)  (turbo_perm_table)
  const uint32_t* crc_lin;   // [W/2][T][2][2]: x^(K-1-n+24) mod g for n = pi((2t + h) * W + 2 q + i), 0 in padding columns
  int ncb_cta;               // code-block slots per CTA
  int slot_words;            // 32-bit words between the exchange arrays of consecutive slots (plane/2 + skew)
  uint4* nii;                // [grid * ncb_cta][2 dec][2 parity][2 kinds][Ppad + 2] records of 8 x int16 (turbo.cu)
  uint4* ckpt;               // [grid][W/8][2][ncb_cta][T]  beta checkpoints (two 16-byte halves per thread)
  int* work_counter;         // dynamic work items handed out so far (zeroed before the launch)
  int work_base;             // first dynamically assigned code block = grid * ncb_cta
  uint32_t ones;             // 0xFFFFFFFF, opaque to the compiler (turbo.cu: vnot)
  int ngroups;               // phase groups per CTA (1 or 2): slots of different groups only meet at named barriers of their own
  int group_threads;         // threads per phase group (multiple of 32); blockDim.x = ngroups * group_threads
  int group_slots;           // code-block slots per phase group
  int phase_delay;           // SM clocks the second group waits before its first block (puts the groups out of phase)
};

__global__ void turbo_decode_kernel(const TurboArgs g);        // fixed iteration count
__global__ void turbo_decode_crc_kernel(const TurboArgs g);    // CRC accumulated on the fly, early stop
__global__ void turbo_decode_wide_kernel(const TurboArgs g);       // the same for code blocks with more than 32 threads (T > 32)
__global__ void turbo_decode_crc_wide_kernel(const TurboArgs g);
__global__ void turbo_decode_t26_kernel(const TurboArgs g);        // T = 26 (K = 5824) as a compile-time constant
__global__ void turbo_decode_crc_t26_kernel(const TurboArgs g);
__global__ void turbo_decode_t24_kernel(const TurboArgs g);        // T = 24 (K = 6144)
__global__ void turbo_decode_crc_t24_kernel(const TurboArgs g);
__global__ void tdec_deinterleave_kernel(const uint8_t* dbits, int dbits_stride, const uint16_t* deint, const int32_t* cb_list,
                                         int n_cb, uint8_t* out, int out_stride, int K, int row_bytes);
__global__ void triples_to_tcb_kernel(const int16_t* in, long long in_stride, int16_t* out, long long out_stride,
                                      int n_cb, TurboGeomDev g);
__global__ void tcb_to_triples_kernel(const int16_t* in, long long in_stride, int16_t* out, long long out_stride,
                                      int n_cb, TurboGeomDev g);

// ---- uplink shared-channel encoder (ulsch.cu) ---------------------------------------------------------
struct UlschArgs {
  const uint8_t* payload;    // [n_tb][payload_stride] transport blocks, tbs / 8 bytes each
  int payload_stride;
  uint8_t* tbcrc;            // [n_tb][4] scratch: the CRC24A bytes of every transport block
  uint8_t* ebits;            // [n_tb][G] scratch: the codeword before the channel interleaver, one bit per byte
  uint8_t* out;              // [n_tb][out_stride] interleaved, scrambled bits, packed MSB first
  int out_stride;
  int n_tb, tbs, C, G, qm, rows, n_symb;      // rows = 12 nof_prb
  const int32_t* cbtab;      // [C][8]: K, F, E, first bit of the block in the codeword, first stream byte, stream bytes,
                             //         offset of its QPP table in perm, offset of its read order in seq
  const int32_t* seq_len;    // [C] entries of the block's read order (circular buffer without <NULL>)
  const uint16_t* perm;      // QPP tables pi(i) of the code-block sizes in use
  const uint16_t* seq;       // circular-buffer read orders: (index k of the triple << 2) | stream
  const uint32_t* crcshift;  // [1 + C][32]: x^(8 * bytes after lane's chunk) mod g for the TB CRC24A, then every block's CRC24B
  const uint8_t* scramble;   // [G / 8] Gold sequence packed MSB first
};
__global__ void ulsch_tbcrc_kernel(const UlschArgs a);
__global__ void ulsch_encode_kernel(const UlschArgs a);
__global__ void ulsch_interleave_kernel(const UlschArgs a);

// ---- front end -------------------------------------------------------------------------------------
struct OfdmArgs {
  const float2* iq;          // [n_sf][15 * nfft]
  float2* sf_symbols;        // [n_sf][14 * nsc]
  const float2* tw;          // nfft/2 twiddles (1536: 256 of the 512-point transforms, then w_1536^k and w_1536^2k, k < 512)
  int n_sf, nfft, log2n, nsc;
  float scale;
  float c3;                  // (float)(sqrt(3)/2), radix-3 stage of the 1536-point transform
  // CFO correction fused into the first pass (ofdm_rx_cfo_kernel only; SPEC.md 14)
  const float2* cexp;        // 4096-entry unit circle
  const int32_t* cfo_steps;  // [n_sf] phase step per sample, or nullptr: cfo_step for every subframe
  int32_t cfo_step;
  // int16 {re, im} input (ofdm_rx_*iq16_kernel only): sample = (float)v * iq16_scale
  const short2* iq16;
  float iq16_scale;
  int cp_ext;                // 1: extended cyclic prefix (grid x = 12 symbols; rows 12, 13 of every grid stay untouched)
};
__global__ void ofdm_rx_kernel(const OfdmArgs a);
__global__ void ofdm_rx_r16_kernel(const OfdmArgs a);          // N = 2048 only: radix 16 x 16 x 8, 128 threads (ofdm.cu)
__global__ void ofdm_rx_r16_iq16_kernel(const OfdmArgs a);
__global__ void ofdm_rx_cfo_kernel(const OfdmArgs a);
__global__ void ofdm_rx_inplace_iq16_kernel(const OfdmArgs a);
__global__ void ofdm_rx_iq16_kernel(const OfdmArgs a);
__global__ void ofdm_rx_cfo_iq16_kernel(const OfdmArgs a);
__global__ void ofdm_rx_inplace_kernel(const OfdmArgs a);

struct ChestArgs {
  const float2* sf_symbols;  // [n_sf][14 * nsc]
  float2* ce;                // [n_sf][ports][14 * nsc]; nullptr: skip the interpolation (fused path)
  float2* pilots;            // optional [n_sf][ports][4][2 * nof_prb] smoothed pilot estimates
  float* meas;               // [n_sf][5]  noise, rsrp, rssi, rsrq, snr
  const int8_t* crs_sign;    // [4 crs symbols (+ the 2 of ports 2 / 3 in a four-port cell)][2 (re, im)][2 * nof_prb]
  int n_sf, nsc, nof_prb, nof_ports;
  int crs_off[4][4];         // first pilot subcarrier per port and CRS symbol (ports 2 / 3: two symbols)
};
__global__ void chest_kernel(const ChestArgs a);       // normal cyclic prefix: CRS in symbols 0, 4, 7, 11 of 14
__global__ void chest_ext_kernel(const ChestArgs a);   // extended cyclic prefix: CRS in symbols 0, 3, 6, 9 of 12
__global__ void chest_p4_kernel(const ChestArgs a);    // four-port cells (ports 2 / 3: symbol 1 of each slot)
__global__ void chest_ext_p4_kernel(const ChestArgs a);

struct DemodArgs {
  const float2* sf_symbols;  // [n_sf][14 * nsc]
  const float2* ce;          // [n_sf][ports][14 * nsc] (unused when pilots != nullptr)
  const float2* pilots;      // optional [n_sf][ports][4][2 * nof_prb]: interpolate the channel on the fly
  int nof_prb;
  int cp_ext;                // fused interpolation only: CRS symbols 0, 3, 6, 9 instead of 0, 4, 7, 11
  int crs_off[2][4];         // first pilot subcarrier per port and CRS symbol
  const float* meas;         // [n_sf][5] (noise estimate when noise_mode == 1)
  int16_t* softbuf;          // [n_sf][C][sb_stride] tcb layout
  const int32_t* re_idx;     // [nof_re] grid index of every PDSCH RE
  const uint32_t* scramble;  // packed Gold sequence, G bits
  const uint16_t* gather;    // [C][gather_stride] per code block: tcb element -> LLR index in the CB's range
  const int32_t* cb_e_start; // [C+1] first LLR index of each code block
  const int32_t* cb_geom;    // [C][4]: cb_elems, N (non-null positions), K, unused
  float2* dbg_d;             // optional [n_sf][nof_re] equalised symbols
  int16_t* dbg_e;            // optional [n_sf][G] descrambled LLRs
  int n_sf, nsc, nof_ports, tm, qm, nof_re, C, gather_stride;
  long long sb_stride;       // elements per code block in softbuf
  float noise_est;
  float k_sqpsk, k_c16, k_c64a, k_c64b, k_sq2;   // demapper constants, rounded on the host (SPEC 5)
  int noise_mode;            // 0: noise_est, 1: meas[0]
  int accumulate;            // 0: new transmission (buffer overwritten), 1: add to existing
  int direct;                // gather tables hold direct LLR indices with sentinels E (zero) and E+1 (filler)
};
__global__ void pdsch_llr_dematch_kernel(const DemodArgs a);

struct PcfichArgs {
  const float2* sf_symbols;  // [n_sf][14 * nsc]
  const float2* ce;          // [n_sf][ports][14 * nsc]
  const float* meas;         // [n_sf][5] (noise estimate when noise_mode == 1)
  int32_t* cfi;              // [n_sf] decoded CFI 1..3
  int32_t* corr;             // optional [n_sf][3] correlations with the three code words
  int re[16];                // subcarriers of d(0..15) in symbol 0
  uint32_t scramble;         // 32 scrambling bits, LSB first
  int n_sf, nsc, nof_ports, noise_mode;
  float noise_est, k_sqpsk, k_sq2;
};
__global__ void pcfich_kernel(const PcfichArgs a);

struct PdcchLlrArgs {
  const float2* sf_symbols;  // [n_sf][14 * nsc]
  const float2* ce;          // [n_sf][ports][14 * nsc]
  const float* meas;         // [n_sf][5]
  const int32_t* re4;        // [n_reg][4] grid index of the data REs of every PDCCH REG, mapping order
  const int32_t* src;        // [n_reg] quadruplet of the PDCCH bit stream carried by the REG
  const uint32_t* scramble;  // 8 * n_reg scrambling bits, packed LSB first
  int16_t* llr;              // [n_sf][8 * n_reg]: 72 LLRs per CCE in CCE order
  int n_sf, nsc, nof_ports, n_reg, noise_mode;
  float noise_est, k_sqpsk, k_sq2;
  const int32_t* row_filter; // optional [n_sf]: subframes whose entry differs from row_want are skipped (blind batches:
  int row_want;              // the control region of a plan belongs to one CFI, the batch mixes them)
};
__global__ void pdcch_llr_kernel(const PdcchLlrArgs a);

constexpr int kPdcchMaxCand = 24;
struct PdcchSearchArgs {
  const int16_t* llr;        // [n_sf][llr_stride]
  long long llr_stride;
  const int32_t* rm_seq;     // [3 D] rate-matching order for D = nof_bits + 16
  int32_t* found;            // [n_sf][4]: found (0/1), L, first CCE, candidate index
  uint8_t* bits;             // [n_sf][64]: payload of the match, one bit per byte
  uint16_t* rem;             // optional [n_sf][n_cand]: RNTI every candidate decodes to
  int n_sf, n_cand, nof_bits, rnti;
  int first_bit;             // -1: any; 0 / 1: a match also needs this value in payload bit 0 (format 0 / 1A flag)
  int cand_L[kPdcchMaxCand], cand_ncce[kPdcchMaxCand];
  const int32_t* row_filter; // as in PdcchLlrArgs; skipped subframes report "not found"
  int row_want;
};
__global__ void pdcch_search_kernel(const PdcchSearchArgs a);

struct PhichArgs {
  const float2* sf_symbols;  // [n_sf][14 * nsc]
  const float2* ce;          // [n_sf][ports][14 * nsc]
  const float* meas;         // [n_sf][5]
  int32_t* ack;              // [n_sf] HARQ indicator: 1 = ACK
  float* metric;             // optional [n_sf]: the decision metric (ACK iff < 0)
  int re[12];                // subcarriers (symbol 0) of the group's 3 REGs
  uint32_t scramble;         // 12 scrambling bits, LSB first
  int n_sf, nsc, nof_ports, n_seq, noise_mode;
  float noise_est, k_sq2;
  int ext;                   // extended cyclic prefix: spreading factor 2, the group owns one half of every quadruplet
  int odd;                   // ext: n_group & 1 (the second half)
  int par0;                  // four ports: quadruplet i uses ports (0, 2) when i + par0 is even, else (1, 3); par0 = n_group (ext: n_group / 2)
};
__global__ void phich_kernel(const PhichArgs a);

struct PbchArgs {
  const float2* sf_symbols;  // [n_sf][14 * nsc] grids of subframes 0
  const float2* ce;          // [n_sf][ports][14 * nsc]
  const float* meas;         // [n_sf][5]
  const int32_t* re;         // [n_re] grid index of the PBCH resource elements
  const uint32_t* scramble;  // 8 n_re scrambling bits (four radio frames), packed LSB first
  int n_re;                  // 240, or 216 with the extended cyclic prefix
  const int32_t* rm_seq;     // [120] rate-matching order for D = 40
  int32_t* result;           // [n_sf][4]: found, transmit ports, frame number mod 4, 0
  uint8_t* mib;              // [n_sf][24] MIB bits, one per byte
  int n_sf, nsc, nof_ports, noise_mode;
  float noise_est, k_sqpsk, k_sq2;
};
__global__ void pbch_kernel(const PbchArgs a);

// result of the cell search on one buffer (mirrors srsue_gpu_sync_result_t of the C ABI)
struct srsue_sync_result { int32_t peak_pos, n_id_2, n_id_1, sf5, valid; float peak, mean_power, cfo, sss_corr; int32_t cp; };
struct SyncArgs {
  const float2* iq;          // [n_bufs][stride] samples at 1.92 Msps
  long long stride;
  int n_samples, n_bufs;
  int nfft, log2n;           // samples per OFDM symbol at the buffer's sampling rate (128 at 1.92 Msps ... 2048)
  int force_n_id_2;          // -1: search the three roots, else only this one
  int first_pos;             // first sample offset searched (137 guarantees that the SSS symbol lies inside the buffer)
  int cp_mode;               // SSS position: 0 normal cyclic prefix, 1 extended, 2 both (the better one is reported)
  const float2* pss_time;    // [3][nfft]
  const float2* pss_freq;    // [3][62]
  const int8_t* sss;         // [3][2][168][62]
  const float2* tw;          // nfft / 2 twiddles of the nfft-point transform
  unsigned long long* peak_key;   // [n_bufs] scratch, zeroed before the launch
  double* power_sum;              // [n_bufs] scratch, zeroed before the launch
  srsue_sync_result* result;      // [n_bufs]
};
__global__ void pss_corr_kernel(const SyncArgs a);
__global__ void sss_detect_kernel(const SyncArgs a);

struct TbArgs {
  const uint8_t* cb_bits;    // [n_sf * C][cb_bits_stride] packed hard bits per code block
  const int32_t* cb_status;  // [n_sf * C]
  uint8_t* payload;          // [n_sf][payload_stride]
  int32_t* tb_status;        // [n_sf][4]: crc_ok, sum of iterations, avg iterations, C
  const int32_t* tbmap;      // [C][4]: source byte offset, payload bytes, TB byte position, bytes per lane
  const uint32_t* tbshift;   // [C][32]: x^(8 * bytes after the lane's chunk) mod CRC24A
  int n_sf, C, Cm, Km, Kp, F, tbs, cb_bits_stride, payload_stride;
};
__global__ void tb_assemble_kernel(const TbArgs a);

}  // namespace srsue

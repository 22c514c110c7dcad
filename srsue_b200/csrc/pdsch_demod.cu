// pdsch_demod.cu -- K3+K4: fused PDSCH RE extraction, ZF/MMSE equaliser (TM1) or Alamouti combiner (TM2),
// QPSK/16QAM/64QAM soft demapper to int16, Gold-sequence descrambler and turbo rate de-matcher with HARQ
// accumulation, writing the soft buffer directly in the decoder's device layout (sm_100a).
//
// Replaces srsLTE's pdsch_get + srslte_predecoding_single/_diversity + srslte_demod_soft_demodulate_s +
// srslte_scrambling_s_offset + srslte_rm_turbo_rx_lut, i.e. the first half of srslte_pdsch_decode_rnti
// (/root/reference/ue/src/phy/phch_worker.cc:347-348; noise_estimate = 0.01 at :340).  Arithmetic
// contract: oracle/SPEC.md sections 4-6.  One CTA per (code block, subframe): the LLRs of the code block
// never leave the SM -- they are produced into shared memory and consumed by a gather over the soft
// buffer, so the de-matching is deterministic (ascending order) without atomics and the only HBM
// traffic is the IQ/estimate read and the coalesced int16 soft-buffer write.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {

__device__ __forceinline__ int16_t q16(float v) {
  // trunc toward zero, NaN -> 0, symmetric saturation (SPEC 5)
  int t = __float2int_rz(v);
  t = max(-32767, min(32767, t));
  return (int16_t)t;
}

template <int QM>
__device__ __forceinline__ void demap(const DemodArgs& a, float2 d, int16_t* out) {
  if (QM == 2) {
    const float s = a.k_sqpsk;                // (float)(100*sqrt(2))
    out[0] = q16(-__fmul_rn(s, d.x));
    out[1] = q16(-__fmul_rn(s, d.y));
  } else if (QM == 4) {
    const float s = 400.0f, c1 = a.k_c16;      // (float)(800/sqrt(10))
    const float tr = __fmul_rn(s, d.x), ti = __fmul_rn(s, d.y);
    out[0] = q16(-tr); out[1] = q16(-ti);
    out[2] = q16(__fsub_rn(fabsf(tr), c1)); out[3] = q16(__fsub_rn(fabsf(ti), c1));
  } else {
    const float s = 700.0f, c1 = a.k_c64a, c2 = a.k_c64b;   // 2800/sqrt(42), 1400/sqrt(42)
    const float tr = __fmul_rn(s, d.x), ti = __fmul_rn(s, d.y);
    const float br = __fsub_rn(fabsf(tr), c1), bi = __fsub_rn(fabsf(ti), c1);
    out[0] = q16(-tr); out[1] = q16(-ti);
    out[2] = q16(br); out[3] = q16(bi);
    out[4] = q16(__fsub_rn(fabsf(br), c2)); out[5] = q16(__fsub_rn(fabsf(bi), c2));
  }
}

__device__ __forceinline__ float dot_rn(float a, float b, float c, float d) { return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }
__device__ __forceinline__ float det_rn(float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }

}  // namespace

template <int QM>
__device__ __forceinline__ void llr_stage(const DemodArgs& a, int sf, int e0, int e1, int16_t* s_e, float n0) {
  const int nsc = a.nsc;
  const float2* y = a.sf_symbols + (size_t)sf * 14 * nsc;
  const float2* h0p = a.ce + (size_t)sf * a.nof_ports * 14 * nsc;
  const int re0 = e0 / QM, re1 = e1 / QM;
  const bool sfbc = (a.tm == 2 && a.nof_ports == 2);
  const int step = sfbc ? 2 : 1;
  for (int i = re0 + step * threadIdx.x; i < re1; i += step * blockDim.x) {
    float2 d[2];
    if (sfbc) {
      const float2* h1p = h0p + 14 * nsc;
      const int g0 = a.re_idx[i], g1 = a.re_idx[i + 1];
      const float2 r0 = y[g0], r1 = y[g1], h0 = h0p[g0], h1 = h1p[g0];
      const float sq2 = a.k_sq2;
      const float den = __fadd_rn(__fadd_rn(dot_rn(h0.x, h0.x, h0.y, h0.y), dot_rn(h1.x, h1.x, h1.y, h1.y)), n0);
      const float a_re = dot_rn(h0.x, r0.x, h0.y, r0.y), a_im = det_rn(h0.x, r0.y, h0.y, r0.x);
      const float b_re = dot_rn(h1.x, r1.x, h1.y, r1.y), b_im = det_rn(h1.y, r1.x, h1.x, r1.y);
      const float c_re = dot_rn(h0.x, r1.x, h0.y, r1.y), c_im = det_rn(h0.x, r1.y, h0.y, r1.x);
      const float e_re = dot_rn(h1.x, r0.x, h1.y, r0.y), e_im = det_rn(h1.y, r0.x, h1.x, r0.y);
      d[0] = make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(a_re, b_re), sq2), den), __fdiv_rn(__fmul_rn(__fadd_rn(a_im, b_im), sq2), den));
      d[1] = make_float2(__fdiv_rn(__fmul_rn(__fsub_rn(c_re, e_re), sq2), den), __fdiv_rn(__fmul_rn(__fsub_rn(c_im, e_im), sq2), den));
    } else {
      const int g0 = a.re_idx[i];
      const float2 r = y[g0], h = h0p[g0];
      const float den = __fadd_rn(dot_rn(h.x, h.x, h.y, h.y), n0);
      d[0] = make_float2(__fdiv_rn(dot_rn(r.x, h.x, r.y, h.y), den), __fdiv_rn(det_rn(r.y, h.x, r.x, h.y), den));
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (q >= step) break;
      const int re = i + q;
      int16_t l[QM];
      demap<QM>(a, d[q], l);
      if (a.dbg_d) a.dbg_d[(size_t)sf * a.nof_re + re] = d[q];
#pragma unroll
      for (int bit = 0; bit < QM; bit++) {
        const int e = re * QM + bit;
        const uint32_t c = (__ldg(a.scramble + (e >> 5)) >> (e & 31)) & 1u;
        const int16_t v = c ? (int16_t)-l[bit] : l[bit];
        s_e[e - e0] = v;
        if (a.dbg_e) a.dbg_e[(size_t)sf * a.nof_re * QM + e] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(512) pdsch_llr_dematch_kernel(const DemodArgs a) {
  extern __shared__ __align__(16) int16_t s_e[];
  const int r = blockIdx.x, sf = blockIdx.y;
  const int e0 = a.cb_e_start[r], e1 = a.cb_e_start[r + 1], E = e1 - e0;
  const float n0 = a.noise_mode ? a.meas[(size_t)sf * 5] : a.noise_est;
  switch (a.qm) {
    case 2: llr_stage<2>(a, sf, e0, e1, s_e, n0); break;
    case 4: llr_stage<4>(a, sf, e0, e1, s_e, n0); break;
    default: llr_stage<6>(a, sf, e0, e1, s_e, n0); break;
  }
  __syncthreads();
  // gather: every soft-buffer element sums the LLRs that the circular buffer maps onto it, in
  // ascending order of the LLR index, saturating at +-C after each addition
  const int cb_elems = a.cb_geom[4 * r], N = a.cb_geom[4 * r + 1];
  const uint16_t* gt = a.gather + (size_t)r * a.gather_stride;
  int16_t* w = a.softbuf + ((size_t)sf * a.C + r) * a.sb_stride;
  for (int m2 = threadIdx.x; m2 < cb_elems / 2; m2 += blockDim.x) {
    const uint32_t gg = __ldg(reinterpret_cast<const uint32_t*>(gt) + m2);
    uint32_t old = a.accumulate ? reinterpret_cast<const uint32_t*>(w)[m2] : 0u;
    uint32_t outw = 0;
#pragma unroll
    for (int hsel = 0; hsel < 2; hsel++) {
      const uint32_t gi = (gg >> (16 * hsel)) & 0xFFFFu;
      int v = (int)(int16_t)((old >> (16 * hsel)) & 0xFFFFu);
      if (gi == 0xFFFEu) {
        v = -kTdC;
      } else if (gi != 0xFFFFu) {
        for (int e = (int)gi; e < E; e += N) v = max(-kTdC, min(kTdC, v + (int)s_e[e]));
      }
      outw |= ((uint32_t)v & 0xFFFFu) << (16 * hsel);
    }
    reinterpret_cast<uint32_t*>(w)[m2] = outw;
  }
}

// ---- K6b: transport-block assembly + CRC24A --------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t gf_mul24(uint32_t x, uint32_t yv, uint32_t poly) {
  uint32_t r = 0;
  for (int i = 23; i >= 0; i--) {
    r <<= 1;
    if (r & 0x1000000u) r ^= poly;
    if ((yv >> i) & 1u) r ^= x;
  }
  return r & 0xFFFFFFu;
}
__device__ __forceinline__ uint32_t gf_xpow(uint32_t e, uint32_t poly) {
  uint32_t result = 1, base = 2;
  while (e) {
    if (e & 1u) result = gf_mul24(result, base, poly);
    base = gf_mul24(base, base, poly);
    e >>= 1;
  }
  return result;
}
}  // namespace

// One CTA per transport block: strips filler and code-block CRCs (all offsets are byte aligned), writes
// the payload MSB-first (what srsUE hands to MAC, phch_worker.cc:305) and checks CRC24A over the
// TBS + 24 bits with a per-thread chunk CRC combined through x^n mod g.
__global__ void __launch_bounds__(256) tb_assemble_kernel(const TbArgs a) {
  __shared__ uint32_t s_crc;
  __shared__ int s_it_sum, s_all_ok;
  const int sf = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { s_crc = 0; s_it_sum = 0; s_all_ok = 1; }
  __syncthreads();
  const int nbytes = (a.tbs + 24) / 8;            // transport block incl. CRC24A
  const int L = (a.C > 1) ? 3 : 0;                // bytes of CRC24B per code block
  // byte b of the TB+CRC stream comes from code block r at byte offset o
  const int per = (nbytes + blockDim.x - 1) / blockDim.x;
  const int b0 = tid * per, b1 = min(nbytes, b0 + per);
  uint32_t crc = 0;
  int r = 0, pos = 0;                              // pos = first stream byte of code block r
  auto cb_payload = [&](int rr) { return (rr < a.Cm ? a.Km : a.Kp) / 8 - L - (rr == 0 ? a.F / 8 : 0); };
  while (r < a.C - 1 && pos + cb_payload(r) <= b0) { pos += cb_payload(r); r++; }
  for (int b = b0; b < b1; b++) {
    while (b - pos >= cb_payload(r)) { pos += cb_payload(r); r++; }
    const int o = (b - pos) + (r == 0 ? a.F / 8 : 0);
    const uint8_t v = a.cb_bits[((size_t)sf * a.C + r) * a.cb_bits_stride + o];
    if (b < a.tbs / 8) a.payload[(size_t)sf * a.payload_stride + b] = v;
    crc ^= (uint32_t)v << 16;
#pragma unroll
    for (int i = 0; i < 8; i++) { crc <<= 1; if (crc & 0x1000000u) crc ^= kCrc24A; }
  }
  if (b1 > b0) {
    const uint32_t sh = gf_xpow((uint32_t)(8 * (nbytes - b1)), kCrc24A);
    atomicXor(&s_crc, gf_mul24(crc & 0xFFFFFFu, sh, kCrc24A));
  }
  for (int rr = tid; rr < a.C; rr += blockDim.x) {
    const int st = a.cb_status[(size_t)sf * a.C + rr];
    atomicAdd(&s_it_sum, st & 0xFF);
    if (!((st >> 8) & 1)) atomicAnd(&s_all_ok, 0);
  }
  __syncthreads();
  if (tid == 0) {
    int32_t* o = a.tb_status + (size_t)sf * 4;
    const int ok = s_all_ok && (a.C == 1 || s_crc == 0);
    o[0] = ok; o[1] = s_it_sum; o[2] = s_it_sum / a.C; o[3] = a.C;
  }
}

}  // namespace srsue

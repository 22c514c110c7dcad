// pdsch_demod.cu -- K3+K4: fused PDSCH RE extraction, ZF/MMSE equaliser (TM1) or Alamouti combiner (TM2),
// QPSK/16QAM/64QAM soft demapper to int16, Gold-sequence descrambler and turbo rate de-matcher with HARQ
// accumulation, writing the soft buffer directly in the decoder's device layout (sm_100a).  Also K6b:
// transport-block assembly + CRC24A.
//
// Replaces srsLTE's pdsch_get + srslte_predecoding_single/_diversity + srslte_demod_soft_demodulate_s +
// srslte_scrambling_s_offset + srslte_rm_turbo_rx_lut, i.e. the first half of srslte_pdsch_decode_rnti
// (/root/reference/ue/src/phy/phch_worker.cc:347-348; noise_estimate = 0.01 at :340).  Arithmetic
// contract: oracle/SPEC.md sections 4-6.  One CTA per (code block, subframe): the LLRs of the code block
// never leave the SM -- they are produced into shared memory and consumed by a gather over the soft
// buffer, so the de-matching is deterministic (ascending order) without atomics and the only HBM
// traffic is the IQ/estimate read and the coalesced 128-bit soft-buffer write.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "viterbi.cuh"

#ifndef SRSUE_DEMOD_MAX_THREADS
#define SRSUE_DEMOD_MAX_THREADS 160
#define SRSUE_DEMOD_MIN_CTAS 10
#endif
#ifndef SRSUE_DEMOD_U2
#define SRSUE_DEMOD_U2 2
#endif

namespace srsue {

namespace {

__device__ __forceinline__ int q16(float v) {
  // trunc toward zero, NaN -> 0, symmetric saturation (SPEC 5)
  return max(-32767, min(32767, __float2int_rz(v)));
}

template <int QM>
__device__ __forceinline__ void demap(const DemodArgs& a, float2 d, int (&out)[6]) {
  if (QM == 2) {
    const float s = a.k_sqpsk;                // (float)(100*sqrt(2))
    out[0] = q16(-__fmul_rn(s, d.x));
    out[1] = q16(-__fmul_rn(s, d.y));
  } else if (QM == 4) {
    const float s = 400.0f, c1 = a.k_c16;     // (float)(800/sqrt(10))
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

// LLRs of one resource element: demap, descramble, store to shared memory (and the optional debug taps)
template <int QM>
__device__ __forceinline__ void emit_re(const DemodArgs& a, int sf, int re, int e0, float2 d, int16_t* s_e) {
  int l[6];
  demap<QM>(a, d, l);
  if (a.dbg_d) a.dbg_d[(size_t)sf * a.nof_re + re] = d;
  const int e = re * QM;
  // QM consecutive scrambling bits starting at bit e (may straddle two words)
  const uint32_t w0 = __ldg(a.scramble + (e >> 5)), w1 = __ldg(a.scramble + (e >> 5) + 1);
  const uint32_t cbits = __funnelshift_r(w0, w1, e & 31);
#pragma unroll
  for (int bit = 0; bit < QM; bit++)
    if ((cbits >> bit) & 1u) l[bit] = -l[bit];
  // e - e0 is even (E_r is a multiple of QM and QM is even): pairs go out as aligned 32-bit stores
  uint32_t* dst = reinterpret_cast<uint32_t*>(s_e + (e - e0));
#pragma unroll
  for (int bit = 0; bit < QM; bit += 2) dst[bit / 2] = ((uint32_t)l[bit] & 0xFFFFu) | ((uint32_t)l[bit + 1] << 16);
  if (a.dbg_e) {
#pragma unroll
    for (int bit = 0; bit < QM; bit++) a.dbg_e[(size_t)sf * a.nof_re * QM + e + bit] = (int16_t)l[bit];
  }
}

__device__ __forceinline__ float lerp_rn(float a, float b, float f) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), f)); }

// Transmit diversity (36.211 6.3.4.3): Alamouti pair number `pair` of a mapped sequence arrives over ports (0, 1) of a
// two-port cell; with four ports even pairs use ports (0, 2) and odd pairs ports (1, 3) (SFBC-FSTD).  `grid` = 14 nsc, the
// distance between the estimates of two ports.
__device__ __forceinline__ void div_ports(const float2* h0p, int np, int pair, int grid, const float2*& ha, const float2*& hb) {
  const int odd = (np == 4) ? (pair & 1) : 0;
  ha = h0p + (size_t)odd * grid;
  hb = h0p + (size_t)(np == 4 ? 2 + odd : 1) * grid;
}

// Channel estimate of port p at grid index g (= l*nsc + k) from the smoothed pilots in shared memory: the
// same frequency and time interpolation, operation for operation, as chest_kernel (SPEC.md 3.3-3.4).
struct ChanInterp {
  const float2* s_pil;     // [ports][4][M]
  const float* s_ftab;     // [17]
  const float* s_ttab;     // [14]
  int nsc, M;
  int c1, c2, c3;          // CRS symbols after symbol 0: 4, 7, 11 (normal cyclic prefix) or 3, 6, 9 (extended)
  int off[2][4];
  __device__ __forceinline__ float2 freq(int p, int si, int k) const {
    const int o = off[p][si];
    int m = (k >= o) ? (k - o) / 6 : 0;
    if (m > M - 2) m = M - 2;
    const float f = s_ftab[k - (6 * m + o) + 5];
    const float2 v0 = s_pil[(p * 4 + si) * M + m], v1 = s_pil[(p * 4 + si) * M + m + 1];
    return make_float2(lerp_rn(v0.x, v1.x, f), lerp_rn(v0.y, v1.y, f));
  }
  __device__ __forceinline__ float2 at(int p, int g) const {
    const int l = g / nsc, k = g - l * nsc;
    if (l == 0) return freq(p, 0, k);
    if (l == c1) return freq(p, 1, k);
    if (l == c2) return freq(p, 2, k);
    if (l == c3) return freq(p, 3, k);
    const int s0 = (l < c1) ? 0 : (l < c2) ? 1 : 2;
    const float2 h0 = freq(p, s0, k), h1 = freq(p, s0 + 1, k);
    const float f = s_ttab[l];
    return make_float2(lerp_rn(h0.x, h1.x, f), lerp_rn(h0.y, h1.y, f));
  }
};

template <int QM, bool FUSED>
__device__ __forceinline__ void llr_stage(const DemodArgs& a, const ChanInterp& ci, int sf, int e0, int e1, int16_t* s_e, float n0) {
  const int nsc = a.nsc;
  const float2* y = a.sf_symbols + (size_t)sf * 14 * nsc;
  const float2* h0p = FUSED ? nullptr : a.ce + (size_t)sf * a.nof_ports * 14 * nsc;
  const int re0 = e0 / QM, re1 = e1 / QM;
  if (a.tm == 2 && a.nof_ports >= 2) {
    const float sq2 = a.k_sq2;
    for (int i = re0 + 2 * threadIdx.x; i < re1; i += 2 * blockDim.x) {
      const int g0 = __ldg(a.re_idx + i), g1 = __ldg(a.re_idx + i + 1);
      const float2 r0 = y[g0], r1 = y[g1];
      const float2 *hap = nullptr, *hbp = nullptr;
      if (!FUSED) div_ports(h0p, a.nof_ports, i >> 1, 14 * nsc, hap, hbp);     // code blocks start at even REs (E_r is a multiple of 2 Qm)
      const float2 h0 = FUSED ? ci.at(0, g0) : hap[g0], h1 = FUSED ? ci.at(1, g0) : hbp[g0];
      const float den = __fadd_rn(__fadd_rn(dot_rn(h0.x, h0.x, h0.y, h0.y), dot_rn(h1.x, h1.x, h1.y, h1.y)), n0);
      const float a_re = dot_rn(h0.x, r0.x, h0.y, r0.y), a_im = det_rn(h0.x, r0.y, h0.y, r0.x);
      const float b_re = dot_rn(h1.x, r1.x, h1.y, r1.y), b_im = det_rn(h1.y, r1.x, h1.x, r1.y);
      const float c_re = dot_rn(h0.x, r1.x, h0.y, r1.y), c_im = det_rn(h0.x, r1.y, h0.y, r1.x);
      const float e_re = dot_rn(h1.x, r0.x, h1.y, r0.y), e_im = det_rn(h1.y, r0.x, h1.x, r0.y);
      const float2 d0 = make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(a_re, b_re), sq2), den), __fdiv_rn(__fmul_rn(__fadd_rn(a_im, b_im), sq2), den));
      const float2 d1 = make_float2(__fdiv_rn(__fmul_rn(__fsub_rn(c_re, e_re), sq2), den), __fdiv_rn(__fmul_rn(__fsub_rn(c_im, e_im), sq2), den));
      emit_re<QM>(a, sf, i, e0, d0, s_e);
      emit_re<QM>(a, sf, i + 1, e0, d1, s_e);
    }
  } else {
    // The kernel is bound by memory latency, not by bandwidth or issue (ncu: 58 % of the stall samples on the long
    // scoreboard at the first use of a load, 88 % occupancy): two REs per trip, all six loads before the first use
    for (int i = re0 + threadIdx.x; i < re1; i += 2 * blockDim.x) {
      const int i2 = i + blockDim.x;
      const bool two = i2 < re1;
      const int g0 = __ldg(a.re_idx + i), g1 = __ldg(a.re_idx + (two ? i2 : i));
      const float2 ra = y[g0], ha = FUSED ? ci.at(0, g0) : h0p[g0];
      const float2 rb = y[g1], hb = FUSED ? ci.at(0, g1) : h0p[g1];
      {
        const float den = __fadd_rn(dot_rn(ha.x, ha.x, ha.y, ha.y), n0);
        const float2 d = make_float2(__fdiv_rn(dot_rn(ra.x, ha.x, ra.y, ha.y), den), __fdiv_rn(det_rn(ra.y, ha.x, ra.x, ha.y), den));
        emit_re<QM>(a, sf, i, e0, d, s_e);
      }
      if (two) {
        const float den = __fadd_rn(dot_rn(hb.x, hb.x, hb.y, hb.y), n0);
        const float2 d = make_float2(__fdiv_rn(dot_rn(rb.x, hb.x, rb.y, hb.y), den), __fdiv_rn(det_rn(rb.y, hb.x, rb.x, hb.y), den));
        emit_re<QM>(a, sf, i2, e0, d, s_e);
      }
    }
  }
}

template <int LIM>
__device__ __forceinline__ uint32_t clamp_pair(uint32_t v) {
  constexpr uint32_t kC = ((uint32_t)LIM << 16) | (uint32_t)LIM;
  constexpr uint32_t kNegC = ((uint32_t)(uint16_t)(-LIM) << 16) | (uint16_t)(-LIM);
  return __vmins2(__vmaxs2(v, kNegC), kC);
}

}  // namespace

__global__ void __launch_bounds__(SRSUE_DEMOD_MAX_THREADS, SRSUE_DEMOD_MIN_CTAS) pdsch_llr_dematch_kernel(const DemodArgs a) {
  extern __shared__ __align__(16) int16_t s_e[];
  const int r = blockIdx.x, sf = blockIdx.y;
  const int e0 = a.cb_e_start[r], e1 = a.cb_e_start[r + 1], E = e1 - e0;
  const float n0 = a.noise_mode ? a.meas[(size_t)sf * 5] : a.noise_est;
  if (threadIdx.x == 0) { s_e[E] = 0; s_e[E + 1] = (int16_t)-kTdC; }      // sentinels of the direct tables
  ChanInterp ci;
  if (a.pilots) {
    // fused channel interpolation: the smoothed pilots of this subframe (6.4 KB per port at 100 PRB) live in
    // shared memory behind the LLR buffer; the full estimate grid never exists in HBM
    __shared__ float s_ftab[17];
    __shared__ float s_ttab[14];
    const int M = 2 * a.nof_prb, np = a.nof_ports;
    float2* s_pil = reinterpret_cast<float2*>(s_e + ((E + 2 + 7) & ~7));
    const float2* src = a.pilots + (size_t)sf * np * 4 * M;
    for (int i = threadIdx.x; i < np * 4 * M; i += blockDim.x) s_pil[i] = src[i];
    if (threadIdx.x < 17) s_ftab[threadIdx.x] = (float)((double)((int)threadIdx.x - 5) / 6.0);
    if (threadIdx.x >= 32 && threadIdx.x < 46) {
      const int l = threadIdx.x - 32;
      const int crs_l[4] = {0, a.cp_ext ? 3 : 4, a.cp_ext ? 6 : 7, a.cp_ext ? 9 : 11};
      const int s0 = (l < crs_l[1]) ? 0 : (l < crs_l[2]) ? 1 : 2;
      s_ttab[l] = (float)((double)(l - crs_l[s0]) / (double)(crs_l[s0 + 1] - crs_l[s0]));
    }
    ci.s_pil = s_pil; ci.s_ftab = s_ftab; ci.s_ttab = s_ttab; ci.nsc = a.nsc; ci.M = M;
    ci.c1 = a.cp_ext ? 3 : 4; ci.c2 = a.cp_ext ? 6 : 7; ci.c3 = a.cp_ext ? 9 : 11;
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int si = 0; si < 4; si++) ci.off[p][si] = a.crs_off[p][si];
    __syncthreads();
    switch (a.qm) {
      case 2: llr_stage<2, true>(a, ci, sf, e0, e1, s_e, n0); break;
      case 4: llr_stage<4, true>(a, ci, sf, e0, e1, s_e, n0); break;
      default: llr_stage<6, true>(a, ci, sf, e0, e1, s_e, n0); break;
    }
  } else {
    switch (a.qm) {
      case 2: llr_stage<2, false>(a, ci, sf, e0, e1, s_e, n0); break;
      case 4: llr_stage<4, false>(a, ci, sf, e0, e1, s_e, n0); break;
      default: llr_stage<6, false>(a, ci, sf, e0, e1, s_e, n0); break;
    }
  }
  __syncthreads();
  const int cb_elems = a.cb_geom[4 * r], N = a.cb_geom[4 * r + 1];
  const uint16_t* gt = a.gather + (size_t)r * a.gather_stride;
  int16_t* w = a.softbuf + ((size_t)sf * a.C + r) * a.sb_stride;
  if (a.direct) {
    // every soft-buffer element takes at most one LLR: branch-free gather, 8 elements (16 bytes) per thread
    const uint4* gt4 = reinterpret_cast<const uint4*>(gt);
    uint4* w4 = reinterpret_cast<uint4*>(w);
    constexpr int U2 = SRSUE_DEMOD_U2;                 // table vectors per thread and trip, all requested before the first use
    const int n8 = cb_elems / 8;
    for (int m8a = threadIdx.x; m8a < n8; m8a += U2 * blockDim.x) {
     uint4 giv[U2], oldv[U2];
#pragma unroll
     for (int u = 0; u < U2; u++) {
       const int mm = m8a + u * blockDim.x;
       giv[u] = __ldg(gt4 + (mm < n8 ? mm : m8a));
       if (a.accumulate) oldv[u] = w4[mm < n8 ? mm : m8a];
     }
#pragma unroll
     for (int half = 0; half < U2; half++) {
      const int m8 = m8a + half * blockDim.x;
      if (m8 >= n8) break;
      const uint4 gi = giv[half];
      const uint32_t g[4] = {gi.x, gi.y, gi.z, gi.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
        o[q] = (uint32_t)(uint16_t)s_e[g[q] & 0xFFFFu] | ((uint32_t)(uint16_t)s_e[g[q] >> 16] << 16);
      if (a.accumulate) {
        const uint4 old = oldv[half];
        // |old| <= C: clamp_C(old + v) == clamp_C(old + clamp_2C(v)), and the packed 16-bit add of values
        // bounded by C and 2C cannot wrap
        o[0] = clamp_pair<kTdC>(__vadd2(old.x, clamp_pair<2 * kTdC>(o[0])));
        o[1] = clamp_pair<kTdC>(__vadd2(old.y, clamp_pair<2 * kTdC>(o[1])));
        o[2] = clamp_pair<kTdC>(__vadd2(old.z, clamp_pair<2 * kTdC>(o[2])));
        o[3] = clamp_pair<kTdC>(__vadd2(old.w, clamp_pair<2 * kTdC>(o[3])));
        // filler positions are forced, not accumulated
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if ((g[q] & 0xFFFFu) == (uint32_t)(E + 1)) o[q] = (o[q] & 0xFFFF0000u) | (uint16_t)(-kTdC);
          if ((g[q] >> 16) == (uint32_t)(E + 1)) o[q] = (o[q] & 0x0000FFFFu) | ((uint32_t)(uint16_t)(-kTdC) << 16);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++) o[q] = clamp_pair<kTdC>(o[q]);
      }
      w4[m8] = make_uint4(o[0], o[1], o[2], o[3]);
     }
    }
  } else {
    // repetition (E > N): every element sums its LLRs in ascending order, saturating after each addition
    for (int m2 = threadIdx.x; m2 < cb_elems / 2; m2 += blockDim.x) {
      const uint32_t gg = __ldg(reinterpret_cast<const uint32_t*>(gt) + m2);
      const uint32_t old = a.accumulate ? reinterpret_cast<const uint32_t*>(w)[m2] : 0u;
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
}

// ---- PCFICH: the CFI srslte_ue_dl_decode_fft_estimate hands back (phch_worker.cc:254).  SPEC.md 9: the 16 symbols of
// OFDM symbol 0 are equalised like PDSCH symbols (MMSE / Alamouti), demapped to int16 QPSK LLRs, descrambled and
// correlated with the three 32-bit code words in integer arithmetic.  One warp per subframe, lane i < 16 owns d(i).
__global__ void __launch_bounds__(128) pcfich_kernel(const PcfichArgs a) {
  const int sf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (sf >= a.n_sf) return;
  const float2* y = a.sf_symbols + (size_t)sf * 14 * a.nsc;
  const float2* h0p = a.ce + (size_t)sf * a.nof_ports * 14 * a.nsc;
  const float n0 = a.noise_mode ? a.meas[(size_t)sf * 5] : a.noise_est;
  int c0 = 0, c1 = 0, c2 = 0;
  if (lane < 16) {
    float2 d;
    if (a.nof_ports >= 2) {
      const float2 *hap, *hbp;
      div_ports(h0p, a.nof_ports, lane >> 1, 14 * a.nsc, hap, hbp);
      const int g0 = a.re[lane & ~1], g1 = a.re[lane | 1];
      const float2 r0 = y[g0], r1 = y[g1], h0 = hap[g0], h1 = hbp[g0];
      const float den = __fadd_rn(__fadd_rn(dot_rn(h0.x, h0.x, h0.y, h0.y), dot_rn(h1.x, h1.x, h1.y, h1.y)), n0);
      if ((lane & 1) == 0) {
        const float a_re = dot_rn(h0.x, r0.x, h0.y, r0.y), a_im = det_rn(h0.x, r0.y, h0.y, r0.x);
        const float b_re = dot_rn(h1.x, r1.x, h1.y, r1.y), b_im = det_rn(h1.y, r1.x, h1.x, r1.y);
        d = make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(a_re, b_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fadd_rn(a_im, b_im), a.k_sq2), den));
      } else {
        const float c_re = dot_rn(h0.x, r1.x, h0.y, r1.y), c_im = det_rn(h0.x, r1.y, h0.y, r1.x);
        const float e_re = dot_rn(h1.x, r0.x, h1.y, r0.y), e_im = det_rn(h1.y, r0.x, h1.x, r0.y);
        d = make_float2(__fdiv_rn(__fmul_rn(__fsub_rn(c_re, e_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fsub_rn(c_im, e_im), a.k_sq2), den));
      }
    } else {
      const int g = a.re[lane];
      const float2 r = y[g], h = h0p[g];
      const float den = __fadd_rn(dot_rn(h.x, h.x, h.y, h.y), n0);
      d = make_float2(__fdiv_rn(dot_rn(r.x, h.x, r.y, h.y), den), __fdiv_rn(det_rn(r.y, h.x, r.x, h.y), den));
    }
    int l[2] = {q16(-__fmul_rn(a.k_sqpsk, d.x)), q16(-__fmul_rn(a.k_sqpsk, d.y))};
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int n = 2 * lane + b;
      if ((a.scramble >> n) & 1u) l[b] = -l[b];
      // code word c has a 0 where n mod 3 == c, a 1 elsewhere; LLR > 0 <=> bit 1
      const int m = n % 3;
      c0 += (m != 0) ? l[b] : -l[b];
      c1 += (m != 1) ? l[b] : -l[b];
      c2 += (m != 2) ? l[b] : -l[b];
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    c0 += __shfl_xor_sync(0xFFFFFFFFu, c0, off);
    c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, off);
    c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, off);
  }
  if (lane == 0) {
    int best = 1, v = c0;
    if (c1 > v) { best = 2; v = c1; }
    if (c2 > v) { best = 3; }
    a.cfi[sf] = best;
    if (a.corr) { a.corr[sf * 3 + 0] = c0; a.corr[sf * 3 + 1] = c1; a.corr[sf * 3 + 2] = c2; }
  }
}

// ---- PBCH / MIB (srslte_ue_mib_decode, phch_recv.cc:247; SPEC.md 12): one CTA of 8 warps per subframe 0.  All threads
// turn the 240 resource elements into 480 LLRs under both transmit-port hypotheses; warp w then decodes hypothesis
// w / 4 (1 or 2 ports) at frame position w % 4 of the 40 ms BCH period: descrambling with that quarter of the sequence,
// exact integer soft combining onto the 120 coded bits, tail-biting Viterbi, CRC16 against the antenna mask.
__global__ void __launch_bounds__(384) pbch_kernel(const PbchArgs a) {
  __shared__ int16_t s_llr[3][480];
  __shared__ int32_t s_soft[12][120];
  __shared__ uint32_t s_surv[12][80][2];
  __shared__ uint8_t s_dec[12][40];
  __shared__ int s_rem[12];
  const int sf = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const float2* y = a.sf_symbols + (size_t)sf * 14 * a.nsc;
  const float2* h0p = a.ce + (size_t)sf * a.nof_ports * 14 * a.nsc;
  const float n0 = a.noise_mode ? a.meas[(size_t)sf * 5] : a.noise_est;
  const int n_hyp = a.nof_ports == 4 ? 3 : a.nof_ports == 2 ? 2 : 1;     // 1, 2 and 4 transmit ports (4 warps each: 256 / 384 threads)
  // hypothesis 1: every thread below 240 equalises one RE; hypothesis 2: every thread below 120 one Alamouti pair
  const int nb = 2 * a.n_re;                     // coded bits per radio frame: 480, or 432 with the extended cyclic prefix
  if (tid < a.n_re) {
    const int g = a.re[tid];
    const float2 r = y[g], h = h0p[g];
    const float den = __fadd_rn(dot_rn(h.x, h.x, h.y, h.y), n0);
    const float2 d = make_float2(__fdiv_rn(dot_rn(r.x, h.x, r.y, h.y), den), __fdiv_rn(det_rn(r.y, h.x, r.x, h.y), den));
    s_llr[0][2 * tid] = (int16_t)q16(-__fmul_rn(a.k_sqpsk, d.x));
    s_llr[0][2 * tid + 1] = (int16_t)q16(-__fmul_rn(a.k_sqpsk, d.y));
  }
  for (int hy = 1; hy < n_hyp; hy++) if (tid < a.n_re / 2) {
    const float2 *hap, *hbp;
    div_ports(h0p, hy == 2 ? 4 : 2, tid, 14 * a.nsc, hap, hbp);
    const int g0 = a.re[2 * tid], g1 = a.re[2 * tid + 1];
    const float2 r0 = y[g0], r1 = y[g1], h0 = hap[g0], h1 = hbp[g0];
    const float den = __fadd_rn(__fadd_rn(dot_rn(h0.x, h0.x, h0.y, h0.y), dot_rn(h1.x, h1.x, h1.y, h1.y)), n0);
    const float a_re = dot_rn(h0.x, r0.x, h0.y, r0.y), a_im = det_rn(h0.x, r0.y, h0.y, r0.x);
    const float b_re = dot_rn(h1.x, r1.x, h1.y, r1.y), b_im = det_rn(h1.y, r1.x, h1.x, r1.y);
    const float c_re = dot_rn(h0.x, r1.x, h0.y, r1.y), c_im = det_rn(h0.x, r1.y, h0.y, r1.x);
    const float e_re = dot_rn(h1.x, r0.x, h1.y, r0.y), e_im = det_rn(h1.y, r0.x, h1.x, r0.y);
    const float2 d0 = make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(a_re, b_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fadd_rn(a_im, b_im), a.k_sq2), den));
    const float2 d1 = make_float2(__fdiv_rn(__fmul_rn(__fsub_rn(c_re, e_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fsub_rn(c_im, e_im), a.k_sq2), den));
    s_llr[hy][4 * tid] = (int16_t)q16(-__fmul_rn(a.k_sqpsk, d0.x));
    s_llr[hy][4 * tid + 1] = (int16_t)q16(-__fmul_rn(a.k_sqpsk, d0.y));
    s_llr[hy][4 * tid + 2] = (int16_t)q16(-__fmul_rn(a.k_sqpsk, d1.x));
    s_llr[hy][4 * tid + 3] = (int16_t)q16(-__fmul_rn(a.k_sqpsk, d1.y));
  }
  __syncthreads();
  const int hyp = w >> 2, q = w & 3;
  if (hyp < n_hyp) {
    int32_t* soft = s_soft[w];
    for (int i = lane; i < 120; i += 32) soft[i] = 0;
    __syncwarp();
    for (int k = lane; k < nb; k += 32) {
      const int bit = nb * q + k;                // place in the block of 4 nb bits: the circular buffer of 120 continues across frames
      const int v = s_llr[hyp][k];
      atomicAdd(&soft[a.rm_seq[bit % 120]], ((a.scramble[bit >> 5] >> (bit & 31)) & 1u) ? -v : v);
    }
    __syncwarp();
    const int rem = viterbi_crc16_warp(soft, 24, &s_surv[w][0][0], s_dec[w], lane);
    if (lane == 0) s_rem[w] = (rem == (hyp == 0 ? 0x0000 : hyp == 1 ? 0xFFFF : 0x5555)) ? 1 : 0;
  } else if (lane == 0) {
    s_rem[w] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    int hit = -1;
    for (int c = 0; c < 4 * n_hyp && hit < 0; c++) if (s_rem[c]) hit = c;
    int32_t* o = a.result + (size_t)sf * 4;
    o[0] = hit >= 0; o[1] = hit >= 0 ? 1 << (hit >> 2) : 0; o[2] = hit >= 0 ? (hit & 3) : 0; o[3] = 0;
    if (hit >= 0) for (int i = 0; i < 24; i++) a.mib[(size_t)sf * 24 + i] = s_dec[hit][i];
  }
}

// ---- PHICH (srslte_ue_dl_decode_phich, phch_worker.cc:381; SPEC.md 11): one thread per subframe equalises the 12 REs
// of the group, removes the orthogonal sequence (conj(w) is a sign swap) and the scrambling sign, and adds re + im of
// the 12 results in index order; ACK iff the sum is negative.
__global__ void __launch_bounds__(128) phich_kernel(const PhichArgs a) {
  const int sf = blockIdx.x * blockDim.x + threadIdx.x;
  if (sf >= a.n_sf) return;
  const float2* y = a.sf_symbols + (size_t)sf * 14 * a.nsc;
  const float2* h0p = a.ce + (size_t)sf * a.nof_ports * 14 * a.nsc;
  const float n0 = a.noise_mode ? a.meas[(size_t)sf * 5] : a.noise_est;
  float metric = 0.0f;
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    float2 d[2];
    if (a.nof_ports >= 2) {
      // four ports (36.211 6.9.2): a whole quadruplet uses ports (0, 2) or (1, 3), alternating with quadruplet + group
      const float2 *hap, *hbp;
      div_ports(h0p, a.nof_ports, (i >> 2) + a.par0, 14 * a.nsc, hap, hbp);
      const float2 r0 = y[a.re[i]], r1 = y[a.re[i + 1]], h0 = hap[a.re[i]], h1 = hbp[a.re[i]];
      const float den = __fadd_rn(__fadd_rn(dot_rn(h0.x, h0.x, h0.y, h0.y), dot_rn(h1.x, h1.x, h1.y, h1.y)), n0);
      const float a_re = dot_rn(h0.x, r0.x, h0.y, r0.y), a_im = det_rn(h0.x, r0.y, h0.y, r0.x);
      const float b_re = dot_rn(h1.x, r1.x, h1.y, r1.y), b_im = det_rn(h1.y, r1.x, h1.x, r1.y);
      const float c_re = dot_rn(h0.x, r1.x, h0.y, r1.y), c_im = det_rn(h0.x, r1.y, h0.y, r1.x);
      const float e_re = dot_rn(h1.x, r0.x, h1.y, r0.y), e_im = det_rn(h1.y, r0.x, h1.x, r0.y);
      d[0] = make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(a_re, b_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fadd_rn(a_im, b_im), a.k_sq2), den));
      d[1] = make_float2(__fdiv_rn(__fmul_rn(__fsub_rn(c_re, e_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fsub_rn(c_im, e_im), a.k_sq2), den));
    } else {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const float2 r = y[a.re[i + j]], h = h0p[a.re[i + j]];
        const float den = __fadd_rn(dot_rn(h.x, h.x, h.y, h.y), n0);
        d[j] = make_float2(__fdiv_rn(dot_rn(r.x, h.x, r.y, h.y), den), __fdiv_rn(det_rn(r.y, h.x, r.x, h.y), den));
      }
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int q = (i + j) & 3;
      float tr, ti;
      bool neg, first;
      if (a.ext) {
        // N_SF = 2 (36.211 Table 6.9.1-2): [1 1], [1 -1], j [1 1], j [1 -1]; six spread symbols, element 2 (i / 4) + q % 2
        if ((q >> 1) != a.odd) continue;
        const int e = 2 * ((i + j) >> 2) + (q & 1);
        if (a.n_seq < 2) { tr = d[j].x; ti = d[j].y; } else { tr = d[j].y; ti = -d[j].x; }
        neg = ((a.n_seq & 1) && (q & 1)) != (((a.scramble >> e) & 1u) != 0);
        first = (i + j) == 2 * a.odd;
      } else {
        // w(q) of sequence n_seq & 3: {1,1,1,1}, {1,-1,1,-1}, {1,1,-1,-1}, {1,-1,-1,1}; sequences 4..7 are j times these
        const int s = a.n_seq & 3;
        const bool neg_w = (s == 1 && (q & 1)) || (s == 2 && (q & 2)) || (s == 3 && (q == 1 || q == 2));
        if (a.n_seq < 4) { tr = d[j].x; ti = d[j].y; } else { tr = d[j].y; ti = -d[j].x; }      // d * conj(j) = (im, -re)
        neg = neg_w != (((a.scramble >> (i + j)) & 1u) != 0);
        first = (i + j) == 0;
      }
      if (neg) { tr = -tr; ti = -ti; }
      const float m = __fadd_rn(tr, ti);
      metric = first ? m : __fadd_rn(metric, m);
    }
  }
  a.ack[sf] = metric < 0.0f;
  if (a.metric) a.metric[sf] = metric;
}

// ---- PDCCH soft bits (srslte_pdcch_extract_llr, phch_worker.cc:260; SPEC.md 10): one thread per resource-element
// group equalises its 4 symbols with the PDSCH formulas, demaps them to 8 int16 QPSK LLRs, descrambles and stores
// them (one 16-byte store) at the quadruplet's place in the de-interleaved PDCCH bit stream.
__global__ void __launch_bounds__(128) pdcch_llr_kernel(const PdcchLlrArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x, sf = blockIdx.y;
  if (a.row_filter && a.row_filter[sf] != a.row_want) return;
  if (m >= a.n_reg) return;
  const float2* y = a.sf_symbols + (size_t)sf * 14 * a.nsc;
  const float2* h0p = a.ce + (size_t)sf * a.nof_ports * 14 * a.nsc;
  const float n0 = a.noise_mode ? a.meas[(size_t)sf * 5] : a.noise_est;
  const int4 g = *reinterpret_cast<const int4*>(a.re4 + 4 * m);
  const int gi[4] = {g.x, g.y, g.z, g.w};
  float2 d[4];
  if (a.nof_ports >= 2) {
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const float2 *hap, *hbp;
      div_ports(h0p, a.nof_ports, i >> 1, 14 * a.nsc, hap, hbp);
      const float2 r0 = y[gi[i]], r1 = y[gi[i + 1]], h0 = hap[gi[i]], h1 = hbp[gi[i]];
      const float den = __fadd_rn(__fadd_rn(dot_rn(h0.x, h0.x, h0.y, h0.y), dot_rn(h1.x, h1.x, h1.y, h1.y)), n0);
      const float a_re = dot_rn(h0.x, r0.x, h0.y, r0.y), a_im = det_rn(h0.x, r0.y, h0.y, r0.x);
      const float b_re = dot_rn(h1.x, r1.x, h1.y, r1.y), b_im = det_rn(h1.y, r1.x, h1.x, r1.y);
      const float c_re = dot_rn(h0.x, r1.x, h0.y, r1.y), c_im = det_rn(h0.x, r1.y, h0.y, r1.x);
      const float e_re = dot_rn(h1.x, r0.x, h1.y, r0.y), e_im = det_rn(h1.y, r0.x, h1.x, r0.y);
      d[i] = make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(a_re, b_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fadd_rn(a_im, b_im), a.k_sq2), den));
      d[i + 1] = make_float2(__fdiv_rn(__fmul_rn(__fsub_rn(c_re, e_re), a.k_sq2), den), __fdiv_rn(__fmul_rn(__fsub_rn(c_im, e_im), a.k_sq2), den));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 r = y[gi[i]], h = h0p[gi[i]];
      const float den = __fadd_rn(dot_rn(h.x, h.x, h.y, h.y), n0);
      d[i] = make_float2(__fdiv_rn(dot_rn(r.x, h.x, r.y, h.y), den), __fdiv_rn(det_rn(r.y, h.x, r.x, h.y), den));
    }
  }
  const int q = a.src[m];
  const uint32_t cb = (a.scramble[q >> 2] >> ((q & 3) * 8)) & 0xFFu;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int l0 = q16(-__fmul_rn(a.k_sqpsk, d[i].x)), l1 = q16(-__fmul_rn(a.k_sqpsk, d[i].y));
    if ((cb >> (2 * i)) & 1u) l0 = -l0;
    if ((cb >> (2 * i + 1)) & 1u) l1 = -l1;
    w[i] = ((uint32_t)l0 & 0xFFFFu) | ((uint32_t)l1 << 16);
  }
  *reinterpret_cast<uint4*>(a.llr + (size_t)sf * 8 * a.n_reg + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- K6b: transport-block assembly + CRC24A --------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t gf_mul24(uint32_t x, uint32_t yv, uint32_t poly) {
  uint32_t r = 0;
#pragma unroll 4
  for (int i = 23; i >= 0; i--) {
    r <<= 1;
    if (r & 0x1000000u) r ^= poly;
    if ((yv >> i) & 1u) r ^= x;
  }
  return r & 0xFFFFFFu;
}
}  // namespace

// One CTA per transport block, one warp per code block: strips filler and code-block CRCs (all offsets
// are byte aligned), writes the payload MSB-first (what srsUE hands to MAC, phch_worker.cc:305) and checks
// CRC24A over the TBS + 24 bits: every lane CRCs a contiguous chunk with a byte table and the chunks
// are combined through the host-computed x^n mod g shifts.
__global__ void __launch_bounds__(1024) tb_assemble_kernel(const TbArgs a) {
  __shared__ uint32_t s_tab[256];
  __shared__ uint32_t s_crc;
  __shared__ int s_it_sum, s_all_ok;
  const int sf = blockIdx.x, tid = threadIdx.x, r = tid >> 5, lane = tid & 31;
  for (int v = tid; v < 256; v += blockDim.x) {
    uint32_t c = (uint32_t)v << 16;
#pragma unroll
    for (int i = 0; i < 8; i++) { c <<= 1; if (c & 0x1000000u) c ^= kCrc24A; }
    s_tab[v] = c & 0xFFFFFFu;
  }
  if (tid == 0) { s_crc = 0; s_it_sum = 0; s_all_ok = 1; }
  __syncthreads();
  if (r < a.C) {
    const int src_off = a.tbmap[4 * r], nb = a.tbmap[4 * r + 1], tb_pos = a.tbmap[4 * r + 2], chunk = a.tbmap[4 * r + 3];
    const uint8_t* __restrict__ src = a.cb_bits + ((size_t)sf * a.C + r) * a.cb_bits_stride + src_off;
    uint8_t* __restrict__ dst = a.payload + (size_t)sf * a.payload_stride;
    const int nb_payload = a.tbs / 8;
    // This lane's contiguous chunk first: all its loads are issued before anything depends on them (a loop that
    // loads and uses one byte per iteration pays one L2 round trip per byte -- 23 in a row at K = 5824).
    constexpr int kMaxChunk = 24;                                  // ceil(6144 / 8 / 32)
    const int b0 = lane * chunk;
    uint8_t v[kMaxChunk];
#pragma unroll
    for (int i = 0; i < kMaxChunk; i++) v[i] = (i < chunk && b0 + i < nb) ? src[b0 + i] : (uint8_t)0;
    // coalesced copy of the payload bytes (the last 3 bytes of the stream are the TB CRC, not payload), 8 loads in flight
#pragma unroll 8
    for (int b = lane; b < nb; b += 32)
      if (tb_pos + b < nb_payload) dst[tb_pos + b] = src[b];
    uint32_t crc = 0;
    const int b1 = min(nb, b0 + chunk);
#pragma unroll
    for (int i = 0; i < kMaxChunk; i++)
      if (i < chunk && b0 + i < nb) crc = ((crc << 8) & 0xFFFFFFu) ^ s_tab[((crc >> 16) ^ v[i]) & 0xFFu];
    uint32_t part = (b1 > b0) ? gf_mul24(crc, a.tbshift[r * 32 + lane], kCrc24A) : 0u;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, off);
    if (lane == 0) {
      atomicXor(&s_crc, part);
      const int st = a.cb_status[(size_t)sf * a.C + r];
      atomicAdd(&s_it_sum, st & 0xFF);
      if (!((st >> 8) & 1)) atomicAnd(&s_all_ok, 0);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int32_t* o = a.tb_status + (size_t)sf * 4;
    const int ok = s_all_ok && (a.C == 1 || s_crc == 0);
    o[0] = ok; o[1] = s_it_sum; o[2] = s_it_sum / a.C; o[3] = a.C;
  }
}

}  // namespace srsue

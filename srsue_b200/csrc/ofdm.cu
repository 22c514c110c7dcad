// ofdm.cu -- K1: batched CP removal + N-point FFT + guard/DC strip + 1/sqrt(N) scaling (sm_100a).
//
// Replaces srsLTE's srslte_ofdm_rx_sf (FFTW) inside srslte_ue_dl_decode_fft_estimate
// (/root/reference/ue/src/phy/phch_worker.cc:254).  Arithmetic contract: oracle/SPEC.md section 2 --
// radix-2 decimation-in-time butterflies t = w*b, a' = a + t, b' = a - t, every float operation rounded
// once (this file is compiled with -fmad=false), twiddles from the shared table.  The schedule is free:
// each thread keeps 8 points in registers and performs three radix-2 stages per pass; passes exchange
// data through skewed shared memory (ping-pong, one barrier per pass).  One CTA transforms one OFDM symbol;
// only the 12*N_PRB used bins are written back, coalesced, already scaled.
//
// The kernel is issue-bound, so all index arithmetic is folded at compile time: every pass is a template
// instance with constant N, L (length of the sub-transforms it combines) and radix, and both the skewed
// shared-memory addresses and the twiddle addresses are "one base register + immediate offsets".
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

#ifndef SRSUE_FFT16_MIN_CTAS
#define SRSUE_FFT16_MIN_CTAS 9     // 56 registers: 0.270 ms per 4096 subframes; 8 CTAs (64 registers) 0.276, 10 (48) 0.282
#endif

namespace srsue {

namespace {

// Packed FP32 (sm_100a FADD2 / FMUL2, PTX add/sub/mul.rn.f32x2): one instruction rounds both halves of a register pair
// exactly like two scalar round-to-nearest operations, so a complex value costs one issue slot instead of two.  The
// kernel is issue-bound, and the operand modifiers of these instructions (scalar broadcast, lane swap, per-lane negate)
// make the shuffles of a complex multiply free.  ptxas contracts mul.f32x2 feeding add.f32x2 into FFMA2 even with
// explicit .rn and -fmad=false (one rounding instead of two -- not the arithmetic of SPEC.md 2), so the two products of
// a complex multiply are packed, their sum/difference stays scalar, and only sums of sums are packed again.
typedef unsigned long long u64_t;
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64_t*>(&r)) : "l"(*reinterpret_cast<u64_t*>(&a)), "l"(*reinterpret_cast<u64_t*>(&b)));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64_t*>(&r)) : "l"(*reinterpret_cast<u64_t*>(&a)), "l"(*reinterpret_cast<u64_t*>(&b)));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64_t*>(&r)) : "l"(*reinterpret_cast<u64_t*>(&a)), "l"(*reinterpret_cast<u64_t*>(&b)));
  return r;
}

// t = w * b with t.re = w.re b.re - w.im b.im, t.im = w.re b.im + w.im b.re, four products and two sums each rounded once
__device__ __forceinline__ float2 cmul(float2 w, float2 b) {
  const float2 p1 = mul2(make_float2(w.x, w.x), b);                       // (w.re b.re, w.re b.im)
  const float2 p2 = mul2(make_float2(w.y, w.y), make_float2(b.y, b.x));   // (w.im b.im, w.im b.re)
  return make_float2(__fsub_rn(p1.x, p2.x), __fadd_rn(p1.y, p2.y));
}
__device__ __forceinline__ void bfly(float2& a, float2& b, float2 w) {
  const float2 t = cmul(w, b);
  const float2 a0 = a;
  a = add2(a0, t);
  b = sub2(a0, t);
}
// butterflies with the exact twiddles 1 and -i (the table holds exactly (1,0) and (0,-1) there, SPEC.md 2):
// w*b is then b resp. (b.im, -b.re) without any rounding, so the products can be skipped
__device__ __forceinline__ void bfly_one(float2& a, float2& b) {
  const float2 a0 = a, b0 = b;
  a = add2(a0, b0);
  b = sub2(a0, b0);
}
__device__ __forceinline__ void bfly_mj(float2& a, float2& b) {
  const float2 a0 = a, t = make_float2(b.y, -b.x);
  a = add2(a0, t);
  b = sub2(a0, t);
}

// Shared-memory skew: element i lives at i + (i >> 4) (one padding float2 per 16), which keeps the strided
// exchanges of the late passes off a single bank group.  The passes below never evaluate this per element;
// they use the closed forms derived in fft_pass.

// The first radix-8 pass (L = 1, k = 0): stage A uses w = 1, stage B w = 1 and -i, stage C w_8^0..3.
template <int N>
__device__ __forceinline__ void combine_first8(float2 (&v)[8], const float2* __restrict__ tw) {
#pragma unroll
  for (int r = 0; r < 4; r++) bfly_one(v[r], v[r + 4]);
  bfly_one(v[0], v[2]); bfly_one(v[1], v[3]);
  bfly_mj(v[4], v[6]); bfly_mj(v[5], v[7]);
  bfly_one(v[0], v[1]);
  bfly_mj(v[2], v[3]);
  bfly(v[4], v[5], __ldg(tw + N / 8));
  bfly(v[6], v[7], __ldg(tw + 3 * (N / 8)));
}

// One combining pass of radix R = 2^LOGR on sub-transforms of length L.  On entry v[r] = F_{n' + (N/L')r}[k];
// on exit register reg_of_u(u) holds F'_{n'}[k + u L] (length L' = R L).  tw is the N/2-entry table of w_N^i.
// Twiddle indices: stage A k N/(2L); stage B (k + hL) N/(4L); stage C (k + hL + q 2L) N/(8L) -- each is a
// multiple of k plus a compile-time constant.
template <int N, int L, int LOGR>
__device__ __forceinline__ void combine(float2 (&v)[8], int k, const float2* __restrict__ tw) {
  constexpr int R = 1 << LOGR;
  {
    const float2 w = __ldg(tw + k * (N / (2 * L)));
#pragma unroll
    for (int r = 0; r < R / 2; r++) bfly(v[r], v[r + R / 2], w);
    // now v[r] = G_r[k], v[r + R/2] = G_r[k + L]
  }
  if (LOGR >= 2) {
    const float2* t2 = tw + k * (N / (4 * L));
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float2 w = __ldg(t2 + h * (N / 4));
#pragma unroll
      for (int r = 0; r < R / 4; r++) bfly(v[h * (R / 2) + r], v[h * (R / 2) + r + R / 4], w);
    }
    // v[h*(R/2) + r] = H_r[k + hL], v[h*(R/2) + r + R/4] = H_r[k + hL + 2L]
  }
  if (LOGR >= 3) {
    const float2* t3 = tw + k * (N / (8 * L));
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int q = 0; q < 2; q++) bfly(v[h * 4 + q * 2], v[h * 4 + q * 2 + 1], __ldg(t3 + h * (N / 8) + q * (N / 4)));
  }
}

// register that holds output u (index k + u L) after combine / combine_first8:
// LOGR == 3: element h*4 + q*2 + e holds u = h + 2q + 4e;  LOGR == 2: element h*2 + e holds u = h + 2e
template <int LOGR>
__host__ __device__ constexpr int reg_of_u(int u) {
  return LOGR == 3 ? ((u & 1) * 4 + ((u >> 1) & 1) * 2 + (u >> 2)) : LOGR == 2 ? ((u & 1) * 2 + (u >> 1)) : u;
}

// One pass for all butterflies of this thread.  Data before a pass sits at index k (N/L) + n (n < N/L), after
// it at k' (N/L') + n'.  With nsub = N/L' and np < nsub:
//   read  idx_r = k (N/L) + np + nsub r   (r < R)
//   write idx_u = (k + u L) nsub + np = (k nsub + np) + u (N/R)
// skew(i) = i + (i >> 4) splits into a base plus a per-r / per-u constant because no addend carries into bit 4:
//   N/R >= 16 is a power of two                          -> skew(idx_u) = skew(k nsub + np) + u (N/R + N/(16R))
//   N/L >= 16: k (N/L) is a multiple of 16, np < nsub    -> (idx_r >> 4) = k (N/L)/16 + (np >> 4) + ((nsub r) >> 4)
//              (np >> 4 is 0 when nsub < 16; for nsub >= 16 nsub r is a multiple of 16)
//   N/L <  16: np + nsub r < N/L, k (N/L) multiple of N/L -> (idx_r >> 4) = (k (N/L)) >> 4
// CFO: the first pass rotates sample i of the symbol by T[((rot_n0 + i) * rot_step) >> 20] as it loads it (SPEC.md 14)
// IQ16: gsrc points at int16 {re, im} pairs (the radio's wire format); a sample is (float)v * iq16_scale, rounded once
template <int N, int L, int LOGR, bool FIRST, bool LAST, bool SYNC_RW = false, bool CFO = false, bool IQ16 = false>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ gsrc, const float2* ssrc, float2* sdst,
                                         float2* __restrict__ gdst, const float2* __restrict__ tw, int tid, int nthreads,
                                         int nsc, float scale, int in_stride = 1, uint32_t rot_n0 = 0, uint32_t rot_step = 0,
                                         const float2* __restrict__ rot_tab = nullptr, float iq16_scale = 0.f) {
  constexpr int R = 1 << LOGR, nsub = N / (L * R), NL = N / L;
  for (int b = tid; b < N / R; b += nthreads) {
    const int k = b / nsub, np = b % nsub;          // powers of two: shift and mask
    float2 v[8];
    if (FIRST) {
#pragma unroll
      for (int r = 0; r < R; r++) {                                                        // L == 1, so k == 0
        if (IQ16) {
          const short2 q = __ldg(reinterpret_cast<const short2*>(gsrc) + (np + nsub * r) * in_stride);
          v[r] = make_float2(__fmul_rn((float)q.x, iq16_scale), __fmul_rn((float)q.y, iq16_scale));
        } else {
          v[r] = __ldg(gsrc + (np + nsub * r) * in_stride);
        }
      }
      if (CFO) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const uint32_t ph = (rot_n0 + (uint32_t)((np + nsub * r) * in_stride)) * rot_step;
          v[r] = cmul(__ldg(rot_tab + (ph >> (32 - kCfoTableLog2))), v[r]);
        }
      }
    } else {
      int base;
      if (NL >= 16) base = k * (NL + NL / 16) + np + ((nsub >= 16) ? (np >> 4) : 0);
      else base = k * NL + ((k * NL) >> 4) + np;
      const float2* p = ssrc + base;
#pragma unroll
      for (int r = 0; r < R; r++) v[r] = p[nsub * r + ((NL >= 16) ? ((nsub * r) >> 4) : 0)];
    }
    if (SYNC_RW) __syncthreads();     // in-place exchange: every thread has read its inputs before anyone overwrites them
    if (FIRST && LOGR == 3) combine_first8<N>(v, tw); else combine<N, L, LOGR>(v, k, tw);
    if (LAST) {
#pragma unroll
      for (int u = 0; u < R; u++) {
        // k + u L is the DFT bin; keep the 12*N_PRB centred bins without DC, scaled by 1/sqrt(N)
        const int kp = k + u * L;
        int ko = -1;
        if (kp >= 1 && kp <= nsc / 2) ko = kp - 1 + nsc / 2;
        else if (kp >= N - nsc / 2) ko = kp - (N - nsc / 2);
        const float2 o = v[reg_of_u<LOGR>(u)];
        if (ko >= 0) gdst[ko] = mul2(o, make_float2(scale, scale));
      }
    } else {
      const int w0 = k * nsub + np;
      float2* p = sdst + w0 + (w0 >> 4);
#pragma unroll
      for (int u = 0; u < R; u++) p[u * (N / R + N / R / 16)] = v[reg_of_u<LOGR>(u)];
    }
  }
}

// all passes of one N-point transform: radix 8 while it fits, then one radix-4 or radix-2 pass.  Two shared
// buffers (ping-pong) avoid the read/write hazard of an in-place exchange, so each pass needs a single barrier.
template <int LOG2N, bool CFO = false, bool IQ16 = false>
__device__ __forceinline__ void fft_symbol(const float2* __restrict__ gin, float2* __restrict__ gout, float2* s0,
                                           float2* s1, const float2* __restrict__ tw, int nsc, float scale,
                                           uint32_t rot_n0 = 0, uint32_t rot_step = 0, const float2* __restrict__ rot_tab = nullptr,
                                           float iq16_scale = 0.f) {
  constexpr int N = 1 << LOG2N;
  static_assert(LOG2N >= 7 && LOG2N <= 11, "LTE transform sizes 128..2048");
  const int tid = threadIdx.x, nt = blockDim.x;
  fft_pass<N, 1, 3, true, false, false, CFO, IQ16>(gin, nullptr, s0, nullptr, tw, tid, nt, nsc, scale, 1, rot_n0, rot_step, rot_tab, iq16_scale);
  __syncthreads();
  fft_pass<N, 8, 3, false, false>(nullptr, s0, s1, nullptr, tw, tid, nt, nsc, scale);
  __syncthreads();
  if constexpr (LOG2N == 7) {             // 8 * 8 * 2
    fft_pass<N, 64, 1, false, true>(nullptr, s1, nullptr, gout, tw, tid, nt, nsc, scale);
  } else if constexpr (LOG2N == 8) {      // 8 * 8 * 4
    fft_pass<N, 64, 2, false, true>(nullptr, s1, nullptr, gout, tw, tid, nt, nsc, scale);
  } else if constexpr (LOG2N == 9) {      // 8 * 8 * 8
    fft_pass<N, 64, 3, false, true>(nullptr, s1, nullptr, gout, tw, tid, nt, nsc, scale);
  } else {                      // 8 * 8 * 8 * (2 | 4)
    fft_pass<N, 64, 3, false, false>(nullptr, s1, s0, nullptr, tw, tid, nt, nsc, scale);
    __syncthreads();
    if constexpr (LOG2N == 10) fft_pass<N, 512, 1, false, true>(nullptr, s0, nullptr, gout, tw, tid, nt, nsc, scale);
    else fft_pass<N, 512, 2, false, true>(nullptr, s0, nullptr, gout, tw, tid, nt, nsc, scale);
  }
}

// N = 1536 (75 PRB) = 3 x 512 (SPEC.md 2): sub-transform r takes the samples 3 i + r through the three radix-8
// passes above (64 threads each, its own pair of shared buffers, result in natural order), then one radix-3 DIT
// stage with the full-circle twiddles w1 = w_1536^k, w2 = w_1536^2k from the table:
//   t1 = w1 f1, t2 = w2 f2, s = t1 + t2, d = t1 - t2, X[k] = f0 + s, m = f0 - s/2,
//   X[k + 512] = m - i c3 d, X[k + 1024] = m + i c3 d, c3 = (float)(sqrt(3)/2), every operation rounded once.
template <bool CFO = false, bool IQ16 = false>
__device__ __forceinline__ void fft1536_symbol(const float2* __restrict__ gin, float2* __restrict__ gout, float2* s_fft,
                                               const float2* __restrict__ tw, int nsc, float scale, float c3,
                                               uint32_t rot_n0 = 0, uint32_t rot_step = 0, const float2* __restrict__ rot_tab = nullptr,
                                               float iq16_scale = 0.f) {
  constexpr int M = 512, SUB = M + M / 16 + 8, N = 1536;
  const int tid = threadIdx.x, r = tid / 64, lt = tid - r * 64;
  if (r < 3) {
    float2* s0 = s_fft + r * SUB;
    float2* s1 = s_fft + (3 + r) * SUB;
    // (IQ16: gin is already the int16-pair address of sample 0, so + r is applied in the element type of the source)
    fft_pass<M, 1, 3, true, false, false, CFO, IQ16>(
        IQ16 ? reinterpret_cast<const float2*>(reinterpret_cast<const short2*>(gin) + r) : gin + r, nullptr, s0, nullptr, tw, lt, 64, nsc,
        scale, 3, rot_n0 + r, rot_step, rot_tab, iq16_scale);
    __syncthreads();
    fft_pass<M, 8, 3, false, false>(nullptr, s0, s1, nullptr, tw, lt, 64, nsc, scale);
    __syncthreads();
    fft_pass<M, 64, 3, false, false>(nullptr, s1, s0, nullptr, tw, lt, 64, nsc, scale);
  } else {
    __syncthreads();
    __syncthreads();
  }
  __syncthreads();
  const float2* w1t = tw + M / 2;
  const float2* w2t = tw + M / 2 + M;
  for (int k = tid; k < M; k += blockDim.x) {
    const int ks = k + (k >> 4);
    const float2 f0 = s_fft[ks], f1 = s_fft[SUB + ks], f2 = s_fft[2 * SUB + ks];
    const float2 t1 = cmul(__ldg(w1t + k), f1), t2 = cmul(__ldg(w2t + k), f2);
    const float2 sm = make_float2(__fadd_rn(t1.x, t2.x), __fadd_rn(t1.y, t2.y));
    const float2 df = make_float2(__fsub_rn(t1.x, t2.x), __fsub_rn(t1.y, t2.y));
    const float2 mm = make_float2(__fsub_rn(f0.x, __fmul_rn(0.5f, sm.x)), __fsub_rn(f0.y, __fmul_rn(0.5f, sm.y)));
    float2 o[3];
    o[0] = make_float2(__fadd_rn(f0.x, sm.x), __fadd_rn(f0.y, sm.y));
    o[1] = make_float2(__fadd_rn(mm.x, __fmul_rn(c3, df.y)), __fsub_rn(mm.y, __fmul_rn(c3, df.x)));
    o[2] = make_float2(__fsub_rn(mm.x, __fmul_rn(c3, df.y)), __fadd_rn(mm.y, __fmul_rn(c3, df.x)));
#pragma unroll
    for (int u = 0; u < 3; u++) {
      const int kp = k + u * M;
      int ko = -1;
      if (kp >= 1 && kp <= nsc / 2) ko = kp - 1 + nsc / 2;
      else if (kp >= N - nsc / 2) ko = kp - (N - nsc / 2);
      if (ko >= 0) gout[ko] = make_float2(__fmul_rn(o[u].x, scale), __fmul_rn(o[u].y, scale));
    }
  }
}

}  // namespace

// Single-buffer variant (N >= 256, exactly N/8 threads, the default): the exchange is done in place with a barrier
// between the reads and the writes of a pass.  Half the shared memory per CTA and 32 registers per thread give 8 CTAs =
// 64 warps per SM instead of 48; the kernel is latency-bound (the first butterflies wait for DRAM, the twiddle loads for
// L1), so the extra warps buy more than the two extra barriers cost: 0.414 -> 0.385 ms per 4096 subframes at N = 2048.
template <int LOG2N, bool IQ16 = false>
__device__ __forceinline__ void fft_symbol_inplace(const float2* __restrict__ gin, float2* __restrict__ gout, float2* s0,
                                                   const float2* __restrict__ tw, int nsc, float scale, float iq16_scale = 0.f) {
  constexpr int N = 1 << LOG2N;
  const int tid = threadIdx.x, nt = blockDim.x;
  fft_pass<N, 1, 3, true, false, false, false, IQ16>(gin, nullptr, s0, nullptr, tw, tid, nt, nsc, scale, 1, 0, 0, nullptr, iq16_scale);
  __syncthreads();
  fft_pass<N, 8, 3, false, false, true>(nullptr, s0, s0, nullptr, tw, tid, nt, nsc, scale);
  __syncthreads();
  if constexpr (LOG2N == 8) {
    fft_pass<N, 64, 2, false, true>(nullptr, s0, nullptr, gout, tw, tid, nt, nsc, scale);
  } else if constexpr (LOG2N == 9) {
    fft_pass<N, 64, 3, false, true>(nullptr, s0, nullptr, gout, tw, tid, nt, nsc, scale);
  } else {
    fft_pass<N, 64, 3, false, false, true>(nullptr, s0, s0, nullptr, tw, tid, nt, nsc, scale);
    __syncthreads();
    if constexpr (LOG2N == 10) fft_pass<N, 512, 1, false, true>(nullptr, s0, nullptr, gout, tw, tid, nt, nsc, scale);
    else fft_pass<N, 512, 2, false, true>(nullptr, s0, nullptr, gout, tw, tid, nt, nsc, scale);
  }
}

// ---- N = 2048 in three passes: radix 16, radix 16, radix 8 -------------------------------------------------------
// The kernels above are bound by the L1/shared-memory pipe (l1tex 90 % busy, DRAM 40 %): three exchanges through shared
// memory and twiddle loads whose lanes hit 4 to 8 different cache lines.  This variant keeps 16 points per thread
// (128 threads per symbol), so the 11 radix-2 stages need only TWO exchanges, and reads its twiddles from tables laid out
// per pass ([twiddle of the pass][k], fft_twiddles appends them for N = 2048) so that the lanes of a warp read
// consecutive entries.  The butterflies, their order inside a stage and the twiddle values are those of SPEC.md 2, so
// the bins are bit-identical to the radix-8 kernels.
//   pass 1 (L = 1,   radix 16): thread b holds x[b + 128 r]; all twiddles are warp-uniform (w_16^j)
//   pass 2 (L = 16,  radix 16): k = b / 8, n' = b % 8; reads k 128 + n' + 8 r, writes b + 128 u
//   pass 3 (L = 256, radix 8, two butterflies per thread: k = b, b + 128): reads 8 k + r, writes bin k + 256 u
// Shared memory: one buffer of 2048 + 127 skewed points, used in place (barrier between the reads and the writes of pass 2).
namespace {

// radix-16 combine of sub-transforms of length L at index k: stages A..D of radix-2 DIT butterflies with the twiddles
// T[j * KS + k]: j = 0 (stage A), 1 + h (B), 3 + h + 2q (C), 7 + h + 2q + 4p (D).  Afterwards element h*8 + q*4 + p*2 + s
// holds output u = h + 2q + 4p + 8s (index k + u L).
template <int KS>
__device__ __forceinline__ void combine16(float2 (&v)[16], const float2* __restrict__ T) {
  {
    const float2 w = __ldg(T);
#pragma unroll
    for (int r = 0; r < 8; r++) bfly(v[r], v[r + 8], w);
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const float2 w = __ldg(T + (1 + h) * KS);
#pragma unroll
    for (int r = 0; r < 4; r++) bfly(v[h * 8 + r], v[h * 8 + r + 4], w);
  }
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const float2 w = __ldg(T + (3 + h + 2 * q) * KS);
#pragma unroll
      for (int r = 0; r < 2; r++) bfly(v[h * 8 + q * 4 + r], v[h * 8 + q * 4 + r + 2], w);
    }
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
      for (int pp = 0; pp < 2; pp++)
        bfly(v[h * 8 + q * 4 + pp * 2], v[h * 8 + q * 4 + pp * 2 + 1], __ldg(T + (7 + h + 2 * q + 4 * pp) * KS));
}

// the first pass (L = 1, k = 0): the same with the exact twiddles 1 and -i taken as such (no products); tw = the N/2 table
template <int N>
__device__ __forceinline__ void combine_first16(float2 (&v)[16], const float2* __restrict__ tw) {
#pragma unroll
  for (int r = 0; r < 8; r++) bfly_one(v[r], v[r + 8]);                                     // A: w = 1
#pragma unroll
  for (int r = 0; r < 4; r++) { bfly_one(v[r], v[r + 4]); bfly_mj(v[8 + r], v[8 + r + 4]); }   // B: 1, -i
  const float2 w8_1 = __ldg(tw + N / 8), w8_3 = __ldg(tw + 3 * (N / 8));
#pragma unroll
  for (int r = 0; r < 2; r++) {                                                              // C: w_8^(h + 2q)
    bfly_one(v[r], v[r + 2]);                    // h = 0, q = 0
    bfly_mj(v[4 + r], v[4 + r + 2]);             // h = 0, q = 1: w_8^2 = -i
    bfly(v[8 + r], v[8 + r + 2], w8_1);          // h = 1, q = 0
    bfly(v[12 + r], v[12 + r + 2], w8_3);        // h = 1, q = 1
  }
  // D: w_16^(h + 2q + 4p) on elements h*8 + q*4 + p*2
  bfly_one(v[0], v[1]);                                             // j = 0
  bfly_mj(v[2], v[3]);                                              // j = 4 (p = 1)
  bfly(v[4], v[5], w8_1);                                           // j = 2 (q = 1): w_16^2 = w_8^1
  bfly(v[6], v[7], w8_3);                                           // j = 6
  bfly(v[8], v[9], __ldg(tw + N / 16));                             // j = 1 (h = 1)
  bfly(v[10], v[11], __ldg(tw + 5 * (N / 16)));                     // j = 5
  bfly(v[12], v[13], __ldg(tw + 3 * (N / 16)));                     // j = 3
  bfly(v[14], v[15], __ldg(tw + 7 * (N / 16)));                     // j = 7
}

__host__ __device__ constexpr int reg_of_u16(int u) { return (u & 1) * 8 + ((u >> 1) & 1) * 4 + ((u >> 2) & 1) * 2 + (u >> 3); }

template <bool IQ16>
__device__ __forceinline__ void fft2048_r16(const float2* __restrict__ gin, float2* __restrict__ gout, float2* s0,
                                            const float2* __restrict__ tw, int nsc, float scale, float iq16_scale) {
  constexpr int N = 2048;
  const int b = threadIdx.x;                        // 128 threads
  const float2* T2 = tw + N / 2;                    // [15][16]
  const float2* T3 = tw + N / 2 + 15 * 16;          // [7][256]
  float2 v[16];
  // ---- pass 1
#pragma unroll
  for (int r = 0; r < 16; r++) {
    if (IQ16) {
      const short2 q = __ldg(reinterpret_cast<const short2*>(gin) + b + 128 * r);
      v[r] = make_float2(__fmul_rn((float)q.x, iq16_scale), __fmul_rn((float)q.y, iq16_scale));
    } else {
      v[r] = __ldg(gin + b + 128 * r);
    }
  }
  combine_first16<N>(v, tw);
  {
    float2* p = s0 + b + (b >> 4);
#pragma unroll
    for (int u = 0; u < 16; u++) p[136 * u] = v[reg_of_u16(u)];
  }
  __syncthreads();
  // ---- pass 2
  {
    const int k = b >> 3, np = b & 7;
    const float2* p = s0 + 136 * k + np;
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = p[8 * r + (r >> 1)];
    __syncthreads();
    combine16<16>(v, T2 + k);
    float2* q = s0 + b + (b >> 4);
#pragma unroll
    for (int u = 0; u < 16; u++) q[136 * u] = v[reg_of_u16(u)];
  }
  __syncthreads();
  // ---- pass 3: two radix-8 butterflies, k = b and b + 128
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int k = b + 128 * half;
    const float2* p = s0 + 8 * k + (k >> 1);
#pragma unroll
    for (int r = 0; r < 8; r++) v[half * 8 + r] = p[r];
  }
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int k = b + 128 * half;
    const float2* T = T3 + k;
    float2* w = v + half * 8;
    {
      const float2 t = __ldg(T);
#pragma unroll
      for (int r = 0; r < 4; r++) bfly(w[r], w[r + 4], t);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float2 t = __ldg(T + (1 + h) * 256);
#pragma unroll
      for (int r = 0; r < 2; r++) bfly(w[h * 4 + r], w[h * 4 + r + 2], t);
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int q = 0; q < 2; q++) bfly(w[h * 4 + q * 2], w[h * 4 + q * 2 + 1], __ldg(T + (3 + h + 2 * q) * 256));
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int kp = k + u * 256;
      int ko = -1;
      if (kp >= 1 && kp <= nsc / 2) ko = kp - 1 + nsc / 2;
      else if (kp >= N - nsc / 2) ko = kp - (N - nsc / 2);
      if (ko >= 0) gout[ko] = mul2(w[reg_of_u<3>(u)], make_float2(scale, scale));
    }
  }
}

// sample offset of symbol l inside its subframe: the cyclic prefixes of symbols 0..l plus l full symbols.  Normal prefix:
// 160 samples (at N = 2048) for the first symbol of a slot, 144 for the six others; extended: 512 for each of the six
__device__ __forceinline__ int symbol_start(int N, int l, int cp_ext) {
  if (cp_ext) return (l + 1) * (N / 4) + l * N;
  const int cp0 = 160 * N / 2048, cp1 = 144 * N / 2048;
  const int slot = l / 7, ls = l % 7;
  return slot * (7 * N + cp0 + 6 * cp1) + ls * N + cp0 + ls * cp1;
}
}  // namespace

__global__ void __launch_bounds__(128, SRSUE_FFT16_MIN_CTAS) ofdm_rx_r16_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  constexpr int N = 2048;
  const int start = symbol_start(N, l, a.cp_ext);
  fft2048_r16<false>(a.iq + (size_t)sf * 15 * N + start, a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc, s_fft, a.tw, a.nsc, a.scale, 0.f);
}
__global__ void __launch_bounds__(128, SRSUE_FFT16_MIN_CTAS) ofdm_rx_r16_iq16_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  constexpr int N = 2048;
  const int start = symbol_start(N, l, a.cp_ext);
  fft2048_r16<true>(reinterpret_cast<const float2*>(a.iq16 + (size_t)sf * 15 * N + start), a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc, s_fft,
                    a.tw, a.nsc, a.scale, a.iq16_scale);
}

__global__ void __launch_bounds__(256, 7) ofdm_rx_inplace_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  const int start = symbol_start(N, l, a.cp_ext);
  const float2* gin = a.iq + (size_t)sf * 15 * N + start;
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  switch (a.log2n) {
    case 8: fft_symbol_inplace<8>(gin, gout, s_fft, a.tw, a.nsc, a.scale); break;
    case 9: fft_symbol_inplace<9>(gin, gout, s_fft, a.tw, a.nsc, a.scale); break;
    case 10: fft_symbol_inplace<10>(gin, gout, s_fft, a.tw, a.nsc, a.scale); break;
    case 11: fft_symbol_inplace<11>(gin, gout, s_fft, a.tw, a.nsc, a.scale); break;
    default: break;
  }
}

__global__ void __launch_bounds__(256) ofdm_rx_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  const int start = symbol_start(N, l, a.cp_ext);
  const float2* gin = a.iq + (size_t)sf * 15 * N + start;
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  float2* s0 = s_fft;
  float2* s1 = s_fft + (N + N / 16 + 8);
  if (N == 1536) { fft1536_symbol(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.c3); return; }
  switch (a.log2n) {
    case 7: fft_symbol<7>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 8: fft_symbol<8>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 9: fft_symbol<9>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 10: fft_symbol<10>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 11: fft_symbol<11>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    default: break;
  }
}

// int16 {re, im} input (what the radio puts on the wire and what capture files usually hold): half the bytes over PCIe
// and out of HBM; the conversion (float)v * scale rides on the loads of the first pass.
__global__ void __launch_bounds__(256, 7) ofdm_rx_inplace_iq16_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  const int start = symbol_start(N, l, a.cp_ext);
  const float2* gin = reinterpret_cast<const float2*>(a.iq16 + (size_t)sf * 15 * N + start);
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  switch (a.log2n) {
    case 8: fft_symbol_inplace<8, true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.iq16_scale); break;
    case 9: fft_symbol_inplace<9, true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.iq16_scale); break;
    case 10: fft_symbol_inplace<10, true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.iq16_scale); break;
    case 11: fft_symbol_inplace<11, true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.iq16_scale); break;
    default: break;
  }
}

__global__ void __launch_bounds__(256) ofdm_rx_iq16_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  const int start = symbol_start(N, l, a.cp_ext);
  const float2* gin = reinterpret_cast<const float2*>(a.iq16 + (size_t)sf * 15 * N + start);
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  float2* s0 = s_fft;
  float2* s1 = s_fft + (N + N / 16 + 8);
  if (N == 1536) { fft1536_symbol<false, true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.c3, 0, 0, nullptr, a.iq16_scale); return; }
  switch (a.log2n) {
    case 7: fft_symbol<7, false, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, 0, 0, nullptr, a.iq16_scale); break;
    case 8: fft_symbol<8, false, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, 0, 0, nullptr, a.iq16_scale); break;
    case 9: fft_symbol<9, false, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, 0, 0, nullptr, a.iq16_scale); break;
    case 10: fft_symbol<10, false, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, 0, 0, nullptr, a.iq16_scale); break;
    case 11: fft_symbol<11, false, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, 0, 0, nullptr, a.iq16_scale); break;
    default: break;
  }
}

// int16 input and the carrier-offset rotation together (a live radio's case): convert, then rotate, both on the loads.
__global__ void __launch_bounds__(256) ofdm_rx_cfo_iq16_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  const int start = symbol_start(N, l, a.cp_ext);
  const float2* gin = reinterpret_cast<const float2*>(a.iq16 + (size_t)sf * 15 * N + start);
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  float2* s0 = s_fft;
  float2* s1 = s_fft + (N + N / 16 + 8);
  const uint32_t step = (uint32_t)(a.cfo_steps ? __ldg(a.cfo_steps + sf) : a.cfo_step);
  if (N == 1536) { fft1536_symbol<true, true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.c3, (uint32_t)start, step, a.cexp, a.iq16_scale); return; }
  switch (a.log2n) {
    case 7: fft_symbol<7, true, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp, a.iq16_scale); break;
    case 8: fft_symbol<8, true, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp, a.iq16_scale); break;
    case 9: fft_symbol<9, true, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp, a.iq16_scale); break;
    case 10: fft_symbol<10, true, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp, a.iq16_scale); break;
    case 11: fft_symbol<11, true, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp, a.iq16_scale); break;
    default: break;
  }
}

// The two-buffer kernel with the carrier-offset rotation fused into the sample loads (all transform sizes).
__global__ void __launch_bounds__(256) ofdm_rx_cfo_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  const int start = symbol_start(N, l, a.cp_ext);
  const float2* gin = a.iq + (size_t)sf * 15 * N + start;
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  float2* s0 = s_fft;
  float2* s1 = s_fft + (N + N / 16 + 8);
  const uint32_t step = (uint32_t)(a.cfo_steps ? __ldg(a.cfo_steps + sf) : a.cfo_step);
  if (N == 1536) { fft1536_symbol<true>(gin, gout, s_fft, a.tw, a.nsc, a.scale, a.c3, (uint32_t)start, step, a.cexp); return; }
  switch (a.log2n) {
    case 7: fft_symbol<7, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp); break;
    case 8: fft_symbol<8, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp); break;
    case 9: fft_symbol<9, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp); break;
    case 10: fft_symbol<10, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp); break;
    case 11: fft_symbol<11, true>(gin, gout, s0, s1, a.tw, a.nsc, a.scale, (uint32_t)start, step, a.cexp); break;
    default: break;
  }
}

}  // namespace srsue

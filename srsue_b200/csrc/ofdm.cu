// ofdm.cu -- K1: batched CP removal + N-point FFT + guard/DC strip + 1/sqrt(N) scaling (sm_100a).
//
// Replaces srsLTE's srslte_ofdm_rx_sf (FFTW) inside srslte_ue_dl_decode_fft_estimate
// (/root/reference/ue/src/phy/phch_worker.cc:254).  Arithmetic contract: oracle/SPEC.md section 2 --
// radix-2 decimation-in-time butterflies t = w*b, a' = a + t, b' = a - t, every float operation rounded
// once (this file is compiled with -fmad=false), twiddles from the shared table.  The schedule is free:
// each thread keeps 8 points in registers and performs three radix-2 stages per pass; passes exchange
// data through (skewed) shared memory.  One CTA transforms one OFDM symbol; only the 12*N_PRB used bins
// are written back, coalesced, already scaled.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {

__device__ __forceinline__ float2 cmul(float2 w, float2 b) {
  return make_float2(__fsub_rn(__fmul_rn(w.x, b.x), __fmul_rn(w.y, b.y)),
                     __fadd_rn(__fmul_rn(w.x, b.y), __fmul_rn(w.y, b.x)));
}
__device__ __forceinline__ void bfly(float2& a, float2& b, float2 w) {
  const float2 t = cmul(w, b);
  const float2 a0 = a;
  a = make_float2(__fadd_rn(a0.x, t.x), __fadd_rn(a0.y, t.y));
  b = make_float2(__fsub_rn(a0.x, t.x), __fsub_rn(a0.y, t.y));
}
// butterflies with the exact twiddles 1 and -i (the table holds exactly (1,0) and (0,-1) there, SPEC.md 2):
// w*b is then b resp. (b.im, -b.re) without any rounding, so the products can be skipped
__device__ __forceinline__ void bfly_one(float2& a, float2& b) {
  const float2 a0 = a, b0 = b;
  a = make_float2(__fadd_rn(a0.x, b0.x), __fadd_rn(a0.y, b0.y));
  b = make_float2(__fsub_rn(a0.x, b0.x), __fsub_rn(a0.y, b0.y));
}
__device__ __forceinline__ void bfly_mj(float2& a, float2& b) {
  const float2 a0 = a, b0 = b;
  a = make_float2(__fadd_rn(a0.x, b0.y), __fsub_rn(a0.y, b0.x));
  b = make_float2(__fsub_rn(a0.x, b0.y), __fadd_rn(a0.y, b0.x));
}
// skew: one padding element per 16 float2 keeps strided exchanges off a single bank group
__device__ __forceinline__ int skew(int i) { return i + (i >> 4); }

// One combining pass of radix R = 2^LOGR.  On entry v[r] = F_{n' + (N/L')r}[k] (sub-transforms of length
// L); on exit v[u] = F'_{n'}[k + u L] (length L' = R L).  tw is the N/2-entry table of w_N^i.
template <int LOGR>
__device__ __forceinline__ void combine(float2 (&v)[8], int k, int L, int N, const float2* __restrict__ tw) {
  constexpr int R = 1 << LOGR;
  // stage A: pairs (r, r + R/2), twiddle w_{2L}^k
  {
    const float2 w = tw[k * (N / (2 * L))];
#pragma unroll
    for (int r = 0; r < R / 2; r++) bfly(v[r], v[r + R / 2], w);
    // now v[r] = G_r[k], v[r + R/2] = G_r[k + L]
  }
  if (LOGR >= 2) {
    // stage B: pairs (G_r, G_{r + R/4}) for k2 in {k, k + L}, twiddle w_{4L}^{k2}
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float2 w = tw[(k + h * L) * (N / (4 * L))];
#pragma unroll
      for (int r = 0; r < R / 4; r++) bfly(v[h * (R / 2) + r], v[h * (R / 2) + r + R / 4], w);
    }
    // v[h*(R/2) + r] = H_r[k + hL], v[h*(R/2) + r + R/4] = H_r[k + hL + 2L]
  }
  if (LOGR >= 3) {
    // stage C: pairs (H_0, H_1) for k4 in {k, k+L, k+2L, k+3L}, twiddle w_{8L}^{k4}
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int k4 = k + h * L + q * 2 * L;
        const float2 w = tw[k4 * (N / (8 * L))];
        bfly(v[h * 4 + q * 2], v[h * 4 + q * 2 + 1], w);
      }
  }
  // gather outputs in order u = 0..R-1 (output index k + uL)
  float2 o[8];
  if (LOGR == 3) {
    // element h*4 + q*2 + e holds index k + hL + q 2L + e 4L  ->  u = h + 2q + 4e
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int e = 0; e < 2; e++) o[h + 2 * q + 4 * e] = v[h * 4 + q * 2 + e];
  } else if (LOGR == 2) {
    // element h*2 + e holds k + hL + e 2L -> u = h + 2e
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int e = 0; e < 2; e++) o[h + 2 * e] = v[h * 2 + e];
  } else {
    o[0] = v[0]; o[1] = v[1];
  }
#pragma unroll
  for (int u = 0; u < R; u++) v[u] = o[u];
}

// The first radix-8 pass (L = 1, k = 0): stage A uses w = 1, stage B w = 1 and -i, stage C w_8^0..3.
__device__ __forceinline__ void combine_first8(float2 (&v)[8], int N, const float2* __restrict__ tw) {
#pragma unroll
  for (int r = 0; r < 4; r++) bfly_one(v[r], v[r + 4]);
  bfly_one(v[0], v[2]); bfly_one(v[1], v[3]);
  bfly_mj(v[4], v[6]); bfly_mj(v[5], v[7]);
  bfly_one(v[0], v[1]);
  bfly_mj(v[2], v[3]);
  bfly(v[4], v[5], tw[N / 8]);
  bfly(v[6], v[7], tw[3 * (N / 8)]);
  float2 o[8];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
      for (int e = 0; e < 2; e++) o[h + 2 * q + 4 * e] = v[h * 4 + q * 2 + e];
#pragma unroll
  for (int u = 0; u < 8; u++) v[u] = o[u];
}

// First pass (global -> shared) or last pass (shared -> global, bin selection + scaling) for all
// butterflies of this thread.  Data before a pass sits at index k*(N/L) + n, after it at k'*(N/L') + n'.
template <int LOGR, bool FIRST, bool LAST>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ gsrc, float2* sbuf, int N, int L,
                                         const float2* __restrict__ tw, int tid, int nthreads, float2* __restrict__ gdst,
                                         int nsc, float scale) {
  constexpr int R = 1 << LOGR;
  const int Lp = L * R, nsub = N / Lp;       // nsub = number of length-L' transforms
  for (int b = tid; b < N / R; b += nthreads) {
    const int k = b / nsub, np = b - k * nsub;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int idx = k * (N / L) + np + nsub * r;
      v[r] = FIRST ? __ldg(gsrc + idx) : sbuf[skew(idx)];
    }
    if (FIRST && LOGR == 3) combine_first8(v, N, tw); else combine<LOGR>(v, k, L, N, tw);
#pragma unroll
    for (int u = 0; u < R; u++) {
      const int kp = k + u * L;
      if (LAST) {
        // kp is the DFT bin; keep the 12*N_PRB centred bins without DC, scaled by 1/sqrt(N)
        int ko = -1;
        if (kp >= 1 && kp <= nsc / 2) ko = kp - 1 + nsc / 2;
        else if (kp >= N - nsc / 2) ko = kp - (N - nsc / 2);
        if (ko >= 0) gdst[ko] = make_float2(__fmul_rn(v[u].x, scale), __fmul_rn(v[u].y, scale));
      } else {
        sbuf[skew(kp * nsub + np)] = v[u];
      }
    }
  }
}

}  // namespace

// Two shared buffers (ping-pong) avoid the read/write hazard of an in-place exchange, so each pass needs
// a single barrier.  N <= 2048 -> 2 * 17 KB.
template <int LOG2N>
__device__ __forceinline__ void fft_symbol(const float2* __restrict__ gin, float2* __restrict__ gout, float2* s0,
                                           float2* s1, const float2* __restrict__ tw, int nsc, float scale) {
  constexpr int N = 1 << LOG2N;
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int NP3 = LOG2N / 3, REM = LOG2N % 3;       // NP3 radix-8 passes, then a radix-2^REM pass
  int L = 1;
  float2* cur = s0; float2* nxt = s1;
  // first pass reads global memory (coalesced: consecutive threads read consecutive samples)
  if (NP3 == 1 && REM == 0) { fft_pass<3, true, true>(gin, nullptr, N, L, tw, tid, nt, gout, nsc, scale); return; }
  fft_pass<3, true, false>(gin, cur, N, L, tw, tid, nt, nullptr, nsc, scale);
  L *= 8;
  __syncthreads();
#pragma unroll
  for (int p = 1; p < NP3; p++) {
    const bool last = (p == NP3 - 1) && (REM == 0);
    if (last) {
      fft_pass<3, false, true>(nullptr, cur, N, L, tw, tid, nt, gout, nsc, scale);
      return;
    }
    // read from cur, write to nxt
    {
      constexpr int R = 8;
      const int Lp = L * R, nsub = N / Lp;
      for (int b = tid; b < N / R; b += nt) {
        const int k = b / nsub, np = b - k * nsub;
        float2 v[8];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = cur[skew(k * (N / L) + np + nsub * r)];
        combine<3>(v, k, L, N, tw);
#pragma unroll
        for (int u = 0; u < R; u++) nxt[skew((k + u * L) * nsub + np)] = v[u];
      }
    }
    L *= 8;
    __syncthreads();
    float2* t = cur; cur = nxt; nxt = t;
  }
  if (REM == 2) fft_pass<2, false, true>(nullptr, cur, N, L, tw, tid, nt, gout, nsc, scale);
  if (REM == 1) fft_pass<1, false, true>(nullptr, cur, N, L, tw, tid, nt, gout, nsc, scale);
}

__global__ void __launch_bounds__(256) ofdm_rx_kernel(const OfdmArgs a) {
  extern __shared__ __align__(16) float2 s_fft[];
  const int l = blockIdx.x, sf = blockIdx.y;
  const int N = a.nfft;
  // sample offset of symbol l: CPs of symbols 0..l plus l full symbols
  const int cp0 = 160 * N / 2048, cp1 = 144 * N / 2048;
  const int slot = l / 7, ls = l % 7;
  const int start = slot * (7 * N + cp0 + 6 * cp1) + ls * N + cp0 + ls * cp1;
  const float2* gin = a.iq + (size_t)sf * 15 * N + start;
  float2* gout = a.sf_symbols + ((size_t)sf * 14 + l) * a.nsc;
  float2* s0 = s_fft;
  float2* s1 = s_fft + (N + N / 16 + 8);
  switch (a.log2n) {
    case 7: fft_symbol<7>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 8: fft_symbol<8>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 9: fft_symbol<9>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 10: fft_symbol<10>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    case 11: fft_symbol<11>(gin, gout, s0, s1, a.tw, a.nsc, a.scale); break;
    default: break;
  }
}

}  // namespace srsue

// chest.cu -- K2: CRS least-squares channel estimate, 3-tap smoothing, frequency + time linear
// interpolation, and the noise / RSRP / RSSI / RSRQ / SNR measurements (sm_100a).
//
// Replaces srsLTE's srslte_chest_dl_estimate inside srslte_ue_dl_decode_fft_estimate
// (/root/reference/ue/src/phy/phch_worker.cc:254) and feeds the getters srsUE reads afterwards
// (phch_worker.cc:359,799,821-823,842,847).  Arithmetic contract: oracle/SPEC.md section 3; every float
// operation is an explicit round-to-nearest intrinsic so that nothing is contracted into an FMA.
// One CTA per subframe: pilots and their smoothed values live in shared memory, each thread then owns
// subcarriers k = tid, tid + blockDim, ... and writes the 14 interpolated symbols coalesced.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {
__device__ __forceinline__ float lerp_rn(float a, float b, float f) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), f)); }
__device__ __forceinline__ float abs2_rn(float2 z) { return __fadd_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y)); }

// SPEC 3.5 reduction order: lane l accumulates elements l, l+32, ... in ascending order, then xor-tree
template <typename F>
__device__ __forceinline__ float warp_sum_ordered(int n, int lane, F elem) {
  float acc = 0.0f;
  for (int i = lane; i < n; i += 32) acc = __fadd_rn(acc, elem(i));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, off));
  return acc;
}
}  // namespace

// EXT = extended cyclic prefix: 12 symbols, CRS in symbols 0, 3, 6, 9 (36.211 6.10.1.2: symbols 0 and N_symb - 3 of a slot),
// symbols 10 and 11 extrapolated from (6, 9) as 12 and 13 are from (7, 11); same operations otherwise (SPEC.md 15b)
// P4 = four-port cell: ports 2 / 3 carry pilots in symbol 1 of each slot only (36.211 6.10.1.2), so they have two pilot
// symbols (rows 0, 1 of their [4][M] block; sign rows 4, 5) and one time segment (1, NSLOT + 1) that every symbol
// interpolates or extrapolates from (SPEC.md 15c).  Ports 0 / 1 are computed exactly as in a two-port cell.
template <bool EXT, bool P4>
__device__ __forceinline__ void chest_body(const ChestArgs& a) {
  extern __shared__ __align__(16) float2 s_ch[];
  __shared__ float s_ftab[17];
  __shared__ float s_ttab[14];
  __shared__ float s_ttab2[14];           // ports 2 / 3: (l - 1) / NSLOT
  __shared__ float s_red[3];
  const int sf = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nsc = a.nsc, M = 2 * a.nof_prb, np = a.nof_ports;
  float2* s_ls = s_ch;                    // [np][4][M]
  float2* s_sm = s_ch + np * 4 * M;       // [np][4][M]
  float* s_pw = reinterpret_cast<float*>(s_ch + 2 * np * 4 * M);   // [4][nsc]: |y|^2 of the four CRS symbols (RSSI)
  const float2* y = a.sf_symbols + (size_t)sf * 14 * nsc;
  constexpr int NSYM = EXT ? 12 : 14, C1 = EXT ? 3 : 4, C2 = EXT ? 6 : 7, C3 = EXT ? 9 : 11, NSLOT = EXT ? 6 : 7;
  const int crs_l[4] = {0, C1, C2, C3};
  const float isq2 = (float)(1.0 / sqrt(2.0));

  if (tid < 17) s_ftab[tid] = (float)((double)(tid - 5) / 6.0);
  if (tid >= 32 && tid < 32 + NSYM) {
    const int l = tid - 32;
    const int s0 = (l < C1) ? 0 : (l < C2) ? 1 : 2;
    s_ttab[l] = (float)((double)(l - crs_l[s0]) / (double)(crs_l[s0 + 1] - crs_l[s0]));
    if (P4) s_ttab2[l] = (float)((double)(l - 1) / (double)NSLOT);
  }
  // ---- RSSI inputs: the ordered sum below is one warp walking 4 nsc values (SPEC 3.5 fixes its order), which as a chain
  // of dependent global loads kept every CTA resident for tens of microseconds; here all threads fetch the values at once
  for (int si = 0; si < 4; si++)
    for (int k = tid; k < nsc; k += nt) s_pw[si * nsc + k] = abs2_rn(y[crs_l[si] * nsc + k]);
  // ---- least squares at the pilots ------------------------------------------------------------------
  for (int i = tid; i < np * 4 * M; i += nt) {
    const int m = i % M, si = (i / M) % 4, p = i / (4 * M);
    if (P4 && p >= 2 && si >= 2) { s_ls[i] = make_float2(0.f, 0.f); continue; }      // ports 2 / 3 have two pilot symbols
    const int lsym = (P4 && p >= 2) ? (si ? NSLOT + 1 : 1) : crs_l[si], srow = (P4 && p >= 2) ? 4 + si : si;
    const float2 v = y[lsym * nsc + a.crs_off[p][si] + 6 * m];
    const int rs = a.crs_sign[(srow * 2 + 0) * M + m], is = a.crs_sign[(srow * 2 + 1) * M + m];
    const float tre = __fadd_rn(rs > 0 ? v.x : -v.x, is > 0 ? v.y : -v.y);
    const float tim = __fsub_rn(rs > 0 ? v.y : -v.y, is > 0 ? v.x : -v.x);
    s_ls[i] = make_float2(__fmul_rn(tre, isq2), __fmul_rn(tim, isq2));
  }
  __syncthreads();
  // ---- 3-tap smoothing (edges unchanged) ------------------------------------------------------------
  for (int i = tid; i < np * 4 * M; i += nt) {
    const int m = i % M;
    float2 o = s_ls[i];
    if (m > 0 && m < M - 1) {
      const float2 l = s_ls[i - 1], c = s_ls[i], r = s_ls[i + 1];
      o.x = __fadd_rn(__fadd_rn(__fmul_rn(0.1f, l.x), __fmul_rn(0.8f, c.x)), __fmul_rn(0.1f, r.x));
      o.y = __fadd_rn(__fadd_rn(__fmul_rn(0.1f, l.y), __fmul_rn(0.8f, c.y)), __fmul_rn(0.1f, r.y));
    }
    s_sm[i] = o;
    if (a.pilots) a.pilots[(size_t)sf * np * 4 * M + i] = o;
  }
  __syncthreads();
  // ---- frequency + time interpolation ---------------------------------------------------------------
  for (int p = 0; p < np && a.ce != nullptr; p++) {
    float2* ce = a.ce + ((size_t)sf * np + p) * 14 * nsc;
    if (P4 && p >= 2) {
      for (int k = tid; k < nsc; k += nt) {
        float2 h[2];
#pragma unroll
        for (int si = 0; si < 2; si++) {
          const int off = a.crs_off[p][si];
          int m = (k >= off) ? (k - off) / 6 : 0;
          if (m > M - 2) m = M - 2;
          const float f = s_ftab[k - (6 * m + off) + 5];
          const float2 v0 = s_sm[(p * 4 + si) * M + m], v1 = s_sm[(p * 4 + si) * M + m + 1];
          h[si] = make_float2(lerp_rn(v0.x, v1.x, f), lerp_rn(v0.y, v1.y, f));
        }
#pragma unroll
        for (int l = 0; l < NSYM; l++) {
          float2 o;
          if (l == 1) o = h[0];
          else if (l == NSLOT + 1) o = h[1];
          else { const float f = s_ttab2[l]; o = make_float2(lerp_rn(h[0].x, h[1].x, f), lerp_rn(h[0].y, h[1].y, f)); }
          ce[l * nsc + k] = o;
        }
      }
      continue;
    }
    for (int k = tid; k < nsc; k += nt) {
      float2 h[4];
#pragma unroll
      for (int si = 0; si < 4; si++) {
        const int off = a.crs_off[p][si];
        int m = (k >= off) ? (k - off) / 6 : 0;
        if (m > M - 2) m = M - 2;
        const float f = s_ftab[k - (6 * m + off) + 5];
        const float2 v0 = s_sm[(p * 4 + si) * M + m], v1 = s_sm[(p * 4 + si) * M + m + 1];
        h[si] = make_float2(lerp_rn(v0.x, v1.x, f), lerp_rn(v0.y, v1.y, f));
      }
#pragma unroll
      for (int l = 0; l < NSYM; l++) {
        const int s0 = (l < C1) ? 0 : (l < C2) ? 1 : 2;
        float2 o;
        if (l == 0 || l == C1 || l == C2 || l == C3) o = h[l == 0 ? 0 : l == C1 ? 1 : l == C2 ? 2 : 3];
        else { const float f = s_ttab[l]; o = make_float2(lerp_rn(h[s0].x, h[s0 + 1].x, f), lerp_rn(h[s0].y, h[s0 + 1].y, f)); }
        ce[l * nsc + k] = o;
      }
    }
  }
  // ---- measurements: three warps, one quantity each -------------------------------------------------
  const int warp = tid >> 5, lane = tid & 31;
  const int n_noise = (P4 ? 12 : np * 4) * (M - 2), n_rsrp = 4 * M, n_rssi = 4 * nsc;
  if (warp == 0) {
    const float s = warp_sum_ordered(n_noise, lane, [&](int i) {
      const int m = i % (M - 2) + 1;
      int q = i / (M - 2);                                  // q = p*4 + si; ports 2 / 3 own rows 8, 9 and 12, 13
      if (P4 && q >= 8) q = 8 + 4 * ((q - 8) >> 1) + ((q - 8) & 1);
      const float2 l = s_ls[q * M + m], t = s_sm[q * M + m];
      const float dr = __fsub_rn(l.x, t.x), di = __fsub_rn(l.y, t.y);
      return __fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di));
    });
    if (lane == 0) s_red[0] = __fdiv_rn(__fdiv_rn(s, (float)n_noise), 0.06f);
  } else if (warp == 1) {
    const float s = warp_sum_ordered(n_rsrp, lane, [&](int i) { return abs2_rn(s_ls[i]); });   // port 0 first
    if (lane == 0) s_red[1] = __fdiv_rn(s, (float)n_rsrp);
  } else if (warp == 2) {
    const float s = warp_sum_ordered(n_rssi, lane, [&](int i) { return s_pw[i]; });
    if (lane == 0) s_red[2] = __fdiv_rn(s, (float)n_rssi);
  }
  __syncthreads();
  if (tid == 0 && a.meas) {
    float* o = a.meas + (size_t)sf * 5;
    o[0] = s_red[0]; o[1] = s_red[1]; o[2] = s_red[2];
    o[3] = __fdiv_rn(__fmul_rn((float)a.nof_prb, s_red[1]), s_red[2]);
    o[4] = __fdiv_rn(s_red[1], s_red[0]);
  }
}

__global__ void __launch_bounds__(128, 10) chest_kernel(const ChestArgs a) { chest_body<false, false>(a); }
__global__ void __launch_bounds__(128, 10) chest_ext_kernel(const ChestArgs a) { chest_body<true, false>(a); }
__global__ void __launch_bounds__(512) chest_p4_kernel(const ChestArgs a) { chest_body<false, true>(a); }
__global__ void __launch_bounds__(512) chest_ext_p4_kernel(const ChestArgs a) { chest_body<true, true>(a); }

}  // namespace srsue

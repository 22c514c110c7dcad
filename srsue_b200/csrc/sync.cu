// sync.cu -- batched cell search at 1.92 Msps: PSS correlation for the three Zadoff-Chu roots at every sample offset,
// peak and CFO estimate, then SSS detection (cell group and subframe 0 / 5) at the peak (sm_100a, compiled with
// -fmad=false).
//
// Replaces what srslte_ue_cellsearch_scan computes per 5 ms of samples (/root/reference/ue/src/phy/phch_recv.cc:146-177;
// srsLTE's sync.c / pss.c / sss.c).  Arithmetic contract: oracle/SPEC.md 13 -- every correlation is a sequential sum in
// sample / subcarrier order owned by ONE thread (no cross-thread float reductions), so powers and correlations are
// bit-identical to the oracle; the peak is an integer-keyed atomic maximum (power bits, then lowest index).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

// c(p) for 256 consecutive positions of one buffer and one root.  grid (ceil(n_pos / 256), 3, n_bufs)
__global__ void __launch_bounds__(256) pss_corr_kernel(const SyncArgs a) {
  extern __shared__ __align__(16) float2 s_sync[];              // samples [256 + nfft - 1], then the replica [nfft]
  const int N = a.nfft;
  float2* s_x = s_sync;
  float2* s_t = s_sync + 256 + N;
  const int u = a.force_n_id_2 >= 0 ? a.force_n_id_2 : blockIdx.y, buf = blockIdx.z, p0 = a.first_pos + blockIdx.x * 256, tid = threadIdx.x;
  const int n_pos = a.n_samples - (N - 1);
  const float2* x = a.iq + (size_t)buf * a.stride;
  for (int i = tid; i < 256 + N - 1; i += 256) s_x[i] = (p0 + i < a.n_samples) ? x[p0 + i] : make_float2(0.f, 0.f);
  for (int i = tid; i < N; i += 256) s_t[i] = a.pss_time[u * N + i];
  __syncthreads();
  const int p = p0 + tid;
  const bool active = p < n_pos;                     // no early return: every lane takes part in the warp reduction below
  float c1r = 0.f, c1i = 0.f, c2r = 0.f, c2i = 0.f;
#pragma unroll 8
  for (int n = 0; n < N / 2; n++) {
    const float2 xv = s_x[tid + n], t = s_t[n];
    c1r = __fadd_rn(c1r, __fadd_rn(__fmul_rn(xv.x, t.x), __fmul_rn(xv.y, t.y)));
    c1i = __fadd_rn(c1i, __fsub_rn(__fmul_rn(xv.y, t.x), __fmul_rn(xv.x, t.y)));
  }
#pragma unroll 8
  for (int n = N / 2; n < N; n++) {
    const float2 xv = s_x[tid + n], t = s_t[n];
    c2r = __fadd_rn(c2r, __fadd_rn(__fmul_rn(xv.x, t.x), __fmul_rn(xv.y, t.y)));
    c2i = __fadd_rn(c2i, __fsub_rn(__fmul_rn(xv.y, t.x), __fmul_rn(xv.x, t.y)));
  }
  const float cr = __fadd_rn(c1r, c2r), ci = __fadd_rn(c1i, c2i);
  const float pw = __fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci));
  // power is non-negative, so its bit pattern orders like its value; ties go to the lowest (root, position)
  const unsigned long long key = ((unsigned long long)__float_as_uint(pw) << 32) | (0xFFFFFFFFu - (uint32_t)(u * n_pos + p));
  if (active) atomicMax(a.peak_key + buf, key);
  // mean power: warp sum in double, one atomic per warp; positions past the end contribute exactly zero
  double v = active ? (double)pw : 0.0;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
  if ((tid & 31) == 0) atomicAdd(a.power_sum + buf, v);
}

// one CTA of 128 threads per buffer: CFO at the peak, two nfft-point FFTs, SSS correlations
__global__ void __launch_bounds__(128) sss_detect_kernel(const SyncArgs a) {
  extern __shared__ __align__(16) float2 s_sync[];              // two transforms of nfft points
  __shared__ float2 s_z[62];
  __shared__ float s_best[128];
  __shared__ int s_idx[128];
  // the SSS symbol starts one symbol + one cyclic prefix before the PSS: 9 N / 128 samples of normal prefix, N / 4 of extended.
  // cp_mode 0 / 1 look at that one place, 2 tries both and keeps the larger metric (SPEC.md 15b.7; normal prefix on a tie)
  const int N = a.nfft, gap_n = N + 9 * N / 128, gap_x = N + N / 4;
  const int gap = a.cp_mode == 1 ? gap_x : gap_n;               // the smallest lead-in that allows a verdict
  float2* s_f0 = s_sync;
  float2* s_f1 = s_sync + N;
  const int buf = blockIdx.x, tid = threadIdx.x;
  const int n_pos = a.n_samples - (N - 1);
  const unsigned long long key = a.peak_key[buf];
  const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu);
  const int u = (int)(flat / (uint32_t)n_pos), pos = (int)(flat % (uint32_t)n_pos);
  const float2* x = a.iq + (size_t)buf * a.stride;
  srsue_sync_result* r = a.result + buf;
  if (tid == 0) {
    // the two half correlations again, for the CFO: angle(conj(c1) c2) / pi in units of the subcarrier spacing
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int h = 0; h < 2; h++)
      for (int n = N / 2 * h; n < N / 2 * h + N / 2; n++) {
        const float2 xv = x[pos + n], t = a.pss_time[u * N + n];
        c[2 * h] = __fadd_rn(c[2 * h], __fadd_rn(__fmul_rn(xv.x, t.x), __fmul_rn(xv.y, t.y)));
        c[2 * h + 1] = __fadd_rn(c[2 * h + 1], __fsub_rn(__fmul_rn(xv.y, t.x), __fmul_rn(xv.x, t.y)));
      }
    const float re = __fadd_rn(__fmul_rn(c[0], c[2]), __fmul_rn(c[1], c[3])), im = __fsub_rn(__fmul_rn(c[0], c[3]), __fmul_rn(c[1], c[2]));
    r->peak_pos = pos; r->n_id_2 = u;
    r->peak = __uint_as_float((uint32_t)(key >> 32));
    r->mean_power = (float)(a.power_sum[buf] / ((a.force_n_id_2 >= 0 ? 1.0 : 3.0) * (double)(n_pos - a.first_pos)));
    r->cfo = (float)(atan2((double)im, (double)re) / 3.14159265358979323846);
    r->valid = pos >= gap;
    if (pos < gap) { r->n_id_1 = -1; r->sf5 = 0; r->sss_corr = 0.f; r->cp = 0; }
  }
  if (pos < gap) return;
  float best_all = 0.f; int bc_all = -1, cp_all = 0;            // thread 0 only
  for (int hyp = 0; hyp < 2; hyp++) {
    if (a.cp_mode != 2 && hyp != a.cp_mode) continue;
    const int gap_h = hyp ? gap_x : gap_n;
    if (pos < gap_h) continue;                                  // uniform over the CTA
    __syncthreads();
  // radix-2 decimation-in-time FFTs (SPEC.md 2) of the PSS symbol (s_f0) and the SSS symbol (s_f1), both by all threads
  const int shift = 32 - a.log2n;
  for (int i = tid; i < N; i += 128) {
    const int rv = (int)(__brev((unsigned)i) >> shift);
    s_f0[rv] = x[pos + i];
    s_f1[rv] = x[pos - gap_h + i];
  }
  __syncthreads();
  for (int m = 2; m <= N; m <<= 1) {
    const int half = m >> 1;
    for (int j = tid; j < N; j += 128) {                        // butterflies 0 .. N/2-1 of each transform
      float2* base = (j < N / 2) ? s_f0 : s_f1;
      const int b = (j < N / 2) ? j : j - N / 2;
      const int g = b / half, jj = b - g * half;
      const float2 w = a.tw[jj * (N / m)];
      float2* pa = base + g * m + jj;
      float2* pb = pa + half;
      const float2 av = *pa, bv = *pb;
      const float tr = __fsub_rn(__fmul_rn(w.x, bv.x), __fmul_rn(w.y, bv.y)), ti = __fadd_rn(__fmul_rn(w.x, bv.y), __fmul_rn(w.y, bv.x));
      *pa = make_float2(__fadd_rn(av.x, tr), __fadd_rn(av.y, ti));
      *pb = make_float2(__fsub_rn(av.x, tr), __fsub_rn(av.y, ti));
    }
    __syncthreads();
  }
  if (tid < 62) {
    const int bin = (tid < 31) ? N + (tid - 31) : tid - 30;
    const float2 d = a.pss_freq[u * 62 + tid], yp = s_f0[bin], ys = s_f1[bin];
    const float hr = __fadd_rn(__fmul_rn(yp.x, d.x), __fmul_rn(yp.y, d.y)), hi = __fsub_rn(__fmul_rn(yp.y, d.x), __fmul_rn(yp.x, d.y));
    s_z[tid] = make_float2(__fadd_rn(__fmul_rn(ys.x, hr), __fmul_rn(ys.y, hi)), __fsub_rn(__fmul_rn(ys.y, hr), __fmul_rn(ys.x, hi)));
  }
  __syncthreads();
  // candidate c = sf5 * 168 + n_id_1; each thread scans its candidates in ascending order and keeps its first maximum
  float best = 0.f; int bi = -1;
  for (int c = tid; c < 336; c += 128) {
    const int8_t* sq = a.sss + ((size_t)u * 336 + c) * 62;
    float ar = 0.f, ai = 0.f;
    for (int i = 0; i < 62; i++) {
      const float2 z = s_z[i];
      ar = __fadd_rn(ar, sq[i] > 0 ? z.x : -z.x);
      ai = __fadd_rn(ai, sq[i] > 0 ? z.y : -z.y);
    }
    const float acc = __fadd_rn(__fmul_rn(ar, ar), __fmul_rn(ai, ai));      // a carrier offset only turns the sum
    if (bi < 0 || acc > best) { best = acc; bi = c; }
  }
  s_best[tid] = best; s_idx[tid] = bi;
  __syncthreads();
  if (tid == 0) {
    float bv = s_best[0]; int bc = s_idx[0];
    for (int t = 1; t < 128; t++)
      if (s_idx[t] >= 0 && (s_best[t] > bv || (s_best[t] == bv && s_idx[t] < bc))) { bv = s_best[t]; bc = s_idx[t]; }
    if (bc_all < 0 || bv > best_all) { best_all = bv; bc_all = bc; cp_all = hyp; }
  }
  }
  if (tid == 0) { r->n_id_1 = bc_all % 168; r->sf5 = bc_all / 168; r->sss_corr = best_all; r->cp = cp_all; }
}

}  // namespace srsue

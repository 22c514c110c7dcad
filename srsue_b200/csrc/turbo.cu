// turbo.cu -- K5/K6a: batched LTE turbo decoder for sm_100a (int16 max-log-MAP, parallel windows with
// next-iteration initialisation, per-code-block CRC early stop).  Arithmetic: oracle/SPEC.md section 7.
//
// Replaces srsLTE's srslte_tdec_* (int16 "gen"/SSE/AVX decoders) that srsUE reaches through
// srslte_pdsch_decode_rnti (/root/reference/ue/src/phy/phch_worker.cc:347-348) and whose iteration
// budget it sets with srslte_sch_set_max_noi (phch_worker.cc:88).
//
// Mapping.  One thread owns TWO adjacent windows of one code block: every 32-bit register holds the
// same trellis state of both windows as packed int16x2, so the whole add-compare-select is
// VIADD.16x2 / VIADDMNMX.S16x2 with no cross-lane traffic.  A CTA keeps `ncb_cta` code blocks resident:
//   shared memory : the extrinsic exchange array A (window-transposed, so natural-order accesses are one
//                   conflict-free LDS.32 per thread and QPP-interleaved accesses are conflict-free
//                   LDS.U16 by the contention-free property), the beta checkpoints of the current pass
//                   and the DEC2 position table of this K;
//   L2 / HBM      : the channel LLRs (read-only, window-transposed "tcb" layout -> coalesced 32-bit
//                   loads), the window-boundary metrics (NII) and the hard-decision bytes.
// Each MAP pass is: backward sweep storing beta every 8 steps, then forward sweep that re-creates
// beta for 8 steps in registers and produces alpha, the extrinsic and (DEC2) the hard decision.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {

constexpr int kSW = 8;                 // sub-window: beta values re-created in registers
constexpr uint32_t kNegInfPair = ((uint32_t)(uint16_t)(-kTdInf) << 16) | (uint16_t)(-kTdInf);
constexpr uint32_t kEPair = ((uint32_t)kTdE << 16) | (uint32_t)kTdE;
constexpr uint32_t kNegEPair = ((uint32_t)(uint16_t)(-kTdE) << 16) | (uint16_t)(-kTdE);

__device__ __forceinline__ uint32_t vadd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t vsub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
// max(a + b, c) per int16 half, the add wrapping (VIADDMNMX.S16x2)
__device__ __forceinline__ uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t vclampE(uint32_t v) { return __vmins2(__vmaxs2(v, kNegEPair), kEPair); }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu); }

// beta_k from beta_{k+1} (SPEC 7.4); branch labels follow the RSC trellis of 36.212 5.1.3.2.1
__device__ __forceinline__ void beta_step(const uint32_t (&bn)[8], uint32_t (&b)[8], uint32_t x, uint32_t y, uint32_t xy) {
  b[0] = vaddmax(bn[4], xy, bn[0]);
  b[1] = vaddmax(bn[0], xy, bn[4]);
  b[2] = vaddmax(bn[5], y, vadd(bn[1], x));
  b[3] = vaddmax(bn[1], y, vadd(bn[5], x));
  b[4] = vaddmax(bn[2], y, vadd(bn[6], x));
  b[5] = vaddmax(bn[6], y, vadd(bn[2], x));
  b[6] = vaddmax(bn[3], xy, bn[7]);
  b[7] = vaddmax(bn[7], xy, bn[3]);
}

__device__ __forceinline__ void alpha_step(uint32_t (&a)[8], uint32_t x, uint32_t y, uint32_t xy) {
  uint32_t n[8];
  n[0] = vaddmax(a[1], xy, a[0]);
  n[1] = vaddmax(a[2], x, vadd(a[3], y));
  n[2] = vaddmax(a[4], y, vadd(a[5], x));
  n[3] = vaddmax(a[6], xy, a[7]);
  n[4] = vaddmax(a[0], xy, a[1]);
  n[5] = vaddmax(a[2], y, vadd(a[3], x));
  n[6] = vaddmax(a[4], x, vadd(a[5], y));
  n[7] = vaddmax(a[7], xy, a[6]);
#pragma unroll
  for (int s = 0; s < 8; s++) a[s] = n[s];
}

__device__ __forceinline__ void normalise(uint32_t (&m)[8]) {
  const uint32_t m0 = m[0];
#pragma unroll
  for (int s = 1; s < 8; s++) m[s] = vsub(m[s], m0);
  m[0] = 0;
}

// extrinsic of one trellis step from alpha_k and beta_{k+1}: max(A10, A11 + y) - max(A00, A01 + y)
__device__ __forceinline__ uint32_t ext_step(const uint32_t (&a)[8], const uint32_t (&bn)[8], uint32_t y) {
  // Pipe balance (measured, profiles/alu_peak_r01.json): VIADD.16x2 issues on the FMA pipe, VIADDMNMX /
  // VIMNMX3 on the ALU pipe, 64 lanes/clk/SM each.  Three plain adds + one 3-input max + one fused add-max
  // per class puts 8 instructions on the ALU pipe instead of 12 (max of 4 sums is order-independent).
  const uint32_t a00 = vaddmax(a[7], bn[3], __vimax3_s16x2(vadd(a[0], bn[0]), vadd(a[1], bn[4]), vadd(a[6], bn[7])));
  const uint32_t a11 = vaddmax(a[7], bn[7], __vimax3_s16x2(vadd(a[0], bn[4]), vadd(a[1], bn[0]), vadd(a[6], bn[3])));
  const uint32_t a01 = vaddmax(a[5], bn[6], __vimax3_s16x2(vadd(a[2], bn[5]), vadd(a[3], bn[1]), vadd(a[4], bn[2])));
  const uint32_t a10 = vaddmax(a[5], bn[2], __vimax3_s16x2(vadd(a[2], bn[1]), vadd(a[3], bn[5]), vadd(a[4], bn[6])));
  const uint32_t l1 = vaddmax(a11, y, a10);
  const uint32_t l0 = vaddmax(a01, y, a00);
  return vsub(l1, l0);
}

// PRMT with the sign-replicate bit of every selector nibble set (__byte_perm masks that bit off)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t sign_fill(uint32_t w) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"(SEL));
  return r;
}

__device__ __forceinline__ uint32_t gf_mul24(uint32_t a, uint32_t b, uint32_t poly) {
  uint32_t r = 0;
#pragma unroll 1
  for (int i = 23; i >= 0; i--) {
    r <<= 1;
    if (r & 0x1000000u) r ^= poly;
    if ((b >> i) & 1u) r ^= a;
  }
  return r & 0xFFFFFFu;
}

struct SlotCtx {
  const uint4* sys4;      // channel LLR planes (global), this thread's first 8-step group: 2 x uint4 per group
  const uint4* p14;
  const uint4* p24;
  const int16_t* tail;    // 12 tail LLRs (global)
  uint32_t* Aw;           // shared: extrinsic exchange array [W][Ppad] as packed pairs, word i*T + t
  uint4* ckpt4;           // shared: beta checkpoints [nsw][T][2 x uint4], this thread's entry of group 0
  const uint16_t* perm16; // shared: DEC2 positions [nsw][T][8 steps][2 windows], this thread's entry of group 0
  int16_t* nii;           // global: [2 dec][2 pp][2 kind][8][NP]
  uint16_t* bits;         // global: [W][Ppad] hard decisions, sign bit of each 16-bit word
};

__device__ __forceinline__ void ld8(const uint4* p, uint32_t (&v)[8]) {
  const uint4 a = __ldg(p), b = __ldg(p + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8(const uint4* p, uint32_t (&v)[8]) {
  const uint4 a = p[0], b = p[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void sts8(uint4* p, const uint32_t (&v)[8]) {
  p[0] = make_uint4(v[0], v[1], v[2], v[3]);
  p[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// One max-log-MAP pass of constituent decoder DEC (0 or 1) for the two windows of this thread.
template <int DEC>
__device__ __forceinline__ void map_pass(const TurboArgs& g, const SlotCtx& c, int t, int it, bool store_bits) {
  const int T = g.T, W = g.W, P = g.P, NP = g.Ppad + 2, nsw = W / kSW;
  const int j0 = 2 * t, j1 = 2 * t + 1;
  const int rd = it & 1, wr = rd ^ 1;
  const int16_t* nii_a_rd = c.nii + ((DEC * 2 + rd) * 2 + 0) * 8 * NP;
  const int16_t* nii_b_rd = c.nii + ((DEC * 2 + rd) * 2 + 1) * 8 * NP;
  uint32_t* nii_a_wr = reinterpret_cast<uint32_t*>(c.nii + ((DEC * 2 + wr) * 2 + 0) * 8 * NP);
  uint32_t* nii_b_wr = reinterpret_cast<uint32_t*>(c.nii + ((DEC * 2 + wr) * 2 + 1) * 8 * NP);
  const uint4* yq = DEC ? c.p24 : c.p14;
  int16_t* A16 = reinterpret_cast<int16_t*>(c.Aw);
  const int gstride = 2 * T;                       // uint4 per 8-step group

  uint32_t b[8];
  // ---- beta at the end of the two windows -------------------------------------------------------
  {
    const bool last0 = (j0 == P - 1), last1 = (j1 == P - 1);
    if (last0 || last1) {
      // trellis termination: beta_{K+3} = (0, -INF, ...) and three ordinary steps over the tail
      uint32_t bt[8], bq[8];
      bt[0] = 0;
#pragma unroll
      for (int s = 1; s < 8; s++) bt[s] = kNegInfPair;
#pragma unroll
      for (int q = 2; q >= 0; q--) {
        const int xv = c.tail[DEC * 6 + 2 * q], yv = c.tail[DEC * 6 + 2 * q + 1];
        const uint32_t x = pack16(xv, xv), y = pack16(yv, yv);
        beta_step(bt, bq, x, y, vadd(x, y));
#pragma unroll
        for (int s = 0; s < 8; s++) bt[s] = bq[s];
      }
      normalise(bt);                                  // metric index K is a multiple of 4
#pragma unroll
      for (int s = 0; s < 8; s++) {
        int lo = 0, hi = 0;
        if (last0) lo = (int)(int16_t)(bt[s] & 0xFFFFu); else if (it) lo = nii_b_rd[s * NP + j0 + 1];
        if (last1) hi = (int)(int16_t)(bt[s] >> 16); else if (it) hi = nii_b_rd[s * NP + j1 + 1];
        b[s] = pack16(lo, hi);
      }
    } else if (it) {
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = pack16(nii_b_rd[s * NP + j0 + 1], nii_b_rd[s * NP + j1 + 1]);
    } else {
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = 0;
    }
  }
  // ---- alpha at the start of the two windows ------------------------------------------------------
  uint32_t a[8];
#pragma unroll
  for (int s = 0; s < 8; s++) {
    int lo, hi;
    if (j0 == 0) lo = s ? -kTdInf : 0; else lo = it ? nii_a_rd[s * NP + j0 - 1] : 0;
    hi = it ? nii_a_rd[s * NP + j0] : 0;
    a[s] = pack16(lo, hi);
  }

  // ---- pass 1: backward sweep, checkpoint beta every kSW steps -----------------------------------
  uint32_t ny[kSW], ns[kSW];                        // register prefetch of the next group's channel LLRs
  ld8(yq + (nsw - 1) * gstride, ny);
  if (DEC == 0) ld8(c.sys4 + (nsw - 1) * gstride, ns);
#pragma unroll 1
  for (int sw = nsw - 1; sw >= 0; sw--) {
    uint32_t x[kSW], y[kSW];
#pragma unroll
    for (int i = 0; i < kSW; i++) y[i] = ny[i];
    if (DEC == 0) {
      const uint32_t* ap = c.Aw + sw * kSW * T + t;
#pragma unroll
      for (int i = 0; i < kSW; i++) x[i] = vadd(ns[i], ap[i * T]);
    } else {
      // position table read as LDS.U16 (immediate offsets, LSU pipe) instead of unpacking pairs on the ALU pipe
      const uint16_t* pq = c.perm16 + sw * (2 * kSW) * T;
#pragma unroll
      for (int i = 0; i < kSW; i++) x[i] = pack16(A16[pq[2 * i]], A16[pq[2 * i + 1]]);
    }
    if (sw > 0) {
      ld8(yq + (sw - 1) * gstride, ny);
      if (DEC == 0) ld8(c.sys4 + (sw - 1) * gstride, ns);
    }
    sts8(c.ckpt4 + sw * gstride, b);
#pragma unroll
    for (int i = kSW - 1; i >= 0; i--) {
      uint32_t nb[8];
      beta_step(b, nb, x[i], y[i], vadd(x[i], y[i]));
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = nb[s];
      if ((i & 3) == 0) normalise(b);
    }
  }
  // beta at the window start feeds the previous window in the next iteration
#pragma unroll
  for (int s = 0; s < 8; s++) nii_b_wr[(s * NP + j0) >> 1] = b[s];

  // ---- pass 2: forward sweep ----------------------------------------------------------------------
  ld8(yq, ny);
  if (DEC == 0) ld8(c.sys4, ns);
#pragma unroll 1
  for (int sw = 0; sw < nsw; sw++) {
    uint32_t x[kSW], y[kSW], aux[kSW];     // aux: DEC1 systematic LLRs
    uint32_t* ap = c.Aw + sw * kSW * T + t;
    const uint16_t* pq = c.perm16 + sw * (2 * kSW) * T;
#pragma unroll
    for (int i = 0; i < kSW; i++) y[i] = ny[i];
    if (DEC == 0) {
#pragma unroll
      for (int i = 0; i < kSW; i++) { aux[i] = ns[i]; x[i] = vadd(ns[i], ap[i * T]); }
    } else {
#pragma unroll
      for (int i = 0; i < kSW; i++) x[i] = pack16(A16[pq[2 * i]], A16[pq[2 * i + 1]]);
    }
    if (sw + 1 < nsw) {
      ld8(yq + (sw + 1) * gstride, ny);
      if (DEC == 0) ld8(c.sys4 + (sw + 1) * gstride, ns);
    }
    uint32_t B[kSW][8];                     // B[i] = beta_{i+1} of this sub-window
    lds8(c.ckpt4 + sw * gstride, B[kSW - 1]);
#pragma unroll
    for (int i = kSW - 1; i >= 1; i--) {
      beta_step(B[i], B[i - 1], x[i], y[i], vadd(x[i], y[i]));
      if ((i & 3) == 0) normalise(B[i - 1]);
    }
#pragma unroll
    for (int i = 0; i < kSW; i++) {
      const uint32_t ext = ext_step(a, B[i], y[i]);
      const uint32_t la = vclampE(ext);
      if (DEC == 0) {
        ap[i * T] = vadd(aux[i], la);
      } else {
        const uint32_t p0 = pq[2 * i], p1 = pq[2 * i + 1];
        A16[p0] = (int16_t)(la & 0xFFFFu);
        A16[p1] = (int16_t)(la >> 16);
        if (store_bits) {
          // decision = (x + ext) > 0  <=>  sign bit of -(x + ext), kept as the sign of a 16-bit word
          const uint32_t nd = vsub(0u, vadd(x[i], ext));
          c.bits[p0] = (uint16_t)nd;
          c.bits[p1] = (uint16_t)(nd >> 16);
        }
      }
      alpha_step(a, x[i], y[i], vadd(x[i], y[i]));
      if ((i & 3) == 3) normalise(a);
    }
  }
#pragma unroll
  for (int s = 0; s < 8; s++) nii_a_wr[(s * NP + j0) >> 1] = a[s];
}

}  // namespace

__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_kernel(const TurboArgs g) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int T = g.T, W = g.W, P = g.P, plane = g.plane, NP = g.Ppad + 2, nsw = W / kSW;
  const int tid = threadIdx.x;
  const int slot = tid / T, t = tid - slot * T;
  const bool valid = slot < g.ncb_cta;

  uint32_t* s_permw = smem;                                  // plane/2 words, [nsw][T][8]
  uint32_t* s_crc = s_permw + plane / 2;                     // ncb_cta words (rounded to 4)
  uint32_t* s_slots = s_crc + ((g.ncb_cta + 3) & ~3);
  const int slot_words = plane / 2 + nsw * 8 * T;            // A + checkpoints

  for (int i = tid; i < plane / 2; i += blockDim.x) s_permw[i] = reinterpret_cast<const uint32_t*>(g.perm_pos)[i];

  SlotCtx c;
  c.Aw = s_slots + (size_t)(valid ? slot : 0) * slot_words;
  c.ckpt4 = reinterpret_cast<uint4*>(c.Aw + plane / 2) + 2 * t;
  c.perm16 = reinterpret_cast<const uint16_t*>(s_permw) + 2 * kSW * t;
  const size_t gslot = (size_t)blockIdx.x * g.ncb_cta + (valid ? slot : 0);
  c.nii = g.nii + gslot * (size_t)(2 * 2 * 2 * 8 * NP);
  c.bits = reinterpret_cast<uint16_t*>(g.bits_scratch) + gslot * (size_t)plane;
  __syncthreads();

  const int n_groups = (g.n_cb + g.ncb_cta - 1) / g.ncb_cta;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int cb = grp * g.ncb_cta + slot;
    const bool active = valid && cb < g.n_cb;
    const long long cbi = active ? (g.cb_list ? g.cb_list[cb] : cb) : 0;
    const int16_t* in_cb = g.in + cbi * g.in_stride;
    c.sys4 = reinterpret_cast<const uint4*>(in_cb) + 2 * t;
    c.p14 = c.sys4 + plane / 8;
    c.p24 = c.p14 + plane / 8;
    c.tail = in_cb + 3 * plane;
    // a-priori LLRs start at zero: each thread clears its own column of A
    if (active)
      for (int i = 0; i < W; i++) c.Aw[i * T + t] = 0u;
    // pull the channel LLRs of this CTA's NEXT code blocks into L2 while the current ones are decoded,
    // so that their first (otherwise HBM-latency) pass finds them on chip
    {
      const int ncb_next = (grp + (int)gridDim.x) * g.ncb_cta + slot;
      if (valid && ncb_next < g.n_cb) {
        const long long nbi = g.cb_list ? g.cb_list[ncb_next] : ncb_next;
        const char* base = reinterpret_cast<const char*>(g.in + nbi * g.in_stride);
        const int bytes = (3 * plane + 16) * 2;
        for (int o = t * 128; o < bytes; o += T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }

    bool done = !active;
    int n_iter = 0, crc_ok = 0;
    for (int it = 0; it < g.max_iter; it++) {
      if (valid && t == 0) s_crc[slot] = 0;
      const bool store_bits = (g.crc_type != 0) || (it == g.max_iter - 1);
      if (!done) map_pass<0>(g, c, t, it, store_bits);
      __syncthreads();
      if (!done) { map_pass<1>(g, c, t, it, store_bits); n_iter = it + 1; }
      __syncthreads();
      if (g.crc_type != 0) {
        if (!done) {
          // per-window CRC: remainder of (window bits * x^24), then shifted to the window's place
          uint32_t c0 = 0, c1 = 0;
          const uint32_t* bp = reinterpret_cast<const uint32_t*>(c.bits) + t;
#pragma unroll 16
          for (int i = 0; i < W; i++) {
            const uint32_t w = bp[i * T];
            const uint32_t u = __ldg(g.crcU + i);
            c0 ^= sign_fill<0x9999>(w) & u;      // all-ones when the low half is negative (bit 15)
            c1 ^= sign_fill<0xBBBB>(w) & u;      // same for the high half (bit 31)
          }
          const uint32_t contrib = gf_mul24(c0, __ldg(g.crcV + 2 * t), g.crc_poly) ^
                                   gf_mul24(c1, __ldg(g.crcV + 2 * t + 1), g.crc_poly);
          atomicXor(&s_crc[slot], contrib);
        }
        __syncthreads();
        if (!done && s_crc[slot] == 0) { done = true; crc_ok = 1; }
      }
      if (__syncthreads_and(done)) break;
    }
    // ---- pack the hard decisions of the last iteration, MSB first, natural order ----------------
    if (active) {
      uint8_t* out = g.out_bits + cbi * (long long)g.out_stride;
      const int wbytes = W / 8;
      const uint32_t* bp = reinterpret_cast<const uint32_t*>(c.bits) + t;
#pragma unroll 2
      for (int bb = 0; bb < wbytes; bb++) {
        uint32_t v0 = 0, v1 = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const uint32_t w = bp[(bb * 8 + q) * T];
          v0 = (v0 << 1) | ((w >> 15) & 1u);
          v1 = (v1 << 1) | (w >> 31);
        }
        out[(2 * t) * wbytes + bb] = (uint8_t)v0;
        if (2 * t + 1 < P) out[(2 * t + 1) * wbytes + bb] = (uint8_t)v1;
      }
      if (t == 0) g.out_status[cbi] = n_iter | (crc_ok << 8);
    }
    __syncthreads();
  }
}

// ---- layout conversion at the API edge -----------------------------------------------------------
// srsLTE decoder-input order (3K+12 interleaved triples) -> tcb layout, clamping to +-C (SPEC 7.2)
__global__ void triples_to_tcb_kernel(const int16_t* __restrict__ in, long long in_stride, int16_t* __restrict__ out,
                                      long long out_stride, int n_cb, TurboGeomDev g) {
  const int cb = blockIdx.y;
  if (cb >= n_cb) return;
  const int16_t* src = in + cb * in_stride;
  int16_t* dst = out + cb * out_stride;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.cb_elems; e += gridDim.x * blockDim.x) {
    int v = 0;
    if (e < 3 * g.plane) {
      // tcb element -> (stream, window j, step i): element ((sw*T + t)*8 + ii)*2 + h, j = 2t + h, i = 8sw + ii
      const int stream = e / g.plane, r = e - stream * g.plane;
      const int h = r & 1, ii = (r >> 1) & 7, q = r >> 4;
      const int sw = q / g.T, t = q - sw * g.T;
      const int i = sw * 8 + ii, j = 2 * t + h;
      if (j < g.P) v = src[3 * (j * g.W + i) + stream];
    } else if (e - 3 * g.plane < 12) {
      v = src[3 * g.K + (e - 3 * g.plane)];
    }
    v = max(-kTdC, min(kTdC, v));
    dst[e] = (int16_t)v;
  }
}

__global__ void tcb_to_triples_kernel(const int16_t* __restrict__ in, long long in_stride, int16_t* __restrict__ out,
                                      long long out_stride, int n_cb, TurboGeomDev g) {
  const int cb = blockIdx.y;
  if (cb >= n_cb) return;
  const int16_t* src = in + cb * in_stride;
  int16_t* dst = out + cb * out_stride;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * g.K + 12; e += gridDim.x * blockDim.x) {
    int off;
    if (e < 3 * g.K) {
      const int k = e / 3, stream = e - 3 * k;
      const int j = k / g.W, i = k - j * g.W;
      off = stream * g.plane + ((((i >> 3) * g.T + (j >> 1)) * 8 + (i & 7)) << 1) + (j & 1);
    } else {
      off = 3 * g.plane + (e - 3 * g.K);
    }
    dst[e] = src[off];
  }
}

}  // namespace srsue

// turbo.cu -- K5/K6a: batched LTE turbo decoder for sm_100a (int16 max-log-MAP, parallel windows with
// next-iteration initialisation, per-code-block CRC early stop).  Arithmetic: oracle/SPEC.md section 7.
//
// Replaces srsLTE's srslte_tdec_* (int16 "gen"/SSE/AVX decoders) that srsUE reaches through
// srslte_pdsch_decode_rnti (/root/reference/ue/src/phy/phch_worker.cc:347-348) and whose iteration
// budget it sets with srslte_sch_set_max_noi (phch_worker.cc:88).
//
// Mapping.  One thread owns TWO adjacent windows of one code block: every 32-bit register holds the
// same trellis state of both windows as packed int16x2, so the whole add-compare-select is
// VIADD.16x2 (FMA pipe) / VIADDMNMX.S16x2 (ALU pipe) with no cross-lane traffic.  A CTA is persistent and
// keeps `ncb_cta` code-block SLOTS of T threads each; a slot that finishes a block (CRC passed or budget
// spent) takes the next one from a global counter, so early stopping saves time per block.
//   shared memory : the extrinsic exchange array A of every slot (window-transposed: natural-order
//                   accesses are one conflict-free LDS.32 per thread, QPP-interleaved accesses conflict-free
//                   LDS.U16 by the contention-free property), the DEC2 position table and the per-position
//                   CRC contribution table of this K;
//   L2 / HBM      : the channel LLRs (read-only, thread-contiguous "tcb" layout -> coalesced LDG.128,
//                   register-prefetched), the thread-private beta checkpoints of the current pass, the
//                   window-boundary metrics (NII) and the hard decisions.
// Each MAP pass is: backward sweep storing beta every 8 steps, then forward sweep that re-creates beta
// for 8 steps in registers and produces alpha, the extrinsic and (DEC2) the hard decision, whose CRC
// contribution x^(position) mod g is accumulated on the fly (the CRC is linear over GF(2)).
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
// lo + hi * 65536 with lo, hi zero-extended 16-bit loads: one IMAD on the FMA pipe (a PRMT would take an
// ALU-pipe slot, and the ALU pipe is the one this kernel saturates)
__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) { return hi * 65536u + lo; }
// store a 16-bit value at base[idx] with the address formed by one IMAD.WIDE (FMA pipe)
__device__ __forceinline__ void st16_wide(uint16_t* base, uint32_t idx, uint32_t v) {
  uint64_t addr;
  asm("mad.wide.u32 %0, %1, 2, %2;" : "=l"(addr) : "r"(idx), "l"(base));
  asm volatile("st.global.u16 [%0], %1;" ::"l"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ uint32_t pack16s(int lo, int hi) { return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu); }

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
  uint32_t a00 = vadd(a[0], bn[0]);
  a00 = vaddmax(a[1], bn[4], a00); a00 = vaddmax(a[6], bn[7], a00); a00 = vaddmax(a[7], bn[3], a00);
  uint32_t a11 = vadd(a[0], bn[4]);
  a11 = vaddmax(a[1], bn[0], a11); a11 = vaddmax(a[6], bn[3], a11); a11 = vaddmax(a[7], bn[7], a11);
  uint32_t a01 = vadd(a[2], bn[5]);
  a01 = vaddmax(a[3], bn[1], a01); a01 = vaddmax(a[4], bn[2], a01); a01 = vaddmax(a[5], bn[6], a01);
  uint32_t a10 = vadd(a[2], bn[1]);
  a10 = vaddmax(a[3], bn[5], a10); a10 = vaddmax(a[4], bn[6], a10); a10 = vaddmax(a[5], bn[2], a10);
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

__device__ __forceinline__ void ld8(const uint4* p, uint32_t (&v)[8]) {
  const uint4 a = __ldg(p), b = __ldg(p + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ldp8(const uint4* p, uint32_t (&v)[8]) {      // plain (coherent) load
  const uint4 a = p[0], b = p[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(uint4* p, const uint32_t (&v)[8]) {
  p[0] = make_uint4(v[0], v[1], v[2], v[3]);
  p[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// What one slot thread needs to find its code block and its scratch; everything else is derived in place
// (keeps the live state across the long unrolled sweeps small).
struct SlotCtx {
  const uint4* in4;       // code block in tcb layout (global)
  uint32_t* Aw;           // shared: extrinsic exchange array [W][Ppad] as packed pairs, word i*T + t
  uint32_t gslot;         // global slot number: selects the NII / checkpoint / decision scratch
};

// One max-log-MAP pass of constituent decoder DEC (0 or 1) for the two windows of this thread.  Returns the
// thread's CRC contribution of the hard decisions (DEC2 with crc_on), else 0.
template <int DEC, bool CRC>
__device__ __forceinline__ uint32_t map_pass(const TurboArgs& g, const SlotCtx& c, const uint16_t* perm16,
                                             const uint32_t* s_tpos, int t, int it) {
  const int T = g.T, W = g.W, P = g.P, NP = g.Ppad + 2, nsw = W / kSW, plane = g.plane;
  const int j0 = 2 * t, j1 = 2 * t + 1;
  const int rd = it & 1, wr = rd ^ 1;
  int16_t* nii = g.nii + (size_t)c.gslot * (size_t)(2 * 2 * 2 * 8 * NP);
  const int16_t* nii_a_rd = nii + ((DEC * 2 + rd) * 2 + 0) * 8 * NP;
  const int16_t* nii_b_rd = nii + ((DEC * 2 + rd) * 2 + 1) * 8 * NP;
  uint32_t* nii_a_wr = reinterpret_cast<uint32_t*>(nii + ((DEC * 2 + wr) * 2 + 0) * 8 * NP);
  uint32_t* nii_b_wr = reinterpret_cast<uint32_t*>(nii + ((DEC * 2 + wr) * 2 + 1) * 8 * NP);
  const int gstride = 2 * T;                       // uint4 per 8-step group
  const uint4* sysq = c.in4 + 2 * t;
  const uint4* yq = sysq + (DEC ? 2 : 1) * (plane / 8);
  // checkpoints of one CTA are laid out [sub-window][slot][thread]: for small code blocks (few threads per slot) the
  // slots a warp spans are then contiguous, one 16-byte access per thread = whole 128-byte lines
  const int slot_in_cta = (int)(c.gslot - blockIdx.x * g.ncb_cta);
  const int cstride = g.ncb_cta * gstride;         // uint4 per sub-window of the whole CTA
  uint4* ckpt4 = g.ckpt + (size_t)blockIdx.x * (size_t)(nsw * cstride) + slot_in_cta * gstride + 2 * t;
  uint16_t* bits = reinterpret_cast<uint16_t*>(g.bits_scratch) + (size_t)c.gslot * (size_t)plane;
  int16_t* A16 = reinterpret_cast<int16_t*>(c.Aw);

  uint32_t b[8];
  // ---- beta at the end of the two windows -------------------------------------------------------
  {
    const bool last0 = (j0 == P - 1), last1 = (j1 == P - 1);
    if (last0 || last1) {
      // trellis termination: beta_{K+3} = (0, -INF, ...) and three ordinary steps over the tail
      const int16_t* tail = reinterpret_cast<const int16_t*>(c.in4) + 3 * plane;
      uint32_t bt[8], bq[8];
      bt[0] = 0;
#pragma unroll
      for (int s = 1; s < 8; s++) bt[s] = kNegInfPair;
#pragma unroll
      for (int q = 2; q >= 0; q--) {
        const int xv = tail[DEC * 6 + 2 * q], yv = tail[DEC * 6 + 2 * q + 1];
        const uint32_t x = pack16s(xv, xv), y = pack16s(yv, yv);
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
        b[s] = pack16s(lo, hi);
      }
    } else if (it) {
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = pack16s(nii_b_rd[s * NP + j0 + 1], nii_b_rd[s * NP + j1 + 1]);
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
    a[s] = pack16s(lo, hi);
  }

  // ---- pass 1: backward sweep, checkpoint beta every kSW steps -----------------------------------
  // Register prefetch TWO groups ahead: all warps of the CTA run this short, load-dominated sweep at the
  // same time (the passes are barrier-separated), so there is no other phase to hide the L2 latency behind.
  // The last group of this sweep (steps 0..7) is the first group of the forward sweep: its channel LLRs (ny, ns)
  // and the beta vector it starts from (nbeta) stay in registers across the pass boundary instead of being
  // re-loaded from L2 with nothing to overlap.
  uint32_t ny[kSW], ns[kSW], nbeta[8];
  {
    uint32_t my[kSW], ms[kSW];
    ld8(yq + (nsw - 1) * gstride, ny);
    if (DEC == 0) ld8(sysq + (nsw - 1) * gstride, ns);
    if (nsw > 1) {
      ld8(yq + (nsw - 2) * gstride, my);
      if (DEC == 0) ld8(sysq + (nsw - 2) * gstride, ms);
    }
    auto group = [&](int sw, const uint32_t (&y)[kSW], const uint32_t (&sv)[kSW]) {
      uint32_t x[kSW];
      if (DEC == 0) {
        const uint32_t* ap = c.Aw + sw * kSW * T + t;
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = vadd(sv[i], ap[i * T]);
      } else {
        // position table read as LDS.U16 (immediate offsets, LSU pipe), not unpacked on the ALU pipe
        const uint16_t* pq = perm16 + sw * (2 * kSW) * T;
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = pack16((uint16_t)A16[pq[2 * i]], (uint16_t)A16[pq[2 * i + 1]]);
      }
      if (sw) st8(ckpt4 + sw * cstride, b);         // thread-private scratch, read back in pass 2 (group 0 stays in registers)
#pragma unroll
      for (int i = kSW - 1; i >= 0; i--) {
        uint32_t nb[8];
        beta_step(b, nb, x[i], y[i], vadd(x[i], y[i]));
#pragma unroll
        for (int s = 0; s < 8; s++) b[s] = nb[s];
        if ((i & 3) == 0) normalise(b);
      }
    };
#pragma unroll 1
    for (int sw = nsw - 1; sw >= 1; sw--) {
      uint32_t y[kSW], sv[kSW];
#pragma unroll
      for (int i = 0; i < kSW; i++) { y[i] = ny[i]; ny[i] = my[i]; if (DEC == 0) { sv[i] = ns[i]; ns[i] = ms[i]; } }
      if (sw > 1) {
        ld8(yq + (sw - 2) * gstride, my);
        if (DEC == 0) ld8(sysq + (sw - 2) * gstride, ms);
      }
      group(sw, y, sv);
    }
#pragma unroll
    for (int s = 0; s < 8; s++) nbeta[s] = b[s];
    group(0, ny, ns);
  }
  // beta at the window start feeds the previous window in the next iteration
#pragma unroll
  for (int s = 0; s < 8; s++) nii_b_wr[(s * NP + j0) >> 1] = b[s];

  // ---- pass 2: forward sweep ----------------------------------------------------------------------
  uint32_t crc = 0;
  {
    // ny / ns / nbeta of sub-window 0 come from the backward sweep; later ones are prefetched one group ahead
#pragma unroll 1
    for (int sw = 0; sw < nsw; sw++) {
      uint32_t x[kSW], y[kSW], aux[kSW];            // aux: DEC1 systematic LLRs
      uint32_t* ap = c.Aw + sw * kSW * T + t;
      const uint16_t* pq = perm16 + sw * (2 * kSW) * T;
#pragma unroll
      for (int i = 0; i < kSW; i++) y[i] = ny[i];
      if (DEC == 0) {
#pragma unroll
        for (int i = 0; i < kSW; i++) { aux[i] = ns[i]; x[i] = vadd(ns[i], ap[i * T]); }
      } else {
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = pack16((uint16_t)A16[pq[2 * i]], (uint16_t)A16[pq[2 * i + 1]]);
      }
      uint32_t B[kSW][8];                           // B[i] = beta_{i+1} of this sub-window
#pragma unroll
      for (int s = 0; s < 8; s++) B[kSW - 1][s] = nbeta[s];
      if (sw + 1 < nsw) {
        ld8(yq + (sw + 1) * gstride, ny);
        if (DEC == 0) ld8(sysq + (sw + 1) * gstride, ns);
        ldp8(ckpt4 + (sw + 1) * cstride, nbeta);
      }
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
          // decision = (x + ext) > 0  <=>  sign bit of -(x + ext), kept as the sign of a 16-bit word
          const uint32_t nd = vsub(0u, vadd(x[i], ext));
          st16_wide(bits, p0, nd);
          st16_wide(bits, p1, nd >> 16);
          if (CRC) crc ^= (sign_fill<0x9999>(nd) & __ldg(s_tpos + p0)) ^ (sign_fill<0xBBBB>(nd) & __ldg(s_tpos + p1));
        }
        alpha_step(a, x[i], y[i], vadd(x[i], y[i]));
        if ((i & 3) == 3) normalise(a);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 8; s++) nii_a_wr[(s * NP + j0) >> 1] = a[s];
  return crc;
}

template <bool CRC>
__device__ __forceinline__ void turbo_decode_body(const TurboArgs& g) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int T = g.T, W = g.W, P = g.P, plane = g.plane;
  const int tid = threadIdx.x;
  const int slot = tid / T, t = tid - slot * T;
  const bool valid = slot < g.ncb_cta;
  constexpr bool crc_on = CRC;
  const int nflag = (g.ncb_cta + 3) & ~3;

  uint32_t* s_permw = smem;                                  // plane/2 words: DEC2 positions [W/8][T][8][2] x u16
  const uint32_t* s_tpos = g.crc_tpos;                       // x^(pos) mod g per A position: 23 KB, read through L1
  uint32_t* s_crc = s_permw + plane / 2;                     // per slot: CRC accumulator
  int* s_next = reinterpret_cast<int*>(s_crc + nflag);       // per slot: next work item; [nflag]: slots with work
  uint32_t* s_slots = s_crc + 2 * nflag + 4;                 // per slot: A, plane/2 words
  int* s_active = s_next + nflag;

  for (int i = tid; i < plane / 2; i += blockDim.x) s_permw[i] = reinterpret_cast<const uint32_t*>(g.perm_pos)[i];

  SlotCtx c;
  // slot stride = plane/2 words plus a skew that makes consecutive slots continue the bank sequence
  // (base(s+1) = base(s) + T mod 32): a warp that straddles two slots then stays conflict-free
  c.Aw = s_slots + (size_t)(valid ? slot : 0) * g.slot_words;
  c.gslot = blockIdx.x * g.ncb_cta + (valid ? slot : 0);
  c.in4 = nullptr;
  const uint16_t* perm16 = reinterpret_cast<const uint16_t*>(s_permw) + 2 * kSW * t;

  // ---- persistent slots ------------------------------------------------------------------------------
  int cur = blockIdx.x * g.ncb_cta + slot;                    // the first assignment is static
  bool have = valid && cur < g.n_cb;
  long long cbi = 0;
  int it = 0;
  auto init_slot = [&]() {
    cbi = g.cb_list ? g.cb_list[cur] : cur;
    c.in4 = reinterpret_cast<const uint4*>(g.in + cbi * g.in_stride);
    // a-priori LLRs start at zero: each thread clears its own column of A
    for (int i = 0; i < W; i++) c.Aw[i * T + t] = 0u;
    // the block comes straight from HBM: request all three planes now (the groups the backward sweep needs
    // first go first) so that only the first sub-window waits for DRAM
    const uint4* p0 = c.in4 + 2 * t;
    for (int sw = W / kSW - 1; sw >= 0; sw--) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + sw * 2 * T));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (plane / 8) + sw * 2 * T));
    }
    for (int sw = W / kSW - 1; sw >= 0; sw--) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + 2 * (plane / 8) + sw * 2 * T));
  };
  if (have) init_slot();
  if (tid == 0) *s_active = 0;
  __syncthreads();
  if (have && t == 0) atomicAdd(s_active, 1);
  __syncthreads();

  while (*s_active > 0) {
    if (valid && t == 0) s_crc[slot] = 0;
    if (have) map_pass<0, CRC>(g, c, perm16, s_tpos, t, it);
    __syncthreads();
    if (have) {
      const uint32_t part = map_pass<1, CRC>(g, c, perm16, s_tpos, t, it);
      if (crc_on) atomicXor(&s_crc[slot], part);
    }
    __syncthreads();
    const bool crc_ok = crc_on && valid && s_crc[slot] == 0;
    const bool fin = have && (crc_ok || it + 1 >= g.max_iter);
    if (fin) {
      // ---- pack the hard decisions of this iteration, MSB first, natural order --------------------
      uint8_t* out = g.out_bits + cbi * (long long)g.out_stride;
      const int wbytes = W / 8;
      const uint32_t* bp = reinterpret_cast<const uint32_t*>(g.bits_scratch) + ((size_t)c.gslot * plane) / 2 + t;
#pragma unroll 4
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
      if (t == 0) {
        g.out_status[cbi] = (it + 1) | ((crc_ok ? 1 : 0) << 8);
        const int nxt = g.work_base + atomicAdd(g.work_counter, 1);
        s_next[slot] = nxt;
        if (nxt >= g.n_cb) atomicSub(s_active, 1);      // this slot retires
      }
    } else if (have) {
      it++;
    }
    __syncthreads();
    if (fin) {
      cur = s_next[slot];
      have = cur < g.n_cb;
      it = 0;
      if (have) init_slot();
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_kernel(const TurboArgs g) { turbo_decode_body<false>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_crc_kernel(const TurboArgs g) { turbo_decode_body<true>(g); }

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

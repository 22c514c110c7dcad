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
//                   LDS.U16 by the contention-free property); the DEC2 position table of this K laid out
//                   [step][window parity][thread] so that a warp reads consecutive 16-bit entries; the hard
//                   decisions of the current iteration, one bit per trellis step in DEC2 order;
//   L2 / HBM      : the channel LLRs (read-only, thread-contiguous "tcb" layout -> coalesced LDG.128,
//                   register-prefetched half a sub-window ahead), the thread-private beta checkpoints of the
//                   current pass, the window-boundary metrics (NII), the per-step CRC contributions.
// Each MAP pass is: backward sweep storing beta every 8 steps, then forward sweep that re-creates beta
// for 8 steps in registers and produces alpha, the extrinsic and (DEC2) the hard decision, whose CRC
// contribution x^(position) mod g is accumulated on the fly (the CRC is linear over GF(2)).  Hard decisions
// leave the SM once per code block: the slot that finishes scatters its decision bits into its (now free)
// exchange array at the natural positions and packs them from there.
//
// Issue budget (per thread and trellis step, both passes of a half iteration; see DESIGN.md "K5 in detail"):
// ALU pipe 8 (beta) + 7 (beta again) + 8 (alpha) + 14 (extrinsic) + 1 (clamp) + DEC2 only: 2 masks, 2 CRC, 1 half
// extraction; FMA pipe: the gamma additions, normalisations and address arithmetic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {

constexpr int kSW = 8;                 // sub-window: beta values re-created in registers
constexpr int kL2Ahead = 3;   // backward sweep: groups between an L2 prefetch and the copy that needs the lines
constexpr uint32_t kNegInfPair = ((uint32_t)(uint16_t)(-kTdInf) << 16) | (uint16_t)(-kTdInf);
constexpr uint32_t pair16(int v) { return ((uint32_t)(uint16_t)v << 16) | (uint32_t)(uint16_t)v; }
constexpr uint32_t k2EPair = pair16(2 * kTdE), kNegEPair = pair16(-kTdE), kEp1Pair = pair16(kTdE + 1);

__device__ __forceinline__ uint32_t vadd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
// max(a + b, c) per int16 half, the add wrapping (VIADDMNMX.S16x2)
__device__ __forceinline__ uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
// max(min(a + b, 2E), 0) per int16 half: with a + b = v + E this is clamp(v, -E, E) + E in ONE instruction
// (VIADDMNMX.S16x2.RELU) instead of a min and a max
__device__ __forceinline__ uint32_t vclamp2E(uint32_t a, uint32_t b) { return __viaddmin_s16x2_relu(a, b, k2EPair); }
// lo + hi * 65536 with lo, hi zero-extended 16-bit loads: one IMAD on the FMA pipe (a PRMT would take an
// ALU-pipe slot, and the ALU pipe is the one this kernel saturates)
__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) { return hi * 65536u + lo; }
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

// Normalisation (SPEC 7.4 subtracts m(0)).  A packed subtraction costs sm_100a a NOT on the ALU pipe plus two adds, so
// the kernel adds ~m(0) = -m(0) - 1 instead: every metric of the vector ends up ONE LOWER than the oracle's, m(0) = -1.
// The extrinsic max(...) - max(...) and the decision only see differences inside one alpha and one beta vector, so the
// outputs are unchanged; the no-wrap bound of SPEC 7.6 has a margin of 1284, the offset costs 2 of it.
// `m1` is the constant (-1, -1) = 0xFFFFFFFF read from the kernel parameters, so that neither nvcc nor ptxas knows its
// value: the bitwise NOT is then written v * m1 + m1 = -v - 1 = ~v, ONE IMAD on the FMA pipe instead of a LOP3 on the ALU
// pipe, which is the pipe this kernel saturates (a literal constant is folded back into LOP3 / IADD3).
__device__ __forceinline__ uint32_t vnot(uint32_t v, uint32_t m1) { return v * m1 + m1; }
__device__ __forceinline__ void normalise(uint32_t (&m)[8], uint32_t m1) {
  const uint32_t n0 = vnot(m[0], m1);
#pragma unroll
  for (int s = 1; s < 8; s++) m[s] = vadd(m[s], n0);
  m[0] = m1;
}

// The four branch-label maxima of one trellis step from alpha_k and beta_{k+1} (SPEC 7.4), combined to
//   l1 = max(A10, A11 + y),  l0 = max(A00, A01 + y);   extrinsic = l1 - l0
__device__ __forceinline__ void ext_parts(const uint32_t (&a)[8], const uint32_t (&bn)[8], uint32_t y, uint32_t& l1, uint32_t& l0) {
#ifdef SRSUE_TURBO_EXT_MAX3
  // Variant: three of the four sums of every label by plain adds (FMA pipe) and one three-input maximum
  // (VIMNMX3.S16x2), the fourth by add-max: 10 ALU-pipe + 13 FMA-pipe instructions instead of 14 + 5.
  const uint32_t a00 = vaddmax(a[7], bn[3], __vimax3_s16x2(vadd(a[0], bn[0]), vadd(a[1], bn[4]), vadd(a[6], bn[7])));
  const uint32_t a11 = vaddmax(a[7], bn[7], __vimax3_s16x2(vadd(a[0], bn[4]), vadd(a[1], bn[0]), vadd(a[6], bn[3])));
  const uint32_t a01 = vaddmax(a[5], bn[6], __vimax3_s16x2(vadd(a[2], bn[5]), vadd(a[3], bn[1]), vadd(a[4], bn[2])));
  const uint32_t a10 = vaddmax(a[5], bn[2], __vimax3_s16x2(vadd(a[2], bn[1]), vadd(a[3], bn[5]), vadd(a[4], bn[6])));
#else
  uint32_t a00 = vadd(a[0], bn[0]);
  a00 = vaddmax(a[1], bn[4], a00); a00 = vaddmax(a[6], bn[7], a00); a00 = vaddmax(a[7], bn[3], a00);
  uint32_t a11 = vadd(a[0], bn[4]);
  a11 = vaddmax(a[1], bn[0], a11); a11 = vaddmax(a[6], bn[3], a11); a11 = vaddmax(a[7], bn[7], a11);
  uint32_t a01 = vadd(a[2], bn[5]);
  a01 = vaddmax(a[3], bn[1], a01); a01 = vaddmax(a[4], bn[2], a01); a01 = vaddmax(a[5], bn[6], a01);
  uint32_t a10 = vadd(a[2], bn[1]);
  a10 = vaddmax(a[3], bn[5], a10); a10 = vaddmax(a[4], bn[6], a10); a10 = vaddmax(a[5], bn[2], a10);
#endif
  l1 = vaddmax(a11, y, a10);
  l0 = vaddmax(a01, y, a00);
}

// PRMT with the sign-replicate bit of every selector nibble set (__byte_perm masks that bit off)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t sign_fill(uint32_t w) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"(SEL));
  return r;
}

// Asynchronous 16-byte copy global -> shared (LDGSTS, L2 only).  The channel LLRs and checkpoints of the NEXT group
// travel this way: a register-destination load would be sunk by ptxas to its first use (observed: zero prefetch
// distance, a third of all stall samples), a copy with no destination register needs no register to be kept live.
// Staging chunks are addressed by their 32-bit shared-window address so that every access is an LDS/LDGSTS, never a
// generic load.
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem) : "memory");
}
// base + idx * 16 bytes as ONE IMAD.WIDE on the FMA pipe (the compiler's own 64-bit address arithmetic is three or four
// LEA / IADD3.X on the ALU pipe, which is the pipe this kernel saturates)
__device__ __forceinline__ const uint4* at16(const uint4* base, uint32_t idx) {
  uint64_t a;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(idx), "l"(base));
  return reinterpret_cast<const uint4*>(a);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void lds8(uint32_t saddr, uint32_t stride, uint32_t (&v)[8]) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(saddr) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(saddr + stride) : "memory");
}
// One word through shared memory with volatile accesses.  ptxas keeps memory operations in order around the copies but
// hoists every arithmetic instruction it can above them, which leaves the copies at the very end of the loop body, right
// before the wait.  Reading one input of the following trellis step back from shared memory AFTER the copies ties the
// arithmetic of the second half of a group to that point, so the copies really are issued half a group ahead.
__device__ __forceinline__ void pin_store(uint32_t saddr, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t pin_load(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds16(const unsigned char* base, uint32_t off) {
  return *reinterpret_cast<const uint16_t*>(base + off);
}
__device__ __forceinline__ void sts16(unsigned char* base, uint32_t off, uint32_t v) {
  *reinterpret_cast<uint16_t*>(base + off) = (uint16_t)v;
}

// What one slot thread needs to find its code block and its scratch; everything else is derived in place
// (keeps the live state across the long unrolled sweeps small).
struct SlotCtx {
  const uint4* in4;       // code block in tcb layout (global)
  uint32_t* Aw;           // shared: extrinsic exchange array [W][Ppad] as packed pairs, word i*T + t
  uint16_t* bits;         // global: hard decisions of the block in DEC2 order, [W/8][T] x (byte of window 2t | byte of window 2t+1 << 8)
  uint32_t pin;           // shared-window address of this thread's scratch word (pin_store / pin_load)
  uint32_t stage;         // shared-window address of this thread's staging chunks: chunk k (16 bytes) at stage + k * 16 * blockDim.x
                          // (forward sweep: y 0,1  sys 2,3  checkpoint 4,5; backward sweep: two sets 0..3 / 4..7)
  uint32_t gslot;         // global slot number: selects the NII / checkpoint scratch
};

// Window-boundary metrics (NII, SPEC 7.3) live in global memory as one 16-byte record per window and kind: the eight
// int16 state metrics alpha reached at the END of window j (record j + 1 of kind 0) or beta reached at its START (record j
// of kind 1).  A thread reads / writes the records of its two windows with 128-bit accesses and transposes them to / from
// the packed (window 2t | window 2t+1 << 16) registers with byte permutes -- no 16-bit gathers, no branches.
__device__ __forceinline__ void nii_unpack(const uint4& lo, const uint4& hi, uint32_t (&m)[8]) {
  m[0] = __byte_perm(lo.x, hi.x, 0x5410); m[1] = __byte_perm(lo.x, hi.x, 0x7632);
  m[2] = __byte_perm(lo.y, hi.y, 0x5410); m[3] = __byte_perm(lo.y, hi.y, 0x7632);
  m[4] = __byte_perm(lo.z, hi.z, 0x5410); m[5] = __byte_perm(lo.z, hi.z, 0x7632);
  m[6] = __byte_perm(lo.w, hi.w, 0x5410); m[7] = __byte_perm(lo.w, hi.w, 0x7632);
}
__device__ __forceinline__ void nii_pack(const uint32_t (&m)[8], uint4& lo, uint4& hi) {
  lo.x = __byte_perm(m[0], m[1], 0x5410); hi.x = __byte_perm(m[0], m[1], 0x7632);
  lo.y = __byte_perm(m[2], m[3], 0x5410); hi.y = __byte_perm(m[2], m[3], 0x7632);
  lo.z = __byte_perm(m[4], m[5], 0x5410); hi.z = __byte_perm(m[4], m[5], 0x7632);
  lo.w = __byte_perm(m[6], m[7], 0x5410); hi.w = __byte_perm(m[6], m[7], 0x7632);
}

// One max-log-MAP pass of constituent decoder DEC (0 or 1) for the two windows of this thread.  Returns the
// thread's CRC contribution of the hard decisions (DEC2 with CRC), else 0.  perm_t = position table + t.
template <int DEC, bool CRC, int TS, int TT, bool TWO>
__device__ __forceinline__ uint32_t map_pass(const TurboArgs& g, const SlotCtx& c, const uint16_t* perm_t, int t, int it) {
  const int T = TT ? TT : g.T, W = g.W, P = g.P, NP = g.Ppad + 2, nsw = W / kSW, plane = g.plane;
  const int j0 = 2 * t, j1 = 2 * t + 1;
  const int rd = it & 1, wr = rd ^ 1;
  // records: [slot][dec][parity][kind][NP] uint4
  uint4* nii = g.nii + (size_t)c.gslot * (size_t)(2 * 2 * 2 * NP);
  const uint4* nii_a_rd = nii + ((DEC * 2 + rd) * 2 + 0) * NP;
  const uint4* nii_b_rd = nii + ((DEC * 2 + rd) * 2 + 1) * NP;
  uint4* nii_a_wr = nii + ((DEC * 2 + wr) * 2 + 0) * NP;
  uint4* nii_b_wr = nii + ((DEC * 2 + wr) * 2 + 1) * NP;
  // tcb planes and checkpoints are laid out [group][half][thread] x 16 bytes: the chunk a warp copies with one LDGSTS
  // is contiguous (whole 32-byte sectors, half as many L2 requests as a thread-contiguous 32-byte layout)
  const int gstride = 2 * T;                       // uint4 per 8-step group
  const uint4* sysq = c.in4 + t;
  const uint4* yq = sysq + (DEC ? 2 : 1) * (plane / 8);
  const int slot_in_cta = (int)(c.gslot - blockIdx.x * g.ncb_cta);
  const int cstride = g.ncb_cta * gstride;         // uint4 per sub-window of the whole CTA: [half][slot][thread]
  const int chalf = g.ncb_cta * T;
  uint4* ckpt4 = g.ckpt + (size_t)blockIdx.x * (size_t)(nsw * cstride) + slot_in_cta * T + t;
  unsigned char* Ab = reinterpret_cast<unsigned char*>(c.Aw);
  const uint32_t m1 = g.ones;                      // (-1, -1), see normalise()

  // The channel LLRs (and, forward, the checkpoint) of a group reach the thread through its private staging chunks in
  // shared memory: requested with cp.async at the top of the previous group (a whole group of issue slots ahead), read
  // into registers with LDS.128 at the top of the group.  Nobody else touches the chunks, so no barrier is involved.
  const uint32_t cstr = 16u * blockDim.x;          // bytes between the chunks of one thread
  const uint32_t stg = c.stage, pin = c.pin;
  auto fetch = [&](int sw, bool with_ckpt) {
    const uint32_t gi = (uint32_t)(sw * gstride);
    cp_async16(stg, at16(yq, gi));
    cp_async16(stg + cstr, at16(yq, gi + T));
    if (DEC == 0) {
      cp_async16(stg + 2 * cstr, at16(sysq, gi));
      cp_async16(stg + 3 * cstr, at16(sysq, gi + T));
    }
    if (with_ckpt) {
      const uint32_t ci = (uint32_t)(sw * cstride);
      cp_async16(stg + 4 * cstr, at16(ckpt4, ci));
      cp_async16(stg + 5 * cstr, at16(ckpt4, ci + chalf));
    }
    cp_async_commit();
  };
  // L2 prefetch of the backward sweep: the two (DEC2: one) plane rows of a group and slot are contiguous runs of
  // 32 T bytes; thread t asks for line t of the first run, thread nl + t for line t of the second
  const int nl = (T + 3) / 4;                      // 128-byte lines per run
  const uint4* pf_base = (t < nl) ? (yq - t) + 8 * t : (sysq - t) + 8 * (t - nl);
  const bool pf_on = t < (DEC == 0 ? 2 * nl : nl) && (t < nl ? 8 * t : 8 * (t - nl)) < 2 * T;
  // Backward sweep: TWO groups in flight (its trip is a third as long as the forward sweep's, one group of distance did not
  // cover an L2 miss).  Two sets of chunks alternate: {0..3} and {4..7} (y at +0, +1; systematic at +2, +3).
  auto fetch_bwd = [&](int sw, uint32_t base) {
    const uint32_t gi = (uint32_t)(sw * gstride);
    cp_async16(base, at16(yq, gi));
    cp_async16(base + cstr, at16(yq, gi + T));
    if (DEC == 0) {
      cp_async16(base + 2 * cstr, at16(sysq, gi));
      cp_async16(base + 3 * cstr, at16(sysq, gi + T));
    }
    cp_async_commit();
  };
  constexpr bool two = TWO;                         // (one group in flight where the eight chunks would cost a code-block slot)
  fetch_bwd(nsw - 1, stg);
  if (two) fetch_bwd(nsw - 2, stg + 4 * cstr);

  // ---- boundary metrics: all four records are requested at once
  uint32_t b[8], a[8];
  {
    if (it) {
      const uint4 ra0 = nii_a_rd[j0], ra1 = nii_a_rd[j0 + 1], rb0 = nii_b_rd[j0 + 1], rb1 = nii_b_rd[j0 + 2];
      nii_unpack(ra0, ra1, a);
      nii_unpack(rb0, rb1, b);
    } else {                                        // first iteration: all boundary metrics are zero, nothing to read
#pragma unroll
      for (int s = 0; s < 8; s++) a[s] = b[s] = 0u;
    }
    if (j0 == 0) {                                  // trellis start: alpha_0 = (0, -INF, ...)
      a[0] &= 0xFFFF0000u;
#pragma unroll
      for (int s = 1; s < 8; s++) a[s] = (a[s] & 0xFFFF0000u) | (uint32_t)(uint16_t)(-kTdInf);
    }
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
      normalise(bt, m1);                              // metric index K is a multiple of 4
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = last0 ? ((b[s] & 0xFFFF0000u) | (bt[s] & 0xFFFFu)) : ((b[s] & 0xFFFFu) | (bt[s] & 0xFFFF0000u));
    }
  }

  uint32_t Cy[kSW], Cs[kSW], Cb[8];

  // ---- pass 1: backward sweep, checkpoint beta every kSW steps -----------------------------------
#pragma unroll 1
  for (int sw = nsw - 1; sw >= 0; sw--) {
    const uint32_t cur = stg + ((two && ((nsw - 1 - sw) & 1)) ? 4 * cstr : 0u);
    if (two && sw) cp_async_wait_but_one(); else cp_async_wait_all();
    lds8(cur, cstr, Cy);
    if (DEC == 0) lds8(cur + 2 * cstr, cstr, Cs);
    pin_store(pin, Cy[kSW - 1]);
    {                                               // checkpoint: thread-private scratch, read back in pass 2
      const uint32_t ci = (uint32_t)(sw * cstride);
      *const_cast<uint4*>(at16(ckpt4, ci)) = make_uint4(b[0], b[1], b[2], b[3]);
      *const_cast<uint4*>(at16(ckpt4, ci + chalf)) = make_uint4(b[4], b[5], b[6], b[7]);
    }
    if (sw == 0) fetch(0, true);                    // after group 0: its data and checkpoint again, for the forward sweep
    else if (!two) fetch_bwd(sw - 1, stg);
    else if (sw >= 2) fetch_bwd(sw - 2, cur);       // into the set just read
    // The resident code blocks of all SMs together exceed what the L2 keeps (hit rate 80 %): ask for the lines of the
    // group three ahead now, so that its copy finds them in L2 instead of waiting for DRAM with nothing to overlap.
    if (sw >= kL2Ahead && pf_on) asm volatile("prefetch.global.L2 [%0];" ::"l"(at16(pf_base, (uint32_t)((sw - kL2Ahead) * gstride))));
    Cy[kSW - 1] = pin_load(pin);                    // ties the arithmetic below to this point (see pin_store)
    // x of a group: DEC1 systematic + a-priori (natural order, one LDS.32 for both windows); DEC2 the two
    // interleaved positions of the exchange array
    uint32_t x[kSW];
    if (DEC == 0) {
      if (it) {                                     // a-priori LLRs are zero in the first iteration: nothing to read
        const uint32_t* ap = c.Aw + sw * kSW * T + t;
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = vadd(Cs[i], ap[i * T]);
      } else {
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = Cs[i];
      }
    } else {
      const uint16_t* pq = perm_t + sw * (2 * kSW) * TS;
#pragma unroll
      for (int i = 0; i < kSW; i++) x[i] = pack16(lds16(Ab, pq[(2 * i) * TS]), lds16(Ab, pq[(2 * i + 1) * TS]));
    }
#pragma unroll
    for (int i = kSW - 1; i >= 0; i--) {
      uint32_t nb[8];
      beta_step(b, nb, x[i], Cy[i], vadd(x[i], Cy[i]));
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = nb[s];
      if ((i & 3) == 0) normalise(b, m1);
    }
  }
  // beta at the window start feeds the previous window in the next iteration
  {
    uint4 lo, hi;
    nii_pack(b, lo, hi);
    nii_b_wr[j0] = lo;
    nii_b_wr[j0 + 1] = hi;
  }

  // ---- pass 2: forward sweep ----------------------------------------------------------------------
  uint32_t crc = 0;
  const uint4* crcq = reinterpret_cast<const uint4*>(g.crc_lin) + t;      // [sw][4][T] uint4
#pragma unroll 1
  for (int sw = 0; sw < nsw; sw++) {
    cp_async_wait_all();
    lds8(stg, cstr, Cy);
    if (DEC == 0) lds8(stg + 2 * cstr, cstr, Cs);
    lds8(stg + 4 * cstr, cstr, Cb);
    pin_store(pin, Cy[kSW - 1]);
    if (sw + 1 < nsw) fetch(sw + 1, true);
    Cy[kSW - 1] = pin_load(pin);
    uint32_t x[kSW];
    uint32_t pa[2 * kSW];                           // DEC2: byte offsets of the two interleaved positions per step
    uint32_t* ap = c.Aw + sw * kSW * T + t;
    if (DEC == 0) {
      if (it) {
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = vadd(Cs[i], ap[i * T]);
      } else {
#pragma unroll
        for (int i = 0; i < kSW; i++) x[i] = Cs[i];
      }
    } else {
      const uint16_t* pq = perm_t + sw * (2 * kSW) * TS;
#pragma unroll
      for (int i = 0; i < 2 * kSW; i++) pa[i] = pq[i * TS];
#pragma unroll
      for (int i = 0; i < kSW; i++) x[i] = pack16(lds16(Ab, pa[2 * i]), lds16(Ab, pa[2 * i + 1]));
    }
    uint32_t B[kSW][8];                             // B[i] = beta_{i+1} of this sub-window
#pragma unroll
    for (int s = 0; s < 8; s++) B[kSW - 1][s] = Cb[s];
#pragma unroll
    for (int i = kSW - 1; i >= 1; i--) {
      beta_step(B[i], B[i - 1], x[i], Cy[i], vadd(x[i], Cy[i]));
      if ((i & 3) == 0) normalise(B[i - 1], m1);
    }
    uint4 tq[4];                                    // CRC contributions of the 16 decisions of this group (L1-resident table)
    uint32_t acc0 = 0xFFu, acc1 = 0xFFu;            // decision bits of the two windows, step i at bit 7-i: start from all
                                                    // ones, every zero decision subtracts its bit
#pragma unroll
    for (int i = 0; i < kSW; i++) {
      if (DEC == 1 && CRC && (i & 3) == 0) {
        tq[i >> 1] = __ldg(crcq + ((size_t)sw * 4 + (i >> 1)) * T);
        tq[(i >> 1) + 1] = __ldg(crcq + ((size_t)sw * 4 + (i >> 1) + 1) * T);
      }
      // extrinsic = l1 - l0 without a packed subtraction (-l0 = ~l0 + 1): q = l1 + ~l0 = ext - 1, exact in 16 bits
      uint32_t l1, l0;
      ext_parts(a, B[i], Cy[i], l1, l0);
      const uint32_t q = vadd(l1, vnot(l0, m1));
      const uint32_t r = vclamp2E(q, kEp1Pair);                // clamp(ext, -E, E) + E
      if (DEC == 0) {
        ap[i * T] = vadd(vadd(Cs[i], kNegEPair), r);
      } else {
        const uint32_t la = vadd(r, kNegEPair);
        sts16(Ab, pa[2 * i], la);
        sts16(Ab, pa[2 * i + 1], la >> 16);
        // decision = (x + ext) > 0  <=>  x + ext - 1 >= 0: the masks below are set where the bit is ZERO
        const uint32_t s1 = vadd(q, x[i]);
        const uint32_t z0 = sign_fill<0x9999>(s1), z1 = sign_fill<0xBBBB>(s1);
        acc0 = z0 * (0x80u >> i) + acc0;
        acc1 = z1 * (0x80u >> i) + acc1;
        if (CRC) {
          const uint32_t t0 = (i & 1) ? tq[i >> 1].z : tq[i >> 1].x, t1 = (i & 1) ? tq[i >> 1].w : tq[i >> 1].y;
          crc ^= (~z0 & t0) ^ (~z1 & t1);
        }
      }
      alpha_step(a, x[i], Cy[i], vadd(x[i], Cy[i]));
      if ((i & 3) == 3) normalise(a, m1);
    }
    if (DEC == 1) c.bits[sw * T + t] = (uint16_t)(acc0 | (acc1 << 8));
  }
  {
    uint4 lo, hi;
    nii_pack(a, lo, hi);
    nii_a_wr[j0 + 1] = lo;
    nii_a_wr[j0 + 2] = hi;
  }
  return crc;
}

// TT: threads per code block as a compile-time constant (0: run time).  With T known, every stride of the exchange array,
// the channel-LLR planes and the checkpoints is an immediate of the load / store that uses it.
template <bool CRC, int TS, int TT, bool TWO>
__device__ __forceinline__ void turbo_decode_body(const TurboArgs& g) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int T = TT ? TT : g.T, W = g.W;
  const int tid = threadIdx.x;
  // Phase groups: the slots of a CTA are split into g.ngroups groups of whole warps that only synchronise among themselves
  // (named barriers).  The second group starts late, so while one group is in a load-dominated backward sweep, at a
  // boundary-metric exchange or refilling slots, the other one is in an arithmetic-bound forward sweep and takes the
  // issue slots the first cannot use.
  const int grp = tid / g.group_threads, lt = tid - grp * g.group_threads;
  const int ls = lt / T, t = lt - ls * T;
  const int slot = grp * g.group_slots + ls;
  const bool valid = ls < g.group_slots && slot < g.ncb_cta;
  const int bar_id = 1 + grp, bar_n = g.group_threads;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory"); };
  constexpr bool crc_on = CRC;
  const int nflag = (g.ncb_cta + 3) & ~3;

  uint32_t* s_permw = smem;                                  // W * TS words: DEC2 positions [W][2][TS] x u16 (byte offsets into A)
  uint32_t* s_crc = s_permw + W * TS;                        // per slot: CRC accumulator
  int* s_next = reinterpret_cast<int*>(s_crc + nflag);       // per slot: next work item; [nflag]: slots with work
  uint32_t* s_slots = s_crc + 2 * nflag + 4;                 // per slot: A, plane/2 words (+ skew)
  int* s_active = s_next + nflag + grp;                      // per phase group: slots with work
  // staging chunks of all threads, 16-byte aligned: [8][blockDim.x] uint4, then [blockDim.x] scratch words
  uint4* s_stage = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(s_slots + (size_t)g.ncb_cta * g.slot_words) + 15) & ~(uintptr_t)15);

  {
    // position table: global [W][2][T] -> shared [W][2][TS] (rows padded to the warp-friendly stride)
    uint16_t* d = reinterpret_cast<uint16_t*>(s_permw);
    for (int e = tid; e < 2 * W * T; e += blockDim.x) {
      const int row = e / T, col = e - row * T;
      d[row * TS + col] = g.perm_tab[e];
    }
  }

  SlotCtx c;
  // slot stride = plane/2 words plus a skew that makes consecutive slots continue the bank sequence
  // (base(s+1) = base(s) + T mod 32): a warp that straddles two slots then stays conflict-free
  c.Aw = s_slots + (size_t)(valid ? slot : 0) * g.slot_words;
  c.bits = nullptr;
  c.stage = (uint32_t)__cvta_generic_to_shared(s_stage + tid);
  c.pin = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<uint32_t*>(s_stage + (TWO ? 8 : 6) * blockDim.x) + tid);
  c.gslot = blockIdx.x * g.ncb_cta + (valid ? slot : 0);
  c.in4 = nullptr;
  const uint16_t* perm_t = reinterpret_cast<const uint16_t*>(s_permw) + t;

  // ---- persistent slots ------------------------------------------------------------------------------
  int cur = blockIdx.x * g.ncb_cta + slot;                    // the first assignment is static
  bool have = valid && cur < g.n_cb;
  long long cbi = 0;
  int it = 0;
  auto init_slot = [&]() {
    cbi = g.cb_list ? g.cb_list[cur] : cur;
    c.in4 = reinterpret_cast<const uint4*>(g.in + cbi * g.in_stride);
    c.bits = reinterpret_cast<uint16_t*>(g.dbits + (size_t)cur * g.dbits_stride);
    // No prefetch of the new block here: asking the L2 for all of it at once (or for its head one pass early) measured
    // SLOWER -- the resident blocks of all SMs already overflow the L2, and lines that arrive long before their use push
    // out lines that are needed sooner.  The backward sweep asks for its groups kL2Ahead groups ahead instead.
  };
  if (lt == 0) *s_active = 0;
  __syncthreads();
  if (have && t == 0) atomicAdd(s_active, 1);
  __syncthreads();
  if (grp && g.phase_delay > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)g.phase_delay) __nanosleep(500);
  }
  if (have) init_slot();

  while (*reinterpret_cast<volatile int*>(s_active) > 0) {
    if (valid && t == 0) s_crc[slot] = 0;
    if (have) map_pass<0, CRC, TS, TT, TWO>(g, c, perm_t, t, it);
    group_sync();
    if (have) {
      const uint32_t part = map_pass<1, CRC, TS, TT, TWO>(g, c, perm_t, t, it);
      if (crc_on) atomicXor(&s_crc[slot], part);
    }
    group_sync();
    const bool crc_ok = crc_on && valid && s_crc[slot] == 0;
    const bool fin = have && ((crc_ok && it + 1 >= g.min_iter) || it + 1 >= g.max_iter);
    if (fin) {
      // The hard decisions of this iteration already lie in global memory in DEC2 order (the forward sweep stores them
      // group by group); tdec_deinterleave_kernel turns them into natural-order bytes after this kernel.
      if (t == 0) {
        g.out_status[cbi] = (it + 1) | ((crc_ok ? 1 : 0) << 8);
        const int nxt = g.work_base + atomicAdd(g.work_counter, 1);
        s_next[slot] = nxt;
        if (nxt >= g.n_cb) atomicSub(s_active, 1);      // this slot retires
      }
    } else if (have) {
      it++;
    }
    group_sync();
    if (fin) {
      cur = s_next[slot];
      have = cur < g.n_cb;
      it = 0;
      if (have) init_slot();
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_kernel(const TurboArgs g) { turbo_decode_body<false, 32, 0, false>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_crc_kernel(const TurboArgs g) { turbo_decode_body<true, 32, 0, false>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_wide_kernel(const TurboArgs g) { turbo_decode_body<false, 64, 0, false>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_crc_wide_kernel(const TurboArgs g) { turbo_decode_body<true, 64, 0, false>(g); }
// the two largest code-block sizes carry most of the bits of a wide-band transport block: K = 5824 (T = 26; its 14 slots leave
// room for the eight staging chunks of a backward sweep with two groups in flight), K = 6144 (T = 24)
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_t26_kernel(const TurboArgs g) { turbo_decode_body<false, 32, 26, true>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_crc_t26_kernel(const TurboArgs g) { turbo_decode_body<true, 32, 26, true>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_t24_kernel(const TurboArgs g) { turbo_decode_body<false, 32, 24, false>(g); }
__global__ void __launch_bounds__(kTurboMaxThreads, 1) turbo_decode_crc_t24_kernel(const TurboArgs g) { turbo_decode_body<true, 32, 24, false>(g); }

// ---- hard decisions: DEC2 order -> natural order ------------------------------------------------------
// The decoder leaves the decisions of a code block as it produced them: one bit per trellis step of the second
// constituent decoder, [W/8][T] x (byte of window 2t | byte of window 2t+1 << 8), step i of a group at bit 7 - i.  This
// kernel applies the inverse QPP permutation at full occupancy (inside the decoder the same work ran at 12 warps per SM
// and held three barriers): a CTA keeps the per-K table `deint` (for natural bit n the index of its source bit, counted
// LSB first within each source byte) in shared memory, spreads the row of a code block to one byte per bit and
// gathers the K / 8 output bytes, MSB first.  cb_list maps launch-local rows to output rows as in the decoder.
__global__ void __launch_bounds__(256) tdec_deinterleave_kernel(const uint8_t* __restrict__ dbits, int dbits_stride,
                                                                const uint16_t* __restrict__ deint, const int32_t* __restrict__ cb_list,
                                                                int n_cb, uint8_t* __restrict__ out, int out_stride, int K, int row_bytes) {
  extern __shared__ __align__(16) uint32_t dsm[];
  uint16_t* s_tab = reinterpret_cast<uint16_t*>(dsm);                        // K entries (K is a multiple of 8)
  uint2* s_bit = reinterpret_cast<uint2*>(s_tab + ((K + 7) & ~7));           // row_bytes x 8 bytes: one byte per bit
  const unsigned char* s_bitb = reinterpret_cast<const unsigned char*>(s_bit);
  uint4* s_row = reinterpret_cast<uint4*>(s_bit + row_bytes);                // the row as it came (16-byte pieces)
  {
    const uint4* src = reinterpret_cast<const uint4*>(deint);
    uint4* dst = reinterpret_cast<uint4*>(s_tab);
    for (int e = threadIdx.x; e < K / 8; e += blockDim.x) dst[e] = __ldg(src + e);
  }
  // The row of the NEXT code block travels while this one is processed (rows are 16-byte aligned, a thread carries at most
  // one 16-byte piece: row_bytes <= 16 * blockDim.x): without it every block started with an exposed L2 / HBM round trip.
  const int row16 = (row_bytes + 15) / 16;
  uint4 piece = make_uint4(0u, 0u, 0u, 0u);
  long long cbi_next = 0;
  if ((int)blockIdx.x < n_cb) {
    if ((int)threadIdx.x < row16) piece = reinterpret_cast<const uint4*>(dbits + (size_t)blockIdx.x * dbits_stride)[threadIdx.x];
    cbi_next = cb_list ? cb_list[blockIdx.x] : (long long)blockIdx.x;
  }
  for (int cb = blockIdx.x; cb < n_cb; cb += gridDim.x) {
    __syncthreads();                                                         // table loaded / previous row consumed
    if ((int)threadIdx.x < row16) s_row[threadIdx.x] = piece;
    __syncthreads();
    for (int e = threadIdx.x; e < row_bytes; e += blockDim.x) {
      const uint32_t b = reinterpret_cast<const unsigned char*>(s_row)[e];
      // nibble * 0x00204081 puts bit q of the nibble at bit 8 q
      s_bit[e] = make_uint2(((b & 15u) * 0x00204081u) & 0x01010101u, ((b >> 4) * 0x00204081u) & 0x01010101u);
    }
    const long long cbi = cbi_next;
    {
      const int nx = cb + gridDim.x;
      if (nx < n_cb) {
        if ((int)threadIdx.x < row16) piece = reinterpret_cast<const uint4*>(dbits + (size_t)nx * dbits_stride)[threadIdx.x];
        cbi_next = cb_list ? cb_list[nx] : (long long)nx;
      }
    }
    __syncthreads();
    uint8_t* o = out + cbi * (long long)out_stride;
    for (int e = threadIdx.x; e < K / 8; e += blockDim.x) {
      const uint4 q = reinterpret_cast<const uint4*>(s_tab)[e];
      uint32_t v = s_bitb[q.x & 0xFFFFu];
      v = v * 2u + s_bitb[q.x >> 16];
      v = v * 2u + s_bitb[q.y & 0xFFFFu];
      v = v * 2u + s_bitb[q.y >> 16];
      v = v * 2u + s_bitb[q.z & 0xFFFFu];
      v = v * 2u + s_bitb[q.z >> 16];
      v = v * 2u + s_bitb[q.w & 0xFFFFu];
      v = v * 2u + s_bitb[q.w >> 16];
      o[e] = (uint8_t)v;
    }
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
      // tcb element -> (stream, window j, step i): element ((hs*T + t)*4 + ii)*2 + h, j = 2t + h, i = 4hs + ii
      const int stream = e / g.plane, r = e - stream * g.plane;
      const int h = r & 1, ii = (r >> 1) & 3, q = r >> 3;
      const int hs = q / g.T, t = q - hs * g.T;
      const int i = hs * 4 + ii, j = 2 * t + h;
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
      off = stream * g.plane + ((((i >> 2) * g.T + (j >> 1)) * 4 + (i & 3)) << 1) + (j & 1);
    } else {
      off = 3 * g.plane + (e - 3 * g.K);
    }
    dst[e] = src[off];
  }
}

}  // namespace srsue

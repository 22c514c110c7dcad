// pdcch.cu -- blind PDCCH decoding on the device: one warp per search-space candidate does the rate de-matching,
// a 64-state tail-biting Viterbi decoder and the RNTI-masked CRC16 (sm_100a).
//
// Replaces srsLTE's srslte_pdcch_decode_msg loop inside srslte_ue_dl_find_dl_dci_type
// (/root/reference/ue/src/phy/phch_worker.cc:293; UL grants :426 use the same search).  Arithmetic contract:
// oracle/SPEC.md section 10 -- everything here is integer, so the result does not depend on the schedule:
//   soft[3D]  = int32 sums of the E = 72 L candidate LLRs over the circular buffer (punctured positions 0);
//   Viterbi   = the D trellis steps three times in a row from all-zero int32 path metrics, branch metric
//               sum_j (code bit ? +soft : -soft), ties keep the predecessor whose dropped bit is 0, traceback from the
//               best final state (lowest index on ties), middle repetition is the output;
//   rem       = CRC16(payload) xor received CRC = the RNTI the message was masked with.
// Lane l owns the two states l (input history bit 5 = 0) and l + 32 (= 1); both have the predecessors 2l and 2l + 1.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {
__device__ __forceinline__ int parity7(int v) { return __popc(v) & 1; }
}  // namespace

// grid (n_sf), block 32 * n_cand: warp w decodes candidate w; thread 0 then picks the first match in candidate order
__global__ void __launch_bounds__(32 * kPdcchMaxCand) pdcch_search_kernel(const PdcchSearchArgs a) {
  // dynamic shared memory per candidate: soft[3D] int32, survivors of the last two repetitions [2D][2] u32, decisions [D]
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ int s_rem[kPdcchMaxCand];
  const int sf = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.nof_bits + 16, T = 3 * D;
  const int per = 3 * D + 4 * D + (D + 3) / 4;                 // words per candidate
  int32_t* soft = reinterpret_cast<int32_t*>(s_dyn + (size_t)w * per);
  uint32_t* surv = s_dyn + (size_t)w * per + 3 * D;            // [t - D][2]
  uint8_t* dec = reinterpret_cast<uint8_t*>(s_dyn + (size_t)w * per + 7 * D);
  if (w < a.n_cand) {
    const int L = a.cand_L[w], E = 72 * L;
    const int16_t* llr = a.llr + (size_t)sf * a.llr_stride + 72 * a.cand_ncce[w];
    for (int i = lane; i < 3 * D; i += 32) soft[i] = 0;
    __syncwarp();
    // circular buffer: position k of the candidate belongs to coded bit rm_seq[k mod 3D]
    for (int k = lane; k < E; k += 32) atomicAdd(&soft[a.rm_seq[k % (3 * D)]], (int32_t)llr[k]);
    __syncwarp();
    // code bits of the four transitions into this lane's two states: reg = (u << 6) | (2 lane + b)
    int sgn[2][2][3];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int reg = (u << 6) | (2 * lane + b);
        sgn[u][b][0] = parity7(reg & 0133); sgn[u][b][1] = parity7(reg & 0171); sgn[u][b][2] = parity7(reg & 0165);
      }
    int32_t pm0 = 0, pm1 = 0;                 // path metrics of states lane and lane + 32
    const int src0 = (2 * lane) & 31, src1 = (2 * lane + 1) & 31, hi = lane >> 4;
    for (int t = 0; t < T; t++) {
      const int k = t % D;
      const int32_t s0 = soft[k], s1 = soft[D + k], s2 = soft[2 * D + k];
      const int32_t a0 = __shfl_sync(0xFFFFFFFFu, pm0, src0), a1 = __shfl_sync(0xFFFFFFFFu, pm1, src0);
      const int32_t b0 = __shfl_sync(0xFFFFFFFFu, pm0, src1), b1 = __shfl_sync(0xFFFFFFFFu, pm1, src1);
      const int32_t pa = hi ? a1 : a0, pb = hi ? b1 : b0;       // metrics of predecessors 2 lane and 2 lane + 1
      int32_t nm[2];
      uint32_t sv[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int32_t bma = (sgn[u][0][0] ? s0 : -s0) + (sgn[u][0][1] ? s1 : -s1) + (sgn[u][0][2] ? s2 : -s2);
        const int32_t bmb = (sgn[u][1][0] ? s0 : -s0) + (sgn[u][1][1] ? s1 : -s1) + (sgn[u][1][2] ? s2 : -s2);
        const int32_t va = pa + bma, vb = pb + bmb;
        const bool take_b = vb > va;
        nm[u] = take_b ? vb : va;
        sv[u] = __ballot_sync(0xFFFFFFFFu, take_b);
      }
      pm0 = nm[0]; pm1 = nm[1];
      if (lane == 0 && t >= D) { surv[2 * (t - D)] = sv[0]; surv[2 * (t - D) + 1] = sv[1]; }
    }
    // best final state, lowest index on ties
    int32_t bv = pm0; int bi = lane;
    if (pm1 > bv) { bv = pm1; bi = lane + 32; }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const int32_t ov = __shfl_xor_sync(0xFFFFFFFFu, bv, off);
      const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncwarp();
    if (lane == 0) {
      int st = bi;
      for (int t = T - 1; t >= D; t--) {
        if (t < 2 * D) dec[t - D] = (uint8_t)(st >> 5);
        const uint32_t word = surv[2 * (t - D) + (st >> 5)];
        st = ((st & 31) << 1) | ((word >> (st & 31)) & 1u);
      }
      // CRC16 (x^16 + x^12 + x^5 + 1) of the payload, xor the received parity bits
      uint32_t reg = 0;
      for (int i = 0; i < a.nof_bits; i++) { reg = (reg << 1) | dec[i]; if (reg & 0x10000u) reg ^= 0x11021u; }
      for (int i = 0; i < 16; i++) { reg <<= 1; if (reg & 0x10000u) reg ^= 0x11021u; }
      uint32_t rx = 0;
      for (int i = 0; i < 16; i++) rx = (rx << 1) | dec[a.nof_bits + i];
      const int rem = (int)((reg ^ rx) & 0xFFFFu);
      if (a.rem) a.rem[(size_t)sf * a.n_cand + w] = (uint16_t)rem;
      // formats 0 and 1A share size and search space and differ in payload bit 0: a wrong flag is not a match
      s_rem[w] = (a.first_bit >= 0 && dec[0] != a.first_bit) ? -1 : rem;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int hit = -1;
    for (int c = 0; c < a.n_cand && hit < 0; c++) if (s_rem[c] == a.rnti) hit = c;
    int32_t* f = a.found + (size_t)sf * 4;
    f[0] = hit >= 0; f[1] = hit >= 0 ? a.cand_L[hit] : 0; f[2] = hit >= 0 ? a.cand_ncce[hit] : 0; f[3] = hit;
    if (hit >= 0) {
      const uint8_t* hd = reinterpret_cast<const uint8_t*>(s_dyn + (size_t)hit * per + 7 * D);
      for (int i = 0; i < a.nof_bits; i++) a.bits[(size_t)sf * 64 + i] = hd[i];
    }
  }
}

}  // namespace srsue

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
#include "viterbi.cuh"

namespace srsue {

// grid (n_sf), block 32 * n_cand: warp w decodes candidate w; thread 0 then picks the first match in candidate order
__global__ void __launch_bounds__(32 * kPdcchMaxCand) pdcch_search_kernel(const PdcchSearchArgs a) {
  // dynamic shared memory per candidate: soft[3D] int32, survivors of the last two repetitions [2D][2] u32, decisions [D]
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ int s_rem[kPdcchMaxCand];
  const int sf = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (a.row_filter && a.row_filter[sf] != a.row_want) {        // whole CTA: this subframe belongs to another control region
    if (threadIdx.x < 4) a.found[(size_t)sf * 4 + threadIdx.x] = 0;
    return;
  }
  const int D = a.nof_bits + 16;
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
    const int rem = viterbi_crc16_warp(soft, a.nof_bits, surv, dec, lane);
    if (lane == 0) {
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

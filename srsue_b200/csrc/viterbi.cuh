// viterbi.cuh -- warp-wide 64-state tail-biting Viterbi decoder of the LTE rate-1/3 convolutional code
// (K = 7, G = 133/171/165 octal) + CRC16, shared by the PDCCH search and the PBCH decoder.  oracle/SPEC.md 10:
// the D trellis steps run three times in a row from all-zero int32 path metrics, branch metric
// sum_j (code bit ? +soft : -soft), ties keep the predecessor whose dropped bit is 0, traceback from the best final
// state (lowest index on ties), the middle repetition is the output.  All-integer: the result is schedule-independent.
// Lane l owns the two states l (most recent input 0) and l + 32 (1); both have the predecessors 2l and 2l + 1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsue {

// soft: [3 D] int32 in shared memory (stream-major), surv: [2 D][2] words, dec: [D] bytes (both shared, per warp).
// Every lane of the warp calls this; returns (on every lane) CRC16(payload) xor received CRC; dec is valid after the
// call for all lanes (written by lane 0, followed by __syncwarp).
__device__ __forceinline__ int viterbi_crc16_warp(const int32_t* soft, int nof_bits, uint32_t* surv, uint8_t* dec, int lane) {
  const int D = nof_bits + 16, T = 3 * D;
  int sgn[2][2][3];
#pragma unroll
  for (int u = 0; u < 2; u++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int reg = (u << 6) | (2 * lane + b);
      sgn[u][b][0] = __popc(reg & 0133) & 1; sgn[u][b][1] = __popc(reg & 0171) & 1; sgn[u][b][2] = __popc(reg & 0165) & 1;
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
  int rem = 0;
  if (lane == 0) {
    int st = bi;
    for (int t = T - 1; t >= D; t--) {
      if (t < 2 * D) dec[t - D] = (uint8_t)(st >> 5);
      const uint32_t word = surv[2 * (t - D) + (st >> 5)];
      st = ((st & 31) << 1) | ((word >> (st & 31)) & 1u);
    }
    // CRC16 (x^16 + x^12 + x^5 + 1) of the payload, xor the received parity bits
    uint32_t reg = 0;
    for (int i = 0; i < nof_bits; i++) { reg = (reg << 1) | dec[i]; if (reg & 0x10000u) reg ^= 0x11021u; }
    for (int i = 0; i < 16; i++) { reg <<= 1; if (reg & 0x10000u) reg ^= 0x11021u; }
    uint32_t rx = 0;
    for (int i = 0; i < 16; i++) rx = (rx << 1) | dec[nof_bits + i];
    rem = (int)((reg ^ rx) & 0xFFFFu);
  }
  rem = __shfl_sync(0xFFFFFFFFu, rem, 0);
  __syncwarp();
  return rem;
}

}  // namespace srsue

// ulsch.cu -- uplink shared-channel encoder for sm_100a: transport block -> CRC24A -> code-block segmentation (+CRC24B) ->
// turbo encoder -> rate matching -> concatenation -> channel interleaver -> PUSCH scrambling (36.212 5.2.2 without control
// information, 36.211 5.3.1), for a batch of transport blocks of one grant shape.
//
// Replaces the bit chain of srslte_ue_ul_pusch_encode_rnti_softbuffer (/root/reference/ue/src/phy/phch_worker.cc:545-590,
// i.e. srslte_ulsch_encode + the scrambling of srslte_pusch_encode).  SURVEY 8 row f4: the turbo ENCODER of the uplink.
//
// The recursive systematic encoder is a linear recurrence over GF(2) with three bits of state, so it parallelises as a
// prefix scan: a thread owns one byte of the code block, computes the state transition of its eight steps as a map on the
// eight states (24 bits), the maps are composed across the CTA (warp shuffles, then one warp over the warp totals), and
// every thread re-runs its eight steps from its true start state to emit a parity byte.  The second constituent encoder
// does the same on the QPP-interleaved bits gathered from shared memory.  Rate matching is a gather through the circular-
// buffer read order of (K, F, rv), one thread per output bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srsue {

namespace {

// (a * b) mod g for 24-bit CRC polynomials
__device__ __forceinline__ uint32_t ul_gf_mul24(uint32_t x, uint32_t yv, uint32_t poly) {
  uint32_t r = 0;
#pragma unroll 4
  for (int i = 23; i >= 0; i--) {
    r <<= 1;
    if (r & 0x1000000u) r ^= poly;
    if ((yv >> i) & 1u) r ^= x;
  }
  return r & 0xFFFFFFu;
}

__device__ __forceinline__ void crc24_table(uint32_t* tab, uint32_t poly) {
  for (int v = threadIdx.x; v < 256; v += blockDim.x) {
    uint32_t c = (uint32_t)v << 16;
#pragma unroll
    for (int i = 0; i < 8; i++) { c <<= 1; if (c & 0x1000000u) c ^= poly; }
    tab[v] = c & 0xFFFFFFu;
  }
}

// CRC of nb bytes by one warp: lane l takes bytes [l * chunk, (l + 1) * chunk), the partial remainders are moved to their
// place with shift[l] = x^(8 * bytes after the chunk) mod g and added
template <typename Load>
__device__ __forceinline__ uint32_t warp_crc24(Load byte_at, int nb, int chunk, const uint32_t* __restrict__ shift, const uint32_t* tab, uint32_t poly,
                                               int lane) {
  const int b0 = lane * chunk, b1 = min(nb, b0 + chunk);
  uint32_t crc = 0;
  for (int i = b0; i < b1; i++) crc = ((crc << 8) & 0xFFFFFFu) ^ tab[((crc >> 16) ^ byte_at(i)) & 0xFFu];
  uint32_t part = (b1 > b0) ? ul_gf_mul24(crc, __ldg(shift + lane), poly) : 0u;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, off);
  return part;
}

// one step of the constituent encoder (36.212 5.1.3.2.1: g0 = 1 + D^2 + D^3 feedback, g1 = 1 + D + D^3); state = (s1 s2 s3)
__device__ __forceinline__ int rsc(int& s, int u) {
  const int s1 = (s >> 2) & 1, s2 = (s >> 1) & 1, s3 = s & 1;
  const int fb = u ^ s2 ^ s3;
  s = (fb << 2) | (s1 << 1) | s2;
  return fb ^ s1 ^ s3;
}
// map of a byte on the eight states: 3 bits per start state
__device__ __forceinline__ uint32_t byte_map(uint32_t byte) {
  uint32_t m = 0;
#pragma unroll
  for (int s0 = 0; s0 < 8; s0++) {
    int s = s0;
#pragma unroll
    for (int q = 7; q >= 0; q--) rsc(s, (byte >> q) & 1);
    m |= (uint32_t)s << (3 * s0);
  }
  return m;
}
// first a, then b
__device__ __forceinline__ uint32_t compose(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll
  for (int s = 0; s < 8; s++) r |= ((b >> (3 * ((a >> (3 * s)) & 7u))) & 7u) << (3 * s);
  return r;
}
constexpr uint32_t kIdentityMap = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12) | (5u << 15) | (6u << 18) | (7u << 21);

// Encodes the nb bytes of `in` (shared memory): parity bytes to `par`, the state after the last step is returned to every
// thread.  blockDim.x >= nb; s_warp: one word per warp.
__device__ __forceinline__ int rsc_encode_bytes(const uint8_t* in, int nb, uint8_t* par, uint32_t* s_warp) {
  const int i = threadIdx.x, lane = i & 31, wid = i >> 5, nw = (blockDim.x + 31) >> 5;
  const uint32_t byte = (i < nb) ? in[i] : 0u;
  uint32_t incl = (i < nb) ? byte_map(byte) : kIdentityMap;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl = compose(prev, incl);
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = (lane < nw) ? s_warp[lane] : kIdentityMap;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, t, off);
      if (lane >= off) t = compose(prev, t);
    }
    if (lane < nw) s_warp[lane] = t;          // inclusive over warps
  }
  __syncthreads();
  const uint32_t before_warp = wid ? s_warp[wid - 1] : kIdentityMap;
  uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
  if (lane == 0) excl = kIdentityMap;
  int s = (int)(compose(before_warp, excl) & 7u);          // start state of this byte (encoder starts in state 0)
  if (i < nb) {
    uint32_t z = 0;
#pragma unroll
    for (int q = 7; q >= 0; q--) z |= (uint32_t)rsc(s, (byte >> q) & 1) << q;
    par[i] = (uint8_t)z;
  }
  const int last = (int)(s_warp[nw - 1] & 7u);
  __syncthreads();
  return last;
}

__device__ __forceinline__ int bit_of(const uint8_t* bytes, int n) { return (bytes[n >> 3] >> (7 - (n & 7))) & 1; }

}  // namespace

// CRC24A of every transport block (one warp each); tbcrc[n][0..2] = the three CRC bytes
__global__ void __launch_bounds__(32) ulsch_tbcrc_kernel(const UlschArgs a) {
  __shared__ uint32_t s_tab[256];
  crc24_table(s_tab, kCrc24A);
  __syncwarp();
  const int n = blockIdx.x, nb = a.tbs / 8;
  const uint8_t* p = a.payload + (size_t)n * a.payload_stride;
  const uint32_t crc = warp_crc24([&](int i) { return (uint32_t)p[i]; }, nb, (nb + 31) / 32, a.crcshift, s_tab, kCrc24A, threadIdx.x);
  if (threadIdx.x < 3) a.tbcrc[(size_t)n * 4 + threadIdx.x] = (uint8_t)(crc >> (16 - 8 * threadIdx.x));
}

// one CTA per (code block, transport block), one thread per byte of the code block
__global__ void __launch_bounds__(768) ulsch_encode_kernel(const UlschArgs a) {
  __shared__ uint8_t s_cb[768], s_ci[768], s_p1[768], s_p2[768];
  __shared__ uint8_t s_tail[12];
  __shared__ uint32_t s_tab[256];
  __shared__ uint32_t s_warp[32];
  const int r = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const int32_t* cbt = a.cbtab + 8 * r;
  const int K = cbt[0], F = cbt[1], E = cbt[2], e_start = cbt[3], pos = cbt[4], nbs = cbt[5], perm_off = cbt[6], seq_off = cbt[7];
  const int nb = K / 8, nfill = F / 8, ncrc = (a.C > 1) ? 3 : 0;
  if (a.C > 1) crc24_table(s_tab, kCrc24B);
  // ---- the code block: filler, its share of (payload || CRC24A), room for CRC24B
  if (tid < nb) {
    uint32_t v = 0;
    if (tid >= nfill && tid < nfill + nbs) {
      const int sp = pos + tid - nfill;             // byte of the stream
      v = (sp < a.tbs / 8) ? a.payload[(size_t)n * a.payload_stride + sp] : a.tbcrc[(size_t)n * 4 + (sp - a.tbs / 8)];
    }
    s_cb[tid] = (uint8_t)v;
  }
  __syncthreads();
  if (ncrc && tid < 32) {
    const int nbc = nb - 3;
    const uint32_t crc = warp_crc24([&](int i) { return (uint32_t)s_cb[i]; }, nbc, (nbc + 31) / 32, a.crcshift + 32 * (1 + r), s_tab, kCrc24B, tid);
    if (tid < 3) s_cb[nbc + tid] = (uint8_t)(crc >> (16 - 8 * tid));
  }
  __syncthreads();
  // ---- second encoder's input: c'(i) = c(pi(i))
  if (tid < nb) {
    const uint16_t* pi = a.perm + perm_off + 8 * tid;
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) v = v * 2u + (uint32_t)bit_of(s_cb, __ldg(pi + q));
    s_ci[tid] = (uint8_t)v;
  }
  __syncthreads();
  int st1 = rsc_encode_bytes(s_cb, nb, s_p1, s_warp);
  int st2 = rsc_encode_bytes(s_ci, nb, s_p2, s_warp);
  // ---- trellis termination (36.212 5.1.3.2.2): three steps with the input taken from the feedback
  if (tid == 0) {
    int x[3], z[3], xp[3], zp[3];
#pragma unroll
    for (int t = 0; t < 3; t++) { x[t] = ((st1 >> 1) ^ st1) & 1; z[t] = rsc(st1, x[t]); }
#pragma unroll
    for (int t = 0; t < 3; t++) { xp[t] = ((st2 >> 1) ^ st2) & 1; zp[t] = rsc(st2, xp[t]); }
    // triples d(0), d(1), d(2) at k = K .. K+3
    s_tail[0] = x[0];  s_tail[1] = z[0];  s_tail[2] = x[1];
    s_tail[3] = z[1];  s_tail[4] = x[2];  s_tail[5] = z[2];
    s_tail[6] = xp[0]; s_tail[7] = zp[0]; s_tail[8] = xp[1];
    s_tail[9] = zp[1]; s_tail[10] = xp[2]; s_tail[11] = zp[2];
  }
  __syncthreads();
  // ---- rate matching: bit j of this block's E is element seq[j mod N] of the circular buffer without <NULL>s
  const uint16_t* seq = a.seq + seq_off;
  const int N = a.seq_len[r];
  uint8_t* e = a.ebits + (size_t)n * a.G + e_start;
  for (int j = tid; j < E; j += blockDim.x) {
    const uint32_t idx = __ldg(seq + (j < N ? j : j % N));
    const int di = (int)(idx >> 2), stream = (int)(idx & 3u);
    int bit;
    if (di >= K) bit = s_tail[3 * (di - K) + stream];
    else bit = bit_of(stream == 0 ? s_cb : stream == 1 ? s_p1 : s_p2, di);
    e[j] = (uint8_t)bit;
  }
}

// channel interleaver (36.212 5.2.2.8: symbol row * n_symb + col -> col * rows + row, qm bits each) + scrambling + packing;
// one thread per output byte
__global__ void __launch_bounds__(256) ulsch_interleave_kernel(const UlschArgs a) {
  const int n = blockIdx.y, ob = blockIdx.x * blockDim.x + threadIdx.x;
  if (ob >= a.G / 8) return;
  const uint8_t* e = a.ebits + (size_t)n * a.G;
  uint32_t v = 0;
#pragma unroll
  for (int q8 = 0; q8 < 8; q8++) {
    const int h = 8 * ob + q8, sym = h / a.qm, q = h - sym * a.qm;
    const int col = sym / a.rows, row = sym - col * a.rows;
    v = v * 2u + e[(row * a.n_symb + col) * a.qm + q];
  }
  a.out[(size_t)n * a.out_stride + ob] = (uint8_t)(v ^ __ldg(a.scramble + ob));
}

}  // namespace srsue

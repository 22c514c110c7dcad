// acs_peak.cu -- the arithmetic of the turbo decoder's forward sweep alone: what the SM sustains for this instruction
// stream when nothing is loaded or stored (registers only), at 1..4 warps per scheduler.  The decoder's measured step
// time over this number says how much of its time is not arithmetic.  Prints JSON.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/acs_peak tools/acs_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t vadd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ void beta_step(const uint32_t (&bn)[8], uint32_t (&b)[8], uint32_t x, uint32_t y, uint32_t xy) {
  b[0] = vaddmax(bn[4], xy, bn[0]); b[1] = vaddmax(bn[0], xy, bn[4]);
  b[2] = vaddmax(bn[5], y, vadd(bn[1], x)); b[3] = vaddmax(bn[1], y, vadd(bn[5], x));
  b[4] = vaddmax(bn[2], y, vadd(bn[6], x)); b[5] = vaddmax(bn[6], y, vadd(bn[2], x));
  b[6] = vaddmax(bn[3], xy, bn[7]); b[7] = vaddmax(bn[7], xy, bn[3]);
}
__device__ __forceinline__ void alpha_step(uint32_t (&a)[8], uint32_t x, uint32_t y, uint32_t xy) {
  uint32_t n[8];
  n[0] = vaddmax(a[1], xy, a[0]); n[1] = vaddmax(a[2], x, vadd(a[3], y)); n[2] = vaddmax(a[4], y, vadd(a[5], x)); n[3] = vaddmax(a[6], xy, a[7]);
  n[4] = vaddmax(a[0], xy, a[1]); n[5] = vaddmax(a[2], y, vadd(a[3], x)); n[6] = vaddmax(a[4], x, vadd(a[5], y)); n[7] = vaddmax(a[7], xy, a[6]);
#pragma unroll
  for (int s = 0; s < 8; s++) a[s] = n[s];
}
__device__ __forceinline__ void normalise(uint32_t (&m)[8], uint32_t m1) {
  const uint32_t n0 = ~m[0];
#pragma unroll
  for (int s = 1; s < 8; s++) m[s] = vadd(m[s], n0);
  m[0] = m1;
}
__device__ __forceinline__ void ext_parts(const uint32_t (&a)[8], const uint32_t (&bn)[8], uint32_t y, uint32_t& l1, uint32_t& l0) {
  uint32_t a00 = vadd(a[0], bn[0]); a00 = vaddmax(a[1], bn[4], a00); a00 = vaddmax(a[6], bn[7], a00); a00 = vaddmax(a[7], bn[3], a00);
  uint32_t a11 = vadd(a[0], bn[4]); a11 = vaddmax(a[1], bn[0], a11); a11 = vaddmax(a[6], bn[3], a11); a11 = vaddmax(a[7], bn[7], a11);
  uint32_t a01 = vadd(a[2], bn[5]); a01 = vaddmax(a[3], bn[1], a01); a01 = vaddmax(a[4], bn[2], a01); a01 = vaddmax(a[5], bn[6], a01);
  uint32_t a10 = vadd(a[2], bn[1]); a10 = vaddmax(a[3], bn[5], a10); a10 = vaddmax(a[4], bn[6], a10); a10 = vaddmax(a[5], bn[2], a10);
  l1 = vaddmax(a11, y, a10); l0 = vaddmax(a01, y, a00);
}

// MODE 0: forward group (7 beta steps + 8 x (extrinsic, clamp, alpha)); MODE 1: backward group (8 beta steps)
template <int MODE>
__global__ void __launch_bounds__(512, 1) acs_kernel(uint32_t* out, int groups, uint32_t seed) {
  uint32_t a[8], b[8], x[8], y[8], m1;
  asm volatile("mov.b32 %0, 0xFFFFFFFF;" : "=r"(m1));
#pragma unroll
  for (int s = 0; s < 8; s++) { a[s] = seed * (s + 1) + threadIdx.x; b[s] = seed * (s + 9); x[s] = (seed >> s) & 0x00FF00FF; y[s] = (seed >> (s + 3)) & 0x003F003F; }
  uint32_t acc = 0;
#pragma unroll 1
  for (int g = 0; g < groups; g++) {
    if (MODE == 0) {
      uint32_t B[8][8];
#pragma unroll
      for (int s = 0; s < 8; s++) B[7][s] = b[s];
#pragma unroll
      for (int i = 7; i >= 1; i--) { beta_step(B[i], B[i - 1], x[i], y[i], vadd(x[i], y[i])); if ((i & 3) == 0) normalise(B[i - 1], m1); }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint32_t l1, l0;
        ext_parts(a, B[i], y[i], l1, l0);
        const uint32_t q = vadd(l1, ~l0);
        const uint32_t r = __viaddmin_s16x2_relu(q, 0x08000800u, 0x0FFE0FFEu);
        acc ^= r;
        x[i] = vadd(x[i], r & 0x00010001u);          // keeps the next group dependent on this one (no hoisting)
        alpha_step(a, x[i], y[i], vadd(x[i], y[i]));
        if ((i & 3) == 3) normalise(a, m1);
      }
#pragma unroll
      for (int s = 0; s < 8; s++) b[s] = B[0][s];
    } else {
#pragma unroll
      for (int i = 7; i >= 0; i--) {
        uint32_t nb[8];
        beta_step(b, nb, x[i], y[i], vadd(x[i], y[i]));
#pragma unroll
        for (int s = 0; s < 8; s++) b[s] = nb[s];
        if ((i & 3) == 0) normalise(b, m1);
      }
      x[0] ^= b[3] & 1u;
    }
  }
#pragma unroll
  for (int s = 0; s < 8; s++) acc ^= a[s] ^ b[s];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  uint32_t* out;
  cudaMalloc(&out, 148 * 512 * 4 * 4);
  const int groups = 4000;
  printf("{\"sms\": %d, \"sm_clock_mhz\": %d, \"groups_of_8_steps\": %d, \"runs\": [", sms, khz / 1000, groups);
  bool first = true;
  for (int mode = 0; mode < 2; mode++)
    for (int warps = 4; warps <= 16; warps += 4) {          // warps per SM: 1..4 per scheduler
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) acs_kernel<0><<<sms, warps * 32>>>(out, groups, 12345u);
        else acs_kernel<1><<<sms, warps * 32>>>(out, groups, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      // clocks per (warp, trellis step) and scheduler: elapsed clocks / (steps per warp * warps per scheduler)
      const double clk = ms * 1e-3 * khz * 1e3;
      const double per = clk / ((double)groups * 8 * (warps / 4));
      printf("%s{\"sweep\": \"%s\", \"warps_per_scheduler\": %d, \"ms\": %.3f, \"clocks_per_warp_step_per_scheduler\": %.1f}", first ? "" : ", ",
             mode == 0 ? "forward (7 beta + 8 alpha/extrinsic per 8 steps)" : "backward (8 beta per 8 steps)", warps / 4, ms, per);
      first = false;
    }
  printf("]}\n");
  cudaFree(out);
  return 0;
}

// f32x2_peak.cu -- issue rate of the packed FP32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2)
// against their scalar forms, to decide whether the no-FMA butterflies of ofdm.cu gain from packing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_peak f32x2_peak.cu && ./f32x2_peak
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096, CHAINS = 8;
typedef unsigned long long u64;

__device__ __forceinline__ float2 add2(float2 a, float2 b) { float2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&r) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return r; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { float2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&r) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return r; }

template <int OP>
__global__ void __launch_bounds__(256) k(float2* out, float seed) {
  float2 a[CHAINS];
  const float2 b = make_float2(seed, seed * 0.5f), c = make_float2(1.0f + seed * 1e-7f, 1.0f - seed * 1e-7f);
#pragma unroll
  for (int i = 0; i < CHAINS; i++) a[i] = make_float2(threadIdx.x + i, i);
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < CHAINS; i++) {
        if (OP == 0) { a[i].x = __fadd_rn(a[i].x, b.x); a[i].y = __fadd_rn(a[i].y, b.y); }        // 2 FADD
        if (OP == 1) a[i] = add2(a[i], b);                                                        // 1 FADD2
        if (OP == 2) { a[i].x = __fmul_rn(a[i].x, c.x); a[i].y = __fmul_rn(a[i].y, c.y); }        // 2 FMUL
        if (OP == 3) a[i] = mul2(a[i], c);                                                        // 1 FMUL2
        if (OP == 4) { a[i] = mul2(a[i], c); a[i] = add2(a[i], b); }                              // FMUL2 + FADD2
        if (OP == 5) { a[i] = add2(a[i], b); a[i].x = __fadd_rn(a[i].x, b.y); }                   // FADD2 + FADD
      }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(float2* d, int sms, int instr_per_inner) {
  const int grid = sms * 8;
  k<OP><<<grid, 256>>>(d, 1.0f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<OP><<<grid, 256>>>(d, (float)r);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = 5.0 * grid * 256 * (double)ITERS * 4 * CHAINS * instr_per_inner;
  return instr / (ms * 1e-3);       // thread-instructions per second
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float2* d;
  cudaMalloc(&d, sizeof(float2) * p.multiProcessorCount * 8 * 256);
  const double per = (double)p.multiProcessorCount * clk * 1e3;     // SM-clocks per second
  printf("{\"sms\": %d, \"clock_khz\": %d, \"lanes_per_clk_per_sm\": {", p.multiProcessorCount, clk);
  printf("\"FADD\": %.1f, ", run<0>(d, p.multiProcessorCount, 2) / per);
  printf("\"FADD2\": %.1f, ", run<1>(d, p.multiProcessorCount, 1) / per);
  printf("\"FMUL\": %.1f, ", run<2>(d, p.multiProcessorCount, 2) / per);
  printf("\"FMUL2\": %.1f, ", run<3>(d, p.multiProcessorCount, 1) / per);
  printf("\"FMUL2+FADD2\": %.1f, ", run<4>(d, p.multiProcessorCount, 2) / per);
  printf("\"FADD2+FADD\": %.1f}}\n", run<5>(d, p.multiProcessorCount, 2) / per);
  return 0;
}

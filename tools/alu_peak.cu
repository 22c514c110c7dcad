// alu_peak.cu -- issue-rate microbenchmark for the packed-int16 instructions the turbo decoder is built
// from (SURVEY.md 8d asks for this number next to hbm_gbs).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_peak alu_peak.cu && ./alu_peak
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int ITERS = 4096;
constexpr int CHAINS = 16;

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t a[CHAINS];
  const uint32_t b = seed ^ 0x00010001u, c = seed ^ 0x00070003u;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) a[i] = threadIdx.x * 0x00010001u + i;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) {
        if (OP == 0) a[i] = __vadd2(a[i], b);                         // VIADD.16x2
        if (OP == 1) a[i] = __viaddmax_s16x2(a[i], b, c);             // VIADDMNMX.S16x2
        if (OP == 2) a[i] = __vmaxs2(a[i], b) ^ c;                        // VIMNMX.S16x2
        if (OP == 3) a[i] = (a[i] + b + it) ^ c;                            // IADD3
        if (OP == 4) a[i] = __vsub2(a[i], b);                         // packed subtract
        if (OP == 5) { a[i] = __viaddmax_s16x2(a[i], b, c); a[i] = __vadd2(a[i], c); }   // 1:1 mix
        if (OP == 6) a[i] = __vmins2(__vmaxs2(a[i], b), c);           // clamp
        if (OP == 7) a[i] = a[i] * 3u + b;                            // IMAD (fma pipe)
        if (OP == 8) { a[i] = __vadd2(a[i], b); a[i] = a[i] * 3u + c; }                  // VIADD.16x2 + IMAD
        if (OP == 9) { a[i] = __viaddmax_s16x2(a[i], b, c); a[i] = a[i] * 3u + c; }      // VIADDMNMX + IMAD
        if (OP == 10) { a[i] = __vmaxs2(a[i], b); a[i] = __vadd2(a[i], c); }             // VIMNMX + VIADD.16x2
        if (OP == 11) { a[i] = __vmaxs2(a[i], b); a[i] = __viaddmax_s16x2(a[i], c, b); } // VIMNMX + VIADDMNMX
        if (OP == 12) { a[i] = __vadd2(a[i], b); a[i] = a[i] ^ (c + it); }               // VIADD.16x2 + LOP3
        if (OP == 13) { a[i] = __viaddmax_s16x2(a[i], b, c); a[i] = a[i] ^ (c + it); }   // VIADDMNMX + LOP3
        if (OP == 14) a[i] = __vimax3_s16x2(a[i], b, c + it);                            // VIMNMX3
        if (OP == 15) { a[i] = __vadd2(a[i], b); a[i] = __vsub2(a[i], c); }              // VIADD + VSUB
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(uint32_t* d, int sms, int ops_per_inner) {
  const int grid = sms * 8;
  k<OP><<<grid, 256>>>(d, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<OP><<<grid, 256>>>(d, r);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = 5.0 * grid * 256.0 * ITERS * 4.0 * CHAINS * ops_per_inner;
  return instr / (ms * 1e-3);      // thread-instructions per second
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  uint32_t* d;
  cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  const int sms = p.multiProcessorCount;
  double r[16];
  r[0] = run<0>(d, sms, 1); r[1] = run<1>(d, sms, 1); r[2] = run<2>(d, sms, 1); r[3] = run<3>(d, sms, 1);
  r[4] = run<4>(d, sms, 1); r[5] = run<5>(d, sms, 2); r[6] = run<6>(d, sms, 2); r[7] = run<7>(d, sms, 1);
  r[8] = run<8>(d, sms, 2); r[9] = run<9>(d, sms, 2); r[10] = run<10>(d, sms, 2); r[11] = run<11>(d, sms, 2);
  r[12] = run<12>(d, sms, 2); r[13] = run<13>(d, sms, 2); r[14] = run<14>(d, sms, 1); r[15] = run<15>(d, sms, 2);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const char* names[16] = {"viadd16x2", "viaddmnmx_s16x2", "vimnmx_s16x2", "iadd3", "vsub2", "mix_addmax_add", "clamp_minmax", "imad",
                           "viadd_imad", "viaddmnmx_imad", "vimnmx_viadd", "vimnmx_viaddmnmx", "viadd_lop3", "viaddmnmx_lop3", "vimnmx3", "viadd_vsub"};
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f", p.name, sms, clk / 1000.0);
  for (int i = 0; i < 16; i++) {
    // lanes per clock per SM at the nominal max clock (the achieved clock may be lower)
    printf(", \"%s_tinstr_per_s\": %.4g, \"%s_lanes_per_clk_sm\": %.2f", names[i], r[i], names[i], r[i] / sms / (clk * 1e3));
  }
  printf("}\n");
  return 0;
}

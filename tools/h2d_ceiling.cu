// h2d_ceiling.cu -- what the box gives: plain cudaMemcpyAsync from pinned host memory to N GPUs at once (one call per
// copy, the only kind of copy libsrsue_gpu issues), for N = 1, 2, 4, 8 of the visible devices.  The end-to-end figure of
// bench.py (`e2e`: 245 760 bytes of IQ per subframe over PCIe) is quoted as a fraction of this ceiling.
//
//   nvcc -O3 -o build/h2d_ceiling tools/h2d_ceiling.cu && build/h2d_ceiling > profiles/h2d_ceiling_r02.json
//
// Variants per N: default pinned memory, write-combined pinned memory, two copy streams per GPU (half the buffer each),
// and copies of the size bench.py's chunked pipeline uses (64 subframes = 15.7 MB) instead of one 256 MB copy.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  const size_t bytes = 256u << 20, chunk = 64u * 245760u;
  const int reps = 8;
  printf("{\"what\": \"aggregate cudaMemcpyAsync host->device rate from pinned memory, all N GPUs concurrently, GB/s\", \"bytes_per_gpu_per_rep\": %zu, \"reps\": %d, \"devices_visible\": %d, \"runs\": [", bytes, reps, nd);
  bool first = true;
  for (int n = 1; n <= nd; n *= 2) {
    for (int variant = 0; variant < 4; variant++) {
      std::vector<void*> h(n), d(n);
      std::vector<cudaStream_t> s(2 * n);
      for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(i));
        CK(cudaHostAlloc(&h[i], bytes, variant == 1 ? cudaHostAllocWriteCombined | cudaHostAllocPortable : cudaHostAllocPortable));
        memset(h[i], 1, bytes);
        CK(cudaMalloc(&d[i], bytes));
        CK(cudaStreamCreateWithFlags(&s[2 * i], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s[2 * i + 1], cudaStreamNonBlocking));
      }
      auto pass = [&]() -> int {
        for (int i = 0; i < n; i++) {
          CK(cudaSetDevice(i));
          if (variant == 2) {
            CK(cudaMemcpyAsync(d[i], h[i], bytes / 2, cudaMemcpyHostToDevice, s[2 * i]));
            CK(cudaMemcpyAsync((char*)d[i] + bytes / 2, (char*)h[i] + bytes / 2, bytes / 2, cudaMemcpyHostToDevice, s[2 * i + 1]));
          } else if (variant == 3) {
            for (size_t off = 0; off < bytes; off += chunk)
              CK(cudaMemcpyAsync((char*)d[i] + off, (char*)h[i] + off, off + chunk <= bytes ? chunk : bytes - off, cudaMemcpyHostToDevice, s[2 * i]));
          } else {
            CK(cudaMemcpyAsync(d[i], h[i], bytes, cudaMemcpyHostToDevice, s[2 * i]));
          }
        }
        for (int i = 0; i < n; i++) { CK(cudaSetDevice(i)); CK(cudaStreamSynchronize(s[2 * i])); CK(cudaStreamSynchronize(s[2 * i + 1])); }
        return 0;
      };
      if (pass()) return 1;
      const double t0 = now();
      for (int r = 0; r < reps; r++) if (pass()) return 1;
      const double dt = now() - t0;
      const char* names[4] = {"pinned", "pinned_write_combined", "pinned_two_streams_per_gpu", "pinned_15.7MB_copies"};
      printf("%s{\"n_gpus\": %d, \"variant\": \"%s\", \"gb_per_s\": %.2f, \"gb_per_s_per_gpu\": %.2f}", first ? "" : ", ", n, names[variant],
             (double)bytes * n * reps / dt / 1e9, (double)bytes * reps / dt / 1e9);
      first = false;
      for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(i));
        cudaStreamDestroy(s[2 * i]); cudaStreamDestroy(s[2 * i + 1]);
        cudaFree(d[i]); cudaFreeHost(h[i]);
      }
    }
  }
  printf("]}\n");
  return 0;
}

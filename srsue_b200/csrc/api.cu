// api.cu -- host side of libsrsue_gpu: context, per-K turbo tables, PDSCH plans and the C ABI declared in
// include/srsue_gpu/srsue_gpu.h.  This file and srslte_shim.cc are the only host code that touches CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"
#include "srsue_gpu/srsue_gpu.h"

using namespace srsue;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU_CHECK(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return fail(SRSUE_GPU_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
cudaError_t upload(T** d, const std::vector<T>& h) {
  cudaError_t e = cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

struct TurboTables {
  TurboGeom g{};
  uint16_t* d_perm = nullptr;
  uint16_t* d_deint = nullptr;                // [K] turbo_deint_table
  uint32_t* d_tpos[2] = {nullptr, nullptr};   // [0] CRC24A, [1] CRC24B
};

struct Scratch {
  int16_t* nii = nullptr; size_t nii_elems = 0;
  uint8_t* bits = nullptr; size_t bits_bytes = 0;
  int16_t* tcb = nullptr; size_t tcb_elems = 0;
  uint4* ckpt = nullptr; size_t ckpt_bytes = 0;
  int* counter = nullptr;
  void release() {
    cudaFree(nii); cudaFree(bits); cudaFree(tcb); cudaFree(ckpt); cudaFree(counter); counter = nullptr;
    nii = nullptr; bits = nullptr; tcb = nullptr; ckpt = nullptr; nii_elems = bits_bytes = tcb_elems = ckpt_bytes = 0;
  }
};

}  // namespace

namespace srsue {
// error reporting for the other translation units of the library (batch.cu)
int internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }

// Pinned host regions the library knows about (srsue_gpu_host_alloc / srsue_gpu_host_register): the batching layer
// lets the GPU fetch subframes that lie inside them directly (zero-copy) instead of issuing one copy per subframe.
namespace {
std::mutex g_regions_mu;
std::map<uintptr_t, std::pair<size_t, bool>> g_regions;    // base -> (bytes, registered by us with cudaHostRegister)
}
bool host_region_contains(const void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_regions_mu);
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  auto it = g_regions.upper_bound(a);
  if (it == g_regions.begin()) return false;
  --it;
  return a >= it->first && a + bytes <= it->first + it->second.first;
}
}  // namespace srsue

struct srsue_gpu_ctx {
  int device = 0, num_sms = 0, smem_optin = 0, smem_sm = 0;
  std::mutex mu;
  std::map<int, TurboTables> turbo;
  // decoder scratch of the device-pointer srsue_gpu_tdec_* calls, one set per stream: launches on one stream are ordered,
  // callers on different streams (or threads) never share buffers.  std::map nodes do not move, so the references stay valid.
  std::map<cudaStream_t, Scratch> scratch;
  int last_grid = 0, last_block = 0, last_smem = 0, last_ncb = 0;
  std::atomic<int> launch_count{0};       // workers on different streams share the context
  std::atomic<bool> attr_set{false};
  std::atomic<bool> pdcch_attr_set{false};
  // cell-search tables and scratch (built on first use)
  float2* d_pss_freq = nullptr; int8_t* d_sss = nullptr;
  std::map<int, std::pair<float2*, float2*>> sync_tabs;      // nfft -> (PSS time replicas [3][nfft], nfft/2 twiddles)
  unsigned long long* d_peak_key = nullptr; double* d_power_sum = nullptr; int sync_cap = 0;
  float2* d_cexp = nullptr;                                  // CFO correction: 4096-entry unit circle (built on first use)
};

namespace {

Scratch& stream_scratch(srsue_gpu_ctx* ctx, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  return ctx->scratch[st];
}

int get_turbo_tables(srsue_gpu_ctx* ctx, int K, const TurboTables** out) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->turbo.find(K);
  if (it != ctx->turbo.end()) { *out = &it->second; return 0; }
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  TurboTables t;
  t.g = turbo_geom(K);
  std::vector<uint16_t> pos;
  turbo_perm_table(t.g, pos);
  CU_CHECK(upload(&t.d_perm, pos));
  turbo_deint_table(t.g, pos);
  CU_CHECK(upload(&t.d_deint, pos));
  const uint32_t polys[2] = {kCrc24A, kCrc24B};
  for (int i = 0; i < 2; i++) {
    std::vector<uint32_t> tpos;
    turbo_crc_table(t.g, polys[i], tpos);
    CU_CHECK(upload(&t.d_tpos[i], tpos));
  }
  auto ins = ctx->turbo.emplace(K, t);
  *out = &ins.first->second;
  return 0;
}

struct TurboLaunchCfg { int ncb, threads, grid, smem, ngroups, group_threads, group_slots, bwd_two; };

// words between the exchange arrays of consecutive slots: plane/2 plus the skew that lets a warp straddling
// two slots keep hitting distinct shared-memory banks ((plane/2 + skew) mod 32 == T mod 32)
int turbo_slot_words(const TurboGeom& g) {
  const int base = g.plane / 2;
  return base + ((g.T - base) % 32 + 32) % 32;
}

// stride of the position-table rows in shared memory (template parameter of the kernels)
int turbo_perm_stride(const TurboGeom& g) { return g.T <= 32 ? 32 : 64; }

int turbo_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// CTAs per SM of the persistent decoder (SRSUE_TURBO_CTAS_PER_SM overrides; 0: choose per code-block size)
int turbo_ctas_per_sm() {
  static const int v = std::max(0, std::min(4, turbo_env_int("SRSUE_TURBO_CTAS_PER_SM", 0)));
  return v;
}

// Phase groups of a CTA with ncb slots (turbo.cu): two groups of whole warps when there are at least two slots
void turbo_groups(const TurboGeom& g, int ncb, int* ngroups, int* group_threads, int* group_slots) {
  static const int want = std::max(1, std::min(2, turbo_env_int("SRSUE_TURBO_GROUPS", 2)));
  const int ng = (ncb >= 2) ? want : 1;
  *ngroups = ng;
  *group_slots = (ncb + ng - 1) / ng;
  *group_threads = ((*group_slots * g.T + 31) / 32) * 32;
}
int turbo_threads(const TurboGeom& g, int ncb) {
  int ng, gt, gs;
  turbo_groups(g, ncb, &ng, &gt, &gs);
  return ng * gt;
}

// Shared memory of one decoder CTA with ncb slots: position table, flags, per slot the exchange array, per thread six
// (eight with two groups in flight in the backward sweep) 16-byte staging chunks and a scratch word
int turbo_smem_bytes(const TurboGeom& g, int ncb, int chunks = 6) {
  const int threads = turbo_threads(g, ncb);
  const int slot_bytes = turbo_slot_words(g) * 4;
  return g.W * turbo_perm_stride(g) * 4 + 2 * ((ncb + 3) & ~3) * 4 + 16 + ncb * slot_bytes + 16 + threads * (16 * chunks + 4);
}

TurboLaunchCfg turbo_launch_cfg(const srsue_gpu_ctx* ctx, const TurboGeom& g, int n_cb, bool crc) {
  (void)crc;
  // CTAs per SM: whichever of 1 and 2 keeps more code-block slots resident (1 KB per CTA is reserved by the driver);
  // SRSUE_TURBO_CTAS_PER_SM overrides
  auto slots_for = [&](int per_sm) {
    const int budget = std::min((ctx->smem_sm - per_sm * 1024) / per_sm, ctx->smem_optin);
    const int max_threads = std::min(kTurboMaxThreads / per_sm, 1024);
    int ncb = max_threads / g.T;
    while (ncb > 0 && (turbo_smem_bytes(g, ncb) > budget || turbo_threads(g, ncb) > max_threads)) ncb--;
    return ncb;
  };
  int per_sm = turbo_ctas_per_sm();
  if (per_sm == 0) per_sm = (2 * slots_for(2) >= slots_for(1)) ? 2 : 1;
  int ncb = slots_for(per_sm);
  if (const char* e = getenv("SRSUE_TURBO_MAX_SLOTS")) ncb = std::min(ncb, std::max(1, atoi(e)));   // tuning knob
  // spread small batches over all SMs rather than filling a few CTAs
  ncb = std::min(ncb, std::max(1, (n_cb + ctx->num_sms * per_sm - 1) / (ctx->num_sms * per_sm)));
  ncb = std::max(ncb, 1);
  TurboLaunchCfg c;
  c.ncb = ncb;
  turbo_groups(g, ncb, &c.ngroups, &c.group_threads, &c.group_slots);
  c.threads = c.ngroups * c.group_threads;
  c.grid = std::min((n_cb + ncb - 1) / ncb, ctx->num_sms * per_sm);
  // the K = 5824 kernels keep two groups of channel LLRs in flight in the backward sweep (eight staging chunks per thread,
  // which fit next to its 14 slots; K = 6144 and the small sizes would lose a slot to them, which costs more than it gains)
  static const bool generic_only = turbo_env_int("SRSUE_TURBO_GENERIC", 0) != 0;      // tuning: skip the T-specific kernels
  const int budget = std::min((ctx->smem_sm - per_sm * 1024) / per_sm, ctx->smem_optin);
  c.bwd_two = (!generic_only && g.T == 26 && turbo_perm_stride(g) == 32 && turbo_smem_bytes(g, ncb, 8) <= budget) ? 1 : 0;
  c.smem = turbo_smem_bytes(g, ncb, c.bwd_two ? 8 : 6);
  return c;
}

int ensure_scratch(Scratch& s, size_t nii_elems, size_t bits_bytes, size_t tcb_elems) {
  if (nii_elems > s.nii_elems) { cudaFree(s.nii); CU_CHECK(cudaMalloc((void**)&s.nii, nii_elems * 2)); s.nii_elems = nii_elems; }
  if (bits_bytes > s.bits_bytes) { cudaFree(s.bits); CU_CHECK(cudaMalloc((void**)&s.bits, bits_bytes)); s.bits_bytes = bits_bytes; }
  if (tcb_elems > s.tcb_elems) { cudaFree(s.tcb); CU_CHECK(cudaMalloc((void**)&s.tcb, tcb_elems * 2)); s.tcb_elems = tcb_elems; }
  return 0;
}

// launches the decoder for n_cb code blocks of size K.  cb_list (device) may be null.
int launch_turbo(srsue_gpu_ctx* ctx, Scratch& scr, const int16_t* d_in, long long in_stride, const int32_t* d_cb_list, int n_cb, int K,
                 int max_iter, int crc_type, uint8_t* d_bits, int out_stride, int32_t* d_status, cudaStream_t st, int min_iter = 1) {
  if (n_cb <= 0) return 0;
  if (max_iter < 1 || crc_type < 0 || crc_type > 2) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "bad max_iter/crc_type");
  const TurboTables* tt = nullptr;
  int rc = get_turbo_tables(ctx, K, &tt);
  if (rc) return rc;
  const TurboGeom& g = tt->g;
  const TurboLaunchCfg lc = turbo_launch_cfg(ctx, g, n_cb, crc_type != 0);
  const size_t slots = (size_t)lc.grid * lc.ncb;
  const int row_bytes = g.plane / 8;                                  // decisions of one block in DEC2 order
  const int dbits_stride = (row_bytes + 15) & ~15;
  rc = ensure_scratch(scr, slots * (size_t)(2 * 2 * 2 * 8 * (g.Ppad + 2)), (size_t)n_cb * dbits_stride, 0);
  if (rc) return rc;
  const size_t ckpt_bytes = slots * (size_t)(g.W / 8) * g.T * 32;
  if (ckpt_bytes > scr.ckpt_bytes) { cudaFree(scr.ckpt); CU_CHECK(cudaMalloc((void**)&scr.ckpt, ckpt_bytes)); scr.ckpt_bytes = ckpt_bytes; }
  typedef void (*TurboKernel)(const TurboArgs);
  static const TurboKernel kernels[4][2] = {{turbo_decode_kernel, turbo_decode_crc_kernel}, {turbo_decode_wide_kernel, turbo_decode_crc_wide_kernel},
                                            {turbo_decode_t26_kernel, turbo_decode_crc_t26_kernel}, {turbo_decode_t24_kernel, turbo_decode_crc_t24_kernel}};
  if (!ctx->attr_set) {
    for (int v = 0; v < 4; v++)
      for (int c = 0; c < 2; c++) CU_CHECK(cudaFuncSetAttribute(kernels[v][c], cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin));
    ctx->attr_set = true;
  }
  TurboArgs a{};
  a.in = d_in; a.in_stride = in_stride; a.cb_list = d_cb_list; a.n_cb = n_cb;
  a.dbits = scr.bits; a.dbits_stride = dbits_stride; a.out_status = d_status;
  a.max_iter = max_iter; a.crc_type = crc_type;
  a.min_iter = std::max(1, std::min(min_iter, max_iter));
  a.K = g.K; a.W = g.W; a.P = g.P; a.Ppad = g.Ppad; a.T = g.T; a.plane = g.plane;
  a.perm_tab = tt->d_perm;
  a.crc_lin = tt->d_tpos[crc_type == 2 ? 1 : 0];
  a.ncb_cta = lc.ncb;
  a.slot_words = turbo_slot_words(g);
  a.nii = reinterpret_cast<uint4*>(scr.nii);
  a.ckpt = scr.ckpt;
  if (!scr.counter) CU_CHECK(cudaMalloc((void**)&scr.counter, 256));
  // work counter of the persistent slots: the first grid * ncb code blocks are assigned statically
  CU_CHECK(cudaMemsetAsync(scr.counter, 0, sizeof(int), st));
  a.work_counter = scr.counter;
  a.work_base = lc.grid * lc.ncb;
  a.ones = 0xFFFFFFFFu;
  a.ngroups = lc.ngroups; a.group_threads = lc.group_threads; a.group_slots = lc.group_slots;
  // the second phase group starts a good half of a MAP pass late (a pass costs roughly 600 SM clocks per trellis step
  // of a window with all slots busy); SRSUE_TURBO_PHASE_DELAY (clocks) overrides
  {
    static const int env_delay = turbo_env_int("SRSUE_TURBO_PHASE_DELAY", -1);
    a.phase_delay = lc.ngroups > 1 ? (env_delay >= 0 ? env_delay : 350 * g.W) : 0;
    if ((long long)n_cb < 2LL * lc.grid * lc.ncb) a.phase_delay = 0;      // short launches: latency matters more than overlap
  }
  static const bool generic_only = turbo_env_int("SRSUE_TURBO_GENERIC", 0) != 0;
  const int variant = turbo_perm_stride(g) == 64 ? 1 : generic_only ? 0 : (g.T == 26 && lc.bwd_two) ? 2 : g.T == 24 ? 3 : 0;
  kernels[variant][crc_type ? 1 : 0]<<<lc.grid, lc.threads, lc.smem, st>>>(a);
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return fail(SRSUE_GPU_ERROR, "turbo decoder launch failed: %s (K=%d grid=%d threads=%d smem=%d slots=%d)", cudaGetErrorString(e), K, lc.grid,
                  lc.threads, lc.smem, lc.ncb);
  }
  {
    // decisions: DEC2 order -> natural-order bytes (persistent CTAs, the table of this K in shared memory)
    const int dsmem = ((g.K + 7) & ~7) * 2 + row_bytes * 8 + ((row_bytes + 15) & ~15);
    static const int deint_ctas = std::max(1, turbo_env_int("SRSUE_DEINT_CTAS_PER_SM", 8));
    const int dgrid = std::min(n_cb, ctx->num_sms * deint_ctas);
    tdec_deinterleave_kernel<<<dgrid, 256, dsmem, st>>>(scr.bits, dbits_stride, tt->d_deint, d_cb_list, n_cb, d_bits, out_stride, g.K, row_bytes);
    CU_CHECK(cudaGetLastError());
    ctx->launch_count++;
  }
  ctx->last_grid = lc.grid; ctx->last_block = lc.threads; ctx->last_smem = lc.smem; ctx->last_ncb = lc.ncb;
  ctx->launch_count++;
  return 0;
}

TurboGeomDev to_dev(const TurboGeom& g) { return TurboGeomDev{g.K, g.W, g.P, g.Ppad, g.T, g.plane, g.cb_elems}; }

}  // namespace

struct srsue_gpu_pdsch_plan {
  srsue_gpu_ctx* ctx = nullptr;
  CellCfg cell{};
  PdschCfg cfg{};
  CbSegm seg{};
  srsue_gpu_plan_info_t info{};
  TurboGeom gp{}, gm{};
  int crs_off[4][4]{};     // ports 2 / 3 (four-port cells): two pilot symbols, columns 0 and 1
  int max_E = 0, gather_stride = 0, direct = 0;
  // device tables
  int32_t* d_re = nullptr; uint32_t* d_scr = nullptr; uint16_t* d_gather = nullptr;
  int32_t* d_e_start = nullptr; int32_t* d_cb_geom = nullptr; int8_t* d_crs = nullptr; float* d_tw = nullptr;
  int32_t* d_list_m = nullptr; int32_t* d_list_p = nullptr; int32_t* d_tbmap = nullptr; uint32_t* d_tbshift = nullptr;
  // device work buffers (max_batch)
  float2* d_sf = nullptr; float2* d_ce = nullptr; float2* d_pil = nullptr; float* d_meas = nullptr; int16_t* d_sb = nullptr;
  uint8_t* d_cb_bits = nullptr; int32_t* d_cb_status = nullptr;
  // staging for the host-pointer call
  float2* d_iq = nullptr; uint8_t* d_payload = nullptr; int32_t* d_tb_status = nullptr;
  const int32_t* cfo_steps = nullptr; int32_t cfo_step_all = 0;            // carrier-offset correction of the batch calls
  int min_iter = 1;                                                          // srsue_gpu_pdsch_plan_set_min_iter
  const int32_t* row_filter = nullptr; int row_want = 0;                     // srsue_gpu_pdsch_plan_set_row_filter
  int iq_format = SRSUE_GPU_IQ_CF32; float iq16_scale = 1.0f / 32768.0f;   // what the d_iq / h_iq arguments of the batch calls point at
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev[8] = {};
  Scratch scratch;             // decoder scratch of this plan (plans may run concurrently on different streams)
  // PDCCH tables of (cell, cfi, sf_idx, ng): built on first use
  int pd_ng = -1, pd_nreg = 0;
  int32_t* d_pd_re4 = nullptr; int32_t* d_pd_src = nullptr; uint32_t* d_pd_scr = nullptr;
  std::map<int, int32_t*> d_pd_rm;   // D -> rate-matching order
  int32_t* d_pbch_re = nullptr; uint32_t* d_pbch_scr = nullptr; int32_t* d_pbch_rm = nullptr;   // PBCH tables, built on first use
  int pbch_nre = 240;
};

extern "C" {

const char* srsue_gpu_last_error(void) { return g_err.c_str(); }
int srsue_gpu_version(void) { return 100; }

int srsue_gpu_ctx_create(srsue_gpu_ctx_t** out, int device) {
  if (!out) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null ctx pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SRSUE_GPU_ERROR, "libsrsue_gpu needs a CUDA device and has no CPU fallback: %s",
                e != cudaSuccess ? cudaGetErrorString(e) : "no device");
  if (device < 0 || device >= n) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "device %d out of range (%d devices)", device, n);
  CU_CHECK(cudaSetDevice(device));
  cudaDeviceProp p;
  CU_CHECK(cudaGetDeviceProperties(&p, device));
  if (p.major < 10) return fail(SRSUE_GPU_ERROR, "device %s is sm_%d%d; libsrsue_gpu is built for sm_100a only", p.name, p.major, p.minor);
  auto* ctx = new srsue_gpu_ctx();
  ctx->device = device;
  ctx->num_sms = p.multiProcessorCount;
  ctx->smem_optin = (int)p.sharedMemPerBlockOptin;
  ctx->smem_sm = (int)p.sharedMemPerMultiprocessor;
  *out = ctx;
  return 0;
}

void srsue_gpu_ctx_destroy(srsue_gpu_ctx_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& kv : ctx->turbo) {
    cudaFree(kv.second.d_perm); cudaFree(kv.second.d_deint);
    for (int i = 0; i < 2; i++) cudaFree(kv.second.d_tpos[i]);
  }
  for (auto& kv : ctx->scratch) kv.second.release();
  cudaFree(ctx->d_pss_freq); cudaFree(ctx->d_sss);
  for (auto& kv : ctx->sync_tabs) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
  cudaFree(ctx->d_peak_key); cudaFree(ctx->d_power_sum); cudaFree(ctx->d_cexp);
  delete ctx;
}

int srsue_gpu_cell_search(srsue_gpu_ctx_t* ctx, const srsue_gpu_cf_t* d_iq, int n_bufs, int n_samples, long long stride, int nfft,
                          int force_n_id_2, int first_pos, srsue_gpu_sync_result_t* d_result, void* stream) {
  return srsue_gpu_cell_search_cp(ctx, d_iq, n_bufs, n_samples, stride, nfft, force_n_id_2, first_pos, 0, d_result, stream);
}

int srsue_gpu_cell_search_cp(srsue_gpu_ctx_t* ctx, const srsue_gpu_cf_t* d_iq, int n_bufs, int n_samples, long long stride, int nfft,
                             int force_n_id_2, int first_pos, int cp_mode, srsue_gpu_sync_result_t* d_result, void* stream) {
  if (cp_mode < 0 || cp_mode > 2) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "cell_search: cp_mode must be 0 (normal), 1 (extended) or 2 (detect)");
  const bool pow2 = nfft >= 128 && nfft <= 2048 && (nfft & (nfft - 1)) == 0;
  if (!ctx || !d_iq || !d_result || n_bufs < 1 || !pow2 || n_samples < 2 * nfft + (cp_mode ? nfft / 4 : 9 * nfft / 128) || stride < n_samples ||
      force_n_id_2 > 2 || first_pos < 0 || first_pos >= n_samples - (nfft - 1))
    return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "cell_search: bad arguments (nfft must be 128, 256, 512, 1024 or 2048)");
  static_assert(sizeof(srsue_gpu_sync_result_t) == sizeof(srsue_sync_result), "result layouts must match");
  CU_CHECK(cudaSetDevice(ctx->device));
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->d_sss) {
    std::vector<float> pf(3 * 62 * 2);
    std::vector<int8_t> ss((size_t)3 * 336 * 62);
    for (int u = 0; u < 3; u++) {
      pss_freq(u, pf.data() + (size_t)u * 124);
      for (int s5 = 0; s5 < 2; s5++)
        for (int n1 = 0; n1 < 168; n1++) sss_seq(n1, u, s5, ss.data() + ((size_t)u * 336 + s5 * 168 + n1) * 62);
    }
    CU_CHECK(upload(reinterpret_cast<float**>(&ctx->d_pss_freq), pf));
    CU_CHECK(upload(&ctx->d_sss, ss));
  }
  auto it = ctx->sync_tabs.find(nfft);
  if (it == ctx->sync_tabs.end()) {
    std::vector<float> pt((size_t)3 * nfft * 2), tw;
    for (int u = 0; u < 3; u++) pss_time_n(u, nfft, pt.data() + (size_t)u * nfft * 2);
    fft_twiddles(nfft, tw);
    std::pair<float2*, float2*> tabs{nullptr, nullptr};
    CU_CHECK(upload(reinterpret_cast<float**>(&tabs.first), pt));
    CU_CHECK(upload(reinterpret_cast<float**>(&tabs.second), tw));
    it = ctx->sync_tabs.emplace(nfft, tabs).first;
  }
  if (n_bufs > ctx->sync_cap) {
    cudaFree(ctx->d_peak_key); cudaFree(ctx->d_power_sum);
    CU_CHECK(cudaMalloc((void**)&ctx->d_peak_key, (size_t)n_bufs * sizeof(unsigned long long)));
    CU_CHECK(cudaMalloc((void**)&ctx->d_power_sum, (size_t)n_bufs * sizeof(double)));
    ctx->sync_cap = n_bufs;
  }
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(cudaMemsetAsync(ctx->d_peak_key, 0, (size_t)n_bufs * sizeof(unsigned long long), st));
  CU_CHECK(cudaMemsetAsync(ctx->d_power_sum, 0, (size_t)n_bufs * sizeof(double), st));
  SyncArgs a{};
  a.iq = reinterpret_cast<const float2*>(d_iq); a.stride = stride; a.n_samples = n_samples; a.n_bufs = n_bufs;
  a.nfft = nfft; a.log2n = 0; while ((1 << a.log2n) < nfft) a.log2n++;
  a.force_n_id_2 = force_n_id_2 < 0 ? -1 : force_n_id_2;
  a.first_pos = first_pos; a.cp_mode = cp_mode;
  a.pss_time = it->second.first; a.pss_freq = ctx->d_pss_freq; a.sss = ctx->d_sss; a.tw = it->second.second;
  a.peak_key = ctx->d_peak_key; a.power_sum = ctx->d_power_sum; a.result = reinterpret_cast<srsue_sync_result*>(d_result);
  const int n_pos = n_samples - (nfft - 1);
  const int smem_pss = (256 + 2 * nfft) * (int)sizeof(float2), smem_sss = 2 * nfft * (int)sizeof(float2);
  for (int done = 0; done < n_bufs; done += 65535) {
    const int n = std::min(65535, n_bufs - done);
    SyncArgs b = a;
    b.iq += (size_t)done * stride; b.peak_key += done; b.power_sum += done; b.result += done; b.n_bufs = n;
    pss_corr_kernel<<<dim3((n_pos - first_pos + 255) / 256, force_n_id_2 < 0 ? 3 : 1, n), 256, smem_pss, st>>>(b);
    sss_detect_kernel<<<n, 128, smem_sss, st>>>(b);
    ctx->launch_count += 2;
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_tdec_geometry(int K, int* W, int* P, int* cb_elems) {
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  const TurboGeom g = turbo_geom(K);
  if (W) *W = g.W;
  if (P) *P = g.P;
  if (cb_elems) *cb_elems = g.cb_elems;
  return 0;
}

int srsue_gpu_tdec_import(srsue_gpu_ctx_t* ctx, const int16_t* d_triples, int n_cb, int K, int16_t* d_tcb, void* stream) {
  if (!ctx || !d_triples || !d_tcb || n_cb < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tdec_import: bad arguments");
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  if (n_cb == 0) return 0;
  const TurboGeom g = turbo_geom(K);
  CU_CHECK(cudaSetDevice(ctx->device));
  for (int done = 0; done < n_cb; done += 65535) {
    const int n = std::min(65535, n_cb - done);
    dim3 grid((g.cb_elems + 255) / 256 > 8 ? 8 : (g.cb_elems + 255) / 256, n);
    triples_to_tcb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_triples + (size_t)done * (3 * K + 12), 3 * K + 12,
                                                                  d_tcb + (size_t)done * g.cb_elems, g.cb_elems, n, to_dev(g));
    ctx->launch_count++;
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_tdec_export(srsue_gpu_ctx_t* ctx, const int16_t* d_tcb, int n_cb, int K, int16_t* d_triples, void* stream) {
  if (!ctx || !d_triples || !d_tcb || n_cb < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tdec_export: bad arguments");
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  if (n_cb == 0) return 0;
  const TurboGeom g = turbo_geom(K);
  CU_CHECK(cudaSetDevice(ctx->device));
  for (int done = 0; done < n_cb; done += 65535) {
    const int n = std::min(65535, n_cb - done);
    dim3 grid(8, n);
    tcb_to_triples_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_tcb + (size_t)done * g.cb_elems, g.cb_elems,
                                                                  d_triples + (size_t)done * (3 * K + 12), 3 * K + 12, n, to_dev(g));
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_tdec_decode(srsue_gpu_ctx_t* ctx, const int16_t* d_tcb, int n_cb, int K, int max_iter, int crc_type,
                          uint8_t* d_bits, int32_t* d_status, void* stream) {
  if (!ctx || !d_tcb || !d_bits || !d_status || n_cb < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tdec_decode: bad arguments");
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  CU_CHECK(cudaSetDevice(ctx->device));
  const TurboGeom g = turbo_geom(K);
  return launch_turbo(ctx, stream_scratch(ctx, (cudaStream_t)stream), d_tcb, g.cb_elems, nullptr, n_cb, K, max_iter, crc_type, d_bits, K / 8, d_status, (cudaStream_t)stream);
}

int srsue_gpu_tdec_run_all(srsue_gpu_ctx_t* ctx, const int16_t* d_triples, int n_cb, int K, int max_iter, int crc_type,
                           uint8_t* d_bits, int32_t* d_status, void* stream) {
  if (!ctx) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null ctx");
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  if (n_cb <= 0) return n_cb == 0 ? 0 : fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "negative n_cb");
  const TurboGeom g = turbo_geom(K);
  CU_CHECK(cudaSetDevice(ctx->device));
  Scratch& scr = stream_scratch(ctx, (cudaStream_t)stream);
  int rc = ensure_scratch(scr, 0, 0, (size_t)n_cb * g.cb_elems);
  if (rc) return rc;
  ctx->launch_count = 0;
  rc = srsue_gpu_tdec_import(ctx, d_triples, n_cb, K, scr.tcb, stream);
  if (rc) return rc;
  return srsue_gpu_tdec_decode(ctx, scr.tcb, n_cb, K, max_iter, crc_type, d_bits, d_status, stream);
}

int srsue_gpu_tdec_run_all_host(srsue_gpu_ctx_t* ctx, const int16_t* h_triples, int n_cb, int K, int max_iter, int crc_type,
                                uint8_t* h_bits, int32_t* h_status) {
  if (!ctx || !h_triples || !h_bits || !h_status || n_cb < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tdec_run_all_host: bad arguments");
  if (qpp_index(K) < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "K=%d is not a valid LTE code-block size", K);
  if (n_cb == 0) return 0;
  static std::mutex host_call_mu;     // the context-level decoder scratch is shared: one host call at a time
  std::lock_guard<std::mutex> lk(host_call_mu);
  CU_CHECK(cudaSetDevice(ctx->device));
  int16_t* d_in = nullptr; uint8_t* d_bits = nullptr; int32_t* d_st = nullptr;
  const size_t in_bytes = (size_t)n_cb * (3 * K + 12) * 2, bits_bytes = (size_t)n_cb * (K / 8);
  CU_CHECK(cudaMalloc((void**)&d_in, in_bytes));
  CU_CHECK(cudaMalloc((void**)&d_bits, bits_bytes));
  CU_CHECK(cudaMalloc((void**)&d_st, (size_t)n_cb * 4));
  int rc = 0;
  cudaError_t e = cudaMemcpy(d_in, h_triples, in_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) rc = srsue_gpu_tdec_run_all(ctx, d_in, n_cb, K, max_iter, crc_type, d_bits, d_st, nullptr);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_bits, d_bits, bits_bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_status, d_st, (size_t)n_cb * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_bits); cudaFree(d_st);
  if (e != cudaSuccess) return fail(SRSUE_GPU_ERROR, "tdec_run_all_host: %s", cudaGetErrorString(e));
  return rc;
}

int srsue_gpu_tdec_last_launch(srsue_gpu_ctx_t* ctx, int* grid, int* block, int* smem_bytes, int* cb_per_cta) {
  if (!ctx) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  if (grid) *grid = ctx->last_grid;
  if (block) *block = ctx->last_block;
  if (smem_bytes) *smem_bytes = ctx->last_smem;
  if (cb_per_cta) *cb_per_cta = ctx->last_ncb;
  return 0;
}

int srsue_gpu_last_launch_count(srsue_gpu_ctx_t* ctx) { return ctx ? ctx->launch_count.load() : 0; }

// ---- PDSCH plan -------------------------------------------------------------------------------------
// +-1 signs of the CRS of the plan's subframe, [row][re, im][2 nof_prb]: rows 0..3 = the four CRS symbols of ports 0 / 1
// (one sequence per symbol, shared by the ports), rows 4, 5 = symbol 1 of each slot (ports 2 / 3 of a four-port cell);
// and the first pilot subcarrier of every (port, pilot symbol)
static void build_crs_tables(srsue_gpu_pdsch_plan_t* p, int sf_idx, std::vector<int8_t>& crs) {
  const int M = 2 * p->cell.nof_prb, cp = p->cell.cp, rows = p->cell.nof_ports == 4 ? 6 : 4;
  const int crs_l[6] = {0, cp ? 3 : 4, cp ? 6 : 7, cp ? 9 : 11, 1, slot_symb(cp) + 1};
  crs.assign((size_t)rows * 2 * M, 0);
  std::vector<int8_t> rs, is;
  for (int r = 0; r < rows; r++) {
    crs_signs(p->cell, sf_idx, crs_l[r], rs, is);
    std::copy(rs.begin(), rs.end(), crs.begin() + (r * 2 + 0) * M);
    std::copy(is.begin(), is.end(), crs.begin() + (r * 2 + 1) * M);
    if (r < 4) for (int port = 0; port < 2; port++) p->crs_off[port][r] = crs_offset(p->cell, port, crs_l[r]);
    else for (int port = 2; port < 4; port++) p->crs_off[port][r - 4] = crs_offset(p->cell, port, crs_l[r]);
  }
}

int srsue_gpu_pdsch_plan_create(srsue_gpu_ctx_t* ctx, const srsue_gpu_cell_t* cell, const srsue_gpu_pdsch_cfg_t* cfg,
                                int max_batch, srsue_gpu_pdsch_plan_t** out) {
  if (!ctx || !cell || !cfg || !out || max_batch < 1) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "plan_create: bad arguments");
  *out = nullptr;
  const int nfft = symbol_sz(cell->nof_prb);
  if (nfft < 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "nof_prb=%d is not an LTE bandwidth", cell->nof_prb);
  if (cell->nof_ports != 1 && cell->nof_ports != 2 && cell->nof_ports != 4) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "nof_ports must be 1, 2 or 4");
  if (cfg->qm != 2 && cfg->qm != 4 && cfg->qm != 6) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "qm must be 2, 4 or 6");
  if (cfg->tm == 2 && cell->nof_ports < 2) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "transmit diversity needs 2 or 4 ports");
  if (cfg->tm == 1 && cell->nof_ports == 4 && cfg->tbs != 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "a four-port cell transmits with diversity (tm 2)");
  if (cfg->tm != 1 && cfg->tm != 2) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tm must be 1 or 2");
  if (cfg->rv < 0 || cfg->rv > 3 || cfg->sf_idx < 0 || cfg->sf_idx > 9 || cfg->cfi < 1 || cfg->cfi > 3)
    return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "rv/sf_idx/cfi out of range");
  CU_CHECK(cudaSetDevice(ctx->device));
  auto* p = new srsue_gpu_pdsch_plan();
  p->ctx = ctx;
  p->cell = CellCfg{cell->nof_prb, cell->nof_ports, cell->cell_id, cell->cp ? 1 : 0};
  static_assert(sizeof(PdschCfg) == sizeof(srsue_gpu_pdsch_cfg_t), "config layouts must match");
  std::memcpy(&p->cfg, cfg, sizeof(PdschCfg));
  if (cfg->tbs == 0) {
    // front-end-only plan: OFDM demodulation and channel estimation for (cell, sf_idx); no grant yet
    // (what srslte_ue_dl_decode_fft_estimate needs before the DCI is known, phch_worker.cc:254)
    const int nsc0 = 12 * cell->nof_prb;
    p->info.nfft = nfft; p->info.nsc = nsc0; p->info.sf_len = 15 * nfft; p->info.max_batch = max_batch;
    std::vector<int8_t> crs0;
    build_crs_tables(p, cfg->sf_idx, crs0);
    std::vector<float> tw0;
    fft_twiddles(nfft, tw0);
    if (upload(&p->d_crs, crs0) != cudaSuccess || upload(&p->d_tw, tw0) != cudaSuccess) {
      srsue_gpu_pdsch_plan_destroy(p);
      return fail(SRSUE_GPU_ERROR, "plan_create: device allocation failed");
    }
    *out = p;
    return 0;
  }
  if (!cbsegm(cfg->tbs, &p->seg)) { delete p; return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tbs=%d cannot be segmented", cfg->tbs); }
  const CbSegm& s = p->seg;
  std::vector<int32_t> re;
  pdsch_re_list(p->cell, p->cfg, re);
  const int nre = (int)re.size(), G = nre * cfg->qm, nl = (cfg->tm == 2) ? 2 : 1;
  if (nre == 0 || G / (nl * cfg->qm) < s.C) { delete p; return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "grant has too few resource elements"); }
  p->gp = turbo_geom(s.Kp);
  p->gm = s.Cm ? turbo_geom(s.Km) : p->gp;
  const int nsc = 12 * cell->nof_prb;
  srsue_gpu_plan_info_t& I = p->info;
  I.nfft = nfft; I.nsc = nsc; I.sf_len = 15 * nfft; I.nof_re = nre; I.G = G;
  I.C = s.C; I.Kp = s.Kp; I.Km = s.Km; I.Cp = s.Cp; I.Cm = s.Cm; I.F = s.F;
  I.sb_cb_stride = std::max(p->gp.cb_elems, p->gm.cb_elems);
  I.sb_sf_stride = I.sb_cb_stride * s.C;
  I.payload_stride = (cfg->tbs + 7) / 8;
  I.max_batch = max_batch;
  // rate-dematching tables
  p->gather_stride = I.sb_cb_stride;
  std::vector<uint16_t> gather((size_t)s.C * p->gather_stride, 0xFFFF), tab;
  std::vector<int32_t> e_start(s.C + 1, 0), geom(4 * s.C, 0);
  // "direct" tables (every soft-buffer element receives at most one LLR, E <= N for all code blocks): the
  // entry is the index of that LLR inside the code block's range, E for "nothing" and E+1 for filler, so the
  // kernel gathers branch-free from a shared array that ends in the two sentinels {0, -C}.
  p->direct = 1;
  for (int r = 0; r < s.C; r++) {
    const TurboGeom g = turbo_geom(cb_len(s, r));
    const int N = 3 * (g.K + 4) - ((r == 0) ? 2 * s.F : 0);
    if (cb_E(s, G, cfg->qm, nl, r) > N) p->direct = 0;
  }
  // transport-block assembly: byte ranges of every code block's payload and CRC24A chunk shifts
  std::vector<int32_t> tbmap(4 * s.C, 0);
  std::vector<uint32_t> tbshift((size_t)s.C * 32, 0);
  const int tb_bytes = (cfg->tbs + 24) / 8;
  int tb_pos = 0;
  for (int r = 0; r < s.C; r++) {
    const int K = cb_len(s, r), F = (r == 0) ? s.F : 0;
    const TurboGeom g = turbo_geom(K);
    const int N = rm_gather_table(g, F, cfg->rv, tab);
    const int E = cb_E(s, G, cfg->qm, nl, r);
    if (p->direct)
      for (auto& v : tab) v = (v == 0xFFFE) ? (uint16_t)(E + 1) : (v >= E ? (uint16_t)E : v);
    std::copy(tab.begin(), tab.end(), gather.begin() + (size_t)r * p->gather_stride);
    e_start[r + 1] = e_start[r] + E;
    p->max_E = std::max(p->max_E, E);
    geom[4 * r] = g.cb_elems; geom[4 * r + 1] = N; geom[4 * r + 2] = K;
    const int src_off = F / 8, nb = K / 8 - (s.C > 1 ? 3 : 0) - F / 8, chunk = (nb + 31) / 32;
    tbmap[4 * r] = src_off; tbmap[4 * r + 1] = nb; tbmap[4 * r + 2] = tb_pos; tbmap[4 * r + 3] = chunk;
    for (int l = 0; l < 32; l++) {
      const int end = std::min(nb, (l + 1) * chunk);           // stream bytes after this lane's chunk
      tbshift[(size_t)r * 32 + l] = crc_xpow(kCrc24A, (uint64_t)8 * (tb_bytes - (tb_pos + end)));
    }
    tb_pos += nb;
  }
  if (e_start[s.C] != G || tb_pos != tb_bytes) { delete p; return fail(SRSUE_GPU_ERROR, "internal: rate-matching sizes do not add up"); }
  std::vector<uint32_t> scr;
  gold_packed(((uint32_t)cfg->rnti << 14) | ((uint32_t)cfg->sf_idx << 9) | (uint32_t)cell->cell_id, G, scr);
  scr.push_back(0u);      // the kernel reads bit groups with a two-word funnel shift
  std::vector<int8_t> crs;
  build_crs_tables(p, cfg->sf_idx, crs);
  std::vector<float> tw;
  fft_twiddles(nfft, tw);
  std::vector<int32_t> list_m((size_t)max_batch * s.Cm), list_p((size_t)max_batch * s.Cp);
  for (int sf = 0; sf < max_batch; sf++) {
    for (int r = 0; r < s.Cm; r++) list_m[(size_t)sf * s.Cm + r] = sf * s.C + r;
    for (int r = 0; r < s.Cp; r++) list_p[(size_t)sf * s.Cp + r] = sf * s.C + s.Cm + r;
  }
  bool ok = upload(&p->d_re, re) == cudaSuccess && upload(&p->d_scr, scr) == cudaSuccess &&
            upload(&p->d_gather, gather) == cudaSuccess && upload(&p->d_e_start, e_start) == cudaSuccess &&
            upload(&p->d_cb_geom, geom) == cudaSuccess && upload(&p->d_crs, crs) == cudaSuccess &&
            upload(&p->d_tw, tw) == cudaSuccess && upload(&p->d_list_m, list_m) == cudaSuccess &&
            upload(&p->d_list_p, list_p) == cudaSuccess && upload(&p->d_tbmap, tbmap) == cudaSuccess &&
            upload(&p->d_tbshift, tbshift) == cudaSuccess;
  const size_t B = (size_t)max_batch;
  ok = ok && cudaMalloc((void**)&p->d_sf, B * 14 * nsc * sizeof(float2)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_ce, B * cell->nof_ports * 14 * nsc * sizeof(float2)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_meas, B * 5 * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_pil, B * cell->nof_ports * 4 * 2 * cell->nof_prb * sizeof(float2)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_sb, B * I.sb_sf_stride * sizeof(int16_t)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_cb_bits, B * s.C * (s.Kp / 8)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_cb_status, B * s.C * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) == cudaSuccess;
  if (!ok) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    srsue_gpu_pdsch_plan_destroy(p);
    return fail(SRSUE_GPU_ERROR, "plan_create: device allocation failed: %s", msg);
  }
  // the kernel also has a few hundred bytes of static shared memory: leave room for it
  CU_CHECK(cudaFuncSetAttribute(pdsch_llr_dematch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin - 1024));
  *out = p;
  return 0;
}

void srsue_gpu_pdsch_plan_destroy(srsue_gpu_pdsch_plan_t* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaFree(p->d_re); cudaFree(p->d_scr); cudaFree(p->d_gather); cudaFree(p->d_e_start); cudaFree(p->d_cb_geom);
  cudaFree(p->d_pd_re4); cudaFree(p->d_pd_src); cudaFree(p->d_pd_scr);
  cudaFree(p->d_pbch_re); cudaFree(p->d_pbch_scr); cudaFree(p->d_pbch_rm);
  for (auto& kv : p->d_pd_rm) cudaFree(kv.second);
  cudaFree(p->d_crs); cudaFree(p->d_tw); cudaFree(p->d_list_m); cudaFree(p->d_list_p); cudaFree(p->d_tbmap); cudaFree(p->d_tbshift);
  cudaFree(p->d_sf); cudaFree(p->d_ce); cudaFree(p->d_pil); cudaFree(p->d_meas); cudaFree(p->d_sb); cudaFree(p->d_cb_bits);
  cudaFree(p->d_cb_status); cudaFree(p->d_iq); cudaFree(p->d_payload); cudaFree(p->d_tb_status);
  p->scratch.release();
  if (p->stream) cudaStreamDestroy(p->stream);
  if (p->stream2) { cudaStreamDestroy(p->stream2); for (auto& e : p->ev) cudaEventDestroy(e); }
  delete p;
}

int srsue_gpu_pdsch_plan_info(const srsue_gpu_pdsch_plan_t* p, srsue_gpu_plan_info_t* info) {
  if (!p || !info) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  *info = p->info;
  return 0;
}

#define PLAN_CHECK(p, n)                                                                               \
  do {                                                                                                 \
    if (!(p)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");                                \
    if ((n) < 0 || (n) > (p)->info.max_batch) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "n_sf=%d outside [0, max_batch=%d]", (n), (p)->info.max_batch); \
    if ((n) == 0) return 0;                                                                            \
    CU_CHECK(cudaSetDevice((p)->ctx->device));                                                         \
  } while (0)

int srsue_gpu_host_cfo_step(float cfo, int nfft) { return nfft > 0 ? cfo_step(cfo, nfft) : 0; }

static int ofdm_launch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_iq, const int16_t* d_iq16, float iq16_scale,
                       srsue_gpu_cf_t* d_sf, const int32_t* d_cfo_steps, int32_t cfo_step_all, void* stream);

int srsue_gpu_ofdm_rx(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_iq, srsue_gpu_cf_t* d_sf, void* stream) {
  return ofdm_launch(p, n_sf, d_iq, nullptr, 0.f, d_sf, nullptr, 0, stream);
}

int srsue_gpu_ofdm_rx_cfo(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_iq, srsue_gpu_cf_t* d_sf,
                          const int32_t* d_cfo_steps, int32_t cfo_step_all, void* stream) {
  return ofdm_launch(p, n_sf, d_iq, nullptr, 0.f, d_sf, d_cfo_steps, cfo_step_all, stream);
}

int srsue_gpu_ofdm_rx_sc16(srsue_gpu_pdsch_plan_t* p, int n_sf, const int16_t* d_iq16, float scale, srsue_gpu_cf_t* d_sf, void* stream) {
  return srsue_gpu_ofdm_rx_sc16_cfo(p, n_sf, d_iq16, scale, d_sf, nullptr, 0, stream);
}

int srsue_gpu_ofdm_rx_sc16_cfo(srsue_gpu_pdsch_plan_t* p, int n_sf, const int16_t* d_iq16, float scale, srsue_gpu_cf_t* d_sf,
                               const int32_t* d_cfo_steps, int32_t cfo_step_all, void* stream) {
  if (!(scale > 0.f)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ofdm_rx_sc16: scale must be positive");
  return ofdm_launch(p, n_sf, nullptr, d_iq16, scale, d_sf, d_cfo_steps, cfo_step_all, stream);
}

int srsue_gpu_pdsch_plan_set_cfo(srsue_gpu_pdsch_plan_t* p, const int32_t* d_cfo_steps, int32_t cfo_step) {
  if (!p) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");
  p->cfo_steps = d_cfo_steps; p->cfo_step_all = cfo_step;
  return 0;
}

int srsue_gpu_pdsch_plan_set_row_filter(srsue_gpu_pdsch_plan_t* p, const int32_t* d_values, int want) {
  if (!p) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");
  p->row_filter = d_values; p->row_want = want;
  return 0;
}

int srsue_gpu_pdsch_plan_set_min_iter(srsue_gpu_pdsch_plan_t* p, int min_iter) {
  if (!p || min_iter < 1) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "set_min_iter: null plan or min_iter < 1");
  p->min_iter = min_iter;
  return 0;
}

int srsue_gpu_pdsch_plan_set_iq_format(srsue_gpu_pdsch_plan_t* p, int format, float scale) {
  if (!p || (format != SRSUE_GPU_IQ_CF32 && format != SRSUE_GPU_IQ_SC16) || (format == SRSUE_GPU_IQ_SC16 && !(scale > 0.f)))
    return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "set_iq_format: format 0 (cf32) or 1 (sc16 with a positive scale)");
  if (p->d_iq && format != p->iq_format) {
    // the staging buffers of srsue_gpu_pdsch_decode_batch_host are sized for the old sample format
    CU_CHECK(cudaSetDevice(p->ctx->device));
    CU_CHECK(cudaDeviceSynchronize());
    cudaFree(p->d_iq); p->d_iq = nullptr;
    cudaFree(p->d_payload); p->d_payload = nullptr;
    cudaFree(p->d_tb_status); p->d_tb_status = nullptr;
  }
  p->iq_format = format;
  if (format == SRSUE_GPU_IQ_SC16) p->iq16_scale = scale;
  return 0;
}

static int ofdm_launch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_iq, const int16_t* d_iq16, float iq16_scale,
                       srsue_gpu_cf_t* d_sf, const int32_t* d_cfo_steps, int32_t cfo_step_all, void* stream) {
  PLAN_CHECK(p, n_sf);
  if ((!d_iq && !d_iq16) || !d_sf) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ofdm_rx: null buffer");
  const bool rotate = d_cfo_steps != nullptr || cfo_step_all != 0;
  OfdmArgs a{};
  a.iq16 = reinterpret_cast<const short2*>(d_iq16); a.iq16_scale = iq16_scale;
  if (rotate) {
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    if (!p->ctx->d_cexp) {
      std::vector<float> tab;
      cfo_table(tab);
      CU_CHECK(upload(reinterpret_cast<float**>(&p->ctx->d_cexp), tab));
    }
    a.cexp = p->ctx->d_cexp; a.cfo_steps = d_cfo_steps; a.cfo_step = cfo_step_all;
  }
  a.iq = reinterpret_cast<const float2*>(d_iq); a.sf_symbols = reinterpret_cast<float2*>(d_sf);
  a.tw = reinterpret_cast<const float2*>(p->d_tw);
  a.nfft = p->info.nfft; a.nsc = p->info.nsc; a.n_sf = n_sf; a.cp_ext = p->cell.cp;
  const int nsym = nof_symb(p->cell.cp);     // grid x: one CTA per OFDM symbol
  a.log2n = 0; while ((1 << a.log2n) < a.nfft) a.log2n++;
  a.scale = (float)(1.0 / std::sqrt((double)a.nfft));
  a.c3 = (float)(std::sqrt(3.0) / 2.0);
  const int threads = std::max(32, a.nfft / 8);       // 8 points per thread (16 and 32 measured slower)
  // 1536 = 3 x 512: three sub-transforms side by side in each of the two buffers
  const int smem = (a.nfft == 1536 ? 2 * 3 * (512 + 512 / 16 + 8) : 2 * (a.nfft + a.nfft / 16 + 8)) * (int)sizeof(float2);
  for (int done = 0; done < n_sf; done += 65535) {
    const int n = std::min(65535, n_sf - done);
    OfdmArgs b = a;
    if (b.iq) b.iq += (size_t)done * 15 * a.nfft;
    if (b.iq16) b.iq16 += (size_t)done * 15 * a.nfft;
    b.sf_symbols += (size_t)done * 14 * a.nsc; b.n_sf = n;
    // single exchange buffer (8 CTAs per SM) wherever the CTA has exactly N/8 threads; SRSUE_FFT_INPLACE=0 selects the
    // two-buffer kernel for comparison
    static const int inplace = getenv("SRSUE_FFT_INPLACE") ? atoi(getenv("SRSUE_FFT_INPLACE")) : 1;
    // N = 2048 without rotation: radix 16 x 16 x 8 with 16 points per thread (two exchanges instead of three, per-pass
    // twiddle tables); SRSUE_FFT_R16=0 selects the radix-8 kernel for comparison
    static const int r16 = getenv("SRSUE_FFT_R16") ? atoi(getenv("SRSUE_FFT_R16")) : 1;
    if (r16 && a.nfft == 2048 && !rotate) {
      const int smem16 = (2048 + 2048 / 16) * (int)sizeof(float2);
      if (b.iq16) ofdm_rx_r16_iq16_kernel<<<dim3(nsym, n), 128, smem16, (cudaStream_t)stream>>>(b);
      else ofdm_rx_r16_kernel<<<dim3(nsym, n), 128, smem16, (cudaStream_t)stream>>>(b);
    } else if (b.iq16 && rotate) {
      if (b.cfo_steps) b.cfo_steps += done;
      ofdm_rx_cfo_iq16_kernel<<<dim3(nsym, n), threads, smem, (cudaStream_t)stream>>>(b);
    } else if (b.iq16) {
      if (inplace && a.nfft != 1536 && a.nfft >= 256) ofdm_rx_inplace_iq16_kernel<<<dim3(nsym, n), threads, smem / 2, (cudaStream_t)stream>>>(b);
      else ofdm_rx_iq16_kernel<<<dim3(nsym, n), threads, smem, (cudaStream_t)stream>>>(b);
    } else if (rotate) {
      if (b.cfo_steps) b.cfo_steps += done;
      ofdm_rx_cfo_kernel<<<dim3(nsym, n), threads, smem, (cudaStream_t)stream>>>(b);
    } else if (inplace && a.nfft != 1536 && a.nfft >= 256) ofdm_rx_inplace_kernel<<<dim3(nsym, n), threads, smem / 2, (cudaStream_t)stream>>>(b);
    else ofdm_rx_kernel<<<dim3(nsym, n), threads, smem, (cudaStream_t)stream>>>(b);
    p->ctx->launch_count++;
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

static int chest_launch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, srsue_gpu_cf_t* d_ce,
                        srsue_gpu_cf_t* d_pilots, float* d_meas, void* stream);

int srsue_gpu_chest(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, srsue_gpu_cf_t* d_ce, float* d_meas, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_sf || !d_ce) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "chest: null buffer");
  return chest_launch(p, n_sf, d_sf, d_ce, nullptr, d_meas, stream);
}

int srsue_gpu_pcfich_decode(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                            const float* d_meas, float noise_est, int noise_mode, int32_t* d_cfi, int32_t* d_corr, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_sf || !d_ce || !d_cfi || (noise_mode && !d_meas)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pcfich_decode: null buffer");
  PcfichArgs a{};
  a.sf_symbols = reinterpret_cast<const float2*>(d_sf); a.ce = reinterpret_cast<const float2*>(d_ce); a.meas = d_meas;
  a.cfi = d_cfi; a.corr = d_corr;
  pcfich_re(p->cell, a.re);
  a.scramble = pcfich_scramble(p->cell, p->cfg.sf_idx);
  a.n_sf = n_sf; a.nsc = p->info.nsc; a.nof_ports = p->cell.nof_ports; a.noise_mode = noise_mode; a.noise_est = noise_est;
  a.k_sqpsk = (float)(100.0 * std::sqrt(2.0));
  a.k_sq2 = (float)std::sqrt(2.0);
  pcfich_kernel<<<(n_sf + 3) / 4, 128, 0, (cudaStream_t)stream>>>(a);
  p->ctx->launch_count++;
  CU_CHECK(cudaGetLastError());
  return 0;
}

namespace {
int pdcch_tables(srsue_gpu_pdsch_plan_t* p, int ng_x6) {
  if (ng_x6 != 1 && ng_x6 != 3 && ng_x6 != 6 && ng_x6 != 12) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ng_x6 must be 1, 3, 6 or 12 (6 x Ng)");
  if (p->pd_ng == ng_x6) return 0;
  // only the PDCCH tables depend on Ng; the PBCH tables of the plan (srsue_gpu_pbch_decode) are left alone
  cudaFree(p->d_pd_re4); cudaFree(p->d_pd_src); cudaFree(p->d_pd_scr);
  p->d_pd_re4 = nullptr; p->d_pd_src = nullptr; p->d_pd_scr = nullptr; p->pd_ng = -1;
  std::vector<int32_t> re4, src;
  const int n_reg = pdcch_regs(p->cell, p->cfg.cfi, ng_x6, re4);
  if (n_reg < 9) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "control region too small for a PDCCH");
  pdcch_quad_perm(n_reg, p->cell.cell_id, src);
  std::vector<uint32_t> scr;
  gold_packed(((uint32_t)p->cfg.sf_idx << 9) + (uint32_t)p->cell.cell_id, 8 * n_reg, scr);
  scr.push_back(0);
  CU_CHECK(upload(&p->d_pd_re4, re4));
  CU_CHECK(upload(&p->d_pd_src, src));
  CU_CHECK(upload(&p->d_pd_scr, scr));
  p->pd_ng = ng_x6; p->pd_nreg = n_reg;
  return 0;
}
}  // namespace

int srsue_gpu_pdcch_info(srsue_gpu_pdsch_plan_t* p, int ng_x6, int* n_reg, int* nof_cce) {
  if (!p) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");
  CU_CHECK(cudaSetDevice(p->ctx->device));
  int rc = pdcch_tables(p, ng_x6);
  if (rc) return rc;
  if (n_reg) *n_reg = p->pd_nreg;
  if (nof_cce) *nof_cce = p->pd_nreg / 9;
  return 0;
}

int srsue_gpu_pdcch_extract_llr(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                                const float* d_meas, float noise_est, int noise_mode, int ng_x6, int16_t* d_llr, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_sf || !d_ce || !d_llr || (noise_mode && !d_meas)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdcch_extract_llr: null buffer");
  int rc = pdcch_tables(p, ng_x6);
  if (rc) return rc;
  PdcchLlrArgs a{};
  a.sf_symbols = reinterpret_cast<const float2*>(d_sf); a.ce = reinterpret_cast<const float2*>(d_ce); a.meas = d_meas;
  a.re4 = p->d_pd_re4; a.src = p->d_pd_src; a.scramble = p->d_pd_scr; a.llr = d_llr;
  a.n_sf = n_sf; a.nsc = p->info.nsc; a.nof_ports = p->cell.nof_ports; a.n_reg = p->pd_nreg; a.noise_mode = noise_mode;
  a.noise_est = noise_est; a.k_sqpsk = (float)(100.0 * std::sqrt(2.0)); a.k_sq2 = (float)std::sqrt(2.0);
  a.row_filter = p->row_filter; a.row_want = p->row_want;
  for (int done = 0; done < n_sf; done += 65535) {
    const int n = std::min(65535, n_sf - done);
    PdcchLlrArgs b = a;
    if (b.row_filter) b.row_filter += done;
    b.sf_symbols += (size_t)done * 14 * a.nsc; b.ce += (size_t)done * a.nof_ports * 14 * a.nsc;
    if (b.meas) b.meas += (size_t)done * 5;
    b.llr += (size_t)done * 8 * a.n_reg; b.n_sf = n;
    pdcch_llr_kernel<<<dim3((a.n_reg + 127) / 128, n), 128, 0, (cudaStream_t)stream>>>(b);
    p->ctx->launch_count++;
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_pdcch_find_dci(srsue_gpu_pdsch_plan_t* p, int n_sf, const int16_t* d_llr, int ng_x6, int rnti, int common, int nof_bits,
                             int first_bit, int32_t* d_found, uint8_t* d_bits, uint16_t* d_rem, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_llr || !d_found || !d_bits) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdcch_find_dci: null buffer");
  if (nof_bits < 8 || nof_bits > 64 || rnti < 0 || rnti > 0xFFFF) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdcch_find_dci: nof_bits must be 8..64");
  int rc = pdcch_tables(p, ng_x6);
  if (rc) return rc;
  const int D = nof_bits + 16;
  auto it = p->d_pd_rm.find(D);
  if (it == p->d_pd_rm.end()) {
    std::vector<int32_t> seq;
    cc_rm_sequence(D, seq);
    int32_t* d = nullptr;
    CU_CHECK(upload(&d, seq));
    it = p->d_pd_rm.emplace(D, d).first;
  }
  PdcchSearchArgs a{};
  a.llr = d_llr; a.llr_stride = 8LL * p->pd_nreg; a.rm_seq = it->second; a.found = d_found; a.bits = d_bits; a.rem = d_rem;
  a.n_sf = n_sf; a.nof_bits = nof_bits; a.rnti = rnti; a.first_bit = first_bit < 0 ? -1 : (first_bit ? 1 : 0);
  a.row_filter = p->row_filter; a.row_want = p->row_want;
  a.n_cand = pdcch_search_space(p->pd_nreg / 9, p->cfg.sf_idx, (uint16_t)rnti, common != 0, a.cand_L, a.cand_ncce);
  if (a.n_cand == 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdcch_find_dci: empty search space");
  const int per_words = 3 * D + 4 * D + (D + 3) / 4;
  const int smem = a.n_cand * per_words * 4;
  if (!p->ctx->pdcch_attr_set) {        // per device: a process may hold contexts on several GPUs
    CU_CHECK(cudaFuncSetAttribute(pdcch_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    p->ctx->pdcch_attr_set = true;
  }
  pdcch_search_kernel<<<n_sf, 32 * a.n_cand, smem, (cudaStream_t)stream>>>(a);
  p->ctx->launch_count++;
  CU_CHECK(cudaGetLastError());
  return a.n_cand;
}

int srsue_gpu_phich_decode(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                           const float* d_meas, float noise_est, int noise_mode, int ng_x6, int n_group, int n_seq, int32_t* d_ack,
                           float* d_metric, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_sf || !d_ce || !d_ack || (noise_mode && !d_meas)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "phich_decode: null buffer");
  if (ng_x6 != 1 && ng_x6 != 3 && ng_x6 != 6 && ng_x6 != 12) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ng_x6 must be 1, 3, 6 or 12");
  if (n_group < 0 || n_group >= (p->cell.cp ? 2 : 1) * phich_groups(p->cell.nof_prb, ng_x6) || n_seq < 0 || n_seq > (p->cell.cp ? 3 : 7))
    return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "phich_decode: group %d / sequence %d out of range", n_group, n_seq);
  PhichArgs a{};
  a.sf_symbols = reinterpret_cast<const float2*>(d_sf); a.ce = reinterpret_cast<const float2*>(d_ce); a.meas = d_meas;
  a.ack = d_ack; a.metric = d_metric;
  phich_res(p->cell, n_group, a.re);
  a.scramble = pcfich_scramble(p->cell, p->cfg.sf_idx) & 0xFFFu;     // the same c_init as the PCFICH (36.211 6.9.1)
  a.n_sf = n_sf; a.nsc = p->info.nsc; a.nof_ports = p->cell.nof_ports; a.n_seq = n_seq; a.noise_mode = noise_mode;
  a.noise_est = noise_est; a.k_sq2 = (float)std::sqrt(2.0);
  a.ext = p->cell.cp; a.odd = n_group & 1; a.par0 = p->cell.cp ? n_group / 2 : n_group;
  phich_kernel<<<(n_sf + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  p->ctx->launch_count++;
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_pbch_decode(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                          const float* d_meas, float noise_est, int noise_mode, int32_t* d_result, uint8_t* d_mib, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_sf || !d_ce || !d_result || !d_mib || (noise_mode && !d_meas)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pbch_decode: null buffer");
  if (p->cfg.sf_idx != 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pbch_decode: the PBCH is in subframe 0 (plan has sf_idx %d)", p->cfg.sf_idx);
  if (!p->d_pbch_re) {
    std::vector<int32_t> re(240), seq;
    p->pbch_nre = pbch_res(p->cell, re.data());
    re.resize(p->pbch_nre);
    std::vector<uint32_t> scr;
    gold_packed((uint32_t)p->cell.cell_id, 8 * p->pbch_nre, scr);
    cc_rm_sequence(40, seq);
    CU_CHECK(upload(&p->d_pbch_re, re));
    CU_CHECK(upload(&p->d_pbch_scr, scr));
    CU_CHECK(upload(&p->d_pbch_rm, seq));
  }
  PbchArgs a{};
  a.sf_symbols = reinterpret_cast<const float2*>(d_sf); a.ce = reinterpret_cast<const float2*>(d_ce); a.meas = d_meas;
  a.re = p->d_pbch_re; a.scramble = p->d_pbch_scr; a.rm_seq = p->d_pbch_rm; a.result = d_result; a.mib = d_mib; a.n_re = p->pbch_nre;
  a.n_sf = n_sf; a.nsc = p->info.nsc; a.nof_ports = p->cell.nof_ports; a.noise_mode = noise_mode; a.noise_est = noise_est;
  a.k_sqpsk = (float)(100.0 * std::sqrt(2.0)); a.k_sq2 = (float)std::sqrt(2.0);
  pbch_kernel<<<n_sf, p->cell.nof_ports == 4 ? 384 : 256, 0, (cudaStream_t)stream>>>(a);      // four warps per transmit-port hypothesis
  p->ctx->launch_count++;
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_host_pbch_res(const srsue_gpu_cell_t* cell, int32_t* g240) {
  if (!cell || !g240 || cell->nof_prb < 6) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  const int n = pbch_res(CellCfg{cell->nof_prb, cell->nof_ports, cell->cell_id, cell->cp ? 1 : 0}, g240);
  for (int i = n; i < 240; i++) g240[i] = -1;       // extended cyclic prefix: 216 elements, the rest marked unused
  return 0;
}

int srsue_gpu_host_phich_index(int nof_prb, int ng_x6, int I_lowest, int n_dmrs, int* n_group, int* n_seq) {
  if (!n_group || !n_seq || nof_prb < 6 || I_lowest < 0 || n_dmrs < 0) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  phich_index(nof_prb, ng_x6, I_lowest, n_dmrs, n_group, n_seq);
  return 0;
}

int srsue_gpu_host_phich_index_cp(int nof_prb, int ng_x6, int cp, int I_lowest, int n_dmrs, int* n_group, int* n_seq) {
  if (!n_group || !n_seq || nof_prb < 6 || I_lowest < 0 || n_dmrs < 0) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  phich_index(nof_prb, ng_x6, I_lowest, n_dmrs, n_group, n_seq, cp ? 1 : 0);
  return 0;
}

int srsue_gpu_host_phich_res(const srsue_gpu_cell_t* cell, int n_group, int32_t* k12) {
  if (!cell || !k12 || n_group < 0) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  phich_res(CellCfg{cell->nof_prb, cell->nof_ports, cell->cell_id, cell->cp ? 1 : 0}, n_group, k12);
  return 0;
}

int srsue_gpu_host_pdcch_regs(const srsue_gpu_cell_t* cell, int cfi, int ng_x6, int32_t* re4) {
  if (!cell || cfi < 1 || cfi > 3) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  std::vector<int32_t> v;
  const int n = pdcch_regs(CellCfg{cell->nof_prb, cell->nof_ports, cell->cell_id, cell->cp ? 1 : 0}, cfi, ng_x6, v);
  if (re4) std::copy(v.begin(), v.end(), re4);
  return n;
}

int srsue_gpu_host_pdcch_quad_perm(int n_quad, int cell_id, int32_t* src) {
  if (n_quad < 1 || !src) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  std::vector<int32_t> v;
  pdcch_quad_perm(n_quad, cell_id, v);
  std::copy(v.begin(), v.end(), src);
  return 0;
}

int srsue_gpu_host_pdcch_search_space(int nof_cce, int sf_idx, int rnti, int common, int32_t* cand_L, int32_t* cand_ncce) {
  if (!cand_L || !cand_ncce || nof_cce < 1) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  return pdcch_search_space(nof_cce, sf_idx, (uint16_t)rnti, common != 0, cand_L, cand_ncce);
}

int srsue_gpu_host_dci_format_sizeof(int fmt, int nof_prb) { return (fmt == 0 || fmt == 1) ? dci_format_sizeof(fmt, nof_prb) : SRSUE_GPU_ERROR_INVALID_INPUTS; }

int srsue_gpu_host_pcfich_re(const srsue_gpu_cell_t* cell, int32_t* k16) {
  if (!cell || !k16 || cell->nof_prb < 6) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  pcfich_re(CellCfg{cell->nof_prb, cell->nof_ports, cell->cell_id, cell->cp ? 1 : 0}, k16);
  return 0;
}

int srsue_gpu_chest_pilots(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, srsue_gpu_cf_t* d_pilots, float* d_meas,
                           void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_sf || !d_pilots) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "chest_pilots: null buffer");
  return chest_launch(p, n_sf, d_sf, nullptr, d_pilots, d_meas, stream);
}

static int chest_launch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, srsue_gpu_cf_t* d_ce,
                        srsue_gpu_cf_t* d_pilots, float* d_meas, void* stream) {
  ChestArgs a{};
  a.sf_symbols = reinterpret_cast<const float2*>(d_sf); a.ce = reinterpret_cast<float2*>(d_ce); a.meas = d_meas;
  a.pilots = reinterpret_cast<float2*>(d_pilots);
  a.crs_sign = p->d_crs; a.n_sf = n_sf; a.nsc = p->info.nsc; a.nof_prb = p->cell.nof_prb; a.nof_ports = p->cell.nof_ports;
  std::memcpy(a.crs_off, p->crs_off, sizeof(a.crs_off));
  const int smem = 2 * p->cell.nof_ports * 4 * 2 * p->cell.nof_prb * (int)sizeof(float2) + 4 * p->info.nsc * (int)sizeof(float);
  // 128 threads measured best on B200 (0.19 ms vs 0.26 ms per 4096 subframes with 512): the kernel is bound by its
  // three ordered reduction warps, smaller CTAs pack more of them per SM.  Needs >= 3 warps.
  static const int chest_threads = std::min(128, std::max(96, getenv("SRSUE_CHEST_THREADS") ? atoi(getenv("SRSUE_CHEST_THREADS")) : 128));
  if (p->cell.nof_ports == 4) {
    if (d_pilots) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "chest_pilots: the fused interpolation serves one and two ports");
    static std::once_flag once;
    std::call_once(once, [] {
      cudaFuncSetAttribute(chest_p4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(chest_ext_p4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    static const int p4_threads = std::min(512, std::max(96, getenv("SRSUE_CHEST_P4_THREADS") ? atoi(getenv("SRSUE_CHEST_P4_THREADS")) : 256));     // 256 measured best (0.63 -> 0.54 ms per 4096 subframes at 100 PRB: 70 KB of shared memory allow 3 CTAs per SM, so larger CTAs are the way to more warps)
    if (p->cell.cp) chest_ext_p4_kernel<<<n_sf, p4_threads, smem, (cudaStream_t)stream>>>(a);
    else chest_p4_kernel<<<n_sf, p4_threads, smem, (cudaStream_t)stream>>>(a);
  } else if (p->cell.cp) chest_ext_kernel<<<n_sf, chest_threads, smem, (cudaStream_t)stream>>>(a);
  else chest_kernel<<<n_sf, chest_threads, smem, (cudaStream_t)stream>>>(a);
  p->ctx->launch_count++;
  CU_CHECK(cudaGetLastError());
  return 0;
}

static int llr_launch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                      const srsue_gpu_cf_t* d_pilots, const float* d_meas, float noise_est, int noise_mode, int accumulate,
                      int16_t* d_softbuf, srsue_gpu_cf_t* d_dbg_d, int16_t* d_dbg_e, void* stream);

int srsue_gpu_pdsch_llr(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                        const float* d_meas, float noise_est, int noise_mode, int accumulate, int16_t* d_softbuf,
                        srsue_gpu_cf_t* d_dbg_d, int16_t* d_dbg_e, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_ce) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_llr: null buffer");
  return llr_launch(p, n_sf, d_sf, d_ce, nullptr, d_meas, noise_est, noise_mode, accumulate, d_softbuf, d_dbg_d, d_dbg_e, stream);
}

int srsue_gpu_pdsch_llr_fused(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_pilots,
                              const float* d_meas, float noise_est, int noise_mode, int accumulate, int16_t* d_softbuf, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_pilots) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_llr_fused: null buffer");
  return llr_launch(p, n_sf, d_sf, nullptr, d_pilots, d_meas, noise_est, noise_mode, accumulate, d_softbuf, nullptr, nullptr, stream);
}

static int llr_launch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_sf, const srsue_gpu_cf_t* d_ce,
                      const srsue_gpu_cf_t* d_pilots, const float* d_meas, float noise_est, int noise_mode, int accumulate,
                      int16_t* d_softbuf, srsue_gpu_cf_t* d_dbg_d, int16_t* d_dbg_e, void* stream) {
  if (!d_sf || !d_softbuf) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_llr: null buffer");
  if (p->info.C == 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_llr: front-end-only plan (tbs == 0)");
  if (noise_mode && !d_meas) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_llr: noise_mode 1 needs d_meas");
  if (d_pilots && p->cell.nof_ports == 4) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_llr_fused: the fused interpolation serves one and two ports");
  DemodArgs a{};
  a.sf_symbols = reinterpret_cast<const float2*>(d_sf); a.ce = reinterpret_cast<const float2*>(d_ce); a.meas = d_meas;
  a.pilots = reinterpret_cast<const float2*>(d_pilots); a.nof_prb = p->cell.nof_prb; a.cp_ext = p->cell.cp;
  std::memcpy(a.crs_off, p->crs_off, sizeof(a.crs_off));       // ports 0 and 1 (the fused interpolation's)
  a.softbuf = d_softbuf; a.re_idx = p->d_re; a.scramble = p->d_scr; a.gather = p->d_gather;
  a.cb_e_start = p->d_e_start; a.cb_geom = p->d_cb_geom;
  a.dbg_d = reinterpret_cast<float2*>(d_dbg_d); a.dbg_e = d_dbg_e;
  a.n_sf = n_sf; a.nsc = p->info.nsc; a.nof_ports = p->cell.nof_ports; a.tm = p->cfg.tm; a.qm = p->cfg.qm;
  a.nof_re = p->info.nof_re; a.C = p->info.C; a.gather_stride = p->gather_stride; a.sb_stride = p->info.sb_cb_stride;
  a.noise_est = noise_est; a.noise_mode = noise_mode; a.accumulate = accumulate; a.direct = p->direct;
  a.k_sqpsk = (float)(100.0 * std::sqrt(2.0));
  a.k_c16 = (float)(2.0 * 400.0 / std::sqrt(10.0));
  a.k_c64a = (float)(4.0 * 700.0 / std::sqrt(42.0));
  a.k_c64b = (float)(2.0 * 700.0 / std::sqrt(42.0));
  a.k_sq2 = (float)std::sqrt(2.0);
  const int pil_elems = p->cell.nof_ports * 4 * 2 * p->cell.nof_prb;
  const int smem = (((p->max_E + 2 + 7) & ~7) * 2 + 15) / 16 * 16 + (d_pilots ? pil_elems * 8 : 0);
  if (smem > p->ctx->smem_optin - 1024) return fail(SRSUE_GPU_ERROR, "code block of %d LLRs does not fit in shared memory", p->max_E);
  for (int done = 0; done < n_sf; done += 65535) {
    const int n = std::min(65535, n_sf - done);
    DemodArgs b = a;
    b.sf_symbols += (size_t)done * 14 * a.nsc;
    if (b.ce) b.ce += (size_t)done * a.nof_ports * 14 * a.nsc;
    if (b.pilots) b.pilots += (size_t)done * pil_elems;
    if (b.meas) b.meas += (size_t)done * 5;
    b.softbuf += (size_t)done * p->info.sb_sf_stride;
    if (b.dbg_d) b.dbg_d += (size_t)done * a.nof_re;
    if (b.dbg_e) b.dbg_e += (size_t)done * p->info.G;
    b.n_sf = n;
    // 128 threads measured best on B200 (0.68 ms vs 0.85 ms per 4096 subframes with 512): finer occupancy granularity
    static const int demod_threads = std::min(160, std::max(32, getenv("SRSUE_DEMOD_THREADS") ? atoi(getenv("SRSUE_DEMOD_THREADS")) : 160));
    pdsch_llr_dematch_kernel<<<dim3(a.C, n), demod_threads, smem, (cudaStream_t)stream>>>(b);
    p->ctx->launch_count++;
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_pdsch_turbo(srsue_gpu_pdsch_plan_t* p, int n_sf, const int16_t* d_softbuf, int max_iter, uint8_t* d_payload,
                          int32_t* d_tb_status, int32_t* d_cb_status, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (!d_softbuf || !d_payload || !d_tb_status) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_turbo: null buffer");
  if (p->info.C == 0) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "pdsch_turbo: front-end-only plan (tbs == 0)");
  const CbSegm& s = p->seg;
  int32_t* cbst = d_cb_status ? d_cb_status : p->d_cb_status;
  const int crc_type = (s.C > 1) ? 2 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  if (s.Cm) rc = launch_turbo(p->ctx, p->scratch, d_softbuf, p->info.sb_cb_stride, p->d_list_m, n_sf * s.Cm, s.Km, max_iter, crc_type,
                              p->d_cb_bits, s.Kp / 8, cbst, st, p->min_iter);
  if (rc) return rc;
  rc = launch_turbo(p->ctx, p->scratch, d_softbuf, p->info.sb_cb_stride, p->d_list_p, n_sf * s.Cp, s.Kp, max_iter, crc_type, p->d_cb_bits,
                    s.Kp / 8, cbst, st, p->min_iter);
  if (rc) return rc;
  TbArgs t{};
  t.cb_bits = p->d_cb_bits; t.cb_status = cbst; t.payload = d_payload; t.tb_status = d_tb_status;
  t.n_sf = n_sf; t.C = s.C; t.Cm = s.Cm; t.Km = s.Km; t.Kp = s.Kp; t.F = s.F; t.tbs = s.tbs;
  t.cb_bits_stride = s.Kp / 8; t.payload_stride = p->info.payload_stride; t.tbmap = p->d_tbmap; t.tbshift = p->d_tbshift;
  tb_assemble_kernel<<<n_sf, 32 * s.C, 0, st>>>(t);
  p->ctx->launch_count++;
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_pdsch_decode_batch(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* d_iq, float noise_est, int noise_mode,
                                 int max_iter, int accumulate, int16_t* d_softbuf, uint8_t* d_payload, int32_t* d_tb_status,
                                 float* d_meas, void* stream) {
  PLAN_CHECK(p, n_sf);
  if (accumulate && !d_softbuf) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "HARQ combining needs a caller-owned soft buffer");
  p->ctx->launch_count = 0;
  float* meas = d_meas ? d_meas : p->d_meas;
  int16_t* sb = d_softbuf ? d_softbuf : p->d_sb;
  int rc = p->iq_format == SRSUE_GPU_IQ_SC16
               ? srsue_gpu_ofdm_rx_sc16_cfo(p, n_sf, reinterpret_cast<const int16_t*>(d_iq), p->iq16_scale, reinterpret_cast<srsue_gpu_cf_t*>(p->d_sf),
                                            p->cfo_steps, p->cfo_step_all, stream)
               : srsue_gpu_ofdm_rx_cfo(p, n_sf, d_iq, reinterpret_cast<srsue_gpu_cf_t*>(p->d_sf), p->cfo_steps, p->cfo_step_all, stream);
  // (the fused variant srsue_gpu_chest_pilots + srsue_gpu_pdsch_llr_fused moves 250 KB less per subframe but was
  // measured SLOWER on B200, 1.25 ms vs 1.08 ms per 4096 subframes: the demapper is issue-bound, not HBM-bound)
  if (!rc) rc = srsue_gpu_chest(p, n_sf, reinterpret_cast<srsue_gpu_cf_t*>(p->d_sf), reinterpret_cast<srsue_gpu_cf_t*>(p->d_ce), meas, stream);
  if (!rc) rc = srsue_gpu_pdsch_llr(p, n_sf, reinterpret_cast<srsue_gpu_cf_t*>(p->d_sf), reinterpret_cast<srsue_gpu_cf_t*>(p->d_ce), meas,
                                    noise_est, noise_mode, accumulate, sb, nullptr, nullptr, stream);
  if (!rc) rc = srsue_gpu_pdsch_turbo(p, n_sf, sb, max_iter, d_payload, d_tb_status, nullptr, stream);
  return rc;
}

int srsue_gpu_pdsch_decode_batch_host(srsue_gpu_pdsch_plan_t* p, int n_sf, const srsue_gpu_cf_t* h_iq, float noise_est,
                                      int noise_mode, int max_iter, uint8_t* h_payload, int32_t* h_tb_status, float* h_meas) {
  PLAN_CHECK(p, n_sf);
  if (!h_iq || !h_payload || !h_tb_status) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "decode_batch_host: null buffer");
  const size_t B = (size_t)p->info.max_batch;
  const size_t esz = p->iq_format == SRSUE_GPU_IQ_SC16 ? sizeof(short2) : sizeof(float2);     // bytes per sample of h_iq
  if (!p->d_iq) {
    CU_CHECK(cudaMalloc((void**)&p->d_iq, B * p->info.sf_len * esz));
    CU_CHECK(cudaMalloc((void**)&p->d_payload, B * p->info.payload_stride));
    CU_CHECK(cudaMalloc((void**)&p->d_tb_status, B * 4 * sizeof(int32_t)));
  }
  // Chunked pipeline: the copy stream uploads chunk c+1 while the compute stream decodes chunk c and returns
  // its results, so with pinned host memory the call runs at PCIe speed (245 760 B of IQ per subframe).
  if (!p->stream2) {
    CU_CHECK(cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking));
    for (auto& e : p->ev) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int kChunks = (int)(sizeof(p->ev) / sizeof(p->ev[0]));
  const int chunk = std::max(1, (n_sf + kChunks - 1) / kChunks);
  cudaStream_t sc = p->stream, sx = p->stream2;
  int n_ch = 0;
  for (int off = 0; off < n_sf; off += chunk, n_ch++) {
    const int n = std::min(chunk, n_sf - off);
    CU_CHECK(cudaMemcpyAsync(reinterpret_cast<char*>(p->d_iq) + (size_t)off * p->info.sf_len * esz,
                             reinterpret_cast<const char*>(h_iq) + (size_t)off * p->info.sf_len * esz,
                             (size_t)n * p->info.sf_len * esz, cudaMemcpyHostToDevice, sx));
    CU_CHECK(cudaEventRecord(p->ev[n_ch], sx));
  }
  n_ch = 0;
  for (int off = 0; off < n_sf; off += chunk, n_ch++) {
    const int n = std::min(chunk, n_sf - off);
    CU_CHECK(cudaStreamWaitEvent(sc, p->ev[n_ch], 0));
    const int32_t* steps_all = p->cfo_steps;              // per-subframe carrier-offset steps follow the chunk
    if (steps_all) p->cfo_steps = steps_all + off;
    int rc = srsue_gpu_pdsch_decode_batch(p, n, reinterpret_cast<srsue_gpu_cf_t*>(reinterpret_cast<char*>(p->d_iq) + (size_t)off * p->info.sf_len * esz), noise_est,
                                          noise_mode, max_iter, 0, nullptr, p->d_payload + (size_t)off * p->info.payload_stride,
                                          p->d_tb_status + (size_t)off * 4, p->d_meas + (size_t)off * 5, sc);
    p->cfo_steps = steps_all;
    if (rc) return rc;
    CU_CHECK(cudaMemcpyAsync(h_payload + (size_t)off * p->info.payload_stride, p->d_payload + (size_t)off * p->info.payload_stride,
                             (size_t)n * p->info.payload_stride, cudaMemcpyDeviceToHost, sc));
    CU_CHECK(cudaMemcpyAsync(h_tb_status + (size_t)off * 4, p->d_tb_status + (size_t)off * 4, (size_t)n * 4 * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, sc));
    if (h_meas) CU_CHECK(cudaMemcpyAsync(h_meas + (size_t)off * 5, p->d_meas + (size_t)off * 5, (size_t)n * 5 * sizeof(float),
                                         cudaMemcpyDeviceToHost, sc));
  }
  CU_CHECK(cudaStreamSynchronize(sc));
  return 0;
}

int srsue_gpu_host_cbsegm(int tbs, int* out) {
  CbSegm s;
  if (!out || !cbsegm(tbs, &s)) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "tbs=%d cannot be segmented", tbs);
  const int v[8] = {s.tbs, s.B, s.C, s.Kp, s.Km, s.Cp, s.Cm, s.F};
  std::memcpy(out, v, sizeof(v));
  return 0;
}

int srsue_gpu_host_pdsch_re(const srsue_gpu_cell_t* cell, const srsue_gpu_pdsch_cfg_t* cfg, int32_t* re_idx) {
  if (!cell || !cfg) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  CellCfg c{cell->nof_prb, cell->nof_ports, cell->cell_id, cell->cp ? 1 : 0};
  PdschCfg pc;
  std::memcpy(&pc, cfg, sizeof(pc));
  std::vector<int32_t> re;
  pdsch_re_list(c, pc, re);
  if (re_idx) std::copy(re.begin(), re.end(), re_idx);
  return (int)re.size();
}

int srsue_gpu_host_rm_sequence(int K, int F, int rv, int32_t* seq) {
  if (qpp_index(K) < 0 || !seq || rv < 0 || rv > 3 || F < 0) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  // invert the gather table: tcb offset -> read index, back to srsLTE order
  const TurboGeom g = turbo_geom(K);
  std::vector<uint16_t> tab;
  const int N = rm_gather_table(g, F, rv, tab);
  for (int tri = 0; tri < 3 * (K + 4); tri++) {
    const uint16_t n = tab[tcb_offset(g, tri)];
    if (n < 0xFFFE) seq[n] = tri;
  }
  return N;
}

int srsue_gpu_host_qpp(int K, uint16_t* pi) {
  int f1, f2;
  if (!pi || !qpp_params(K, &f1, &f2)) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  for (int64_t i = 0; i < K; i++) pi[i] = (uint16_t)((f1 * i + (int64_t)f2 * i * i) % K);
  return 0;
}

int srsue_gpu_host_gold(uint32_t c_init, int n, uint8_t* c) {
  if (!c || n < 0) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  gold_bits(c_init, n, c);
  return 0;
}

void* srsue_gpu_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(srsue::g_regions_mu);
  srsue::g_regions[reinterpret_cast<uintptr_t>(p)] = {(size_t)bytes, false};
  return p;
}
void srsue_gpu_host_free(void* p) {
  if (!p) return;
  { std::lock_guard<std::mutex> lk(srsue::g_regions_mu); srsue::g_regions.erase(reinterpret_cast<uintptr_t>(p)); }
  cudaFreeHost(p);
}
int srsue_gpu_host_register(void* p, uint64_t bytes) {
  if (!p || !bytes) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "host_register: null region");
  CU_CHECK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  std::lock_guard<std::mutex> lk(srsue::g_regions_mu);
  srsue::g_regions[reinterpret_cast<uintptr_t>(p)] = {(size_t)bytes, true};
  return 0;
}
int srsue_gpu_host_unregister(void* p) {
  {
    std::lock_guard<std::mutex> lk(srsue::g_regions_mu);
    auto it = srsue::g_regions.find(reinterpret_cast<uintptr_t>(p));
    if (it == srsue::g_regions.end() || !it->second.second) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "host_unregister: unknown region");
    srsue::g_regions.erase(it);
  }
  CU_CHECK(cudaHostUnregister(p));
  return 0;
}

}  // extern "C"

// ---- uplink shared-channel encoder (ulsch.cu) -------------------------------------------------------
struct srsue_gpu_ulsch_plan {
  srsue_gpu_ctx* ctx = nullptr;
  srsue_gpu_ulsch_cfg_t cfg{};
  CbSegm seg{};
  int G = 0, max_batch = 0, threads = 0;
  int32_t* d_cbtab = nullptr; int32_t* d_seq_len = nullptr; uint16_t* d_perm = nullptr; uint16_t* d_seq = nullptr;
  uint32_t* d_crcshift = nullptr; uint8_t* d_scramble = nullptr; uint8_t* d_tbcrc = nullptr; uint8_t* d_ebits = nullptr;
  uint8_t* d_payload = nullptr; uint8_t* d_out = nullptr;          // staging of the host-pointer call
};

int srsue_gpu_ulsch_plan_create(srsue_gpu_ctx_t* ctx, const srsue_gpu_ulsch_cfg_t* cfg, int max_batch, srsue_gpu_ulsch_plan_t** plan) {
  if (!ctx || !cfg || !plan || max_batch < 1) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch_plan_create: bad arguments");
  if (cfg->qm != 2 && cfg->qm != 4 && cfg->qm != 6) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch: qm=%d", cfg->qm);
  if (cfg->nof_prb < 1 || cfg->nof_prb > 110 || (cfg->n_symb != 12 && cfg->n_symb != 11 && cfg->n_symb != 10 && cfg->n_symb != 9))
    return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch: nof_prb=%d n_symb=%d", cfg->nof_prb, cfg->n_symb);
  if (cfg->tbs < 8 || cfg->tbs % 8 || cfg->rv < 0 || cfg->rv > 3 || cfg->sf_idx < 0 || cfg->sf_idx > 9)
    return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch: tbs=%d rv=%d sf_idx=%d", cfg->tbs, cfg->rv, cfg->sf_idx);
  CU_CHECK(cudaSetDevice(ctx->device));
  auto* p = new srsue_gpu_ulsch_plan;
  p->ctx = ctx; p->cfg = *cfg; p->max_batch = max_batch;
  if (!cbsegm(cfg->tbs, &p->seg)) { delete p; return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch: tbs=%d cannot be segmented", cfg->tbs); }
  const CbSegm& s = p->seg;
  const int rows = 12 * cfg->nof_prb, G = rows * cfg->n_symb * cfg->qm;
  p->G = G;
  std::vector<int32_t> cbtab(8 * s.C), seq_len(s.C);
  std::vector<uint16_t> perm, seq, one;
  std::vector<uint32_t> shift((size_t)(1 + s.C) * 32, 0u);
  // TB CRC24A: lane l of the warp takes bytes [l chunk, (l + 1) chunk) of the payload
  {
    const int nb = cfg->tbs / 8, chunk = (nb + 31) / 32;
    for (int l = 0; l < 32; l++) shift[l] = crc_xpow(kCrc24A, (uint64_t)8 * (nb - std::min(nb, (l + 1) * chunk)));
  }
  std::map<int, int> perm_of_K;
  std::map<std::pair<int, int>, std::pair<int, int>> seq_of;      // (K, F) -> (offset, length)
  int e_start = 0, pos = 0, maxK = 0;
  for (int r = 0; r < s.C; r++) {
    const int K = cb_len(s, r), F = (r == 0) ? s.F : 0, E = cb_E(s, G, cfg->qm, 1, r);
    const int nbs = K / 8 - F / 8 - (s.C > 1 ? 3 : 0);
    if (!perm_of_K.count(K)) {
      int f1 = 0, f2 = 0;
      qpp_params(K, &f1, &f2);
      perm_of_K[K] = (int)perm.size();
      for (int64_t i = 0; i < K; i++) perm.push_back((uint16_t)((f1 * i + (int64_t)f2 * i * i) % K));
    }
    if (!seq_of.count({K, F})) {
      rm_tx_sequence(K, F, cfg->rv, one);
      seq_of[{K, F}] = {(int)seq.size(), (int)one.size()};
      seq.insert(seq.end(), one.begin(), one.end());
    }
    const int32_t row[8] = {K, F, E, e_start, pos, nbs, perm_of_K[K], seq_of[{K, F}].first};
    std::copy(row, row + 8, cbtab.begin() + 8 * r);
    seq_len[r] = seq_of[{K, F}].second;
    if (s.C > 1) {
      const int nbc = K / 8 - 3, chunk = (nbc + 31) / 32;
      for (int l = 0; l < 32; l++) shift[(size_t)(1 + r) * 32 + l] = crc_xpow(kCrc24B, (uint64_t)8 * (nbc - std::min(nbc, (l + 1) * chunk)));
    }
    e_start += E; pos += nbs; maxK = std::max(maxK, K);
  }
  if (e_start != G || pos != (cfg->tbs + 24) / 8) { delete p; return fail(SRSUE_GPU_ERROR, "internal: uplink rate-matching sizes do not add up"); }
  p->threads = ((maxK / 8 + 31) / 32) * 32;
  std::vector<uint8_t> gold(G), scr(G / 8, 0);
  gold_bits(((uint32_t)cfg->rnti << 14) | ((uint32_t)cfg->sf_idx << 9) | (uint32_t)cfg->cell_id, G, gold.data());
  for (int i = 0; i < G; i++) scr[i >> 3] |= (uint8_t)(gold[i] << (7 - (i & 7)));
  bool ok = upload(&p->d_cbtab, cbtab) == cudaSuccess && upload(&p->d_seq_len, seq_len) == cudaSuccess &&
            upload(&p->d_perm, perm) == cudaSuccess && upload(&p->d_seq, seq) == cudaSuccess &&
            upload(&p->d_crcshift, shift) == cudaSuccess && upload(&p->d_scramble, scr) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_tbcrc, (size_t)max_batch * 4) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&p->d_ebits, (size_t)max_batch * G) == cudaSuccess;
  if (!ok) { srsue_gpu_ulsch_plan_destroy(p); return fail(SRSUE_GPU_ERROR, "ulsch_plan_create: device allocation failed"); }
  *plan = p;
  return 0;
}

void srsue_gpu_ulsch_plan_destroy(srsue_gpu_ulsch_plan_t* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaFree(p->d_cbtab); cudaFree(p->d_seq_len); cudaFree(p->d_perm); cudaFree(p->d_seq); cudaFree(p->d_crcshift);
  cudaFree(p->d_scramble); cudaFree(p->d_tbcrc); cudaFree(p->d_ebits); cudaFree(p->d_payload); cudaFree(p->d_out);
  delete p;
}

int srsue_gpu_ulsch_plan_info(const srsue_gpu_ulsch_plan_t* p, int* G, int* C, int* Kp, int* Km) {
  if (!p) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");
  if (G) *G = p->G;
  if (C) *C = p->seg.C;
  if (Kp) *Kp = p->seg.Kp;
  if (Km) *Km = p->seg.Km;
  return 0;
}

int srsue_gpu_ulsch_encode(srsue_gpu_ulsch_plan_t* p, int n_tb, const uint8_t* d_payload, uint8_t* d_bits, void* stream) {
  if (!p) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");
  if (n_tb < 0 || n_tb > p->max_batch) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "n_tb=%d outside [0, max_batch=%d]", n_tb, p->max_batch);
  if (n_tb == 0) return 0;
  if (!d_payload || !d_bits) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch_encode: null buffer");
  CU_CHECK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  UlschArgs a{};
  a.payload = d_payload; a.payload_stride = p->cfg.tbs / 8; a.tbcrc = p->d_tbcrc; a.ebits = p->d_ebits; a.out = d_bits;
  a.out_stride = p->G / 8; a.n_tb = n_tb; a.tbs = p->cfg.tbs; a.C = p->seg.C; a.G = p->G; a.qm = p->cfg.qm;
  a.rows = 12 * p->cfg.nof_prb; a.n_symb = p->cfg.n_symb;
  a.cbtab = p->d_cbtab; a.seq_len = p->d_seq_len; a.perm = p->d_perm; a.seq = p->d_seq; a.crcshift = p->d_crcshift; a.scramble = p->d_scramble;
  for (int done = 0; done < n_tb; done += 65535) {
    const int n = std::min(65535, n_tb - done);
    UlschArgs b = a;
    b.payload += (size_t)done * a.payload_stride; b.tbcrc += (size_t)done * 4; b.ebits += (size_t)done * a.G; b.out += (size_t)done * a.out_stride;
    b.n_tb = n;
    ulsch_tbcrc_kernel<<<n, 32, 0, st>>>(b);
    ulsch_encode_kernel<<<dim3(a.C, n), p->threads, 0, st>>>(b);
    ulsch_interleave_kernel<<<dim3((a.G / 8 + 255) / 256, n), 256, 0, st>>>(b);
    p->ctx->launch_count += 3;
  }
  CU_CHECK(cudaGetLastError());
  return 0;
}

int srsue_gpu_ulsch_encode_host(srsue_gpu_ulsch_plan_t* p, int n_tb, const uint8_t* h_payload, uint8_t* h_bits) {
  if (!p) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "null plan");
  if (n_tb < 0 || n_tb > p->max_batch) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "n_tb=%d outside [0, max_batch=%d]", n_tb, p->max_batch);
  if (n_tb == 0) return 0;
  if (!h_payload || !h_bits) return fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "ulsch_encode_host: null buffer");
  CU_CHECK(cudaSetDevice(p->ctx->device));
  const size_t pb = (size_t)p->cfg.tbs / 8, ob = (size_t)p->G / 8;
  if (!p->d_payload) CU_CHECK(cudaMalloc((void**)&p->d_payload, (size_t)p->max_batch * pb));
  if (!p->d_out) CU_CHECK(cudaMalloc((void**)&p->d_out, (size_t)p->max_batch * ob));
  CU_CHECK(cudaMemcpy(p->d_payload, h_payload, (size_t)n_tb * pb, cudaMemcpyHostToDevice));
  const int rc = srsue_gpu_ulsch_encode(p, n_tb, p->d_payload, p->d_out, nullptr);
  if (rc) return rc;
  CU_CHECK(cudaMemcpy(h_bits, p->d_out, (size_t)n_tb * ob, cudaMemcpyDeviceToHost));
  return 0;
}

// batch.cu -- the batching layer of libsrsue_gpu: takes an arbitrary stream of subframe descriptors (any mix of
// bandwidths, grants, redundancy versions and UEs), packs the ones that share a (cell, grant) shape into one launch
// each, and keeps per-(UE, HARQ process) soft buffers resident on the device.
//
// It replaces, for offline / batched operation, the hand-off of ONE subframe to ONE worker thread
// (/root/reference/ue/src/phy/phch_recv.cc:309-369 acquiring a worker, ue/src/common/thread_pool.cc:72-82,206-254
// running it) and the per-process soft-buffer ownership of the MAC (ue/src/mac/dl_harq.cc:169-174 init,
// :232 reset on new data, :216-259 rv selection).  Everything here sits above the public C ABI of api.cu
// (plans + srsue_gpu_pdsch_decode_batch); the only device code is the row gather/scatter of soft buffers.
#include <cuda_runtime.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <list>
#include <map>
#include <string>
#include <vector>

#include "srsue_gpu/srslte_compat.h"
#include "srsue_gpu/srsue_gpu.h"

namespace srsue {
int internal_fail(int code, const char* msg);   // api.cu: sets srsue_gpu_last_error()
bool host_region_contains(const void* p, size_t bytes);   // api.cu: inside a known pinned region?
}

namespace {

#define B_FAIL(code, ...)                                   \
  do {                                                      \
    char buf_[384];                                         \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);              \
    return srsue::internal_fail(code, buf_);                \
  } while (0)

#define B_CU(call)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) B_FAIL(SRSUE_GPU_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// copies rows between a packed [n][row_elems] array and per-row device buffers (rows[i] == nullptr: skip).
// One CTA per row; 16-byte accesses (row buffers come from cudaMalloc, row_elems is a multiple of 8).
__global__ void __launch_bounds__(256) softbuffer_rows_kernel(int16_t* packed, int16_t* const* rows, int row_elems, int to_rows) {
  int16_t* r = rows[blockIdx.x];
  if (r == nullptr) return;
  uint4* a = reinterpret_cast<uint4*>(packed + (size_t)blockIdx.x * row_elems);
  uint4* b = reinterpret_cast<uint4*>(r);
  const int n = row_elems / 8;
  if (to_rows) for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = a[i];
  else for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = b[i];
}

// Zero-copy gather of scattered subframes: rows[i] points into pinned host memory (UVA), the GPU pulls each row over
// PCIe into the packed device array.  V = 16 or 8 bytes per access.  The kernel is bound by the PCIe round trip, not
// by the SMs, and it runs NEXT TO the decoding chain of the previous chunk -- whose turbo decoder needs whole SMs (all
// registers of an SM per CTA).  So it is a small persistent grid (a few dozen CTAs walking the (row, slice) items with eight
// loads per thread in flight: enough bytes on the wire to fill the link) instead of one CTA per row and slice, which put
// long-lived CTAs on every SM and kept the decoder waiting (11.9 vs 15.4 Gbit/s against the copy-engine path).
template <typename V>
__global__ void __launch_bounds__(256) gather_host_rows_kernel(V* __restrict__ packed, const V* const* __restrict__ rows, int n_rows,
                                                               int row_vecs, int slices) {
  const int per = (row_vecs + slices - 1) / slices;
  for (int item = blockIdx.x; item < n_rows * slices; item += gridDim.x) {
    const int row = item / slices, sl = item - row * slices;
    const V* __restrict__ src = rows[row];
    V* dst = packed + (size_t)row * row_vecs;
    const int lo = sl * per, hi = min(row_vecs, lo + per);
    constexpr int U = 8;                                  // loads in flight per thread
    for (int i = lo + threadIdx.x; i < hi; i += U * blockDim.x) {
      V v[U];
#pragma unroll
      for (int u = 0; u < U; u++)
        if (i + u * (int)blockDim.x < hi) v[u] = src[i + u * blockDim.x];
#pragma unroll
      for (int u = 0; u < U; u++)
        if (i + u * (int)blockDim.x < hi) dst[i + u * blockDim.x] = v[u];
    }
  }
}
// CTAs of the gather grid (SRSUE_GATHER_CTAS overrides)
static int gather_ctas(int items) {
  static const int want = getenv("SRSUE_GATHER_CTAS") ? std::max(1, atoi(getenv("SRSUE_GATHER_CTAS"))) : 48;
  return std::max(1, std::min(items, want));
}

// Zero-copy scatter of the payload rows into the callers' pinned buffers.  WORD = 4 when every row and pointer is
// 4-byte aligned, else 1.
template <typename W>
__global__ void __launch_bounds__(128) scatter_host_rows_kernel(const W* __restrict__ packed, W* const* __restrict__ rows, int row_words) {
  W* dst = rows[blockIdx.x];
  const W* src = packed + (size_t)blockIdx.x * row_words;
  for (int i = threadIdx.x; i < row_words; i += blockDim.x) dst[i] = src[i];
}

struct PlanEntry {
  srsue_gpu_pdsch_plan_t* plan = nullptr;
  srsue_gpu_plan_info_t info{};
  std::list<std::string>::iterator lru;
  bool front = false;                             // front-end-only plan (no grant)
};

struct SoftBuffer { int16_t* d = nullptr; int elems = 0; };

}  // namespace

struct srsue_gpu_batch {
  srsue_gpu_ctx_t* ctx = nullptr;
  int max_subframes = 0, chunk_cap = 0, max_iter = 4, noise_mode = 0, max_plans = 16;
  float noise_est = 0.01f;
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_up[2] = {}, ev_free[2] = {};
  std::vector<cudaEvent_t> ev_blind;               // blind submissions: samples of group g are on the device
  cudaStream_t s_aux = nullptr;                    // blind submissions: the "copy" stream of the PDSCH phase (device-side gathers), next to the uploads on s_copy
  cudaStream_t s_ctrl = nullptr;                   // blind submissions: the control-channel passes (the host waits for each; the PDSCH chains of earlier groups run next to them)
  float* d_bmeas = nullptr;                        // their measurements (d_meas belongs to the PDSCH chains)
  size_t pl_off = 0;                               // payload staging used by the chunks submitted so far
  // plan cache keyed by the bytes of (cell, cfg)
  std::map<std::string, PlanEntry> plans;
  std::list<std::string> lru;
  // device staging (grown on demand)
  srsue_gpu_cf_t* d_iq[2] = {nullptr, nullptr}; size_t iq_elems[2] = {0, 0};
  uint8_t* d_payload = nullptr; size_t payload_bytes = 0;
  int32_t* d_status = nullptr; float* d_meas = nullptr;
  int16_t* d_sb = nullptr; size_t sb_elems = 0;
  int16_t** d_rows = nullptr;
  int iq_format = SRSUE_GPU_IQ_CF32; float iq16_scale = 1.0f / 32768.0f;   // what every descriptor's iq points at
  int32_t *h_cfo = nullptr, *d_cfo = nullptr;      // per-row carrier-offset phase steps of the chunk being launched
  // pinned result staging, in processing order
  int32_t* h_status = nullptr; float* h_meas = nullptr; int16_t** h_rows = nullptr;
  const void** h_iq_rows = nullptr; void** h_pl_rows = nullptr;    // pinned pointer tables read by the zero-copy kernels
  // pinned staging for scattered host buffers: small IQ rows are gathered by the CPU and uploaded with one copy per
  // chunk, payload rows come back with one copy per chunk and are scattered in srsue_gpu_batch_wait
  srsue_gpu_cf_t* h_iq[2] = {nullptr, nullptr}; size_t h_iq_elems[2] = {0, 0};
  uint8_t* h_pl = nullptr; size_t h_pl_bytes = 0;
  struct PlChunk { size_t pos0, stage_off, row_bytes; int rows; };
  std::vector<PlChunk> pl_chunks;
  std::vector<int> order;                 // processing position -> descriptor index
  srsue_gpu_sf_desc_t* pending = nullptr; int n_pending = 0;
  std::map<int64_t, SoftBuffer> softbuffers;
  int launches = 0, plan_launches = 0;
  int half = 0;
  // ---- blind submissions (srsue_gpu_batch_submit_blind): the samples of the whole submission stay on the device between
  // the control-channel phase and the PDSCH phase; phase 2 gathers its chunks from there instead of from host memory
  bool iq_on_device = false;
  char* d_biq = nullptr; size_t biq_bytes = 0;
  srsue_gpu_cf_t* d_bsf = nullptr; size_t bsf_elems = 0;
  srsue_gpu_cf_t* d_bce = nullptr; size_t bce_elems = 0;
  int16_t* d_bllr = nullptr; size_t bllr_elems = 0;
  int32_t* d_bcfi = nullptr; int32_t* d_bfound = nullptr; uint8_t* d_bbits = nullptr;     // [cap], [9][cap][4], [9][cap][64]
  int32_t* h_bcfi = nullptr; int32_t* h_bfound = nullptr; uint8_t* h_bbits = nullptr;     // pinned mirrors
  std::vector<srsue_gpu_sf_desc_t> blind_descs;   // phase-2 descriptors (grants filled in, iq = device row)
  std::vector<int> blind_index;                   // their positions in the caller's array
  srsue_gpu_sf_desc_t* blind_orig = nullptr;
  int blind_n = 0, blind_rows_used = 0;
  // ---- multi-GPU front (srsue_gpu_batch_create_multi): this object then owns no device state of its own.  Every device
  // has its own context, single-device batch (streams, pinned staging, plan cache, resident soft buffers) and host
  // thread; a submission is partitioned here, the parts run concurrently, results are gathered in batch_wait.
  // Reference analogue: N phch_workers behind one thread_pool (thread_pool.cc:206-254, phy.h:118-119).
  struct Dev {
    int device = 0;
    srsue_gpu_ctx_t* ctx = nullptr;
    srsue_gpu_batch* b = nullptr;
    std::thread th;
    std::vector<srsue_gpu_sf_desc_t> descs;       // this device's share of the submission (copies)
    std::vector<int> index;                       // position of each in the caller's array
    int rc = 0;
    std::string err;
    double work = 0;                              // estimated turbo work of the share (sum of C * K)
    int id = 0, share = 0;                        // position in devs; descriptors of the current submission
  };
  std::vector<int> where;                         // current submission: descriptor -> device (the workers copy their own shares)
  srsue_gpu_sf_desc_t* src = nullptr; int n_src = 0;
  int multi_blind = 0, multi_blind_run = 0, multi_ng_x6 = 6, multi_payload_cap = 0;   // how the workers submit their share
  std::vector<Dev*> devs;
  std::map<int64_t, int> affinity;                // soft buffer id -> device that holds it
  std::mutex mu;
  std::condition_variable cv;
  uint64_t generation = 0;                        // bumped by submit; workers run one submission per generation
  int running = 0;                                // workers that have not finished the current generation
  bool quit = false;
};

namespace {

// same launch shape, field for field (what plan_key() below strings together); false negatives only cost the slow path
bool same_shape(const srsue_gpu_sf_desc_t& a, const srsue_gpu_sf_desc_t& b) {
  return a.cell.nof_prb == b.cell.nof_prb && a.cell.nof_ports == b.cell.nof_ports && a.cell.cell_id == b.cell.cell_id &&
         (a.cell.cp ? 1 : 0) == (b.cell.cp ? 1 : 0) && a.cfg.sf_idx == b.cfg.sf_idx && a.cfg.cfi == b.cfg.cfi && a.cfg.rnti == b.cfg.rnti &&
         a.cfg.qm == b.cfg.qm && a.cfg.tbs == b.cfg.tbs && a.cfg.rv == b.cfg.rv && a.cfg.tm == b.cfg.tm &&
         std::memcmp(a.cfg.prb_mask, b.cfg.prb_mask, sizeof(a.cfg.prb_mask)) == 0;
}

std::string plan_key(const srsue_gpu_sf_desc_t& d) {
  // explicit fields only (struct padding must not split buckets)
  const int v[11] = {d.cell.nof_prb, d.cell.nof_ports, d.cell.cell_id, d.cell.cp ? 1 : 0, d.cfg.sf_idx, d.cfg.cfi, d.cfg.rnti, d.cfg.qm, d.cfg.tbs, d.cfg.rv, d.cfg.tm};
  std::string k(reinterpret_cast<const char*>(v), sizeof(v));
  for (int i = 0; i < 110; i++) k.push_back((char)(d.cfg.prb_mask[i] & 7));
  return k;
}

int get_plan(srsue_gpu_batch* b, const std::string& key, const srsue_gpu_sf_desc_t& d, PlanEntry** out) {
  auto it = b->plans.find(key);
  if (it != b->plans.end()) {
    b->lru.erase(it->second.lru);
    b->lru.push_front(key);
    it->second.lru = b->lru.begin();
    *out = &it->second;
    return 0;
  }
  // Two budgets: plans with a grant own per-subframe work buffers (hundreds of MB at 20 MHz) and are limited to max_plans;
  // front-end-only plans (tbs = 0: tables for FFT, estimate, PCFICH, PDCCH of one (cell, subframe, CFI)) are small and a
  // blind submission needs three per cell and subframe number, so they get their own, larger budget.
  const bool front = d.cfg.tbs == 0;
  int n_kind = 0;
  for (const auto& kv : b->plans) n_kind += kv.second.front == front;
  if (n_kind >= (front ? 8 * b->max_plans : b->max_plans)) {
    // evict the least recently used plan of this kind; its work may still be in flight on the compute stream
    B_CU(cudaStreamSynchronize(b->s_compute));
    auto it2 = b->lru.end();
    do { --it2; } while (b->plans[*it2].front != front);
    const std::string victim = *it2;
    b->lru.erase(it2);
    srsue_gpu_pdsch_plan_destroy(b->plans[victim].plan);
    b->plans.erase(victim);
  }
  PlanEntry e;
  e.front = front;
  int rc = srsue_gpu_pdsch_plan_create(b->ctx, &d.cell, &d.cfg, b->chunk_cap, &e.plan);
  if (rc) return rc;
  srsue_gpu_pdsch_plan_info(e.plan, &e.info);
  b->lru.push_front(key);
  e.lru = b->lru.begin();
  auto ins = b->plans.emplace(key, e);
  *out = &ins.first->second;
  return 0;
}

template <typename T>
int grow_pinned(T** p, size_t* have, size_t want, cudaStream_t s) {
  if (want <= *have) return 0;
  if (*p) { B_CU(cudaStreamSynchronize(s)); B_CU(cudaFreeHost(*p)); *p = nullptr; }
  B_CU(cudaMallocHost((void**)p, want * sizeof(T)));
  *have = want;
  return 0;
}

// IQ rows up to this size are gathered on the host when they are scattered in memory (one cudaMemcpyAsync call
// costs a few microseconds, about the CPU time to copy 32-64 KB)
constexpr size_t kGatherRowBytes = 32 * 1024;

template <typename T>
int grow(T** p, size_t* have, size_t want, cudaStream_t s) {
  if (want <= *have) return 0;
  if (*p) { B_CU(cudaStreamSynchronize(s)); B_CU(cudaFree(*p)); *p = nullptr; }
  B_CU(cudaMalloc((void**)p, want * sizeof(T)));
  *have = want;
  return 0;
}

}  // namespace

extern "C" {

int srsue_gpu_batch_create(srsue_gpu_ctx_t* ctx, int max_subframes, float noise_est, int noise_mode, int max_iter,
                           srsue_gpu_batch_t** out) {
  if (!ctx || !out || max_subframes < 1 || max_iter < 1) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_create: bad arguments");
  *out = nullptr;
  auto* b = new srsue_gpu_batch();
  b->ctx = ctx;
  b->max_subframes = max_subframes;
  b->chunk_cap = std::min(max_subframes, 1024);
  b->noise_est = noise_est; b->noise_mode = noise_mode; b->max_iter = max_iter;
  B_CU(cudaStreamCreateWithFlags(&b->s_compute, cudaStreamNonBlocking));
  B_CU(cudaStreamCreateWithFlags(&b->s_copy, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    B_CU(cudaEventCreateWithFlags(&b->ev_up[i], cudaEventDisableTiming));
    B_CU(cudaEventCreateWithFlags(&b->ev_free[i], cudaEventDisableTiming));
  }
  B_CU(cudaMalloc((void**)&b->d_status, (size_t)b->chunk_cap * 4 * sizeof(int32_t)));
  B_CU(cudaMalloc((void**)&b->d_meas, (size_t)b->chunk_cap * 5 * sizeof(float)));
  B_CU(cudaMalloc((void**)&b->d_rows, (size_t)b->chunk_cap * sizeof(int16_t*)));
  B_CU(cudaMalloc((void**)&b->d_cfo, (size_t)b->chunk_cap * sizeof(int32_t)));
  B_CU(cudaMallocHost((void**)&b->h_cfo, (size_t)max_subframes * sizeof(int32_t)));
  B_CU(cudaMallocHost((void**)&b->h_status, (size_t)max_subframes * 4 * sizeof(int32_t)));
  B_CU(cudaMallocHost((void**)&b->h_meas, (size_t)max_subframes * 5 * sizeof(float)));
  B_CU(cudaMallocHost((void**)&b->h_rows, (size_t)max_subframes * sizeof(int16_t*)));
  B_CU(cudaMallocHost((void**)&b->h_iq_rows, (size_t)max_subframes * sizeof(void*)));
  B_CU(cudaMallocHost((void**)&b->h_pl_rows, (size_t)max_subframes * sizeof(void*)));
  *out = b;
  return 0;
}

namespace {

// host thread of one device of a multi-GPU batch: runs its share of every submission through the device's own batch
void multi_worker(srsue_gpu_batch* front, srsue_gpu_batch::Dev* d) {
  cudaSetDevice(d->device);
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(front->mu);
      front->cv.wait(lk, [&] { return front->quit || front->generation != seen; });
      if (front->quit) return;
      seen = front->generation;
    }
    d->rc = 0;
    d->err.clear();
    // every worker picks its own share out of the caller's array (the copies were 0.9 ms of one thread for 32 768 descriptors)
    d->descs.clear(); d->index.clear();
    d->descs.reserve((size_t)d->share); d->index.reserve((size_t)d->share);
    for (int i = 0; i < front->n_src; i++)
      if (front->where[i] == d->id) { d->descs.push_back(front->src[i]); d->index.push_back(i); }
    if (!d->descs.empty()) {
      d->rc = front->multi_blind_run ? srsue_gpu_batch_submit_blind(d->b, d->descs.data(), (int)d->descs.size(), front->multi_ng_x6, front->multi_payload_cap)
                                     : srsue_gpu_batch_submit(d->b, d->descs.data(), (int)d->descs.size());
      if (!d->rc) d->rc = srsue_gpu_batch_wait(d->b);
      if (d->rc) d->err = srsue_gpu_last_error();          // the error text is thread-local: hand it to the caller's thread
    }
    {
      std::lock_guard<std::mutex> lk(front->mu);
      front->running--;
    }
    front->cv.notify_all();
  }
}

double turbo_work(int tbs) {                                // sum of C * K of a transport block: what the decoder's time follows
  int seg[8];
  if (tbs <= 0 || srsue_gpu_host_cbsegm(tbs, seg)) return 1.0;
  return (double)seg[5] * seg[3] + (double)seg[6] * seg[4] + 1.0;
}

int multi_submit(srsue_gpu_batch* f, srsue_gpu_sf_desc_t* descs, int n) {
  const int nd = (int)f->devs.size();
  for (auto* d : f->devs) { d->work = 0; d->share = 0; }
  // HARQ state is device-resident: a soft buffer id stays on the device that first saw it.  Everything else is cut
  // into contiguous runs (neighbouring host buffers keep merging into single copies) balanced by estimated turbo work.
  std::vector<int>& where = f->where;
  where.assign((size_t)n, -1);
  std::vector<double> work((size_t)n);
  double total = 0;
  int memo_tbs = -1;
  double memo_w = 0;
  for (int i = 0; i < n; i++) {
    if (descs[i].cfg.tbs != memo_tbs) { memo_tbs = descs[i].cfg.tbs; memo_w = turbo_work(memo_tbs); }     // runs of one size
    const double w = work[i] = memo_w;
    total += w;
    if (descs[i].softbuffer_id < 0) continue;
    auto it = f->affinity.find(descs[i].softbuffer_id);
    if (it != f->affinity.end()) { where[i] = it->second; f->devs[it->second]->work += w; }
  }
  const double target = total / nd;
  int cur = 0;
  for (int i = 0; i < n; i++) {
    if (where[i] >= 0) continue;
    const double w = work[i];
    while (cur + 1 < nd && f->devs[cur]->work + 0.5 * w > target) cur++;
    where[i] = cur;
    f->devs[cur]->work += w;
    if (descs[i].softbuffer_id >= 0) f->affinity[descs[i].softbuffer_id] = cur;
  }
  for (int i = 0; i < n; i++) f->devs[where[i]]->share++;
  for (auto* d : f->devs)
    if (d->share > d->b->max_subframes) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit: share of device %d exceeds max_subframes", d->device);
  f->src = descs;
  f->n_src = n;
  {
    std::lock_guard<std::mutex> lk(f->mu);
    f->running = nd;
    f->multi_blind_run = f->multi_blind;
    f->generation++;
  }
  f->cv.notify_all();
  f->pending = descs;
  f->n_pending = n;
  return 0;
}

int multi_wait(srsue_gpu_batch* f) {
  if (!f->pending) return 0;
  {
    std::unique_lock<std::mutex> lk(f->mu);
    f->cv.wait(lk, [&] { return f->running == 0; });
  }
  int rc = 0;
  std::string err;
  f->launches = 0;
  for (auto* d : f->devs) {
    if (d->rc && !rc) { rc = d->rc; err = "device " + std::to_string(d->device) + ": " + d->err; }
    for (size_t k = 0; k < d->descs.size(); k++) {
      srsue_gpu_sf_desc_t& o = f->pending[d->index[k]];
      o.crc_ok = d->descs[k].crc_ok;
      o.n_iter = d->descs[k].n_iter;
      if (f->multi_blind_run) o.cfg = d->descs[k].cfg;
      std::memcpy(o.meas, d->descs[k].meas, sizeof(o.meas));
    }
    f->launches += d->b->launches;
  }
  f->pending = nullptr;
  f->n_pending = 0;
  if (rc) return srsue::internal_fail(rc, err.c_str());
  return 0;
}

void multi_destroy(srsue_gpu_batch* f) {
  {
    std::lock_guard<std::mutex> lk(f->mu);
    f->quit = true;
  }
  f->cv.notify_all();
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto* d : f->devs) {
    if (d->th.joinable()) d->th.join();
    cudaSetDevice(d->device);
    if (d->b) srsue_gpu_batch_destroy(d->b);
    if (d->ctx) srsue_gpu_ctx_destroy(d->ctx);
    delete d;
  }
  cudaSetDevice(prev);
  delete f;
}

}  // namespace

int srsue_gpu_batch_create_multi(const int* devices, int n_devices, int max_subframes, float noise_est, int noise_mode, int max_iter,
                                 srsue_gpu_batch_t** out) {
  if (!devices || !out || n_devices < 1 || n_devices > 64 || max_subframes < 1 || max_iter < 1)
    B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_create_multi: bad arguments");
  *out = nullptr;
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_create_multi: device %d listed twice", devices[i]);
  int prev = 0;
  cudaGetDevice(&prev);
  auto* f = new srsue_gpu_batch();
  f->max_subframes = max_subframes;
  int rc = 0;
  for (int i = 0; i < n_devices && !rc; i++) {
    auto* d = new srsue_gpu_batch::Dev();
    d->device = devices[i];
    d->id = (int)f->devs.size();
    f->devs.push_back(d);
    rc = srsue_gpu_ctx_create(&d->ctx, d->device);            // (sets the current device)
    // any device may end up with the whole submission (all of it pinned to one device by HARQ affinity)
    if (!rc) rc = srsue_gpu_batch_create(d->ctx, max_subframes, noise_est, noise_mode, max_iter, &d->b);
  }
  cudaSetDevice(prev);
  if (rc) {
    const std::string err = srsue_gpu_last_error();
    multi_destroy(f);
    return srsue::internal_fail(rc, err.c_str());
  }
  for (auto* d : f->devs) d->th = std::thread(multi_worker, f, d);
  *out = f;
  return 0;
}

void srsue_gpu_batch_destroy(srsue_gpu_batch_t* b) {
  if (!b) return;
  if (!b->devs.empty()) { multi_destroy(b); return; }
  cudaStreamSynchronize(b->s_compute);
  cudaStreamSynchronize(b->s_copy);
  for (auto& kv : b->plans) srsue_gpu_pdsch_plan_destroy(kv.second.plan);
  for (auto& kv : b->softbuffers) cudaFree(kv.second.d);
  cudaFree(b->d_iq[0]); cudaFree(b->d_iq[1]); cudaFree(b->d_payload); cudaFree(b->d_status); cudaFree(b->d_meas);
  cudaFree(b->d_sb); cudaFree(b->d_rows); cudaFree(b->d_cfo); cudaFreeHost(b->h_cfo);
  cudaFreeHost(b->h_status); cudaFreeHost(b->h_meas); cudaFreeHost(b->h_rows);
  cudaFreeHost(b->h_iq[0]); cudaFreeHost(b->h_iq[1]); cudaFreeHost(b->h_pl);
  cudaFreeHost(b->h_iq_rows); cudaFreeHost(b->h_pl_rows);
  cudaFree(b->d_biq); cudaFree(b->d_bsf); cudaFree(b->d_bce); cudaFree(b->d_bllr); cudaFree(b->d_bcfi); cudaFree(b->d_bfound); cudaFree(b->d_bbits);
  cudaFreeHost(b->h_bcfi); cudaFreeHost(b->h_bfound); cudaFreeHost(b->h_bbits);
  for (int i = 0; i < 2; i++) { cudaEventDestroy(b->ev_up[i]); cudaEventDestroy(b->ev_free[i]); }
  for (cudaEvent_t e : b->ev_blind) cudaEventDestroy(e);
  cudaStreamDestroy(b->s_compute); cudaStreamDestroy(b->s_copy);
  if (b->s_aux) cudaStreamDestroy(b->s_aux);
  if (b->s_ctrl) cudaStreamDestroy(b->s_ctrl);
  cudaFree(b->d_bmeas);
  delete b;
}

int srsue_gpu_batch_set_iq_format(srsue_gpu_batch_t* b, int format, float scale) {
  if (!b || (format != SRSUE_GPU_IQ_CF32 && format != SRSUE_GPU_IQ_SC16) || (format == SRSUE_GPU_IQ_SC16 && !(scale > 0.f)))
    return srsue::internal_fail(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_set_iq_format: format 0 (cf32) or 1 (sc16 with a positive scale)");
  for (auto* d : b->devs) srsue_gpu_batch_set_iq_format(d->b, format, scale);
  b->iq_format = format;
  if (format == SRSUE_GPU_IQ_SC16) b->iq16_scale = scale;
  return 0;
}

static int batch_submit_impl(srsue_gpu_batch_t* b, srsue_gpu_sf_desc_t* descs, int n, int first = 0);

int srsue_gpu_batch_submit(srsue_gpu_batch_t* b, srsue_gpu_sf_desc_t* descs, int n) {
  if (!b || !descs || n < 0 || n > b->max_subframes) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit: bad arguments (n=%d)", n);
  if (b->pending) B_FAIL(SRSUE_GPU_ERROR, "batch_submit: the previous submission has not been waited for");
  if (!b->devs.empty()) return multi_submit(b, descs, n);
  b->launches = 0;
  const int rc = batch_submit_impl(b, descs, n);
  if (rc) {
    // Every descriptor is validated before the first launch, so a failure here is a CUDA error in the middle of the
    // submission: copies and scatter kernels of earlier chunks may still be writing into the callers' payload buffers.
    // Let them finish before the caller gets its buffers back; nothing is delivered (batch_wait has nothing pending).
    cudaStreamSynchronize(b->s_copy);
    cudaStreamSynchronize(b->s_compute);
  }
  return rc;
}

// descs[first .. n) join the submission (first > 0: a blind submission feeding its PDSCH phase group by group; the payload
// staging was sized for the whole submission by the caller, nothing that is in flight may move)
static int batch_submit_impl(srsue_gpu_batch_t* b, srsue_gpu_sf_desc_t* descs, int n, int first) {
  if (first == 0) {
    b->order.clear();
    b->pl_chunks.clear();
    b->pl_off = 0;
  }
  size_t& pl_off = b->pl_off;
  size_t pl_total = pl_off;
  for (int i = first; i < n; i++) pl_total += (size_t)(descs[i].cfg.tbs + 7) / 8;
  if (pl_total > b->h_pl_bytes && first) B_FAIL(SRSUE_GPU_ERROR, "batch_submit: payload staging too small for an appended group");
  { int rc = grow_pinned(&b->h_pl, &b->h_pl_bytes, pl_total, b->s_compute); if (rc) return rc; }
  // ---- bucket the descriptors by launch shape, keeping arrival order inside a bucket -----------------------
  std::map<std::string, std::vector<int>> groups;
  std::vector<std::string> group_order;
  std::map<int64_t, int> seen;
  struct Recent { int rep; std::vector<int>* bucket; char mode; };
  Recent recent[8];
  int n_recent = 0;
  for (int i = first; i < n; i++) {
    const srsue_gpu_sf_desc_t& d = descs[i];
    if (!d.iq || !d.payload) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit: descriptor %d has a null buffer", i);
    if (!(d.cfo > -1.0f && d.cfo < 1.0f)) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: cfo %g outside (-1, 1) subcarrier spacings", i, (double)d.cfo);
    if (d.softbuffer_id < 0 && !d.new_data) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: HARQ combining needs a softbuffer_id", i);
    // buckets are launched one after the other, not in arrival order, so one HARQ process may appear only once
    // per submission (its transmissions are 8 ms apart on the air anyway)
    if (d.softbuffer_id >= 0 && !seen.emplace(d.softbuffer_id, i).second)
      B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit: descriptors %d and %d use soft buffer %lld in one submission",
             seen[d.softbuffer_id], i, (long long)d.softbuffer_id);
    // new transmissions and HARQ combines use different launches (reset vs accumulate), as do tracked and
    // untracked soft buffers
    const char mode = d.softbuffer_id >= 0 ? (d.new_data ? 1 : 2) : 0;
    // a descriptor shaped like a recent one joins that bucket without building a key
    // (150 bytes of key + a map lookup per descriptor were 0.2 us each: 8 % of a 20 MHz step before the first copy started)
    // (a mixed stream interleaves a handful of shapes: the last eight distinct ones are remembered, most recent first)
    int hit = -1;
    for (int c = 0; c < n_recent && hit < 0; c++)
      if (mode == recent[c].mode && same_shape(d, descs[recent[c].rep])) hit = c;
    if (hit >= 0) {
      recent[hit].bucket->push_back(i);
      if (hit) std::swap(recent[hit], recent[hit - 1]);          // drifts to the front as it keeps being hit
      continue;
    }
    std::string k = plan_key(d);
    k.push_back(mode);
    auto it = groups.find(k);
    if (it == groups.end()) { group_order.push_back(k); it = groups.emplace(k, std::vector<int>()).first; }
    it->second.push_back(i);
    if (n_recent < 8) n_recent++;
    for (int c = n_recent - 1; c > 0; c--) recent[c] = recent[c - 1];
    recent[0] = Recent{i, &it->second, mode};          // (std::map nodes do not move)
  }
  // ---- everything that can be refused is refused here, before the first launch: the plan of every bucket exists and
  // every combine finds an earlier transmission of the same size ------------------------------------------------------
  for (const std::string& gk : group_order) {
    const std::vector<int>& idx = groups[gk];
    PlanEntry* pe = nullptr;
    const int rc = get_plan(b, gk.substr(0, gk.size() - 1), descs[idx[0]], &pe);   // (looked up again below: with more
    if (rc) return rc;                                                             // buckets than cached plans it may be evicted)
    if (gk.back() != 2) continue;
    for (int i : idx) {
      auto it = b->softbuffers.find(descs[i].softbuffer_id);
      if (it == b->softbuffers.end() || !it->second.d)
        B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: soft buffer %lld has no earlier transmission", i, (long long)descs[i].softbuffer_id);
      if (it->second.elems != pe->info.sb_sf_stride)
        B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: soft buffer %lld holds a different transport block size", i, (long long)descs[i].softbuffer_id);
    }
  }
  // ---- one chain launch per bucket chunk --------------------------------------------------------------------
  for (const std::string& gk : group_order) {
    const std::vector<int>& idx = groups[gk];
    const int mode = gk.back();                       // 0 untracked, 1 tracked new data, 2 tracked combine
    PlanEntry* pe = nullptr;
    int rc = get_plan(b, gk.substr(0, gk.size() - 1), descs[idx[0]], &pe);
    if (rc) return rc;
    const srsue_gpu_plan_info_t& info = pe->info;
    // chunk size: uploads of chunk c + 1 overlap the decoding of chunk c (two staging halves, two streams), so a large
    // bucket is cut into pieces of about 96 MB of samples (400 subframes at 20 MHz) instead of going up in one copy
    // that nothing overlaps; small buckets stay whole (a chunk costs a dozen launches, and a decoder launch with fewer
    // code blocks than slots cannot balance early-stopping blocks against late ones: 24 MB chunks measured 18 % slower
    // on the HARQ workload and 9 % slower on the mixed stream, 250 MB 10 % slower on a single large bucket;
    // SRSUE_BATCH_CHUNK_MB overrides)
    const size_t sample_bytes = (size_t)info.sf_len * (b->iq_format == SRSUE_GPU_IQ_SC16 ? 4 : sizeof(srsue_gpu_cf_t));
    static const size_t chunk_mb = getenv("SRSUE_BATCH_CHUNK_MB") ? (size_t)std::max(1, atoi(getenv("SRSUE_BATCH_CHUNK_MB"))) : 96;
    const size_t cap = std::min<size_t>((size_t)b->chunk_cap, std::max<size_t>(32, (chunk_mb << 20) / sample_bytes));
    for (int i = 0; i < 2; i++) { rc = grow(&b->d_iq[i], &b->iq_elems[i], cap * info.sf_len, b->s_compute); if (rc) return rc; }
    rc = grow(&b->d_payload, &b->payload_bytes, cap * info.payload_stride, b->s_compute);
    if (rc) return rc;
    if (mode) { rc = grow(&b->d_sb, &b->sb_elems, cap * info.sb_sf_stride, b->s_compute); if (rc) return rc; }
    for (size_t off = 0; off < idx.size(); off += cap) {
      const int m = (int)std::min(cap, idx.size() - off);
      const int h = b->half;
      b->half ^= 1;
      // upload: subframes that are adjacent in host memory go up in one copy; scattered small ones are gathered
      // into pinned staging first
      int runs = 1;
      // (all addresses in bytes: a sample is 8 bytes as cf32, 4 as int16 pairs)
      const size_t row_bytes = (size_t)info.sf_len * (b->iq_format == SRSUE_GPU_IQ_SC16 ? 4 : sizeof(srsue_gpu_cf_t));
      auto row_ptr = [&](int r) { return reinterpret_cast<const char*>(descs[idx[off + r]].iq); };
      for (int r = 1; r < m; r++) runs += row_ptr(r) != row_ptr(r - 1) + row_bytes;
      const size_t pos_up = b->order.size();
      bool zero_copy = runs > 4 || b->iq_on_device;
      uintptr_t align_or = 0;
      for (int r = 0; r < m && zero_copy; r++) {
        zero_copy = b->iq_on_device || srsue::host_region_contains(descs[idx[off + r]].iq, row_bytes);
        align_or |= reinterpret_cast<uintptr_t>(descs[idx[off + r]].iq);
      }
      if (zero_copy) {
        for (int r = 0; r < m; r++) b->h_iq_rows[pos_up + r] = descs[idx[off + r]].iq;
        B_CU(cudaStreamWaitEvent(b->s_copy, b->ev_free[h], 0));
        const int slices = (int)std::max<size_t>(1, row_bytes / (32 * 1024));
        if ((align_or & 15) == 0)
          gather_host_rows_kernel<uint4><<<gather_ctas(m * slices), 256, 0, b->s_copy>>>(reinterpret_cast<uint4*>(b->d_iq[h]),
              reinterpret_cast<const uint4* const*>(b->h_iq_rows + pos_up), m, (int)(row_bytes / 16), slices);
        else
          gather_host_rows_kernel<uint2><<<gather_ctas(m * slices), 256, 0, b->s_copy>>>(reinterpret_cast<uint2*>(b->d_iq[h]),
              reinterpret_cast<const uint2* const*>(b->h_iq_rows + pos_up), m, (int)(row_bytes / 8), slices);
        B_CU(cudaGetLastError());
        b->launches++;
      } else if (runs > 4 && row_bytes <= kGatherRowBytes) {
        B_CU(cudaEventSynchronize(b->ev_up[h]));            // the previous upload from this staging half has finished
        rc = grow_pinned(&b->h_iq[h], &b->h_iq_elems[h], cap * info.sf_len, b->s_copy);
        if (rc) return rc;
        for (int r = 0; r < m; r++) std::memcpy(reinterpret_cast<char*>(b->h_iq[h]) + (size_t)r * row_bytes, descs[idx[off + r]].iq, row_bytes);
        B_CU(cudaStreamWaitEvent(b->s_copy, b->ev_free[h], 0));
        B_CU(cudaMemcpyAsync(b->d_iq[h], b->h_iq[h], (size_t)m * row_bytes, cudaMemcpyHostToDevice, b->s_copy));
      } else {
        B_CU(cudaStreamWaitEvent(b->s_copy, b->ev_free[h], 0));
        for (int r = 0; r < m;) {
          int e = r + 1;
          while (e < m && row_ptr(e) == row_ptr(e - 1) + row_bytes) e++;
          B_CU(cudaMemcpyAsync(reinterpret_cast<char*>(b->d_iq[h]) + (size_t)r * row_bytes, descs[idx[off + r]].iq, (size_t)(e - r) * row_bytes,
                               cudaMemcpyHostToDevice, b->s_copy));
          r = e;
        }
      }
      B_CU(cudaEventRecord(b->ev_up[h], b->s_copy));
      B_CU(cudaStreamWaitEvent(b->s_compute, b->ev_up[h], 0));
      const size_t pos0 = b->order.size();
      int16_t* d_sb = nullptr;
      if (mode) {
        // resident soft buffers of the (UE, HARQ process) ids of this chunk
        for (int r = 0; r < m; r++) {
          const srsue_gpu_sf_desc_t& d = descs[idx[off + r]];
          SoftBuffer& sb = b->softbuffers[d.softbuffer_id];
          if (sb.d && sb.elems != info.sb_sf_stride) {
            if (mode == 2) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: soft buffer %lld holds a different transport block size", idx[off + r], (long long)d.softbuffer_id);
            B_CU(cudaStreamSynchronize(b->s_compute));
            B_CU(cudaFree(sb.d)); sb.d = nullptr;
          }
          if (!sb.d) {
            if (mode == 2) {
              b->softbuffers.erase(d.softbuffer_id);
              B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: soft buffer %lld has no earlier transmission", idx[off + r], (long long)d.softbuffer_id);
            }
            B_CU(cudaMalloc((void**)&sb.d, (size_t)info.sb_sf_stride * sizeof(int16_t)));
            sb.elems = info.sb_sf_stride;
          }
          b->h_rows[pos0 + r] = sb.d;
        }
        B_CU(cudaMemcpyAsync(b->d_rows, b->h_rows + pos0, (size_t)m * sizeof(int16_t*), cudaMemcpyHostToDevice, b->s_compute));
        d_sb = b->d_sb;
        if (mode == 2) { softbuffer_rows_kernel<<<m, 256, 0, b->s_compute>>>(d_sb, b->d_rows, info.sb_sf_stride, 0); b->launches++; }
      }
      bool any_cfo = false;
      for (int r = 0; r < m; r++) {
        const float cfo = descs[idx[off + r]].cfo;
        if (!(cfo > -1.0f && cfo < 1.0f)) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "descriptor %d: cfo %g outside (-1, 1) subcarrier spacings", idx[off + r], (double)cfo);
        b->h_cfo[pos0 + r] = cfo != 0.0f ? srsue_gpu_host_cfo_step(cfo, info.nfft) : 0;
        any_cfo |= b->h_cfo[pos0 + r] != 0;
      }
      if (any_cfo) B_CU(cudaMemcpyAsync(b->d_cfo, b->h_cfo + pos0, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, b->s_compute));
      srsue_gpu_pdsch_plan_set_cfo(pe->plan, any_cfo ? b->d_cfo : nullptr, 0);
      srsue_gpu_pdsch_plan_set_iq_format(pe->plan, b->iq_format, b->iq16_scale);
      rc = srsue_gpu_pdsch_decode_batch(pe->plan, m, b->d_iq[h], b->noise_est, b->noise_mode, b->max_iter, mode == 2, d_sb, b->d_payload,
                                        b->d_status, b->d_meas, b->s_compute);
      srsue_gpu_pdsch_plan_set_cfo(pe->plan, nullptr, 0);
      if (rc) return rc;
      b->launches += srsue_gpu_last_launch_count(b->ctx);
      B_CU(cudaEventRecord(b->ev_free[h], b->s_compute));
      if (mode) { softbuffer_rows_kernel<<<m, 256, 0, b->s_compute>>>(d_sb, b->d_rows, info.sb_sf_stride, 1); b->launches++; }
      // results: payload rows to the callers' buffers (adjacent buffers merged), status and measurements to staging
      const size_t pbytes = (size_t)(info.payload_stride);
      int pruns = 1;
      for (int r = 1; r < m; r++) pruns += descs[idx[off + r]].payload != descs[idx[off + r - 1]].payload + pbytes;
      bool pl_zero_copy = pruns > 4;
      uintptr_t pl_align = pbytes;
      for (int r = 0; r < m && pl_zero_copy; r++) {
        pl_zero_copy = srsue::host_region_contains(descs[idx[off + r]].payload, pbytes);
        pl_align |= reinterpret_cast<uintptr_t>(descs[idx[off + r]].payload);
      }
      if (pl_zero_copy) {
        for (int r = 0; r < m; r++) b->h_pl_rows[pos0 + r] = descs[idx[off + r]].payload;
        if ((pl_align & 3) == 0)
          scatter_host_rows_kernel<uint32_t><<<m, 128, 0, b->s_compute>>>(reinterpret_cast<const uint32_t*>(b->d_payload),
              reinterpret_cast<uint32_t* const*>(b->h_pl_rows + pos0), (int)(pbytes / 4));
        else
          scatter_host_rows_kernel<uint8_t><<<m, 128, 0, b->s_compute>>>(b->d_payload, reinterpret_cast<uint8_t* const*>(b->h_pl_rows + pos0), (int)pbytes);
        B_CU(cudaGetLastError());
        b->launches++;
      } else if (pruns > 4) {
        B_CU(cudaMemcpyAsync(b->h_pl + pl_off, b->d_payload, (size_t)m * pbytes, cudaMemcpyDeviceToHost, b->s_compute));
        b->pl_chunks.push_back({pos0, pl_off, pbytes, m});
        pl_off += (size_t)m * pbytes;
      } else {
        for (int r = 0; r < m;) {
          int e = r + 1;
          while (e < m && descs[idx[off + e]].payload == descs[idx[off + e - 1]].payload + pbytes) e++;
          B_CU(cudaMemcpyAsync(descs[idx[off + r]].payload, b->d_payload + (size_t)r * pbytes, (size_t)(e - r) * pbytes,
                               cudaMemcpyDeviceToHost, b->s_compute));
          r = e;
        }
      }
      B_CU(cudaMemcpyAsync(b->h_status + pos0 * 4, b->d_status, (size_t)m * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, b->s_compute));
      B_CU(cudaMemcpyAsync(b->h_meas + pos0 * 5, b->d_meas, (size_t)m * 5 * sizeof(float), cudaMemcpyDeviceToHost, b->s_compute));
      for (int r = 0; r < m; r++) b->order.push_back(idx[off + r]);
    }
  }
  b->pending = descs;
  b->n_pending = n;
  return 0;
}

int srsue_gpu_batch_wait(srsue_gpu_batch_t* b) {
  if (!b) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_wait: null batch");
  if (!b->devs.empty()) return multi_wait(b);
  if (!b->pending) {
    if (b->blind_orig) { b->blind_orig = nullptr; b->blind_n = 0; }       // a blind submission in which no grant was found
    return 0;
  }
  B_CU(cudaStreamSynchronize(b->s_compute));
  for (const auto& c : b->pl_chunks)
    for (int r = 0; r < c.rows; r++)
      std::memcpy(b->pending[b->order[c.pos0 + r]].payload, b->h_pl + c.stage_off + (size_t)r * c.row_bytes, c.row_bytes);
  for (size_t pos = 0; pos < b->order.size(); pos++) {
    srsue_gpu_sf_desc_t& d = b->pending[b->order[pos]];
    d.crc_ok = b->h_status[pos * 4 + 0];
    d.n_iter = b->h_status[pos * 4 + 2];
    std::memcpy(d.meas, b->h_meas + pos * 5, 5 * sizeof(float));
  }
  if (b->blind_orig && b->pending == b->blind_descs.data()) {
    // blind submission: hand the results (and the grants that were found) to the caller's descriptors
    for (size_t k = 0; k < b->blind_index.size(); k++) {
      srsue_gpu_sf_desc_t& o = b->blind_orig[b->blind_index[k]];
      const srsue_gpu_sf_desc_t& r = b->blind_descs[k];
      o.crc_ok = r.crc_ok; o.n_iter = r.n_iter;
      std::memcpy(o.meas, r.meas, sizeof(o.meas));
    }
    b->blind_orig = nullptr;
    b->blind_n = 0;
    b->iq_on_device = false;
  }
  b->pending = nullptr;
  b->n_pending = 0;
  return 0;
}

// ---- blind submission: PCFICH -> CFI, PDCCH search for the descriptor's RNTI, DCI -> grant, then the PDSCH chain --------
// (phch_worker.cc:254-297 for a whole batch).  In: cell, cfg.sf_idx, cfg.rnti, iq, payload (payload_cap bytes).  Out:
// the rest of cfg (cfi, qm, tbs, rv, tm, prb_mask, nof_prb_alloc) and the usual results; a subframe without a DCI for its
// RNTI comes back with cfg.tbs = 0, crc_ok = 0, n_iter = 0.
static int symbol_sz_of(int nof_prb) {          // srslte_symbol_sz: FFT size of a bandwidth
  static const int lim[6] = {6, 15, 25, 50, 75, 110}, sz[6] = {128, 256, 512, 1024, 1536, 2048};
  if (nof_prb <= 0) return -1;
  for (int i = 0; i < 6; i++)
    if (nof_prb <= lim[i]) return sz[i];
  return -1;
}

int srsue_gpu_batch_submit_blind(srsue_gpu_batch_t* b, srsue_gpu_sf_desc_t* descs, int n, int ng_x6, int payload_cap) {
  if (!b || !descs || n < 0 || n > b->max_subframes || payload_cap < 1) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit_blind: bad arguments (n=%d)", n);
  if (ng_x6 != 1 && ng_x6 != 3 && ng_x6 != 6 && ng_x6 != 12) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit_blind: ng_x6 must be 1, 3, 6 or 12");
  if (b->pending || b->blind_orig) B_FAIL(SRSUE_GPU_ERROR, "batch_submit_blind: the previous submission has not been waited for");
  if (!b->devs.empty()) {
    b->multi_blind = 1; b->multi_ng_x6 = ng_x6; b->multi_payload_cap = payload_cap;
    const int rc = multi_submit(b, descs, n);
    b->multi_blind = 0;
    return rc;
  }
  b->launches = 0;
  b->blind_rows_used = 0;
  if (!b->s_ctrl) B_CU(cudaStreamCreateWithFlags(&b->s_ctrl, cudaStreamNonBlocking));
  if (!b->d_bmeas) B_CU(cudaMalloc((void**)&b->d_bmeas, (size_t)b->chunk_cap * 5 * sizeof(float)));
  cudaStream_t st = b->s_ctrl;
  const size_t esz = b->iq_format == SRSUE_GPU_IQ_SC16 ? 4 : sizeof(srsue_gpu_cf_t);
  // ---- 1. control channels run once per (cell, subframe number, RNTI): group the descriptors first ------------------
  for (int i = 0; i < n; i++) {
    const srsue_gpu_sf_desc_t& d = descs[i];
    if (!d.iq || !d.payload) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit_blind: descriptor %d has a null buffer", i);
    if (d.cfg.sf_idx < 0 || d.cfg.sf_idx > 9 || d.softbuffer_id >= 0) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit_blind: descriptor %d: bad sf_idx, or a soft buffer id (blind subframes are decoded as new transmissions)", i);
    if (symbol_sz_of(d.cell.nof_prb) <= 0) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "batch_submit_blind: descriptor %d: bad cell", i);
  }
  std::map<std::string, std::vector<int>> groups;
  std::vector<std::string> group_order;
  for (int i = 0; i < n; i++) {
    const int v[6] = {descs[i].cell.nof_prb, descs[i].cell.nof_ports, descs[i].cell.cell_id, descs[i].cell.cp ? 1 : 0, descs[i].cfg.sf_idx, descs[i].cfg.rnti};
    std::string k(reinterpret_cast<const char*>(v), sizeof(v));
    auto it = groups.find(k);
    if (it == groups.end()) { group_order.push_back(k); it = groups.emplace(k, std::vector<int>()).first; }
    it->second.push_back(i);
  }
  // ---- 2. all samples to the device, once, group by group: the rows of a group are packed one after the other, so the
  // control pass reads them in place.  Scattered rows in pinned memory the library knows are pulled by the GPU (one
  // kernel per group instead of one copy per subframe); neighbouring rows go up in one copy. ---------------------------
  std::vector<size_t> row_off(n + 1, 0);
  {
    size_t pos = 0;
    for (const std::string& gk : group_order) {
      const std::vector<int>& idx = groups[gk];
      const size_t rb = (size_t)15 * symbol_sz_of(descs[idx[0]].cell.nof_prb) * esz;
      pos = (pos + 255) / 256 * 256;
      for (int i : idx) { row_off[i] = pos; pos += rb; }
    }
    int rc = grow(&b->d_biq, &b->biq_bytes, pos + 256, st); if (rc) return rc;
  }
  // The uploads go out on the copy stream, group by group, each followed by an event; the control pass of a group (compute
  // stream, below) waits for its own event only, so it runs next to the uploads of the groups behind it.
  while (b->ev_blind.size() < group_order.size()) {
    cudaEvent_t e;
    B_CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    b->ev_blind.push_back(e);
  }
  size_t g_no = 0;
  for (const std::string& gk : group_order) {
    cudaStream_t st = b->s_copy;                          // (shadows the compute stream inside this loop)
    struct Mark { srsue_gpu_batch* b; size_t g; ~Mark() { cudaEventRecord(b->ev_blind[g], b->s_copy); } } mark{b, g_no++};
    const std::vector<int>& idx = groups[gk];
    const int m = (int)idx.size();
    const size_t rb = (size_t)15 * symbol_sz_of(descs[idx[0]].cell.nof_prb) * esz;
    auto row_ptr = [&](int r) { return reinterpret_cast<const char*>(descs[idx[r]].iq); };
    int runs = 1;
    for (int r = 1; r < m; r++) runs += row_ptr(r) != row_ptr(r - 1) + rb;
    bool pull = runs > 4 && rb % 16 == 0 && m <= b->max_subframes;
    for (int r = 0; r < m && pull; r++) pull = srsue::host_region_contains(descs[idx[r]].iq, rb) && (reinterpret_cast<uintptr_t>(descs[idx[r]].iq) & 15) == 0;
    if (pull) {
      // (the pointer table lives in pinned memory the GPU reads directly; one table region per group)
      const void** tab = b->h_iq_rows + b->blind_rows_used;
      for (int r = 0; r < m; r++) tab[r] = descs[idx[r]].iq;
      b->blind_rows_used += m;
      const int slices = (int)std::max<size_t>(1, rb / (32 * 1024));
      gather_host_rows_kernel<uint4><<<gather_ctas(m * slices), 256, 0, st>>>(reinterpret_cast<uint4*>(b->d_biq + row_off[idx[0]]),
          reinterpret_cast<const uint4* const*>(tab), m, (int)(rb / 16), slices);
      B_CU(cudaGetLastError());
      b->launches++;
    } else {
      for (int r = 0; r < m;) {
        int e = r + 1;
        while (e < m && row_ptr(e) == row_ptr(e - 1) + rb) e++;
        B_CU(cudaMemcpyAsync(b->d_biq + row_off[idx[r]], descs[idx[r]].iq, (size_t)(e - r) * rb, cudaMemcpyHostToDevice, st));
        r = e;
      }
    }
  }
  const size_t cap = (size_t)b->chunk_cap;
  if (!b->d_bcfi) {
    B_CU(cudaMalloc((void**)&b->d_bcfi, cap * sizeof(int32_t)));
    B_CU(cudaMalloc((void**)&b->d_bfound, 9 * cap * 4 * sizeof(int32_t)));
    B_CU(cudaMalloc((void**)&b->d_bbits, 9 * cap * 64));
    B_CU(cudaMallocHost((void**)&b->h_bcfi, cap * sizeof(int32_t)));
    B_CU(cudaMallocHost((void**)&b->h_bfound, 9 * cap * 4 * sizeof(int32_t)));
    B_CU(cudaMallocHost((void**)&b->h_bbits, 9 * cap * 64));
  }
  b->blind_descs.clear();
  b->blind_descs.reserve((size_t)n);            // (the PDSCH phase keeps a pointer into it while later groups are appended)
  b->blind_index.clear();
  // The PDSCH phase of a group starts as soon as its grants are known (step 4 below), while the samples of later groups are still
  // on their way: its device-side gathers use a stream of their own (they would queue up behind every upload on s_copy), and
  // the payload staging is sized now for the whole submission because nothing in flight may move later.
  if (!b->s_aux) B_CU(cudaStreamCreateWithFlags(&b->s_aux, cudaStreamNonBlocking));
  { const int rc = grow_pinned(&b->h_pl, &b->h_pl_bytes, (size_t)n * (size_t)payload_cap, st); if (rc) return rc; }
  b->blind_orig = descs;
  b->blind_n = n;
  size_t fed = 0;
  auto feed = [&]() -> int {                     // hands blind_descs[fed ..) to the PDSCH chain
    if (b->blind_descs.size() == fed) return 0;
    cudaStream_t keep = b->s_copy;
    b->s_copy = b->s_aux;
    b->iq_on_device = true;
    const int rc = batch_submit_impl(b, b->blind_descs.data(), (int)b->blind_descs.size(), (int)fed);
    b->s_copy = keep;
    fed = b->blind_descs.size();
    if (rc) {
      cudaStreamSynchronize(b->s_copy); cudaStreamSynchronize(b->s_aux); cudaStreamSynchronize(b->s_compute);
      b->iq_on_device = false; b->blind_orig = nullptr; b->blind_n = 0; b->pending = nullptr; b->n_pending = 0;
    }
    return rc;
  };
  for (int i = 0; i < n; i++) { descs[i].crc_ok = 0; descs[i].n_iter = 0; descs[i].cfg.tbs = 0; std::memset(descs[i].meas, 0, sizeof(descs[i].meas)); }
  struct Try { int common, fmt, first_bit; };
  g_no = 0;
  for (const std::string& gk : group_order) {
    B_CU(cudaStreamWaitEvent(st, b->ev_blind[g_no++], 0));
    const std::vector<int>& idx = groups[gk];
    const srsue_gpu_sf_desc_t& d0 = descs[idx[0]];
    const int rnti = d0.cfg.rnti, nof_prb = d0.cell.nof_prb;
    const bool user = rnti >= SRSLTE_CRNTI_START && rnti <= SRSLTE_CRNTI_END;
    // the searches of srslte_ue_dl_find_dl_dci_type: C-RNTI -- 1A then 1 in the UE-specific space, 1A in the common space;
    // SI / RA / P-RNTI -- 1A in the common space
    const Try tries_user[3] = {{0, 0, 1}, {0, 1, -1}, {1, 0, 1}}, tries_bc[1] = {{1, 0, 1}};
    const Try* tries = user ? tries_user : tries_bc;
    const int n_tries = user ? 3 : 1;
    PlanEntry* fp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int cfi = 1; cfi <= 3; cfi++) {
      srsue_gpu_sf_desc_t fd = d0;
      std::memset(&fd.cfg, 0, sizeof(fd.cfg));
      fd.cfg.sf_idx = d0.cfg.sf_idx; fd.cfg.cfi = cfi; fd.cfg.qm = 2; fd.cfg.tm = d0.cell.nof_ports == 1 ? 1 : 2; fd.cfg.tbs = 0;
      int rc = get_plan(b, plan_key(fd), fd, &fp[cfi]);
      if (rc) return rc;
    }
    const srsue_gpu_plan_info_t& info = fp[1]->info;
    const size_t grid = (size_t)14 * info.nsc;
    int n_reg_max = 0;
    for (int cfi = 1; cfi <= 3; cfi++) { int nr = 0, nc = 0; if (srsue_gpu_pdcch_info(fp[cfi]->plan, ng_x6, &nr, &nc)) return SRSUE_GPU_ERROR; n_reg_max = std::max(n_reg_max, nr); }
    int rc = grow(&b->d_bsf, &b->bsf_elems, cap * grid, st); if (rc) return rc;
    rc = grow(&b->d_bce, &b->bce_elems, cap * grid * d0.cell.nof_ports, st); if (rc) return rc;
    rc = grow(&b->d_bllr, &b->bllr_elems, 3 * cap * 8 * (size_t)n_reg_max, st); if (rc) return rc;
    for (size_t off = 0; off < idx.size(); off += cap) {
      const int m = (int)std::min(cap, idx.size() - off);
      // the rows of this group lie packed in d_biq in group order: the FFT reads them in place
      const srsue_gpu_cf_t* rows_in = reinterpret_cast<const srsue_gpu_cf_t*>(b->d_biq + row_off[idx[off]]);
      srsue_gpu_pdsch_plan_set_iq_format(fp[1]->plan, b->iq_format, b->iq16_scale);
      srsue_gpu_pdsch_plan_set_cfo(fp[1]->plan, nullptr, 0);
      if (b->iq_format == SRSUE_GPU_IQ_SC16) rc = srsue_gpu_ofdm_rx_sc16(fp[1]->plan, m, reinterpret_cast<const int16_t*>(rows_in), b->iq16_scale, b->d_bsf, st);
      else rc = srsue_gpu_ofdm_rx(fp[1]->plan, m, rows_in, b->d_bsf, st);
      if (!rc) rc = srsue_gpu_chest(fp[1]->plan, m, b->d_bsf, b->d_bce, b->d_bmeas, st);
      // PCFICH and PDCCH with the channel estimator's noise figure, as srslte_ue_dl_decode does
      if (!rc) rc = srsue_gpu_pcfich_decode(fp[1]->plan, m, b->d_bsf, b->d_bce, b->d_bmeas, 0.0f, 1, b->d_bcfi, nullptr, st);
      if (rc) return rc;
      bool have_try[9] = {};
      for (int cfi = 1; cfi <= 3; cfi++) {
        int16_t* llr = b->d_bllr + (size_t)(cfi - 1) * cap * 8 * n_reg_max;
        srsue_gpu_pdsch_plan_set_row_filter(fp[cfi]->plan, b->d_bcfi, cfi);     // only the subframes whose PCFICH said this CFI
        rc = srsue_gpu_pdcch_extract_llr(fp[cfi]->plan, m, b->d_bsf, b->d_bce, b->d_bmeas, 0.0f, 1, ng_x6, llr, st);
        if (rc) return rc;
        for (int t = 0; t < n_tries; t++) {
          const int slot = (cfi - 1) * 3 + t;
          const int nof_bits = srsue_gpu_host_dci_format_sizeof(tries[t].fmt, nof_prb);
          const int r2 = srsue_gpu_pdcch_find_dci(fp[cfi]->plan, m, llr, ng_x6, rnti, tries[t].common, nof_bits, tries[t].first_bit,
                                                  b->d_bfound + (size_t)slot * cap * 4, b->d_bbits + (size_t)slot * cap * 64, nullptr, st);
          have_try[slot] = r2 >= 0;                       // (< 0: a control region too small for a common search space)
          if (t == n_tries - 1) srsue_gpu_pdsch_plan_set_row_filter(fp[cfi]->plan, nullptr, 0);
          if (have_try[slot]) {
            B_CU(cudaMemcpyAsync(b->h_bfound + (size_t)slot * cap * 4, b->d_bfound + (size_t)slot * cap * 4, (size_t)m * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            B_CU(cudaMemcpyAsync(b->h_bbits + (size_t)slot * cap * 64, b->d_bbits + (size_t)slot * cap * 64, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
          }
        }
      }
      B_CU(cudaMemcpyAsync(b->h_bcfi, b->d_bcfi, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      B_CU(cudaStreamSynchronize(st));
      b->launches += 1 + 1 + 1 + 3 * (1 + n_tries);
      // ---- 3. DCI -> grant on the host (ra.cc), descriptor by descriptor ---------------------------------------------
      for (int r = 0; r < m; r++) {
        srsue_gpu_sf_desc_t& d = descs[idx[off + r]];
        const int cfi = b->h_bcfi[r];
        if (cfi < 1 || cfi > 3) continue;
        d.cfg.cfi = cfi;
        for (int t = 0; t < n_tries; t++) {
          const int slot = (cfi - 1) * 3 + t;
          if (!have_try[slot] || !b->h_bfound[((size_t)slot * cap + r) * 4]) continue;
          srslte_dci_msg_t msg;
          std::memset(&msg, 0, sizeof(msg));
          const int nof_bits = srsue_gpu_host_dci_format_sizeof(tries[t].fmt, nof_prb);
          std::memcpy(msg.data, b->h_bbits + ((size_t)slot * cap + r) * 64, (size_t)nof_bits);
          msg.nof_bits = (uint32_t)nof_bits;
          msg.format = tries[t].fmt ? SRSLTE_DCI_FORMAT1 : SRSLTE_DCI_FORMAT1A;
          srslte_ra_dl_dci_t dci;
          srslte_ra_dl_grant_t grant;
          if (srslte_dci_msg_to_dl_grant(&msg, (uint16_t)rnti, (uint32_t)nof_prb, &dci, &grant) || grant.mcs.tbs <= 0) break;
          if ((grant.mcs.tbs + 7) / 8 > payload_cap) break;                    // would not fit the caller's buffer
          d.cfg.qm = (int)grant.Qm; d.cfg.tbs = grant.mcs.tbs; d.cfg.rv = (int)dci.rv_idx;
          d.cfg.tm = d.cell.nof_ports == 1 ? 1 : 2;
          int na = 0;
          for (int p = 0; p < nof_prb; p++) {
            const bool s0 = grant.prb_idx[0][p], s1 = grant.prb_idx[1][p];
            d.cfg.prb_mask[p] = (s0 && s1) ? 1 : (uint8_t)((s0 ? 2 : 0) | (s1 ? 4 : 0));
            na += s0 ? 1 : 0;
          }
          d.cfg.nof_prb_alloc = na;
          break;
        }
        if (d.cfg.tbs > 0) {
          srsue_gpu_sf_desc_t p2 = d;
          p2.iq = reinterpret_cast<const srsue_gpu_cf_t*>(b->d_biq + row_off[idx[off + r]]);
          p2.softbuffer_id = -1; p2.new_data = 1; p2.cfo = 0.f;
          b->blind_descs.push_back(p2);
          b->blind_index.push_back(idx[off + r]);
        }
      }
    }
    // ---- 4. the PDSCH chain for the subframes of this group that have a grant, bucketed by grant as usual ----------------
    { const int rc4 = feed(); if (rc4) return rc4; }
  }
  return 0;
}

int srsue_gpu_batch_softbuffer_release(srsue_gpu_batch_t* b, int64_t softbuffer_id) {
  if (!b) B_FAIL(SRSUE_GPU_ERROR_INVALID_INPUTS, "softbuffer_release: null batch");
  if (!b->devs.empty()) {
    if (b->pending) B_FAIL(SRSUE_GPU_ERROR, "softbuffer_release: a submission is in flight");
    auto a = b->affinity.find(softbuffer_id);
    if (a == b->affinity.end()) return 0;
    auto* d = b->devs[a->second];
    b->affinity.erase(a);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d->device);
    const int rc = srsue_gpu_batch_softbuffer_release(d->b, softbuffer_id);
    cudaSetDevice(prev);
    return rc;
  }
  auto it = b->softbuffers.find(softbuffer_id);
  if (it == b->softbuffers.end()) return 0;
  B_CU(cudaStreamSynchronize(b->s_compute));
  cudaFree(it->second.d);
  b->softbuffers.erase(it);
  return 0;
}

int srsue_gpu_batch_stats(const srsue_gpu_batch_t* b, int* n_plans, int* n_softbuffers, int* launches) {
  if (!b) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  if (!b->devs.empty()) {
    int np = 0, ns = 0;
    for (auto* d : b->devs) { np += (int)d->b->plans.size(); ns += (int)d->b->softbuffers.size(); }
    if (n_plans) *n_plans = np;
    if (n_softbuffers) *n_softbuffers = ns;
    if (launches) *launches = b->launches;
    return 0;
  }
  if (n_plans) *n_plans = (int)b->plans.size();
  if (n_softbuffers) *n_softbuffers = (int)b->softbuffers.size();
  if (launches) *launches = b->launches;
  return 0;
}

int srsue_gpu_batch_device_shares(const srsue_gpu_batch_t* b, int* n_devices, int* subframes, double* work, int cap) {
  if (!b) return SRSUE_GPU_ERROR_INVALID_INPUTS;
  const int nd = (int)b->devs.size();
  if (n_devices) *n_devices = nd;
  for (int i = 0; i < nd && i < cap; i++) {
    if (subframes) subframes[i] = b->devs[i]->share;
    if (work) work[i] = b->devs[i]->work;
  }
  return 0;
}

}  // extern "C"

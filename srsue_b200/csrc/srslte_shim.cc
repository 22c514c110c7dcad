// srslte_shim.cc -- srsLTE-shaped C entry points (include/srsue_gpu/srslte_compat.h) over the batch API
// with a batch of one subframe.  Mirrors the call sequence of phch_worker::work_imp
// (/root/reference/ue/src/phy/phch_worker.cc:132-243): decode_fft_estimate -> cfg_grant ->
// pdsch_decode_rnti, with MAC owning the soft buffer and the payload pointer
// (/root/reference/ue/src/mac/dl_harq.cc:216-279).  All buffers crossing this boundary are caller-owned
// HOST memory; device state is private.  No CPU fallback: without a usable GPU every call fails.
#define SRSUE_GPU_PLAIN_CF 1
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "lte_tables.h"
#include "srsue_gpu/srslte_compat.h"

using namespace srsue;

namespace {

std::mutex g_mu;
srsue_gpu_ctx_t* g_ctx = nullptr;

srsue_gpu_ctx_t* shared_ctx() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx) {
    const char* dev = getenv("SRSUE_GPU_DEVICE");
    if (srsue_gpu_ctx_create(&g_ctx, dev ? atoi(dev) : 0) != 0) {
      fprintf(stderr, "libsrsue_gpu: %s\n", srsue_gpu_last_error());
      g_ctx = nullptr;
    }
  }
  return g_ctx;
}

struct SbShadow {
  int16_t* d_buf = nullptr;
  size_t elems = 0;
  int valid_tbs = 0;      // 0: empty (reset); otherwise the TBS whose LLRs are accumulated
};

struct UeDlGpu {
  srsue_gpu_ctx_t* ctx = nullptr;
  srsue_gpu_cell_t cell{};
  int nsc = 0, sf_len = 0;
  uint32_t cfi = 0;       // 0: take the CFI from the PCFICH of every subframe; 1..3: forced by srsue_gpu_ue_dl_set_cfi
  int32_t* d_cfi = nullptr;
  srsue_gpu_pdsch_plan_t* front[10][4] = {{nullptr}};      // front-end plans per (sf_idx, cfi): FFT, estimate, PCFICH, PDCCH
  int16_t* d_pdcch_llr = nullptr; size_t pdcch_llr_elems = 0;
  int32_t* d_dci_found = nullptr; uint8_t* d_dci_bits = nullptr;
  int pdcch_sf = -1, pdcch_cfi = 0, ng_x6 = 6;
  std::map<std::string, srsue_gpu_pdsch_plan_t*> plans;
  srsue_gpu_cf_t *d_iq = nullptr, *d_sf = nullptr, *d_ce = nullptr;
  float* d_meas = nullptr;
  uint8_t* d_payload = nullptr;
  int32_t* d_tb_status = nullptr;
  bool dev_valid = false;           // d_sf/d_ce hold what q->sf_symbols/q->ce mirror
  const void* h_sf = nullptr;       // addresses of those host mirrors
  const void* h_ce[SRSLTE_MAX_PORTS] = {nullptr};
  bool have_grant = false;
  int32_t cfo_step = 0;             // carrier-offset correction applied while the samples are transformed (0: none)
  uint32_t grant_cfi = 1, grant_rv = 0;
  srslte_ra_dl_grant_t grant{};
  cudaStream_t stream = nullptr;
};

// srsue_gpu_pdsch_cfg_t::prb_mask value of one PRB from the grant's two slot masks
uint8_t slot_mask(bool s0, bool s1) { return (s0 && s1) ? 1 : (uint8_t)((s0 ? 2 : 0) | (s1 ? 4 : 0)); }

// one received block: measure, and every `period` blocks move the radio's gain towards the target power
void agc_process(srslte_agc_t* a, void* handler, const srsue_gpu_cf_t* x, size_t n) {
  if (!a || !a->set_gain_callback || !x || n == 0) return;
  if (!a->handler) a->handler = handler;
  if (a->count++ < a->period) return;
  a->count = 0;
  double acc = 0.0;
  size_t m = 0;
  for (size_t i = 0; i < n; i += 4, m++) acc += (double)x[i].re * x[i].re + (double)x[i].im * x[i].im;    // every 4th sample
  const double p = acc / (double)m;
  a->last_power = (float)p;
  if (!(p > 0.0)) return;
  double g = a->gain + a->bandwidth * 10.0 * std::log10((double)a->target / p);
  g = std::min((double)a->max_gain, std::max(0.0, g));
  a->gain = a->set_gain_callback(a->handler, g);
  a->nof_updates++;
}

int ng_x6_of(const srslte_cell_t& c) {
  switch (c.phich_resources) { case SRSLTE_PHICH_R_1_6: return 1; case SRSLTE_PHICH_R_1_2: return 3; case SRSLTE_PHICH_R_1: return 6; default: return 12; }
}

// front-end-only plan (no grant) of (sf_idx, cfi)
srsue_gpu_pdsch_plan_t* front_plan(UeDlGpu* u, uint32_t sf_idx, uint32_t cfi) {
  if (sf_idx > 9 || cfi < 1 || cfi > 3) return nullptr;
  if (!u->front[sf_idx][cfi]) {
    srsue_gpu_pdsch_cfg_t c{};
    c.sf_idx = (int)sf_idx; c.cfi = (int)cfi; c.qm = 2; c.tm = u->cell.nof_ports == 1 ? 1 : 2; c.tbs = 0;
    if (srsue_gpu_pdsch_plan_create(u->ctx, &u->cell, &c, 1, &u->front[sf_idx][cfi]) != 0) return nullptr;
  }
  return u->front[sf_idx][cfi];
}

int to_gpu_cfg(const srslte_cell_t& cell, const srslte_pdsch_cfg_t* cfg, uint16_t rnti, srsue_gpu_pdsch_cfg_t* out) {
  std::memset(out, 0, sizeof(*out));
  out->sf_idx = (int)cfg->sf_idx;
  out->cfi = (int)cfg->nbits.lstart - (cell.nof_prb <= 10 ? 1 : 0);
  out->rnti = rnti;
  out->qm = (int)cfg->grant.Qm;
  out->tbs = cfg->grant.mcs.tbs;
  out->rv = (int)cfg->rv;
  out->tm = cell.nof_ports == 1 ? 1 : 2;      // srsLTE: single antenna port or transmit diversity
  int n = 0;
  for (uint32_t i = 0; i < cell.nof_prb; i++) {
    out->prb_mask[i] = slot_mask(cfg->grant.prb_idx[0][i], cfg->grant.prb_idx[1][i]);
    n += cfg->grant.prb_idx[0][i] ? 1 : 0;
  }
  out->nof_prb_alloc = n;
  return 0;
}

srsue_gpu_pdsch_plan_t* get_plan(UeDlGpu* u, const srsue_gpu_pdsch_cfg_t& c) {
  std::string key(reinterpret_cast<const char*>(&c), sizeof(c));
  auto it = u->plans.find(key);
  if (it != u->plans.end()) return it->second;
  if (u->plans.size() >= 64) {               // bounded cache: grants repeat in practice
    for (auto& kv : u->plans) srsue_gpu_pdsch_plan_destroy(kv.second);
    u->plans.clear();
  }
  srsue_gpu_pdsch_plan_t* p = nullptr;
  if (srsue_gpu_pdsch_plan_create(u->ctx, &u->cell, &c, 1, &p) != 0) return nullptr;
  u->plans[key] = p;
  return p;
}

int sb_total_elems(uint32_t max_cb) { return (int)max_cb * (3 * kMaxK + 16); }

}  // namespace

extern "C" {

int srslte_symbol_sz(uint32_t nof_prb) { return symbol_sz((int)nof_prb); }

void* srslte_vec_malloc(uint32_t size) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, size ? size : 64)) return nullptr;
  return p;
}
void srslte_vec_free(void* p) { free(p); }

// ---- UE DL ------------------------------------------------------------------------------------------
int srslte_ue_dl_init(srslte_ue_dl_t* q, srslte_cell_t cell) {
  if (!q) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(q, 0, sizeof(*q));
  if (symbol_sz((int)cell.nof_prb) < 0 || (cell.nof_ports != 1 && cell.nof_ports != 2 && cell.nof_ports != 4) ||
      (cell.cp != SRSLTE_CP_NORM && cell.cp != SRSLTE_CP_EXT))
    return SRSLTE_ERROR_INVALID_INPUTS;
  srsue_gpu_ctx_t* ctx = shared_ctx();
  if (!ctx) return SRSLTE_ERROR;
  auto* u = new UeDlGpu();
  u->ctx = ctx;
  u->cell = srsue_gpu_cell_t{(int)cell.nof_prb, (int)cell.nof_ports, (int)cell.id, cell.cp == SRSLTE_CP_EXT ? 1 : 0};
  u->ng_x6 = ng_x6_of(cell);
  u->nsc = 12 * (int)cell.nof_prb;
  u->sf_len = 15 * symbol_sz((int)cell.nof_prb);
  const size_t grid = (size_t)14 * u->nsc;
  bool ok = cudaMalloc((void**)&u->d_iq, u->sf_len * sizeof(srsue_gpu_cf_t)) == cudaSuccess &&
            cudaMalloc((void**)&u->d_sf, grid * sizeof(srsue_gpu_cf_t)) == cudaSuccess &&
            cudaMalloc((void**)&u->d_ce, grid * cell.nof_ports * sizeof(srsue_gpu_cf_t)) == cudaSuccess &&
            cudaMalloc((void**)&u->d_meas, 5 * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&u->d_payload, 19200) == cudaSuccess &&     /* demux.h:60: 150*1024/8 bytes */
            cudaMalloc((void**)&u->d_tb_status, 4 * sizeof(int32_t)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&u->stream, cudaStreamNonBlocking) == cudaSuccess;
  q->gpu = u;
  q->cell = cell;
  q->pdsch.cell = cell; q->pdsch.gpu = u; q->pdsch.dl_sch.max_iterations = 4;   /* ue.conf.example:83 */
  q->chest.cell = cell; q->chest.gpu = u;
  q->pdcch.gpu = u;
  q->sf_symbols = (cf_t*)srslte_vec_malloc((uint32_t)(grid * sizeof(cf_t)));
  for (uint32_t p = 0; p < cell.nof_ports; p++) q->ce[p] = (cf_t*)srslte_vec_malloc((uint32_t)(grid * sizeof(cf_t)));
  u->h_sf = q->sf_symbols;
  for (uint32_t p = 0; p < cell.nof_ports; p++) u->h_ce[p] = q->ce[p];
  ok = ok && q->sf_symbols && q->ce[0] && srslte_softbuffer_rx_init(&q->softbuffer, cell.nof_prb) == 0;
  if (!ok) { srslte_ue_dl_free(q); return SRSLTE_ERROR; }
  return SRSLTE_SUCCESS;
}

void srslte_ue_dl_free(srslte_ue_dl_t* q) {
  if (!q) return;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  if (u) {
    for (auto& fs : u->front) for (auto& f : fs) srsue_gpu_pdsch_plan_destroy(f);
    cudaFree(u->d_pdcch_llr); cudaFree(u->d_dci_found); cudaFree(u->d_dci_bits);
    for (auto& kv : u->plans) srsue_gpu_pdsch_plan_destroy(kv.second);
    cudaFree(u->d_iq); cudaFree(u->d_sf); cudaFree(u->d_ce); cudaFree(u->d_meas); cudaFree(u->d_cfi); cudaFree(u->d_payload); cudaFree(u->d_tb_status);
    if (u->stream) cudaStreamDestroy(u->stream);
    delete u;
  }
  srslte_vec_free(q->sf_symbols);
  for (int p = 0; p < SRSLTE_MAX_PORTS; p++) srslte_vec_free(q->ce[p]);
  srslte_softbuffer_rx_free(&q->softbuffer);
  std::memset(q, 0, sizeof(*q));
}

void srslte_ue_dl_set_rnti(srslte_ue_dl_t* q, uint16_t rnti) {
  // called from the MAC thread while workers run (phy.cc:253-257): a plain field write; the scrambling
  // sequence of the rnti is generated lazily when a plan for it is first built
  if (q) { q->current_rnti = rnti; q->pdsch.rnti = rnti; }
}

void srsue_gpu_ue_dl_set_cfi(srslte_ue_dl_t* q, uint32_t cfi) {
  if (q && q->gpu && cfi <= 3) static_cast<UeDlGpu*>(q->gpu)->cfi = cfi;     // 0: decode the PCFICH (default)
}

// The downlink half of phch_worker::set_cfo (phch_worker.cc:120, fed from srslte_ue_sync_get_cfo at phch_recv.cc:328-329):
// srsLTE's synchroniser rotates the samples on the CPU before the worker sees them; here the rotation rides on the
// sample loads of the FFT.  cfo in subcarrier spacings (what set_cfo receives); 0 switches the correction off.
int srsue_gpu_ue_dl_set_cfo(srslte_ue_dl_t* q, float cfo) {
  if (!q || !q->gpu || !(cfo > -1.0f && cfo < 1.0f)) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  u->cfo_step = srsue_gpu_host_cfo_step(cfo, srslte_symbol_sz(q->cell.nof_prb));
  return SRSLTE_SUCCESS;
}

int srslte_ue_dl_decode_fft_estimate(srslte_ue_dl_t* q, cf_t* input, uint32_t sf_idx, uint32_t* cfi) {
  if (!q || !q->gpu || !input || sf_idx > 9) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  srsue_gpu_pdsch_plan_t* fp = front_plan(u, sf_idx, 1);
  if (!fp) return SRSLTE_ERROR;
  const size_t grid = (size_t)14 * u->nsc;
  const size_t glen = (size_t)nof_symb(u->cell.cp) * u->nsc;   // what the caller's buffers hold (12 symbols with the extended prefix)
  // the IQ buffer is reused by the caller right after we return (phch_worker.cc:254 vs :559,610,641):
  // it is fully consumed (H2D) before the synchronise below
  if (cudaMemcpyAsync(u->d_iq, input, u->sf_len * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, u->stream) != cudaSuccess) return SRSLTE_ERROR;
  if (srsue_gpu_ofdm_rx_cfo(fp, 1, u->d_iq, u->d_sf, nullptr, u->cfo_step, u->stream)) return SRSLTE_ERROR;
  if (srsue_gpu_chest(fp, 1, u->d_sf, u->d_ce, u->d_meas, u->stream)) return SRSLTE_ERROR;
  float meas[5];
  int32_t cfi_dec = 0;
  if (u->cfi == 0) {
    // srsLTE decodes the PCFICH with the channel estimator's noise figure
    if (!u->d_cfi && cudaMalloc((void**)&u->d_cfi, 4 * sizeof(int32_t)) != cudaSuccess) return SRSLTE_ERROR;
    if (srsue_gpu_pcfich_decode(fp, 1, u->d_sf, u->d_ce, u->d_meas, 0.0f, 1, u->d_cfi, u->d_cfi + 1, u->stream)) return SRSLTE_ERROR;
    cudaMemcpyAsync(&cfi_dec, u->d_cfi, sizeof(cfi_dec), cudaMemcpyDeviceToHost, u->stream);
  }
  cudaMemcpyAsync(q->sf_symbols, u->d_sf, glen * sizeof(srsue_gpu_cf_t), cudaMemcpyDeviceToHost, u->stream);
  for (int p = 0; p < u->cell.nof_ports; p++)
    cudaMemcpyAsync(q->ce[p], u->d_ce + (size_t)p * grid, glen * sizeof(srsue_gpu_cf_t), cudaMemcpyDeviceToHost, u->stream);
  cudaMemcpyAsync(meas, u->d_meas, sizeof(meas), cudaMemcpyDeviceToHost, u->stream);
  if (cudaStreamSynchronize(u->stream) != cudaSuccess) return SRSLTE_ERROR;
  q->chest.noise_estimate = meas[0]; q->chest.rsrp = meas[1]; q->chest.rssi = meas[2]; q->chest.rsrq = meas[3];
  u->dev_valid = true;
  if (cfi) *cfi = u->cfi ? u->cfi : (uint32_t)cfi_dec;
  return SRSLTE_SUCCESS;
}

int srslte_ue_dl_cfg_grant(srslte_ue_dl_t* q, srslte_ra_dl_grant_t* grant, uint32_t cfi, uint32_t sf_idx, uint32_t rvidx) {
  if (!q || !grant || cfi < 1 || cfi > 3 || sf_idx > 9 || rvidx > 3) return SRSLTE_ERROR_INVALID_INPUTS;
  srslte_pdsch_cfg_t* c = &q->pdsch_cfg;
  std::memset(c, 0, sizeof(*c));
  c->grant = *grant;
  c->rv = rvidx;
  c->sf_idx = sf_idx;
  c->nbits.lstart = cfi + (q->cell.nof_prb <= 10 ? 1 : 0);
  c->nbits.nof_symb = (q->cell.cp == SRSLTE_CP_EXT ? 12 : 14) - c->nbits.lstart;
  CellCfg cell{(int)q->cell.nof_prb, (int)q->cell.nof_ports, (int)q->cell.id, q->cell.cp == SRSLTE_CP_EXT ? 1 : 0};
  PdschCfg pc{};
  pc.sf_idx = (int)sf_idx; pc.cfi = (int)cfi;
  for (uint32_t i = 0; i < q->cell.nof_prb; i++) pc.prb_mask[i] = slot_mask(grant->prb_idx[0][i], grant->prb_idx[1][i]);
  std::vector<int32_t> re;
  pdsch_re_list(cell, pc, re);
  c->nbits.nof_re = (uint32_t)re.size();
  c->nbits.nof_bits = c->nbits.nof_re * grant->Qm;
  if (grant->mcs.tbs > 0) {
    CbSegm s;
    if (!cbsegm(grant->mcs.tbs, &s)) return SRSLTE_ERROR;
    c->cb_segm.F = s.F; c->cb_segm.C = s.C; c->cb_segm.K1 = s.Kp; c->cb_segm.K2 = s.Km;
    c->cb_segm.C1 = s.Cp; c->cb_segm.C2 = s.Cm; c->cb_segm.tbs = s.tbs;
  }
  return SRSLTE_SUCCESS;
}

int srsue_gpu_ue_dl_set_grant(srslte_ue_dl_t* q, const srslte_ra_dl_grant_t* grant, uint32_t cfi, uint32_t rvidx) {
  if (!q || !q->gpu || !grant) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  u->grant = *grant; u->grant_cfi = cfi; u->grant_rv = rvidx; u->have_grant = true; u->cfi = cfi;
  return SRSLTE_SUCCESS;
}

int srslte_pdsch_decode_rnti(srslte_pdsch_t* q, srslte_pdsch_cfg_t* cfg, srslte_softbuffer_rx_t* softbuffer, cf_t* sf_symbols,
                             cf_t* ce[SRSLTE_MAX_PORTS], float noise_estimate, uint16_t rnti, uint8_t* data) {
  if (!q || !q->gpu || !cfg || !softbuffer || !sf_symbols || !ce || !data) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  auto* sh = static_cast<SbShadow*>(softbuffer->gpu_shadow);
  if (!sh || cfg->grant.mcs.tbs <= 0) return SRSLTE_ERROR_INVALID_INPUTS;
  srsue_gpu_pdsch_cfg_t c;
  if (to_gpu_cfg(q->cell, cfg, rnti, &c)) return SRSLTE_ERROR_INVALID_INPUTS;
  srsue_gpu_pdsch_plan_t* p = get_plan(u, c);
  if (!p) return SRSLTE_ERROR;
  srsue_gpu_plan_info_t info;
  srsue_gpu_pdsch_plan_info(p, &info);
  if ((size_t)info.sb_sf_stride > sh->elems || info.payload_stride > 19200) return SRSLTE_ERROR;
  const size_t grid = (size_t)14 * u->nsc;
  const size_t glen = (size_t)nof_symb(u->cell.cp) * u->nsc;   // what the caller's buffers hold (12 symbols with the extended prefix)
  // reuse the device copies of this subframe's grid when the caller hands back our own mirrors
  // (phch_worker.cc:347-348 passes ue_dl.sf_symbols / ue_dl.ce), otherwise upload what we were given
  bool own = u->dev_valid && sf_symbols == u->h_sf;
  for (int pt = 0; pt < u->cell.nof_ports; pt++) own = own && (ce[pt] == u->h_ce[pt]);
  if (!own) {
    u->dev_valid = false;
    cudaMemcpyAsync(u->d_sf, sf_symbols, glen * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, u->stream);
    for (int pt = 0; pt < u->cell.nof_ports; pt++)
      cudaMemcpyAsync(u->d_ce + (size_t)pt * grid, ce[pt], glen * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, u->stream);
  }
  const int accumulate = (sh->valid_tbs == c.tbs) ? 1 : 0;
  if (srsue_gpu_pdsch_llr(p, 1, u->d_sf, u->d_ce, u->d_meas, noise_estimate, 0, accumulate, sh->d_buf, nullptr, nullptr, u->stream))
    return SRSLTE_ERROR;
  sh->valid_tbs = c.tbs;
  const int max_it = q->dl_sch.max_iterations ? (int)q->dl_sch.max_iterations : 4;
  if (srsue_gpu_pdsch_turbo(p, 1, sh->d_buf, max_it, u->d_payload, u->d_tb_status, nullptr, u->stream)) return SRSLTE_ERROR;
  int32_t st[4];
  // payload must be complete before we return: MAC pushes it to the PDU queue next (dl_harq.cc:296-309)
  cudaMemcpyAsync(data, u->d_payload, info.payload_stride, cudaMemcpyDeviceToHost, u->stream);
  cudaMemcpyAsync(st, u->d_tb_status, sizeof(st), cudaMemcpyDeviceToHost, u->stream);
  if (cudaStreamSynchronize(u->stream) != cudaSuccess) return SRSLTE_ERROR;
  q->dl_sch.nof_iterations = (uint32_t)st[2];
  return st[0] ? SRSLTE_SUCCESS : SRSLTE_ERROR;
}

// ---- PDCCH ------------------------------------------------------------------------------------------------
int srslte_pdcch_extract_llr(srslte_pdcch_t* q, cf_t* sf_symbols, cf_t* ce[SRSLTE_MAX_PORTS], float noise_estimate, uint32_t nsubframe,
                             uint32_t cfi) {
  if (!q || !q->gpu || !sf_symbols || !ce || nsubframe > 9 || cfi < 1 || cfi > 3) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  srsue_gpu_pdsch_plan_t* fp = front_plan(u, nsubframe, cfi);
  if (!fp) return SRSLTE_ERROR;
  const size_t grid = (size_t)14 * u->nsc;
  const size_t glen = (size_t)nof_symb(u->cell.cp) * u->nsc;   // what the caller's buffers hold (12 symbols with the extended prefix)
  bool own = u->dev_valid && sf_symbols == u->h_sf;
  for (int pt = 0; pt < u->cell.nof_ports; pt++) own = own && (ce[pt] == u->h_ce[pt]);
  if (!own) {
    u->dev_valid = false;
    cudaMemcpyAsync(u->d_sf, sf_symbols, glen * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, u->stream);
    for (int pt = 0; pt < u->cell.nof_ports; pt++)
      cudaMemcpyAsync(u->d_ce + (size_t)pt * grid, ce[pt], glen * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, u->stream);
  }
  int n_reg = 0, n_cce = 0;
  if (srsue_gpu_pdcch_info(fp, u->ng_x6, &n_reg, &n_cce)) return SRSLTE_ERROR;
  if ((size_t)8 * n_reg > u->pdcch_llr_elems) {
    cudaFree(u->d_pdcch_llr);
    if (cudaMalloc((void**)&u->d_pdcch_llr, (size_t)8 * n_reg * sizeof(int16_t)) != cudaSuccess) return SRSLTE_ERROR;
    u->pdcch_llr_elems = (size_t)8 * n_reg;
  }
  u->pdcch_sf = -1;
  if (srsue_gpu_pdcch_extract_llr(fp, 1, u->d_sf, u->d_ce, u->d_meas, noise_estimate, 0, u->ng_x6, u->d_pdcch_llr, u->stream)) return SRSLTE_ERROR;
  u->pdcch_sf = (int)nsubframe; u->pdcch_cfi = (int)cfi;
  return SRSLTE_SUCCESS;
}

namespace {
// one blind search; returns 1 found / 0 not found / < 0 error and fills dci_msg + the location fields of q
int dci_search(srslte_ue_dl_t* q, UeDlGpu* u, srsue_gpu_pdsch_plan_t* fp, srslte_dci_msg_t* dci_msg, uint16_t rnti, int common,
               srslte_dci_format_t fmt, int first_bit) {
  const int nof_bits = dci_format_sizeof(fmt == SRSLTE_DCI_FORMAT1 ? 1 : 0, u->cell.nof_prb);     // format 0 is padded to 1A's size
  const int rc = srsue_gpu_pdcch_find_dci(fp, 1, u->d_pdcch_llr, u->ng_x6, rnti, common, nof_bits, first_bit, u->d_dci_found,
                                          u->d_dci_bits, nullptr, u->stream);
  if (rc < 0) return common ? 0 : SRSLTE_ERROR;          // a control region smaller than 4 CCEs has no common space
  int32_t found[4];
  uint8_t bits[64];
  cudaMemcpyAsync(found, u->d_dci_found, sizeof(found), cudaMemcpyDeviceToHost, u->stream);
  cudaMemcpyAsync(bits, u->d_dci_bits, sizeof(bits), cudaMemcpyDeviceToHost, u->stream);
  if (cudaStreamSynchronize(u->stream) != cudaSuccess) return SRSLTE_ERROR;
  if (!found[0]) return 0;
  std::memset(dci_msg, 0, sizeof(*dci_msg));
  std::memcpy(dci_msg->data, bits, (size_t)nof_bits);
  dci_msg->nof_bits = (uint32_t)nof_bits;
  dci_msg->format = fmt;
  q->last_location.L = (uint32_t)found[1]; q->last_location.ncce = (uint32_t)found[2];
  q->last_n_cce = (uint32_t)found[2];
  q->nof_detected++;
  return 1;
}
}  // namespace

// uplink grants (phch_worker.cc:426): DCI format 0 in the UE-specific space, same size as 1A, flag bit 0
int srslte_ue_dl_find_ul_dci(srslte_ue_dl_t* q, srslte_dci_msg_t* dci_msg, uint32_t cfi, uint32_t sf_idx, uint16_t rnti) {
  if (!q || !q->gpu || !dci_msg || sf_idx > 9 || cfi < 1 || cfi > 3) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  if (u->pdcch_sf != (int)sf_idx || u->pdcch_cfi != (int)cfi) return SRSLTE_ERROR;
  srsue_gpu_pdsch_plan_t* fp = front_plan(u, sf_idx, cfi);
  if (!fp) return SRSLTE_ERROR;
  if (!u->d_dci_found && (cudaMalloc((void**)&u->d_dci_found, 4 * sizeof(int32_t)) != cudaSuccess ||
                          cudaMalloc((void**)&u->d_dci_bits, 64) != cudaSuccess)) return SRSLTE_ERROR;
  return dci_search(q, u, fp, dci_msg, rnti, 0, SRSLTE_DCI_FORMAT0, 0);
}

int srslte_ue_dl_find_dl_dci_type(srslte_ue_dl_t* q, srslte_dci_msg_t* dci_msg, uint32_t cfi, uint32_t sf_idx, uint16_t rnti,
                                  srslte_rnti_type_t rnti_type) {
  if (!q || !q->gpu || !dci_msg || sf_idx > 9 || cfi < 1 || cfi > 3) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  if (u->pdcch_sf != (int)sf_idx || u->pdcch_cfi != (int)cfi) return SRSLTE_ERROR;      // srslte_pdcch_extract_llr comes first
  srsue_gpu_pdsch_plan_t* fp = front_plan(u, sf_idx, cfi);
  if (!fp) return SRSLTE_ERROR;
  if (!u->d_dci_found && (cudaMalloc((void**)&u->d_dci_found, 4 * sizeof(int32_t)) != cudaSuccess ||
                          cudaMalloc((void**)&u->d_dci_bits, 64) != cudaSuccess)) return SRSLTE_ERROR;
  struct Try { int common; srslte_dci_format_t fmt; };
  const Try user[3] = {{0, SRSLTE_DCI_FORMAT1A}, {0, SRSLTE_DCI_FORMAT1}, {1, SRSLTE_DCI_FORMAT1A}};
  const Try bcast[1] = {{1, SRSLTE_DCI_FORMAT1A}};
  const bool is_user = (rnti_type == SRSLTE_RNTI_USER || rnti_type == SRSLTE_RNTI_TEMP || rnti_type == SRSLTE_RNTI_SPS);
  const Try* tries = is_user ? user : bcast;
  const int n_tries = is_user ? 3 : 1;
  for (int i = 0; i < n_tries; i++) {
    const int rc = dci_search(q, u, fp, dci_msg, rnti, tries[i].common, tries[i].fmt, tries[i].fmt == SRSLTE_DCI_FORMAT1A ? 1 : -1);
    if (rc) return rc;
  }
  return 0;
}

uint32_t srslte_ue_dl_get_ncce(srslte_ue_dl_t* q) { return q ? q->last_n_cce : 0; }

bool srslte_ue_dl_decode_phich(srslte_ue_dl_t* q, uint32_t sf_idx, uint32_t n_prb_lowest, uint32_t n_dmrs) {
  if (!q || !q->gpu || sf_idx > 9) return false;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  if (!u->dev_valid) return false;                   // the grid of this subframe must be on the device (decode_fft_estimate)
  srsue_gpu_pdsch_plan_t* fp = front_plan(u, sf_idx, 1);
  if (!fp) return false;
  int n_group = 0, n_seq = 0;
  phich_index(u->cell.nof_prb, u->ng_x6, (int)n_prb_lowest, (int)n_dmrs, &n_group, &n_seq, u->cell.cp);
  if (!u->d_cfi && cudaMalloc((void**)&u->d_cfi, 4 * sizeof(int32_t)) != cudaSuccess) return false;
  // srsLTE decodes the PHICH with the channel estimator's noise figure
  if (srsue_gpu_phich_decode(fp, 1, u->d_sf, u->d_ce, u->d_meas, 0.0f, 1, u->ng_x6, n_group, n_seq, u->d_cfi + 2, nullptr, u->stream)) return false;
  int32_t ack = 0;
  cudaMemcpyAsync(&ack, u->d_cfi + 2, sizeof(ack), cudaMemcpyDeviceToHost, u->stream);
  if (cudaStreamSynchronize(u->stream) != cudaSuccess) return false;
  return ack != 0;
}

// ---- cell search ---------------------------------------------------------------------------------------------
namespace {
struct CellSearchGpu {
  srsue_gpu_ctx_t* ctx = nullptr;
  int (*recv)(void*, void*, uint32_t, srslte_timestamp_t*) = nullptr;
  void* handler = nullptr;
  srsue_gpu_cf_t* h_iq = nullptr;     // pinned staging for the scanned frames
  srsue_gpu_cf_t* d_iq = nullptr;
  srsue_gpu_sync_result_t* d_res = nullptr;
  uint32_t cap_frames = 0;
  cudaStream_t stream = nullptr;
};
constexpr int kHalfFrame = 9600;       // 5 ms at 1.92 Msps

int cellsearch_run(srslte_ue_cellsearch_t* q, int force, srslte_ue_cellsearch_result_t* out, uint32_t* n_id_2_out) {
  auto* g = static_cast<CellSearchGpu*>(q->gpu);
  const uint32_t nf = q->nof_frames_to_scan ? q->nof_frames_to_scan : 8;
  if (nf > g->cap_frames) {
    if (g->h_iq) cudaFreeHost(g->h_iq);
    cudaFree(g->d_iq); cudaFree(g->d_res);
    g->h_iq = nullptr; g->d_iq = nullptr; g->d_res = nullptr; g->cap_frames = 0;
    if (cudaMallocHost((void**)&g->h_iq, (size_t)nf * kHalfFrame * sizeof(srsue_gpu_cf_t)) != cudaSuccess ||
        cudaMalloc((void**)&g->d_iq, (size_t)nf * kHalfFrame * sizeof(srsue_gpu_cf_t)) != cudaSuccess ||
        cudaMalloc((void**)&g->d_res, (size_t)nf * sizeof(srsue_gpu_sync_result_t)) != cudaSuccess) return SRSLTE_ERROR;
    g->cap_frames = nf;
  }
  srslte_timestamp_t ts;
  for (uint32_t f = 0; f < nf; f++)
  {
    if (g->recv(g->handler, g->h_iq + (size_t)f * kHalfFrame, kHalfFrame, &ts) < 0) return SRSLTE_ERROR;
    agc_process(&q->ue_sync.agc, g->handler, g->h_iq + (size_t)f * kHalfFrame, kHalfFrame);
  }
  if (cudaMemcpyAsync(g->d_iq, g->h_iq, (size_t)nf * kHalfFrame * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, g->stream) != cudaSuccess) return SRSLTE_ERROR;
  // the SSS is looked for behind both cyclic-prefix lengths; every frame reports the better one (SPEC.md 15b.7)
  if (srsue_gpu_cell_search_cp(g->ctx, g->d_iq, (int)nf, kHalfFrame, kHalfFrame, 128, force, 0, 2, g->d_res, g->stream)) return SRSLTE_ERROR;
  std::vector<srsue_gpu_sync_result_t> res(nf);
  cudaMemcpyAsync(res.data(), g->d_res, nf * sizeof(srsue_gpu_sync_result_t), cudaMemcpyDeviceToHost, g->stream);
  if (cudaStreamSynchronize(g->stream) != cudaSuccess) return SRSLTE_ERROR;
  // detections: frames whose peak-to-side ratio reaches the threshold and whose SSS lies inside the frame; the cell
  // most of them agree on wins (srsLTE takes the mode of the per-frame candidates as well)
  std::map<int, int> votes;
  for (const auto& r : res)
    if (r.valid && r.mean_power > 0.f && r.peak / r.mean_power >= q->detect_threshold) votes[3 * r.n_id_1 + r.n_id_2]++;
  int best_id = -1, best_votes = 0;
  for (const auto& kv : votes) if (kv.second > best_votes) { best_votes = kv.second; best_id = kv.first; }
  if (best_id < 0) return 0;
  double psr = 0, cfo = 0;
  int ext_votes = 0;
  for (const auto& r : res)
    if (r.valid && 3 * r.n_id_1 + r.n_id_2 == best_id && r.peak / r.mean_power >= q->detect_threshold) { psr += r.peak / r.mean_power; cfo += r.cfo; ext_votes += r.cp ? 1 : 0; }
  std::memset(out, 0, sizeof(*out));
  out->cell_id = (uint32_t)best_id;
  out->cp = 2 * ext_votes > best_votes ? SRSLTE_CP_EXT : SRSLTE_CP_NORM;      // the prefix most detections of this cell agree on
  out->peak = out->psr = (float)(psr / best_votes);
  out->mode = (float)best_votes / (float)nf;
  out->cfo = (float)(cfo / best_votes * 15000.0);
  if (n_id_2_out) *n_id_2_out = (uint32_t)(best_id % 3);
  return best_votes;
}
}  // namespace

int srslte_ue_cellsearch_init(srslte_ue_cellsearch_t* q, int (*recv_callback)(void*, void*, uint32_t, srslte_timestamp_t*), void* stream_handler) {
  if (!q || !recv_callback) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(q, 0, sizeof(*q));
  srsue_gpu_ctx_t* ctx = shared_ctx();
  if (!ctx) return SRSLTE_ERROR;
  auto* g = new CellSearchGpu();
  g->ctx = ctx; g->recv = recv_callback; g->handler = stream_handler;
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { delete g; return SRSLTE_ERROR; }
  q->gpu = g;
  q->nof_frames_to_scan = 8;
  q->detect_threshold = 10.0f;
  return SRSLTE_SUCCESS;
}

void srslte_ue_cellsearch_free(srslte_ue_cellsearch_t* q) {
  if (!q || !q->gpu) return;
  auto* g = static_cast<CellSearchGpu*>(q->gpu);
  if (g->h_iq) cudaFreeHost(g->h_iq);
  cudaFree(g->d_iq); cudaFree(g->d_res);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  q->gpu = nullptr;
}

int srslte_ue_cellsearch_set_nof_frames_to_scan(srslte_ue_cellsearch_t* q, uint32_t nof_frames) {
  if (!q || nof_frames == 0 || nof_frames > 1024) return SRSLTE_ERROR_INVALID_INPUTS;
  q->nof_frames_to_scan = nof_frames;
  return SRSLTE_SUCCESS;
}
void srslte_ue_cellsearch_set_threshold(srslte_ue_cellsearch_t* q, float threshold) { if (q) q->detect_threshold = threshold; }

int srslte_ue_cellsearch_scan(srslte_ue_cellsearch_t* q, srslte_ue_cellsearch_result_t found_cells[3], uint32_t* max_N_id_2) {
  if (!q || !q->gpu || !found_cells) return SRSLTE_ERROR_INVALID_INPUTS;
  srslte_ue_cellsearch_result_t r;
  uint32_t n2 = 0;
  const int rc = cellsearch_run(q, -1, &r, &n2);
  if (rc > 0) { found_cells[n2] = r; if (max_N_id_2) *max_N_id_2 = n2; }
  return rc;
}

int srslte_ue_cellsearch_scan_N_id_2(srslte_ue_cellsearch_t* q, uint32_t N_id_2, srslte_ue_cellsearch_result_t* found_cell) {
  if (!q || !q->gpu || !found_cell || N_id_2 > 2) return SRSLTE_ERROR_INVALID_INPUTS;
  return cellsearch_run(q, (int)N_id_2, found_cell, nullptr);
}

// ---- automatic gain control (host logic; srsLTE runs srslte_agc_process on every block its synchroniser receives) --------
int srslte_ue_sync_start_agc(srslte_ue_sync_t* q, double (*set_gain_callback)(void*, double), float init_gain_value) {
  if (!q) return SRSLTE_ERROR_INVALID_INPUTS;
  srslte_agc_t* a = &q->agc;
  a->gain = init_gain_value;
  a->set_gain_callback = set_gain_callback;
  a->target = 0.1f;            // mean |x|^2: 10 dB of headroom below full scale for the OFDM peaks
  a->bandwidth = 0.5f;
  a->max_gain = 90.f;
  a->count = 0; a->nof_updates = 0; a->last_power = 0.f;
  return SRSLTE_SUCCESS;
}
float srslte_agc_get_gain(srslte_agc_t* q) { return q ? (float)q->gain : 0.f; }

// ---- subframe synchroniser -------------------------------------------------------------------------------------
namespace {
struct UeSyncGpu {
  srsue_gpu_ctx_t* ctx = nullptr;
  int (*recv)(void*, void*, uint32_t, srslte_timestamp_t*) = nullptr;
  void* handler = nullptr;
  uint32_t cell_id = 0;
  int cp = 0;                          // the cell's cyclic prefix: where the SSS lies in front of the PSS
  int nfft = 0, sf_len = 0, off_pss = 0, win = 0;
  bool tracking = false, sss_on_track = true;
  uint32_t sf_idx = 0;
  int pending_shift = 0;               // samples to skip (> 0) or to re-use (< 0) before the next subframe
  int lost = 0;
  float cfo_mean = 0.f;                // subcarrier spacings
  double drift_samples = 0.0, tracked_sf = 0.0;
  std::vector<srsue_gpu_cf_t> window;  // two half frames while searching
  std::vector<srsue_gpu_cf_t> tail;    // the end of the previous subframe (for negative timing corrections)
  std::vector<srsue_gpu_cf_t> own;     // buffer of srslte_ue_sync_get_buffer
  srsue_gpu_cf_t* d_iq = nullptr;
  srsue_gpu_sync_result_t* d_res = nullptr;
  cudaStream_t stream = nullptr;
  srslte_timestamp_t last_ts{};
};

// one cell search on host samples: upload, search root n_id_2 from first_pos on, read the result back
int sync_search(UeSyncGpu* g, const srsue_gpu_cf_t* x, int n_samples, int first_pos, srsue_gpu_sync_result_t* r) {
  if (cudaMemcpyAsync(g->d_iq, x, (size_t)n_samples * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, g->stream) != cudaSuccess) return SRSLTE_ERROR;
  if (srsue_gpu_cell_search_cp(g->ctx, g->d_iq, 1, n_samples, n_samples, g->nfft, (int)(g->cell_id % 3), first_pos, g->cp, g->d_res, g->stream)) return SRSLTE_ERROR;
  cudaMemcpyAsync(r, g->d_res, sizeof(*r), cudaMemcpyDeviceToHost, g->stream);
  return cudaStreamSynchronize(g->stream) == cudaSuccess ? SRSLTE_SUCCESS : SRSLTE_ERROR;
}

int sync_find(srslte_ue_sync_t* q, UeSyncGpu* g) {
  const int H = 5 * g->sf_len;
  // slide the two-half-frame window by one half frame
  std::memmove(g->window.data(), g->window.data() + H, (size_t)H * sizeof(srsue_gpu_cf_t));
  if (g->recv(g->handler, g->window.data() + H, (uint32_t)H, &g->last_ts) < 0) return SRSLTE_ERROR;
  agc_process(&q->agc, g->handler, g->window.data() + H, (size_t)H);
  srsue_gpu_sync_result_t r;
  if (sync_search(g, g->window.data(), g->off_pss + H + g->nfft - 1, g->off_pss, &r)) return SRSLTE_ERROR;
  const float thr = q->strack.threshold > 0.f ? q->strack.threshold : 10.f;
  if (!r.valid || r.mean_power <= 0.f || r.peak / r.mean_power < thr || 3 * r.n_id_1 + r.n_id_2 != (int)g->cell_id) return 0;
  // the subframe that holds this PSS starts at s; align the stream to the next subframe boundary after the window
  const int s = r.peak_pos - g->off_pss;
  const int have = 2 * H - s;                                   // samples of the stream from s to the end of the window
  const int k = (have + g->sf_len - 1) / g->sf_len;             // subframes begun by then
  const int skip = k * g->sf_len - have;
  if (skip > 0) {
    std::vector<srsue_gpu_cf_t> junk((size_t)skip);
    if (g->recv(g->handler, junk.data(), (uint32_t)skip, &g->last_ts) < 0) return SRSLTE_ERROR;
  }
  g->sf_idx = (uint32_t)(((r.sf5 ? 5 : 0) + k - 1) % 10);        // index of the last subframe already past
  g->tracking = true;
  g->lost = 0;
  g->pending_shift = 0;
  g->cfo_mean = r.cfo;
  std::fill(g->tail.begin(), g->tail.end(), srsue_gpu_cf_t{0.f, 0.f});
  return 0;
}

int sync_track(srslte_ue_sync_t* q, UeSyncGpu* g, srsue_gpu_cf_t* buf) {
  const int L = g->sf_len, T = (int)g->tail.size();
  int shift = g->pending_shift;
  g->pending_shift = 0;
  if (shift > 0) {                                              // we are early: drop samples
    std::vector<srsue_gpu_cf_t> junk((size_t)shift);
    if (g->recv(g->handler, junk.data(), (uint32_t)shift, &g->last_ts) < 0) return SRSLTE_ERROR;
    shift = 0;
  }
  const int reuse = std::min(-shift, T);                         // we are late: the subframe began in the previous read
  if (reuse > 0) std::memcpy(buf, g->tail.data() + (T - reuse), (size_t)reuse * sizeof(srsue_gpu_cf_t));
  if (g->recv(g->handler, buf + reuse, (uint32_t)(L - reuse), &g->last_ts) < 0) return SRSLTE_ERROR;
  agc_process(&q->agc, g->handler, buf, (size_t)L);
  std::memcpy(g->tail.data(), buf + (L - T), (size_t)T * sizeof(srsue_gpu_cf_t));
  g->sf_idx = (g->sf_idx + 1) % 10;
  g->tracked_sf += 1.0;
  if (g->sf_idx == 0 || g->sf_idx == 5) {
    srsue_gpu_sync_result_t r;
    if (sync_search(g, buf, g->off_pss + g->win + g->nfft, g->off_pss - g->win, &r)) return SRSLTE_ERROR;
    const float thr = q->strack.threshold > 0.f ? q->strack.threshold : 10.f;
    const bool good = r.valid && r.mean_power > 0.f && r.peak / r.mean_power >= thr &&
                      (!g->sss_on_track || (3 * r.n_id_1 + r.n_id_2 == (int)g->cell_id && (r.sf5 != 0) == (g->sf_idx == 5)));
    if (good) {
      g->lost = 0;
      g->pending_shift = r.peak_pos - g->off_pss;
      g->drift_samples += g->pending_shift;
      const float alpha = q->strack.em_alpha > 0.f ? q->strack.em_alpha : 0.1f;
      g->cfo_mean = (1.f - alpha) * g->cfo_mean + alpha * r.cfo;
    } else if (++g->lost >= 10) {
      g->tracking = false;                                        // back to searching; the window restarts empty
      std::fill(g->window.begin(), g->window.end(), srsue_gpu_cf_t{0.f, 0.f});
    }
  }
  return 1;
}
}  // namespace

int srslte_ue_sync_init(srslte_ue_sync_t* q, srslte_cell_t cell, int (*recv_callback)(void*, void*, uint32_t, srslte_timestamp_t*),
                        void* stream_handler) {
  if (!q || !recv_callback || cell.id > 503 || (cell.cp != SRSLTE_CP_NORM && cell.cp != SRSLTE_CP_EXT)) return SRSLTE_ERROR_INVALID_INPUTS;
  const int n = symbol_sz((int)cell.nof_prb);
  if (n < 0 || n == 1536) return SRSLTE_ERROR_INVALID_INPUTS;      // the synchroniser needs a power-of-two symbol size
  std::memset(q, 0, sizeof(*q));
  srsue_gpu_ctx_t* ctx = shared_ctx();
  if (!ctx) return SRSLTE_ERROR;
  auto* g = new UeSyncGpu();
  g->ctx = ctx; g->recv = recv_callback; g->handler = stream_handler; g->cell_id = cell.id; g->cp = cell.cp == SRSLTE_CP_EXT ? 1 : 0;
  g->nfft = n; g->sf_len = 15 * n; g->off_pss = 832 * n / 128; g->win = n / 8;
  g->window.assign((size_t)10 * g->sf_len, srsue_gpu_cf_t{0.f, 0.f});
  g->tail.assign((size_t)g->win, srsue_gpu_cf_t{0.f, 0.f});
  g->own.resize((size_t)g->sf_len);
  const size_t max_search = (size_t)g->off_pss + 5 * g->sf_len + n;
  if (cudaMalloc((void**)&g->d_iq, max_search * sizeof(srsue_gpu_cf_t)) != cudaSuccess ||
      cudaMalloc((void**)&g->d_res, sizeof(srsue_gpu_sync_result_t)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaFree(g->d_iq); cudaFree(g->d_res);
    delete g;
    return SRSLTE_ERROR;
  }
  q->strack.threshold = 10.f; q->strack.em_alpha = 0.1f;
  q->gpu = g;
  return SRSLTE_SUCCESS;
}

void srslte_ue_sync_free(srslte_ue_sync_t* q) {
  if (!q || !q->gpu) return;
  auto* g = static_cast<UeSyncGpu*>(q->gpu);
  cudaFree(g->d_iq); cudaFree(g->d_res);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  q->gpu = nullptr;
}

int srslte_ue_sync_zerocopy(srslte_ue_sync_t* q, cf_t* input_buffer) {
  if (!q || !q->gpu || !input_buffer) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* g = static_cast<UeSyncGpu*>(q->gpu);
  return g->tracking ? sync_track(q, g, reinterpret_cast<srsue_gpu_cf_t*>(input_buffer)) : sync_find(q, g);
}

int srslte_ue_sync_get_buffer(srslte_ue_sync_t* q, cf_t** sf_symbols) {
  if (!q || !q->gpu || !sf_symbols) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* g = static_cast<UeSyncGpu*>(q->gpu);
  *sf_symbols = reinterpret_cast<cf_t*>(g->own.data());
  return srslte_ue_sync_zerocopy(q, *sf_symbols);
}

uint32_t srslte_ue_sync_get_sfidx(srslte_ue_sync_t* q) { return (q && q->gpu) ? static_cast<UeSyncGpu*>(q->gpu)->sf_idx : 0; }
float srslte_ue_sync_get_cfo(srslte_ue_sync_t* q) { return (q && q->gpu) ? 15000.f * static_cast<UeSyncGpu*>(q->gpu)->cfo_mean : 0.f; }
float srslte_ue_sync_get_sfo(srslte_ue_sync_t* q) {
  if (!q || !q->gpu) return 0.f;
  auto* g = static_cast<UeSyncGpu*>(q->gpu);
  // accumulated timing corrections per elapsed time: samples per second of sampling-clock offset
  return g->tracked_sf > 0.0 ? (float)(g->drift_samples / (g->tracked_sf * 1e-3)) : 0.f;
}
void srslte_ue_sync_set_cfo(srslte_ue_sync_t* q, float cfo) { if (q && q->gpu) static_cast<UeSyncGpu*>(q->gpu)->cfo_mean = cfo / 15000.f; }
void srslte_ue_sync_set_agc_period(srslte_ue_sync_t* q, uint32_t period) { if (q) q->agc.period = period; }
void srslte_ue_sync_decode_sss_on_track(srslte_ue_sync_t* q, bool enabled) { if (q && q->gpu) static_cast<UeSyncGpu*>(q->gpu)->sss_on_track = enabled; }
void srslte_ue_sync_get_last_timestamp(srslte_ue_sync_t* q, srslte_timestamp_t* timestamp) {
  if (q && q->gpu && timestamp) *timestamp = static_cast<UeSyncGpu*>(q->gpu)->last_ts;
}
void srslte_sync_set_threshold(srslte_sync_t* q, float threshold) { if (q) q->threshold = threshold; }
void srslte_sync_set_em_alpha(srslte_sync_t* q, float alpha) { if (q) q->em_alpha = alpha; }

// ---- MIB search over the air interface ----------------------------------------------------------------------
namespace {
struct MibSyncGpu {
  srslte_ue_cellsearch_t cs;       // the frame puller + batched cell search
  srslte_ue_mib_t mib;             // 6-PRB front end + PBCH decoder
  int (*recv)(void*, void*, uint32_t, srslte_timestamp_t*) = nullptr;
  void* handler = nullptr;
  std::vector<srsue_gpu_cf_t> win; // two consecutive 5 ms frames
  int cp = 0;
};
}  // namespace

int srslte_ue_mib_sync_init(srslte_ue_mib_sync_t* q, uint32_t cell_id, srslte_cp_t cp,
                            int (*recv_callback)(void*, void*, uint32_t, srslte_timestamp_t*), void* stream_handler) {
  if (!q || !recv_callback || cell_id > 503 || (cp != SRSLTE_CP_NORM && cp != SRSLTE_CP_EXT)) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(q, 0, sizeof(*q));
  auto* m = new MibSyncGpu();
  srslte_cell_t c{};
  c.nof_prb = 6; c.nof_ports = 1; c.id = cell_id; c.cp = cp;        // the PBCH sits in the six central PRBs: 1.92 Msps suffice
  if (srslte_ue_cellsearch_init(&m->cs, recv_callback, stream_handler) != SRSLTE_SUCCESS || srslte_ue_mib_init(&m->mib, c) != SRSLTE_SUCCESS) {
    srslte_ue_cellsearch_free(&m->cs);
    srslte_ue_mib_free(&m->mib);
    delete m;
    return SRSLTE_ERROR;
  }
  m->recv = recv_callback; m->handler = stream_handler; m->cp = cp == SRSLTE_CP_EXT ? 1 : 0;
  m->win.resize(2 * kHalfFrame);
  q->cell_id = cell_id;
  q->gpu = m;
  return SRSLTE_SUCCESS;
}

void srslte_ue_mib_sync_free(srslte_ue_mib_sync_t* q) {
  if (!q || !q->gpu) return;
  auto* m = static_cast<MibSyncGpu*>(q->gpu);
  srslte_ue_cellsearch_free(&m->cs);
  srslte_ue_mib_free(&m->mib);
  delete m;
  q->gpu = nullptr;
}

int srslte_ue_mib_sync_decode(srslte_ue_mib_sync_t* q, uint32_t max_frames_timeout, uint8_t bch_payload[SRSLTE_BCH_PAYLOAD_LEN],
                              uint32_t* nof_tx_ports, uint32_t* sfn_offset) {
  if (!q || !q->gpu || !bch_payload) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* m = static_cast<MibSyncGpu*>(q->gpu);
  auto* g = static_cast<CellSearchGpu*>(m->cs.gpu);
  if (!g->d_res) {          // scratch for one-frame searches
    if (cudaMalloc((void**)&g->d_iq, (size_t)2 * kHalfFrame * sizeof(srsue_gpu_cf_t)) != cudaSuccess ||
        cudaMalloc((void**)&g->d_res, sizeof(srsue_gpu_sync_result_t)) != cudaSuccess) return SRSLTE_ERROR;
    g->cap_frames = 0;      // not usable by the scan path: it reallocates when it needs more
  }
  srslte_timestamp_t ts;
  // sliding window of two 5 ms frames: the PSS of the older one is searched, and a subframe 0 that starts in it lies
  // completely inside the window
  if (m->recv(m->handler, m->win.data() + kHalfFrame, kHalfFrame, &ts) < 0) return SRSLTE_ERROR;
  agc_process(&q->ue_sync.agc, m->handler, m->win.data() + kHalfFrame, kHalfFrame);
  for (uint32_t f = 0; f < max_frames_timeout; f++) {
    std::memcpy(m->win.data(), m->win.data() + kHalfFrame, kHalfFrame * sizeof(srsue_gpu_cf_t));
    if (m->recv(m->handler, m->win.data() + kHalfFrame, kHalfFrame, &ts) < 0) return SRSLTE_ERROR;
    agc_process(&q->ue_sync.agc, m->handler, m->win.data() + kHalfFrame, kHalfFrame);
    if (cudaMemcpyAsync(g->d_iq, m->win.data(), (size_t)2 * kHalfFrame * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, g->stream) != cudaSuccess) return SRSLTE_ERROR;
    // exactly one PSS period, starting at offset 832: the subframe that contains a candidate then begins inside the window
    if (srsue_gpu_cell_search_cp(g->ctx, g->d_iq, 1, 832 + kHalfFrame + 127, 2 * kHalfFrame, 128, (int)(q->cell_id % 3), 832, m->cp, g->d_res, g->stream)) return SRSLTE_ERROR;
    srsue_gpu_sync_result_t r;
    cudaMemcpyAsync(&r, g->d_res, sizeof(r), cudaMemcpyDeviceToHost, g->stream);
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return SRSLTE_ERROR;
    if (!r.valid || 3 * r.n_id_1 + r.n_id_2 != (int)q->cell_id || r.sf5) continue;      // not our cell, or the PSS of subframe 5
    if (r.mean_power <= 0.f || r.peak / r.mean_power < 10.0f) continue;
    // subframe 0 starts 832 samples before the PSS symbol body, the last 128 samples of the 960-sample slot with either
    // prefix (normal: CP 10 + 128, five more symbols of 9 + 128, CP 9)
    const int start = r.peak_pos - 832;
    if (start < 0 || start + 1920 > 2 * kHalfFrame) continue;
    const int rc = srslte_ue_mib_decode(&m->mib, reinterpret_cast<cf_t*>(m->win.data() + start), bch_payload, nof_tx_ports, sfn_offset);
    if (rc == SRSLTE_UE_MIB_FOUND) return 1;
    if (rc < 0) return rc;
  }
  return 0;
}

// ---- small utilities ----------------------------------------------------------------------------------------
void srslte_bit_pack_vector(uint8_t* unpacked, uint8_t* packed, int nof_bits) {
  for (int i = 0; i < (nof_bits + 7) / 8; i++) packed[i] = 0;
  for (int i = 0; i < nof_bits; i++) packed[i >> 3] |= (uint8_t)((unpacked[i] & 1) << (7 - (i & 7)));
}
void srslte_bit_unpack_vector(uint8_t* packed, uint8_t* unpacked, int nof_bits) {
  for (int i = 0; i < nof_bits; i++) unpacked[i] = (packed[i >> 3] >> (7 - (i & 7))) & 1;
}
uint32_t srslte_bit_pack(uint8_t** bits, int nof_bits) {
  uint32_t v = 0;
  for (int i = 0; i < nof_bits; i++) v = (v << 1) | ((*bits)[i] & 1);
  *bits += nof_bits;
  return v;
}
void srslte_bit_unpack(uint32_t value, uint8_t** bits, int nof_bits) {
  for (int i = 0; i < nof_bits; i++) (*bits)[i] = (uint8_t)((value >> (nof_bits - 1 - i)) & 1);
  *bits += nof_bits;
}
const char* srslte_cp_string(srslte_cp_t cp) { return cp == SRSLTE_CP_NORM ? "Normal  " : "Extended"; }
void srslte_cell_fprint(FILE* stream, srslte_cell_t* cell, uint32_t sfn) {
  if (!stream || !cell) return;
  static const char* ng[4] = {"1/6", "1/2", "1", "2"};
  fprintf(stream, " - Cell ID:         %d\n - Nof ports:       %d\n - CP:              %s\n - PRB:             %d\n"
                  " - PHICH Length:    %s\n - PHICH Resources: %s\n - SFN:             %d\n",
          cell->id, cell->nof_ports, srslte_cp_string(cell->cp), cell->nof_prb, cell->phich_length == SRSLTE_PHICH_EXT ? "Extended" : "Normal",
          ng[cell->phich_resources & 3], sfn);
}
int srslte_sampling_freq_hz(uint32_t nof_prb) {
  const int n = symbol_sz((int)nof_prb);
  return n > 0 ? 15000 * n : SRSLTE_ERROR;
}
void srslte_timestamp_copy(srslte_timestamp_t* dest, srslte_timestamp_t* src) { if (dest && src) *dest = *src; }
int srslte_timestamp_add(srslte_timestamp_t* t, uint32_t full_secs, double frac_secs) {
  if (!t || frac_secs < 0.0 || frac_secs >= 1.0) return SRSLTE_ERROR;
  t->full_secs += full_secs;
  t->frac_secs += frac_secs;
  if (t->frac_secs >= 1.0) { t->frac_secs -= 1.0; t->full_secs++; }
  return SRSLTE_SUCCESS;
}
uint32_t srslte_tti_interval(uint32_t tti1, uint32_t tti2) { return tti1 >= tti2 ? tti1 - tti2 : 10240 - tti2 + tti1; }

// ---- MIB ---------------------------------------------------------------------------------------------------
namespace {
struct MibGpu {
  srslte_ue_dl_t ue_dl;          // front end (FFT + estimator) for two ports, so that both port hypotheses can be tried
  int32_t* d_result = nullptr;
  uint8_t* d_mib = nullptr;
};
}  // namespace

int srslte_ue_mib_init(srslte_ue_mib_t* q, srslte_cell_t cell) {
  if (!q) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(q, 0, sizeof(*q));
  auto* m = new MibGpu();
  srslte_cell_t c2 = cell;
  c2.nof_ports = 4;            // the estimator runs for all four ports; the PBCH decoder tries 1, 2 and 4 (CRC mask)
  if (srslte_ue_dl_init(&m->ue_dl, c2) != SRSLTE_SUCCESS ||
      cudaMalloc((void**)&m->d_result, 4 * sizeof(int32_t)) != cudaSuccess || cudaMalloc((void**)&m->d_mib, 24) != cudaSuccess) {
    srslte_ue_dl_free(&m->ue_dl);
    cudaFree(m->d_result); cudaFree(m->d_mib);
    delete m;
    return SRSLTE_ERROR;
  }
  q->cell = cell;
  q->gpu = m; q->pbch.gpu = m;
  return SRSLTE_SUCCESS;
}

void srslte_ue_mib_free(srslte_ue_mib_t* q) {
  if (!q || !q->gpu) return;
  auto* m = static_cast<MibGpu*>(q->gpu);
  srslte_ue_dl_free(&m->ue_dl);
  cudaFree(m->d_result); cudaFree(m->d_mib);
  delete m;
  q->gpu = nullptr; q->pbch.gpu = nullptr;
}

void srslte_pbch_decode_reset(srslte_pbch_t*) {}      // nothing is accumulated between frames

int srslte_ue_mib_decode(srslte_ue_mib_t* q, cf_t* input, uint8_t bch_payload[SRSLTE_BCH_PAYLOAD_LEN], uint32_t* nof_tx_ports,
                         uint32_t* sfn_offset) {
  if (!q || !q->gpu || !input || !bch_payload) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* m = static_cast<MibGpu*>(q->gpu);
  auto* u = static_cast<UeDlGpu*>(m->ue_dl.gpu);
  srsue_gpu_pdsch_plan_t* fp = front_plan(u, 0, 1);
  if (!fp) return SRSLTE_ERROR;
  if (cudaMemcpyAsync(u->d_iq, input, u->sf_len * sizeof(srsue_gpu_cf_t), cudaMemcpyHostToDevice, u->stream) != cudaSuccess) return SRSLTE_ERROR;
  if (srsue_gpu_ofdm_rx_cfo(fp, 1, u->d_iq, u->d_sf, nullptr, u->cfo_step, u->stream)) return SRSLTE_ERROR;
  if (srsue_gpu_chest(fp, 1, u->d_sf, u->d_ce, u->d_meas, u->stream)) return SRSLTE_ERROR;
  if (srsue_gpu_pbch_decode(fp, 1, u->d_sf, u->d_ce, u->d_meas, 0.0f, 1, m->d_result, m->d_mib, u->stream)) return SRSLTE_ERROR;
  int32_t res[4];
  uint8_t mib[24];
  cudaMemcpyAsync(res, m->d_result, sizeof(res), cudaMemcpyDeviceToHost, u->stream);
  cudaMemcpyAsync(mib, m->d_mib, sizeof(mib), cudaMemcpyDeviceToHost, u->stream);
  if (cudaStreamSynchronize(u->stream) != cudaSuccess) return SRSLTE_ERROR;
  u->dev_valid = false;
  if (!res[0]) return SRSLTE_UE_MIB_NOTFOUND;
  std::memcpy(bch_payload, mib, 24);
  if (nof_tx_ports) *nof_tx_ports = (uint32_t)res[1];
  if (sfn_offset) *sfn_offset = (uint32_t)res[2];
  return SRSLTE_UE_MIB_FOUND;
}

// 36.331 MasterInformationBlock: dl-Bandwidth (3), phich-Duration (1), phich-Resource (2), systemFrameNumber (8), spare (10)
void srslte_pbch_mib_unpack(uint8_t* msg, srslte_cell_t* cell, uint32_t* sfn) {
  if (!msg) return;
  auto field = [&](int pos, int n) { uint32_t v = 0; for (int i = 0; i < n; i++) v = (v << 1) | (msg[pos + i] & 1); return v; };
  static const uint32_t bw[8] = {6, 15, 25, 50, 75, 100, 0, 0};
  if (cell) {
    cell->nof_prb = bw[field(0, 3)];
    cell->phich_length = field(3, 1) ? SRSLTE_PHICH_EXT : SRSLTE_PHICH_NORM;
    cell->phich_resources = (srslte_phich_resources_t)field(4, 2);
  }
  if (sfn) *sfn = field(6, 8) << 2;
}

void srslte_pbch_mib_pack(srslte_cell_t* cell, uint32_t sfn, uint8_t* msg) {
  if (!cell || !msg) return;
  uint32_t b = 0;
  for (; b < 6; b++) { static const uint32_t bw[6] = {6, 15, 25, 50, 75, 100}; if (bw[b] == cell->nof_prb) break; }
  const uint32_t v = (b << 21) | ((cell->phich_length == SRSLTE_PHICH_EXT ? 1u : 0u) << 20) | ((uint32_t)cell->phich_resources << 18) |
                     (((sfn >> 2) & 0xFFu) << 10);
  for (int i = 0; i < 24; i++) msg[i] = (uint8_t)((v >> (23 - i)) & 1u);
}

void srslte_sch_set_max_noi(srslte_sch_t* q, uint32_t max_iterations) { if (q) q->max_iterations = max_iterations; }
uint32_t srslte_pdsch_last_noi(srslte_pdsch_t* q) { return q ? q->dl_sch.nof_iterations : 0; }

int srslte_ue_dl_decode_rnti(srslte_ue_dl_t* q, cf_t* input, uint8_t* data, uint32_t tti, uint16_t rnti) {
  if (!q || !q->gpu || !input || !data) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* u = static_cast<UeDlGpu*>(q->gpu);
  uint32_t cfi = 0;
  const uint32_t sf_idx = tti % 10;
  int rc = srslte_ue_dl_decode_fft_estimate(q, input, sf_idx, &cfi);
  if (rc < 0) return rc;
  uint32_t rv = u->grant_rv;
  if (u->have_grant) {
    rc = srslte_ue_dl_cfg_grant(q, &u->grant, u->grant_cfi ? u->grant_cfi : cfi, sf_idx, rv);   // cfi 0: from the PCFICH
  } else {
    // srsLTE's own sequence: PDCCH soft bits, blind search for this rnti, DCI -> grant (needs the installed TBS table)
    if (!srsue_gpu_ra_have_tbs_table()) return 0;      // no grant source at all
    rc = srslte_pdcch_extract_llr(&q->pdcch, q->sf_symbols, q->ce, srslte_chest_dl_get_noise_estimate(&q->chest), sf_idx, cfi);
    if (rc) return rc;
    srslte_dci_msg_t dci_msg;
    const srslte_rnti_type_t type = rnti == SRSLTE_SIRNTI ? SRSLTE_RNTI_SI : rnti == SRSLTE_PRNTI ? SRSLTE_RNTI_PCH
                                    : (rnti >= SRSLTE_RARNTI_START && rnti <= SRSLTE_RARNTI_END) ? SRSLTE_RNTI_RAR : SRSLTE_RNTI_USER;
    rc = srslte_ue_dl_find_dl_dci_type(q, &dci_msg, cfi, sf_idx, rnti, type);
    if (rc != 1) return rc < 0 ? rc : 0;               // no DCI for this rnti
    srslte_ra_dl_dci_t dci_unpacked;
    srslte_ra_dl_grant_t grant;
    if (srslte_dci_msg_to_dl_grant(&dci_msg, rnti, q->cell.nof_prb, &dci_unpacked, &grant)) return SRSLTE_ERROR;
    if (grant.mcs.tbs <= 0) return 0;                  // MCS 29..31: the size belongs to MAC's HARQ entity, not to this wrapper
    rv = (uint32_t)dci_unpacked.rv_idx;
    rc = srslte_ue_dl_cfg_grant(q, &grant, cfi, sf_idx, rv);
  }
  if (rc) return rc;
  if (rv == 0) srslte_softbuffer_rx_reset(&q->softbuffer);
  q->pkts_total++;
  // srslte_ue_dl_decode uses the channel estimator's noise figure (srsUE's worker passes 0.01 instead)
  rc = srslte_pdsch_decode_rnti(&q->pdsch, &q->pdsch_cfg, &q->softbuffer, q->sf_symbols, q->ce,
                                srslte_chest_dl_get_noise_estimate(&q->chest), rnti, data);
  if (rc == SRSLTE_SUCCESS) return q->pdsch_cfg.grant.mcs.tbs;
  q->pkt_errors++;
  return (rc == SRSLTE_ERROR) ? 0 : rc;
}

int srslte_ue_dl_decode(srslte_ue_dl_t* q, cf_t* input, uint8_t* data, uint32_t tti) {
  return srslte_ue_dl_decode_rnti(q, input, data, tti, q ? q->current_rnti : 0);
}

// ---- channel estimator getters -----------------------------------------------------------------------
float srslte_chest_dl_get_noise_estimate(srslte_chest_dl_t* q) { return q ? q->noise_estimate : 0.f; }
float srslte_chest_dl_get_snr(srslte_chest_dl_t* q) { return (q && q->noise_estimate > 0.f) ? q->rsrp / q->noise_estimate : 0.f; }
float srslte_chest_dl_get_rssi(srslte_chest_dl_t* q) { return q ? q->rssi : 0.f; }
float srslte_chest_dl_get_rsrq(srslte_chest_dl_t* q) { return q ? q->rsrq : 0.f; }
float srslte_chest_dl_get_rsrp(srslte_chest_dl_t* q) { return q ? q->rsrp : 0.f; }

// ---- soft buffer ---------------------------------------------------------------------------------------
int srslte_softbuffer_rx_init(srslte_softbuffer_rx_t* q, uint32_t nof_prb) {
  if (!q || nof_prb == 0 || nof_prb > SRSLTE_MAX_PRB) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(q, 0, sizeof(*q));
  if (!shared_ctx()) return SRSLTE_ERROR;
  // largest transport block of the bandwidth: I_TBS 26 row of 36.213 Table 7.1.7.2.1-1 scales ~ 753.76 bits/PRB
  const uint32_t max_tbs = (nof_prb >= 100) ? 75376u : (uint32_t)(nof_prb * 760u + 24u);
  q->max_cb = (max_tbs + 24 + 6119) / 6120 + 1;
  q->buffer_f = (int16_t**)calloc(q->max_cb, sizeof(int16_t*));
  if (!q->buffer_f) return SRSLTE_ERROR;
  for (uint32_t i = 0; i < q->max_cb; i++) {
    q->buffer_f[i] = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (3 * kMaxK + 12));
    if (!q->buffer_f[i]) { srslte_softbuffer_rx_free(q); return SRSLTE_ERROR; }
    std::memset(q->buffer_f[i], 0, sizeof(int16_t) * (3 * kMaxK + 12));
  }
  auto* sh = new SbShadow();
  sh->elems = (size_t)sb_total_elems(q->max_cb);
  if (cudaMalloc((void**)&sh->d_buf, sh->elems * sizeof(int16_t)) != cudaSuccess) { delete sh; srslte_softbuffer_rx_free(q); return SRSLTE_ERROR; }
  q->gpu_shadow = sh;
  return SRSLTE_SUCCESS;
}

void srslte_softbuffer_rx_free(srslte_softbuffer_rx_t* q) {
  if (!q) return;
  if (q->buffer_f) {
    for (uint32_t i = 0; i < q->max_cb; i++) srslte_vec_free(q->buffer_f[i]);
    free(q->buffer_f);
  }
  auto* sh = static_cast<SbShadow*>(q->gpu_shadow);
  if (sh) { cudaFree(sh->d_buf); delete sh; }
  std::memset(q, 0, sizeof(*q));
}

void srslte_softbuffer_rx_reset(srslte_softbuffer_rx_t* q) {
  // new data (dl_harq.cc:232): the next decode overwrites the device buffer instead of accumulating
  if (!q) return;
  auto* sh = static_cast<SbShadow*>(q->gpu_shadow);
  if (sh) sh->valid_tbs = 0;
  if (q->buffer_f)
    for (uint32_t i = 0; i < q->max_cb; i++)
      if (q->buffer_f[i]) std::memset(q->buffer_f[i], 0, sizeof(int16_t) * (3 * kMaxK + 12));
}
void srslte_softbuffer_rx_reset_tbs(srslte_softbuffer_rx_t* q, uint32_t) { srslte_softbuffer_rx_reset(q); }
void srslte_softbuffer_rx_reset_cb(srslte_softbuffer_rx_t* q, uint32_t) { srslte_softbuffer_rx_reset(q); }

int srsue_gpu_softbuffer_rx_sync_host(srslte_softbuffer_rx_t* q, uint32_t tbs) {
  if (!q || !q->gpu_shadow) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* sh = static_cast<SbShadow*>(q->gpu_shadow);
  CbSegm s;
  if (!cbsegm((int)tbs, &s) || (uint32_t)s.C > q->max_cb) return SRSLTE_ERROR_INVALID_INPUTS;
  const int stride = std::max(turbo_geom(s.Kp).cb_elems, s.Cm ? turbo_geom(s.Km).cb_elems : 0);
  srsue_gpu_ctx_t* ctx = shared_ctx();
  int16_t* d_tmp = nullptr;
  if (cudaMalloc((void**)&d_tmp, sizeof(int16_t) * (3 * kMaxK + 12)) != cudaSuccess) return SRSLTE_ERROR;
  int rc = SRSLTE_SUCCESS;
  for (int r = 0; r < s.C && rc == SRSLTE_SUCCESS; r++) {
    const int K = cb_len(s, r);
    if (srsue_gpu_tdec_export(ctx, sh->d_buf + (size_t)r * stride, 1, K, d_tmp, nullptr)) rc = SRSLTE_ERROR;
    else if (cudaMemcpy(q->buffer_f[r], d_tmp, sizeof(int16_t) * (3 * K + 12), cudaMemcpyDeviceToHost) != cudaSuccess) rc = SRSLTE_ERROR;
  }
  cudaFree(d_tmp);
  return rc;
}

// ---- turbo decoder object --------------------------------------------------------------------------------
namespace {
struct TdecGpu {
  srsue_gpu_ctx_t* ctx = nullptr;
  std::vector<int16_t> input;     // copy of the block handed to srslte_tdec_iteration
  uint32_t K = 0;
};
}

int srslte_tdec_init(srslte_tdec_t* h, uint32_t max_long_cb) {
  if (!h || max_long_cb == 0 || max_long_cb > kMaxK) return SRSLTE_ERROR_INVALID_INPUTS;
  std::memset(h, 0, sizeof(*h));
  srsue_gpu_ctx_t* ctx = shared_ctx();
  if (!ctx) return SRSLTE_ERROR;
  auto* g = new TdecGpu();
  g->ctx = ctx;
  h->gpu = g;
  h->max_long_cb = max_long_cb;
  return SRSLTE_SUCCESS;
}

void srslte_tdec_free(srslte_tdec_t* h) {
  if (!h) return;
  delete static_cast<TdecGpu*>(h->gpu);
  std::memset(h, 0, sizeof(*h));
}

int srslte_tdec_reset(srslte_tdec_t* h, uint32_t long_cb) {
  if (!h || !h->gpu || long_cb > h->max_long_cb || qpp_index((int)long_cb) < 0) return SRSLTE_ERROR_INVALID_INPUTS;
  h->n_iter = 0;
  static_cast<TdecGpu*>(h->gpu)->K = long_cb;
  return SRSLTE_SUCCESS;
}

// The decoder state after n calls is a pure function of (input, n); the device decoder keeps no state
// between launches, so iteration() records the request and decision*() runs the n iterations.
void srslte_tdec_iteration(srslte_tdec_t* h, int16_t* input, uint32_t long_cb) {
  if (!h || !h->gpu || !input) return;
  auto* g = static_cast<TdecGpu*>(h->gpu);
  g->K = long_cb;
  g->input.assign(input, input + 3 * long_cb + 12);
  h->n_iter++;
}

void srslte_tdec_decision_byte(srslte_tdec_t* h, uint8_t* output, uint32_t long_cb) {
  if (!h || !h->gpu || !output) return;
  auto* g = static_cast<TdecGpu*>(h->gpu);
  if (g->input.empty() || g->K != long_cb || h->n_iter == 0) return;
  int32_t st = 0;
  if (srsue_gpu_tdec_run_all_host(g->ctx, g->input.data(), 1, (int)long_cb, (int)h->n_iter, 0, output, &st))
    fprintf(stderr, "libsrsue_gpu: %s\n", srsue_gpu_last_error());
}

void srslte_tdec_decision(srslte_tdec_t* h, uint8_t* output, uint32_t long_cb) {
  std::vector<uint8_t> packed(long_cb / 8);
  srslte_tdec_decision_byte(h, packed.data(), long_cb);
  for (uint32_t i = 0; i < long_cb; i++) output[i] = (packed[i >> 3] >> (7 - (i & 7))) & 1;
}

int srslte_tdec_run_all(srslte_tdec_t* h, int16_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb) {
  if (!h || !h->gpu || !input || !output || nof_iterations == 0) return SRSLTE_ERROR_INVALID_INPUTS;
  if (srslte_tdec_reset(h, long_cb)) return SRSLTE_ERROR_INVALID_INPUTS;
  auto* g = static_cast<TdecGpu*>(h->gpu);
  int32_t st = 0;
  if (srsue_gpu_tdec_run_all_host(g->ctx, input, 1, (int)long_cb, (int)nof_iterations, 0, output, &st)) return SRSLTE_ERROR;
  h->n_iter = nof_iterations;
  return SRSLTE_SUCCESS;
}

}  // extern "C"

// pdsch_offline.cc -- offline downlink decoder: the C++ host side above the C ABI.
//
// Mode "worker" replays, subframe by subframe, exactly the call sequence of srsUE's PHY worker
// (/root/reference/ue/src/phy/phch_worker.cc:132-243: extract_fft_and_pdcch_llr -> new_grant_dl ->
// decode_pdsch -> tb_decoded) through the srsLTE-shaped symbols of libsrsue_gpu, with a MAC stand-in that owns
// the soft buffer and the payload buffer the way dl_harq does (/root/reference/ue/src/mac/dl_harq.cc:216-279;
// the same minimal caller as ue/test/phy/ue_itf_test_sib1.cc:108-123).  Mode "batch" hands all subframes to
// srsue_gpu_pdsch_decode_batch_host in one call -- the batching layer that replaces the one-subframe-per-
// thread hand-off of thread_pool::start_worker (ue/src/common/thread_pool.cc:246-254).
//
//   pdsch_offline worker|batch <in.bin> <out.bin>
//   pdsch_offline acquire <capture.bin> <out.txt>     cell search -> MIB -> subframe synchronisation on a 1.92 Msps capture
// in.bin : 12 int32 {magic 0x53525355 (normal cyclic prefix) or 0x53525345 (extended), nof_prb, nof_ports (1, 2, 4), cell_id, sf_idx,
//          cfi, rnti, qm, tbs, rv, n_sf, max_iter} followed by n_sf * SRSLTE_SF_LEN_PRB(nof_prb) cf_t samples
// out.bin: per subframe {int32 ack, int32 n_iter, float snr} followed by tbs/8 payload bytes
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <vector>

#include "srsue_gpu/srslte_compat.h"

namespace {

struct Header { int32_t magic, nof_prb, nof_ports, cell_id, sf_idx, cfi, rnti, qm, tbs, rv, n_sf, max_iter; };
constexpr int32_t kMagicNorm = 0x53525355, kMagicExt = 0x53525345;
inline int header_cp(const Header& h) { return h.magic == kMagicExt ? 1 : 0; }

// what mac->new_grant_dl() hands back to the worker (mac_interface.h:62-74), reduced to the fields the DL path reads
struct tb_action_dl_t {
  bool decode_enabled;
  uint32_t rv;
  uint16_t rnti;
  uint8_t* payload_ptr;
  srslte_softbuffer_rx_t* softbuffer;
  srslte_ra_dl_grant_t phy_grant;
};

class mac_stub {   // one HARQ process: owns the soft buffer and the PDU buffer
 public:
  bool init(uint32_t nof_prb, uint32_t tbs) {
    payload.assign(tbs / 8 + 8, 0);
    return srslte_softbuffer_rx_init(&softbuffer, nof_prb) == SRSLTE_SUCCESS;
  }
  ~mac_stub() { srslte_softbuffer_rx_free(&softbuffer); }
  void new_grant_dl(const srslte_ra_dl_grant_t& grant, uint32_t rv, uint16_t rnti, tb_action_dl_t* action) {
    if (rv == 0) srslte_softbuffer_rx_reset_tbs(&softbuffer, (uint32_t)grant.mcs.tbs);   // new data: dl_harq.cc:232
    action->decode_enabled = true;
    action->rv = rv;
    action->rnti = rnti;
    action->payload_ptr = payload.data();
    action->softbuffer = &softbuffer;
    action->phy_grant = grant;
  }
  std::vector<uint8_t> payload;
  srslte_softbuffer_rx_t softbuffer;
};

int run_worker(const Header& h, cf_t* iq, FILE* out) {
  srslte_cell_t cell;
  std::memset(&cell, 0, sizeof(cell));
  cell.nof_prb = h.nof_prb; cell.nof_ports = h.nof_ports; cell.id = h.cell_id; cell.cp = header_cp(h) ? SRSLTE_CP_EXT : SRSLTE_CP_NORM;
  srslte_ue_dl_t ue_dl;
  if (srslte_ue_dl_init(&ue_dl, cell)) { fprintf(stderr, "Initiating UE DL: %s\n", srsue_gpu_last_error()); return 1; }
  srslte_ue_dl_set_rnti(&ue_dl, (uint16_t)h.rnti);
  if (h.max_iter > 0) srslte_sch_set_max_noi(&ue_dl.pdsch.dl_sch, h.max_iter);     // phch_worker.cc:87-89
  srsue_gpu_ue_dl_set_cfi(&ue_dl, h.cfi);                                          // 0: decode the PCFICH; 1..3: capture without one
  mac_stub mac;
  if (!mac.init(h.nof_prb, h.tbs)) return 1;
  srslte_ra_dl_grant_t grant;
  std::memset(&grant, 0, sizeof(grant));
  for (int i = 0; i < h.nof_prb; i++) grant.prb_idx[0][i] = grant.prb_idx[1][i] = true;
  grant.nof_prb = h.nof_prb; grant.Qm = h.qm; grant.mcs.tbs = h.tbs;
  grant.mcs.mod = h.qm == 2 ? SRSLTE_MOD_QPSK : h.qm == 4 ? SRSLTE_MOD_16QAM : SRSLTE_MOD_64QAM;
  const int sf_len = SRSLTE_SF_LEN_PRB(h.nof_prb);
  std::vector<double> lat_us;
  for (int n = 0; n < h.n_sf; n++) {
    cf_t* signal_buffer = iq + (size_t)n * sf_len;
    uint32_t cfi = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (srslte_ue_dl_decode_fft_estimate(&ue_dl, signal_buffer, h.sf_idx, &cfi) < 0) { fprintf(stderr, "Getting PDCCH FFT estimate\n"); return 1; }
    tb_action_dl_t dl_action;
    mac.new_grant_dl(grant, h.rv, (uint16_t)h.rnti, &dl_action);
    bool dl_ack = false;
    if (dl_action.decode_enabled && !srslte_ue_dl_cfg_grant(&ue_dl, &dl_action.phy_grant, cfi, h.sf_idx, dl_action.rv)) {
      if (ue_dl.pdsch_cfg.grant.mcs.mod > 0 && ue_dl.pdsch_cfg.grant.mcs.tbs >= 0) {
        const float noise_estimate = 0.01f;                                        // phch_worker.cc:340
        dl_ack = srslte_pdsch_decode_rnti(&ue_dl.pdsch, &ue_dl.pdsch_cfg, dl_action.softbuffer, ue_dl.sf_symbols, ue_dl.ce,
                                          noise_estimate, dl_action.rnti, dl_action.payload_ptr) == 0;
      }
    }
    lat_us.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    const int32_t ack = dl_ack, n_iter = (int32_t)srslte_pdsch_last_noi(&ue_dl.pdsch);
    const float snr = srslte_chest_dl_get_snr(&ue_dl.chest);
    fwrite(&ack, 4, 1, out); fwrite(&n_iter, 4, 1, out); fwrite(&snr, 4, 1, out);
    fwrite(dl_action.payload_ptr, 1, h.tbs / 8, out);
  }
  // per-subframe latency of the worker sequence (the real-time budget of a phch_worker is a few ms, phy.h:118-119)
  if (lat_us.size() > 2) {
    std::vector<double> v(lat_us.begin() + 2, lat_us.end());        // the first calls build plans and tables
    std::sort(v.begin(), v.end());
    double sum = 0;
    for (double x : v) sum += x;
    fprintf(stderr, "worker latency per subframe over %zu subframes: mean %.0f us, median %.0f us, max %.0f us\n", v.size(),
            sum / v.size(), v[v.size() / 2], v.back());
  }
  srslte_ue_dl_free(&ue_dl);
  return 0;
}

int run_batch(const Header& h, cf_t* iq, FILE* out) {
  srsue_gpu_ctx_t* ctx = nullptr;
  if (srsue_gpu_ctx_create(&ctx, 0)) { fprintf(stderr, "%s\n", srsue_gpu_last_error()); return 1; }
  srsue_gpu_cell_t cell = {h.nof_prb, h.nof_ports, h.cell_id, header_cp(h)};
  srsue_gpu_pdsch_cfg_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.sf_idx = h.sf_idx; cfg.cfi = h.cfi; cfg.rnti = h.rnti; cfg.qm = h.qm; cfg.tbs = h.tbs; cfg.rv = h.rv;
  cfg.tm = h.nof_ports == 1 ? 1 : 2; cfg.nof_prb_alloc = h.nof_prb;
  for (int i = 0; i < h.nof_prb; i++) cfg.prb_mask[i] = 1;
  srsue_gpu_pdsch_plan_t* plan = nullptr;
  if (srsue_gpu_pdsch_plan_create(ctx, &cell, &cfg, h.n_sf, &plan)) { fprintf(stderr, "%s\n", srsue_gpu_last_error()); return 1; }
  srsue_gpu_plan_info_t info;
  srsue_gpu_pdsch_plan_info(plan, &info);
  std::vector<uint8_t> payload((size_t)h.n_sf * info.payload_stride);
  std::vector<int32_t> status((size_t)h.n_sf * 4);
  std::vector<float> meas((size_t)h.n_sf * 5);
  if (srsue_gpu_pdsch_decode_batch_host(plan, h.n_sf, reinterpret_cast<const srsue_gpu_cf_t*>(iq), 0.01f, 0, h.max_iter > 0 ? h.max_iter : 4,
                                        payload.data(), status.data(), meas.data())) {
    fprintf(stderr, "%s\n", srsue_gpu_last_error());
    return 1;
  }
  for (int n = 0; n < h.n_sf; n++) {
    const int32_t ack = status[4 * n], n_iter = status[4 * n + 2];
    const float snr = meas[5 * n + 4];
    fwrite(&ack, 4, 1, out); fwrite(&n_iter, 4, 1, out); fwrite(&snr, 4, 1, out);
    fwrite(payload.data() + (size_t)n * info.payload_stride, 1, h.tbs / 8, out);
  }
  srsue_gpu_pdsch_plan_destroy(plan);
  srsue_gpu_ctx_destroy(ctx);
  return 0;
}

// ---- mode "batch --gpus N": the same subframes as descriptors through ONE multi-GPU handle (srsue_gpu_batch_create_multi):
// the library splits the submission over devices 0..N-1 by estimated turbo work and gathers the results -- what N
// phch_workers behind the reference's thread pool do one subframe at a time (thread_pool.cc:206-254, phy.h:118-119)
int run_batch_multi(const Header& h, cf_t* iq, FILE* out, int n_gpus) {
  std::vector<int> devs(n_gpus);
  for (int i = 0; i < n_gpus; i++) devs[i] = i;
  srsue_gpu_batch_t* b = nullptr;
  if (srsue_gpu_batch_create_multi(devs.data(), n_gpus, h.n_sf, 0.01f, 0, h.max_iter > 0 ? h.max_iter : 4, &b)) {
    fprintf(stderr, "%s\n", srsue_gpu_last_error());
    return 1;
  }
  const size_t sf_len = SRSLTE_SF_LEN_PRB(h.nof_prb), pl = (size_t)h.tbs / 8;
  std::vector<uint8_t> payload((size_t)h.n_sf * pl);
  std::vector<srsue_gpu_sf_desc_t> d((size_t)h.n_sf);
  for (int n = 0; n < h.n_sf; n++) {
    std::memset(&d[n], 0, sizeof(d[n]));
    d[n].cell = {h.nof_prb, h.nof_ports, h.cell_id, header_cp(h)};
    d[n].cfg.sf_idx = h.sf_idx; d[n].cfg.cfi = h.cfi; d[n].cfg.rnti = h.rnti; d[n].cfg.qm = h.qm; d[n].cfg.tbs = h.tbs; d[n].cfg.rv = h.rv;
    d[n].cfg.tm = h.nof_ports == 1 ? 1 : 2; d[n].cfg.nof_prb_alloc = h.nof_prb;
    for (int i = 0; i < h.nof_prb; i++) d[n].cfg.prb_mask[i] = 1;
    d[n].iq = reinterpret_cast<const srsue_gpu_cf_t*>(iq + (size_t)n * sf_len);
    d[n].payload = payload.data() + (size_t)n * pl;
    d[n].softbuffer_id = -1;
    d[n].new_data = 1;
  }
  if (srsue_gpu_batch_submit(b, d.data(), h.n_sf) || srsue_gpu_batch_wait(b)) { fprintf(stderr, "%s\n", srsue_gpu_last_error()); return 1; }
  int nd = 0, share[64];
  srsue_gpu_batch_device_shares(b, &nd, share, nullptr, 64);
  for (int i = 0; i < nd; i++) fprintf(stderr, "device %d: %d subframes\n", devs[i], share[i]);
  for (int n = 0; n < h.n_sf; n++) {
    const int32_t ack = d[n].crc_ok, n_iter = d[n].n_iter;
    const float snr = d[n].meas[4];
    fwrite(&ack, 4, 1, out); fwrite(&n_iter, 4, 1, out); fwrite(&snr, 4, 1, out);
    fwrite(d[n].payload, 1, pl, out);
  }
  srsue_gpu_batch_destroy(b);
  return 0;
}

// ---- mode "acquire": what phch_recv does before any subframe reaches a worker (phch_recv.cc:136-264): cell search,
// MIB search, then subframe synchronisation with the system frame number taken from the MIB of subframes 0.  The
// "radio" is a capture file at 1.92 Msps that wraps around (in.bin: int32 magic 0x53525355, int32 n_samples, cf_t...).
struct FileRadio { const cf_t* x; size_t n, pos; };
int file_recv(void* h, void* data, uint32_t nsamples, srslte_timestamp_t* ts) {
  auto* r = static_cast<FileRadio*>(h);
  cf_t* out = static_cast<cf_t*>(data);
  for (uint32_t i = 0; i < nsamples; i++) { out[i] = r->x[r->pos]; r->pos = (r->pos + 1) % r->n; }
  if (ts) { ts->full_secs = 0; ts->frac_secs = 0.0; }
  return (int)nsamples;
}

int run_acquire(const cf_t* iq, size_t n, FILE* out) {
  FileRadio radio{iq, n, 0};
  srslte_ue_cellsearch_t cs;
  srslte_ue_cellsearch_result_t found_cells[3];
  std::memset(found_cells, 0, sizeof(found_cells));
  if (srslte_ue_cellsearch_init(&cs, file_recv, &radio)) { fprintf(stderr, "Initiating UE cell search\n"); return 1; }
  srslte_ue_cellsearch_set_nof_frames_to_scan(&cs, 8);
  srslte_ue_cellsearch_set_threshold(&cs, 15.0f);
  uint32_t max_peak_cell = 0;
  const int ret = srslte_ue_cellsearch_scan(&cs, found_cells, &max_peak_cell);
  srslte_ue_cellsearch_free(&cs);
  if (ret <= 0) { fprintf(stderr, "Could not find any PSS in this capture\n"); return 1; }
  srslte_cell_t cell;
  std::memset(&cell, 0, sizeof(cell));
  cell.id = found_cells[max_peak_cell].cell_id;
  cell.cp = found_cells[max_peak_cell].cp;
  const float cellsearch_cfo = found_cells[max_peak_cell].cfo;
  srslte_ue_mib_sync_t ue_mib_sync;
  if (srslte_ue_mib_sync_init(&ue_mib_sync, cell.id, cell.cp, file_recv, &radio)) { fprintf(stderr, "Initiating UE MIB synchronization\n"); return 1; }
  uint8_t bch_payload[SRSLTE_BCH_PAYLOAD_LEN];
  uint32_t sfn = 0, sfn_offset = 0;
  const int mret = srslte_ue_mib_sync_decode(&ue_mib_sync, 40, bch_payload, &cell.nof_ports, &sfn_offset);
  srslte_ue_mib_sync_free(&ue_mib_sync);
  if (mret != 1) { fprintf(stderr, "Error decoding MIB\n"); return 1; }
  srslte_pbch_mib_unpack(bch_payload, &cell, &sfn);
  fprintf(out, "cell_id %u ports %u prb %u phich_ng %d cfo_hz %.0f\n", cell.id, cell.nof_ports, cell.nof_prb, (int)cell.phich_resources,
          cellsearch_cfo);
  // the capture is at 1.92 Msps: synchronise on its six central PRBs and follow the frame number (sync_sfn, phch_recv.cc:230-264)
  srslte_cell_t view = cell;
  view.nof_prb = 6;
  srslte_ue_sync_t ue_sync;
  srslte_ue_mib_t ue_mib;
  if (srslte_ue_sync_init(&ue_sync, view, file_recv, &radio) || srslte_ue_mib_init(&ue_mib, view)) { fprintf(stderr, "Initiating ue_sync\n"); return 1; }
  srslte_ue_sync_set_cfo(&ue_sync, cellsearch_cfo);
  cf_t* sf_buffer = nullptr;
  int delivered = 0, mib_ok = 0, sf_errors = 0;
  int expect = -1;
  for (int i = 0; i < 80 && delivered < 40; i++) {
    const int r = srslte_ue_sync_get_buffer(&ue_sync, &sf_buffer);
    if (r < 0) { fprintf(stderr, "ue_sync failed\n"); return 1; }
    if (r == 0) continue;
    const int sf = (int)srslte_ue_sync_get_sfidx(&ue_sync);
    if (expect >= 0 && sf != expect) sf_errors++;
    expect = (sf + 1) % 10;
    delivered++;
    if (sf == 0) {
      uint32_t off = 0, s2 = 0;
      srslte_pbch_decode_reset(&ue_mib.pbch);
      if (srslte_ue_mib_decode(&ue_mib, sf_buffer, bch_payload, nullptr, &off) == SRSLTE_UE_MIB_FOUND) {
        srslte_pbch_mib_unpack(bch_payload, nullptr, &s2);
        fprintf(out, "tti %u\n", (s2 + off) * 10);
        mib_ok++;
      }
    }
  }
  fprintf(out, "delivered %d sf_errors %d mib_decoded %d\n", delivered, sf_errors, mib_ok);
  srslte_ue_mib_free(&ue_mib);
  srslte_ue_sync_free(&ue_sync);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  int n_gpus = 0;                                            // "batch <in> <out> --gpus N": the multi-GPU dispatcher
  if (argc == 6 && std::strcmp(argv[4], "--gpus") == 0) { n_gpus = atoi(argv[5]); argc = 4; }
  if (argc != 4 || n_gpus < 0 || n_gpus > 64) { fprintf(stderr, "usage: %s worker|batch|acquire <in.bin> <out> [--gpus N (batch only)]\n", argv[0]); return 2; }
  if (std::strcmp(argv[1], "acquire") == 0) {
    FILE* in = fopen(argv[2], "rb");
    if (!in) { perror(argv[2]); return 1; }
    int32_t hdr[2];
    if (fread(hdr, sizeof(hdr), 1, in) != 1 || hdr[0] != 0x53525355 || hdr[1] < 19200) { fprintf(stderr, "bad capture header\n"); return 1; }
    std::vector<cf_t> x((size_t)hdr[1]);
    if (fread(x.data(), sizeof(cf_t), x.size(), in) != x.size()) { fprintf(stderr, "short read\n"); return 1; }
    fclose(in);
    FILE* out = fopen(argv[3], "w");
    if (!out) { perror(argv[3]); return 1; }
    const int rc = run_acquire(x.data(), x.size(), out);
    fclose(out);
    return rc;
  }
  FILE* in = fopen(argv[2], "rb");
  if (!in) { perror(argv[2]); return 1; }
  Header h;
  if (fread(&h, sizeof(h), 1, in) != 1 || (h.magic != kMagicNorm && h.magic != kMagicExt)) { fprintf(stderr, "bad header\n"); return 1; }
  const size_t n = (size_t)h.n_sf * SRSLTE_SF_LEN_PRB(h.nof_prb);
  cf_t* iq = (cf_t*)srslte_vec_malloc((uint32_t)(n * sizeof(cf_t)));
  if (!iq || fread(iq, sizeof(cf_t), n, in) != n) { fprintf(stderr, "short read\n"); return 1; }
  fclose(in);
  FILE* out = fopen(argv[3], "wb");
  if (!out) { perror(argv[3]); return 1; }
  const int rc = std::strcmp(argv[1], "batch") == 0 ? (n_gpus ? run_batch_multi(h, iq, out, n_gpus) : run_batch(h, iq, out)) : run_worker(h, iq, out);
  fclose(out);
  srslte_vec_free(iq);
  return rc;
}

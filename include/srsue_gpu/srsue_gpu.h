/*
 * srsue_gpu.h -- C ABI of libsrsue_gpu: the batched, device-resident form of srsUE's downlink PDSCH
 * receive chain for NVIDIA B200 (sm_100a).  Plain pointers and sizes only; every function returns 0
 * (SRSLTE_SUCCESS) or a negative error (SRSLTE_ERROR = -1, or -2 for invalid arguments) like the
 * srsLTE calls it stands in for.  The per-subframe srsLTE-shaped entry points that phch_worker calls
 * (include/srsue_gpu/srslte_compat.h) are thin wrappers over this API with a batch of one.
 *
 * What each group replaces in the reference (kevinmel2000/srsUE, paths under /root/reference):
 *   srsue_gpu_ofdm_rx + srsue_gpu_chest   srslte_ue_dl_decode_fft_estimate   ue/src/phy/phch_worker.cc:254
 *   srsue_gpu_pdsch_plan_create           srslte_ue_dl_init / srslte_ue_dl_cfg_grant   phch_worker.cc:74,337
 *   srsue_gpu_pdsch_llr + _pdsch_turbo    srslte_pdsch_decode_rnti           phch_worker.cc:347-348
 *   srsue_gpu_pdsch_decode_batch[_host]   srslte_ue_dl_decode (the wrapper the north star names) and the
 *                                         thread_pool hand-off it replaces   ue/src/common/thread_pool.cc:72-82
 *   srsue_gpu_tdec_*                      srslte_tdec_run_all & co (turbo sweep, BASELINE config 4)
 *   srsue_gpu_ulsch_*                     srslte_ue_ul_pusch_encode_rnti_softbuffer's bit chain   phch_worker.cc:545-590
 *   tb_status / meas read-back            srslte_pdsch_last_noi, srslte_chest_dl_get_*  phch_worker.cc:359-360,799-848
 *
 * There is NO CPU fallback: every entry point fails (negative return, message in
 * srsue_gpu_last_error) when no CUDA device is usable.
 */
#ifndef SRSUE_GPU_H
#define SRSUE_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRSUE_GPU_SUCCESS 0
#define SRSUE_GPU_ERROR (-1)
#define SRSUE_GPU_ERROR_INVALID_INPUTS (-2)

typedef struct srsue_gpu_ctx srsue_gpu_ctx_t;           /* one per process and device */
typedef struct srsue_gpu_pdsch_plan srsue_gpu_pdsch_plan_t; /* one per (cell, grant shape) */

typedef struct { float re, im; } srsue_gpu_cf_t;        /* layout of srsLTE's cf_t (float _Complex) */

typedef struct {
  int nof_prb;      /* 6, 15, 25, 50, 75, 100 */
  int nof_ports;    /* 1, 2 or 4 (four: CRS of ports 2 / 3 in symbol 1 of each slot, PDSCH and control channels in SFBC-FSTD; tm 2 only) */
  int cell_id;      /* physical cell id */
  int cp;           /* cyclic prefix: 0 = normal (srsLTE's SRSLTE_CP_NORM), 1 = extended (SRSLTE_CP_EXT): 12 symbols per
                     * subframe, CRS in symbols 0 and 3 of each slot.  Device grids (sf_symbols, ce) keep a stride of 14
                     * symbols per subframe and port either way; with the extended prefix rows 12 and 13 are unused. */
} srsue_gpu_cell_t;

typedef struct {
  int sf_idx;       /* subframe 0..9 (tti % 10) */
  int cfi;          /* control format indicator 1..3 */
  int rnti;
  int qm;           /* bits per symbol: 2, 4, 6 */
  int tbs;          /* transport block size in bits */
  int rv;           /* redundancy version 0..3 */
  int tm;           /* 1: single antenna port, 2: transmit diversity */
  int nof_prb_alloc;
  uint8_t prb_mask[110];  /* 1 = PRB allocated in both slots; 2 = in slot 0 only, 4 = in slot 1 only (distributed
                           * virtual resource blocks, 36.211 6.2.3.2; 6 = 1) */
} srsue_gpu_pdsch_cfg_t;

typedef struct {
  int nfft, nsc, sf_len;          /* FFT size, 12*nof_prb, samples per subframe */
  int nof_re, G;                  /* PDSCH resource elements and coded bits per subframe */
  int C, Kp, Km, Cp, Cm, F;       /* code-block segmentation */
  int sb_cb_stride;               /* int16 elements per code block in the device soft buffer */
  int sb_sf_stride;               /* int16 elements per subframe (C * sb_cb_stride) */
  int payload_stride;             /* bytes per subframe in payload arrays */
  int max_batch;
} srsue_gpu_plan_info_t;

/* ---- context ------------------------------------------------------------------------------- */
int srsue_gpu_ctx_create(srsue_gpu_ctx_t **ctx, int device);
void srsue_gpu_ctx_destroy(srsue_gpu_ctx_t *ctx);
const char *srsue_gpu_last_error(void);
int srsue_gpu_version(void);

/* ---- cell search (srslte_ue_cellsearch_scan, ue/src/phy/phch_recv.cc:146-177) -------------------- */
typedef struct {
  int32_t peak_pos;     /* first sample of the PSS symbol body (after its cyclic prefix) inside the buffer */
  int32_t n_id_2;       /* 0..2 */
  int32_t n_id_1;       /* 0..167, -1 when the SSS symbol does not lie inside the buffer (peak_pos < 137 nfft / 128) */
  int32_t sf5;          /* 0: the PSS belongs to subframe 0, 1: to subframe 5 */
  int32_t valid;
  float peak;           /* |correlation|^2 at the peak */
  float mean_power;     /* mean of |correlation|^2 over all roots and positions (peak / mean_power = peak-to-side ratio) */
  float cfo;            /* carrier frequency offset in units of the 15 kHz subcarrier spacing */
  float sss_corr;
  int32_t cp;           /* cyclic prefix the SSS was found with: 0 normal, 1 extended (srsue_gpu_cell_search_cp) */
} srsue_gpu_sync_result_t;
/* d_iq: n_bufs buffers of n_samples (typically a 5 ms half frame = 75 * nfft), `stride` samples apart, at the sampling
 * rate whose OFDM symbol has nfft samples (128 = 1.92 Msps, the cell-search rate; up to 2048 = 30.72 Msps for tracking at
 * the cell's own rate).  force_n_id_2: -1 searches the three PSS roots, 0..2 only that one.  first_pos: first sample offset searched (0, or 137
 * scaled by nfft / 128 so that the SSS symbol in front of every candidate lies inside the buffer).  peak_pos is relative
 * to the buffer start.  Cell id = 3 * n_id_1 + n_id_2. */
int srsue_gpu_cell_search(srsue_gpu_ctx_t *ctx, const srsue_gpu_cf_t *d_iq, int n_bufs, int n_samples, long long stride, int nfft,
                          int force_n_id_2, int first_pos, srsue_gpu_sync_result_t *d_result, void *stream);
/* The same with the cyclic prefix as a parameter (what srslte_ue_cellsearch_scan reports in found_cells[].cp,
 * phch_recv.cc:189): cp_mode 0 looks for the SSS behind a normal prefix (= srsue_gpu_cell_search), 1 behind an extended
 * one (160 nfft / 128 samples before the PSS), 2 tries both and reports the better one in result.cp. */
int srsue_gpu_cell_search_cp(srsue_gpu_ctx_t *ctx, const srsue_gpu_cf_t *d_iq, int n_bufs, int n_samples, long long stride, int nfft,
                             int force_n_id_2, int first_pos, int cp_mode, srsue_gpu_sync_result_t *d_result, void *stream);

/* ---- turbo decoder (device pointers; `stream` is a cudaStream_t passed as void*) ------------ */
/* geometry of the windowed decoder for code-block size K: window length, windows, tcb elements */
int srsue_gpu_tdec_geometry(int K, int *W, int *P, int *cb_elems);
/* srsLTE decoder-input order [n_cb][3K+12] int16 -> device layout [n_cb][cb_elems], clamped to +-511 */
int srsue_gpu_tdec_import(srsue_gpu_ctx_t *ctx, const int16_t *d_triples, int n_cb, int K, int16_t *d_tcb, void *stream);
int srsue_gpu_tdec_export(srsue_gpu_ctx_t *ctx, const int16_t *d_tcb, int n_cb, int K, int16_t *d_triples, void *stream);
/* decode n_cb code blocks of size K from the device layout.  crc_type: 0 none (run max_iter
 * iterations), 1 CRC24A, 2 CRC24B (stop as soon as the remainder is zero).  d_bits: [n_cb][K/8]
 * bytes, MSB first.  d_status[cb] = iterations | crc_ok << 8. */
int srsue_gpu_tdec_decode(srsue_gpu_ctx_t *ctx, const int16_t *d_tcb, int n_cb, int K, int max_iter, int crc_type,
                          uint8_t *d_bits, int32_t *d_status, void *stream);
/* import + decode in one call (what srslte_tdec_run_all does for one block) */
int srsue_gpu_tdec_run_all(srsue_gpu_ctx_t *ctx, const int16_t *d_triples, int n_cb, int K, int max_iter, int crc_type,
                           uint8_t *d_bits, int32_t *d_status, void *stream);
/* host-pointer convenience of the same (copies in and out, synchronises) */
int srsue_gpu_tdec_run_all_host(srsue_gpu_ctx_t *ctx, const int16_t *h_triples, int n_cb, int K, int max_iter,
                                int crc_type, uint8_t *h_bits, int32_t *h_status);
/* launch statistics of the most recent decode on this context */
int srsue_gpu_tdec_last_launch(srsue_gpu_ctx_t *ctx, int *grid, int *block, int *smem_bytes, int *cb_per_cta);

/* ---- PDSCH plan ------------------------------------------------------------------------------ */
int srsue_gpu_pdsch_plan_create(srsue_gpu_ctx_t *ctx, const srsue_gpu_cell_t *cell, const srsue_gpu_pdsch_cfg_t *cfg,
                                int max_batch, srsue_gpu_pdsch_plan_t **plan);
void srsue_gpu_pdsch_plan_destroy(srsue_gpu_pdsch_plan_t *plan);
int srsue_gpu_pdsch_plan_info(const srsue_gpu_pdsch_plan_t *plan, srsue_gpu_plan_info_t *info);

/* ---- stages (device pointers) ---------------------------------------------------------------- */
/* d_iq [n_sf][sf_len] -> d_sf_symbols [n_sf][14*nsc] */
int srsue_gpu_ofdm_rx(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_iq, srsue_gpu_cf_t *d_sf_symbols,
                      void *stream);
/* The same with the carrier-frequency-offset correction srsLTE's synchroniser applies to the samples before the worker
 * sees them (srslte_cfo_correct inside srslte_ue_sync_zerocopy, ue/src/phy/phch_recv.cc:322), fused into the sample
 * loads of the transform.  Phase steps come from srsue_gpu_host_cfo_step(cfo in subcarrier spacings, nfft): one per
 * subframe in d_cfo_steps, or cfo_step for all of them when d_cfo_steps is NULL (0: no rotation). */
int srsue_gpu_ofdm_rx_cfo(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_iq, srsue_gpu_cf_t *d_sf_symbols,
                          const int32_t *d_cfo_steps, int32_t cfo_step, void *stream);
int srsue_gpu_host_cfo_step(float cfo, int nfft);
/* int16 {re, im} samples (the radio's wire format, the usual capture-file format; srsLTE converts them to float on the
 * CPU before its receiver sees them): sample = (float)v * scale, converted as the transform loads them.  Half the bytes
 * over PCIe and out of HBM.  d_iq16 [n_sf][sf_len][2]. */
int srsue_gpu_ofdm_rx_sc16(srsue_gpu_pdsch_plan_t *plan, int n_sf, const int16_t *d_iq16, float scale, srsue_gpu_cf_t *d_sf_symbols,
                           void *stream);
/* both: convert, then rotate by the carrier offset (steps as in srsue_gpu_ofdm_rx_cfo) */
int srsue_gpu_ofdm_rx_sc16_cfo(srsue_gpu_pdsch_plan_t *plan, int n_sf, const int16_t *d_iq16, float scale, srsue_gpu_cf_t *d_sf_symbols,
                               const int32_t *d_cfo_steps, int32_t cfo_step, void *stream);
/* What the d_iq / h_iq arguments of srsue_gpu_pdsch_decode_batch[_host] point at for this plan: SRSUE_GPU_IQ_CF32 (default)
 * or SRSUE_GPU_IQ_SC16 with its scale (e.g. 1/32768). */
/* carrier-offset correction for the batch calls of this plan: per-subframe steps on the device (row i of the
 * next srsue_gpu_pdsch_decode_batch call), or one step for all subframes when d_cfo_steps is NULL; (NULL, 0) switches it off */
int srsue_gpu_pdsch_plan_set_cfo(srsue_gpu_pdsch_plan_t *plan, const int32_t *d_cfo_steps, int32_t cfo_step);
enum { SRSUE_GPU_IQ_CF32 = 0, SRSUE_GPU_IQ_SC16 = 1 };
int srsue_gpu_pdsch_plan_set_iq_format(srsue_gpu_pdsch_plan_t *plan, int format, float scale);
/* First turbo iteration after which a passing code-block CRC ends the block (default 1: stop as soon as the CRC passes,
 * what srslte_pdsch_decode_rnti does with the budget of srslte_sch_set_max_noi, phch_worker.cc:88).  min_iter = max_iter
 * gives a fixed iteration count with the CRC verdicts still reported -- the "fixed 4 iterations" measurement of SURVEY 7.3. */
int srsue_gpu_pdsch_plan_set_min_iter(srsue_gpu_pdsch_plan_t *plan, int min_iter);
/* PDCCH calls of this plan (srsue_gpu_pdcch_extract_llr / _find_dci) skip every subframe i with d_values[i] != want and
 * report "not found" for it; (NULL, 0) switches the filter off.  A plan's control region belongs to one CFI: a batch
 * whose subframes carry different CFIs (d_values = the output of srsue_gpu_pcfich_decode) runs each plan on its own rows. */
int srsue_gpu_pdsch_plan_set_row_filter(srsue_gpu_pdsch_plan_t *plan, const int32_t *d_values, int want);
/* d_ce [n_sf][ports][14*nsc]; d_meas [n_sf][5] = noise, rsrp, rssi, rsrq, snr */
int srsue_gpu_chest(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols, srsue_gpu_cf_t *d_ce,
                    float *d_meas, void *stream);
/* PCFICH: the control format indicator srslte_ue_dl_decode_fft_estimate returns through *cfi (phch_worker.cc:254).
 * d_cfi [n_sf] = 1..3 (largest correlation, lowest CFI on ties); d_corr optional [n_sf][3] integer correlations of the
 * 32 descrambled int16 LLRs with the three code words.  Uses the plan's cell and sf_idx only. */
int srsue_gpu_pcfich_decode(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                            const srsue_gpu_cf_t *d_ce, const float *d_meas, float noise_est, int noise_mode,
                            int32_t *d_cfi, int32_t *d_corr, void *stream);
/* PDCCH (srslte_pdcch_extract_llr + srslte_ue_dl_find_dl_dci_type, phch_worker.cc:260,293).  The control region is
 * that of the plan's cell, cfi and sf_idx; ng_x6 = 6 x the PHICH Ng of the MIB (1, 3, 6, 12), normal PHICH duration.
 * _info: REGs and CCEs available to PDCCH.  _extract_llr: d_llr [n_sf][8 * n_reg] int16, 72 per CCE in CCE order.
 * _find_dci: blind search of the UE-specific (common = 0) or common search space for a payload of nof_bits bits whose
 * CRC is masked with rnti; d_found [n_sf][4] = {found, L, first CCE, candidate index}, d_bits [n_sf][64] one bit per
 * byte, d_rem optional [n_sf][candidates] = the RNTI each candidate decodes to.  first_bit: -1, or the value payload
 * bit 0 must have (formats 0 and 1A share size and search space; the flag tells them apart).  Returns the number of
 * candidates. */
int srsue_gpu_pdcch_info(srsue_gpu_pdsch_plan_t *plan, int ng_x6, int *n_reg, int *nof_cce);
int srsue_gpu_pdcch_extract_llr(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                                const srsue_gpu_cf_t *d_ce, const float *d_meas, float noise_est, int noise_mode, int ng_x6,
                                int16_t *d_llr, void *stream);
int srsue_gpu_pdcch_find_dci(srsue_gpu_pdsch_plan_t *plan, int n_sf, const int16_t *d_llr, int ng_x6, int rnti, int common,
                             int nof_bits, int first_bit, int32_t *d_found, uint8_t *d_bits, uint16_t *d_rem, void *stream);
/* PBCH / MIB (srslte_ue_mib_decode, phch_recv.cc:247): blind decode of the master information block from subframes 0
 * (plan with sf_idx 0): transmit-port hypotheses 1, 2 and 4 (as many as the plan's cell has ports, i.e. as were
 * estimated; CRC masks 0x0000 / 0xFFFF / 0x5555) x the four positions in the 40 ms BCH period.  d_result [n_sf][4] = {found, ports, frame number mod 4, 0},
 * d_mib [n_sf][24] one bit per byte (dl-Bandwidth 3, phich-Duration 1, phich-Resource 2, SFN/4 8, spare 10). */
int srsue_gpu_pbch_decode(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                          const srsue_gpu_cf_t *d_ce, const float *d_meas, float noise_est, int noise_mode,
                          int32_t *d_result, uint8_t *d_mib, void *stream);
/* PHICH (srslte_ue_dl_decode_phich, phch_worker.cc:381): HARQ indicator of (n_group, n_seq) in every subframe of the
 * batch; d_ack [n_sf] = 1 for ACK, d_metric optional [n_sf] (ACK iff < 0).  Normal PHICH duration; with the extended
 * cyclic prefix (spreading factor 2) n_group runs to twice the mapping units and n_seq to 3. */
int srsue_gpu_phich_decode(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                           const srsue_gpu_cf_t *d_ce, const float *d_meas, float noise_est, int noise_mode, int ng_x6,
                           int n_group, int n_seq, int32_t *d_ack, float *d_metric, void *stream);
/* equalise + demap + descramble + rate-dematch into d_softbuf [n_sf][sb_sf_stride].  noise_mode 0: use
 * noise_est (srsUE passes 0.01), 1: use d_meas[.][0].  accumulate 0: new transmission, 1: HARQ combine.
 * d_dbg_d [n_sf][nof_re] / d_dbg_e [n_sf][G] optional taps of the equalised symbols / descrambled LLRs. */
int srsue_gpu_pdsch_llr(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                        const srsue_gpu_cf_t *d_ce, const float *d_meas, float noise_est, int noise_mode, int accumulate,
                        int16_t *d_softbuf, srsue_gpu_cf_t *d_dbg_d, int16_t *d_dbg_e, void *stream);
/* fused variants used by the whole-chain call: the channel-estimate grid is never written to HBM.
 * srsue_gpu_chest_pilots leaves only the smoothed pilot estimates d_pilots [n_sf][ports][4][2*nof_prb] (and
 * the measurements); srsue_gpu_pdsch_llr_fused interpolates them per resource element (same arithmetic). */
int srsue_gpu_chest_pilots(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                           srsue_gpu_cf_t *d_pilots, float *d_meas, void *stream);
int srsue_gpu_pdsch_llr_fused(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_sf_symbols,
                              const srsue_gpu_cf_t *d_pilots, const float *d_meas, float noise_est, int noise_mode,
                              int accumulate, int16_t *d_softbuf, void *stream);
/* turbo decode + CRC + transport-block assembly.  d_payload [n_sf][payload_stride] bytes MSB first;
 * d_tb_status [n_sf][4] = {crc_ok, sum of iterations, floor(avg iterations), C};
 * d_cb_status optional [n_sf][C]. */
int srsue_gpu_pdsch_turbo(srsue_gpu_pdsch_plan_t *plan, int n_sf, const int16_t *d_softbuf, int max_iter,
                          uint8_t *d_payload, int32_t *d_tb_status, int32_t *d_cb_status, void *stream);

/* ---- whole chain ------------------------------------------------------------------------------ */
/* d_softbuf may be NULL (plan-owned scratch, new transmission).  d_meas may be NULL. */
int srsue_gpu_pdsch_decode_batch(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *d_iq, float noise_est,
                                 int noise_mode, int max_iter, int accumulate, int16_t *d_softbuf, uint8_t *d_payload,
                                 int32_t *d_tb_status, float *d_meas, void *stream);
/* same with HOST buffers (pinned or pageable): H2D of the IQ, the chain, D2H of payload/status/meas,
 * synchronised on return.  This is the call the offline driver and the e2e benchmark make. */
int srsue_gpu_pdsch_decode_batch_host(srsue_gpu_pdsch_plan_t *plan, int n_sf, const srsue_gpu_cf_t *h_iq, float noise_est,
                                      int noise_mode, int max_iter, uint8_t *h_payload, int32_t *h_tb_status,
                                      float *h_meas);
/* number of kernels the most recent chain call launched (for the benchmark's gpu_launches) */
int srsue_gpu_last_launch_count(srsue_gpu_ctx_t *ctx);

/* ---- batching layer: heterogeneous streams of subframes ------------------------------------------
 * Replaces, for batched/offline operation, the one-subframe-per-worker hand-off of the reference
 * (ue/src/phy/phch_recv.cc:309-369, ue/src/common/thread_pool.cc:72-82) and the MAC's ownership of one soft
 * buffer per HARQ process (ue/src/mac/dl_harq.cc:169-174,232).  A submission may mix any cells, bandwidths,
 * grants, redundancy versions and UEs; descriptors with the same launch shape are packed into one launch of
 * the chain each (plans are cached, least recently used evicted).  All pointers are HOST memory. */
typedef struct srsue_gpu_batch srsue_gpu_batch_t;
typedef struct {
  srsue_gpu_cell_t cell;
  srsue_gpu_pdsch_cfg_t cfg;       /* the grant as srslte_ue_dl_cfg_grant would configure it (incl. rv) */
  const srsue_gpu_cf_t *iq;        /* in:  one subframe of time samples (sf_len = 15*nfft), caller-owned; int16 {re, im}
                                    * pairs behind the same pointer after srsue_gpu_batch_set_iq_format(SC16) */
  uint8_t *payload;                /* out: tbs/8 bytes, MSB first, caller-owned */
  int64_t softbuffer_id;           /* < 0: no HARQ state (new transmission decoded from scratch);
                                    * >= 0: device-resident soft buffer of this (UE, HARQ process) id */
  int32_t new_data;                /* 1: reset the soft buffer first (dl_harq.cc:232); 0: combine into it */
  int32_t crc_ok;                  /* out: 1 when the transport-block CRC passed (ack) */
  int32_t n_iter;                  /* out: srslte_pdsch_last_noi() of this subframe */
  float meas[5];                   /* out: noise, rsrp, rssi, rsrq, snr of this subframe */
  float cfo;                       /* in:  carrier offset of this capture in subcarrier spacings (|cfo| < 1), removed while the
                                    * samples are transformed (oracle/SPEC.md 14); 0: none */
} srsue_gpu_sf_desc_t;

/* max_subframes bounds one submission.  noise_est / noise_mode / max_iter as in srsue_gpu_pdsch_decode_batch. */
int srsue_gpu_batch_create(srsue_gpu_ctx_t *ctx, int max_subframes, float noise_est, int noise_mode, int max_iter,
                           srsue_gpu_batch_t **batch);
/* The same batching layer over SEVERAL GPUs of one box (north star item 4; the reference's analogue is its pool
 * of phch_workers behind one object, ue/src/common/thread_pool.cc:206-254, ue/hdr/phy/phy.h:118-119).  The library creates
 * one context per listed device, each with its own streams, pinned staging, plan cache and resident soft buffers, and
 * drives each with its own host thread.  srsue_gpu_batch_submit then splits a submission: a soft buffer id stays on the
 * device that first saw it (HARQ state never moves), everything else is cut into contiguous runs balanced by estimated
 * turbo work (sum of C * K); the shares run concurrently, srsue_gpu_batch_wait gathers the results on the host.  No
 * collective, no peer traffic.  Every other srsue_gpu_batch_* call works on the returned handle unchanged. */
int srsue_gpu_batch_create_multi(const int *devices, int n_devices, int max_subframes, float noise_est, int noise_mode, int max_iter,
                                 srsue_gpu_batch_t **batch);
/* multi-GPU handles: how the last submission was split (subframes and estimated work per device, first `cap` entries);
 * n_devices = 0 for a single-device batch */
int srsue_gpu_batch_device_shares(const srsue_gpu_batch_t *batch, int *n_devices, int *subframes, double *work, int cap);
void srsue_gpu_batch_destroy(srsue_gpu_batch_t *batch);
/* what the `iq` pointer of every descriptor points at from the next submission on: SRSUE_GPU_IQ_CF32 (default) or
 * SRSUE_GPU_IQ_SC16 (int16 {re, im} pairs, sample = (float)v * scale) */
int srsue_gpu_batch_set_iq_format(srsue_gpu_batch_t *batch, int format, float scale);
/* enqueues uploads, launches and downloads for n descriptors and returns; `descs`, the IQ and the payload
 * buffers must stay valid until srsue_gpu_batch_wait, which also fills the out fields of every descriptor */
int srsue_gpu_batch_submit(srsue_gpu_batch_t *batch, srsue_gpu_sf_desc_t *descs, int n);
int srsue_gpu_batch_wait(srsue_gpu_batch_t *batch);
/* The reference's own sequence for a whole batch (phch_worker.cc:254-297): the caller knows the cell, the subframe number
 * and the RNTI of every capture, nothing else.  Phase 1 runs FFT + channel estimate + PCFICH (-> CFI) + the PDCCH blind
 * search for the RNTI (formats 1A then 1 in the UE-specific space, 1A in the common space; SI / RA / P-RNTI: 1A in the
 * common space) over the whole batch; the DCIs become grants on the host (srslte_dci_msg_to_dl_grant: needs the sizes of
 * srsue_gpu_ra_set_tbs_table or the built-in columns); phase 2 is the PDSCH chain bucketed by grant, fed from the samples
 * that are already on the device.  In: cell, cfg.sf_idx, cfg.rnti, iq, payload (payload_cap bytes each).  Out: the rest of
 * cfg and the usual results after srsue_gpu_batch_wait; no DCI for the RNTI: cfg.tbs = 0, crc_ok = 0.  ng_x6 = 6 x the
 * PHICH Ng of the cell's MIB.  Subframes are decoded as new transmissions (no soft-buffer ids). */
int srsue_gpu_batch_submit_blind(srsue_gpu_batch_t *batch, srsue_gpu_sf_desc_t *descs, int n, int ng_x6, int payload_cap);
/* frees the device soft buffer of one id (a HARQ process that was acknowledged or flushed) */
int srsue_gpu_batch_softbuffer_release(srsue_gpu_batch_t *batch, int64_t softbuffer_id);
/* cached plans, resident soft buffers, kernels launched by the last submission */
int srsue_gpu_batch_stats(const srsue_gpu_batch_t *batch, int *n_plans, int *n_softbuffers, int *launches);

/* ---- host-side bookkeeping, usable without a GPU (what srslte_ue_dl_cfg_grant computes on the host,
 * phch_worker.cc:337; exported so that the tables can be checked on a CPU-only machine) ------------- */
/* code-block segmentation of a transport block: out[8] = tbs, B, C, K+, K-, C+, C-, F */
int srsue_gpu_host_cbsegm(int tbs, int *out);
/* number of PDSCH resource elements of a grant; re_idx (optional) receives l*nsc + k for each */
int srsue_gpu_host_pdsch_re(const srsue_gpu_cell_t *cell, const srsue_gpu_pdsch_cfg_t *cfg, int32_t *re_idx);
/* PDCCH bookkeeping: data REs (grid index l*nsc + k, 4 per REG) of the REGs that carry PDCCH, in mapping order (returns
 * the number of REGs); quadruplet carried by each mapping position; search-space candidates (returns their number);
 * payload size of DCI format 1A (fmt 0) / 1 (fmt 1) */
int srsue_gpu_host_pdcch_regs(const srsue_gpu_cell_t *cell, int cfi, int ng_x6, int32_t *re4);
int srsue_gpu_host_pdcch_quad_perm(int n_quad, int cell_id, int32_t *src);
int srsue_gpu_host_pdcch_search_space(int nof_cce, int sf_idx, int rnti, int common, int32_t *cand_L, int32_t *cand_ncce);
int srsue_gpu_host_dci_format_sizeof(int fmt, int nof_prb);
/* grid indices (l * nsc + k) of the 240 PBCH resource elements of a subframe 0; with the extended cyclic prefix there
 * are 216 and g240[216..239] = -1 */
int srsue_gpu_host_pbch_res(const srsue_gpu_cell_t *cell, int32_t *g240);
/* PHICH bookkeeping: (group, sequence) answering an uplink transmission with lowest PRB I_lowest and DMRS cyclic shift
 * n_dmrs (36.213 9.1.2); the 12 subcarriers of a group in OFDM symbol 0 */
int srsue_gpu_host_phich_index(int nof_prb, int ng_x6, int I_lowest, int n_dmrs, int *n_group, int *n_seq);
/* the same for either cyclic prefix (cp = 1: twice the groups, n_seq modulo 4; 36.213 9.1.2) */
int srsue_gpu_host_phich_index_cp(int nof_prb, int ng_x6, int cp, int I_lowest, int n_dmrs, int *n_group, int *n_seq);
int srsue_gpu_host_phich_res(const srsue_gpu_cell_t *cell, int n_group, int32_t *k12);
/* subcarriers (in OFDM symbol 0) of the 16 PCFICH symbols d(0..15) */
int srsue_gpu_host_pcfich_re(const srsue_gpu_cell_t *cell, int32_t *k16);
/* rate-matching read order for (K, F, rv): seq[n] = index 3k+stream of the n-th non-null circular-buffer
 * position (srsLTE decoder-input order); returns the number of entries (<= 3K+12) */
int srsue_gpu_host_rm_sequence(int K, int F, int rv, int32_t *seq);
/* QPP interleaver pi(i) for code-block size K */
int srsue_gpu_host_qpp(int K, uint16_t *pi);
/* n bits of the Gold sequence with the given c_init, one per byte */
int srsue_gpu_host_gold(uint32_t c_init, int n, uint8_t *c);

/* ---- uplink shared-channel encoder ------------------------------------------------
 * The bit chain of srslte_ue_ul_pusch_encode_rnti_softbuffer (ue/src/phy/phch_worker.cc:545-590: srslte_ulsch_encode and
 * the scrambling of srslte_pusch_encode) for a batch of transport blocks of one grant: CRC24A, segmentation + CRC24B,
 * turbo ENCODER, rate matching for redundancy version rv over the whole circular buffer, concatenation, channel
 * interleaver (36.212 5.2.2.8; no control information multiplexed), scrambling (36.211 5.3.1).  The caller's modulation
 * mapper, DFT precoding and SC-FDMA generation stay where they are.  n_symb = PUSCH symbols per subframe: 12 with the
 * normal cyclic prefix, 11 when the last one gives way to SRS (10 / 9 with the extended prefix). */
typedef struct {
  int tbs;        /* transport-block size in bits (multiple of 8) */
  int qm;         /* 2, 4, 6 */
  int nof_prb;    /* L_prb of the grant */
  int n_symb;
  int rv;
  int rnti, sf_idx, cell_id;
} srsue_gpu_ulsch_cfg_t;
typedef struct srsue_gpu_ulsch_plan srsue_gpu_ulsch_plan_t;
int srsue_gpu_ulsch_plan_create(srsue_gpu_ctx_t *ctx, const srsue_gpu_ulsch_cfg_t *cfg, int max_batch, srsue_gpu_ulsch_plan_t **plan);
void srsue_gpu_ulsch_plan_destroy(srsue_gpu_ulsch_plan_t *plan);
/* G = 12 nof_prb n_symb qm coded bits per transport block; C code blocks of sizes Km (the first ones) and Kp */
int srsue_gpu_ulsch_plan_info(const srsue_gpu_ulsch_plan_t *plan, int *G, int *C, int *Kp, int *Km);
/* d_payload [n_tb][tbs / 8]; d_bits [n_tb][G / 8], packed MSB first (bit 0 of the codeword is bit 7 of byte 0) */
int srsue_gpu_ulsch_encode(srsue_gpu_ulsch_plan_t *plan, int n_tb, const uint8_t *d_payload, uint8_t *d_bits, void *stream);
/* host-pointer convenience of the same (copies in and out, synchronises) */
int srsue_gpu_ulsch_encode_host(srsue_gpu_ulsch_plan_t *plan, int n_tb, const uint8_t *h_payload, uint8_t *h_bits);

/* ---- pinned host memory helpers (so callers need not link the CUDA runtime) ------------------- */
void *srsue_gpu_host_alloc(uint64_t bytes);
void srsue_gpu_host_free(void *p);
/* page-locks a caller-owned region (e.g. the capture buffer).  Subframes and payload buffers that lie inside
 * regions from srsue_gpu_host_alloc / srsue_gpu_host_register are fetched and written by the GPU directly when
 * they go through srsue_gpu_batch_submit; anything else is copied subframe by subframe. */
int srsue_gpu_host_register(void *p, uint64_t bytes);
int srsue_gpu_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif

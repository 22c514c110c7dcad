/*
 * srslte_compat.h -- the slice of the srsLTE C API that srsUE's downlink path uses, served by
 * libsrsue_gpu.  A maintainer builds ue/src/phy/phch_worker.cc, ue/src/mac/dl_harq.cc and
 * ue/src/mac/proc_ra.cc against this header instead of "srslte/srslte.h" for the symbols below
 * (INTEGRATION.md shows the CMake change); names, argument meaning and return conventions are
 * those of the call sites cited next to each declaration (paths under /root/reference).
 *
 * Struct layouts are NOT binary compatible with a stock libsrslte build: callers are recompiled
 * against this header.  Fields the reference dereferences directly keep their names
 * (ue_dl.pdsch.dl_sch, ue_dl.sf_symbols, ue_dl.ce, ue_dl.chest, ue_dl.pdsch_cfg.grant.mcs.{mod,tbs,idx},
 * ue_dl.last_n_cce, ue_dl.last_location.{ncce,L}: phch_worker.cc:88,260,320,338,347-348,356-357,446).
 */
#ifndef SRSUE_GPU_SRSLTE_COMPAT_H
#define SRSUE_GPU_SRSLTE_COMPAT_H
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "srsue_gpu/srsue_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SRSLTE_API
#define SRSLTE_API __attribute__((visibility("default")))
#endif

#define SRSLTE_SUCCESS 0
#define SRSLTE_ERROR (-1)
#define SRSLTE_ERROR_INVALID_INPUTS (-2)

#define SRSLTE_MAX_PORTS 4
#define SRSLTE_MAX_PRB 110
#define SRSLTE_MAX_CODEBLOCKS 32
#define SRSLTE_TCOD_MAX_LEN_CB 6144
#define SRSLTE_TCOD_TOTALTAIL 12
#define SRSLTE_SIRNTI 0xFFFF
#define SRSLTE_PRNTI 0xFFFE
#define SRSLTE_RARNTI_START 0x0001
#define SRSLTE_RARNTI_END 0x003C
#define SRSLTE_CRNTI_START 0x003D
#define SRSLTE_CRNTI_END 0xFFF3

/* cf_t: srsLTE uses C99 `_Complex float`; {re, im} pairs have the same layout */
#ifdef SRSUE_GPU_PLAIN_CF
typedef srsue_gpu_cf_t cf_t;
#else
typedef _Complex float cf_t;
#endif

typedef enum { SRSLTE_CP_NORM = 0, SRSLTE_CP_EXT } srslte_cp_t;
typedef enum { SRSLTE_PHICH_NORM = 0, SRSLTE_PHICH_EXT } srslte_phich_length_t;
typedef enum { SRSLTE_PHICH_R_1_6 = 0, SRSLTE_PHICH_R_1_2, SRSLTE_PHICH_R_1, SRSLTE_PHICH_R_2 } srslte_phich_resources_t;
typedef enum { SRSLTE_MOD_BPSK = 0, SRSLTE_MOD_QPSK, SRSLTE_MOD_16QAM, SRSLTE_MOD_64QAM } srslte_mod_t;
typedef enum { SRSLTE_RNTI_USER = 0, SRSLTE_RNTI_SI, SRSLTE_RNTI_RAR, SRSLTE_RNTI_TEMP, SRSLTE_RNTI_SPS,
               SRSLTE_RNTI_PCH, SRSLTE_RNTI_NOF_TYPES } srslte_rnti_type_t;

typedef struct SRSLTE_API {
  uint32_t nof_prb;
  uint32_t nof_ports;
  uint32_t bw_idx;
  uint32_t id;
  srslte_cp_t cp;
  srslte_phich_length_t phich_length;
  srslte_phich_resources_t phich_resources;
} srslte_cell_t;

typedef struct SRSLTE_API { srslte_mod_t mod; int tbs; uint32_t idx; } srslte_ra_mcs_t;

typedef struct SRSLTE_API {
  bool prb_idx[2][SRSLTE_MAX_PRB];
  uint32_t nof_prb;
  uint32_t Qm;
  srslte_ra_mcs_t mcs;
} srslte_ra_dl_grant_t;

/* Unpacked downlink DCI (phch_worker.cc:288 `srslte_ra_dl_dci_t dci_unpacked`; its ndi, harq_process and rv_idx feed
 * the MAC grant at :303-307).  Allocation types follow 36.213 7.1.6.1-7.1.6.3. */
typedef enum { SRSLTE_RA_ALLOC_TYPE0 = 0, SRSLTE_RA_ALLOC_TYPE1, SRSLTE_RA_ALLOC_TYPE2 } srslte_ra_type_t;
typedef struct SRSLTE_API { uint32_t rbg_bitmask; } srslte_ra_type0_t;
typedef struct SRSLTE_API { uint32_t vrb_bitmask; uint32_t rbg_subset; bool shift; } srslte_ra_type1_t;
typedef enum { SRSLTE_RA_TYPE2_NPRB1A_2 = 0, SRSLTE_RA_TYPE2_NPRB1A_3 } srslte_ra_type2_nprb1a_t;
typedef enum { SRSLTE_RA_TYPE2_NG1 = 0, SRSLTE_RA_TYPE2_NG2 } srslte_ra_type2_ngap_t;
typedef enum { SRSLTE_RA_TYPE2_LOC = 0, SRSLTE_RA_TYPE2_DIST } srslte_ra_type2_mode_t;
typedef struct SRSLTE_API {
  uint32_t riv, L_crb, RB_start;
  srslte_ra_type2_nprb1a_t n_prb1a;
  srslte_ra_type2_ngap_t n_gap;
  srslte_ra_type2_mode_t mode;
} srslte_ra_type2_t;
typedef struct SRSLTE_API {
  srslte_ra_type_t alloc_type;
  srslte_ra_type0_t type0_alloc;
  srslte_ra_type1_t type1_alloc;
  srslte_ra_type2_t type2_alloc;
  uint32_t mcs_idx;
  uint32_t harq_process;
  int rv_idx;
  bool ndi;
  bool dci_is_1a;
  uint32_t tpc;
} srslte_ra_dl_dci_t;

/* uplink grant: only present so that srslte_phy_grant_t keeps its shape (mac_interface.h:59) */
typedef struct SRSLTE_API {
  uint32_t n_prb[2], n_prb_tilde[2], L_prb, freq_hopping, M_sc, M_sc_init, Qm;
  srslte_ra_mcs_t mcs;
  uint32_t ncs_dmrs;
} srslte_ra_ul_grant_t;

typedef union { srslte_ra_dl_grant_t dl; srslte_ra_ul_grant_t ul; } srslte_phy_grant_t;

typedef struct SRSLTE_API { uint32_t F, C, K1, K2, C1, C2, tbs; } srslte_cbsegm_t;
typedef struct SRSLTE_API { uint32_t lstart, nof_symb, nof_bits, nof_re; } srslte_ra_nbits_t;

typedef struct SRSLTE_API {
  srslte_cbsegm_t cb_segm;
  srslte_ra_dl_grant_t grant;
  srslte_ra_nbits_t nbits;
  uint32_t rv;
  uint32_t sf_idx;
} srslte_pdsch_cfg_t;

/* HARQ soft buffer, owned by MAC (ue/hdr/mac/dl_harq.h:88).  buffer_f keeps srsLTE's shape (one
 * int16 array of 3*6144+12 per code block, decoder-input order); the authoritative copy between a
 * reset and the next reset lives on the device (gpu_shadow) and is only mirrored into buffer_f by
 * srsue_gpu_softbuffer_rx_sync_host(). */
typedef struct SRSLTE_API {
  uint32_t max_cb;
  int16_t **buffer_f;
  void *gpu_shadow;
} srslte_softbuffer_rx_t;

typedef struct SRSLTE_API { uint32_t max_iterations; uint32_t nof_iterations; } srslte_sch_t;

typedef struct SRSLTE_API {
  srslte_cell_t cell;
  uint16_t rnti;
  srslte_sch_t dl_sch;
  void *gpu;                 /* back pointer to the owning ue_dl device state */
} srslte_pdsch_t;

typedef struct SRSLTE_API {
  srslte_cell_t cell;
  float noise_estimate, rsrp, rssi, rsrq;
  void *gpu;
} srslte_chest_dl_t;

typedef struct SRSLTE_API { void *gpu; } srslte_pdcch_t;      /* back pointer to the owning ue_dl device state */

/* DCI message as found on the PDCCH: payload bits (one per byte) without the CRC (phch_worker.cc:288,312) */
#define SRSLTE_DCI_MAX_BITS 128
typedef enum { SRSLTE_DCI_FORMAT0 = 0, SRSLTE_DCI_FORMAT1, SRSLTE_DCI_FORMAT1A, SRSLTE_DCI_FORMAT1C, SRSLTE_DCI_FORMAT_ERROR } srslte_dci_format_t;
typedef struct SRSLTE_API {
  uint8_t data[SRSLTE_DCI_MAX_BITS];
  uint32_t nof_bits;
  srslte_dci_format_t format;
} srslte_dci_msg_t;

typedef struct SRSLTE_API { uint32_t L; uint32_t ncce; } srslte_dci_location_t;

typedef struct SRSLTE_API {
  srslte_pdcch_t pdcch;
  srslte_pdsch_t pdsch;
  srslte_chest_dl_t chest;
  srslte_pdsch_cfg_t pdsch_cfg;
  srslte_softbuffer_rx_t softbuffer;      /* used by srslte_ue_dl_decode[_rnti] only */
  srslte_cell_t cell;
  cf_t *sf_symbols;                       /* host mirror, 14 * 12 * nof_prb */
  cf_t *ce[SRSLTE_MAX_PORTS];             /* host mirrors, one per configured port */
  uint16_t current_rnti;
  srslte_dci_location_t last_location;
  uint32_t last_n_cce;
  uint64_t pkt_errors, pkts_total, nof_detected;
  void *gpu;                              /* opaque device state */
} srslte_ue_dl_t;

typedef struct SRSLTE_API {
  uint32_t max_long_cb;
  uint32_t n_iter;                        /* iterations requested through srslte_tdec_iteration */
  void *gpu;
} srslte_tdec_t;

/* ---- helpers the DL callers use ------------------------------------------------------------------ */
SRSLTE_API int srslte_symbol_sz(uint32_t nof_prb);                       /* ue/src/phy/prach.cc:77 */
#define SRSLTE_SF_LEN(symbol_sz) ((symbol_sz) * 15)
#define SRSLTE_SF_LEN_PRB(nof_prb) (SRSLTE_SF_LEN(srslte_symbol_sz(nof_prb)))   /* phch_worker.cc:69 */
SRSLTE_API void *srslte_vec_malloc(uint32_t size);                      /* phch_worker.cc:69 (pinned here) */
SRSLTE_API void srslte_vec_free(void *p);

/* ---- UE DL object (phch_worker.cc:74,104,127,254,337) -------------------------------------------- */
SRSLTE_API int srslte_ue_dl_init(srslte_ue_dl_t *q, srslte_cell_t cell);
SRSLTE_API void srslte_ue_dl_free(srslte_ue_dl_t *q);
SRSLTE_API void srslte_ue_dl_set_rnti(srslte_ue_dl_t *q, uint16_t rnti);
SRSLTE_API int srslte_ue_dl_decode_fft_estimate(srslte_ue_dl_t *q, cf_t *input, uint32_t sf_idx, uint32_t *cfi);
SRSLTE_API int srslte_ue_dl_cfg_grant(srslte_ue_dl_t *q, srslte_ra_dl_grant_t *grant, uint32_t cfi, uint32_t sf_idx,
                                      uint32_t rvidx);
/* PDCCH (phch_worker.cc:260,293,320): soft bits of the whole control region on the device, then a blind search.
 * find_dl_dci_type returns 1 when a DCI for rnti was found (UE-specific space: formats 1A and 1, then the common space:
 * format 1A; SI/RA/P-RNTI: common space only), 0 when not, < 0 on error; it fills dci_msg, q->last_location and
 * q->last_n_cce.  srslte_dci_msg_to_dl_grant (below) turns the message into a grant. */
SRSLTE_API int srslte_pdcch_extract_llr(srslte_pdcch_t *q, cf_t *sf_symbols, cf_t *ce[SRSLTE_MAX_PORTS], float noise_estimate,
                                        uint32_t nsubframe, uint32_t cfi);
SRSLTE_API int srslte_ue_dl_find_dl_dci_type(srslte_ue_dl_t *q, srslte_dci_msg_t *dci_msg, uint32_t cfi, uint32_t sf_idx,
                                             uint16_t rnti, srslte_rnti_type_t rnti_type);
/* uplink grant search (phch_worker.cc:426): DCI format 0 in the UE-specific space; 1 found / 0 / < 0 */
/* DCI payload -> unpacked fields -> grant (phch_worker.cc:297).  Formats 1A and 1, FDD; allocation types 0, 1 and 2,
 * localized and distributed (36.211 6.2.3.2: the grant's two slot masks then differ).  Transport-block sizes come from the 27 x 110 table of 36.213
 * 7.1.7.2.1: the caller installs it once per process with srsue_gpu_ra_set_tbs_table(); until then the thirteen built-in
 * columns serve (srsue_gpu_ra_builtin_tbs_columns) and a call that needs any other size returns SRSLTE_ERROR with a message.  Returns 0 on success (as srsLTE does). */
SRSLTE_API int srsue_gpu_ra_set_tbs_table(const int32_t *table, uint32_t nof_rows /* 27 */, uint32_t nof_cols /* 110 */);
SRSLTE_API int srsue_gpu_ra_have_tbs_table(void);
/* columns of the table this library carries itself (N_PRB = 1..6, 10, 15, 25, 50, 75, 100, 110; tbs_table.inc, written from
 * memory and checked structurally) -- used until the whole table is installed; returns their number */
SRSLTE_API int srsue_gpu_ra_builtin_tbs_columns(int32_t *n_prb, int cap);
SRSLTE_API int srslte_dci_msg_to_dl_grant(srslte_dci_msg_t *msg, uint16_t msg_rnti, uint32_t nof_prb, srslte_ra_dl_dci_t *dl_dci,
                                          srslte_ra_dl_grant_t *grant);
SRSLTE_API int srslte_dci_msg_unpack_pdsch(srslte_dci_msg_t *msg, srslte_ra_dl_dci_t *data, uint32_t nof_prb, bool crc_is_crnti);
SRSLTE_API int srslte_dci_msg_pack_pdsch(srslte_ra_dl_dci_t *data, srslte_dci_format_t format, srslte_dci_msg_t *msg, uint32_t nof_prb,
                                        bool crc_is_crnti);
SRSLTE_API int srslte_ra_dl_dci_to_grant(srslte_ra_dl_dci_t *dci, uint32_t nof_prb, bool crc_is_crnti, srslte_ra_dl_grant_t *grant);
SRSLTE_API int srslte_ra_dl_dci_to_grant_prb_allocation(srslte_ra_dl_dci_t *dci, srslte_ra_dl_grant_t *grant, uint32_t nof_prb);
SRSLTE_API int srslte_ra_tbs_idx_from_mcs(uint32_t mcs_idx);
SRSLTE_API srslte_mod_t srslte_ra_mod_from_mcs(uint32_t mcs_idx);
SRSLTE_API int srslte_ra_tbs_from_idx(uint32_t tbs_idx, uint32_t n_prb);
SRSLTE_API uint32_t srslte_ra_type0_P(uint32_t nof_prb);
SRSLTE_API uint32_t srslte_ra_type2_n_rb(uint32_t nof_prb);
SRSLTE_API uint32_t srslte_ra_type2_to_riv(uint32_t L_crb, uint32_t RB_start, uint32_t nof_prb);
SRSLTE_API void srslte_ra_type2_from_riv(uint32_t riv, uint32_t *L_crb, uint32_t *RB_start, uint32_t nof_prb, uint32_t nof_vrb);
/* CQI reporting (phch_worker.cc:504-527): chest SNR -> CQI index, report timing, UCI bit packing */
typedef enum { SRSLTE_CQI_TYPE_WIDEBAND = 0, SRSLTE_CQI_TYPE_SUBBAND } srslte_cqi_type_t;
typedef struct SRSLTE_API { uint8_t wideband_cqi; } srslte_cqi_wideband_t;
typedef struct SRSLTE_API { uint8_t subband_cqi; uint8_t subband_label; } srslte_cqi_subband_t;
typedef struct SRSLTE_API {
  srslte_cqi_type_t type;
  srslte_cqi_wideband_t wideband;
  srslte_cqi_subband_t subband;
} srslte_cqi_value_t;
#define SRSLTE_CQI_MAX_BITS 64
SRSLTE_API uint8_t srslte_cqi_from_snr(float snr_db);
SRSLTE_API bool srslte_cqi_send(uint32_t I_cqi_pmi, uint32_t tti);
SRSLTE_API int srslte_cqi_value_pack(srslte_cqi_value_t *value, uint8_t *buff);
#define SRSLTE_VEC_EMA(data, average, alpha) ((alpha) * (data) + (1 - (alpha)) * (average))     /* phch_worker.cc:512 */
/* distributed virtual -> physical resource block of one slot (0xFFFFFFFF when out of range) and the number of VRBs */
SRSLTE_API uint32_t srsue_gpu_host_dvrb_to_prb(uint32_t nof_prb, int gap2, uint32_t n_vrb, int slot);
SRSLTE_API uint32_t srsue_gpu_host_n_vrb_dl(uint32_t nof_prb, int gap2);
SRSLTE_API char *srslte_ra_dl_dci_string(srslte_ra_dl_dci_t *dci);                       /* phch_worker.cc:317 */
SRSLTE_API int srslte_ue_dl_find_ul_dci(srslte_ue_dl_t *q, srslte_dci_msg_t *dci_msg, uint32_t cfi, uint32_t sf_idx, uint16_t rnti);
SRSLTE_API uint32_t srslte_ue_dl_get_ncce(srslte_ue_dl_t *q);
/* HARQ indicator of the uplink transmission (phch_worker.cc:381); needs srslte_ue_dl_decode_fft_estimate of this subframe */
SRSLTE_API bool srslte_ue_dl_decode_phich(srslte_ue_dl_t *q, uint32_t sf_idx, uint32_t n_prb_lowest, uint32_t n_dmrs);
/* wrappers the north star names; the DCI search is a "next" row, so the grant to use is the one last
 * installed with srsue_gpu_ue_dl_set_grant().  Return decoded bits (tbs) / 0 (no grant) / < 0. */
SRSLTE_API int srslte_ue_dl_decode(srslte_ue_dl_t *q, cf_t *input, uint8_t *data, uint32_t tti);
SRSLTE_API int srslte_ue_dl_decode_rnti(srslte_ue_dl_t *q, cf_t *input, uint8_t *data, uint32_t tti, uint16_t rnti);

/* ---- PDSCH / SCH (phch_worker.cc:88,347-348,360,848) -------------------------------------------- */
SRSLTE_API int srslte_pdsch_decode_rnti(srslte_pdsch_t *q, srslte_pdsch_cfg_t *cfg, srslte_softbuffer_rx_t *softbuffer,
                                        cf_t *sf_symbols, cf_t *ce[SRSLTE_MAX_PORTS], float noise_estimate, uint16_t rnti,
                                        uint8_t *data);
SRSLTE_API void srslte_sch_set_max_noi(srslte_sch_t *q, uint32_t max_iterations);
SRSLTE_API uint32_t srslte_pdsch_last_noi(srslte_pdsch_t *q);

/* ---- channel-estimate getters (phch_worker.cc:359,799,821-823,842,847) --------------------------- */
SRSLTE_API float srslte_chest_dl_get_noise_estimate(srslte_chest_dl_t *q);
SRSLTE_API float srslte_chest_dl_get_snr(srslte_chest_dl_t *q);
SRSLTE_API float srslte_chest_dl_get_rssi(srslte_chest_dl_t *q);
SRSLTE_API float srslte_chest_dl_get_rsrq(srslte_chest_dl_t *q);
SRSLTE_API float srslte_chest_dl_get_rsrp(srslte_chest_dl_t *q);

/* ---- soft buffer (dl_harq.cc:169,174,232; proc_ra.cc:60,254; ue_itf_test_sib1.cc:121,136) -------- */
SRSLTE_API int srslte_softbuffer_rx_init(srslte_softbuffer_rx_t *q, uint32_t nof_prb);
SRSLTE_API void srslte_softbuffer_rx_free(srslte_softbuffer_rx_t *q);
SRSLTE_API void srslte_softbuffer_rx_reset(srslte_softbuffer_rx_t *q);
SRSLTE_API void srslte_softbuffer_rx_reset_tbs(srslte_softbuffer_rx_t *q, uint32_t tbs);
SRSLTE_API void srslte_softbuffer_rx_reset_cb(srslte_softbuffer_rx_t *q, uint32_t nof_cb);

/* ---- turbo decoder object (BASELINE config 4) ----------------------------------------------------- */
SRSLTE_API int srslte_tdec_init(srslte_tdec_t *h, uint32_t max_long_cb);
SRSLTE_API void srslte_tdec_free(srslte_tdec_t *h);
SRSLTE_API int srslte_tdec_reset(srslte_tdec_t *h, uint32_t long_cb);
SRSLTE_API void srslte_tdec_iteration(srslte_tdec_t *h, int16_t *input, uint32_t long_cb);
SRSLTE_API void srslte_tdec_decision(srslte_tdec_t *h, uint8_t *output, uint32_t long_cb);       /* one bit per byte */
SRSLTE_API void srslte_tdec_decision_byte(srslte_tdec_t *h, uint8_t *output, uint32_t long_cb);  /* packed, MSB first */
SRSLTE_API int srslte_tdec_run_all(srslte_tdec_t *h, int16_t *input, uint8_t *output, uint32_t nof_iterations,
                                   uint32_t long_cb);

/* ---- cell search (ue/src/phy/phch_recv.cc:140-188) --------------------------------------------------- */
typedef struct SRSLTE_API { uint32_t full_secs; double frac_secs; } srslte_timestamp_t;
/* Automatic gain control of the synchroniser objects (phch_recv.cc:111,152,202 start it; :174,212 read the gain back).
 * gain is what the radio callback was last given / answered, in dB as phch_recv.cc:87-90 passes it to
 * radio::set_rx_gain_th.  Every `period` received blocks (0: every block) the mean sample power of the block is
 * compared with `target` and the gain moved by `bandwidth` of the error in dB, inside [0, max_gain]. */
typedef struct SRSLTE_API {
  double gain;
  double (*set_gain_callback)(void *, double);
  void *handler;
  float target, bandwidth, max_gain, last_power;
  uint32_t period, count, nof_updates;
} srslte_agc_t;
typedef struct SRSLTE_API { float threshold; float em_alpha; } srslte_sync_t;
/* subframe synchroniser (phch_recv.cc:100-113,236-258,302-334): what the callers dereference are .agc and .strack */
typedef struct SRSLTE_API {
  srslte_agc_t agc;
  srslte_sync_t strack;
  void *gpu;
} srslte_ue_sync_t;
typedef struct SRSLTE_API {
  uint32_t cell_id;
  srslte_cp_t cp;
  float peak;          /* mean peak-to-side ratio of the frames that agreed on this cell */
  float mode;          /* fraction of the scanned frames that agreed */
  float psr;
  float cfo;           /* Hz */
} srslte_ue_cellsearch_result_t;
typedef struct SRSLTE_API {
  srslte_ue_sync_t ue_sync;
  uint32_t nof_frames_to_scan;
  float detect_threshold;
  void *gpu;
} srslte_ue_cellsearch_t;
/* recv_callback(handler, buffer, nsamples, rx_time) delivers nsamples cf_t at 1.92 Msps (radio_recv_wrapper_cs,
 * phch_recv.cc:73): the scan pulls nof_frames_to_scan x 5 ms through it, searches them in one batch on the device and
 * returns the cell most frames agree on.  scan: > 0 found (found_cells[*max_N_id_2] filled), 0 nothing, < 0 error. */
SRSLTE_API int srslte_ue_cellsearch_init(srslte_ue_cellsearch_t *q,
                                         int (*recv_callback)(void *, void *, uint32_t, srslte_timestamp_t *), void *stream_handler);
SRSLTE_API void srslte_ue_cellsearch_free(srslte_ue_cellsearch_t *q);
SRSLTE_API int srslte_ue_cellsearch_set_nof_frames_to_scan(srslte_ue_cellsearch_t *q, uint32_t nof_frames);
SRSLTE_API void srslte_ue_cellsearch_set_threshold(srslte_ue_cellsearch_t *q, float threshold);
SRSLTE_API int srslte_ue_cellsearch_scan(srslte_ue_cellsearch_t *q, srslte_ue_cellsearch_result_t found_cells[3], uint32_t *max_N_id_2);
SRSLTE_API int srslte_ue_cellsearch_scan_N_id_2(srslte_ue_cellsearch_t *q, uint32_t N_id_2, srslte_ue_cellsearch_result_t *found_cell);
SRSLTE_API int srslte_ue_sync_start_agc(srslte_ue_sync_t *q, double (*set_gain_callback)(void *, double), float init_gain_value);
/* Subframe synchronisation at the cell's own sampling rate.  Every call pulls samples through the radio callback.  While
 * searching (a 5 ms PSS/SSS search for the cell's id) the calls return 0; once aligned every call returns 1 with exactly
 * one subframe in the buffer, checks the PSS of subframes 0 and 5 in a window around its expected place, follows timing
 * drift by reading a few samples more or less, averages the carrier offset, and falls back to searching after ten
 * consecutive missed peaks.  < 0 on error.  The samples are handed out as received: where srsLTE rotates them on the
 * CPU by the tracked offset, here the worker's transform does it on its loads -- pass srslte_ue_sync_get_cfo()/15000
 * to srsue_gpu_ue_dl_set_cfo(), which is what phch_recv.cc:328-329 already hands to phch_worker::set_cfo. */
SRSLTE_API int srslte_ue_sync_init(srslte_ue_sync_t *q, srslte_cell_t cell,
                                   int (*recv_callback)(void *, void *, uint32_t, srslte_timestamp_t *), void *stream_handler);
SRSLTE_API void srslte_ue_sync_free(srslte_ue_sync_t *q);
SRSLTE_API int srslte_ue_sync_zerocopy(srslte_ue_sync_t *q, cf_t *input_buffer);            /* phch_recv.cc:322 */
SRSLTE_API int srslte_ue_sync_get_buffer(srslte_ue_sync_t *q, cf_t **sf_symbols);           /* phch_recv.cc:237 */
SRSLTE_API uint32_t srslte_ue_sync_get_sfidx(srslte_ue_sync_t *q);
SRSLTE_API float srslte_ue_sync_get_cfo(srslte_ue_sync_t *q);                               /* Hz */
SRSLTE_API float srslte_ue_sync_get_sfo(srslte_ue_sync_t *q);                               /* Hz, from the timing corrections */
SRSLTE_API void srslte_ue_sync_set_cfo(srslte_ue_sync_t *q, float cfo);
SRSLTE_API void srslte_ue_sync_set_agc_period(srslte_ue_sync_t *q, uint32_t period);
SRSLTE_API void srslte_ue_sync_decode_sss_on_track(srslte_ue_sync_t *q, bool enabled);
SRSLTE_API void srslte_ue_sync_get_last_timestamp(srslte_ue_sync_t *q, srslte_timestamp_t *timestamp);
SRSLTE_API void srslte_sync_set_threshold(srslte_sync_t *q, float threshold);
SRSLTE_API void srslte_sync_set_em_alpha(srslte_sync_t *q, float alpha);
SRSLTE_API float srslte_agc_get_gain(srslte_agc_t *q);

/* ---- MIB decode (ue/src/phy/phch_recv.cc:98,246-253) ------------------------------------------------- */
#define SRSLTE_BCH_PAYLOAD_LEN 24
#define SRSLTE_UE_MIB_FOUND 1
#define SRSLTE_UE_MIB_NOTFOUND 0
typedef struct SRSLTE_API { void *gpu; } srslte_pbch_t;
typedef struct SRSLTE_API {
  srslte_pbch_t pbch;
  srslte_cell_t cell;
  void *gpu;
} srslte_ue_mib_t;
SRSLTE_API int srslte_ue_mib_init(srslte_ue_mib_t *q, srslte_cell_t cell);            /* 0 = ok (phch_recv.cc:98) */
SRSLTE_API void srslte_ue_mib_free(srslte_ue_mib_t *q);
SRSLTE_API void srslte_pbch_decode_reset(srslte_pbch_t *q);                            /* phch_recv.cc:246 */
/* input: the samples of a subframe 0.  Returns SRSLTE_UE_MIB_FOUND and the 24 MIB bits (one per byte), the number of
 * transmit ports and the frame's position in the 40 ms BCH period; every call decodes from this subframe alone */
SRSLTE_API int srslte_ue_mib_decode(srslte_ue_mib_t *q, cf_t *input, uint8_t bch_payload[SRSLTE_BCH_PAYLOAD_LEN],
                                    uint32_t *nof_tx_ports, uint32_t *sfn_offset);
/* MIB search on the air interface at 1.92 Msps (phch_recv.cc:194-213): pulls 5 ms frames through the radio callback, finds
 * the PSS/SSS of cell_id, and decodes the PBCH of the subframes 0 it sees; returns 1 (found), 0 (not within
 * max_frames_timeout frames) or < 0 */
typedef struct SRSLTE_API {
  srslte_ue_sync_t ue_sync;
  uint32_t cell_id;
  void *gpu;
} srslte_ue_mib_sync_t;
SRSLTE_API int srslte_ue_mib_sync_init(srslte_ue_mib_sync_t *q, uint32_t cell_id, srslte_cp_t cp,
                                       int (*recv_callback)(void *, void *, uint32_t, srslte_timestamp_t *), void *stream_handler);
SRSLTE_API void srslte_ue_mib_sync_free(srslte_ue_mib_sync_t *q);
SRSLTE_API int srslte_ue_mib_sync_decode(srslte_ue_mib_sync_t *q, uint32_t max_frames_timeout,
                                         uint8_t bch_payload[SRSLTE_BCH_PAYLOAD_LEN], uint32_t *nof_tx_ports, uint32_t *sfn_offset);
SRSLTE_API void srslte_pbch_mib_unpack(uint8_t *msg, srslte_cell_t *cell, uint32_t *sfn);   /* phch_recv.cc:216,253 */
SRSLTE_API void srslte_pbch_mib_pack(srslte_cell_t *cell, uint32_t sfn, uint8_t *msg);

/* ---- small utilities the DL callers use (phch_recv.cc:192,218,220,275,335-338; phch_worker.cc:305) ---- */
SRSLTE_API void srslte_bit_pack_vector(uint8_t *unpacked, uint8_t *packed, int nof_bits);      /* bits (one per byte) -> bytes, MSB first */
SRSLTE_API void srslte_bit_unpack_vector(uint8_t *packed, uint8_t *unpacked, int nof_bits);
SRSLTE_API uint32_t srslte_bit_pack(uint8_t **bits, int nof_bits);                           /* reads nof_bits, advances *bits */
SRSLTE_API void srslte_bit_unpack(uint32_t value, uint8_t **bits, int nof_bits);
SRSLTE_API const char *srslte_cp_string(srslte_cp_t cp);
SRSLTE_API void srslte_cell_fprint(FILE *stream, srslte_cell_t *cell, uint32_t sfn);
SRSLTE_API int srslte_sampling_freq_hz(uint32_t nof_prb);
SRSLTE_API void srslte_timestamp_copy(srslte_timestamp_t *dest, srslte_timestamp_t *src);
SRSLTE_API int srslte_timestamp_add(srslte_timestamp_t *t, uint32_t full_secs, double frac_secs);
SRSLTE_API uint32_t srslte_tti_interval(uint32_t tti1, uint32_t tti2);                       /* (tti1 - tti2) mod 10240 */

/* ---- extensions (not in srsLTE) ------------------------------------------------------------------- */
/* srslte_ue_dl_decode_fft_estimate decodes the PCFICH on the device and returns its CFI; set_cfi(1..3) forces a
 * value instead (captures without a control region), set_cfi(0) returns to decoding.  The grant still comes from
 * the caller until PDCCH blind decoding lands on the device. */
SRSLTE_API void srsue_gpu_ue_dl_set_cfi(srslte_ue_dl_t *q, uint32_t cfi);
/* cfi 0: use the CFI decoded from the PCFICH of each subframe */
/* carrier-offset correction of the downlink samples, in subcarrier spacings (the value phch_worker::set_cfo gets at
 * phch_recv.cc:329); applied inside srslte_ue_dl_decode_fft_estimate from the next subframe on.  0: off. */
SRSLTE_API int srsue_gpu_ue_dl_set_cfo(srslte_ue_dl_t *q, float cfo);
SRSLTE_API int srsue_gpu_ue_dl_set_grant(srslte_ue_dl_t *q, const srslte_ra_dl_grant_t *grant, uint32_t cfi, uint32_t rvidx);
/* copy the device-resident soft buffer into buffer_f (srsLTE decoder-input order) for `tbs` bits */
SRSLTE_API int srsue_gpu_softbuffer_rx_sync_host(srslte_softbuffer_rx_t *q, uint32_t tbs);

#ifdef __cplusplus
}
#endif
#endif

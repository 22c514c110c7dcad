/*
 * lte_oracle.h -- CPU ORACLE for the srsUE downlink PDSCH receive chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * executed from the product (libsrsue_gpu / srsue_b200).  Only tests/, the smoke
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in srsLTE (github.com/srsLTE/srsLTE,
 * pinned only as ">= 1.0.0" by /root/reference/ue/hdr/srslte_version_check.h:30-41 and
 * /root/reference/CMakeLists.txt:101), which is neither vendored in the reference tree nor
 * installable offline, and the reference holds no golden vector for this path
 * (ue/test/phy/CMakeLists.txt:20-24).  This file restates the chain from 3GPP TS 36.211 /
 * 36.212 (Rel-8) plus the srsLTE-flavoured fixed-point conventions frozen in oracle/SPEC.md;
 * behaviour at the API boundary follows the reference call sites
 * (ue/src/phy/phch_worker.cc:246-374, ue/src/mac/dl_harq.cc:191-259).
 */
#ifndef LTE_ORACLE_H
#define LTE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } lteo_cf_t;
typedef struct { double re, im; } lteo_cd_t;

#define LTEO_CRC24A 0x1864CFBu
#define LTEO_CRC24B 0x1800063u
#define LTEO_CRC16  0x11021u

/* fixed-point turbo decoder constants (SPEC.md section 7) */
#define LTEO_TD_C     511     /* clamp on channel LLRs at the decoder input          */
#define LTEO_TD_E     2047    /* clamp on the extrinsic / a-priori LLR              */
#define LTEO_TD_INF   10000   /* finite "minus infinity" for known trellis states   */
#define LTEO_TD_NORM  4       /* state metrics re-normalised when k % 4 == 0        */
#define LTEO_MAX_K    6144
#define LTEO_LLR_MAX  32767   /* symmetric int16 saturation of demapper LLRs           */
#define LTEO_SB_MAX   LTEO_TD_C /* the soft buffer IS the decoder input: it saturates at +-C */
#define LTEO_FILLER_LLR (-LTEO_TD_C)

typedef struct {
  int nof_prb;      /* 6,15,25,50,75,100 */
  int nof_ports;    /* 1, 2 or 4 */
  int cell_id;      /* 0..503 */
  int cp;           /* 0 = normal cyclic prefix (7 symbols per slot), 1 = extended (6 symbols per slot, SPEC.md 15b) */
} lteo_cell_t;
/* Grids (sf_symbols, ce, TX grids) always have a stride of 14 symbols per port; with the extended cyclic prefix only
 * rows 0..11 are used. */
/* Transmit diversity (36.211 6.3.4.3): Alamouti pair number `pair` of a mapped sequence goes out on ports (0, 1) of a
 * two-port cell; with four ports even pairs use ports (0, 2) and odd pairs ports (1, 3) (SFBC-FSTD).  SPEC.md 15c. */
#define LTEO_DIV_PA(np, pair) ((np) == 4 ? ((pair) & 1) : 0)
#define LTEO_DIV_PB(np, pair) ((np) == 4 ? 2 + ((pair) & 1) : 1)
#define LTEO_NSYMB(cp) ((cp) ? 12 : 14)
#define LTEO_NSLOT(cp) ((cp) ? 6 : 7)

typedef struct {
  int sf_idx;       /* 0..9 */
  int cfi;          /* 1..3 */
  int rnti;
  int qm;           /* 2,4,6 */
  int tbs;          /* transport block size in bits */
  int rv;           /* 0..3 */
  int tm;           /* 1 = single port, 2 = transmit diversity (needs nof_ports == 2) */
  int nof_prb_alloc;
  uint8_t prb_mask[110]; /* 1 = PRB allocated in both slots; 2 = slot 0 only; 4 = slot 1 only */
} lteo_pdsch_cfg_t;

typedef struct { int tbs, B, C, Kp, Km, Cp, Cm, F; } lteo_cbsegm_t;

/* ---- tables / small helpers ------------------------------------------------------------ */
int      lteo_qpp_params(int K, int *f1, int *f2);         /* 0 = ok, -1 = K not in table      */
int      lteo_qpp_table_size(void);
int      lteo_qpp_K(int idx);
void     lteo_qpp_perm(int K, uint16_t *pi);                 /* pi[i] = (f1 i + f2 i^2) mod K   */
int      lteo_window_len(int K);                             /* W(K), SPEC.md 7.3               */
int      lteo_symbol_sz(int nof_prb);
int      lteo_cp_len(int nfft, int symbol_in_slot);
int      lteo_cp_len_x(int nfft, int symbol_in_slot, int cp);   /* cp = 1: 512 samples at 2048 for every symbol */
uint32_t lteo_crc_bits(const uint8_t *bits, int n, uint32_t poly, int order);
void     lteo_gold(uint32_t c_init, int n, uint8_t *c);
int      lteo_cbsegm(int tbs, lteo_cbsegm_t *s);
int      lteo_cb_len(const lteo_cbsegm_t *s, int r);
int      lteo_cb_E(const lteo_cbsegm_t *s, int G, int qm, int nl, int r);
int      lteo_pdsch_re_list(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, int32_t *re_idx);
int      lteo_crs_positions(const lteo_cell_t *cell, int port, int l, int32_t *k_out);
void     lteo_crs_values(const lteo_cell_t *cell, int sf_idx, int l, int8_t *re_sign, int8_t *im_sign);
void     lteo_pcfich_re(const lteo_cell_t *cell, int32_t *k16);    /* subcarriers of d(0..15) in symbol 0 */
void     lteo_pcfich_bits(const lteo_cell_t *cell, int sf_idx, int cfi, uint8_t *b32);
void     lteo_fft_twiddles(int n, lteo_cf_t *tw);           /* tw[k] = exp(-2 pi i k / n), k < n/2 */

/* ---- TX side (test-vector generator; double precision) ---------------------------------- */
void lteo_turbo_encode(const uint8_t *c, int K, uint8_t *d);            /* d: 3*(K+4) triples   */
int  lteo_rm_sequence(int K, int F, int rv, int32_t *seq);              /* returns N_nd         */
int  lteo_rm_tx(const uint8_t *d, int K, int F, int E, int rv, uint8_t *e);
int  lteo_pdsch_encode_bits(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg,
                            const uint8_t *tb_bytes, uint8_t *e_bits, int *G_out);
/* uplink UL-SCH coding without control information + PUSCH scrambling; returns G or < 0 (lteo_tx.c) */
int  lteo_ulsch_encode(int tbs, int qm, int nof_prb, int n_symb, int rv, int rnti, int sf_idx, int cell_id,
                       const uint8_t *tb_bytes, uint8_t *out_bits /* 12 nof_prb n_symb qm */);
int  lteo_pdsch_tx_grid(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg,
                        const uint8_t *tb_bytes, lteo_cd_t *grid /* [ports][14][12*nof_prb] */);
void lteo_ofdm_tx(int nof_prb, const lteo_cd_t *grid, lteo_cd_t *iq /* 15*nfft samples */);
void lteo_ofdm_tx_cp(int nof_prb, int cp, const lteo_cd_t *grid, lteo_cd_t *iq);
/* adds the PCFICH of `cfi` (1..3) to a grid built by lteo_pdsch_tx_grid (symbol 0; TX diversity with 2 ports) */
void lteo_pcfich_tx(const lteo_cell_t *cell, int sf_idx, int cfi, lteo_cd_t *grid);

/* ---- RX side (the restated hot path) ---------------------------------------------------- */
void lteo_fft(const lteo_cf_t *in, lteo_cf_t *out, int n);
void lteo_ofdm_rx(int nof_prb, const lteo_cf_t *iq, lteo_cf_t *sf_symbols);
void lteo_ofdm_rx_cp(int nof_prb, int cp, const lteo_cf_t *iq, lteo_cf_t *sf_symbols);
/* meas[5] = noise, rsrp, rssi, rsrq, snr */
void lteo_chest(const lteo_cell_t *cell, int sf_idx, const lteo_cf_t *sf_symbols,
                lteo_cf_t *ce /* [ports][14*nsc] */, float *meas);
void lteo_equalize(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const lteo_cf_t *sf_symbols,
                   const lteo_cf_t *ce, float noise_est, lteo_cf_t *d, int *nof_re);
void lteo_demod(const lteo_cf_t *d, int nof_re, int qm, int16_t *llr);
void lteo_descramble(int16_t *llr, int n, uint32_t c_init);
void lteo_rm_rx(const int16_t *e, int E, int K, int F, int rv, int16_t *w /* 3K+12 triples */);
/* crc_type: 0 none, 1 CRC24A, 2 CRC24B.  returns iterations run.  bits: K hard bits.        */
/* ---- PDCCH (SPEC.md 10; lteo_pdcch.c) ---- */
typedef struct { const uint8_t *bits; int nof_bits; uint16_t rnti; int L, ncce; } lteo_dci_tx_t;
int  lteo_ctrl_symbols(int nof_prb, int cfi);
int  lteo_phich_groups(int nof_prb, int ng_x6);                         /* ng_x6 = 6 Ng: 1, 3, 6, 12 */
int  lteo_pdcch_regs(const lteo_cell_t *cell, int cfi, int ng_x6, int32_t *reg_k, int32_t *reg_l);
void lteo_reg_res(const lteo_cell_t *cell, int k0, int l, int32_t *k4);
int  lteo_cc_interleaver(int D, int32_t *out);
void lteo_pdcch_quad_perm(int n_quad, int cell_id, int32_t *src);
void lteo_conv_encode(const uint8_t *c, int D, uint8_t *d);
int  lteo_cc_rm_sequence(int D, int32_t *seq);
void lteo_dci_encode(const uint8_t *bits, int nof_bits, uint16_t rnti, int E, uint8_t *e);
int  lteo_pdcch_search_space(int nof_cce, int sf_idx, uint16_t rnti, int common, int32_t *cand_L, int32_t *cand_ncce);
/* adds the PDCCHs of `list` to a grid built by lteo_pdsch_tx_grid; returns the number of CCEs or < 0 */
int  lteo_pdcch_tx(const lteo_cell_t *cell, int sf_idx, int cfi, int ng_x6, int n_dci, const lteo_dci_tx_t *list,
                   lteo_cd_t *grid);
int  lteo_pdcch_extract_llr(const lteo_cell_t *cell, int sf_idx, int cfi, int ng_x6, const lteo_cf_t *sf_symbols,
                            const lteo_cf_t *ce, float noise_est, int16_t *llr /* 8 * n_reg */);
uint16_t lteo_pdcch_decode_candidate(const int16_t *llr, int L, int nof_bits, uint8_t *bits_out);
int  lteo_pdcch_find_dci(const int16_t *llr, int nof_cce, int sf_idx, uint16_t rnti, int common, int nof_bits,
                         uint8_t *bits_out, int *found_L, int *found_ncce);
/* ---- cell search: PSS / SSS (SPEC.md 13; lteo_sync.c) ---- */
void  lteo_pss_seq(int n_id_2, lteo_cd_t *d62);
void  lteo_sss_seq(int n_id_1, int n_id_2, int sf5, int8_t *d62);
void  lteo_sync_tx(const lteo_cell_t *cell, int sf_idx, lteo_cd_t *grid);
void  lteo_pss_time(int n_id_2, lteo_cf_t *t128);
float lteo_pss_search(const lteo_cf_t *x, int n_samples, int *peak_pos, int *n_id_2, float *cfo, float *mean_power);
int   lteo_sss_detect(const lteo_cf_t *x, int peak_pos, int n_id_2, int *sf5, float *corr);
void  lteo_pss_time_n(int n_id_2, int nfft, lteo_cf_t *t);
float lteo_pss_search_n(const lteo_cf_t *x, int n_samples, int nfft, int force_n_id_2, int first_pos, int *peak_pos, int *n_id_2,
                        float *cfo, float *mean_power);
int   lteo_sss_detect_n(const lteo_cf_t *x, int peak_pos, int n_id_2, int nfft, int *sf5, float *corr);
/* cp_mode 0: normal prefix, 1: extended, 2: try both and report the better one in *cp (SPEC.md 15b.7) */
int   lteo_sss_detect_cp(const lteo_cf_t *x, int peak_pos, int n_id_2, int nfft, int cp_mode, int *sf5, float *corr, int *cp);
/* CFO correction (SPEC.md 14): y[n] = x[n] * T[(n * step mod 2^32) >> 20], T = 4096-entry unit circle */
#define LTEO_CFO_TABLE_LOG2 12
#define LTEO_CFO_TABLE (1 << LTEO_CFO_TABLE_LOG2)
void  lteo_cfo_table(lteo_cf_t *tab);
int32_t lteo_cfo_step(float cfo, int nfft);
void  lteo_cfo_correct(const lteo_cf_t *in, lteo_cf_t *out, int n_samples, int32_t step);
/* ---- PBCH / MIB (SPEC.md 12) ---- */
uint16_t lteo_viterbi_crc16(const int32_t *soft, int nof_bits, uint8_t *bits_out);
void lteo_pbch_res(const lteo_cell_t *cell, int32_t *g240);
int  lteo_pbch_res_n(const lteo_cell_t *cell, int32_t *g240);           /* returns 240, or 216 with the extended prefix */
void lteo_pbch_tx(const lteo_cell_t *cell, const uint8_t *mib24, int frame_idx, lteo_cd_t *grid);
void lteo_pbch_llr(const lteo_cell_t *cell, int hyp_ports, const lteo_cf_t *sf_symbols, const lteo_cf_t *ce, float noise_est,
                   int16_t *llr480);
int  lteo_pbch_decode(const lteo_cell_t *cell, const lteo_cf_t *sf_symbols, const lteo_cf_t *ce, float noise_est,
                      uint8_t *mib24, int *nof_ports, int *sfn_offset);
/* ---- PHICH (SPEC.md 11) ---- */
void lteo_phich_res(const lteo_cell_t *cell, int ng_x6, int n_group, int32_t *k12);
void lteo_phich_index(int nof_prb, int ng_x6, int I_lowest, int n_dmrs, int *n_group, int *n_seq);
void lteo_phich_index_cp(int nof_prb, int ng_x6, int cp, int I_lowest, int n_dmrs, int *n_group, int *n_seq);
void lteo_phich_tx(const lteo_cell_t *cell, int sf_idx, int ng_x6, int n_group, int n_seq, int ack, lteo_cd_t *grid);
int  lteo_phich_decode(const lteo_cell_t *cell, int sf_idx, int ng_x6, const lteo_cf_t *sf_symbols, const lteo_cf_t *ce,
                       float noise_est, int n_group, int n_seq, float *metric);
/* PCFICH (SPEC.md 9): returns the CFI 1..3 with the largest correlation; corr[3] = the three integer correlations */
int  lteo_pcfich_decode(const lteo_cell_t *cell, int sf_idx, const lteo_cf_t *sf_symbols, const lteo_cf_t *ce,
                        float noise_est, int32_t *corr);
int  lteo_tdec(const int16_t *in, int K, int max_iter, int crc_type, uint8_t *bits, int *crc_ok);
/* Debug variant: also returns ext/La after the last iteration and per-iteration bits        */
int  lteo_tdec_dbg(const int16_t *in, int K, int max_iter, int crc_type, uint8_t *bits, int *crc_ok,
                   int16_t *la_out, int window_override);
/* softbuf: [C][3*6144+12] int16 (srsLTE triples order), accumulated in place (caller resets
 * for a new transmission).  Outputs (optional, may be NULL): d, e (descrambled LLRs), cb_iters,
 * cb_crc.  Returns 0 iff TB CRC passes (srslte_pdsch_decode_rnti convention), -1 otherwise.  */
int  lteo_pdsch_decode(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg,
                       const lteo_cf_t *sf_symbols, const lteo_cf_t *ce, float noise_est,
                       int max_iter, int16_t *softbuf, uint8_t *payload,
                       lteo_cf_t *d_out, int16_t *e_out, int *cb_iters, int *cb_crc);
/* whole chain from time-domain IQ: srslte_ue_dl_decode-like with the grant supplied.
 * noise_mode: 0 = use noise_est argument, 1 = use channel-estimator noise.                    */
int  lteo_ue_dl_decode(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const lteo_cf_t *iq,
                       float noise_est, int noise_mode, int max_iter, int16_t *softbuf,
                       uint8_t *payload, float *meas, int *avg_iters);

#ifdef __cplusplus
}
#endif
#endif

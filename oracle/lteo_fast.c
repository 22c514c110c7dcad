/*
 * lteo_fast.c -- multi-threaded batch drivers of the CPU oracle, used ONLY by bench.py's cpu_baseline /
 * --impl reference legs (TEST INFRASTRUCTURE, see lte_oracle.h).  One subframe (or code block) per
 * thread at a time, the way srsUE's phch_worker pool hands one subframe to each worker
 * (/root/reference/ue/src/common/thread_pool.cc:72-82, ue/hdr/phy/phy.h:118-119).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "lte_oracle.h"

#define SB_STRIDE (3 * LTEO_MAX_K + 12)

typedef struct {
  const lteo_cell_t *cell;
  const lteo_pdsch_cfg_t *cfg;
  const lteo_cf_t *iq;
  int sf_len, n_sf, max_iter, payload_stride, nthreads, tid, noise_mode;
  float noise_est;
  uint8_t *payload;
  int32_t *status;      /* [n_sf][2]: crc ok, avg iterations */
  int *next;
  pthread_mutex_t *mu;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  lteo_cbsegm_t s;
  lteo_cbsegm(j->cfg->tbs, &s);
  int16_t *sb = (int16_t *)malloc(sizeof(int16_t) * (size_t)s.C * SB_STRIDE);
  for (;;) {
    pthread_mutex_lock(j->mu);
    int i = (*j->next)++;
    pthread_mutex_unlock(j->mu);
    if (i >= j->n_sf) break;
    memset(sb, 0, sizeof(int16_t) * (size_t)s.C * SB_STRIDE);
    float meas[5];
    int avg = 0;
    int rc = lteo_ue_dl_decode(j->cell, j->cfg, j->iq + (size_t)i * j->sf_len, j->noise_est, j->noise_mode, j->max_iter, sb,
                               j->payload + (size_t)i * j->payload_stride, meas, &avg);
    j->status[2 * i] = (rc == 0);
    j->status[2 * i + 1] = avg;
  }
  free(sb);
  return 0;
}

/* decodes n_sf subframes with nthreads worker threads; returns the number of CRC-passing blocks */
int lteo_ue_dl_decode_mt(const lteo_cell_t *cell, const lteo_pdsch_cfg_t *cfg, const lteo_cf_t *iq, int n_sf, float noise_est,
                         int noise_mode, int max_iter, int nthreads, uint8_t *payload, int32_t *status) {
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  int next = 0;
  for (int t = 0; t < nthreads; t++) {
    job_t j = {cell, cfg, iq, 15 * lteo_symbol_sz(cell->nof_prb), n_sf, max_iter, (cfg->tbs + 7) / 8, nthreads, t, noise_mode,
               noise_est, payload, status, &next, &mu};
    jobs[t] = j;
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
  int ok = 0;
  for (int i = 0; i < n_sf; i++) ok += status[2 * i];
  free(th); free(jobs);
  return ok;
}

typedef struct {
  const int16_t *in;
  int K, n_cb, max_iter, crc_type;
  uint8_t *bits;
  int32_t *iters;
  int *next;
  pthread_mutex_t *mu;
} tjob_t;

static void *tworker(void *arg) {
  tjob_t *j = (tjob_t *)arg;
  for (;;) {
    pthread_mutex_lock(j->mu);
    int i = (*j->next);
    *j->next += 16;
    pthread_mutex_unlock(j->mu);
    if (i >= j->n_cb) break;
    for (int c = i; c < i + 16 && c < j->n_cb; c++) {
      int ok = 0;
      j->iters[c] = lteo_tdec(j->in + (size_t)c * (3 * j->K + 12), j->K, j->max_iter, j->crc_type, j->bits + (size_t)c * j->K, &ok);
    }
  }
  return 0;
}

/* turbo sweep on the host: n_cb code blocks of size K, bits out one per byte */
void lteo_tdec_mt(const int16_t *in, int n_cb, int K, int max_iter, int crc_type, int nthreads, uint8_t *bits, int32_t *iters) {
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  tjob_t *jobs = (tjob_t *)malloc(sizeof(tjob_t) * nthreads);
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  int next = 0;
  for (int t = 0; t < nthreads; t++) {
    tjob_t j = {in, K, n_cb, max_iter, crc_type, bits, iters, &next, &mu};
    jobs[t] = j;
    pthread_create(&th[t], 0, tworker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
  free(th); free(jobs);
}

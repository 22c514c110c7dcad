/* float_ref.c -- an INDEPENDENT receiver back end for the fixed-point anchor tests (tests/test_fixed_point_anchor.py).
 *
 * Test infrastructure, written from the text of 3GPP TS 36.212 5.1.3.2 (turbo code, trellis termination, QPP
 * interleaver) and TS 36.211 7.1 (modulation mapper); it shares no code and no arithmetic choice with oracle/ or
 * srsue_b200/: double precision, full-length forward and backward recursions (no windows, no next-iteration
 * initialisation), no clamps, no quantisation, exact max-log demapper (minimum distances over the whole constellation
 * instead of the piecewise-linear int16 formulas).  What it has in common with the receiver under test is the
 * algorithm family only (max-log-MAP without extrinsic scaling), so the BLER gap between the two measures what the
 * frozen choices of oracle/SPEC.md 5-7 cost: int16 scaling, the +-511 soft-buffer clamp, windows with NII.
 *
 * LLR convention as in SPEC 1: positive = bit 1. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG (-1.0e30)

/* constituent encoder of 36.212 5.1.3.2.1: g0 = 1 + D^2 + D^3 (feedback), g1 = 1 + D + D^3.
 * state s = (r1 r2 r3), r1 the newest register; input u: a = u ^ r2 ^ r3, parity z = a ^ r1 ^ r3, next = (a r1 r2) */
static void trellis(int s, int u, int *next, int *z) {
  const int r1 = (s >> 2) & 1, r2 = (s >> 1) & 1, r3 = s & 1;
  const int a = u ^ r2 ^ r3;
  *z = a ^ r1 ^ r3;
  *next = (a << 2) | (r1 << 1) | r2;
}

/* one max-log-MAP pass over K steps plus the 3 termination steps; x = systematic + a-priori, y = parity,
 * tx/ty = the three tail pairs of this constituent code; out[k] = a-posteriori LLR of step k (x included) */
static void map_pass(int K, const double *x, const double *y, const double *tx, const double *ty, double *out, double *alpha) {
  static int nxt[8][2], par[8][2], init = 0;
  if (!init) {
    for (int s = 0; s < 8; s++)
      for (int u = 0; u < 2; u++) trellis(s, u, &nxt[s][u], &par[s][u]);
    init = 1;
  }
  const int N = K + 3;
  /* forward */
  for (int s = 0; s < 8; s++) alpha[s] = s ? NEG : 0.0;
  for (int k = 0; k < N; k++) {
    const double xv = k < K ? x[k] : tx[k - K], yv = k < K ? y[k] : ty[k - K];
    double *a = alpha + 8 * k, *an = a + 8;
    for (int s = 0; s < 8; s++) an[s] = NEG;
    for (int s = 0; s < 8; s++) {
      if (a[s] <= NEG / 2) continue;
      for (int u = 0; u < 2; u++) {
        if (k >= K && u != (((s >> 1) ^ s) & 1)) continue;     /* termination: the input equals r2 ^ r3, so that a = 0 */
        const double m = a[s] + u * xv + par[s][u] * yv;
        if (m > an[nxt[s][u]]) an[nxt[s][u]] = m;
      }
    }
  }
  /* backward, combining on the way */
  double beta[8], bn[8];
  for (int s = 0; s < 8; s++) beta[s] = s ? NEG : 0.0;          /* the terminated trellis ends in state 0 */
  for (int k = N - 1; k >= 0; k--) {
    const double xv = k < K ? x[k] : tx[k - K], yv = k < K ? y[k] : ty[k - K];
    const double *a = alpha + 8 * k;
    double m1 = NEG, m0 = NEG;
    for (int s = 0; s < 8; s++) {
      bn[s] = NEG;
      for (int u = 0; u < 2; u++) {
        if (k >= K && u != (((s >> 1) ^ s) & 1)) continue;
        const double b = beta[nxt[s][u]];
        if (b <= NEG / 2) continue;
        const double g = u * xv + par[s][u] * yv + b;
        if (g > bn[s]) bn[s] = g;
        if (a[s] > NEG / 2) {
          const double t = a[s] + g;
          if (u) { if (t > m1) m1 = t; } else { if (t > m0) m0 = t; }
        }
      }
    }
    if (k < K) out[k] = m1 - m0;
    memcpy(beta, bn, sizeof(beta));
  }
}

/* in: 3K+12 LLRs in the order d0_k d1_k d2_k (k < K), then x_K z_K x_K+1 z_K+1 x_K+2 z_K+2 x'_K z'_K x'_K+1 z'_K+1 x'_K+2 z'_K+2
 * (the 12 termination bits of 36.212 5.1.3.2.2 regrouped per constituent code).  bits: K hard decisions after `iters`
 * iterations; llr_out (optional): the a-posteriori LLRs in natural order. */
int fr_turbo_decode(const double *in, int K, int f1, int f2, int iters, uint8_t *bits, double *llr_out) {
  double *sys = malloc(sizeof(double) * K * 8), *p1 = sys + K, *p2 = p1 + K, *la = p2 + K, *x = la + K, *y = x + K, *o = y + K, *e2 = o + K;
  double *alpha = malloc(sizeof(double) * 8 * (K + 4));
  int *pi = malloc(sizeof(int) * K);
  if (!sys || !alpha || !pi) return -1;
  for (int k = 0; k < K; k++) {
    sys[k] = in[3 * k]; p1[k] = in[3 * k + 1]; p2[k] = in[3 * k + 2]; la[k] = 0.0;
    pi[k] = (int)(((int64_t)f1 * k + (int64_t)f2 * k * k) % K);
  }
  const double *t = in + 3 * K;
  const double t1x[3] = {t[0], t[2], t[4]}, t1y[3] = {t[1], t[3], t[5]}, t2x[3] = {t[6], t[8], t[10]}, t2y[3] = {t[7], t[9], t[11]};
  for (int it = 0; it < iters; it++) {
    for (int k = 0; k < K; k++) x[k] = sys[k] + la[k];
    map_pass(K, x, p1, t1x, t1y, o, alpha);
    for (int k = 0; k < K; k++) o[k] -= la[k];                   /* o = systematic + extrinsic of decoder 1 */
    for (int k = 0; k < K; k++) x[k] = o[pi[k]];
    for (int k = 0; k < K; k++) y[k] = p2[k];
    map_pass(K, x, y, t2x, t2y, e2, alpha);
    for (int k = 0; k < K; k++) {
      la[pi[k]] = e2[k] - x[k];                                  /* extrinsic of decoder 2, natural order */
      bits[pi[k]] = e2[k] > 0.0;
      if (llr_out) llr_out[pi[k]] = e2[k];
    }
  }
  free(sys); free(alpha); free(pi);
  return 0;
}

/* exact max-log LLRs of the 36.211 7.1 mappers.  d: n complex symbols (re, im interleaved), llr: n*qm, bit order b0 b1 ...
 * LLR_b = (min over points with b = 0 of |d - s|^2) - (min over points with b = 1 of |d - s|^2), times `gain`. */
void fr_demap(const double *d, int n, int qm, double gain, double *llr) {
  const int half = qm / 2, L = 1 << half;
  double lev[8];
  int lb[8][3];
  for (int v = 0; v < L; v++) {                                  /* v = bits of one axis, most significant first */
    int b[3] = {0, 0, 0};
    for (int i = 0; i < half; i++) b[i] = (v >> (half - 1 - i)) & 1;
    double a;
    if (half == 1) a = (1 - 2 * b[0]) / sqrt(2.0);
    else if (half == 2) a = (1 - 2 * b[0]) * (2 - (1 - 2 * b[1])) / sqrt(10.0);
    else a = (1 - 2 * b[0]) * (4 - (1 - 2 * b[1]) * (2 - (1 - 2 * b[2]))) / sqrt(42.0);
    lev[v] = a;
    for (int i = 0; i < 3; i++) lb[v][i] = b[i];
  }
  for (int i = 0; i < n; i++)
    for (int ax = 0; ax < 2; ax++) {
      const double c = d[2 * i + ax];
      for (int j = 0; j < half; j++) {
        double m0 = 1e300, m1 = 1e300;
        for (int v = 0; v < L; v++) {
          const double e = (c - lev[v]) * (c - lev[v]);
          if (lb[v][j]) { if (e < m1) m1 = e; } else { if (e < m0) m0 = e; }
        }
        llr[i * qm + 2 * j + ax] = gain * (m0 - m1);              /* axis bits interleave: b0 (I) b1 (Q) b2 (I) ... */
      }
    }
}

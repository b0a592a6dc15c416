/*
 * svgd_oracle.c -- CPU restatement (plain C, fp64) of the SVGD-ICP registration loop, the
 * `class_type = SVGDICP` branch of the same registration interface.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as svn_oracle.c): loaded by tests/, smoke() and the bench's
 * CPU-baseline legs, never by the product path.
 *
 * Parity status: pinned against the reference's own SVGDICP.cpp compiled in this container
 * (oracle/_ref, `ref_svgd_scan` in oracle/ref_driver.cpp); fixtures tests/golden/svgd_*.npz.
 *
 * Citations are relative to /root/reference/svn-icp/.  The torch::optim update rules are those of
 * the libtorch the reference links (torch/csrc/api/src/optim/{adam,rmsprop,sgd,adagrad}.cpp,
 * v2.11), configured as SVGDICP.cpp:142-170 does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* from svn_oracle.c */
void oracle_knn_mink(const double *q, int64_t n_q, const double *tgt, int64_t n_t, int K, int64_t *idx, double *d2);
void oracle_transform_q0(const double *src, int64_t n_s, const double R0[9], const double t0[3], double *q0);
double oracle_rbf_kernel(const double *x, int P, double *Kmat);

enum { OPT_ADAM = 0, OPT_RMSPROP = 1, OPT_SGD = 2, OPT_ADAGRAD = 3 };

typedef struct {
  int iterations;               /* SVGDICP.h:44 */
  double lr;                    /* SVGDICP.h:47 */
  double max_dist;              /* SVGDICP.h:48 */
  int check_early_stop;         /* SVGDICP.h:51 */
  double convergence_threshold; /* SVGDICP.h:53 */
  int knn_count;                /* SVGDICP.h:54 */
  int optimizer;                /* SVGDICP.h:50 -> OPT_*, anything else = "No optimizer chosen" */
} oracle_svgd_params;

typedef struct {
  double *grad;      /* [I][P][6] sgd_grad output (scaled)                  */
  double *stein;     /* [I][P][6] svgd_grad output                          */
  double *bandwidth; /* [I]                                                 */
  double *x_after;   /* [I][P][6] parameters after the optimizer step       */
  int32_t *corr_idx; /* [I][P][N_s]                                         */
  uint8_t *corr_mask;/* [I][P][N_s]                                         */
} oracle_svgd_dumps;

/* SVGDICP::to_rotation_tensor, SVGDICP.cpp:226-260 (ZYX Euler) */
void oracle_euler_R(double r, double p, double y, double R[9]) {
  const double cy = cos(y), sy = sin(y), cp = cos(p), sp = sin(p), cr = cos(r), sr = sin(r);
  R[0] = cp * cy; R[1] = sr * sp * cy - cr * sy; R[2] = sr * sy + cr * sp * cy;
  R[3] = cp * sy; R[4] = cr * cy + sr * sp * sy; R[5] = cr * sp * sy - sr * cy;
  R[6] = -sp;     R[7] = sr * cp;                R[8] = cr * cp;
}

/* SVGDICP::partial_derivative, SVGDICP.cpp:335-396, WITHOUT the leading R0 (applied by the caller):
 * dR[0] = dR/droll, dR[1] = dR/dpitch, dR[2] = dR/dyaw, each 3x3 row-major. */
void oracle_euler_partials(double roll, double pitch, double yaw, double dR[27]) {
  const double A = cos(yaw), B = sin(yaw), C = cos(pitch), D = sin(pitch), E = cos(roll), F = sin(roll);
  const double DE = D * E, DF = D * F, AC = A * C, AF = A * F, AE = A * E;
  const double ADE = A * DE, ADF = A * DF, BC = B * C, BE = B * E, BF = B * F, BDE = B * DE;
  double *r = dR, *p = dR + 9, *y = dR + 18;
  r[0] = 0; r[1] = ADE + BF;      r[2] = BE - ADF;               /* :360-364 */
  r[3] = 0; r[4] = -AF + BDE;     r[5] = B * (-DF) - AE;         /* :365-368 */
  r[6] = 0; r[7] = C * E;         r[8] = C * (-F);               /* :369 */
  p[0] = A * -D; p[1] = AC * F;   p[2] = AC * E;                 /* :375 */
  p[3] = B * -D; p[4] = BC * F;   p[5] = BC * E;                 /* :376 */
  p[6] = -C;     p[7] = -DF;      p[8] = -DE;                    /* :377 */
  y[0] = -BC;    y[1] = -B * DF - AE; y[2] = AF - BDE;           /* :383-386 */
  y[3] = AC;     y[4] = -BE + ADF;    y[5] = ADE + BF;           /* :387-390 */
  y[6] = 0;      y[7] = 0;            y[8] = 0;                  /* :391 */
}

static void m3mul(const double *A, const double *B, double *C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* One particle of SVGDICP::sgd_grad (SVGDICP.cpp:398-455) after get_correspondence_fast (:300-329)
 * and point_filter (:331-333).  pose = (x,y,z,roll,pitch,yaw).  out[6] is already multiplied by
 * gradient_scaling_factor_ = N_s (:58, :454). */
static void particle_sgd_grad(const double pose[6], const double R0[9], const double t0[3], const double *src,
                              int64_t n_s, const double *tgt, const int64_t *cand, int K, double max_dist,
                              double out[6], int32_t *corr_out, uint8_t *mask_out, const int32_t *given_idx,
                              const uint8_t *given_mask) {
  double R[9], Rt[9], tt[3], dR[27], M[27];
  oracle_euler_R(pose[3], pose[4], pose[5], R);                    /* :88 */
  m3mul(R0, R, Rt);                                                /* :90 */
  for (int r = 0; r < 3; r++) tt[r] = t0[r] + (R0[3 * r] * pose[0] + R0[3 * r + 1] * pose[1] + R0[3 * r + 2] * pose[2]); /* :91 */
  oracle_euler_partials(pose[3], pose[4], pose[5], dR);
  for (int k = 0; k < 3; k++) m3mul(R0, dR + 9 * k, M + 9 * k);    /* :395 */
  double esum[3] = {0, 0, 0}, g[3] = {0, 0, 0};
  double count = 0;
  for (int64_t i = 0; i < n_s; i++) {
    const double *s = src + 3 * i;
    double q[3];
    for (int r = 0; r < 3; r++) q[r] = Rt[3 * r] * s[0] + Rt[3 * r + 1] * s[1] + Rt[3 * r + 2] * s[2] + tt[r]; /* :94 */
    double best = 0;
    int bestk = -1;
    int64_t gi;
    double mu;
    if (given_idx) {                                               /* test mode: correspondences chosen elsewhere */
      gi = given_idx[i];
      mu = given_mask[i] ? 1.0 : 0.0;
    } else {
      for (int k = 0; k < K; k++) {                                /* knn.cu 1-NN, strict '<' */
        const double *m = tgt + 3 * cand[i * K + k];
        double dx = q[0] - m[0], dy = q[1] - m[1], dz = q[2] - m[2];
        double d = fma(dz, dz, fma(dy, dy, dx * dx));
        if (bestk < 0 || d < best) { best = d; bestk = k; }
      }
      gi = cand[i * K + bestk];
      mu = (best < max_dist) ? 1.0 : 0.0;                          /* :332 (Q1) */
    }
    if (corr_out) corr_out[i] = (int32_t)gi;
    if (mask_out) mask_out[i] = (uint8_t)mu;
    double sp[3] = {mu * s[0], mu * s[1], mu * s[2]};
    double tq[3] = {mu * q[0], mu * q[1], mu * q[2]};
    if ((tq[0] + tq[1]) + tq[2] != 0.0) count += 1.0;              /* :404 count_nonzero(sum(2)) */
    double e[3];
    for (int r = 0; r < 3; r++) e[r] = tq[r] - mu * tgt[3 * gi + r]; /* :407 */
    const double en = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]); /* :410 */
    double w = max_dist / (max_dist + 3.0 * en);
    w = w * w;                                                     /* :411 */
    for (int r = 0; r < 3; r++) e[r] *= w;
    for (int r = 0; r < 3; r++) esum[r] += e[r];                   /* :414 */
    for (int k = 0; k < 3; k++) {                                  /* :418-452 */
      const double *Mk = M + 9 * k;
      double acc = 0;
      for (int r = 0; r < 3; r++) acc += e[r] * (Mk[3 * r] * sp[0] + Mk[3 * r + 1] * sp[1] + Mk[3 * r + 2] * sp[2]);
      g[k] += acc;
    }
  }
  const double den = count + 1.0;
  for (int c = 0; c < 3; c++)
    out[c] = (esum[0] * R0[c] + esum[1] * R0[3 + c] + esum[2] * R0[6 + c]) / den * (double)n_s; /* :414, :454 */
  for (int k = 0; k < 3; k++) out[3 + k] = g[k] / den / 1.0 * (double)n_s;                        /* normalize_factor_ = 1 */
}

/* SVGDICP::svgd_grad + rbf_kernel, SVGDICP.cpp:457-474.  x [P][6], g = -sgd_gradient [P][6]. */
void oracle_svgd_step(const double *x, const double *g, int P, double *out, double *h_out) {
  double *Kmat = (double *)malloc(sizeof(double) * (size_t)P * P);
  const double h = oracle_rbf_kernel(x, P, Kmat);                  /* same form as :464-474 */
  if (h_out) *h_out = h;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    double rep[6] = {0, 0, 0, 0, 0, 0}, kg[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < P; j++) {
      const double kij = Kmat[(size_t)i * P + j];
      for (int d = 0; d < 6; d++) {
        rep[d] += (x[6 * i + d] - x[6 * j + d]) * kij;              /* :459-460 */
        kg[d] += kij * g[6 * (size_t)j + d];                        /* :461 */
      }
    }
    for (int d = 0; d < 6; d++) out[6 * i + d] = (kg[d] + 2.0 / h * rep[d]) / (double)P;
  }
  free(Kmat);
}

/* torch::optim step for one scalar parameter.  st: [0]=exp_avg/square_avg/sum, [1]=exp_avg_sq/momentum buffer.
 * step = 1-based step count. */
void oracle_opt_step(int opt, double lr, int step, double *p, double grad, double st[2]) {
  switch (opt) {
    case OPT_ADAM: {                                               /* betas (0.9, 0.999), eps 1e-8: SVGDICP.cpp:146-148 */
      const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
      const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
      st[0] = st[0] * b1 + grad * (1.0 - b1);
      st[1] = st[1] * b2 + grad * grad * (1.0 - b2);
      const double denom = sqrt(st[1]) / sqrt(bc2) + eps;
      *p += -(lr / bc1) * (st[0] / denom);
      break;
    }
    case OPT_RMSPROP: {                                            /* weight_decay 1e-8, momentum 0.9: :152-155 */
      const double alpha = 0.99, eps = 1e-8, wd = 1e-8, mom = 0.9;
      grad = grad + wd * (*p);
      st[0] = st[0] * alpha + grad * grad * (1.0 - alpha);
      const double avg = sqrt(st[0]) + eps;
      st[1] = st[1] * mom + grad / avg;
      *p += -lr * st[1];
      break;
    }
    case OPT_SGD:                                                  /* :159 */
      *p += -lr * grad;
      break;
    case OPT_ADAGRAD: {                                            /* :163-164, eps 1e-10, lr_decay 0 */
      st[0] += grad * grad;
      const double sd = sqrt(st[0]) + 1e-10;
      *p += -lr * (grad / sd);
      break;
    }
    default: break;
  }
}

/* n scalars at once (tests of the sharded host logic): x [n], grad [n], st [n][2] updated in place */
void oracle_opt_step_array(int opt, double lr, int step, double *x, const double *grad, double *st, int64_t n) {
  for (int64_t i = 0; i < n; i++) oracle_opt_step(opt, lr, step, x + i, grad[i], st + 2 * i);
}

static void svgd_getters(const double *x /*[P][6]*/, int P, double *particles, double *mean, double *var,
                         double *cov, double *weights) {
  for (int p = 0; p < P; p++)
    for (int c = 0; c < 6; c++) particles[c * P + p] = x[6 * p + c];  /* :515-520 */
  for (int c = 0; c < 6; c++) {
    double s = 0;
    for (int p = 0; p < P; p++) s += x[6 * p + c];
    mean[c] = s / (double)P;                                       /* :497-499 */
  }
  for (int c = 0; c < 6; c++) {
    double s = 0;
    for (int p = 0; p < P; p++) { double d = x[6 * p + c] - mean[c]; s += d * d; }
    var[c] = s / (double)(P - 1);                                  /* :501-503 torch::var is unbiased; P=1 -> NaN */
  }
  for (int a = 0; a < 6; a++)
    for (int c = 0; c < 6; c++) {
      double s = 0;
      for (int p = 0; p < P; p++) s += (x[6 * p + a] - mean[a]) * (x[6 * p + c] - mean[c]);
      cov[6 * a + c] = s / (double)P;                              /* :505-513 */
    }
  if (weights) for (int p = 0; p < P; p++) weights[p] = 1.0;       /* :522-524 */
}

/* add_cloud -> set_initial_mean -> stein_align -> getters, SVGDICP.cpp:46-140.
 * prev_pose [6][P]: the object's pose_particles_ when stein_align starts -- the constructor's
 * init_pose for the first scan, the previous scan's result afterwards: add_cloud (:46-62) does NOT
 * refresh it, so iteration 0 evaluates the kernel (and the early-stop difference) on it.
 * init_pose [6][P]: add_cloud's init_pose.  Returns 1 (ALIGN_SUCCESS) or 2 (NO_OPTIMIZER). */
int oracle_svgd_align(const oracle_svgd_params *prm, const double *src, int64_t n_s, const double *tgt, int64_t n_t,
                      const double *prev_pose, const double *init_pose, int P, const double R0[9], const double t0[3],
                      double *particles, double *mean, double *var, double *cov, double *weights, float *history,
                      int *iters_done, oracle_svgd_dumps *dmp) {
  const int K = prm->knn_count, I = prm->iterations;
  double *pp = (double *)malloc(sizeof(double) * 6 * P);   /* pose_particles_  [P][6] */
  double *x = (double *)malloc(sizeof(double) * 6 * P);    /* x_,y_,z_,rx_,ry_,rz_ [P][6] */
  for (int p = 0; p < P; p++)
    for (int c = 0; c < 6; c++) { pp[6 * p + c] = prev_pose[c * P + p]; x[6 * p + c] = init_pose[c * P + p]; }
  if (prm->optimizer < OPT_ADAM || prm->optimizer > OPT_ADAGRAD) { /* :73-75 */
    svgd_getters(pp, P, particles, mean, var, cov, weights);
    if (iters_done) *iters_done = 0;
    free(pp); free(x);
    return 2;
  }
  if (history) memset(history, 0, sizeof(float) * (size_t)I * 6 * P); /* :172-174 */
  double *q0 = (double *)malloc(sizeof(double) * 3 * (n_s > 0 ? n_s : 1));
  int64_t *cand = (int64_t *)calloc((size_t)(n_s > 0 ? n_s : 1) * K, sizeof(int64_t));
  oracle_transform_q0(src, n_s, R0, t0, q0);                       /* :201-215 */
  oracle_knn_mink(q0, n_s, tgt, n_t, K, cand, NULL);
  double *grad = (double *)malloc(sizeof(double) * 6 * P), *gneg = (double *)malloc(sizeof(double) * 6 * P);
  double *stein = (double *)malloc(sizeof(double) * 6 * P), *old = (double *)malloc(sizeof(double) * 6 * P);
  double *st = (double *)calloc((size_t)12 * P, sizeof(double));   /* optimizer state, fresh per scan (:73) */
  int done = 0;
  for (int epoch = 0; epoch < I; epoch++) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < P; p++) {
      int32_t *co = (dmp && dmp->corr_idx) ? dmp->corr_idx + ((size_t)epoch * P + p) * n_s : NULL;
      uint8_t *mo = (dmp && dmp->corr_mask) ? dmp->corr_mask + ((size_t)epoch * P + p) * n_s : NULL;
      particle_sgd_grad(x + 6 * p, R0, t0, src, n_s, tgt, cand, K, prm->max_dist, grad + 6 * p, co, mo, NULL, NULL); /* :106 */
    }
    double h = 0;
    for (int i = 0; i < 6 * P; i++) gneg[i] = -grad[i];
    if (P > 1) oracle_svgd_step(pp, gneg, P, stein, &h);           /* :109-110 */
    else memcpy(stein, gneg, sizeof(double) * 6);                  /* :112 */
    memcpy(old, pp, sizeof(double) * 6 * P);                       /* :114 */
    for (int i = 0; i < 6 * P; i++)                                /* :115, :476-494: parameter grad = -stein_grad */
      oracle_opt_step(prm->optimizer, prm->lr, epoch + 1, x + i, -stein[i], st + 2 * i);
    memcpy(pp, x, sizeof(double) * 6 * P);                         /* :118-121 */
    done = epoch + 1;
    if (dmp) {
      if (dmp->grad) memcpy(dmp->grad + (size_t)epoch * 6 * P, grad, sizeof(double) * 6 * P);
      if (dmp->stein) memcpy(dmp->stein + (size_t)epoch * 6 * P, stein, sizeof(double) * 6 * P);
      if (dmp->bandwidth) dmp->bandwidth[epoch] = h;
      if (dmp->x_after) memcpy(dmp->x_after + (size_t)epoch * 6 * P, x, sizeof(double) * 6 * P);
    }
    if (prm->check_early_stop) {                                   /* :123-131 */
      double s = 0;
      for (int p = 0; p < P; p++) {
        double n2 = 0;
        for (int c = 0; c < 6; c++) { double d = pp[6 * p + c] - old[6 * p + c]; n2 += d * d; }
        s += sqrt(n2);
      }
      if (s / (double)P < prm->convergence_threshold) break;
    }
    if (history)                                                   /* :133 */
      for (int p = 0; p < P; p++)
        for (int c = 0; c < 6; c++) history[((size_t)epoch * 6 + c) * P + p] = (float)pp[6 * p + c];
  }
  if (iters_done) *iters_done = done;
  svgd_getters(pp, P, particles, mean, var, cov, weights);
  free(pp); free(x); free(q0); free(cand); free(grad); free(gneg); free(stein); free(old); free(st);
  return 1;
}

/* gradient alone, for the kernel tests: poses [P][6] -> grad [P][6] (scaled), given a MinK table */
void oracle_svgd_grad(const double *poses, int P, const double R0[9], const double t0[3], const double *src,
                      int64_t n_s, const double *tgt, const int64_t *cand, int K, double max_dist, double *grad) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int p = 0; p < P; p++)
    particle_sgd_grad(poses + 6 * p, R0, t0, src, n_s, tgt, cand, K, max_dist, grad + 6 * p, NULL, NULL, NULL, NULL);
}

/* same, on GIVEN correspondences (corr_idx [P][n_s] global map index, corr_mask [P][n_s]): separates the arithmetic of the
 * gradient from fp32-vs-fp64 near-tie / near-threshold decisions; corr_out/mask_out optional [P][n_s] = the fp64 choices */
void oracle_svgd_grad_given_corr(const double *poses, int P, const double R0[9], const double t0[3], const double *src,
                                 int64_t n_s, const double *tgt, const int32_t *corr_idx, const uint8_t *corr_mask,
                                 double max_dist, double *grad) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int p = 0; p < P; p++)
    particle_sgd_grad(poses + 6 * p, R0, t0, src, n_s, tgt, NULL, 0, max_dist, grad + 6 * p, NULL, NULL,
                      corr_idx + (size_t)p * n_s, corr_mask + (size_t)p * n_s);
}

/* the fp64 correspondences alone, for flip statistics */
void oracle_svgd_corr(const double *poses, int P, const double R0[9], const double t0[3], const double *src, int64_t n_s,
                      const double *tgt, const int64_t *cand, int K, double max_dist, int32_t *corr_idx, uint8_t *corr_mask) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int p = 0; p < P; p++) {
    double g[6];
    particle_sgd_grad(poses + 6 * p, R0, t0, src, n_s, tgt, cand, K, max_dist, g, corr_idx + (size_t)p * n_s,
                      corr_mask + (size_t)p * n_s, NULL, NULL);
  }
}

/*
 * preprocess_oracle.c -- sequential CPU restatement (plain C) of the node's scan pre-processing:
 * OdometryPipeline::crop_pointcloud (svn-icp/src/core/OdometryPipeline.cpp:692-704) and
 * OdometryPipeline::downsample_uniform (:684-690), which is pcl::UniformSampling<PointXYZI>::filter with
 * setRadiusSearch(voxel_size).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as svn_oracle.c).
 *
 * PARITY UNPINNED for the down-sampling: PCL is absent from this image and OdometryPipeline.cpp cannot be compiled here
 * (ROS 2, GTSAM, PCL), so neither golden vectors nor reference outputs exist.  The restatement follows the published
 * source of PCL 1.12 (filters/include/pcl/filters/impl/uniform_sampling.hpp, applyFilter):
 *   inverse_leaf_size = 1.0f / float(radius);  ijk = floor(p * inverse_leaf_size) per axis (float multiply);
 *   a leaf keeps the first point that lands in it; a later point replaces it iff
 *       ||p_new.getVector4fMap() - ijk.cast<float>()||^2  <  ||p_kept.getVector4fMap() - ijk.cast<float>()||^2
 *   i.e. the distance to the leaf's INTEGER INDEX vector (not its centre), with the constant 4th component (1 - 0)^2 on
 *   both sides.  Restated here as dx*dx + dy*dy + dz*dz in float, left to right, WITHOUT the common +1 (Eigen's vectorised
 *   squaredNorm may associate differently; points whose distances differ by an ulp may therefore be picked differently --
 *   that is the documented tie tolerance).  Output order: PCL iterates a std::unordered_map -- not part of the contract.
 * The crop is fully specified by the reference's own lines and is exact.
 */
#define _GNU_SOURCE /* M_PI under -std=c11 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* crop_pointcloud, :692-704.  Returns the kept count; *max_sq = max over all points of the float squared norm (:699). */
int64_t oracle_crop(const float *xyz, int64_t n, double min_range, double max_range, float *out, double *max_sq) {
  int64_t m = 0;
  float mx = 0.f;
  for (int64_t i = 0; i < n; i++) {
    const float *p = xyz + 3 * i;
    const float norm = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];           /* :697 */
    if (norm > mx) mx = norm;                                              /* :699 */
    if ((double)norm < max_range * max_range && (double)norm > min_range * min_range) { /* :700 */
      if (out) memcpy(out + 3 * m, p, 3 * sizeof(float));
      m++;
    }
  }
  if (max_sq) *max_sq = (double)mx;
  return m;
}

typedef struct { int32_t ijk[3]; int64_t idx; int used; } us_leaf;

static size_t us_hash(const int32_t v[3], size_t mask) {
  uint64_t h = (uint32_t)v[0] * 73856093u ^ (uint32_t)v[1] * 19349669u ^ (uint32_t)v[2] * 83492791u;
  h ^= h >> 15; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
  return (size_t)h & mask;
}

/* pcl::UniformSampling::applyFilter with leaf = radius.  out: [<= n][3], in order of first appearance of each leaf. */
int64_t oracle_downsample_uniform(const float *xyz, int64_t n, double radius, float *out) {
  size_t slots = 1024;
  while (slots < (size_t)(2 * n + 2)) slots <<= 1;
  us_leaf *tab = (us_leaf *)calloc(slots, sizeof(us_leaf));
  int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int64_t n_leaves = 0;
  const float inv = 1.0f / (float)radius;
  for (int64_t cp = 0; cp < n; cp++) {
    const float *p = xyz + 3 * cp;
    const int32_t ijk[3] = {(int32_t)floorf(p[0] * inv), (int32_t)floorf(p[1] * inv), (int32_t)floorf(p[2] * inv)};
    size_t s = us_hash(ijk, slots - 1);
    while (tab[s].used && (tab[s].ijk[0] != ijk[0] || tab[s].ijk[1] != ijk[1] || tab[s].ijk[2] != ijk[2])) s = (s + 1) & (slots - 1);
    if (!tab[s].used) {                      /* leaf.idx == -1: first point of the leaf */
      tab[s].used = 1;
      memcpy(tab[s].ijk, ijk, sizeof(ijk));
      tab[s].idx = cp;
      order[n_leaves++] = (int64_t)s;
      continue;
    }
    const float *q = xyz + 3 * tab[s].idx;
    const float ax = p[0] - (float)ijk[0], ay = p[1] - (float)ijk[1], az = p[2] - (float)ijk[2];
    const float bx = q[0] - (float)ijk[0], by = q[1] - (float)ijk[1], bz = q[2] - (float)ijk[2];
    const float diff_cur = ax * ax + ay * ay + az * az, diff_prev = bx * bx + by * by + bz * bz;
    if (diff_cur < diff_prev) tab[s].idx = cp;
  }
  for (int64_t l = 0; l < n_leaves; l++) memcpy(out + 3 * l, xyz + 3 * tab[order[l]].idx, 3 * sizeof(float));
  free(tab);
  free(order);
  return n_leaves;
}

/* ---------------------------------------------------------------------------------------------------------------
 * De-skewing: OdometryPipeline::deskew_pointcloud (svn-icp/src/core/OdometryPipeline.cpp:357-447).
 *
 * The per-point arithmetic lives in GTSAM (gtsam::Pose3::Logmap / Expmap / transformFrom), a third-party dependency that
 * is absent from this image (svn-icp/CMakeLists.txt: find_package(GTSAM REQUIRED), version unpinned; the ROS 2 Humble
 * era release is GTSAM 4.2).  PARITY UNPINNED against GTSAM itself: the functions below restate the published GTSAM 4.2
 * closed forms (gtsam/geometry/Pose3.cpp, SO3.cpp) and are pinned against the mathematical definition instead
 * (scipy.linalg.expm / logm of the 4x4 twist, tests/test_deskew.py).  The control flow (timestamp normalisation, the
 * (t - 0.5) * delta_pose twist per point, the KITTI branch) follows the reference lines cited per statement.
 * Twist order is GTSAM's: xi = [omega (3) ; v (3)].
 * --------------------------------------------------------------------------------------------------------------- */

/* gtsam::SO3 Expmap (SO3.cpp, so3::ExpmapFunctor): Rodrigues; near zero (theta^2 <= eps) first order I + W */
static void gtsam_rot_expmap(const double w[3], double R[9]) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  if (theta2 <= 2.220446049250313e-16) {
    for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + W[i];
    return;
  }
  const double theta = sqrt(theta2), s = sin(theta), s2 = sin(0.5 * theta), omc = 2.0 * s2 * s2;
  double K[9], KK[9];
  for (int i = 0; i < 9; i++) K[i] = W[i] / theta;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) KK[3 * r + c] = K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
  for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + s * K[i] + omc * KK[i];
}

/* gtsam::SO3 Logmap (SO3.cpp): trace-based, Taylor near the identity; the theta ~ pi branch picks the largest column */
static void gtsam_rot_logmap(const double R[9], double w[3]) {
  const double tr = R[0] + R[4] + R[8];
  if (tr + 1.0 < 1e-10) {
    if (fabs(R[8] + 1.0) > 1e-5) {
      const double f = M_PI / sqrt(2.0 + 2.0 * R[8]);
      w[0] = f * R[2]; w[1] = f * R[5]; w[2] = f * (1.0 + R[8]);
    } else if (fabs(R[4] + 1.0) > 1e-5) {
      const double f = M_PI / sqrt(2.0 + 2.0 * R[4]);
      w[0] = f * R[1]; w[1] = f * (1.0 + R[4]); w[2] = f * R[7];
    } else {
      const double f = M_PI / sqrt(2.0 + 2.0 * R[0]);
      w[0] = f * (1.0 + R[0]); w[1] = f * R[3]; w[2] = f * R[6];
    }
    return;
  }
  double mag;
  const double tr_3 = tr - 3.0;
  if (tr_3 < -1e-7) {
    const double theta = acos((tr - 1.0) / 2.0);
    mag = theta / (2.0 * sin(theta));
  } else {
    mag = 0.5 - tr_3 / 12.0;
  }
  w[0] = mag * (R[7] - R[5]); w[1] = mag * (R[2] - R[6]); w[2] = mag * (R[3] - R[1]);
}

/* gtsam::Pose3::Expmap (Pose3.cpp): t = (w x v - R (w x v) + w (w.v)) / theta^2 */
void oracle_pose3_expmap(const double xi[6], double R[9], double t[3]) {
  const double *w = xi, *v = xi + 3;
  gtsam_rot_expmap(w, R);
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > 2.220446049250313e-16) {
    const double wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
    const double c[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
    for (int r = 0; r < 3; r++) {
      const double Rc = R[3 * r] * c[0] + R[3 * r + 1] * c[1] + R[3 * r + 2] * c[2];
      t[r] = (c[r] - Rc + w[r] * wv) / theta2;
    }
  } else {
    t[0] = v[0]; t[1] = v[1]; t[2] = v[2];
  }
}

/* gtsam::Pose3::Logmap (Pose3.cpp, Agrawal06iros eq. 14) */
void oracle_pose3_logmap(const double R[9], const double T[3], double xi[6]) {
  double w[3];
  gtsam_rot_logmap(R, w);
  const double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
  if (t < 1e-10) {
    xi[3] = T[0]; xi[4] = T[1]; xi[5] = T[2];
    return;
  }
  const double k[3] = {w[0] / t, w[1] / t, w[2] / t};
  const double WT[3] = {k[1] * T[2] - k[2] * T[1], k[2] * T[0] - k[0] * T[2], k[0] * T[1] - k[1] * T[0]};
  const double WWT[3] = {k[1] * WT[2] - k[2] * WT[1], k[2] * WT[0] - k[0] * WT[2], k[0] * WT[1] - k[1] * WT[0]};
  const double Tan = tan(0.5 * t);
  for (int i = 0; i < 3; i++) xi[3 + i] = T[i] - (0.5 * t) * WT[i] + (1.0 - t / (2.0 * Tan)) * WWT[i];
}

/* delta_pose = Logmap(start^-1 * finish), :424 */
void oracle_deskew_delta(const double Rs[9], const double ts[3], const double Rf[9], const double tf[3], double xi[6]) {
  double R[9], T[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[3 * r + c] = Rs[r] * Rf[c] + Rs[3 + r] * Rf[3 + c] + Rs[6 + r] * Rf[6 + c];
  const double d[3] = {tf[0] - ts[0], tf[1] - ts[1], tf[2] - ts[2]};
  for (int r = 0; r < 3; r++) T[r] = Rs[r] * d[0] + Rs[3 + r] * d[1] + Rs[6 + r] * d[2];
  oracle_pose3_logmap(R, T, xi);
}

/* deskew_pointcloud, :357-447.  stamps: per-point time stamps (any unit; :400-410), ignored when kitti != 0 (the KITTI
 * branch :385-399 first tilts every point by 0.205 deg about p x z and derives the stamp from its azimuth).
 * Returns 0 when all stamps are equal (the cloud is returned unchanged, :415), else 1. */
int oracle_deskew(const float *xyz, const double *stamps, int64_t n, int kitti, const double Rs[9], const double ts[3],
                  const double Rf[9], const double tf[3], float *out) {
  double *st = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  float *pts = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  memcpy(pts, xyz, sizeof(float) * 3 * (size_t)n);
  if (kitti) {
    const double off = (0.205 * M_PI) / 180.0;                                   /* :385 */
    for (int64_t i = 0; i < n; i++) {
      const double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
      double a[3] = {p[1], -p[0], 0.0};                                          /* pt.cross(z), :389 */
      const double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      if (an > 0) { a[0] /= an; a[1] /= an; a[2] /= an; }                        /* .normalized(): zero vector stays zero */
      /* Eigen::AngleAxisd(off, a) * p = Rodrigues rotation of p about a, :390-391 */
      const double c = cos(off), s = sin(off), ad = a[0] * p[0] + a[1] * p[1] + a[2] * p[2];
      const double cr[3] = {a[1] * p[2] - a[2] * p[1], a[2] * p[0] - a[0] * p[2], a[0] * p[1] - a[1] * p[0]};
      for (int k = 0; k < 3; k++) pts[3 * i + k] = (float)(c * p[k] + s * cr[k] + (1.0 - c) * ad * a[k]);  /* :392-394 */
      const double yaw = -atan2((double)pts[3 * i + 1], (double)pts[3 * i]);    /* :395-397 (float x, y promoted) */
      st[i] = 0.5 * (yaw / M_PI + 1.0);                                          /* :398 */
    }
  } else {
    for (int64_t i = 0; i < n; i++) st[i] = stamps[i];
  }
  double mn = INFINITY, mx = -INFINITY;
  for (int64_t i = 0; i < n; i++) { if (st[i] < mn) mn = st[i]; if (st[i] > mx) mx = st[i]; }   /* :411-414 */
  /* :415 returns *frame: the cloud as decoded from the message, i.e. WITHOUT the KITTI tilt (that edits a copy, :359) */
  if (n == 0 || mn == mx) { memcpy(out, xyz, sizeof(float) * 3 * (size_t)n); free(st); free(pts); return 0; }
  double xi[6];
  oracle_deskew_delta(Rs, ts, Rf, tf, xi);                                        /* :419-424 */
  for (int64_t i = 0; i < n; i++) {
    const double f = (st[i] - mn) / (mx - mn) - 0.5;                              /* :417-420, :436 */
    double tw[6], R[9], t[3];
    for (int k = 0; k < 6; k++) tw[k] = f * xi[k];
    oracle_pose3_expmap(tw, R, t);
    const double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    for (int r = 0; r < 3; r++) out[3 * i + r] = (float)(R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2] + t[r]);  /* :437-439 */
  }
  free(st);
  free(pts);
  return 1;
}

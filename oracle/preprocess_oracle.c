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

// ref_driver_map.cpp -- C entry points around the reference's OWN svnicp::VoxelHashMap (VoxelHashMap.cpp compiled
// unmodified from /root/reference over the stand-in PCL / Eigen / tsl / gtsam types of oracle/ref_shim_map).
// TEST INFRASTRUCTURE ONLY: pins the control flow of oracle/voxelmap_oracle.c (first-come voxel fill, front-point range
// tests, strict comparisons).  Built by oracle/build_ref.sh into oracle/_ref/libvmap_ref.so.
#include <cstdint>
#include <cstring>
#include "core/VoxelHashMap.h"

static gtsam::Pose3 make_pose(const double *R, const double *t) {
  gtsam::Pose3 p;
  std::memcpy(p.R, R, sizeof(p.R));
  for (int i = 0; i < 3; i++) p.t[i] = t[i];
  return p;
}

extern "C" {

void *ref_vmap_create(double voxel_size, double max_range, int max_pointscount) {
  return new svnicp::VoxelHashMap(voxel_size, max_range, max_pointscount);
}
void ref_vmap_destroy(void *m) { delete static_cast<svnicp::VoxelHashMap *>(m); }

void ref_vmap_add(void *m, const float *xyz, int64_t n, const double *R, const double *t) {
  pcl::PointCloud<svnicp::data_types::Point_t> cloud;
  cloud.points.resize((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    cloud.points[i].x = xyz[3 * i];
    cloud.points[i].y = xyz[3 * i + 1];
    cloud.points[i].z = xyz[3 * i + 2];
  }
  static_cast<svnicp::VoxelHashMap *>(m)->AddPointCloud(cloud, make_pose(R, t));
}

int64_t ref_vmap_get(void *m, const double *pos, double max_range, double *out) {
  auto *map = static_cast<svnicp::VoxelHashMap *>(m);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const auto cloud = pos ? map->GetMap(make_pose(I, pos), max_range) : map->GetMap();
  if (out)
    for (size_t i = 0; i < cloud.size(); i++) {
      out[3 * i] = cloud.points[i].x;
      out[3 * i + 1] = cloud.points[i].y;
      out[3 * i + 2] = cloud.points[i].z;
    }
  return (int64_t)cloud.size();
}

int64_t ref_vmap_size(void *m) { return (int64_t) static_cast<svnicp::VoxelHashMap *>(m)->Size(); }

}  // extern "C"

// svnicp/ScanPreprocessor.hpp -- header-only C++ mirror of the node's scan pre-processing helpers
// (svn-icp/src/core/OdometryPipeline.cpp: crop_pointcloud :692-704, downsample_uniform :684-690) over the C ABI of
// include/svnicp_b200.h (svnicp_pre_*).  Clouds are float xyz triples; results are DEVICE clouds owned by this object
// (valid until the next-but-one call), so crop -> downsample -> downsample chains without host copies and the result
// goes straight into svnicp::VoxelHashMap::AddPointCloud(..., on_device) or, via to_f64(), into add_cloud.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "svnicp/SVNICP.hpp"

namespace svnicp {

struct DeviceCloudF32 {
  const float *xyz = nullptr;  // device pointer, [n][3]
  int64_t n = 0;
};

class ScanPreprocessor {
 public:
  explicit ScanPreprocessor(int64_t max_points, int device = -1) {
    if (svnicp_pre_create(&p_, max_points, device) != SVNICP_OK) throw Error(std::string("svnicp_pre_create: ") + svnicp_pre_last_error(nullptr));
  }
  ~ScanPreprocessor() {
    if (p_) svnicp_pre_destroy(p_);
  }
  ScanPreprocessor(const ScanPreprocessor &) = delete;
  ScanPreprocessor &operator=(const ScanPreprocessor &) = delete;

  /** the reference's running maximum of the SQUARED point norm (OdometryPipeline.cpp:699), updated by crop_pointcloud */
  double scan_max_range_ = 0;

  /** deskew_pointcloud (OdometryPipeline.cpp:357-447): stamps = the message's per-point time field widened to double (host
   *  array of n; nullptr with kitti = true, the `/kitti/velo/pointcloud` branch :385-399); start / finish = the two newest
   *  poses of the node's pose buffer (:419-422), rotation row-major + translation.  moved (optional) = false when all stamps
   *  are equal and the cloud is returned unchanged (:415). */
  DeviceCloudF32 deskew_pointcloud(const float *xyz, int64_t n, const double *stamps, const double R_start[9], const double t_start[3],
                                   const double R_finish[9], const double t_finish[3], bool kitti = false, bool on_device = false,
                                   bool *moved = nullptr) {
    DeviceCloudF32 out;
    int32_t mv = 0;
    check(svnicp_pre_deskew(p_, xyz, n, on_device, stamps, 0, kitti, R_start, t_start, R_finish, t_finish, &out.xyz, &mv), "deskew_pointcloud");
    out.n = n;
    if (moved) *moved = mv != 0;
    return out;
  }
  /** crop_pointcloud (OdometryPipeline.cpp:692-704) */
  DeviceCloudF32 crop_pointcloud(const float *xyz, int64_t n, double min_range, double max_range, bool on_device = false) {
    DeviceCloudF32 out;
    double mx = 0;
    check(svnicp_pre_crop(p_, xyz, n, on_device, min_range, max_range, &out.xyz, &out.n, &mx), "crop_pointcloud");
    if (mx > scan_max_range_) scan_max_range_ = mx;
    return out;
  }
  /** downsample_uniform (OdometryPipeline.cpp:684-690): pcl::UniformSampling with leaf = voxel_size */
  DeviceCloudF32 downsample_uniform(const DeviceCloudF32 &cloud, double voxel_size) {
    DeviceCloudF32 out;
    check(svnicp_pre_downsample_uniform(p_, cloud.xyz, cloud.n, 1, voxel_size, &out.xyz, &out.n), "downsample_uniform");
    return out;
  }
  /** double copy for add_cloud({view.xyz, view.n, true}, ...) */
  CloudView to_f64(const DeviceCloudF32 &cloud) {
    const double *d = nullptr;
    check(svnicp_pre_to_f64(p_, cloud.xyz, cloud.n, &d), "to_f64");
    return CloudView{d, cloud.n, true};
  }
  std::vector<float> download(const DeviceCloudF32 &cloud) {
    std::vector<float> v((size_t)cloud.n * 3);
    check(svnicp_pre_download(p_, cloud.xyz, cloud.n, v.data()), "download");
    return v;
  }

 private:
  void check(int rc, const char *what) const {
    if (rc < 0) throw Error(std::string(what) + ": " + svnicp_pre_last_error(p_));
  }
  svnicp_pre p_ = nullptr;
};

}  // namespace svnicp

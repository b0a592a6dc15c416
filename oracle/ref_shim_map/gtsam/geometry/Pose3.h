// Stand-in for gtsam::Pose3 (oracle build of VoxelHashMap only): rotation row-major R, translation t.
#pragma once
#include <Eigen/Eigen>
namespace gtsam {
using Point3 = Eigen::Vector3d;
struct Pose3 {
  double R[9];
  Point3 t;
  Eigen::Matrix4d matrix() const {
    Eigen::Matrix4d M{};
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) M.m[r][c] = R[3 * r + c];
      M.m[r][3] = t[r];
    }
    M.m[3][3] = 1.0;
    return M;
  }
  const Point3 &translation() const { return t; }
};
}  // namespace gtsam

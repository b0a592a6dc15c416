// Minimal stand-in for gtsam::Pose3 (oracle build only).  Storage is COLUMN-major like Eigen's
// default and lives inside the Pose3 object: under the CPU device swap the reference's
// from_blob(...).to(kCPU) aliases this memory, so the Pose3 must outlive the scan.
#pragma once
#include <Eigen/Eigen>
namespace gtsam {
using Vector6 = Eigen::Matrix<double, 6, 1>;
using Matrix6 = Eigen::Matrix<double, 6, 6>;
using Matrix3 = Eigen::Matrix<double, 3, 3>;
using Point3 = Eigen::Matrix<double, 3, 1>;
struct Rot3 {
  mutable Matrix3 m;
  Matrix3 &matrix() const { return m; }
};
struct Pose3 {
  Rot3 R;
  Point3 t;
  const Rot3 &rotation() const { return R; }
  const Point3 &translation() const { return t; }
};
}  // namespace gtsam

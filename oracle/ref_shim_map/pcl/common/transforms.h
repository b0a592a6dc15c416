// Stand-in for the parts of PCL that svnicp::VoxelHashMap touches (oracle build only; PCL is absent from the image).
// transformPointCloud restates PCL 1.12 common/impl/transforms.hpp, detail::Transformer<double>::se3:
//   out = float(m00*x + m01*y + m02*z + m03), evaluated in double.
#pragma once
#include <Eigen/Eigen>
#include <iostream>
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZI {
  float x = 0, y = 0, z = 0, pad = 1, intensity = 0;
  Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(x, y, z); }
};
template <class P>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<P>>;
  std::vector<P> points;
  void push_back(const P &p) { points.push_back(p); }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  const P &front() const { return points.front(); }
  PointCloud &operator+=(const PointCloud &o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
template <class P>
void transformPointCloud(const PointCloud<P> &in, PointCloud<P> &out, const Eigen::Matrix4d &T) {
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); i++) {
    const double x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    P q = in.points[i];
    q.x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
    q.y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
    q.z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
    out.points[i] = q;
  }
}
}  // namespace pcl

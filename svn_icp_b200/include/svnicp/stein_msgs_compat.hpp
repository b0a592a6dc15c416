// svnicp/stein_msgs_compat.hpp -- the producer side of the reference's stein_msgs interface, kept verbatim.
//
// The fill_* templates below write the fields of the reference's messages exactly as OdometryPipeline does
// (svn-icp/src/core/OdometryPipeline.cpp:942-1000, message definitions stein_msgs/msg/*.msg).  They are templates over the
// message type, so inside a ROS 2 workspace they take the generated stein_msgs::msg::* types unchanged; where ROS 2 is
// absent (this repo's tests) the plain structs of namespace svnicp::stein_msgs_plain -- same field names, same types --
// stand in.  No ROS dependency is introduced by including this file.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

#include "svnicp/SVNICP.hpp"

namespace svnicp {
namespace stein_msgs_plain {
struct Header {  // std_msgs/Header (stamp left to the caller)
  std::string frame_id;
};
struct SteinParticle {  // stein_msgs/msg/SteinParticle.msg:1-8
  Header header;
  std::vector<double> x, y, z, roll, pitch, yaw, weights;
};
struct SteinParticleArray {  // SteinParticleArray.msg:1-2
  Header header;
  std::vector<SteinParticle> stein_particle_array;
};
struct Runtime {  // Runtime.msg:1-6
  Header header;
  double steinicp_time = 0, preprocessing_time = 0, knn_time = 0, update_time = 0, finish_iter = 0;
};
struct Variance {  // Variance.msg:1-5
  Header header;
  std::array<double, 6> var_icp{}, var_mean_filtered{}, var_maxsliding_filtered{}, var_random_walk{};
};
struct SteinParameters {  // SteinParameters.msg:1-20
  Header header;
  std::string optimizer;
  int64_t iterations = 0, batch_size = 0, particle_count = 0;
  bool normalize = false;
  double learning_rate = 0, correspondence_distance = 0;
  bool early_stop = false;
  int16_t converge_steps = 0;
  double converge_threshold = 0;
  bool deskew_cloud = false, voxelization = false;
  double voxel_size = 0, map_voxel_size = 0, map_voxel_max_points = 0;
  std::array<double, 2> point_range{};
  bool weight_mean = false;
  double md_learning_rate = 0;
  int64_t md_iterations = 0;
};
}  // namespace stein_msgs_plain

/** publish_particle_info (OdometryPipeline.cpp:942-965): get_particles() is [6][P] component-major, so component k is the
 *  slice [k*P, (k+1)*P) -- the layout the getters of this library keep. */
template <class SteinParticleT, class Scalar>
void fill_stein_particle(SteinParticleT &msg, const std::vector<Scalar> &particles_6xP, const std::vector<double> &weights) {
  const size_t P = particles_6xP.size() / 6;
  auto slice = [&](size_t k) { return std::vector<double>(particles_6xP.begin() + k * P, particles_6xP.begin() + (k + 1) * P); };
  msg.x = slice(0); msg.y = slice(1); msg.z = slice(2);
  msg.roll = slice(3); msg.pitch = slice(4); msg.yaw = slice(5);
  msg.weights = weights;
}

/** publish_all_particles (OdometryPipeline.cpp:967-987): one SteinParticle per iteration of get_particle_history() */
template <class SteinParticleArrayT>
void fill_stein_particle_array(SteinParticleArrayT &msg, const std::vector<std::vector<float>> &history) {
  msg.header.frame_id = "odom_svnicp";
  msg.stein_particle_array.clear();
  for (const auto &row : history) {
    typename std::remove_reference<decltype(msg.stein_particle_array)>::type::value_type p;
    fill_stein_particle(p, row, std::vector<double>());
    msg.stein_particle_array.push_back(p);
  }
}

/** publish_runtime (OdometryPipeline.cpp:989-1002; the reference leaves the three get_runtime() fields commented out, :998-1000,
 *  because SVNICP never fills them -- here they carry CUDA-event times): get_runtime() = {knn s, update s, finish_iter} */
template <class RuntimeT>
void fill_runtime(RuntimeT &msg, SVGDICP &icp, double steinicp_time, double preprocessing_time) {
  const auto rt = icp.get_runtime();
  msg.steinicp_time = steinicp_time;
  msg.preprocessing_time = preprocessing_time;
  msg.knn_time = rt[0];
  msg.update_time = rt[1];
  msg.finish_iter = rt[2];
}

/** Variance.var_icp: get_distribution() (6) */
template <class VarianceT>
void fill_variance_icp(VarianceT &msg, SVGDICP &icp) {
  const auto v = icp.get_distribution();
  for (int i = 0; i < 6; i++) msg.var_icp[i] = v[i];
}

/** the registration-class part of publish_stein_param (OdometryPipeline.cpp:839-859); the node adds its own ranges / voxel sizes */
template <class SteinParametersT>
void fill_stein_parameters(SteinParametersT &msg, const SteinICPParam &p, int particle_count, const ParticleWeightOpt &opt = {}) {
  msg.optimizer = p.optimizer;
  msg.iterations = p.iterations;
  msg.batch_size = p.batch_size;
  msg.particle_count = particle_count;
  msg.learning_rate = p.lr;
  msg.correspondence_distance = p.max_dist;
  msg.early_stop = p.check_early_stop;
  msg.converge_steps = (int16_t)p.convergence_steps;
  msg.converge_threshold = p.convergence_threshold;
  msg.weight_mean = opt.use_weight_mean;
}

}  // namespace svnicp

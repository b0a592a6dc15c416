// Stand-in for the reference's data/DataTypes.h (which drags in rclcpp/PCL/GTSAM navigation):
// only the three names the registration classes use.  Also:
//  * polyfills torch::linalg::{solve,inv}, which left the C++ frontend in torch 2.11;
//  * with -DSVN_ORACLE_CPU rewrites torch::kCUDA to kCPU AFTER all torch headers were parsed, so
//    the reference's unmodified sources run on host cores.
#pragma once
#include <torch/torch.h>
#include <gtsam/geometry/Pose3.h>
namespace svnicp::data_types {
struct Cloud_t {
  using Ptr = void *;
};
using Device_type = c10::DeviceType;
using at::indexing::Slice;
}  // namespace svnicp::data_types
namespace torch::linalg {
inline at::Tensor solve(const at::Tensor &A, const at::Tensor &B, bool left) { return at::linalg_solve(A, B, left); }
inline at::Tensor inv(const at::Tensor &A) { return at::linalg_inv(A); }
}  // namespace torch::linalg
#ifdef SVN_ORACLE_CPU
#define kCUDA kCPU
#endif

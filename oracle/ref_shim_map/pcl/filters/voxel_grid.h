// Stand-in: only VoxelHashMap::MergePoints (private, never called) names pcl::VoxelGrid.
#pragma once
#include <pcl/common/transforms.h>
namespace pcl {
template <class P>
struct VoxelGrid {
  void setInputCloud(const typename PointCloud<P>::Ptr &) {}
  void setLeafSize(float, float, float) {}
  void filter(PointCloud<P> &) {}
};
}  // namespace pcl

// Stand-in for the reference's data/DataTypes.h (which drags in rclcpp/GTSAM navigation): the two names VoxelHashMap uses
// (data/DataTypes.h:25-26 of the reference).
#pragma once
#include <pcl/common/transforms.h>
namespace svnicp::data_types {
using Point_t = pcl::PointXYZI;
using Cloud_t = pcl::PointCloud<Point_t>;
}  // namespace svnicp::data_types

// host-only check of the header-only mirrors: the stein_msgs producer mapping runs here; the other two headers only have
// to compile and link against the C ABI (they need a GPU to run: tests/test_cpp_mirror.py, tests/test_preprocess.py)
#include "svnicp/ScanPreprocessor.hpp"
#include "svnicp/VoxelHashMap.hpp"
#include "svnicp/stein_msgs_compat.hpp"
int main() {
  using namespace svnicp;
  stein_msgs_plain::SteinParticle p;
  std::vector<double> parts(12); for (int i = 0; i < 12; i++) parts[i] = i;
  fill_stein_particle(p, parts, {0.5, 0.5});
  if (p.x[1] != 1 || p.yaw[0] != 10 || p.weights.size() != 2) return 1;
  stein_msgs_plain::SteinParticleArray a;
  fill_stein_particle_array(a, {{0.f,1.f,2.f,3.f,4.f,5.f},{6.f,7.f,8.f,9.f,10.f,11.f}});
  if (a.stein_particle_array.size() != 2 || a.stein_particle_array[1].pitch[0] != 10.0) return 2;
  stein_msgs_plain::SteinParameters sp; SteinICPParam prm; fill_stein_parameters(sp, prm, 100);
  return sp.particle_count == 100 ? 0 : 3;
}

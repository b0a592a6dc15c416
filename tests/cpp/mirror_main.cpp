// Drives the header-only C++ mirror (svn_icp_b200/include/svnicp/SVNICP.hpp) exactly like the reference's caller
// (OdometryPipeline.cpp:282-288, 582-607): construct once, then add_cloud -> set_initial_mean -> stein_align -> getters.
// usage: mirror_main problem.bin result.bin     (binary layout written by tests/test_cpp_mirror.py)
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "svnicp/SVNICP.hpp"
#include "svnicp/VoxelHashMap.hpp"

template <class T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n);
  if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  const auto hdr = rd<int64_t>(f, 6);  // n_s, n_t, P, iterations, K, svn_full_grad
  const int64_t n_s = hdr[0], n_t = hdr[1];
  const int P = (int)hdr[2];
  const auto scal = rd<double>(f, 2);  // lr, max_dist
  const auto src = rd<double>(f, 3 * n_s), tgt = rd<double>(f, 3 * n_t), init = rd<double>(f, 6 * P), R0 = rd<double>(f, 9), t0 = rd<double>(f, 3);
  fclose(f);

  svnicp::SteinICPParam config;
  config.iterations = (int)hdr[3];
  config.KNN_count = (int)hdr[4];
  config.SVN_full_grad = hdr[5] != 0;
  config.lr = scal[0];
  config.max_dist = scal[1];
  // the node holds the base-class pointer (OdometryPipeline.h:125)
  // ... and picks the class from the `class_type` parameter (OdometryPipeline.cpp:282-288)
  const bool svgd = argc > 3 && std::string(argv[3]) == "SVGDICP";
  std::unique_ptr<svnicp::SVGDICP> icp;
  if (svgd) icp = std::make_unique<svnicp::SVGDICP>(config, init);
  else icp = std::make_unique<svnicp::SVNICP>(config, init, svnicp::ParticleWeightOpt{});
  svnicp::InitialMean guess;
  for (int i = 0; i < 9; i++) guess.R[i] = R0[i];
  for (int i = 0; i < 3; i++) guess.t[i] = t0[i];
  const bool via_map = argc > 4 && std::string(argv[4]) == "MAP";
  std::unique_ptr<svnicp::VoxelHashMap> local_map;
  if (via_map) {
    // the node's flow (OdometryPipeline.cpp:576-581, :630): the map is fed world-frame points at the identity pose and
    // hands the target back as a DEVICE cloud.  cap 32 / huge range: keeps every map point of this small problem
    local_map = std::make_unique<svnicp::VoxelHashMap>(1.0, 1e6, 32, 1 << 16);
    std::vector<float> tf(tgt.begin(), tgt.end());
    local_map->AddPointCloud(tf, svnicp::InitialMean{});
    const svnicp::CloudView target = local_map->GetMap(guess, 1e5);
    icp->add_cloud({src.data(), n_s, false}, target, init);
  } else {
    icp->add_cloud({src.data(), n_s, false}, {tgt.data(), n_t, false}, init);
  }
  icp->set_initial_mean(guess);
  if (icp->stein_align() != svnicp::ALIGN_SUCCESS) return 3;
  const auto mean = icp->get_transformation();
  const auto var = icp->get_distribution();
  const auto cov = icp->get_cov_matrix();
  const auto particles = icp->get_particles();
  const auto weight = icp->get_particle_weight();
  const auto hist = icp->get_particle_history();
  const auto rt = icp->get_runtime();
  FILE *o = fopen(argv[2], "wb");
  fwrite(mean.data(), 8, 6, o);
  fwrite(var.data(), 8, 6, o);
  fwrite(cov.data(), 8, 36, o);
  fwrite(particles.data(), 8, particles.size(), o);
  fwrite(weight.data(), 8, weight.size(), o);
  for (const auto &row : hist) fwrite(row.data(), 4, row.size(), o);
  fclose(o);
  printf("mirror ok: mean %.6f %.6f %.6f  finish_iter %.0f  history rows %zu\n", mean[0], mean[1], mean[2], rt[2], hist.size());
  // error behaviour: the mirror throws where the reference would throw c10::Error
  try {
    icp->add_cloud({src.data(), 0, false}, {tgt.data(), n_t, false}, init);
    return 4;
  } catch (const svnicp::Error &) {
  }
  return 0;
}

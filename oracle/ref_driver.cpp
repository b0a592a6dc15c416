// ref_driver.cpp -- C entry points around the UNMODIFIED reference registration classes
// (compiled from /root/reference/svn-icp/src/core/{SVGDICP,SVNICP}.cpp by oracle/build_ref.sh).
//
// TEST INFRASTRUCTURE ONLY (see oracle/svn_oracle.c header).  Outputs go to oracle/_ref/.
// This file is ours; it contains no reference code.  It supplies
//   * KNearestNeighborIdx for fp64 CPU tensors with the semantics of the reference's CUDA
//     kernels (knn.cu:68-111 global-memory MinK for K>32, knn.cu:204-251 RegisterMinK for
//     K==1; mink.cuh:62-83,132-153) -- the reference's own CPU KNN (knn_cpu.cpp:28-33) is
//     float32-only and priority-queue ordered, so it cannot serve the fp64 hot path;
//   * ref_scan(): add_cloud -> set_initial_mean -> stein_align -> getters, the call order of
//     OdometryPipeline.cpp:582-607;
//   * ref_scan_steps(): the same scan advanced one iteration at a time (iterations = 1,
//     SURVEY.md App. B) so fp64 per-iteration particles, H, b and matched target points can be
//     recorded as golden vectors.
// NOTE (quirk Q8): function-static tensors in SVNICP.cpp:42,167 freeze the particle count and
// the early-stop threshold at first use -> ONE (P, threshold) per process.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <tuple>
#include <vector>
#include <torch/torch.h>

#define private public
#define protected public
#include "core/SVNICP.h"
#undef private
#undef protected

// -DREF_CUDA (oracle/build_ref_cuda.sh): no device swap and the reference's OWN vendored CUDA K-NN (knn.cu) -- the
// reference as it runs on a GPU, for bench.py's `reference_on_gpu` comparator.  Inputs are moved to the device, results
// are read back through host(), and the timers synchronise the device.
#ifdef REF_CUDA
#include <torch/cuda.h>
static inline at::Tensor host(const at::Tensor &t) { return t.to(torch::kCPU).contiguous(); }
static inline at::Tensor dev(const at::Tensor &t) { return t.to(torch::kCUDA); }
static inline void dsync() { torch::cuda::synchronize(); }
#else
static inline at::Tensor host(const at::Tensor &t) { return t.contiguous(); }
static inline at::Tensor dev(const at::Tensor &t) { return t; }
static inline void dsync() {}
#endif

#ifndef REF_CUDA
// ------------------------------------------------------------------------------------------
std::tuple<at::Tensor, at::Tensor> KNearestNeighborIdx(const at::Tensor &p1_, const at::Tensor &p2_,
                                                       const at::Tensor &lengths1, const at::Tensor &lengths2,
                                                       const int norm, const int K, const int /*version*/) {
  TORCH_CHECK(norm == 2, "oracle KNN: only the squared-L2 path of the hot loop is provided");
  TORCH_CHECK(p1_.scalar_type() == at::kDouble && p2_.scalar_type() == at::kDouble, "oracle KNN: fp64 only");
  const auto p1 = p1_.contiguous(), p2 = p2_.contiguous();
  const int64_t N = p1.size(0), P1 = p1.size(1), P2 = p2.size(1);
  const auto l1 = lengths1.to(at::kLong).contiguous().view({-1});
  const auto l2 = lengths2.to(at::kLong).contiguous().view({-1});
  auto idxs = at::zeros({N, P1, K}, at::TensorOptions().dtype(at::kLong));
  auto dists = at::zeros({N, P1, K}, p1.options());
  const double *a = p1.data_ptr<double>(), *b = p2.data_ptr<double>();
  const int64_t *len1 = l1.data_ptr<int64_t>(), *len2 = l2.data_ptr<int64_t>();
  int64_t *io = idxs.data_ptr<int64_t>();
  double *dop = dists.data_ptr<double>();
  at::parallel_for(0, N * P1, 16, [&](int64_t lo, int64_t hi) {
    for (int64_t w = lo; w < hi; w++) {
      const int64_t n = w / P1, i = w % P1;
      if (i >= len1[n]) continue;
      double *keys = dop + (n * P1 + i) * K;
      int64_t *vals = io + (n * P1 + i) * K;
      const double *q = a + (n * P1 + i) * 3;
      int size = 0, max_idx = 0;
      double max_key = 0;
      for (int64_t j = 0; j < len2[n]; j++) {
        const double *m = b + (n * P2 + j) * 3;
        const double dx = q[0] - m[0], dy = q[1] - m[1], dz = q[2] - m[2];
        const double d = std::fma(dz, dz, std::fma(dy, dy, dx * dx));
        if (size < K) {
          keys[size] = d; vals[size] = j;
          if (size == 0 || d > max_key) { max_key = d; max_idx = size; }
          size++;
        } else if (d < max_key) {
          keys[max_idx] = d; vals[max_idx] = j; max_key = d;
          for (int k = 0; k < K; k++) if (keys[k] > max_key) { max_key = keys[k]; max_idx = k; }
        }
      }
    }
  });
  return std::make_tuple(idxs, dists);
}
#endif  // !REF_CUDA

// ------------------------------------------------------------------------------------------
extern "C" {

struct ref_params {
  int iterations;
  double lr;
  double max_dist;
  int check_early_stop;
  double convergence_threshold;
  int knn_count;
  int svn_full_grad;
};

int ref_num_threads() { return at::get_num_threads(); }
void ref_set_num_threads(int n) { if (n > 0) at::set_num_threads(n); }

static svnicp::SteinICPParam to_param(const ref_params *p) {
  svnicp::SteinICPParam c;
  c.iterations = p->iterations;
  c.lr = p->lr;
  c.max_dist = p->max_dist;
  c.check_early_stop = p->check_early_stop != 0;
  c.convergence_threshold = p->convergence_threshold;
  c.KNN_count = p->knn_count;
  c.SVN_full_grad = p->svn_full_grad != 0;
  c.optimizer = "Adam";
  return c;
}

static gtsam::Pose3 *make_pose(const double *R0_rowmajor, const double *t0) {
  auto *pose = new gtsam::Pose3();  // leaked on purpose: the reference aliases this storage
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) pose->R.m.d[c * 3 + r] = R0_rowmajor[r * 3 + c];  // column-major
  for (int r = 0; r < 3; r++) pose->t.d[r] = t0[r];
  return pose;
}

static at::Tensor blob(const double *p, std::vector<int64_t> shape) {
  return dev(torch::from_blob(const_cast<double *>(p), shape, torch::TensorOptions().dtype(torch::kFloat64)).clone());
}

static void getters(svnicp::SVNICP &icp, int P, int I, double *particles, double *mean, double *var,
                    double *cov, double *weights, float *history) {
  const auto m = host(icp.get_transformation());
  const auto v = host(icp.get_distribution());
  const auto c = icp.get_cov_matrix();
  const auto pp = icp.get_particles();
  const auto w = icp.get_particle_weight();
  std::memcpy(mean, m.data_ptr<double>(), 6 * sizeof(double));
  std::memcpy(var, v.data_ptr<double>(), 6 * sizeof(double));
  std::memcpy(cov, c.data(), 36 * sizeof(double));
  std::memcpy(particles, pp.data(), sizeof(double) * 6 * P);
  if (weights) std::memcpy(weights, w.data(), sizeof(double) * P);
  if (history) {
    const auto h = icp.get_particle_history();
    for (int i = 0; i < I; i++) std::memcpy(history + (size_t)i * 6 * P, h[i].data(), sizeof(float) * 6 * P);
  }
}

// Whole scan through the public interface only.  seconds[0] = add_cloud+set_initial_mean,
// seconds[1] = stein_align, seconds[2] = getters.
int ref_scan(const ref_params *prm, const double *src, int64_t n_s, const double *tgt, int64_t n_t,
             const double *init_pose /*[6][P]*/, int P, const double *R0, const double *t0,
             double *particles, double *mean, double *var, double *cov, double *weights, float *history,
             double *seconds) {
  torch::NoGradGuard ng;
  const auto cfg = to_param(prm);
  const auto init = blob(init_pose, {6, P, 1});
  svnicp::ParticleWeightOpt opt;
  svnicp::SVNICP icp(cfg, init, opt);
  const auto s = blob(src, {n_s, 3}), t = blob(tgt, {n_t, 3});
  dsync();
  const auto T0 = std::chrono::steady_clock::now();
  icp.add_cloud(s, t, init.clone());
  icp.set_initial_mean(*make_pose(R0, t0));
  dsync();
  const auto T1 = std::chrono::steady_clock::now();
  const int state = icp.stein_align();
  dsync();
  const auto T2 = std::chrono::steady_clock::now();
  getters(icp, P, prm->iterations, particles, mean, var, cov, weights, history);
  const auto T3 = std::chrono::steady_clock::now();
  if (seconds) {
    seconds[0] = std::chrono::duration<double>(T1 - T0).count();
    seconds[1] = std::chrono::duration<double>(T2 - T1).count();
    seconds[2] = std::chrono::duration<double>(T3 - T2).count();
  }
  return state;
}

// Scan advanced one iteration per stein_align() call.  Optional dumps (NULL to skip):
//   x_after [I][6][P] fp64 get_particles() after each iteration
//   H [I][P][36], b [I][P][6]   from Newton_grad_right on the state BEFORE that iteration
//   tgt_paired [I][P][n_s][3], src_tr [I][P][n_s][3]  matched target point / transformed source (masked)
//   cand_idx [n_s][K] int64 MinK-ordered candidate table
int ref_scan_steps(const ref_params *prm, const double *src, int64_t n_s, const double *tgt, int64_t n_t,
                   const double *init_pose, int P, const double *R0, const double *t0, int steps,
                   double *x_after, double *Hd, double *bd, double *tgt_paired, double *src_tr, int64_t *cand_idx,
                   double *particles, double *mean, double *var, double *cov) {
  torch::NoGradGuard ng;
  auto cfg = to_param(prm);
  cfg.iterations = 1;
  cfg.check_early_stop = false;
  const auto init = blob(init_pose, {6, P, 1});
  svnicp::ParticleWeightOpt opt;
  svnicp::SVNICP icp(cfg, init, opt);
  const auto s = blob(src, {n_s, 3}), t = blob(tgt, {n_t, 3});
  icp.add_cloud(s, t, init.clone());
  icp.set_initial_mean(*make_pose(R0, t0));
  int state = 0;
  for (int it = 0; it < steps; it++) {
    if (Hd || bd || tgt_paired || src_tr || (cand_idx && it == 0)) {
      // replay the head of the iteration with the object's own methods (SVNICP.cpp:50-71)
      const auto [mb, tb] = icp.mini_batch_pair_generator();
      if (cand_idx && it == 0)
        std::memcpy(cand_idx, host(icp.sourceKNN_idx_).data_ptr<int64_t>(), sizeof(int64_t) * n_s * cfg.KNN_count);
      const auto mbe = mb[0].expand({P, (int64_t)n_s, 3});
      const auto Rtot = icp.R0_.matmul(icp.R_);
      const auto ttot = icp.t0_ + icp.R0_.matmul(icp.t_);
      icp.R_total_ = Rtot;
      icp.t_total_ = ttot;
      const auto tr = mbe.matmul(Rtot.transpose(1, 2)) + ttot.view({P, 1, 3});
      const auto [sp, trp, tp] = icp.get_correspondence_fast(mbe, tr, tb[0]);
      const auto [ng_, H, b] = icp.Newton_grad_right(sp, trp, tp);
      if (Hd) std::memcpy(Hd + (size_t)it * P * 36, host(H).data_ptr<double>(), sizeof(double) * P * 36);
      if (bd) std::memcpy(bd + (size_t)it * P * 6, host(b).data_ptr<double>(), sizeof(double) * P * 6);
      if (tgt_paired) std::memcpy(tgt_paired + (size_t)it * P * n_s * 3, host(tp).data_ptr<double>(), sizeof(double) * P * n_s * 3);
      if (src_tr) std::memcpy(src_tr + (size_t)it * P * n_s * 3, host(trp).data_ptr<double>(), sizeof(double) * P * n_s * 3);
    }
    state = icp.stein_align();
    if (x_after) {
      const auto pp = icp.get_particles();
      std::memcpy(x_after + (size_t)it * 6 * P, pp.data(), sizeof(double) * 6 * P);
    }
  }
  getters(icp, P, 1, particles, mean, var, cov, nullptr, nullptr);
  return state;
}

// SVGD-ICP class through its public interface (SVGDICP.h:64-110).  ctor_pose is what the
// constructor sees (it seeds pose_particles_, which add_cloud does not refresh), init_pose what
// add_cloud sees.  Two consecutive scans when init_pose2 != NULL (same clouds), outputs from the last.
int ref_svgd_scan(const ref_params *prm, const char *optimizer, const double *src, int64_t n_s, const double *tgt,
                  int64_t n_t, const double *ctor_pose, const double *init_pose, const double *init_pose2, int P,
                  const double *R0, const double *t0, double *particles, double *mean, double *var, double *cov,
                  double *weights, float *history, double *runtime3) {
  torch::NoGradGuard ng;
  auto cfg = to_param(prm);
  cfg.optimizer = optimizer;
  const auto ctor = blob(ctor_pose, {6, P, 1});
  svnicp::SVGDICP icp(cfg, ctor);
  const auto s = blob(src, {n_s, 3}), t = blob(tgt, {n_t, 3});
  int state = 0;
  for (int scan = 0; scan < (init_pose2 ? 2 : 1); scan++) {
    const auto init = blob(scan == 0 ? init_pose : init_pose2, {6, P, 1});
    icp.add_cloud(s, t, init);
    icp.set_initial_mean(*make_pose(R0, t0));
    state = icp.stein_align();
  }
  const auto m = host(icp.get_transformation());
  const auto v = host(icp.get_distribution());
  const auto c = icp.get_cov_matrix();
  const auto pp = icp.get_particles();
  const auto w = icp.get_particle_weight();
  std::memcpy(mean, m.data_ptr<double>(), 6 * sizeof(double));
  std::memcpy(var, v.data_ptr<double>(), 6 * sizeof(double));
  std::memcpy(cov, c.data(), 36 * sizeof(double));
  std::memcpy(particles, pp.data(), sizeof(double) * 6 * P);
  if (weights) std::memcpy(weights, w.data(), sizeof(double) * P);
  if (history && state == 1) {
    const auto h = icp.get_particle_history();
    for (int i = 0; i < prm->iterations; i++) std::memcpy(history + (size_t)i * 6 * P, h[i].data(), sizeof(float) * 6 * P);
  }
  if (runtime3) {
    const auto r = icp.get_runtime();
    for (int i = 0; i < 3; i++) runtime3[i] = r[i];
  }
  return state;
}

}  // extern "C"

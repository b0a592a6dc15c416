// stein_kernels.cu -- particle initialisation, the separate decide / radix-median kernels (SVGD-ICP class) and the getters.
// (The SVN-ICP class runs its Stein phase in tail2.cu: k_head + k_tail.)
//
// Kept here (both classes / SVGD-ICP class):
//   k_particles_init     SVGDICP::add_cloud :46-61 with SVNICP's Exp (SVNICP.cpp:166-194)
//   k_align_reset        head of every stein_align
//   k_decide             early stop :95-101, history row :103-107, SoA copy of the gathered record   (SVGD-ICP class)
//   k_median_pass x 5    rbf kernel bandwidth: LOWER median over all P^2 entries (torch::median), exact radix select
//   k_stats              getters :281-308  weighted mean / variance / covariance
// All fp64.
#include "common.cuh"
#include "kernels.h"

namespace svn {

__device__ __forceinline__ int pass_bits(int s) { return s == 0 ? 11 : 13; }
__device__ __forceinline__ int bits_before(int s) { return s == 0 ? 0 : 11 + 13 * (s - 1); }

// ---------------------------------------------------------------------------------------------
// k_particles_init: SVGDICP::add_cloud (SVGDICP.cpp:46-61) with SVNICP's Exp (SVNICP.cpp:166-194)
// ---------------------------------------------------------------------------------------------
__global__ void k_particles_init(double *R, double *t, const double *init_pose, int P, double *dnorm, int p_lo, int P_l, Ctrl *ctrl) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) {
    ctrl->stop = 0; ctrl->iter = 0; ctrl->iters_done = 0; ctrl->bandwidth = 0.0; ctrl->kept_total = 0ull;
  }
  if (p >= P) return;
  const double r[3] = {init_pose[3 * P + p], init_pose[4 * P + p], init_pose[5 * P + p]};
  double Rm[9];
  so3_exp(r, Rm, nullptr);
  for (int i = 0; i < 9; i++) R[9 * (size_t)p + i] = Rm[i];
  for (int i = 0; i < 3; i++) t[3 * (size_t)p + i] = init_pose[i * P + p];
  if (p >= p_lo && p < p_lo + P_l) dnorm[p - p_lo] = 0.0;
}

// k_align_reset: head of every stein_align.  A second scan without add_cloud continues from the current poses (as the
// reference does) but with fresh iteration counters, stop flag and -- SVGD-ICP class -- optimizer moments (SVGDICP.cpp:73).
__global__ void k_align_reset(Ctrl *ctrl, double *opt_state, size_t n_opt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    ctrl->stop = 0; ctrl->iter = 0; ctrl->iters_done = 0; ctrl->kept_total = 0ull;
    ctrl->error = 0; ctrl->fin_ticket = 0u; ctrl->tail_ticket = 0u;
  }
  for (size_t k = i; k < n_opt; k += (size_t)gridDim.x * blockDim.x) opt_state[k] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// k_decide: early-stop decision for the PREVIOUS iteration + its history row; builds the SoA copy of the gathered
// record the Stein kernels read; resets the median select.  Runs after the all-gather, so every rank (and every CTA:
// the decision is recomputed redundantly from the same record) decides identically.
// recT rows: 0..32 = rec[0..32] (x, b, H), 33..38 = rec[34..39] (g).
// ---------------------------------------------------------------------------------------------
constexpr int RT_ROWS = 39;
constexpr int DECIDE_THREADS = 512;

__global__ void __launch_bounds__(DECIDE_THREADS) k_decide(SteinArgs a, int epilogue) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  __shared__ double s_part[32];
  __shared__ int s_stop;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it = c->iter;  // updates applied so far
  // mean_p |delta_p| of the last applied update, fixed summation order (SVNICP.cpp:96)
  double s = 0.0;
  for (int p = tid; p < a.P; p += blockDim.x) s += a.rec[(size_t)p * REC + REC_DNORM];
  s = warp_sum(s);
  if (lane == 0) s_part[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_part[w];
    int stop = 0;
    if (a.check_early_stop && it > 0 && tot / (double)a.P < a.threshold) stop = 1;
    s_stop = stop;
    if (blockIdx.x == 0) {
      if (stop) { c->stop = 1; c->iters_done = it; }
      else if (epilogue) c->iters_done = it;
    }
  }
  __syncthreads();
  if (s_stop) return;  // break BEFORE the history row of that iteration (Q9)
  const int gtid = blockIdx.x * blockDim.x + tid, gn = gridDim.x * blockDim.x;
  if (it > 0 && it - 1 < a.I) {
    float *row = a.history + (size_t)(it - 1) * 6 * a.P;
    for (int i = gtid; i < 6 * a.P; i += gn) {
      const int comp = i / a.P, p = i % a.P;
      row[i] = (float)a.rec[(size_t)p * REC + REC_X + comp];
    }
  }
  if (epilogue) return;
  for (int i = gtid; i < RT_ROWS * a.P; i += gn) {
    const int row = i / a.P, p = i % a.P;
    a.xs[i] = a.rec[(size_t)p * REC + (row < 33 ? row : row + 1)];
  }
  for (int i = gtid; i < MED_PASSES * MED_BINS; i += gn) a.hist[i] = 0u;
  if (gtid < MED_PASSES) c->med_ticket[gtid] = 0u;
  if (gtid == 0) {
    if (a.kept_hist && it <= a.I) a.kept_hist[it] = c->kept_total;
    c->sel_prefix[0] = 0ull;
    c->sel_rank[0] = ((unsigned long long)a.P * (unsigned long long)a.P - 1ull) / 2ull;  // lower median
  }
}

// derive (prefix, rank) after pass s-1 from its histogram; every CTA computes the same values
__device__ void select_from_hist(const unsigned *hist, unsigned long long prefix_in, unsigned long long rank_in, int nbins, int bits,
                                 unsigned long long *prefix_out, unsigned long long *rank_out) {
  // parallel: every thread sums its own run of bins, a block-wide exclusive scan of the run sums locates the run
  // that contains the rank, and only that thread walks its (<= 32, cache-hot) bins.
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_res[2];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int per = (nbins + nt - 1) / nt;
  unsigned vals[16];  // per <= 16 (8192 bins / 512 threads): all loads in flight at once, kept in registers
  unsigned long long loc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int b = tid * per + i;
    vals[i] = (i < per && b < nbins) ? __ldcg(hist + b) : 0u;  // written by other CTAs' atomics: read at L2
    loc += vals[i];
  }
  unsigned long long incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  if (tid == 0) { s_res[0] = (prefix_in << bits) | (unsigned long long)(nbins - 1); s_res[1] = 0ull; }
  __syncthreads();
  unsigned long long wbase = 0;
  for (int w = 0; w < warp; w++) wbase += s_warp[w];
  const unsigned long long excl = wbase + incl - loc;
  if (loc > 0 && excl <= rank_in && rank_in < excl + loc) {  // exactly one thread owns the rank
    unsigned long long cum = excl;
    int b = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (cum + vals[i] > rank_in) break;
      cum += vals[i];
      b = i + 1;
    }
    b += tid * per;
    s_res[0] = (prefix_in << bits) | (unsigned long long)b;
    s_res[1] = rank_in - cum;
  }
  __syncthreads();
  *prefix_out = s_res[0];
  *rank_out = s_res[1];
  __syncthreads();
}

__device__ __forceinline__ double pair_d2(const double *__restrict__ xs, int P, int i, int j) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < 6; d++) {
    const double df = xs[d * P + i] - xs[d * P + j];
    s += df * df;  // SVNICP.cpp:257-260
  }
  return s;
}

// one radix-select pass over the upper triangle of D (D_ij = D_ji counted twice, diagonal once).  Few fat CTAs; the LAST
// CTA to finish (ticket counter) turns the histogram into the next (prefix, rank) -- and after the final pass into the
// bandwidth h = median / log(P + 1) (SVNICP.cpp:262, Q4) -- so no other kernel has to rescan histograms.
constexpr int MED_THREADS = 512;

__global__ void __launch_bounds__(MED_THREADS) k_median_pass(SteinArgs a, int s) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  __shared__ unsigned s_hist[MED_BINS];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const unsigned long long prefix = c->sel_prefix[s];
  const int nb = 1 << pass_bits(s);
  for (int i = tid; i < nb; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const int consumed = bits_before(s);
  const int shift = 63 - consumed - pass_bits(s);
  const unsigned bmask = (unsigned)(nb - 1);
  const int P = a.P;
  for (int i = blockIdx.x; i < P; i += gridDim.x) {
    for (int j = i + tid; j < P; j += blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(pair_d2(a.xs, P, i, j));
      const bool match = (consumed == 0) || ((key >> (63 - consumed)) == prefix);
      if (match) atomicAdd(&s_hist[(unsigned)(key >> shift) & bmask], (j == i) ? 1u : 2u);
    }
  }
  __syncthreads();
  unsigned *gh = a.hist + (size_t)s * MED_BINS;
  for (int i = tid; i < nb; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&gh[i], s_hist[i]);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&c->med_ticket[s], 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  unsigned long long np, nr;
  select_from_hist(gh, prefix, c->sel_rank[s], nb, pass_bits(s), &np, &nr);
  if (tid == 0) {
    c->sel_prefix[s + 1] = np;
    c->sel_rank[s + 1] = nr;
    if (s == MED_PASSES - 1) c->bandwidth = __longlong_as_double((long long)np) / log((double)(a.P + 1));
  }
}

// ---------------------------------------------------------------------------------------------
// k_stats: getters (SVNICP.cpp:281-308) from the gathered record.  One CTA.
// weights = float32(1)/P promoted to double (SVNICP.cpp:46, :281-284).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_stats(SteinArgs a, int parity_from_ctrl) {
  // SVN-ICP class: the final poses sit in the record buffer of parity (number of updates applied & 1)
  if (parity_from_ctrl) a.rec += (size_t)(a.ctrl->iters_done & 1) * a.rec_stride;
  __shared__ double s_red[32][36];
  __shared__ double s_mean[6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const double w = (double)(1.0f / (float)a.P);
  for (int i = tid; i < 6 * a.P; i += blockDim.x) {
    const int comp = i / a.P, p = i % a.P;
    a.particles[i] = a.rec[(size_t)p * REC + REC_X + comp];  // get_particles: [6][P] (SVGDICP.cpp:515-520)
  }
  double acc[36];
#pragma unroll
  for (int q = 0; q < 6; q++) acc[q] = 0.0;
  for (int p = tid; p < a.P; p += blockDim.x)
#pragma unroll
    for (int q = 0; q < 6; q++) acc[q] += a.rec[(size_t)p * REC + REC_X + q] * w;  // :288
#pragma unroll
  for (int q = 0; q < 6; q++) acc[q] = warp_sum(acc[q]);
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < 6; q++) s_red[warp][q] = acc[q];
  __syncthreads();
  if (tid < 6) {
    double s = 0.0;
    for (int ww = 0; ww < nw; ww++) s += s_red[ww][tid];
    s_mean[tid] = s;
    a.stats[tid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 36; q++) acc[q] = 0.0;
  for (int p = tid; p < a.P; p += blockDim.x) {
    double d[6];
#pragma unroll
    for (int q = 0; q < 6; q++) d[q] = a.rec[(size_t)p * REC + REC_X + q] - s_mean[q];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc < 6; cc++) acc[6 * r + cc] += w * (d[r] * d[cc]);  // :302-304
  }
#pragma unroll
  for (int q = 0; q < 36; q++) acc[q] = warp_sum(acc[q]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < 36; q++) s_red[warp][q] = acc[q];
  __syncthreads();
  if (tid < 36) {
    double s = 0.0;
    for (int ww = 0; ww < nw; ww++) s += s_red[ww][tid];
    a.stats[12 + tid] = s;                     // covariance, row-major 36 (:305-306)
    if (tid % 7 == 0) a.stats[6 + tid / 7] = s;  // variance (:294-295) = diagonal of the same weighted sum
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int launch_align_reset(Ctrl *ctrl, double *opt_state, size_t n_opt, cudaStream_t st) {
  k_align_reset<<<n_opt > 4096 ? 32 : 1, 256, 0, st>>>(ctrl, opt_state, n_opt);
  return 1;
}

int launch_init_particles(double *R, double *t, const double *init_pose_dev, int P, double *dnorm, int p_lo, int P_l, Ctrl *ctrl,
                          cudaStream_t st) {
  k_particles_init<<<cdiv(P, 128), 128, 0, st>>>(R, t, init_pose_dev, P, dnorm, p_lo, P_l, ctrl);
  return 1;
}

int launch_decide(const SteinArgs &a, cudaStream_t st, int epilogue) {
  k_decide<<<epilogue ? 4 : 32, DECIDE_THREADS, 0, st>>>(a, epilogue);
  return 1;
}

int launch_median(const SteinArgs &a, cudaStream_t st) {
  if (a.P < 2) return 0;
  const int grid = a.P < a.sm_count ? a.P : a.sm_count;
  for (int s = 0; s < MED_PASSES; s++) k_median_pass<<<grid, MED_THREADS, 0, st>>>(a, s);
  return MED_PASSES;
}

int launch_stats(const SteinArgs &a, int parity_from_ctrl, cudaStream_t st) {
  k_stats<<<1, 1024, 0, st>>>(a, parity_from_ctrl);
  return 1;
}

}  // namespace svn

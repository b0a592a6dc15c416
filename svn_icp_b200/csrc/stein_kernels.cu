// stein_kernels.cu -- the Stein Variational Newton step (north_star kernel (c)) and the scan epilogue.
//
// Replaces, per iteration of SVNICP::stein_align (reference svn-icp/src/core/SVNICP.cpp):
//   rbf_hessian_kernel   :254-266  pairwise squared distances, LOWER median over all P^2 entries
//                                  (torch::median), K = exp(-D/h)   -> exact 5-pass radix select, the
//                                  [P,P] matrices are never materialised
//   svn_full_grad        :229-252  kernel-weighted Hessians, one 6x6 solve per particle in registers
//   svgd_grad            :218-227  pre-conditioned SVGD step (shipped default SVNFullGrad=false)
//   pose_update          :268-279  right-multiplicative update with the left Jacobian
//   early stop           :95-101   decided on the device; later kernels become no-ops
//   history row          :103-107  float32 [6][P]
//   getters              :281-308  weighted mean / variance / covariance
// All fp64.  Summation order over j is fixed (lane-strided, xor tree) so that an N-GPU run reproduces
// the 1-GPU Stein step bit for bit.
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
  if (i == 0) { ctrl->stop = 0; ctrl->iter = 0; ctrl->iters_done = 0; ctrl->kept_total = 0ull; }
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
// k_stein_full: SVNICP::svn_full_grad (SVNICP.cpp:229-252).  A CTA owns ST_NI particles i; each is served by ST_JQ warps
// that split every 128-wide j tile of the 33-double record (staged in shared memory) into quarters; lane = j.
// Fixed summation order (lane-strided per quarter, xor tree, quarters 0..3), identical on every rank.
// ---------------------------------------------------------------------------------------------
constexpr int ST_WARPS = 8;
constexpr int ST_TJ = 32;    // j-tile of the pre-conditioned SVGD kernel
constexpr int ST_TJF = 128;  // j-tile of the full SVN kernel
constexpr int ST_JQ = 2;     // warps per particle
constexpr int ST_NI = ST_WARPS / ST_JQ;

__global__ void __launch_bounds__(ST_WARPS * 32) k_stein_full(SteinArgs a) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  __shared__ double s_rec[33][ST_TJF + 1];
  __shared__ double s_part[ST_WARPS][28];
  const double h = c->bandwidth;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ii = warp / ST_JQ, jq = warp % ST_JQ;
  const int l = blockIdx.x * ST_NI + ii;
  const bool active = l < a.P_l;
  const int i = a.p_lo + (active ? l : 0);
  double xi[6];
#pragma unroll
  for (int d = 0; d < 6; d++) xi[d] = a.rec[(size_t)i * REC + REC_X + d];
  double Hm[21], v[6];
#pragma unroll
  for (int q = 0; q < 21; q++) Hm[q] = 0.0;
#pragma unroll
  for (int q = 0; q < 6; q++) v[q] = 0.0;
  const double two_over_h = 2.0 / h;
  for (int j0 = 0; j0 < a.P; j0 += ST_TJF) {
    __syncthreads();
    {
      // coalesced rows of the SoA record copy; all loads are issued before the first store (one L2 round trip per tile)
      constexpr int NLD = (33 * ST_TJF + ST_WARPS * 32 - 1) / (ST_WARPS * 32);
      double tmp[NLD];
#pragma unroll
      for (int u = 0; u < NLD; u++) {
        const int e = tid + u * ST_WARPS * 32;
        const int q = e / ST_TJF, jj = e % ST_TJF;
        tmp[u] = (e < 33 * ST_TJF && j0 + jj < a.P) ? __ldg(a.xs + (size_t)q * a.P + j0 + jj) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < NLD; u++) {
        const int e = tid + u * ST_WARPS * 32;
        if (e < 33 * ST_TJF) s_rec[e / ST_TJF][e % ST_TJF] = tmp[u];
      }
    }
    __syncthreads();
    for (int u = 0; u < ST_TJF / 32 / ST_JQ; u++) {  // this warp's share of the tile, ascending j (fixed order)
    const int jj = (jq * (ST_TJF / 32 / ST_JQ) + u) * 32 + lane;
    if (active && j0 + jj < a.P) {
      double dl[6], D = 0.0;
#pragma unroll
      for (int d = 0; d < 6; d++) { dl[d] = xi[d] - s_rec[REC_X + d][jj]; D += dl[d] * dl[d]; }
      const double kij = exp(-D / h);          // :264
      const double k2 = kij * kij;             // :238
      double g[6];
#pragma unroll
      for (int d = 0; d < 6; d++) g[d] = two_over_h * (dl[d] * kij);  // :233
      int q = 0;
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = r; cc < 6; cc++, q++) Hm[q] += k2 * s_rec[REC_H + q][jj] + g[r] * g[cc];  // :236-242
#pragma unroll
      for (int d = 0; d < 6; d++) v[d] += g[d] - kij * s_rec[REC_B + d][jj];  // :244 with b' = -b
    }
    }
  }
#pragma unroll
  for (int q = 0; q < 21; q++) Hm[q] = warp_sum(Hm[q]);
#pragma unroll
  for (int q = 0; q < 6; q++) v[q] = warp_sum(v[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 21; q++) s_part[warp][q] = Hm[q];
#pragma unroll
    for (int q = 0; q < 6; q++) s_part[warp][21 + q] = v[q];
  }
  __syncthreads();
  if (active && jq == 0 && lane == 0) {
    double A[36], x[6];
    for (int r = 0; r < 6; r++)
      for (int cc = r; cc < 6; cc++) {
        double sum = 0.0;
        for (int w = 0; w < ST_JQ; w++) sum += s_part[ii * ST_JQ + w][tri(r, cc)];
        A[6 * r + cc] = sum / (double)a.P;
        A[6 * cc + r] = A[6 * r + cc];
      }
    for (int d = 0; d < 6; d++) {
      double sum = 0.0;
      for (int w = 0; w < ST_JQ; w++) sum += s_part[ii * ST_JQ + w][21 + d];
      x[d] = sum / (double)a.P;
    }
    lu_solve6(A, x, 1);  // :250 (the reference forms the explicit inverse; tolerance-level difference)
    for (int d = 0; d < 6; d++) a.delta[(size_t)l * 6 + d] = a.lr * x[d];
  }
}

// mean Hessian and its inverse for the pre-conditioned SVGD step (SVNICP.cpp:85, :225). One CTA.
__global__ void __launch_bounds__(1024) k_mean_hessian(SteinArgs a) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  __shared__ double s_sum[32][21];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[21];
#pragma unroll
  for (int q = 0; q < 21; q++) acc[q] = 0.0;
  for (int p = tid; p < a.P; p += blockDim.x)
#pragma unroll
    for (int q = 0; q < 21; q++) acc[q] += a.rec[(size_t)p * REC + REC_H + q];
#pragma unroll
  for (int q = 0; q < 21; q++) acc[q] = warp_sum(acc[q]);
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < 21; q++) s_sum[warp][q] = acc[q];
  __syncthreads();
  if (tid == 0) {
    double A[36], Inv[36];
    for (int r = 0; r < 6; r++)
      for (int cc = r; cc < 6; cc++) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += s_sum[w][tri(r, cc)];
        A[6 * r + cc] = s / (double)a.P;
        A[6 * cc + r] = A[6 * r + cc];
      }
    for (int q = 0; q < 36; q++) Inv[q] = (q % 7 == 0) ? 1.0 : 0.0;
    lu_solve6(A, Inv, 6);
    for (int q = 0; q < 36; q++) a.Hbar_inv[q] = Inv[q];
  }
}

// k_stein_svgd: SVNICP::svgd_grad (SVNICP.cpp:218-227), no lr (Q5)
__global__ void __launch_bounds__(ST_WARPS * 32) k_stein_svgd(SteinArgs a) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  __shared__ double s_rec[12][ST_TJ + 1];
  const double h = c->bandwidth;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = blockIdx.x * ST_WARPS + warp;
  const bool active = l < a.P_l;
  const int i = a.p_lo + (active ? l : 0);
  double xi[6];
#pragma unroll
  for (int d = 0; d < 6; d++) xi[d] = a.rec[(size_t)i * REC + REC_X + d];
  double gs[6], kn[6], ks = 0.0;
#pragma unroll
  for (int d = 0; d < 6; d++) { gs[d] = 0.0; kn[d] = 0.0; }
  for (int j0 = 0; j0 < a.P; j0 += ST_TJ) {
    __syncthreads();
    for (int e = tid; e < 12 * ST_TJ; e += blockDim.x) {
      const int q = e / ST_TJ, jj = e % ST_TJ;
      const int row = (q < 6) ? q : (33 + q - 6);  // x rows 0..5, g rows 33..38 of the SoA record copy
      s_rec[q][jj] = (j0 + jj < a.P) ? a.xs[(size_t)row * a.P + j0 + jj] : 0.0;
    }
    __syncthreads();
    if (active && j0 + lane < a.P) {
      double dl[6], D = 0.0;
#pragma unroll
      for (int d = 0; d < 6; d++) { dl[d] = xi[d] - s_rec[d][lane]; D += dl[d] * dl[d]; }
      const double kij = exp(-D / h);
      ks += kij;                                                       // :226
#pragma unroll
      for (int d = 0; d < 6; d++) { gs[d] += dl[d] * kij; kn[d] -= kij * s_rec[6 + d][lane]; }  // :221-224 (newton passed negated, :86)
    }
  }
  ks = warp_sum(ks);
#pragma unroll
  for (int d = 0; d < 6; d++) { gs[d] = warp_sum(gs[d]); kn[d] = warp_sum(kn[d]); }
  if (active && lane == 0) {
    const double f = 2.0 / h;
    for (int r = 0; r < 6; r++) {
      double s = 0.0;
      for (int cc = 0; cc < 6; cc++) s += a.Hbar_inv[6 * r + cc] * (f * gs[cc]);
      a.delta[(size_t)l * 6 + r] = (kn[r] + s) / ks;
    }
  }
}

// P == 1: stein_grad = -H^-1 b (SVNICP.cpp:88-89)
__global__ void k_stein_single(SteinArgs a) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  if (threadIdx.x != 0 || blockIdx.x != 0 || a.P_l < 1) return;
  double A[36], x[6];
  const double *rec = a.rec + (size_t)a.p_lo * REC;
  for (int r = 0; r < 6; r++)
    for (int cc = r; cc < 6; cc++) { A[6 * r + cc] = rec[REC_H + tri(r, cc)]; A[6 * cc + r] = A[6 * r + cc]; }
  for (int d = 0; d < 6; d++) x[d] = rec[REC_B + d];
  lu_solve6(A, x, 1);
  for (int d = 0; d < 6; d++) a.delta[d] = -x[d];
}

// ---------------------------------------------------------------------------------------------
// k_update: SVNICP::pose_update (SVNICP.cpp:268-279) for the local slice
// ---------------------------------------------------------------------------------------------
__global__ void k_update(SteinArgs a) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < a.P_l) {
    const int p = a.p_lo + l;
    double d[6];
    for (int i = 0; i < 6; i++) d[i] = a.delta[(size_t)l * 6 + i];
    double dR[9], Jl[9], R[9], Rn[9], dt[3];
    so3_exp(d + 3, dR, Jl);  // :269-271 (J_l side effect :188-192)
    for (int r = 0; r < 3; r++) dt[r] = Jl[3 * r] * d[0] + Jl[3 * r + 1] * d[1] + Jl[3 * r + 2] * d[2];  // :275
    for (int i = 0; i < 9; i++) R[i] = a.R[9 * (size_t)p + i];
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) Rn[3 * r + cc] = R[3 * r] * dR[cc] + R[3 * r + 1] * dR[3 + cc] + R[3 * r + 2] * dR[6 + cc];  // :277
    for (int i = 0; i < 9; i++) a.R[9 * (size_t)p + i] = Rn[i];
    for (int r = 0; r < 3; r++)
      a.t[3 * (size_t)p + r] = (Rn[3 * r] * dt[0] + Rn[3 * r + 1] * dt[1] + Rn[3 * r + 2] * dt[2]) + a.t[3 * (size_t)p + r];  // :278 (Q6)
    a.dnorm[l] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);  // :96 norm(2,1)
  }
  if (l == 0) c->iter = c->iter + 1;
}

// ---------------------------------------------------------------------------------------------
// k_stats: getters (SVNICP.cpp:281-308) from the gathered record.  One CTA.
// weights = float32(1)/P promoted to double (SVNICP.cpp:46, :281-284).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_stats(SteinArgs a) {
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

int launch_stein(const SteinArgs &a, cudaStream_t st) {
  if (a.P < 2) {
    k_stein_single<<<1, 32, 0, st>>>(a);
    return 1;
  }
  const int grid = cdiv(a.P_l, ST_WARPS);
  if (a.svn_full_grad) {
    k_stein_full<<<cdiv(a.P_l, ST_NI), ST_WARPS * 32, 0, st>>>(a);
    return 1;
  }
  k_mean_hessian<<<1, 1024, 0, st>>>(a);
  k_stein_svgd<<<grid, ST_WARPS * 32, 0, st>>>(a);
  return 2;
}

int launch_update(const SteinArgs &a, cudaStream_t st) {
  k_update<<<cdiv(a.P_l, 128), 128, 0, st>>>(a);
  return 1;
}

int launch_stats(const SteinArgs &a, cudaStream_t st) {
  k_stats<<<1, 1024, 0, st>>>(a);
  return 1;
}

}  // namespace svn

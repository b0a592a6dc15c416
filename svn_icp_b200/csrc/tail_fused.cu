// tail_fused.cu -- the whole "Stein phase" of one SVN iteration in ONE cooperative kernel.
//
//   decide (early stop for the previous update, history row, SoA record copy)        SVNICP.cpp:95-107
//   exact lower median of the P^2 pairwise squared distances -> bandwidth h            SVNICP.cpp:254-266
//   Stein step: full SVN / pre-conditioned SVGD / P == 1                              SVNICP.cpp:218-252, 88-89
//   pose update                                                                       SVNICP.cpp:268-279
//   next iteration's fp32 transforms + exact-pruning ball (what k_prep does)           SVNICP.cpp:58-59,74-77
//
// The separate kernels (stein_kernels.cu / k_prep) cost 9 launches = ~190 us per iteration at configs[1], almost all
// of it launch ramp and single-wave latency; here the phases are separated by cooperative-groups grid barriers
// (9 per iteration).  Arithmetic and summation orders are those of the separate kernels, so the Stein step is
// bit-identical to them (tests/test_gpu_parity.py::test_fused_tail_same_bits); only the pruning ball's centre is
// summed in a different order, which cannot change any result (pruning is exact for any centre).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace svn {

constexpr int TF_THREADS = 512;
constexpr int TF_WARPS = TF_THREADS / 32;
constexpr int TF_TJ = 128;  // j tile of the full SVN step
constexpr int TF_JQ = 2;    // warps per particle (same split as k_stein_full): 8 particles per CTA pass
constexpr int TF_NI = TF_WARPS / TF_JQ;
constexpr int TF_RT_ROWS = 39;

__device__ __forceinline__ int tf_pass_bits(int s) { return s == 0 ? 11 : 13; }
__device__ __forceinline__ int tf_bits_before(int s) { return s == 0 ? 0 : 11 + 13 * (s - 1); }

// (prefix, rank) after a pass from its finished global histogram; every CTA computes the same values
__device__ void tf_select(const unsigned *hist, unsigned long long prefix_in, unsigned long long rank_in, int nbins, int bits,
                          unsigned long long *prefix_out, unsigned long long *rank_out, unsigned long long *s_warp,
                          unsigned long long *s_res) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int per = (nbins + nt - 1) / nt;  // <= 16 (8192 bins / 512 threads)
  unsigned vals[16];
  unsigned long long loc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int b = tid * per + i;
    vals[i] = (i < per && b < nbins) ? __ldcg(hist + b) : 0u;  // all loads in flight at once; kept in registers
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
  if (loc > 0 && excl <= rank_in && rank_in < excl + loc) {
    unsigned long long cum = excl;
    int b = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (cum + vals[i] > rank_in) break;
      cum += vals[i];
      b = i + 1;
    }
    s_res[0] = (prefix_in << bits) | (unsigned long long)(tid * per + b);
    s_res[1] = rank_in - cum;
  }
  __syncthreads();
  *prefix_out = s_res[0];
  *rank_out = s_res[1];
  __syncthreads();
}

__device__ __forceinline__ unsigned long long tf_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// phase time stamps of CTA 0 (ns), written behind the per-CTA partial sums: [sm_count*12 + k]
#define TF_STAMP(k) do { if (gtid == 0) a.prep_scratch_d[(size_t)gridDim.x * 12 + (k)] = (double)tf_now(); } while (0)

__global__ void __launch_bounds__(TF_THREADS, 1) k_tail_fused(SteinArgs a, IterArgs ia, int xs_smem_bytes) {
  cg::grid_group grid = cg::this_grid();
  Ctrl *c = a.ctrl;
  if (c->stop) return;  // set by an earlier launch: identical for every CTA
  extern __shared__ __align__(16) unsigned char s_dyn[];  // optional copy of x [6][P] for the median passes
  __shared__ __align__(16) unsigned char s_raw[33 * (TF_TJ + 1) * 8];  // histogram (32 KB) / record tile (34 KB), never live together
  __shared__ double s_part[TF_WARPS][28];
  __shared__ double s_red[32][21];
  __shared__ unsigned long long s_warp[32], s_res[2];
  __shared__ double s_Hinv[36];
  __shared__ float s_center[12];
  __shared__ int s_env[PRUNE_BINS + 2];
  __shared__ int s_flag[2];
  unsigned *s_hist = reinterpret_cast<unsigned *>(s_raw);
  double(*s_rec)[TF_TJ + 1] = reinterpret_cast<double(*)[TF_TJ + 1]>(s_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gtid = blockIdx.x * blockDim.x + tid, gn = gridDim.x * blockDim.x;
  const int P = a.P;
  const int it = c->iter;
  // previous iteration's median (bandwidth is rewritten only after the barriers below): the guess of the fast median path
  const double med_guess = c->bandwidth * log((double)(P + 1));
  TF_STAMP(0);

  // ------------------------------------------------------------------ decide (as k_decide, redundantly per CTA)
  {
    double s = 0.0;
    for (int p = tid; p < P; p += blockDim.x) s += a.rec[(size_t)p * REC + REC_DNORM];
    s = warp_sum(s);
    if (lane == 0) s_red[warp][0] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < TF_WARPS; w++) tot += s_red[w][0];
      const int stop = (a.check_early_stop && it > 0 && tot / (double)P < a.threshold) ? 1 : 0;
      s_flag[0] = stop;
      if (stop && blockIdx.x == 0) { c->stop = 1; c->iters_done = it; }
    }
    __syncthreads();
    if (s_flag[0]) return;  // every CTA takes the same decision from the same record: nobody waits at a barrier
    if (it > 0 && it - 1 < a.I) {
      float *row = a.history + (size_t)(it - 1) * 6 * P;
      for (int i = gtid; i < 6 * P; i += gn) {
        const int comp = i / P, p = i % P;
        row[i] = (float)a.rec[(size_t)p * REC + REC_X + comp];
      }
    }
    for (int i = gtid; i < TF_RT_ROWS * P; i += gn) {
      const int row = i / P, p = i % P;
      a.xs[i] = a.rec[(size_t)p * REC + (row < 33 ? row : row + 1)];
    }
    for (int i = gtid; i < MED_PASSES * MED_BINS; i += gn) a.hist[i] = 0u;
    for (int i = gtid; i < PRUNE_BINS + 3; i += gn) a.prep_scratch_i[i] = 0;
    if (gtid == 0 && a.kept_hist && it <= a.I) a.kept_hist[it] = c->kept_total;
  }
  TF_STAMP(1);
  grid.sync();
  TF_STAMP(2);

  // ------------------------------------------------------------------ bandwidth: exact lower median, 5 radix passes
  double h = 0.0;
  if (P > 1) {
    // the 5 passes re-read x for every pair: keep it in shared memory when it fits (no L2 round trip per row)
    const double *X = a.xs;
    if (xs_smem_bytes > 0) {
      double *sx = reinterpret_cast<double *>(s_dyn);
      for (int i = tid; i < 6 * P; i += blockDim.x) sx[i] = __ldcg(a.xs + i);
      __syncthreads();
      X = sx;
    }
    unsigned long long prefix = 0ull, rank = ((unsigned long long)P * (unsigned long long)P - 1ull) / 2ull;
    bool have_median = false;
    double median = 0.0;
    // ---- fast path (2 passes instead of 5): the median moves little between iterations, so histogram LINEARLY around the
    // previous one (8190 bins over [0.5, 1.5) x previous median, one bin below, one above), then gather the few dozen
    // values of the bin that holds the rank and pick the exact order statistic.  Same value as the radix select (the
    // lower median is unique), so the bandwidth is bit-identical to the 5-pass path; any miss (median left the window,
    // degenerate bin) falls back to it.  Every branch below depends only on data all CTAs see identically.
    constexpr int TF_COLLECT_CAP = 4096;  // doubles; staged in s_raw (32 KB)
    if (P >= 64 && it > 0 && med_guess > 0.0 && med_guess < INFINITY) {
      const double lo = 0.5 * med_guess, hi = 1.5 * med_guess, scale = (double)(MED_BINS - 2) / (hi - lo);
      auto bin_of = [&](double d) -> unsigned {  // monotone non-decreasing in d; NaN sorts last like its bit pattern
        if (d < lo) return 0u;
        if (!(d < hi)) return (unsigned)(MED_BINS - 1);
        const int b = (int)((d - lo) * scale);
        return 1u + (unsigned)min(b, MED_BINS - 3);
      };
      for (int i = tid; i < MED_BINS; i += blockDim.x) s_hist[i] = 0u;
      __syncthreads();
      for (int i = blockIdx.x; i < P; i += gridDim.x)
        for (int j0 = i + 1; j0 < P; j0 += blockDim.x) {
          const int j = j0 + tid;
          unsigned bin = 0;
          if (j < P) {
            double d2 = 0.0;
#pragma unroll
            for (int d = 0; d < 6; d++) {
              const double df = X[d * P + i] - X[d * P + j];
              d2 += df * df;
            }
            bin = bin_of(d2);
          }
          const unsigned mm = __ballot_sync(0xffffffffu, j < P);
          if (j < P) {
            const unsigned peers = __match_any_sync(mm, bin);
            if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_hist[bin], 2u * (unsigned)__popc(peers));
          }
        }
      if (blockIdx.x == 0 && tid == 0) atomicAdd(&s_hist[0], (unsigned)P);  // the diagonal: exact zeros, below lo
      __syncthreads();
      for (int i = tid; i < MED_BINS; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&a.hist[i], s_hist[i]);
      grid.sync();
      unsigned long long tbin = 0ull, trank = 0ull;
      tf_select(a.hist, 0ull, rank, MED_BINS, 13, &tbin, &trank, s_warp, s_res);
      if (tbin != 0ull && tbin != (unsigned long long)(MED_BINS - 1)) {
        // gather the values of that bin once per unordered pair (each stands for two entries of the P x P matrix)
        unsigned *cursor = a.hist + MED_BINS;                                   // zeroed with the histograms
        double *list = reinterpret_cast<double *>(a.hist + 2 * MED_BINS);       // rows 2..4: 12288 doubles
        for (int i = blockIdx.x; i < P; i += gridDim.x)
          for (int j = i + 1 + tid; j < P; j += blockDim.x) {
            double d2 = 0.0;
#pragma unroll
            for (int d = 0; d < 6; d++) {
              const double df = X[d * P + i] - X[d * P + j];
              d2 += df * df;
            }
            if (bin_of(d2) == (unsigned)tbin) {
              const unsigned k = atomicAdd(cursor, 1u);
              if (k < (unsigned)TF_COLLECT_CAP) list[k] = d2;
            }
          }
        grid.sync();
        const unsigned n_c = __ldcg(cursor);
        if (n_c >= 1u && n_c <= (unsigned)TF_COLLECT_CAP) {
          double *sl = reinterpret_cast<double *>(s_raw);
          for (unsigned k = tid; k < n_c; k += blockDim.x) sl[k] = __ldcg(list + k);
          if (tid == 0) s_flag[1] = 0;
          __syncthreads();
          const unsigned target = (unsigned)(trank >> 1);  // index among the distinct pairs of the bin, ascending
          for (unsigned k = tid; k < n_c; k += blockDim.x) {
            const double v = sl[k];
            unsigned less = 0, eq = 0;
            for (unsigned u = 0; u < n_c; u++) {
              const double w = sl[u];
              less += (w < v) ? 1u : 0u;
              eq += (w == v) ? 1u : 0u;
            }
            if (less <= target && target < less + eq) { s_res[0] = (unsigned long long)__double_as_longlong(v); s_flag[1] = 1; }
          }
          __syncthreads();
          if (s_flag[1]) { have_median = true; median = __longlong_as_double((long long)s_res[0]); }
          __syncthreads();
        }
      }
      if (!have_median) {  // miss: clean the scratch the radix passes expect to be zero
        for (int i = gtid; i < MED_PASSES * MED_BINS; i += gn) a.hist[i] = 0u;
        grid.sync();
      }
    }
    for (int s = 0; s < MED_PASSES && !have_median; s++) {
      const int nb = 1 << tf_pass_bits(s);
      for (int i = tid; i < nb; i += blockDim.x) s_hist[i] = 0u;
      __syncthreads();
      const int consumed = tf_bits_before(s);
      const int shift = 63 - consumed - tf_pass_bits(s);
      const unsigned bmask = (unsigned)(nb - 1);
      // upper triangle, each D_ij = D_ji counted twice; lanes that fall into the same bin are aggregated with
      // match.any before the shared-memory atomic (the first pass puts nearly every pair into 2-3 exponent bins)
      for (int i = blockIdx.x; i < P; i += gridDim.x)
        for (int j0 = i + 1; j0 < P; j0 += blockDim.x) {
          const int j = j0 + tid;
          bool match = false;
          unsigned bin = 0;
          if (j < P) {
            double d2 = 0.0;
#pragma unroll
            for (int d = 0; d < 6; d++) {
              const double df = X[d * P + i] - X[d * P + j];
              d2 += df * df;  // SVNICP.cpp:257-260
            }
            const unsigned long long key = (unsigned long long)__double_as_longlong(d2);
            match = (consumed == 0) || ((key >> (63 - consumed)) == prefix);
            bin = (unsigned)(key >> shift) & bmask;
          }
          const unsigned mm = __ballot_sync(0xffffffffu, match);
          if (match) {
            const unsigned peers = __match_any_sync(mm, bin);
            if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_hist[bin], 2u * (unsigned)__popc(peers));
          }
        }
      // the P diagonal entries are exact zeros (key 0): they match only an all-zero prefix and fall into bin 0
      if (blockIdx.x == 0 && tid == 0 && (consumed == 0 || prefix == 0ull)) atomicAdd(&s_hist[0], (unsigned)P);
      __syncthreads();
      unsigned *gh = a.hist + (size_t)s * MED_BINS;
      for (int i = tid; i < nb; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&gh[i], s_hist[i]);
      grid.sync();
      tf_select(gh, prefix, rank, nb, tf_pass_bits(s), &prefix, &rank, s_warp, s_res);
    }
    if (!have_median) median = __longlong_as_double((long long)prefix);
    h = median / log((double)(P + 1));  // SVNICP.cpp:262 (Q4)
    if (gtid == 0) c->bandwidth = h;
  }

  TF_STAMP(3);
  // ------------------------------------------------------------------ Stein step for the local slice
  if (P < 2) {
    if (gtid == 0 && a.P_l >= 1) {  // SVNICP.cpp:88-89
      double A[36], x[6];
      const double *rec = a.rec + (size_t)a.p_lo * REC;
      for (int r = 0; r < 6; r++)
        for (int cc = r; cc < 6; cc++) { A[6 * r + cc] = rec[REC_H + tri(r, cc)]; A[6 * cc + r] = A[6 * r + cc]; }
      for (int d = 0; d < 6; d++) x[d] = rec[REC_B + d];
      lu_solve6(A, x, 1);
      for (int d = 0; d < 6; d++) a.delta[d] = -x[d];
    }
  } else if (a.svn_full_grad) {
    const int ii = warp / TF_JQ, jq = warp % TF_JQ;
    const int n_groups = (a.P_l + TF_NI - 1) / TF_NI;
    const double two_over_h = 2.0 / h;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
      const int l = g * TF_NI + ii;
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
      for (int j0 = 0; j0 < P; j0 += TF_TJ) {
        __syncthreads();
        {
          constexpr int NLD = (33 * TF_TJ + TF_THREADS - 1) / TF_THREADS;
          double tmp[NLD];
#pragma unroll
          for (int u = 0; u < NLD; u++) {
            const int e = tid + u * TF_THREADS;
            const int q = e / TF_TJ, jj = e % TF_TJ;
            tmp[u] = (e < 33 * TF_TJ && j0 + jj < P) ? __ldcg(a.xs + (size_t)q * P + j0 + jj) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < NLD; u++) {
            const int e = tid + u * TF_THREADS;
            if (e < 33 * TF_TJ) s_rec[e / TF_TJ][e % TF_TJ] = tmp[u];
          }
        }
        __syncthreads();
        for (int u = 0; u < TF_TJ / 32 / TF_JQ; u++) {  // this warp's share of the tile, ascending j (fixed order)
        const int jj = (jq * (TF_TJ / 32 / TF_JQ) + u) * 32 + lane;
        if (active && j0 + jj < P) {
          double dl[6], D = 0.0;
#pragma unroll
          for (int d = 0; d < 6; d++) { dl[d] = xi[d] - s_rec[REC_X + d][jj]; D += dl[d] * dl[d]; }
          const double kij = exp(-D / h);  // :264
          const double k2 = kij * kij;     // :238
          double gv[6];
#pragma unroll
          for (int d = 0; d < 6; d++) gv[d] = two_over_h * (dl[d] * kij);  // :233
          int q = 0;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = r; cc < 6; cc++, q++) Hm[q] += k2 * s_rec[REC_H + q][jj] + gv[r] * gv[cc];  // :236-242
#pragma unroll
          for (int d = 0; d < 6; d++) v[d] += gv[d] - kij * s_rec[REC_B + d][jj];  // :244 with b' = -b
        }
        }
      }
#pragma unroll
      for (int q = 0; q < 21; q++) Hm[q] = warp_sum(Hm[q]);
#pragma unroll
      for (int q = 0; q < 6; q++) v[q] = warp_sum(v[q]);
      __syncthreads();
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
            for (int w = 0; w < TF_JQ; w++) sum += s_part[ii * TF_JQ + w][tri(r, cc)];
            A[6 * r + cc] = sum / (double)P;
            A[6 * cc + r] = A[6 * r + cc];
          }
        for (int d = 0; d < 6; d++) {
          double sum = 0.0;
          for (int w = 0; w < TF_JQ; w++) sum += s_part[ii * TF_JQ + w][21 + d];
          x[d] = sum / (double)P;
        }
        lu_solve6(A, x, 1);  // :250
        for (int d = 0; d < 6; d++) a.delta[(size_t)l * 6 + d] = a.lr * x[d];
      }
    }
  } else {
    // pre-conditioned SVGD (SVNICP.cpp:85, :218-227).  Mean Hessian with the summation tree of k_mean_hessian
    // (1024 virtual threads: this thread plays virtual threads tid and tid + 512), inverse per CTA.
    {
      double acc0[21], acc1[21];
#pragma unroll
      for (int q = 0; q < 21; q++) { acc0[q] = 0.0; acc1[q] = 0.0; }
      for (int p = tid; p < P; p += 1024)
#pragma unroll
        for (int q = 0; q < 21; q++) acc0[q] += a.rec[(size_t)p * REC + REC_H + q];
      for (int p = tid + 512; p < P; p += 1024)
#pragma unroll
        for (int q = 0; q < 21; q++) acc1[q] += a.rec[(size_t)p * REC + REC_H + q];
#pragma unroll
      for (int q = 0; q < 21; q++) { acc0[q] = warp_sum(acc0[q]); acc1[q] = warp_sum(acc1[q]); }
      if (lane == 0)
#pragma unroll
        for (int q = 0; q < 21; q++) { s_red[warp][q] = acc0[q]; s_red[warp + 16][q] = acc1[q]; }
      __syncthreads();
      if (tid == 0) {
        double A[36], Inv[36];
        for (int r = 0; r < 6; r++)
          for (int cc = r; cc < 6; cc++) {
            double s = 0.0;
            for (int w = 0; w < 32; w++) s += s_red[w][tri(r, cc)];
            A[6 * r + cc] = s / (double)P;
            A[6 * cc + r] = A[6 * r + cc];
          }
        for (int q = 0; q < 36; q++) Inv[q] = (q % 7 == 0) ? 1.0 : 0.0;
        lu_solve6(A, Inv, 6);
        for (int q = 0; q < 36; q++) s_Hinv[q] = Inv[q];
      }
      __syncthreads();
    }
    double(*s12)[33] = reinterpret_cast<double(*)[33]>(s_raw);  // [12][32+1]
    const int n_groups = (a.P_l + TF_WARPS - 1) / TF_WARPS;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
      const int l = g * TF_WARPS + warp;
      const bool active = l < a.P_l;
      const int i = a.p_lo + (active ? l : 0);
      double xi[6];
#pragma unroll
      for (int d = 0; d < 6; d++) xi[d] = a.rec[(size_t)i * REC + REC_X + d];
      double gs[6], kn[6], ks = 0.0;
#pragma unroll
      for (int d = 0; d < 6; d++) { gs[d] = 0.0; kn[d] = 0.0; }
      for (int j0 = 0; j0 < P; j0 += 32) {
        __syncthreads();
        for (int e = tid; e < 12 * 32; e += blockDim.x) {
          const int q = e / 32, jj = e % 32;
          const int row = (q < 6) ? q : (33 + q - 6);
          s12[q][jj] = (j0 + jj < P) ? __ldcg(a.xs + (size_t)row * P + j0 + jj) : 0.0;
        }
        __syncthreads();
        if (active && j0 + lane < P) {
          double dl[6], D = 0.0;
#pragma unroll
          for (int d = 0; d < 6; d++) { dl[d] = xi[d] - s12[d][lane]; D += dl[d] * dl[d]; }
          const double kij = exp(-D / h);
          ks += kij;  // :226
#pragma unroll
          for (int d = 0; d < 6; d++) { gs[d] += dl[d] * kij; kn[d] -= kij * s12[6 + d][lane]; }  // :221-224
        }
      }
      ks = warp_sum(ks);
#pragma unroll
      for (int d = 0; d < 6; d++) { gs[d] = warp_sum(gs[d]); kn[d] = warp_sum(kn[d]); }
      if (active && lane == 0) {
        const double f = 2.0 / h;
        for (int r = 0; r < 6; r++) {
          double s = 0.0;
          for (int cc = 0; cc < 6; cc++) s += s_Hinv[6 * r + cc] * (f * gs[cc]);
          a.delta[(size_t)l * 6 + r] = (kn[r] + s) / ks;  // no lr (Q5)
        }
      }
    }
  }
  TF_STAMP(4);
  grid.sync();
  TF_STAMP(5);

  // ------------------------------------------------------------------ pose update + next iteration's transforms
  const double *R0 = ia.sc.R0;
  double accc[12];
#pragma unroll
  for (int i = 0; i < 12; i++) accc[i] = 0.0;
  // particle l -> CTA l % grid, thread l / grid: a long serial fp64 chain per particle (Exp, J_l, Log, two 3x3 products), so
  // spread the particles over ALL SMs (7 threads on each of 148 SMs at P = 1000) instead of filling two CTAs
  for (int l = blockIdx.x + gridDim.x * tid; l < a.P_l; l += gn) {
    const int p = a.p_lo + l;
    double d[6];
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = a.delta[(size_t)l * 6 + i];
    double dR[9], Jl[9], R[9], Rn[9], dt[3], t[3], w[3];
    so3_exp(d + 3, dR, Jl);  // :269-271
#pragma unroll
    for (int r = 0; r < 3; r++) dt[r] = Jl[3 * r] * d[0] + Jl[3 * r + 1] * d[1] + Jl[3 * r + 2] * d[2];  // :275
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = a.R[9 * (size_t)p + i];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) Rn[3 * r + cc] = R[3 * r] * dR[cc] + R[3 * r + 1] * dR[3 + cc] + R[3 * r + 2] * dR[6 + cc];  // :277
#pragma unroll
    for (int i = 0; i < 9; i++) a.R[9 * (size_t)p + i] = Rn[i];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      t[r] = (Rn[3 * r] * dt[0] + Rn[3 * r + 1] * dt[1] + Rn[3 * r + 2] * dt[2]) + a.t[3 * (size_t)p + r];  // :278 (Q6)
      a.t[3 * (size_t)p + r] = t[r];
    }
    const double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);  // :96
    a.dnorm[l] = dn;
    // head of the next iteration (k_prep): x = [t ; Log R], fp32 transforms relative to q0
    so3_log(Rn, w);
    double *rec = a.rec + (size_t)p * REC;
#pragma unroll
    for (int i = 0; i < 3; i++) { rec[REC_X + i] = t[i]; rec[REC_X + 3 + i] = w[i]; }
    rec[REC_DNORM] = dn;
    double D[9], T[9], M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) D[i] = Rn[i] - ((i % 4 == 0) ? 1.0 : 0.0);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) T[3 * r + cc] = R0[3 * r] * D[cc] + R0[3 * r + 1] * D[3 + cc] + R0[3 * r + 2] * D[6 + cc];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) M[3 * r + cc] = T[3 * r] * R0[3 * cc] + T[3 * r + 1] * R0[3 * cc + 1] + T[3 * r + 2] * R0[3 * cc + 2];
    float *xf = ia.xf + (size_t)l * 12;
#pragma unroll
    for (int i = 0; i < 9; i++) { const float f = __double2float_rn(M[i]); xf[i] = f; accc[i] += (double)f; }
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const float f = __double2float_rn(R0[3 * r] * t[0] + R0[3 * r + 1] * t[1] + R0[3 * r + 2] * t[2]);
      xf[9 + r] = f;
      accc[9 + r] += (double)f;
    }
  }
#pragma unroll
  for (int i = 0; i < 12; i++) accc[i] = warp_sum(accc[i]);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 12; i++) s_red[warp][i] = accc[i];
  __syncthreads();
  if (tid < 12) {
    double s = 0;
    for (int w = 0; w < TF_WARPS; w++) s += s_red[w][tid];
    a.prep_scratch_d[(size_t)blockIdx.x * 12 + tid] = s;
  }
  if (gtid == 0) c->iter = it + 1;
  TF_STAMP(6);
  grid.sync();
  TF_STAMP(7);

  // centre of the slice's transforms (any centre is valid: pruning is exact around whatever centre is used)
  if (tid < 12) {
    double s = 0;
    for (int b = 0; b < (int)gridDim.x; b++) s += __ldcg(a.prep_scratch_d + (size_t)b * 12 + tid);
    s_center[tid] = __double2float_rn(s / (double)a.P_l);
  }
  for (int i = tid; i < PRUNE_BINS + 2; i += blockDim.x) s_env[i] = 0;
  if (tid == 0) s_flag[1] = 0;
  __syncthreads();
  for (int l = gtid; l < a.P_l; l += gn) {
    const float *xf = ia.xf + (size_t)l * 12;
    double M[9], db = 0;
    for (int i = 0; i < 9; i++) M[i] = (double)xf[i] - (double)s_center[i];
    for (int i = 9; i < 12; i++) { const double dd = (double)xf[i] - (double)s_center[i]; db += dd * dd; }
    const double da = sym3_max_eig_MtM(M);
    if (da != da || db != db) { s_flag[1] = 1; continue; }  // NaN state must poison the radius
    const double sg = sqrt(da) * (1.0 + 1e-6), bt = sqrt(db) * (1.0 + 1e-6);
    atomicMax(&s_env[PRUNE_BINS], __float_as_int(__double2float_ru(sg)));
    atomicMax(&s_env[PRUNE_BINS + 1], __float_as_int(__double2float_ru(bt)));
    for (int i = 0; i < PRUNE_BINS; i++)
      atomicMax(&s_env[i], __float_as_int(__double2float_ru(sg * ((double)(i + 1) * PRUNE_BIN_W) + bt)));
  }
  __syncthreads();
  for (int i = tid; i < PRUNE_BINS + 2; i += blockDim.x)
    if (s_env[i]) atomicMax(&a.prep_scratch_i[i], s_env[i]);
  if (tid == 0 && s_flag[1]) atomicMax(&a.prep_scratch_i[PRUNE_BINS + 2], 1);
  TF_STAMP(8);
  grid.sync();
  TF_STAMP(9);
  if (blockIdx.x == 0) {
    const bool nan = __ldcg(a.prep_scratch_i + PRUNE_BINS + 2) != 0;
    for (int i = tid; i < PRUNE_BINS; i += blockDim.x) c->env[i] = nan ? NAN : __int_as_float(__ldcg(a.prep_scratch_i + i));
    if (tid < 9) c->Abar[tid] = s_center[tid];
    if (tid < 3) c->taubar[tid] = s_center[9 + tid];
    if (tid == 0) {
      c->alpha = nan ? NAN : __int_as_float(__ldcg(a.prep_scratch_i + PRUNE_BINS));
      c->beta = nan ? NAN : __int_as_float(__ldcg(a.prep_scratch_i + PRUNE_BINS + 1));
      c->kept_total = 0ull;
    }
  }
}

int launch_tail_fused(const SteinArgs &a, const IterArgs &ia, cudaStream_t st) {
  static int max_blocks_per_sm = -1;
  if (max_blocks_per_sm < 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, k_tail_fused, TF_THREADS, 0);
  if (max_blocks_per_sm < 1) return -1;
  int grid = a.sm_count;  // one CTA per SM: all co-resident, as the grid barriers require
  cudaFuncSetAttribute(k_tail_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);  // per device; cheap, so every launch
  int xs_bytes = 6 * a.P * (int)sizeof(double);
  if (xs_bytes > 160 * 1024 || a.P < 2) xs_bytes = 0;  // large P: read x through L2 instead
  void *args[] = {(void *)&a, (void *)&ia, (void *)&xs_bytes};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_tail_fused, dim3(grid), dim3(TF_THREADS), args, (size_t)xs_bytes, st);
  return e == cudaSuccess ? 1 : -1;
}

}  // namespace svn

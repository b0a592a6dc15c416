// tail2.cu -- the "Stein phase" of one SVN iteration as two kernels without a grid-wide barrier on the critical path.
//
//   k_head  (side stream, small cooperative grid; overlaps the correspondence + Gauss-Newton pass of the same iteration)
//           early-stop decision for the previous update + its history row          SVNICP.cpp:95-107
//           exact lower median of the P^2 pairwise squared distances -> bandwidth h   SVNICP.cpp:254-266
//           Everything here depends only on the particle positions x, which are final as soon as the previous update is.
//   k_tail  (main stream, ordinary launch, one CTA per NI particles)
//           Stein step: full SVN / pre-conditioned SVGD / P == 1                    SVNICP.cpp:218-252, 88-89
//           pose update                                                             SVNICP.cpp:268-279
//           next iteration's x = [t ; Log R], fp32 transforms, exact-pruning ball   SVNICP.cpp:58-59,74-77
//           Records of all particles stream through shared memory as AoS tiles (1-D bulk TMA + mbarrier ring); the CTA
//           that finishes last publishes the ball and the iteration counter.
//
// Sharded handles: k_finalize stores (b, H, g) and k_tail stores (x, |delta|) of the local particles straight into every
// rank's record buffer over NVLink (PeerTable, common.cuh), followed by a sequence number in the peer's flag block; k_tail /
// k_head spin on their local flags.  Records are double buffered by iteration parity, which is what makes the exchange
// safe without a second handshake.  No NCCL call on the per-iteration path.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace svn {

// ---------------------------------------------------------------------------------------------
// k_head
// ---------------------------------------------------------------------------------------------
constexpr int HD_THREADS = 512;
constexpr int HD_WARPS = HD_THREADS / 32;

__device__ __forceinline__ int hd_pass_bits(int s) { return s == 0 ? 11 : 13; }
__device__ __forceinline__ int hd_bits_before(int s) { return s == 0 ? 0 : 11 + 13 * (s - 1); }

// (prefix, rank) after a pass from its finished global histogram; every CTA computes the same values
__device__ void hd_select(const unsigned *hist, unsigned long long prefix_in, unsigned long long rank_in, int nbins, int bits,
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

__global__ void __launch_bounds__(HD_THREADS, 1) k_head(SteinArgs a, PeerTable pt, unsigned seq_x, int epilogue, int xs_smem_bytes) {
  cg::grid_group grid = cg::this_grid();
  Ctrl *c = a.ctrl;
  if (c->stop) return;  // set by an earlier launch: identical for every CTA
  extern __shared__ __align__(16) unsigned char s_dyn[];  // optional copy of x [6][P] for the median passes
  __shared__ __align__(16) unsigned s_hist[MED_BINS];     // 32 KB: histogram / collected candidates of the median bin
  __shared__ double s_red[HD_WARPS];
  __shared__ unsigned long long s_warp[32], s_res[2];
  __shared__ int s_flag[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gtid = blockIdx.x * blockDim.x + tid, gn = gridDim.x * blockDim.x;
  const int P = a.P;
  // x of every particle after the previous update must have arrived (the first iteration's x is computed locally by k_prep)
  if (seq_x) peer_wait(pt, FLAG_X, seq_x, c);
  const int it = c->iter;
  const double *rec = a.rec + (size_t)(it & 1) * a.rec_stride;
  // previous iteration's median (the bandwidth is rewritten only at the very end): the guess of the fast median path
  const double med_guess = c->bandwidth * log((double)(P + 1));

  // ------------------------------------------------------------------ decide (redundantly per CTA: same data, same order)
  {
    double s = 0.0;
    for (int p = tid; p < P; p += blockDim.x) s += __ldcg(rec + (size_t)p * REC + REC_DNORM);
    s = warp_sum(s);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < HD_WARPS; w++) tot += s_red[w];
      const int stop = (a.check_early_stop && it > 0 && tot / (double)P < a.threshold) ? 1 : 0;  // SVNICP.cpp:95-101
      s_flag[0] = stop;
      if (blockIdx.x == 0) {
        if (stop) { c->stop = 1; c->iters_done = it; }
        else if (epilogue) c->iters_done = it;
      }
    }
    __syncthreads();
    if (s_flag[0]) return;  // break BEFORE the history row of that iteration (Q9); every CTA decides identically
    if (it > 0 && it - 1 < a.I) {
      float *row = a.history + (size_t)(it - 1) * 6 * P;  // SVNICP.cpp:103-107
      for (int i = gtid; i < 6 * P; i += gn) {
        const int comp = i / P, p = i % P;
        row[i] = (float)__ldcg(rec + (size_t)p * REC + REC_X + comp);
      }
    }
    if (epilogue || P < 2) return;
    for (int i = gtid; i < 6 * P; i += gn) {
      const int comp = i / P, p = i % P;
      a.xs[i] = __ldcg(rec + (size_t)p * REC + REC_X + comp);
    }
    for (int i = gtid; i < MED_PASSES * MED_BINS; i += gn) a.hist[i] = 0u;
  }
  grid.sync();

  // ------------------------------------------------------------------ bandwidth: exact lower median
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
  constexpr int COLLECT_CAP = 4096;  // doubles; staged in s_hist (32 KB)
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
    hd_select(a.hist, 0ull, rank, MED_BINS, 13, &tbin, &trank, s_warp, s_res);
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
            if (k < (unsigned)COLLECT_CAP) list[k] = d2;
          }
        }
      grid.sync();
      const unsigned n_c = __ldcg(cursor);
      if (n_c >= 1u && n_c <= (unsigned)COLLECT_CAP) {
        double *sl = reinterpret_cast<double *>(s_hist);
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
    const int nb = 1 << hd_pass_bits(s);
    for (int i = tid; i < nb; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const int consumed = hd_bits_before(s);
    const int shift = 63 - consumed - hd_pass_bits(s);
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
    hd_select(gh, prefix, rank, nb, hd_pass_bits(s), &prefix, &rank, s_warp, s_res);
  }
  if (!have_median) median = __longlong_as_double((long long)prefix);
  if (gtid == 0) c->bandwidth = median / log((double)(P + 1));  // SVNICP.cpp:262 (Q4)
}

// ---------------------------------------------------------------------------------------------
// k_tail
// ---------------------------------------------------------------------------------------------
constexpr int TL_CONSUMERS = 256;
constexpr int TL_WARPS = TL_CONSUMERS / 32;
constexpr int TL_THREADS = TL_CONSUMERS + 32;  // + one producer warp (bulk TMA)
constexpr int TL_ENV_CHUNK = 1024;             // particles per pass of the envelope reduction (last CTA)

// dynamic shared memory: [stages][tile_records * REC] doubles, then the mbarriers
__global__ void __launch_bounds__(TL_THREADS, 1) k_tail(SteinArgs a, IterArgs ia, PeerTable pt, unsigned seq_h, unsigned seq_x_out, int NI,
                                                        int JQ, int stages) {
  Ctrl *c = a.ctrl;
  if (c->stop) return;  // decided by k_head of this iteration (ordered before this launch) or earlier
  extern __shared__ __align__(128) unsigned char s_dyn[];
  __shared__ double s_part[TL_WARPS][28];
  __shared__ double s_Hbar[TL_WARPS][21];
  __shared__ double s_Hinv[36];
  __shared__ double s_xf[TL_WARPS][12];
  __shared__ float s_center[12];
  __shared__ int s_envmax[4][PRUNE_BINS];
  __shared__ int s_ab[2];
  __shared__ int s_flag[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P;
  // phase stamps (ns, tuning aid): CTA 0 stamps 0..3, the last CTA 4..5
#define TL_STAMP(k) do { if (a.stamps && tid == 0) a.stamps[k] = (double)global_timer_ns(); } while (0)
  if (blockIdx.x == 0) TL_STAMP(0);
  const int TJ = 32 * JQ;                                   // records per tile
  const size_t tile_bytes = (size_t)TJ * REC * sizeof(double);
  uint64_t *full = reinterpret_cast<uint64_t *>(s_dyn + (size_t)stages * tile_bytes);
  uint64_t *empty = full + stages;
  const int n_tiles = (P + TJ - 1) / TJ;

  if (tid == 0) {
    for (int s = 0; s < stages; s++) { mbar_init(full + s, 1); mbar_init(empty + s, TL_WARPS); }
    fence_barrier_init();
  }
  // (b, H, g) of every particle must have arrived; x arrived before k_head of this iteration ran
  peer_wait(pt, FLAG_H, seq_h, c);  // ends with __syncthreads: also publishes the barrier init
  const int it = c->iter;
  const double *rec = a.rec + (size_t)(it & 1) * a.rec_stride;
  const size_t nxt_off = (size_t)((it + 1) & 1) * a.rec_stride;
  const double h = c->bandwidth;
  if (blockIdx.x == 0) TL_STAMP(1);

  if (warp == TL_WARPS) {
    // ---------------- producer warp: AoS record tiles, one bulk copy each ----------------
    if (P >= 2)
      for (int t = 0; t < n_tiles; t++) {
        const int s = t % stages, k = t / stages;
        mbar_wait(empty + s, (uint32_t)((k & 1) ^ 1));
        if (lane == 0) {
          mbar_expect_tx(full + s, (uint32_t)tile_bytes);
          bulk_g2s(s_dyn + (size_t)s * tile_bytes, rec + (size_t)t * TJ * REC, (uint32_t)tile_bytes, full + s);
        }
        __syncwarp();
      }
  } else {
    // ---------------- consumers: warp = (particle ii, j-quarter jq); lane = j ----------------
    const int ii = warp / JQ, jq = warp % JQ;
    const int l = blockIdx.x * NI + ii;
    const bool active = ii < NI && l < a.P_l;
    const int i = a.p_lo + (active ? l : 0);
    double xi[6];
#pragma unroll
    for (int d = 0; d < 6; d++) xi[d] = __ldcg(rec + (size_t)i * REC + REC_X + d);
    if (P >= 2 && a.svn_full_grad) {
      double Hm[21], v[6];
#pragma unroll
      for (int q = 0; q < 21; q++) Hm[q] = 0.0;
#pragma unroll
      for (int q = 0; q < 6; q++) v[q] = 0.0;
      const double two_over_h = 2.0 / h;
      for (int t = 0; t < n_tiles; t++) {
        const int s = t % stages, k = t / stages;
        mbar_wait(full + s, (uint32_t)(k & 1));
        const double *tile = reinterpret_cast<const double *>(s_dyn + (size_t)s * tile_bytes);
        const int jj = jq * 32 + lane;
        if (active && t * TJ + jj < P) {
          const double *r = tile + (size_t)jj * REC;
          double dl[6], D = 0.0;
#pragma unroll
          for (int d = 0; d < 6; d++) { dl[d] = xi[d] - r[REC_X + d]; D += dl[d] * dl[d]; }
          const double kij = exp(-D / h);  // :264
          const double k2 = kij * kij;     // :238
          double gv[6];
#pragma unroll
          for (int d = 0; d < 6; d++) gv[d] = two_over_h * (dl[d] * kij);  // :233
          int q = 0;
#pragma unroll
          for (int rr = 0; rr < 6; rr++)
#pragma unroll
            for (int cc = rr; cc < 6; cc++, q++) Hm[q] += k2 * r[REC_H + q] + gv[rr] * gv[cc];  // :236-242
#pragma unroll
          for (int d = 0; d < 6; d++) v[d] += gv[d] - kij * r[REC_B + d];  // :244 with b' = -b
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
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
    } else if (P >= 2) {
      // pre-conditioned SVGD (SVNICP.cpp:85, :218-227); the warps of particle 0 of the CTA also sum H over all j (mean Hessian)
      double gs[6], kn[6], ks = 0.0, Hs[21];
#pragma unroll
      for (int d = 0; d < 6; d++) { gs[d] = 0.0; kn[d] = 0.0; }
#pragma unroll
      for (int q = 0; q < 21; q++) Hs[q] = 0.0;
      for (int t = 0; t < n_tiles; t++) {
        const int s = t % stages, k = t / stages;
        mbar_wait(full + s, (uint32_t)(k & 1));
        const double *tile = reinterpret_cast<const double *>(s_dyn + (size_t)s * tile_bytes);
        const int jj = jq * 32 + lane;
        if (t * TJ + jj < P) {
          const double *r = tile + (size_t)jj * REC;
          if (ii == 0) {
#pragma unroll
            for (int q = 0; q < 21; q++) Hs[q] += r[REC_H + q];
          }
          if (active) {
            double dl[6], D = 0.0;
#pragma unroll
            for (int d = 0; d < 6; d++) { dl[d] = xi[d] - r[REC_X + d]; D += dl[d] * dl[d]; }
            const double kij = exp(-D / h);
            ks += kij;  // :226
#pragma unroll
            for (int d = 0; d < 6; d++) { gs[d] += dl[d] * kij; kn[d] -= kij * r[REC_G + d]; }  // :221-224
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
      }
      ks = warp_sum(ks);
#pragma unroll
      for (int d = 0; d < 6; d++) { gs[d] = warp_sum(gs[d]); kn[d] = warp_sum(kn[d]); }
      if (ii == 0) {
#pragma unroll
        for (int q = 0; q < 21; q++) Hs[q] = warp_sum(Hs[q]);
      }
      if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 6; d++) { s_part[warp][d] = gs[d]; s_part[warp][6 + d] = kn[d]; }
        s_part[warp][12] = ks;
        if (ii == 0)
#pragma unroll
          for (int q = 0; q < 21; q++) s_Hbar[jq][q] = Hs[q];
      }
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) TL_STAMP(2);
  if (P >= 2 && !a.svn_full_grad && tid == 0) {
    double A0[36];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = r; cc < 6; cc++) {
        double s = 0.0;
        for (int w = 0; w < JQ; w++) s += s_Hbar[w][tri(r, cc)];
        A0[6 * r + cc] = s / (double)P;  // :85 mean over particles
        A0[6 * cc + r] = A0[6 * r + cc];
      }
    // :225 inverse, column by column (each solve register resident; the pivot sequence is the same for every column)
#pragma unroll
    for (int col = 0; col < 6; col++) {
      double A[36], e[6];
#pragma unroll
      for (int q = 0; q < 36; q++) A[q] = A0[q];
#pragma unroll
      for (int q = 0; q < 6; q++) e[q] = (q == col) ? 1.0 : 0.0;
      lu_solve6_reg(A, e);
#pragma unroll
      for (int q = 0; q < 6; q++) s_Hinv[6 * q + col] = e[q];
    }
  }
  if (P >= 2 && !a.svn_full_grad) __syncthreads();

  // ---------------- one thread per particle: solve, pose update, head of the next iteration ----------------
  const double *R0 = ia.sc.R0;
  if (warp < TL_WARPS && lane == 0 && (warp % JQ) == 0) {
    const int ii = warp / JQ;
    const int l = blockIdx.x * NI + ii;
    double xfd[12];
#pragma unroll
    for (int k = 0; k < 12; k++) xfd[k] = 0.0;
    if (ii < NI && l < a.P_l) {
      const int p = a.p_lo + l;
      double d[6];
      if (P < 2) {  // SVNICP.cpp:88-89
        double A[36];
        const double *r = rec + (size_t)p * REC;
#pragma unroll
        for (int rr = 0; rr < 6; rr++)
#pragma unroll
          for (int cc = rr; cc < 6; cc++) { A[6 * rr + cc] = __ldcg(r + REC_H + tri(rr, cc)); A[6 * cc + rr] = A[6 * rr + cc]; }
#pragma unroll
        for (int q = 0; q < 6; q++) d[q] = __ldcg(r + REC_B + q);
        lu_solve6_reg(A, d);
        for (int q = 0; q < 6; q++) d[q] = -d[q];
      } else if (a.svn_full_grad) {
        double A[36];
#pragma unroll
        for (int rr = 0; rr < 6; rr++)
#pragma unroll
          for (int cc = rr; cc < 6; cc++) {
            double sum = 0.0;
            for (int w = 0; w < JQ; w++) sum += s_part[ii * JQ + w][tri(rr, cc)];
            A[6 * rr + cc] = sum / (double)P;
            A[6 * cc + rr] = A[6 * rr + cc];
          }
#pragma unroll
        for (int q = 0; q < 6; q++) {
          double sum = 0.0;
          for (int w = 0; w < JQ; w++) sum += s_part[ii * JQ + w][21 + q];
          d[q] = sum / (double)P;
        }
        lu_solve6_reg(A, d);  // :250 (the reference forms the explicit inverse; tolerance-level difference)
        for (int q = 0; q < 6; q++) d[q] = a.lr * d[q];
      } else {
        double gs[6], kn[6], ks = 0.0;
        for (int q = 0; q < 6; q++) { gs[q] = 0.0; kn[q] = 0.0; }
        for (int w = 0; w < JQ; w++) {
          for (int q = 0; q < 6; q++) { gs[q] += s_part[ii * JQ + w][q]; kn[q] += s_part[ii * JQ + w][6 + q]; }
          ks += s_part[ii * JQ + w][12];
        }
        const double f = 2.0 / h;
        for (int rr = 0; rr < 6; rr++) {
          double s = 0.0;
          for (int cc = 0; cc < 6; cc++) s += s_Hinv[6 * rr + cc] * (f * gs[cc]);
          d[rr] = (kn[rr] + s) / ks;  // no lr (Q5)
        }
      }
#pragma unroll
      for (int q = 0; q < 6; q++) a.delta[(size_t)l * 6 + q] = d[q];
      // pose update, SVNICP.cpp:268-279
      double dR[9], Jl[9], R[9], Rn[9], dt[3], t[3], w[3];
      so3_exp(d + 3, dR, Jl);  // :269-271
#pragma unroll
      for (int r = 0; r < 3; r++) dt[r] = Jl[3 * r] * d[0] + Jl[3 * r + 1] * d[1] + Jl[3 * r + 2] * d[2];  // :275
#pragma unroll
      for (int q = 0; q < 9; q++) R[q] = a.R[9 * (size_t)p + q];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) Rn[3 * r + cc] = R[3 * r] * dR[cc] + R[3 * r + 1] * dR[3 + cc] + R[3 * r + 2] * dR[6 + cc];  // :277
#pragma unroll
      for (int q = 0; q < 9; q++) a.R[9 * (size_t)p + q] = Rn[q];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        t[r] = (Rn[3 * r] * dt[0] + Rn[3 * r + 1] * dt[1] + Rn[3 * r + 2] * dt[2]) + a.t[3 * (size_t)p + r];  // :278 (Q6)
        a.t[3 * (size_t)p + r] = t[r];
      }
      const double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);  // :96
      a.dnorm[l] = dn;
      // head of the next iteration: x = [t ; Log R] and |delta| into every rank's NEXT record buffer
      so3_log(Rn, w);
      for (int r = 0; r < pt.n_ranks; r++) {
        double *o = pt.rec[r] + nxt_off + (size_t)p * REC;
#pragma unroll
        for (int q = 0; q < 3; q++) { o[REC_X + q] = t[q]; o[REC_X + 3 + q] = w[q]; }
        o[REC_DNORM] = dn;
      }
      if (pt.n_ranks > 1) __threadfence_system();
      // fp32 transforms relative to q0 (as k_prep)
      double D[9], T[9], M[9];
#pragma unroll
      for (int q = 0; q < 9; q++) D[q] = Rn[q] - ((q % 4 == 0) ? 1.0 : 0.0);
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
      for (int q = 0; q < 9; q++) { const float f = __double2float_rn(M[q]); xf[q] = f; xfd[q] = (double)f; }
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float f = __double2float_rn(R0[3 * r] * t[0] + R0[3 * r + 1] * t[1] + R0[3 * r + 2] * t[2]);
        xf[9 + r] = f;
        xfd[9 + r] = (double)f;
      }
    }
    if (ii < TL_WARPS)
#pragma unroll
      for (int k = 0; k < 12; k++) s_xf[ii][k] = xfd[k];
  }
  __syncthreads();
  if (blockIdx.x == 0) TL_STAMP(3);
  // per-CTA partial of the centre transform, fixed order; then the ticket: the last CTA publishes
  if (tid < 12) {
    double s = 0.0;
    for (int ii = 0; ii < NI; ii++) s += s_xf[ii][tid];
    a.prep_scratch_d[(size_t)blockIdx.x * 12 + tid] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag[0] = (atomicAdd(&c->tail_ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_flag[0]) return;
  __threadfence();
  TL_STAMP(4);

  // ---------------- last CTA: exact-pruning ball of the local slice for the next iteration (as k_prep) ----------------
  if (tid < 12) {
    double s = 0.0;
    for (int b = 0; b < (int)gridDim.x; b++) s += __ldcg(a.prep_scratch_d + (size_t)b * 12 + tid);
    s_center[tid] = __double2float_rn(s / (double)a.P_l);
  }
  for (int q = tid; q < 4 * PRUNE_BINS; q += blockDim.x) s_envmax[q / PRUNE_BINS][q % PRUNE_BINS] = 0;
  if (tid < 2) { s_ab[tid] = 0; s_flag[1] = 0; }
  __syncthreads();
  {
    float *s_sg = reinterpret_cast<float *>(s_dyn);  // the tile ring is idle now: [TL_ENV_CHUNK] sigma, [TL_ENV_CHUNK] beta
    float *s_bt = s_sg + TL_ENV_CHUNK;
    int emax = 0;  // this thread's bin (tid % 64) over its share (tid / 64) of the chunk
    for (int l0 = 0; l0 < a.P_l; l0 += TL_ENV_CHUNK) {
      const int n = min(TL_ENV_CHUNK, a.P_l - l0);
      for (int u = tid; u < n; u += blockDim.x) {
        const float *xf = ia.xf + (size_t)(l0 + u) * 12;
        double M[9], db = 0;
        for (int q = 0; q < 9; q++) M[q] = (double)__ldcg(xf + q) - (double)s_center[q];
        for (int q = 9; q < 12; q++) { const double dd = (double)__ldcg(xf + q) - (double)s_center[q]; db += dd * dd; }
        const double da = sym3_max_eig_MtM(M);
        float sg = 0.f, bt = 0.f;
        if (da != da || db != db) s_flag[1] = 1;  // NaN state must poison the radius
        else { sg = __double2float_ru(sqrt(da) * (1.0 + 1e-6)); bt = __double2float_ru(sqrt(db) * (1.0 + 1e-6)); }
        s_sg[u] = sg;
        s_bt[u] = bt;
        atomicMax(&s_ab[0], __float_as_int(sg));  // non-negative floats order like their bit patterns
        atomicMax(&s_ab[1], __float_as_int(bt));
      }
      __syncthreads();
      if (tid < 4 * PRUNE_BINS) {
        const int bin = tid % PRUNE_BINS, part = tid / PRUNE_BINS;
        const double r = (double)(bin + 1) * PRUNE_BIN_W;
        for (int u = part; u < n; u += 4) {
          const float vv = __double2float_ru((double)s_sg[u] * r + (double)s_bt[u]);
          emax = max(emax, __float_as_int(vv));
        }
      }
      __syncthreads();
    }
    if (tid < 4 * PRUNE_BINS) s_envmax[tid / PRUNE_BINS][tid % PRUNE_BINS] = emax;
    __syncthreads();
  }
  const bool nan = s_flag[1] != 0;
  for (int q = tid; q < PRUNE_BINS; q += blockDim.x) {
    const int m = max(max(s_envmax[0][q], s_envmax[1][q]), max(s_envmax[2][q], s_envmax[3][q]));
    c->env[q] = nan ? NAN : __int_as_float(m);
  }
  if (tid < 9) c->Abar[tid] = s_center[tid];
  if (tid < 3) c->taubar[tid] = s_center[9 + tid];
  if (tid == 0) {
    c->alpha = nan ? NAN : __int_as_float(s_ab[0]);
    c->beta = nan ? NAN : __int_as_float(s_ab[1]);
    if (a.kept_hist && it <= a.I) a.kept_hist[it] = c->kept_total;  // candidates kept by this iteration's pruning pass
    c->kept_total = 0ull;
    c->tail_ticket = 0u;
    c->iter = it + 1;
  }
  TL_STAMP(5);
#undef TL_STAMP
  // every CTA fenced its peer stores before taking its ticket: the new x of the whole slice is out -> tell the peers
  if (pt.n_ranks > 1 && tid < pt.n_ranks && tid != pt.rank) {
    __threadfence_system();
    st_release_sys(pt.flag[tid] + FLAG_X * MAX_RANKS + pt.rank, seq_x_out);
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int head_grid(int P, int sm_count) {
  // the median sweeps P^2/2 pairs: a few fat CTAs while that is small (cheap grid barrier, leaves the SMs to k_gn), all SMs for large P
  long long pairs = (long long)P * P / 2;
  long long g = (pairs + 32767) / 32768;
  if (g < 8) g = 8;
  if (g > sm_count) g = sm_count;
  if (P < 2) g = 1;
  return (int)g;
}

int launch_head(const SteinArgs &a, const PeerTable &pt, unsigned seq_x, int epilogue, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_head, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024); attr = true; }
  int grid = epilogue ? 4 : head_grid(a.P, a.sm_count);
  int xs_bytes = 6 * a.P * (int)sizeof(double);
  if (xs_bytes > 144 * 1024 || a.P < 2 || epilogue) xs_bytes = 0;  // large P: read x through L2 instead
  SteinArgs aa = a;
  PeerTable pp = pt;
  void *args[] = {(void *)&aa, (void *)&pp, (void *)&seq_x, (void *)&epilogue, (void *)&xs_bytes};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_head, dim3(grid), dim3(HD_THREADS), args, (size_t)xs_bytes, st);
  return e == cudaSuccess ? 1 : -1;
}

void tail_shape(int P_l, int sm_count, int *NI, int *JQ, int *stages, size_t *smem) {
  // NI particles x JQ warps per CTA (NI * JQ = 8): the largest NI that still gives ~one CTA per SM -- small slices get more
  // warps per particle (shorter chain over j), large ones share each record tile among more particles (less L2 traffic)
  int ni = 8;
  while (ni > 1 && (P_l + ni - 1) / ni < (sm_count * 4) / 5) ni >>= 1;
  const int jq = TL_WARPS / ni;
  const size_t tile = (size_t)32 * jq * REC * sizeof(double);
  int s = (int)((150 * 1024) / tile);
  if (s > 12) s = 12;
  if (s < 1) s = 1;
  size_t bytes = (size_t)s * tile;
  if (bytes < (size_t)2 * TL_ENV_CHUNK * sizeof(float)) bytes = (size_t)2 * TL_ENV_CHUNK * sizeof(float);
  *NI = ni; *JQ = jq; *stages = s;
  *smem = bytes + 2 * (size_t)s * sizeof(uint64_t) + 128;
}

int launch_tail(const SteinArgs &a, const IterArgs &ia, const PeerTable &pt, unsigned seq_h, unsigned seq_x_out, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  int NI, JQ, stages;
  size_t smem;
  tail_shape(a.P_l, a.sm_count, &NI, &JQ, &stages, &smem);
  const int grid = (a.P_l + NI - 1) / NI;
  if (grid < 1) return 0;
  k_tail<<<grid, TL_THREADS, smem, st>>>(a, ia, pt, seq_h, seq_x_out, NI, JQ, stages);
  return 1;
}

}  // namespace svn

// gn_pair.cu -- packed-fp32x2 variant of the correspondence + Gauss-Newton pass (north_star kernels (a)+(b)).
//
// Same arithmetic, operation for operation, as k_filter / k_gn in iter_kernels.cu (and therefore the same bit-exact
// correspondence indices; oracle_corr_f32 restates it), but every thread processes TWO source points at once in the
// two halves of Blackwell's packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2: one issue slot, two IEEE-rn results).
// k_gn is issue-bound (profiles/: issue slots 67 % busy, FMA pipe 44 %), so halving the FP issue slots is the lever.
//
// Layout ("pair mode"): rows 2m and 2m+1 form a pair.
//   spair [n_pad/2][2] float4 : (sx0,sx1,sy0,sy1) (sz0,sz1,sw0,sw1)
//   plist [n_pad/2][Kp][2] float4 : slot k = (x0,x1,y0,y1) (z0,z1,w0,w1), the k-th KEPT candidate of each row; the shorter
//                                  list is padded with +inf sentinels; pcount[m] = padded common length (even), 0 = no rows
// Replaces, per iteration: SVNICP.cpp:58-71,116-164; SVGDICP.cpp:300-333; knn.cu:204-251 (reference svn-icp/src/core).
#include "common.cuh"
#include "kernels.h"
#include "ptx_helpers.cuh"

namespace svn {

constexpr float PRUNE_MARGIN_P = 1e-4f;  // metres (same bound as k_filter)

// sp (row layout) -> interleaved pairs, once per scan
__global__ void k_spair(const float4 *__restrict__ sp, float4 *__restrict__ spair, int n_pairs) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_pairs) return;
  const float4 a = sp[2 * m], b = sp[2 * m + 1];
  spair[2 * m] = make_float4(a.x, b.x, a.y, b.y);
  spair[2 * m + 1] = make_float4(a.z, b.z, a.w, b.w);
}

// ---------------------------------------------------------------------------------------------
// k_filter_pair: exact candidate pruning (see k_filter), one warp per row PAIR, interleaved output
// ---------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256) k_filter_pair(IterArgs a) {
  if (a.ctrl->stop) return;
  const int lane = lane_id();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const Ctrl *c = a.ctrl;
  float A[9], tb[3];
#pragma unroll
  for (int i = 0; i < 9; i++) A[i] = c->Abar[i];
#pragma unroll
  for (int i = 0; i < 3; i++) tb[i] = c->taubar[i];
  const float alpha = c->alpha, beta = c->beta;
  const int K = a.K, Kp = a.Kp;
  const unsigned lt = (1u << lane) - 1u;
  const int n_pairs = a.n_pad >> 1;
  unsigned long long kept = 0;
  for (int m = gw; m < n_pairs; m += nw) {
    float *out = reinterpret_cast<float *>(a.clist + (size_t)m * Kp * 2);
    int cnt[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int b = 2 * m + h;
      if (b >= a.n_s) continue;  // warp-uniform
      const float4 s = a.sp[b];
      const float qx = fmaf(A[0], s.x, fmaf(A[1], s.y, fmaf(A[2], s.z, tb[0])));
      const float qy = fmaf(A[3], s.x, fmaf(A[4], s.y, fmaf(A[5], s.z, tb[1])));
      const float qz = fmaf(A[6], s.x, fmaf(A[7], s.y, fmaf(A[8], s.z, tb[2])));
      float rho = fmaf(alpha, s.w, beta);
      const int rbin = (int)(s.w * (float)(1.0 / PRUNE_BIN_W));
      if (rbin < PRUNE_BINS) rho = fminf(rho, c->env[rbin]);
      const float4 *row = a.cand + (size_t)b * K;
      float4 e[NCH];
      float d2[NCH];
      float dmin = INFINITY;
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        const int k = ch * 32 + lane;
        if (k < K) {
          e[ch] = __ldcs(row + k);
          const float dx = qx - e[ch].x, dy = qy - e[ch].y, dz = qz - e[ch].z;
          d2[ch] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          dmin = fminf(dmin, d2[ch]);
        } else {
          d2[ch] = INFINITY;
        }
      }
      dmin = warp_min(dmin);
      const float lim = sqrtf(dmin) + 2.0f * rho + PRUNE_MARGIN_P;
      const float thr = lim * lim * (1.0f + 1e-5f);
      int base = 0;
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        const int k = ch * 32 + lane;
        const bool keep = (k < K) && !(d2[ch] > thr);  // NaN anywhere keeps everything (NaN must propagate)
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          float *o = out + (size_t)(base + __popc(mk & lt)) * 8 + h;
          o[0] = e[ch].x; o[2] = e[ch].y; o[4] = e[ch].z; o[6] = e[ch].w;
        }
        base += __popc(mk);
      }
      cnt[h] = base;
      kept += (unsigned long long)base;
    }
    const int L = max(cnt[0], cnt[1]);
    const int Lp = (L + 1) & ~1;
    // +inf sentinels can never win (d = inf, strict '<') nor block the exit test (w = inf).  A row beyond the cloud
    // (odd N_s: second half of the last pair) gets FINITE far coordinates instead, so that its (discarded) residual
    // stays finite; k_gn_pair zeroes that half's contribution.
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float far = (2 * m + h >= a.n_s) ? 1e18f : INFINITY;
      for (int r = cnt[h] + lane; r < Lp; r += 32) {
        float *o = out + (size_t)r * 8 + h;
        o[0] = far; o[2] = far; o[4] = far; o[6] = INFINITY;
      }
    }
    if (lane == 0) a.ccount[m] = Lp;
  }
  if (lane == 0 && kept) atomicAdd(&a.ctrl->kept_total, kept);
}

// ---------------------------------------------------------------------------------------------
// k_gn_pair: fused transform + 1-NN + robust weight + Gauss-Newton reduction, two rows per thread step
// ---------------------------------------------------------------------------------------------
constexpr int GP_CONSUMERS = 512;
constexpr int GP_THREADS = GP_CONSUMERS;  // 16 warps x 128 registers fill the four 16K-register SMSP files exactly;
                                          // a 17th (producer) warp would not fit, so warp 0 also issues the TMA copies
constexpr int GP_FLUSH_PAIRS = 16;  // 32 rows per first-level flush
constexpr int GP_FLUSH2 = 64;
constexpr int NACCP = 19;           // 16 sums of k_gn, the cross-product sums split into + and - parts

// issue the 1-D TMA bulk copies of tile j of this CTA into its stage (executed by all lanes of warp 0)
__device__ __forceinline__ void gp_produce(const IterArgs &a, unsigned char *smem, uint64_t *full, uint64_t *empty, int j, int S,
                                           size_t stage_bytes, int slice, int n_slices, int TB, int TP, int Kp, int lane) {
  const int s = j % S, k = j / S;
  mb_wait(empty + s, (uint32_t)((k & 1) ^ 1));  // every consumer warp has released the previous tile of this stage
  const int pair0 = (slice + j * n_slices) * TP;
  unsigned char *st = smem + (size_t)s * stage_bytes;
  const int cnt = (lane < TP) ? a.ccount[pair0 + lane] : 0;
  const int bytes = cnt * 32;
  int total = bytes;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  total += TP * 32 + TP * 4;
  if (lane == 0) mb_expect_tx(full + s, (uint32_t)total);
  __syncwarp();
  if (lane < TP && cnt > 0) tma_bulk_g2s(st + (size_t)lane * Kp * 32, a.clist + (size_t)(pair0 + lane) * Kp * 2, (uint32_t)bytes, full + s);
  if (lane == 0) {
    tma_bulk_g2s(st + (size_t)TB * Kp * 16, a.spair + (size_t)pair0 * 2, (uint32_t)(TP * 32), full + s);
    tma_bulk_g2s(st + (size_t)TB * Kp * 16 + (size_t)TB * 16, a.ccount + pair0, (uint32_t)(TP * 4), full + s);
  }
}

template <bool DBG>
__global__ void __launch_bounds__(GP_THREADS, 1) k_gn_pair(IterArgs a) {
  if (a.ctrl->stop) return;
  extern __shared__ __align__(128) unsigned char smem[];
  const int TB = a.TB, Kp = a.Kp, S = a.stages;
  const int TP = TB >> 1;  // row pairs per tile
  const size_t stage_bytes = gn_stage_bytes(TB, Kp);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes);
  uint64_t *empty = full + S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = a.n_pad / TB;
  const int slice = blockIdx.x, n_slices = gridDim.x;
  const int n_my = (slice < n_tiles) ? (n_tiles - slice + n_slices - 1) / n_slices : 0;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mb_init(full + s, 1); mb_init(empty + s, GP_CONSUMERS / 32); }
    mb_fence_init();
  }
  __syncthreads();
  // prologue: S-1 tiles in flight
  if (warp == 0)
    for (int j = 0; j < S - 1 && j < n_my; j++) gp_produce(a, smem, full, empty, j, S, stage_bytes, slice, n_slices, TB, TP, Kp, lane);

  // ---------------- consumers: one thread = one particle (x RG pair groups) ----------------
  const int PG = a.PG, RG = a.RG;
  const int pl = tid % PG, rg = tid / PG;
  const int l = blockIdx.y * PG + pl;  // local particle index
  const bool active = l < a.P_l;
  f2 A0, A1, A2, A3, A4, A5, A6, A7, A8, T0, T1, T2;
  {
    float x[12];
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = active ? a.xf[(size_t)l * 12 + i] : 0.f;
    A0 = pk(x[0], x[0]); A1 = pk(x[1], x[1]); A2 = pk(x[2], x[2]); A3 = pk(x[3], x[3]); A4 = pk(x[4], x[4]); A5 = pk(x[5], x[5]);
    A6 = pk(x[6], x[6]); A7 = pk(x[7], x[7]); A8 = pk(x[8], x[8]); T0 = pk(x[9], x[9]); T1 = pk(x[10], x[10]); T2 = pk(x[11], x[11]);
  }
  const float Dm = a.max_dist;
  f2 acc[NACCP];
  float acc2[NACC];
#pragma unroll
  for (int i = 0; i < NACCP; i++) acc[i] = 0ull;
#pragma unroll
  for (int i = 0; i < NACC; i++) acc2[i] = 0.f;
  int pairs_in_acc = 0, flushes2 = 0;
  bool wrote = false;
  double *out = a.part + (((size_t)slice * RG + rg) * a.P_l + (active ? l : 0)) * NACC;

// one slot = the k-th kept candidate of both rows; fixed operation order (index parity with oracle_corr_f32);
// strict '<', first slot wins (mink.cuh:141)
#define SVN_EVAL2(CXY, CZW, SLOT)                                                           \
  {                                                                                         \
    const f2 dx_ = sub2(qx, (CXY).x), dy_ = sub2(qy, (CXY).y), dz_ = sub2(qz, (CZW).x);     \
    const f2 d_ = fma2(dz_, dz_, fma2(dy_, dy_, mul2(dx_, dx_)));                           \
    const float d0_ = lo32(d_), d1_ = hi32(d_);                                             \
    if (d0_ < best0) { best0 = d0_; bi0 = (SLOT); }                                         \
    if (d1_ < best1) { best1 = d1_; bi1 = (SLOT); }                                         \
  }

  for (int i = 0; i < n_my; i++) {
    if (warp == 0 && i + S - 1 < n_my) gp_produce(a, smem, full, empty, i + S - 1, S, stage_bytes, slice, n_slices, TB, TP, Kp, lane);
    const int s = i % S, k = i / S;
    mb_wait(full + s, (uint32_t)(k & 1));
    const unsigned char *st = smem + (size_t)s * stage_bytes;
    const ulonglong2 *srcp = reinterpret_cast<const ulonglong2 *>(st + (size_t)TB * Kp * 16);
    const int *cnt = reinterpret_cast<const int *>(st + (size_t)TB * Kp * 16 + (size_t)TB * 16);
    for (int pr = rg; pr < TP; pr += RG) {
      const int n = cnt[pr];  // common padded length (even); 0 = padding pair
      if (n == 0) continue;
      const ulonglong2 *e = reinterpret_cast<const ulonglong2 *>(st) + (size_t)pr * Kp * 2;
      const ulonglong2 S01 = srcp[2 * pr], S23 = srcp[2 * pr + 1];
      ulonglong2 c0xy = e[0], c0zw = e[1], c1xy = e[2], c1zw = e[3];
      const f2 sx = S01.x, sy = S01.y, sz = S23.x;
      // a = A' s' ; q = a + tau : queries relative to q0_b.  Order fixed (index parity).
      const f2 ax = fma2(A0, sx, fma2(A1, sy, mul2(A2, sz)));
      const f2 ay = fma2(A3, sx, fma2(A4, sy, mul2(A5, sz)));
      const f2 az = fma2(A6, sx, fma2(A7, sy, mul2(A8, sz)));
      const f2 qx = add2(ax, T0), qy = add2(ay, T1), qz = add2(az, T2);
      float best0 = INFINITY, best1 = INFINITY;
      int bi0 = 0, bi1 = 0;
      SVN_EVAL2(c0xy, c0zw, 0) SVN_EVAL2(c1xy, c1zw, 1)
      if (n > 2) {
        // upper bounds of |q| (distance of each query from its initial-guess query q0_b)
        const f2 qq = fma2(qz, qz, fma2(qy, qy, mul2(qx, qx)));
        const float qn0 = fmaf(sqrt_ftz(lo32(qq)), 1.00001f, 1e-7f), qn1 = fmaf(sqrt_ftz(hi32(qq)), 1.00001f, 1e-7f);
        for (int k0 = 2; k0 < n; k0 += 2) {
          c0xy = e[2 * k0]; c0zw = e[2 * k0 + 1]; c1xy = e[2 * k0 + 2]; c1zw = e[2 * k0 + 3];
          // exact early exit (see k_gn): slots ascend in |c| per row (w = lower bound), |q - c| >= |c| - |q|
          const float t0 = lo32(c0zw.y) - qn0, t1 = hi32(c0zw.y) - qn1;
          const bool done = (t0 > 0.f) && (t0 * t0 * 0.99999f > best0) && (t1 > 0.f) && (t1 * t1 * 0.99999f > best1);
          if (__all_sync(0xffffffffu, done)) break;
          SVN_EVAL2(c0xy, c0zw, k0) SVN_EVAL2(c1xy, c1zw, k0 + 1)
        }
      }
      const ulonglong2 w0xy = e[2 * bi0], w0zw = e[2 * bi0 + 1], w1xy = e[2 * bi1], w1zw = e[2 * bi1 + 1];
      const float c0x = lo32(w0xy.x), c0y = lo32(w0xy.y), c0z = lo32(w0zw.x);
      const float c1x = hi32(w1xy.x), c1y = hi32(w1xy.y), c1z = hi32(w1zw.x);
      const f2 ex = sub2(qx, pk(c0x, c1x)), ey = sub2(qy, pk(c0y, c1y)), ez = sub2(qz, pk(c0z, c1z));
      const bool valid0 = best0 < Dm, valid1 = best1 < Dm;  // SVGDICP.cpp:332 (Q1)
      if (DBG) {
        if (active) {
          const int row0 = ((slice + i * n_slices) * TP + pr) * 2;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int row = row0 + h;
            if (row >= a.n_s) continue;
            const float wx_ = h ? c1x : c0x, wy_ = h ? c1y : c0y, wz_ = h ? c1z : c0z;
            const float4 *full_row = a.cand + (size_t)row * a.K;
            int slot = 0;
            for (int kk = 0; kk < a.K; kk++) {
              const float4 f = full_row[kk];
              if (f.x == wx_ && f.y == wy_ && f.z == wz_) { slot = kk; break; }
            }
            a.dbg_idx[(size_t)l * a.n_s + row] = a.cand_idx[(size_t)row * a.K + slot];
            a.dbg_mask[(size_t)l * a.n_s + row] = (h ? valid1 : valid0) ? 1 : 0;
          }
        }
      }
      // rho = (D / (D + 3 |e|))^2 (SVNICP.cpp:120-122); masked pairs: rho' = 0 and +1 on the translation block (Q2)
      const float en0 = sqrt_ftz(best0), en1 = sqrt_ftz(best1);
      const float wq0 = Dm * rcp_ftz(fmaf(3.0f, en0, Dm)), wq1 = Dm * rcp_ftz(fmaf(3.0f, en1, Dm));
      const float rho0 = wq0 * wq0, rho1 = wq1 * wq1;
      const bool ex1 = (((slice + i * n_slices) * TP + pr) * 2 + 1) < a.n_s;  // odd N_s: the last pair has one real row
      const f2 rp = pk(valid0 ? rho0 : 0.0f, valid1 ? rho1 : 0.0f);
      acc[0] = add2(acc[0], pk(valid0 ? rho0 : 1.0f, ex1 ? (valid1 ? rho1 : 1.0f) : 0.0f));
      const f2 gx = mul2(rp, sx), gy = mul2(rp, sy), gz = mul2(rp, sz);
      acc[1] = add2(acc[1], gx); acc[2] = add2(acc[2], gy); acc[3] = add2(acc[3], gz);
      acc[4] = fma2(gx, sx, acc[4]); acc[5] = fma2(gx, sy, acc[5]); acc[6] = fma2(gx, sz, acc[6]);
      acc[7] = fma2(gy, sy, acc[7]); acc[8] = fma2(gy, sz, acc[8]); acc[9] = fma2(gz, sz, acc[9]);
      const f2 fx = mul2(rp, ex), fy = mul2(rp, ey), fz = mul2(rp, ez);  // multiplication (not select): NaN must propagate
      acc[10] = add2(acc[10], fx); acc[11] = add2(acc[11], fy); acc[12] = add2(acc[12], fz);
      const f2 wx = add2(ax, sx), wy = add2(ay, sy), wz = add2(az, sz);  // R~ s in the world-oriented frame
      // C = sum w x f, kept as C+ - C- (no packed negate): [13..15] += (wy fz, wz fx, wx fy), [16..18] += (wz fy, wx fz, wy fx)
      acc[13] = fma2(wy, fz, acc[13]); acc[14] = fma2(wz, fx, acc[14]); acc[15] = fma2(wx, fy, acc[15]);
      acc[16] = fma2(wz, fy, acc[16]); acc[17] = fma2(wx, fz, acc[17]); acc[18] = fma2(wy, fx, acc[18]);
      pairs_in_acc++;
    }
    if (pairs_in_acc >= GP_FLUSH_PAIRS) {
#pragma unroll
      for (int j = 0; j < 13; j++) { acc2[j] += lo32(acc[j]) + hi32(acc[j]); acc[j] = 0ull; }
#pragma unroll
      for (int j = 0; j < 3; j++) {
        acc2[13 + j] += (lo32(acc[13 + j]) + hi32(acc[13 + j])) - (lo32(acc[16 + j]) + hi32(acc[16 + j]));
        acc[13 + j] = 0ull; acc[16 + j] = 0ull;
      }
      pairs_in_acc = 0;
      if (++flushes2 >= GP_FLUSH2 && active) {
#pragma unroll
        for (int j = 0; j < NACC; j++) { out[j] = (wrote ? out[j] : 0.0) + (double)acc2[j]; acc2[j] = 0.f; }
        wrote = true;
        flushes2 = 0;
      }
    }
    __syncwarp();
    if (lane == 0) mb_arrive(empty + s);
  }
#undef SVN_EVAL2
  if (active) {
#pragma unroll
    for (int j = 0; j < 13; j++) out[j] = (wrote ? out[j] : 0.0) + ((double)acc2[j] + ((double)lo32(acc[j]) + (double)hi32(acc[j])));
#pragma unroll
    for (int j = 0; j < 3; j++)
      out[13 + j] = (wrote ? out[13 + j] : 0.0) + ((double)acc2[13 + j] + (((double)lo32(acc[13 + j]) + (double)hi32(acc[13 + j])) -
                                                                        ((double)lo32(acc[16 + j]) + (double)hi32(acc[16 + j]))));
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline int cdivp(long long a, long long b) { return (int)((a + b - 1) / b); }

void init_pair_kernels() {
  cudaFuncSetAttribute(k_gn_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_gn_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

int launch_spair(const float4 *sp, float4 *spair, int n_pad, cudaStream_t st) {
  k_spair<<<cdivp(n_pad / 2, 256), 256, 0, st>>>(sp, spair, n_pad / 2);
  return 1;
}

int launch_filter_pair(const IterArgs &a, cudaStream_t st) {
  const int nch = (a.K + 31) / 32;
  int grid = cdivp((long long)(a.n_pad / 2) * 32, 256);
  const int max_grid = a.sm_count * 8;
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  switch (nch) {
    case 1: k_filter_pair<1><<<grid, 256, 0, st>>>(a); break;
    case 2: k_filter_pair<2><<<grid, 256, 0, st>>>(a); break;
    case 3: k_filter_pair<3><<<grid, 256, 0, st>>>(a); break;
    case 4: k_filter_pair<4><<<grid, 256, 0, st>>>(a); break;
    case 5: k_filter_pair<5><<<grid, 256, 0, st>>>(a); break;
    case 6: k_filter_pair<6><<<grid, 256, 0, st>>>(a); break;
    case 7: k_filter_pair<7><<<grid, 256, 0, st>>>(a); break;
    default: k_filter_pair<8><<<grid, 256, 0, st>>>(a); break;
  }
  return 1;
}

int launch_gn_pair(const IterArgs &a, cudaStream_t st) {
  dim3 grid(a.n_slices, a.n_pgroups);
  if (a.dbg_idx) k_gn_pair<true><<<grid, GP_THREADS, a.gn_smem, st>>>(a);
  else k_gn_pair<false><<<grid, GP_THREADS, a.gn_smem, st>>>(a);
  return 1;
}

}  // namespace svn

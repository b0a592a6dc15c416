// preprocess.cu -- the scan pre-processing that runs immediately before add_cloud in the reference's node, on the device
// (SURVEY.md 8(f) row 3): OdometryPipeline::crop_pointcloud (svn-icp/src/core/OdometryPipeline.cpp:692-704) and
// OdometryPipeline::downsample_uniform (:684-690 = pcl::UniformSampling with leaf size = the "radius", applied twice:
// 0.5 * voxel_size for the cloud that goes into the map, then 1.5 * voxel_size for the registration source, :559-560).
//
// Semantics kept:
//   crop       keep p iff min_range^2 < x*x + y*y + z*z < max_range^2 (norm in float, compared in double), input order kept;
//              also reports max over ALL points of that float norm -- the reference stores this SQUARED value in
//              scan_max_range_ (:699) and later uses it as a range in metres (:577): the caller keeps the running maximum.
//   uniform    PCL 1.12 filters/impl/uniform_sampling.hpp (PCL is absent from this image: restated from its published source):
//              ijk = floor(p * (1.0f / float(leaf))) per axis; per leaf the point with the smallest
//              (x - i)^2 + (y - j)^2 + (z - k)^2 wins -- PCL measures the distance to the leaf's INTEGER INDEX, not to its
//              centre -- the first point on ties (strict '<' in cloud order).  Output order is PCL's unordered_map iteration
//              order, i.e. not part of the contract; here it is hash-slot order.
// All of it is streaming integer/byte work: one pass to mark or vote (atomicMin on (distance bits, point index) packs the
// sequential rule "strictly closer, else the earlier point" into one 64-bit key), a block scan, one pass to emit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/svnicp_b200.h"

namespace {

constexpr int PB = 1024;  // scan block
constexpr unsigned long long EMPTY = ~0ull;

__device__ int pre_block_excl_scan(int v, int *total) {
  __shared__ int s_w[32];
  __shared__ int s_tot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int x = s_w[lane], xi = x;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += u;
    }
    s_w[lane] = xi - x;
    if (lane == 31) s_tot = xi;
  }
  __syncthreads();
  const int r = s_w[warp] + incl - v;
  if (total) *total = s_tot;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(PB) k_pre_scan_sums(int *sums, int n_blocks, long long *total_out) {
  int carry = 0;
  for (int base = 0; base < n_blocks; base += PB) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? sums[i] : 0;
    int tot;
    const int ex = pre_block_excl_scan(v, &tot);
    if (i < n_blocks) sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__device__ __forceinline__ bool crop_keep(const float *p, double lo2, double hi2, float *norm_out) {
  // :697, evaluated as the host compiler does for a baseline x86-64 build: three products, two adds, no fused multiply-add
  const float norm = __fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2]));
  *norm_out = norm;
  return (double)norm < hi2 && (double)norm > lo2;               // :700
}

__global__ void __launch_bounds__(PB) k_crop_count(const float *xyz, int n, double lo2, double hi2, int *sums, unsigned *max_bits) {
  __shared__ int s_w[32];
  const int i = blockIdx.x * PB + threadIdx.x;
  int c = 0;
  float norm = 0.f;
  if (i < n) c = crop_keep(xyz + 3 * (size_t)i, lo2, hi2, &norm) ? 1 : 0;
  // max of the (non-negative) float norms: their bit patterns order like the values; NaN (bits above +inf) propagates as in a max
  unsigned nb = (i < n) ? __float_as_uint(norm) : 0u;
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, o));
  }
  if ((threadIdx.x & 31) == 0) { s_w[threadIdx.x >> 5] = c; atomicMax(max_bits, nb); }
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = s_w[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(PB) k_crop_emit(const float *xyz, int n, double lo2, double hi2, const int *offs, float *out) {
  const int i = blockIdx.x * PB + threadIdx.x;
  float norm;
  const int c = (i < n && crop_keep(xyz + 3 * (size_t)i, lo2, hi2, &norm)) ? 1 : 0;
  const int off = offs[blockIdx.x] + pre_block_excl_scan(c, nullptr);
  if (c)
    for (int k = 0; k < 3; k++) out[3 * (size_t)off + k] = xyz[3 * (size_t)i + k];
}

__device__ __forceinline__ unsigned long long leaf_key(int i, int j, int k) {
  const unsigned long long B = 1ull << 20;
  return ((unsigned long long)(i + (long long)B) & 0x1FFFFFull) | (((unsigned long long)(j + (long long)B) & 0x1FFFFFull) << 21) |
         (((unsigned long long)(k + (long long)B) & 0x1FFFFFull) << 42);
}
__device__ __forceinline__ unsigned leaf_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}

// first pass of UniformSampling::applyFilter: every point votes for its leaf with (distance bits, index)
__global__ void k_us_vote(const float *xyz, int n, float inv_leaf, unsigned long long *keys, unsigned long long *best, unsigned mask) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float x = xyz[3 * (size_t)p], y = xyz[3 * (size_t)p + 1], z = xyz[3 * (size_t)p + 2];
  const int i = (int)floorf(x * inv_leaf), j = (int)floorf(y * inv_leaf), k = (int)floorf(z * inv_leaf);
  const float dx = x - (float)i, dy = y - (float)j, dz = z - (float)k;
  // distance to the leaf INDEX (PCL's rule), unfused left-to-right like the restatement in oracle/preprocess_oracle.c;
  // NaN points vote last
  const float diff = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const unsigned long long vote = ((unsigned long long)__float_as_uint(diff) << 32) | (unsigned)p;
  const unsigned long long key = leaf_key(i, j, k);
  unsigned slot = leaf_hash(key) & mask;
  for (;;) {
    const unsigned long long cur = keys[slot];
    if (cur == key) break;
    if (cur == EMPTY) {
      const unsigned long long old = atomicCAS(&keys[slot], EMPTY, key);
      if (old == EMPTY || old == key) break;
    }
    slot = (slot + 1) & mask;
  }
  atomicMin(&best[slot], vote);  // smaller distance wins; equal distances: the earlier point (strict '<' in cloud order)
}

__global__ void __launch_bounds__(PB) k_us_count(const unsigned long long *keys, int slots, int *sums) {
  __shared__ int s_w[32];
  const int s = blockIdx.x * PB + threadIdx.x;
  int c = (s < slots && keys[s] != EMPTY) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = s_w[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(PB) k_us_emit(const float *xyz, const unsigned long long *keys, const unsigned long long *best, int slots,
                                                const int *offs, float *out) {
  const int s = blockIdx.x * PB + threadIdx.x;
  const int c = (s < slots && keys[s] != EMPTY) ? 1 : 0;
  const int off = offs[blockIdx.x] + pre_block_excl_scan(c, nullptr);
  if (c) {
    const unsigned p = (unsigned)(best[s] & 0xFFFFFFFFull);
    for (int k = 0; k < 3; k++) out[3 * (size_t)off + k] = xyz[3 * (size_t)p + k];
  }
}

__global__ void k_f32_to_f64(const float *in, double *out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

}  // namespace

struct svnicp_pre_t {
  int device = 0;
  size_t cap = 0;  // points
  float *in = nullptr, *buf[2] = {nullptr, nullptr};
  double *f64 = nullptr;
  unsigned long long *keys = nullptr, *best = nullptr;
  size_t slots = 0;
  int *sums = nullptr;
  long long *d_total = nullptr;
  unsigned *d_max = nullptr;
  long long *h_total = nullptr;  // pinned: [0] count, [1] max bits
  int last = 1;                  // buffer holding the last output
  cudaStream_t stream = nullptr;
  std::string err;
};

static thread_local std::string g_pre_create_error;

static int pfail(svnicp_pre p, int code, const char *what, cudaError_t e = cudaSuccess) {
  char b[256];
  snprintf(b, sizeof(b), "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
  if (p) p->err = b;
  else g_pre_create_error = b;
  return code;
}
#define PCU(call)                                                                                                  \
  do {                                                                                                             \
    cudaError_t e__ = (call);                                                                                      \
    if (e__ != cudaSuccess) return pfail(p, e__ == cudaErrorMemoryAllocation ? SVNICP_ERR_OOM : SVNICP_ERR_CUDA, #call, e__); \
  } while (0)

// input staging: host clouds are copied in; a device cloud is used where it lies.  Output goes to the buffer that is not the input.
static int stage_input(svnicp_pre p, const float *xyz, int64_t n, int on_device, const float **src, int *out_buf) {
  if ((size_t)n > p->cap) return pfail(p, SVNICP_ERR_INVALID, "cloud larger than max_points given to svnicp_pre_create");
  if (!on_device) {
    PCU(cudaMemcpyAsync(p->in, xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    *src = p->in;
    *out_buf = p->last ^ 1;
  } else {
    *src = xyz;
    *out_buf = (xyz == p->buf[0]) ? 1 : (xyz == p->buf[1]) ? 0 : (p->last ^ 1);
  }
  return SVNICP_OK;
}

extern "C" {

int svnicp_pre_create(svnicp_pre *out, int64_t max_points, int device) {
  if (!out) return SVNICP_ERR_INVALID;
  *out = nullptr;
  if (max_points < 1 || max_points > (1ll << 30)) return pfail(nullptr, SVNICP_ERR_INVALID, "svnicp_pre_create: bad max_points");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return pfail(nullptr, SVNICP_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  if (device < 0) cudaGetDevice(&device);
  if (device >= ndev) return pfail(nullptr, SVNICP_ERR_NO_DEVICE, "device out of range");
  svnicp_pre p = new svnicp_pre_t();
  p->device = device;
  p->cap = (size_t)max_points;
  size_t slots = 1024;
  while (slots < 2 * p->cap) slots <<= 1;
  p->slots = slots;
  auto body = [&]() -> int {
    PCU(cudaSetDevice(device));
    PCU(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PCU(cudaMalloc((void **)&p->in, p->cap * 3 * sizeof(float)));
    for (int i = 0; i < 2; i++) PCU(cudaMalloc((void **)&p->buf[i], p->cap * 3 * sizeof(float)));
    PCU(cudaMalloc((void **)&p->f64, p->cap * 3 * sizeof(double)));
    PCU(cudaMalloc((void **)&p->keys, slots * sizeof(unsigned long long)));
    PCU(cudaMalloc((void **)&p->best, slots * sizeof(unsigned long long)));
    PCU(cudaMalloc((void **)&p->sums, ((slots + PB - 1) / PB + 2) * sizeof(int)));
    PCU(cudaMalloc((void **)&p->d_total, sizeof(long long)));
    PCU(cudaMalloc((void **)&p->d_max, sizeof(unsigned)));
    PCU(cudaMallocHost((void **)&p->h_total, 2 * sizeof(long long)));
    return SVNICP_OK;
  };
  const int rc = body();
  if (rc != SVNICP_OK) {
    g_pre_create_error = p->err;
    svnicp_pre_destroy(p);
    return rc;
  }
  *out = p;
  return SVNICP_OK;
}

void svnicp_pre_destroy(svnicp_pre p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  cudaFree(p->in); cudaFree(p->buf[0]); cudaFree(p->buf[1]); cudaFree(p->f64); cudaFree(p->keys); cudaFree(p->best);
  cudaFree(p->sums); cudaFree(p->d_total); cudaFree(p->d_max);
  if (p->h_total) cudaFreeHost(p->h_total);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

const char *svnicp_pre_last_error(svnicp_pre p) { return p ? p->err.c_str() : g_pre_create_error.c_str(); }

int svnicp_pre_crop(svnicp_pre p, const float *xyz, int64_t n, int on_device, double min_range, double max_range, const float **dev_out,
                    int64_t *n_out, double *max_sq_norm) {
  if (!p || !n_out || (n > 0 && !xyz) || n < 0) return SVNICP_ERR_INVALID;
  PCU(cudaSetDevice(p->device));
  const float *src = nullptr;
  int ob = 0;
  int rc = stage_input(p, xyz, n, on_device, &src, &ob);
  if (rc) return rc;
  const int nb = (int)((n + PB - 1) / PB);
  PCU(cudaMemsetAsync(p->d_max, 0, sizeof(unsigned), p->stream));
  PCU(cudaMemsetAsync(p->d_total, 0, sizeof(long long), p->stream));
  const double lo2 = min_range * min_range, hi2 = max_range * max_range;
  if (n > 0) {
    k_crop_count<<<nb, PB, 0, p->stream>>>(src, (int)n, lo2, hi2, p->sums, p->d_max);
    k_pre_scan_sums<<<1, PB, 0, p->stream>>>(p->sums, nb, p->d_total);
    k_crop_emit<<<nb, PB, 0, p->stream>>>(src, (int)n, lo2, hi2, p->sums, p->buf[ob]);
    PCU(cudaGetLastError());
  }
  PCU(cudaMemcpyAsync(&p->h_total[0], p->d_total, sizeof(long long), cudaMemcpyDeviceToHost, p->stream));
  PCU(cudaMemcpyAsync(&p->h_total[1], p->d_max, sizeof(unsigned), cudaMemcpyDeviceToHost, p->stream));
  PCU(cudaStreamSynchronize(p->stream));
  p->last = ob;
  *n_out = p->h_total[0];
  if (dev_out) *dev_out = p->buf[ob];
  if (max_sq_norm) {
    const unsigned bits = (unsigned)(p->h_total[1] & 0xFFFFFFFFll);
    float f;
    memcpy(&f, &bits, sizeof(f));
    *max_sq_norm = (double)f;
  }
  return SVNICP_OK;
}

int svnicp_pre_downsample_uniform(svnicp_pre p, const float *xyz, int64_t n, int on_device, double leaf, const float **dev_out, int64_t *n_out) {
  if (!p || !n_out || (n > 0 && !xyz) || n < 0) return SVNICP_ERR_INVALID;
  if (!(leaf > 0)) return pfail(p, SVNICP_ERR_INVALID, "svnicp_pre_downsample_uniform: leaf size must be positive");
  PCU(cudaSetDevice(p->device));
  const float *src = nullptr;
  int ob = 0;
  int rc = stage_input(p, xyz, n, on_device, &src, &ob);
  if (rc) return rc;
  PCU(cudaMemsetAsync(p->keys, 0xFF, p->slots * sizeof(unsigned long long), p->stream));
  PCU(cudaMemsetAsync(p->best, 0xFF, p->slots * sizeof(unsigned long long), p->stream));
  PCU(cudaMemsetAsync(p->d_total, 0, sizeof(long long), p->stream));
  if (n > 0) {
    const float inv_leaf = 1.0f / (float)leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
    const int nb = (int)((p->slots + PB - 1) / PB);
    k_us_vote<<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(src, (int)n, inv_leaf, p->keys, p->best, (unsigned)(p->slots - 1));
    k_us_count<<<nb, PB, 0, p->stream>>>(p->keys, (int)p->slots, p->sums);
    k_pre_scan_sums<<<1, PB, 0, p->stream>>>(p->sums, nb, p->d_total);
    k_us_emit<<<nb, PB, 0, p->stream>>>(src, p->keys, p->best, (int)p->slots, p->sums, p->buf[ob]);
    PCU(cudaGetLastError());
  }
  PCU(cudaMemcpyAsync(&p->h_total[0], p->d_total, sizeof(long long), cudaMemcpyDeviceToHost, p->stream));
  PCU(cudaStreamSynchronize(p->stream));
  p->last = ob;
  *n_out = p->h_total[0];
  if (dev_out) *dev_out = p->buf[ob];
  return SVNICP_OK;
}

int svnicp_pre_to_f64(svnicp_pre p, const float *dev_xyz, int64_t n, const double **dev_out) {
  if (!p || !dev_out || n < 0 || (size_t)n > p->cap) return SVNICP_ERR_INVALID;
  PCU(cudaSetDevice(p->device));
  if (n > 0) {
    k_f32_to_f64<<<(unsigned)(((size_t)n * 3 + 255) / 256), 256, 0, p->stream>>>(dev_xyz, p->f64, (size_t)n * 3);
    PCU(cudaGetLastError());
  }
  PCU(cudaStreamSynchronize(p->stream));
  *dev_out = p->f64;
  return SVNICP_OK;
}

int svnicp_pre_download(svnicp_pre p, const float *dev_xyz, int64_t n, float *out) {
  if (!p || n < 0 || (n > 0 && (!dev_xyz || !out))) return SVNICP_ERR_INVALID;
  PCU(cudaSetDevice(p->device));
  if (n > 0) PCU(cudaMemcpy(out, dev_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  return SVNICP_OK;
}

}  // extern "C"

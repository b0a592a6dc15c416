// voxel_map.cu -- device-resident local map: the reference's svnicp::VoxelHashMap (svn-icp/include/core/VoxelHashMap.h:28-72,
// src/core/VoxelHashMap.cpp:22-101) kept in HBM, so that the per-scan target cloud handed to add_cloud never leaves the GPU
// (the reference rebuilds it on the host and re-uploads it every scan, OdometryPipeline.cpp:577-581).  SURVEY.md 8(f) row 1.
//
// Semantics kept (VoxelHashMap.cpp):
//   AddPointCloud :22-43   p_map = float(T * p) (pcl::transformPointCloud with a double matrix); voxel index =
//                          (p_map / float(voxel_size)).cast<int>() -- truncation toward zero; a voxel keeps its FIRST
//                          max_pointscount points in arrival order; then RemoveFarPointCloud(translation)
//   GetMap()      :45-51   all points;  GetMap(pose, r) :53-63: voxels whose FIRST point is closer than r to the position
//   RemoveFar     :93-101  voxels whose first point is farther than max_range are dropped
// "Arrival order" is made deterministic on the GPU by ranking the points of one AddPointCloud batch by their index in the
// cloud (what the reference's sequential loop does): pass 1 links every point into its voxel's list and elects the lowest
// index as the voxel's walker; pass 2 lets the walker keep the (cap - count) lowest indices, ascending.
// The table is open addressing (linear probing) over packed 3x21-bit voxel indices; RemoveFar rebuilds into the twin table
// (no tombstones).  Point output order of GetMap is slot order: arbitrary like the reference's robin_map iteration.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/svnicp_b200.h"

namespace {

constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int MAX_CAP = 32;
constexpr int SCAN_BLOCK = 1024;

struct MapTable {
  unsigned long long *keys;  // [slots]
  int *count;                // [slots]
  float *pts;                // [slots][cap][3]
};

struct MapDev {
  MapTable tab;
  int *head, *walker;        // [slots] per-batch list head / elected walker (restored to -1 / INT_MAX after every batch)
  long long *counters;       // [0] voxels, [1] points, [2] far voxels (scratch), [3] dropped (table full)
  int slots_mask, cap;
  float inv_guard;           // unused
  float voxel_size_f;
};

__device__ __forceinline__ unsigned long long pack_key(int vx, int vy, int vz) {
  const unsigned long long B = 1ull << 20;
  return ((unsigned long long)(vx + (long long)B) & 0x1FFFFFull) | (((unsigned long long)(vy + (long long)B) & 0x1FFFFFull) << 21) |
         (((unsigned long long)(vz + (long long)B) & 0x1FFFFFull) << 42);
}
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}

// find-or-create; returns the slot or -1 when the table is full
__device__ int find_or_create(const MapDev &m, unsigned long long key, bool *created) {
  unsigned slot = hash_key(key) & (unsigned)m.slots_mask;
  *created = false;
  for (int probe = 0; probe <= m.slots_mask; probe++) {
    const unsigned long long k = m.tab.keys[slot];
    if (k == key) return (int)slot;
    if (k == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&m.tab.keys[slot], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { *created = true; return (int)slot; }
      if (old == key) return (int)slot;
    }
    slot = (slot + 1) & (unsigned)m.slots_mask;
  }
  return -1;
}

// pass 1: transform (VoxelHashMap.cpp:23-26), voxel index (:30), slot lookup, link into the voxel's batch list
template <typename T>
__global__ void k_map_insert_link(MapDev m, const T *xyz, int n, const double *Rt /*R[9], t[3]*/, float *wxyz, int *pt_slot, int *next) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {(double)xyz[3 * (size_t)i], (double)xyz[3 * (size_t)i + 1], (double)xyz[3 * (size_t)i + 2]};
  float q[3];
#pragma unroll
  for (int r = 0; r < 3; r++) q[r] = (float)(Rt[3 * r] * p[0] + Rt[3 * r + 1] * p[1] + Rt[3 * r + 2] * p[2] + Rt[9 + r]);
#pragma unroll
  for (int r = 0; r < 3; r++) wxyz[3 * (size_t)i + r] = q[r];
  const int vx = (int)(q[0] / m.voxel_size_f), vy = (int)(q[1] / m.voxel_size_f), vz = (int)(q[2] / m.voxel_size_f);
  bool created;
  const int slot = find_or_create(m, pack_key(vx, vy, vz), &created);
  pt_slot[i] = slot;
  if (slot < 0) { atomicAdd((unsigned long long *)&m.counters[3], 1ull); return; }
  if (created) atomicAdd((unsigned long long *)&m.counters[0], 1ull);
  if (m.tab.count[slot] >= m.cap) { pt_slot[i] = -2; return; }  // voxel already full before this cloud (:33)
  next[i] = atomicExch(&m.head[slot], i);
  atomicMin(&m.walker[slot], i);
}

// pass 2: the lowest-index point of each touched voxel appends the (cap - count) lowest indices in ascending order
__global__ void k_map_insert_place(MapDev m, const float *wxyz, const int *pt_slot, const int *next, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = pt_slot[i];
  if (slot < 0 || m.walker[slot] != i) return;
  const int have = m.tab.count[slot];
  const int need = m.cap - have;
  int best[MAX_CAP];
  int nb = 0;
  for (int j = m.head[slot]; j >= 0; j = next[j]) {
    if (nb < need) {
      int k = nb++;
      while (k > 0 && best[k - 1] > j) { best[k] = best[k - 1]; k--; }
      best[k] = j;
    } else if (j < best[nb - 1]) {
      int k = nb - 1;
      while (k > 0 && best[k - 1] > j) { best[k] = best[k - 1]; k--; }
      best[k] = j;
    }
  }
  float *dst = m.tab.pts + ((size_t)slot * m.cap + have) * 3;
  for (int k = 0; k < nb; k++)
    for (int c = 0; c < 3; c++) dst[3 * k + c] = wxyz[3 * (size_t)best[k] + c];
  m.tab.count[slot] = have + nb;
  m.head[slot] = -1;
  m.walker[slot] = INT_MAX;
  atomicAdd((unsigned long long *)&m.counters[1], (unsigned long long)nb);
}

__device__ __forceinline__ bool front_within(const MapTable &t, int cap, int slot, const double *pos, double r2, bool strict_less) {
  const float *f = t.pts + (size_t)slot * cap * 3;
  const double dx = (double)f[0] - pos[0], dy = (double)f[1] - pos[1], dz = (double)f[2] - pos[2];
  const double d2 = dx * dx + dy * dy + dz * dz;
  return strict_less ? (d2 < r2) : !(d2 > r2);
}

// RemoveFarPointCloud (:93-101): count the voxels to drop
__global__ void k_map_count_far(MapDev m, const double *pos, double r2) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > (size_t)m.slots_mask) return;
  if (m.tab.keys[s] == EMPTY_KEY || m.tab.count[s] == 0) return;
  if (!front_within(m.tab, m.cap, s, pos, r2, false)) atomicAdd((unsigned long long *)&m.counters[2], 1ull);
}

// ... and rebuild the survivors into the twin table (one warp per old slot: coalesced copy of the voxel's points)
__global__ void k_map_rebuild(MapDev old, MapDev neu, const double *pos, double r2) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // 64-bit: slots * 32 exceeds 2^32 beyond 2^27 slots
  const int lane = threadIdx.x & 31;
  if (w > (size_t)old.slots_mask) return;
  const unsigned long long key = old.tab.keys[w];
  const int cnt = old.tab.count[w];
  if (key == EMPTY_KEY || cnt == 0) return;
  if (!front_within(old.tab, old.cap, w, pos, r2, false)) return;
  int slot = 0;
  if (lane == 0) {
    bool created;
    slot = find_or_create(neu, key, &created);
    if (slot >= 0) {
      neu.tab.count[slot] = cnt;
      atomicAdd((unsigned long long *)&neu.counters[0], 1ull);
      atomicAdd((unsigned long long *)&neu.counters[1], (unsigned long long)cnt);
    }
  }
  slot = __shfl_sync(0xffffffffu, slot, 0);
  if (slot < 0) return;
  const float *src = old.tab.pts + (size_t)w * old.cap * 3;
  float *dst = neu.tab.pts + (size_t)slot * neu.cap * 3;
  for (int e = lane; e < cnt * 3; e += 32) dst[e] = src[e];
}

// GetMap (:45-63): three-step deterministic compaction in slot order; output fp64 xyz ready for add_cloud
__device__ __forceinline__ int slot_emit_count(const MapDev &m, int s, const double *pos, double r2, int all) {
  if (s > m.slots_mask || m.tab.keys[s] == EMPTY_KEY) return 0;
  const int c = m.tab.count[s];
  if (c == 0) return 0;
  return (all || front_within(m.tab, m.cap, s, pos, r2, true)) ? c : 0;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_getmap_count(MapDev m, const double *pos, double r2, int all, int *block_sums) {
  __shared__ int s_w[32];
  const int s = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  int c = slot_emit_count(m, s, pos, r2, all);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = s_w[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
  }
}

__device__ int block_exclusive_scan(int v, int *total) {  // SCAN_BLOCK threads
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

__global__ void __launch_bounds__(SCAN_BLOCK) k_getmap_scan(int *block_sums, int n_blocks, long long *total_out) {
  // one block; n_blocks <= SCAN_BLOCK * 8
  int carry = 0;
  for (int base = 0; base < n_blocks; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_sums[i] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, &tot);
    if (i < n_blocks) block_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_getmap_emit(MapDev m, const double *pos, double r2, int all, const int *block_offs, double *out) {
  const int s = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int c = slot_emit_count(m, s, pos, r2, all);
  const int off = block_offs[blockIdx.x] + block_exclusive_scan(c, nullptr);
  if (c == 0) return;
  const float *src = m.tab.pts + (size_t)s * m.cap * 3;
  double *dst = out + (size_t)off * 3;
  for (int e = 0; e < c * 3; e++) dst[e] = (double)src[e];
}

__global__ void k_fill_i32(int *p, int v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

struct svnicp_map_t {
  int device = 0;
  double voxel_size = 1.0, max_range = 80.0;
  int cap = 20;
  size_t slots = 0;
  MapTable tab[2] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  int cur = 0;
  int *head = nullptr, *walker = nullptr;
  long long *counters = nullptr;  // device [4]
  long long *h_counters = nullptr;  // pinned [4]
  // batch scratch
  void *batch_in = nullptr;
  float *wxyz = nullptr;
  int *pt_slot = nullptr, *next = nullptr;
  size_t batch_cap = 0;
  double *d_pose = nullptr;  // [12] R,t  + [3] position at offset 12
  // GetMap output
  double *out = nullptr;
  size_t out_cap = 0;
  int *block_sums = nullptr;
  long long last_n = 0;
  cudaStream_t stream = nullptr;
  std::string err;
};

static thread_local std::string g_map_create_error;

static int mfail(svnicp_map m, int code, const char *what, cudaError_t e = cudaSuccess) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
  if (m) m->err = buf;
  else g_map_create_error = buf;
  return code;
}
#define MCU(call)                                                                                                  \
  do {                                                                                                             \
    cudaError_t e__ = (call);                                                                                      \
    if (e__ != cudaSuccess) return mfail(m, e__ == cudaErrorMemoryAllocation ? SVNICP_ERR_OOM : SVNICP_ERR_CUDA, #call, e__); \
  } while (0)

static MapDev dev_view(svnicp_map m, int which) {
  MapDev d;
  d.tab = m->tab[which];
  d.head = m->head; d.walker = m->walker; d.counters = m->counters;
  d.slots_mask = (int)(m->slots - 1); d.cap = m->cap; d.inv_guard = 0.f;
  d.voxel_size_f = (float)m->voxel_size;  // Eigen promotes the double scalar to the vector's float (VoxelHashMap.cpp:30)
  return d;
}

static int clear_table(svnicp_map m, int which) {
  MCU(cudaMemsetAsync(m->tab[which].keys, 0xFF, m->slots * sizeof(unsigned long long), m->stream));
  MCU(cudaMemsetAsync(m->tab[which].count, 0, m->slots * sizeof(int), m->stream));
  return SVNICP_OK;
}

extern "C" {

int svnicp_map_create(svnicp_map *out, double voxel_size, double max_range, int max_pointscount, int64_t capacity_voxels, int device) {
  if (!out) return SVNICP_ERR_INVALID;
  *out = nullptr;
  if (!(voxel_size > 0) || max_pointscount < 1 || max_pointscount > MAX_CAP || capacity_voxels < 1)
    return mfail(nullptr, SVNICP_ERR_INVALID, "svnicp_map_create: need voxel_size > 0, 1 <= max_pointscount <= 32, capacity_voxels >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return mfail(nullptr, SVNICP_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  if (device < 0) cudaGetDevice(&device);
  if (device >= ndev) return mfail(nullptr, SVNICP_ERR_NO_DEVICE, "device out of range");
  svnicp_map m = new svnicp_map_t();
  m->device = device; m->voxel_size = voxel_size; m->max_range = max_range; m->cap = max_pointscount;
  size_t slots = 1024;
  while (slots < (size_t)capacity_voxels * 2) slots <<= 1;
  if (slots > (1ull << 30)) { delete m; return mfail(nullptr, SVNICP_ERR_INVALID, "capacity_voxels too large"); }
  m->slots = slots;
  auto body = [&]() -> int {
    MCU(cudaSetDevice(device));
    MCU(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    for (int w = 0; w < 2; w++) {
      MCU(cudaMalloc((void **)&m->tab[w].keys, slots * sizeof(unsigned long long)));
      MCU(cudaMalloc((void **)&m->tab[w].count, slots * sizeof(int)));
      MCU(cudaMalloc((void **)&m->tab[w].pts, slots * (size_t)m->cap * 3 * sizeof(float)));
    }
    MCU(cudaMalloc((void **)&m->head, slots * sizeof(int)));
    MCU(cudaMalloc((void **)&m->walker, slots * sizeof(int)));
    MCU(cudaMalloc((void **)&m->counters, 4 * sizeof(long long)));
    MCU(cudaMalloc((void **)&m->d_pose, 16 * sizeof(double)));
    MCU(cudaMalloc((void **)&m->block_sums, ((slots + SCAN_BLOCK - 1) / SCAN_BLOCK + 1) * sizeof(int)));
    MCU(cudaMallocHost((void **)&m->h_counters, 4 * sizeof(long long)));
    memset(m->h_counters, 0, 4 * sizeof(long long));
    return SVNICP_OK;
  };
  int rc = body();
  if (rc == SVNICP_OK) rc = svnicp_map_clear(m);
  if (rc != SVNICP_OK) {
    g_map_create_error = m->err;
    svnicp_map_destroy(m);
    return rc;
  }
  *out = m;
  return SVNICP_OK;
}

void svnicp_map_destroy(svnicp_map m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (int w = 0; w < 2; w++) { cudaFree(m->tab[w].keys); cudaFree(m->tab[w].count); cudaFree(m->tab[w].pts); }
  cudaFree(m->head); cudaFree(m->walker); cudaFree(m->counters); cudaFree(m->d_pose); cudaFree(m->block_sums);
  cudaFree(m->batch_in); cudaFree(m->wxyz); cudaFree(m->pt_slot); cudaFree(m->next); cudaFree(m->out);
  if (m->h_counters) cudaFreeHost(m->h_counters);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

const char *svnicp_map_last_error(svnicp_map m) { return m ? m->err.c_str() : g_map_create_error.c_str(); }

int svnicp_map_clear(svnicp_map m) {  // VoxelHashMap::Clear, VoxelHashMap.h:54
  if (!m) return SVNICP_ERR_INVALID;
  MCU(cudaSetDevice(m->device));
  m->cur = 0;
  int rc = clear_table(m, 0);
  if (rc) return rc;
  const unsigned grid = (unsigned)((m->slots + 255) / 256);
  k_fill_i32<<<grid, 256, 0, m->stream>>>(m->head, -1, m->slots);
  k_fill_i32<<<grid, 256, 0, m->stream>>>(m->walker, INT_MAX, m->slots);
  MCU(cudaMemsetAsync(m->counters, 0, 4 * sizeof(long long), m->stream));
  MCU(cudaGetLastError());
  MCU(cudaStreamSynchronize(m->stream));
  memset(m->h_counters, 0, 4 * sizeof(long long));
  m->last_n = 0;
  return SVNICP_OK;
}

static int sync_counters(svnicp_map m) {
  MCU(cudaMemcpyAsync(m->h_counters, m->counters, 4 * sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  return SVNICP_OK;
}

int svnicp_map_add_cloud(svnicp_map m, const void *xyz, int64_t n, int dtype_f64, int on_device, const double R[9], const double t[3]) {
  if (!m || !R || !t || (n > 0 && !xyz)) return SVNICP_ERR_INVALID;
  if (n < 0 || n > (1ll << 30)) return mfail(m, SVNICP_ERR_INVALID, "svnicp_map_add_cloud: bad point count");
  MCU(cudaSetDevice(m->device));
  const size_t esz = dtype_f64 ? sizeof(double) : sizeof(float);
  if ((size_t)n > m->batch_cap) {
    cudaFree(m->batch_in); cudaFree(m->wxyz); cudaFree(m->pt_slot); cudaFree(m->next);
    m->batch_in = nullptr; m->wxyz = nullptr; m->pt_slot = nullptr; m->next = nullptr;
    const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
    m->batch_cap = 0;
    MCU(cudaMalloc(&m->batch_in, cap * 3 * sizeof(double)));
    MCU(cudaMalloc((void **)&m->wxyz, cap * 3 * sizeof(float)));
    MCU(cudaMalloc((void **)&m->pt_slot, cap * sizeof(int)));
    MCU(cudaMalloc((void **)&m->next, cap * sizeof(int)));
    m->batch_cap = cap;
  }
  double pose[15];
  memcpy(pose, R, 9 * sizeof(double));
  memcpy(pose + 9, t, 3 * sizeof(double));
  memcpy(pose + 12, t, 3 * sizeof(double));  // current_pos = new_pose.translation() (:26)
  MCU(cudaMemcpyAsync(m->d_pose, pose, sizeof(pose), cudaMemcpyHostToDevice, m->stream));
  const MapDev d = dev_view(m, m->cur);
  // points that found no free voxel slot in THIS call (reported below; not sticky: RemoveFar may free room for the next cloud)
  MCU(cudaMemsetAsync(m->counters + 3, 0, sizeof(long long), m->stream));
  if (n > 0) {
    const void *src = xyz;
    if (!on_device) {
      MCU(cudaMemcpyAsync(m->batch_in, xyz, (size_t)n * 3 * esz, cudaMemcpyHostToDevice, m->stream));
      src = m->batch_in;
    }
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype_f64) k_map_insert_link<double><<<grid, 256, 0, m->stream>>>(d, (const double *)src, (int)n, m->d_pose, m->wxyz, m->pt_slot, m->next);
    else k_map_insert_link<float><<<grid, 256, 0, m->stream>>>(d, (const float *)src, (int)n, m->d_pose, m->wxyz, m->pt_slot, m->next);
    k_map_insert_place<<<grid, 256, 0, m->stream>>>(d, m->wxyz, m->pt_slot, m->next, (int)n);
    MCU(cudaGetLastError());
  }
  // RemoveFarPointCloud(current_pos) (:42, :93-101)
  MCU(cudaMemsetAsync(m->counters + 2, 0, sizeof(long long), m->stream));
  const double r2 = m->max_range * m->max_range;
  k_map_count_far<<<(unsigned)((m->slots + 255) / 256), 256, 0, m->stream>>>(d, m->d_pose + 12, r2);
  MCU(cudaGetLastError());
  int rc = sync_counters(m);  // also makes the host copy safe to reuse (the caller may free xyz)
  if (rc) return rc;
  const long long dropped = m->h_counters[3];
  if (m->h_counters[2] > 0) {
    const int nxt = m->cur ^ 1;
    rc = clear_table(m, nxt);
    if (rc) return rc;
    MCU(cudaMemsetAsync(m->counters, 0, 2 * sizeof(long long), m->stream));
    const MapDev dn = dev_view(m, nxt);
    k_map_rebuild<<<(unsigned)((m->slots * 32 + 255) / 256), 256, 0, m->stream>>>(d, dn, m->d_pose + 12, r2);
    MCU(cudaGetLastError());
    m->cur = nxt;
    rc = sync_counters(m);
    if (rc) return rc;
  }
  // the far voxels are gone and the counters are current either way; the points that did not fit are simply not in the map
  if (dropped > 0) return mfail(m, SVNICP_ERR_OOM, "voxel table full: some points of this cloud were dropped; create the map with a larger capacity_voxels");
  return SVNICP_OK;
}

int svnicp_map_get(svnicp_map m, const double position[3], double max_range, const double **device_xyz, int64_t *n_out) {
  if (!m || !n_out) return SVNICP_ERR_INVALID;
  MCU(cudaSetDevice(m->device));
  const size_t need = (size_t)(m->h_counters[1] > 0 ? m->h_counters[1] : 1);
  if (need > m->out_cap) {
    cudaFree(m->out);
    m->out = nullptr;
    m->out_cap = 0;
    MCU(cudaMalloc((void **)&m->out, (need + need / 4 + 1024) * 3 * sizeof(double)));
    m->out_cap = need + need / 4 + 1024;
  }
  const int all = position ? 0 : 1;  // GetMap() (:45-51) vs GetMap(pose, max_range) (:53-63)
  if (position) MCU(cudaMemcpyAsync(m->d_pose + 12, position, 3 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  const MapDev d = dev_view(m, m->cur);
  const int n_blocks = (int)((m->slots + SCAN_BLOCK - 1) / SCAN_BLOCK);
  const double r2 = max_range * max_range;
  k_getmap_count<<<n_blocks, SCAN_BLOCK, 0, m->stream>>>(d, m->d_pose + 12, r2, all, m->block_sums);
  k_getmap_scan<<<1, SCAN_BLOCK, 0, m->stream>>>(m->block_sums, n_blocks, m->counters + 2);
  k_getmap_emit<<<n_blocks, SCAN_BLOCK, 0, m->stream>>>(d, m->d_pose + 12, r2, all, m->block_sums, m->out);
  MCU(cudaGetLastError());
  int rc = sync_counters(m);
  if (rc) return rc;
  m->last_n = m->h_counters[2];
  *n_out = m->last_n;
  if (device_xyz) *device_xyz = m->out;
  return SVNICP_OK;
}

int svnicp_map_download(svnicp_map m, double *out_xyz, int64_t n) {
  if (!m || (n > 0 && !out_xyz)) return SVNICP_ERR_INVALID;
  if (n > m->last_n) return mfail(m, SVNICP_ERR_INVALID, "svnicp_map_download: more points requested than the last svnicp_map_get produced");
  MCU(cudaSetDevice(m->device));
  if (n > 0) MCU(cudaMemcpy(out_xyz, m->out, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  return SVNICP_OK;
}

int svnicp_map_size(svnicp_map m, int64_t *voxels, int64_t *points) {  // VoxelHashMap::Size / Empty, VoxelHashMap.h:55-56
  if (!m) return SVNICP_ERR_INVALID;
  if (voxels) *voxels = m->h_counters[0];
  if (points) *points = m->h_counters[1];
  return SVNICP_OK;
}

}  // extern "C"

/*
 * voxelmap_oracle.c -- sequential CPU restatement (plain C) of the reference's local map,
 * svnicp::VoxelHashMap (svn-icp/include/core/VoxelHashMap.h:28-72, src/core/VoxelHashMap.cpp:22-101).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as svn_oracle.c).
 *
 * Parity status: the control flow of this file is pinned against the reference's own VoxelHashMap.cpp
 * compiled in this container over stand-in PCL / Eigen / tsl types (oracle/ref_shim, `ref_vmap_*` in
 * oracle/ref_driver_map.cpp; fixture tests/golden/vmap_sequence.npz).  The arithmetic inside the absent
 * third-party types is restated from their documented behaviour and is NOT pinned:
 *   - pcl::transformPointCloud with a double 4x4 (PCL 1.12 common/impl/transforms.hpp, Transformer<double>):
 *     out = float(m00*x + m01*y + m02*z + m03) evaluated in double;
 *   - Eigen `Vector3f / double`: the scalar is converted to float, the division is float (Eigen 3.4
 *     promote_scalar_arg), then `.cast<int>()` truncates toward zero.
 *   - RemoveFarPointCloud erases from a tsl::robin_map inside a range-for over it (VoxelHashMap.cpp:94-100);
 *     what the iterator does after the erase is the container's business.  Restated as intended: every voxel whose
 *     first point is farther than max_range goes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t v[3];
  int count;
  float *pts; /* [cap][3] */
  int used;   /* 0 empty, 1 live, 2 erased */
} vm_entry;

typedef struct {
  double voxel_size, max_range;
  int cap;
  vm_entry *tab;
  size_t slots, live, filled;
} vmap;

static size_t vm_hash(const int32_t v[3]) { /* any hash: order of iteration is not part of the contract */
  uint64_t h = (uint32_t)v[0] * 73856093u ^ (uint32_t)v[1] * 19349669u ^ (uint32_t)v[2] * 83492791u; /* VoxelHashMap.h:47-50 */
  h ^= h >> 15; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
  return (size_t)h;
}

vmap *oracle_vmap_create(double voxel_size, double max_range, int cap) { /* VoxelHashMap.h:40-43 */
  vmap *m = (vmap *)calloc(1, sizeof(vmap));
  m->voxel_size = voxel_size; m->max_range = max_range; m->cap = cap;
  m->slots = 1 << 12;
  m->tab = (vm_entry *)calloc(m->slots, sizeof(vm_entry));
  return m;
}

static void vm_free_tab(vm_entry *tab, size_t slots) {
  for (size_t i = 0; i < slots; i++) free(tab[i].pts);
  free(tab);
}

void oracle_vmap_destroy(vmap *m) {
  if (!m) return;
  vm_free_tab(m->tab, m->slots);
  free(m);
}

void oracle_vmap_clear(vmap *m) { /* VoxelHashMap.h:54 */
  vm_free_tab(m->tab, m->slots);
  m->slots = 1 << 12;
  m->tab = (vm_entry *)calloc(m->slots, sizeof(vm_entry));
  m->live = m->filled = 0;
}

static vm_entry *vm_find(vmap *m, const int32_t v[3], int create) {
  size_t s = vm_hash(v) & (m->slots - 1);
  vm_entry *grave = NULL;
  for (;;) {
    vm_entry *e = &m->tab[s];
    if (e->used == 0) {
      if (!create) return NULL;
      if (grave) e = grave; else m->filled++;
      e->used = 1; e->count = 0;
      memcpy(e->v, v, sizeof(e->v));
      if (!e->pts) e->pts = (float *)malloc(sizeof(float) * 3 * (size_t)m->cap);
      m->live++;
      return e;
    }
    if (e->used == 1 && e->v[0] == v[0] && e->v[1] == v[1] && e->v[2] == v[2]) return e;
    if (e->used == 2 && !grave) grave = e;
    s = (s + 1) & (m->slots - 1);
  }
}

static void vm_grow(vmap *m) {
  vm_entry *old = m->tab;
  const size_t os = m->slots;
  m->slots = os * 2;
  m->tab = (vm_entry *)calloc(m->slots, sizeof(vm_entry));
  m->live = m->filled = 0;
  for (size_t i = 0; i < os; i++) {
    if (old[i].used == 1) {
      vm_entry *e = vm_find(m, old[i].v, 1);
      free(e->pts);
      e->pts = old[i].pts;
      e->count = old[i].count;
      old[i].pts = NULL;
    }
  }
  vm_free_tab(old, os);
}

/* RemoveFarPointCloud, VoxelHashMap.cpp:93-101 (strict '>') */
static void vm_remove_far(vmap *m, const double pos[3]) {
  const double r2 = m->max_range * m->max_range;
  for (size_t i = 0; i < m->slots; i++) {
    vm_entry *e = &m->tab[i];
    if (e->used != 1) continue;
    const double dx = (double)e->pts[0] - pos[0], dy = (double)e->pts[1] - pos[1], dz = (double)e->pts[2] - pos[2];
    if (dx * dx + dy * dy + dz * dz > r2) { e->used = 2; m->live--; }
  }
}

/* AddPointCloud, VoxelHashMap.cpp:22-43.  xyz [n][3]: float (is_f64 = 0) or double. */
void oracle_vmap_add(vmap *m, const void *xyz, int64_t n, int is_f64, const double R[9], const double t[3]) {
  const float vs = (float)m->voxel_size;
  for (int64_t i = 0; i < n; i++) {
    double p[3];
    for (int c = 0; c < 3; c++) p[c] = is_f64 ? ((const double *)xyz)[3 * i + c] : (double)((const float *)xyz)[3 * i + c];
    float q[3];
    for (int r = 0; r < 3; r++) q[r] = (float)(R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2] + t[r]); /* :25 */
    const int32_t v[3] = {(int32_t)(q[0] / vs), (int32_t)(q[1] / vs), (int32_t)(q[2] / vs)};                      /* :30 */
    if (m->filled * 2 >= m->slots) vm_grow(m);
    vm_entry *e = vm_find(m, v, 1);                                                                             /* :31-39 */
    if (e->count < m->cap) { memcpy(e->pts + 3 * e->count, q, sizeof(q)); e->count++; }                         /* :32-33 */
  }
  vm_remove_far(m, t);                                                                                          /* :42 */
}

/* GetMap() (pos == NULL, :45-51) / GetMap(pose, max_range) (:53-63, strict '<').  out may be NULL to count only. */
int64_t oracle_vmap_get(vmap *m, const double *pos, double max_range, double *out) {
  int64_t n = 0;
  const double r2 = max_range * max_range;
  for (size_t i = 0; i < m->slots; i++) {
    vm_entry *e = &m->tab[i];
    if (e->used != 1 || e->count == 0) continue;
    if (pos) {
      const double dx = (double)e->pts[0] - pos[0], dy = (double)e->pts[1] - pos[1], dz = (double)e->pts[2] - pos[2];
      if (!(dx * dx + dy * dy + dz * dz < r2)) continue;
    }
    if (out)
      for (int k = 0; k < 3 * e->count; k++) out[3 * n + k] = (double)e->pts[k];
    n += e->count;
  }
  return n;
}

int64_t oracle_vmap_size(vmap *m) { return (int64_t)m->live; } /* VoxelHashMap.h:56 */

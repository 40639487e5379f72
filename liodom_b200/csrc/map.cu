// liodom::Map on sm_100a: coarse-cell hash grid with per-cell voxel-centroid clouds (replaces
// src/map.cc:24-189 and include/liodom/map.h:39-116 of the reference).
//
// Reference semantics kept:
//  - updateMap: transform the cloud with the pose (double math, float store), cell key per axis
//    int(floor(p * inv_size) * size + size / 2) (src/map.cc:103-105), find-or-create cells in
//    order of first appearance (:108-118), then re-filter EVERY modified cell with
//    pcl::VoxelGrid(leaf = resolution) over all its points, old centroids counting as single
//    points (:124-128, :56-60).  VoxelGrid output = one centroid per occupied voxel in ascending
//    (z, y, x) lattice order, lattice = floor(p * (1/leaf)) in float.
//  - getMap: all cells concatenated in creation order (:131-139).
//  - getLocalMap: pose translation truncated to int, (2 cells_xy + 1)^2 cells of the pose's z
//    layer (i outer, j inner) and the z column with the reference's bounds (:141-189).
//
// HBM layout: the whole map is ONE contiguous float4 array in cell-creation order (so getMap is a
// plain copy and a cell is a [offset, count) range); every update rewrites it into the other half
// of a ping-pong pool — touched cells from the sorted work set, untouched cells copied.  At
// 16 B/point that rewrite is HBM-bound and cheap next to the reference's per-cell std::sort.
//
// Touched cells are re-voxelised by a stable LSD radix sort of (touched rank | lattice z | y | x)
// over [old points in stored order..., new points in input order...], so the in-voxel float
// accumulation order is exactly "old centroid first, then new points in arrival order".
#include "../../include/liodom_b200.h"
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int kTile = 2048;            // keys per radix-sort block (256 threads x 8)
constexpr int kLatBits = 10;           // lattice bits per axis inside a coarse cell
constexpr int kRankShift = 3 * kLatBits;
constexpr unsigned long long kEmptyKey = ~0ull;

struct MapState {      // device-resident scalars
  int num_cells;
  int num_points;
  int n_new_slots;
  int n_touched;
  int w_old;           // old points of touched cells
  int n_groups;        // output voxels of touched cells
  int error;           // bit0: coordinate out of packable range, bit1: capacity, bit2: lattice overflow
  int n_dropped;       // non-finite input points (the reference has UB there)
};

struct MapDev {
  double xy, inv_xy, xy_half, zs, inv_z, z_half;
  float inv_leaf;
  int cap_points, cap_cells, hcap, cap_new;
  float4* pool[2];
  int* cell_count;       // [cap_cells] points per cell (creation order)
  int* cell_off;         // [cap_cells + 1] exclusive prefix of cell_count
  int* cell_newoff;      // [cap_cells + 1]
  int* cell_key;         // [cap_cells][3] reference key ints
  int* touched;          // [cap_cells] 0/1
  int* rank;             // [cap_cells] exclusive prefix of touched
  int* touched_list;     // [cap_cells] cell id per rank
  int* woff;             // [cap_cells + 1] work offsets of old points per rank
  int* group_first;      // [cap_cells + 1] first output group per rank
  unsigned long long* htab;  // [hcap] packed cell key
  int* hval;             // [hcap] cell id
  int* hfirst;           // [hcap] first input index that created the slot (pending cells)
  int* new_slots;        // [cap_new]
  float4* newpts;        // [cap_new] transformed input points
  int* pt_slot;          // [cap_new]
  int* pt_cell;          // [cap_new]
  unsigned long long* keys[2];   // [cap_points + cap_new]
  unsigned* vals[2];
  int* head;             // [cap_points + cap_new] group head flags -> exclusive scan
  int* hist;             // [256 * nblk_max]
  int* scan_tmp;         // block sums of the generic scan
  MapState* st;
  double* pose;          // [12]
};

__device__ __forceinline__ bool pack_key(int kx, int ky, int kz, unsigned long long* out) {
  const int lim = 1 << 20;
  if (kx <= -lim || kx >= lim || ky <= -lim || ky >= lim || kz <= -lim || kz >= lim) return false;
  *out = ((unsigned long long)(unsigned)(kx + lim) << 42) | ((unsigned long long)(unsigned)(ky + lim) << 21) | (unsigned long long)(unsigned)(kz + lim);
  return true;
}
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}
// src/map.cc:103-105, double math then int truncation
__device__ __forceinline__ int cell_key_axis(double v, double inv, double size, double half) {
  return (int)__dadd_rn(__dmul_rn(floor(__dmul_rn(v, inv)), size), half);
}
__device__ __forceinline__ float xform_row(const double* m, double x, double y, double z) {
  return __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z)), m[3]));
}
__device__ __forceinline__ int find_slot(const MapDev& m, unsigned long long key) {
  unsigned slot = hash_key(key) & (unsigned)(m.hcap - 1);
  for (;;) {
    const unsigned long long cur = m.htab[slot];
    if (cur == key) return (int)slot;
    if (cur == kEmptyKey) return -1;
    slot = (slot + 1) & (unsigned)(m.hcap - 1);
  }
}

// ---- update, step 1: transform, key, find-or-create hash slot -------------------------------
__global__ void __launch_bounds__(256) k_map_insert(MapDev m, const float4* in, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 s = in[i];
  float4 p;
  p.x = xform_row(m.pose, s.x, s.y, s.z); p.y = xform_row(m.pose + 4, s.x, s.y, s.z); p.z = xform_row(m.pose + 8, s.x, s.y, s.z);
  p.w = s.w;
  m.newpts[i] = p;
  m.pt_slot[i] = -1;
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { atomicAdd(&m.st->n_dropped, 1); return; }
  const int kx = cell_key_axis(p.x, m.inv_xy, m.xy, m.xy_half), ky = cell_key_axis(p.y, m.inv_xy, m.xy, m.xy_half);
  const int kz = cell_key_axis(p.z, m.inv_z, m.zs, m.z_half);
  unsigned long long key;
  if (!pack_key(kx, ky, kz, &key)) { atomicOr(&m.st->error, 1); return; }
  unsigned slot = hash_key(key) & (unsigned)(m.hcap - 1);
  for (;;) {
    unsigned long long cur = ((volatile unsigned long long*)m.htab)[slot];
    if (cur == kEmptyKey) {
      cur = atomicCAS(&m.htab[slot], kEmptyKey, key);
      if (cur == kEmptyKey) {  // created: remember it for id assignment
        const int k = atomicAdd(&m.st->n_new_slots, 1);
        if (k < m.cap_new) m.new_slots[k] = (int)slot; else atomicOr(&m.st->error, 2);
        cur = key;
      }
    }
    if (cur == key) break;
    slot = (slot + 1) & (unsigned)(m.hcap - 1);
  }
  atomicMin(&m.hfirst[slot], i);
  m.pt_slot[i] = (int)slot;
}

// step 2: new cells get ids in order of first appearance in the input (cells_vector_ order)
__global__ void k_map_assign(MapDev m) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  MapState& st = *m.st;
  const int nn = min(st.n_new_slots, m.cap_new);
  for (int a = 1; a < nn; ++a) {  // insertion sort by first input index (a handful of cells per update)
    const int s = m.new_slots[a], f = m.hfirst[s];
    int b = a - 1;
    while (b >= 0 && m.hfirst[m.new_slots[b]] > f) { m.new_slots[b + 1] = m.new_slots[b]; --b; }
    m.new_slots[b + 1] = s;
  }
  const int lim = 1 << 20;
  for (int a = 0; a < nn; ++a) {
    const int s = m.new_slots[a];
    if (st.num_cells >= m.cap_cells) { st.error |= 2; break; }
    const int id = st.num_cells++;
    m.hval[s] = id;
    const unsigned long long k = m.htab[s];
    m.cell_key[id * 3 + 0] = (int)((k >> 42) & 0x1fffff) - lim;
    m.cell_key[id * 3 + 1] = (int)((k >> 21) & 0x1fffff) - lim;
    m.cell_key[id * 3 + 2] = (int)(k & 0x1fffff) - lim;
    m.cell_count[id] = 0;
  }
  st.n_new_slots = 0;
}

// step 3: point -> cell id, touched flags
__global__ void __launch_bounds__(256) k_map_mark(MapDev m, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = m.pt_slot[i];
  int cid = -1;
  if (s >= 0) { cid = m.hval[s]; if (cid >= 0) m.touched[cid] = 1; }
  m.pt_cell[i] = cid;
}

// step 4 (one CTA): prefix sums over the cells: old offsets, touched ranks, work offsets
__global__ void __launch_bounds__(1024) k_map_plan(MapDev m) {
  __shared__ int s_a[1024], s_b[1024], s_c[1024];
  __shared__ int carry[3];
  const int nc = m.st->num_cells, tid = threadIdx.x;
  if (tid < 3) carry[tid] = 0;
  __syncthreads();
  for (int base = 0; base < nc; base += 1024) {
    const int c = base + tid;
    const int cnt = c < nc ? m.cell_count[c] : 0, t = c < nc ? m.touched[c] : 0;
    s_a[tid] = cnt; s_b[tid] = t; s_c[tid] = t ? cnt : 0;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int va = tid >= o ? s_a[tid - o] : 0, vb = tid >= o ? s_b[tid - o] : 0, vc = tid >= o ? s_c[tid - o] : 0;
      __syncthreads();
      s_a[tid] += va; s_b[tid] += vb; s_c[tid] += vc;
      __syncthreads();
    }
    if (c < nc) {
      m.cell_off[c] = carry[0] + s_a[tid] - cnt;
      const int r = carry[1] + s_b[tid] - t;
      m.rank[c] = r;
      if (t) { m.touched_list[r] = c; m.woff[r] = carry[2] + s_c[tid] - cnt; }
    }
    __syncthreads();
    if (tid == 1023) { carry[0] += s_a[1023]; carry[1] += s_b[1023]; carry[2] += s_c[1023]; }
    __syncthreads();
  }
  if (tid == 0) {
    m.cell_off[nc] = carry[0];
    m.woff[carry[1]] = carry[2];
    m.st->n_touched = carry[1];
    m.st->w_old = carry[2];
  }
}

// step 5: sort keys of the work set [old points of touched cells (rank order, stored order) | new points]
__device__ __forceinline__ int lattice(float v, float inv_leaf) { return (int)floorf(v * inv_leaf); }
__global__ void __launch_bounds__(256) k_map_workkeys(MapDev m, int cur, int w_old, int n_new, int n_touched) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= w_old + n_new) return;
  float4 p; int r, cid;
  if (w < w_old) {
    int lo = 0, hi = n_touched;  // last rank with woff[rank] <= w
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (m.woff[mid] <= w) lo = mid; else hi = mid; }
    r = lo; cid = m.touched_list[r];
    p = m.pool[cur][m.cell_off[cid] + (w - m.woff[r])];
  } else {
    const int j = w - w_old;
    cid = m.pt_cell[j];
    p = m.newpts[j];
    if (cid < 0) { m.keys[0][w] = kEmptyKey; m.vals[0][w] = (unsigned)w; return; }  // dropped point: sorts last
    r = m.rank[cid];
  }
  // lattice relative to the cell's lower corner (with a 2-voxel margin for float rounding at the faces)
  const float fx = (float)((double)m.cell_key[cid * 3 + 0] - m.xy_half), fy = (float)((double)m.cell_key[cid * 3 + 1] - m.xy_half);
  const float fz = (float)((double)m.cell_key[cid * 3 + 2] - m.z_half);
  const int lx = lattice(p.x, m.inv_leaf) - (lattice(fx, m.inv_leaf) - 2), ly = lattice(p.y, m.inv_leaf) - (lattice(fy, m.inv_leaf) - 2);
  const int lz = lattice(p.z, m.inv_leaf) - (lattice(fz, m.inv_leaf) - 2);
  const int lim = 1 << kLatBits;
  if (lx < 0 || lx >= lim || ly < 0 || ly >= lim || lz < 0 || lz >= lim) atomicOr(&m.st->error, 4);
  m.keys[0][w] = ((unsigned long long)r << kRankShift) | ((unsigned long long)(lz & (lim - 1)) << (2 * kLatBits)) |
                 ((unsigned long long)(ly & (lim - 1)) << kLatBits) | (unsigned long long)(lx & (lim - 1));
  m.vals[0][w] = (unsigned)w;
}

// ---- stable LSD radix sort, 8-bit digits -------------------------------------------------------
__global__ void __launch_bounds__(256) k_rs_hist(const unsigned long long* keys, int n, int shift, int* hist, int nblk) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kTile;
  for (int k = threadIdx.x; k < kTile; k += 256) {
    const int i = base + k;
    if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `n` ints by one CTA (n = 256 * blocks of the sort, or the head flags via 2 levels)
__global__ void __launch_bounds__(1024) k_scan1(int* data, int n, int* total) {
  __shared__ int s[1024];
  __shared__ int carry;
  const int tid = threadIdx.x;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    int v[4], sum = 0;
    for (int k = 0; k < 4; ++k) { const int i = base + tid * 4 + k; v[k] = i < n ? data[i] : 0; sum += v[k]; }
    s[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = tid >= o ? s[tid - o] : 0;
      __syncthreads();
      s[tid] += t;
      __syncthreads();
    }
    int run = carry + s[tid] - sum;
    for (int k = 0; k < 4; ++k) { const int i = base + tid * 4 + k; if (i < n) data[i] = run; run += v[k]; }
    __syncthreads();
    if (tid == 1023) carry += s[1023];
    __syncthreads();
  }
  if (tid == 0 && total) *total = carry;
}

// two-level scan for long arrays: per-block (4096 items) sums, scan of the sums, then local scans
__global__ void __launch_bounds__(1024) k_scan_blocksum(const int* data, int n, int* sums) {
  __shared__ int s[32];
  const int base = blockIdx.x * 4096, tid = threadIdx.x;
  int sum = 0;
  for (int k = 0; k < 4; ++k) { const int i = base + tid * 4 + k; if (i < n) sum += data[i]; }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) s[tid >> 5] = sum;
  __syncthreads();
  if (tid < 32) {
    int v = s[tid];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tid == 0) sums[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(1024) k_scan_local(int* data, int n, const int* sums) {
  __shared__ int s[1024];
  const int base = blockIdx.x * 4096, tid = threadIdx.x;
  int v[4], sum = 0;
  for (int k = 0; k < 4; ++k) { const int i = base + tid * 4 + k; v[k] = i < n ? data[i] : 0; sum += v[k]; }
  s[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int t = tid >= o ? s[tid - o] : 0;
    __syncthreads();
    s[tid] += t;
    __syncthreads();
  }
  int run = sums[blockIdx.x] + s[tid] - sum;
  for (int k = 0; k < 4; ++k) { const int i = base + tid * 4 + k; if (i < n) data[i] = run; run += v[k]; }
}

// stable scatter: each warp owns a contiguous 256-key segment of the tile, processed in 8 rounds
__global__ void __launch_bounds__(256) k_rs_scatter(const unsigned long long* kin, const unsigned* vin, unsigned long long* kout,
                                                     unsigned* vout, int n, int shift, const int* hist, int nblk) {
  __shared__ int wcnt[8][256];
  for (int k = threadIdx.x; k < 8 * 256; k += 256) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
  const int base = blockIdx.x * kTile;
  int dg[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = base + w * 256 + r * 32 + ln;
    const int d = i < n ? (int)((unsigned)(kin[i] >> shift) & 255u) : -1;
    dg[r] = d;
    const unsigned mm = __match_any_sync(0xffffffffu, d);
    if (d >= 0 && (__ffs(mm) - 1) == ln) wcnt[w][d] += __popc(mm);
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;
    int run = hist[d * nblk + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { const int c = wcnt[ww][d]; wcnt[ww][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = base + w * 256 + r * 32 + ln;
    const int d = dg[r];
    const unsigned mm = __match_any_sync(0xffffffffu, d);
    if (d >= 0) {
      const int pos = wcnt[w][d] + __popc(mm & ((1u << ln) - 1u));
      kout[pos] = kin[i]; vout[pos] = vin[i];
    }
    __syncwarp();
    if (d >= 0 && (__ffs(mm) - 1) == ln) wcnt[w][d] += __popc(mm);
    __syncwarp();
  }
}

// step 6: group heads over the sorted keys (a group = one voxel of one touched cell)
__global__ void __launch_bounds__(256) k_map_heads(MapDev m, int sb, int wtot) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= wtot) return;
  const unsigned long long k = m.keys[sb][w];
  m.head[w] = (k != kEmptyKey && (w == 0 || m.keys[sb][w - 1] != k)) ? 1 : 0;
}
// first group of every rank (binary search for the first key of the rank in the sorted keys)
__global__ void __launch_bounds__(256) k_map_rankfirst(MapDev m, int sb, int wtot, int n_touched, int n_groups) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_touched) return;
  if (r == n_touched) { m.group_first[r] = n_groups; return; }
  const unsigned long long target = (unsigned long long)r << kRankShift;
  int lo = 0, hi = wtot;  // first w with key >= target
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (m.keys[sb][mid] < target) lo = mid + 1; else hi = mid; }
  m.group_first[r] = lo < wtot ? m.head[lo] : n_groups;   // head[] holds the exclusive scan = group index
}
// new per-cell counts (touched: its groups; untouched: unchanged) — then scanned into cell_newoff
__global__ void __launch_bounds__(256) k_map_newcounts(MapDev m, int nc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int cnt = m.cell_count[c];
  if (m.touched[c]) { const int r = m.rank[c]; cnt = m.group_first[r + 1] - m.group_first[r]; }
  m.cell_newoff[c] = cnt;
}
// step 7: centroids of the touched cells' voxels, sequential float accumulation in sorted order
// (pcl::CentroidPoint: sum of x, y, z, intensity divided by the count as float).
__global__ void __launch_bounds__(256) k_map_centroids(MapDev m, int cur, int sb, int wtot, int w_old, const int* head_flag_src) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= wtot) return;
  const unsigned long long k = m.keys[sb][w];
  if (k == kEmptyKey || (w > 0 && m.keys[sb][w - 1] == k)) return;   // not a group head
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int cnt = 0;
  for (int u = w; u < wtot && m.keys[sb][u] == k; ++u) {
    const int src = (int)m.vals[sb][u];
    float4 p;
    if (src < w_old) {
      int lo = 0, hi = m.st->n_touched;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (m.woff[mid] <= src) lo = mid; else hi = mid; }
      p = m.pool[cur][m.cell_off[m.touched_list[lo]] + (src - m.woff[lo])];
    } else p = m.newpts[src - w_old];
    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
    ++cnt;
  }
  const float fc = (float)cnt;
  const int r = (int)(k >> kRankShift);
  const int g = m.head[w];   // exclusive scan of the head flags = global group index
  const int cid = m.touched_list[r];
  m.pool[cur ^ 1][m.cell_newoff[cid] + (g - m.group_first[r])] = make_float4(__fdiv_rn(sx, fc), __fdiv_rn(sy, fc), __fdiv_rn(sz, fc), __fdiv_rn(si, fc));
}
// step 8: untouched cells move to their new offsets; counts/offsets are committed
__global__ void __launch_bounds__(256) k_map_copy(MapDev m, int cur, int nc, int total_old) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_old) return;
  int lo = 0, hi = nc;  // cell containing old point i
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (m.cell_off[mid] <= i) lo = mid; else hi = mid; }
  if (m.touched[lo]) return;
  m.pool[cur ^ 1][m.cell_newoff[lo] + (i - m.cell_off[lo])] = m.pool[cur][i];
}
__global__ void __launch_bounds__(256) k_map_commit(MapDev m, int nc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc) {
    m.cell_count[c] = m.cell_newoff[c + 1] - m.cell_newoff[c];
    m.cell_off[c] = m.cell_newoff[c];
    m.touched[c] = 0;
  }
  if (c == 0) { m.cell_off[nc] = m.cell_newoff[nc]; m.st->num_points = m.cell_newoff[nc]; }
}
__global__ void __launch_bounds__(256) k_map_reset_first(MapDev m, int n) {   // hfirst back to INT_MAX for the slots used
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && m.pt_slot[i] >= 0) m.hfirst[m.pt_slot[i]] = INT_MAX;
}

// ---- extraction --------------------------------------------------------------------------------
// keys3: nq query keys (reference ints).  seg[q] = {offset, count} of the cell or {0, 0}.
__global__ void k_map_lookup(MapDev m, const int* keys3, int nq, int* seg_off, int* seg_cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  unsigned long long key;
  int off = 0, cnt = 0;
  if (pack_key(keys3[q * 3], keys3[q * 3 + 1], keys3[q * 3 + 2], &key)) {
    const int s = find_slot(m, key);
    if (s >= 0) { const int cid = m.hval[s]; if (cid >= 0) { off = m.cell_off[cid]; cnt = m.cell_count[cid]; } }
  }
  seg_off[q] = off; seg_cnt[q] = cnt;
}
__global__ void __launch_bounds__(256) k_map_gather(const float4* pool, const int* seg_off, const int* seg_pre, int nq, int total, float4* out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = nq;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_pre[mid] <= i) lo = mid; else hi = mid; }
    out[i] = pool[seg_off[lo] + (i - seg_pre[lo])];
  }
}

}  // namespace

// ===================================================================================================
struct liodom_map {
  MapDev m{};
  int device = 0;
  int cur = 0;
  int nblk_max = 0;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  std::string err;
  float4* stage_in = nullptr;    // device copy of the input cloud
  int* q_keys = nullptr;         // device scratch of getLocalMap
  int* q_off = nullptr; int* q_cnt = nullptr; int* q_pre = nullptr;
  float4* gather_out = nullptr;  // [cap_points]
  int h_cells = 0, h_points = 0;
  long long launches = 0;
};

static thread_local std::string g_map_err;
static int mfail(liodom_map* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf; else g_map_err = buf;
  return code;
}
#define MCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return mfail(c, LIODOM_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
static cudaError_t malloc_dev(liodom_map* c, T** p, size_t count, int fill = 0) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) return e;
  c->allocs.push_back(q);
  e = cudaMemsetAsync(q, fill, count * sizeof(T) + 256, c->stream);
  *p = static_cast<T*>(q);
  return e;
}

// exclusive scan of data[0..n) in place (n+1-th element = total written by the caller's layout)
static void scan_inplace(liodom_map* c, int* data, int n, int* total) {
  if (n <= 1 << 16) { k_scan1<<<1, 1024, 0, c->stream>>>(data, n, total); c->launches += 1; return; }
  const int nb = (n + 4095) / 4096;
  k_scan_blocksum<<<nb, 1024, 0, c->stream>>>(data, n, c->m.scan_tmp);
  k_scan1<<<1, 1024, 0, c->stream>>>(c->m.scan_tmp, nb, total);
  k_scan_local<<<nb, 1024, 0, c->stream>>>(data, n, c->m.scan_tmp);
  c->launches += 3;
}

extern "C" {

const char* liodom_map_last_error(const liodom_map* m) { return m ? m->err.c_str() : g_map_err.c_str(); }

int liodom_map_create(double voxel_xysize, double voxel_zsize, double resolution, int device, int max_points, liodom_map** out) {
  liodom_map* c = nullptr;
  if (!out || !(voxel_xysize > 0) || !(voxel_zsize > 0) || !(resolution > 0) || max_points < 1)
    return mfail(nullptr, LIODOM_E_INVALID, "bad arguments");
  if (voxel_xysize < 1.0 || voxel_zsize < 1.0)   // the reference's `int += double` cell loops (src/map.cc:157-186) never advance below 1 m
    return mfail(nullptr, LIODOM_E_INVALID, "voxel sizes below 1 m are not supported (the reference's getLocalMap loops do not terminate)");
  const double vox = std::max(voxel_xysize, voxel_zsize) / resolution;
  if (vox > (1 << kLatBits) - 8) return mfail(nullptr, LIODOM_E_INVALID, "cell size / resolution = %.0f exceeds %d voxels per axis", vox, (1 << kLatBits) - 8);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    return mfail(nullptr, LIODOM_E_NODEVICE, "no usable CUDA device (count=%d, requested %d): liodom_b200 has no CPU fallback", ndev, device);
  c = new liodom_map;
  c->device = device;
#define MCC(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      mfail(nullptr, LIODOM_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      liodom_map_destroy(c);                                                                       \
      return LIODOM_E_CUDA;                                                                        \
    }                                                                                              \
  } while (0)
  MCC(cudaSetDevice(device));
  MCC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  MapDev& m = c->m;
  m.xy = voxel_xysize; m.inv_xy = 1.0 / voxel_xysize; m.xy_half = voxel_xysize / 2.0;   // src/map.cc:70-81
  m.zs = voxel_zsize; m.inv_z = 1.0 / voxel_zsize; m.z_half = voxel_zsize / 2.0;
  m.inv_leaf = 1.0f / (float)resolution;   // pcl::VoxelGrid: inverse_leaf_size_ in float
  m.cap_points = max_points;
  m.cap_new = 1 << 16;
  m.cap_cells = 1 << 16;
  m.hcap = 1 << 18;
  const size_t capw = (size_t)m.cap_points + m.cap_new;
  c->nblk_max = (int)((capw + kTile - 1) / kTile);
  for (int k = 0; k < 2; ++k) {
    MCC(malloc_dev(c, &m.pool[k], (size_t)m.cap_points));
    MCC(malloc_dev(c, &m.keys[k], capw));
    MCC(malloc_dev(c, &m.vals[k], capw));
  }
  MCC(malloc_dev(c, &m.cell_count, (size_t)m.cap_cells));
  MCC(malloc_dev(c, &m.cell_off, (size_t)m.cap_cells + 1));
  MCC(malloc_dev(c, &m.cell_newoff, (size_t)m.cap_cells + 1));
  MCC(malloc_dev(c, &m.cell_key, (size_t)m.cap_cells * 3));
  MCC(malloc_dev(c, &m.touched, (size_t)m.cap_cells));
  MCC(malloc_dev(c, &m.rank, (size_t)m.cap_cells));
  MCC(malloc_dev(c, &m.touched_list, (size_t)m.cap_cells));
  MCC(malloc_dev(c, &m.woff, (size_t)m.cap_cells + 1));
  MCC(malloc_dev(c, &m.group_first, (size_t)m.cap_cells + 1));
  MCC(malloc_dev(c, &m.htab, (size_t)m.hcap, 0xff));
  MCC(malloc_dev(c, &m.hval, (size_t)m.hcap, 0xff));
  MCC(malloc_dev(c, &m.hfirst, (size_t)m.hcap, 0x7f));   // 0x7f7f7f7f > any input index
  MCC(malloc_dev(c, &m.new_slots, (size_t)m.cap_new));
  MCC(malloc_dev(c, &m.newpts, (size_t)m.cap_new));
  MCC(malloc_dev(c, &m.pt_slot, (size_t)m.cap_new));
  MCC(malloc_dev(c, &m.pt_cell, (size_t)m.cap_new));
  MCC(malloc_dev(c, &m.head, capw));
  MCC(malloc_dev(c, &m.hist, (size_t)256 * c->nblk_max));
  MCC(malloc_dev(c, &m.scan_tmp, capw / 4096 + 8));
  MCC(malloc_dev(c, &m.st, 1));
  MCC(malloc_dev(c, &m.pose, 12));
  MCC(malloc_dev(c, &c->stage_in, (size_t)m.cap_new));
  MCC(malloc_dev(c, &c->q_keys, 3 * 4096));
  MCC(malloc_dev(c, &c->q_off, 4096));
  MCC(malloc_dev(c, &c->q_cnt, 4096));
  MCC(malloc_dev(c, &c->q_pre, 4097));
  MCC(malloc_dev(c, &c->gather_out, (size_t)m.cap_points));
  MCC(cudaStreamSynchronize(c->stream));
#undef MCC
  *out = c;
  return 0;
}

void liodom_map_destroy(liodom_map* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (void* p : c->allocs) cudaFree(p);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// Map::updateMap (src/map.cc:90-129)
int liodom_map_update(liodom_map* c, const float* pts_xyzi, int n, const double* pose16) {
  if (!c || !pose16 || n < 0) return mfail(c, LIODOM_E_INVALID, "bad arguments");
  MCK(cudaSetDevice(c->device));
  MapDev& m = c->m;
  if (n > m.cap_new) return mfail(c, LIODOM_E_CAPACITY, "cloud of %d points exceeds the per-update capacity %d", n, m.cap_new);
  if (n == 0) return 0;
  cudaStream_t s = c->stream;
  MCK(cudaMemcpyAsync(c->stage_in, pts_xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(m.pose, pose16, sizeof(double) * 12, cudaMemcpyHostToDevice, s));
  const int gb = (n + 255) / 256;
  k_map_insert<<<gb, 256, 0, s>>>(m, c->stage_in, n);
  k_map_assign<<<1, 32, 0, s>>>(m);
  k_map_mark<<<gb, 256, 0, s>>>(m, n);
  k_map_plan<<<1, 1024, 0, s>>>(m);
  c->launches += 4;
  MapState st;
  MCK(cudaMemcpyAsync(&st, m.st, sizeof(st), cudaMemcpyDeviceToHost, s));
  MCK(cudaStreamSynchronize(s));
  if (st.error & 1) return mfail(c, LIODOM_E_INVALID, "map coordinates beyond the +-2^20 m key range");
  if (st.error & 2) return mfail(c, LIODOM_E_CAPACITY, "cell capacity exceeded (%d cells)", m.cap_cells);
  const int wtot = st.w_old + n, nc = st.num_cells, total_old = st.num_points;
  if ((size_t)total_old + n > (size_t)m.cap_points) return mfail(c, LIODOM_E_CAPACITY, "map of %d points + %d new exceeds max_points %d", total_old, n, m.cap_points);
  k_map_workkeys<<<(wtot + 255) / 256, 256, 0, s>>>(m, c->cur, st.w_old, n, st.n_touched);
  c->launches += 1;
  // key bits in use: 30 lattice bits + ceil(log2(n_touched))
  int bits = kRankShift;
  while ((1 << (bits - kRankShift)) < st.n_touched) ++bits;
  const int nblk = (wtot + kTile - 1) / kTile;
  int sb = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    k_rs_hist<<<nblk, 256, 0, s>>>(m.keys[sb], wtot, shift, m.hist, nblk);
    scan_inplace(c, m.hist, 256 * nblk, nullptr);
    k_rs_scatter<<<nblk, 256, 0, s>>>(m.keys[sb], m.vals[sb], m.keys[sb ^ 1], m.vals[sb ^ 1], wtot, shift, m.hist, nblk);
    c->launches += 2;
    sb ^= 1;
  }
  k_map_heads<<<(wtot + 255) / 256, 256, 0, s>>>(m, sb, wtot);
  scan_inplace(c, m.head, wtot, &m.st->n_groups);
  int n_groups = 0;
  MCK(cudaMemcpyAsync(&n_groups, &m.st->n_groups, sizeof(int), cudaMemcpyDeviceToHost, s));
  MCK(cudaStreamSynchronize(s));
  k_map_rankfirst<<<(st.n_touched + 256) / 256, 256, 0, s>>>(m, sb, wtot, st.n_touched, n_groups);
  k_map_newcounts<<<(nc + 255) / 256, 256, 0, s>>>(m, nc);
  scan_inplace(c, m.cell_newoff, nc, &m.cell_newoff[nc]);
  k_map_centroids<<<(wtot + 255) / 256, 256, 0, s>>>(m, c->cur, sb, wtot, st.w_old, nullptr);
  if (total_old > 0) k_map_copy<<<(total_old + 255) / 256, 256, 0, s>>>(m, c->cur, nc, total_old);
  k_map_commit<<<(nc + 255) / 256, 256, 0, s>>>(m, nc);
  k_map_reset_first<<<gb, 256, 0, s>>>(m, n);
  c->launches += 7;
  MCK(cudaGetLastError());
  MCK(cudaMemcpyAsync(&st, m.st, sizeof(st), cudaMemcpyDeviceToHost, s));
  MCK(cudaStreamSynchronize(s));
  if (st.error & 4) return mfail(c, LIODOM_E_INVALID, "voxel lattice overflow inside a cell (cell size / resolution too large)");
  c->cur ^= 1;
  c->h_cells = st.num_cells; c->h_points = st.num_points;
  return 0;
}

int liodom_map_size(liodom_map* c, int* n_points, int* n_cells) {
  if (!c) return LIODOM_E_INVALID;
  if (n_points) *n_points = c->h_points;
  if (n_cells) *n_cells = c->h_cells;
  return 0;
}

// Map::getMap (src/map.cc:131-139): the pool already is the concatenation in creation order
int liodom_map_get(liodom_map* c, float* xyzi, int cap, int* n_points) {
  if (!c) return LIODOM_E_INVALID;
  MCK(cudaSetDevice(c->device));
  if (n_points) *n_points = c->h_points;
  if (xyzi && c->h_points > 0) {
    if (cap < c->h_points) return mfail(c, LIODOM_E_CAPACITY, "output capacity %d < map size %d", cap, c->h_points);
    MCK(cudaMemcpyAsync(xyzi, c->m.pool[c->cur], (size_t)c->h_points * 16, cudaMemcpyDeviceToHost, c->stream));
    MCK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

// Map::getLocalMap (src/map.cc:141-189).  The key list is host integer arithmetic restated from
// the reference (translation truncated to int; the z loop bounds use voxel_xysize_ and `int += double`).
// The cloud is gathered on the device into `dev_out` (capacity `cap` points); returns its size.
static int map_local_to_device(liodom_map* c, const double* pose16, int cells_xy, int cells_z, float4* dev_out, int cap, int* n_points) {
  const MapDev& m = c->m;
  const int x = (int)pose16[3], y = (int)pose16[7], z = (int)pose16[11];
  const int vx = (int)(std::floor(x * m.inv_xy) * m.xy + m.xy_half), vy = (int)(std::floor(y * m.inv_xy) * m.xy + m.xy_half);
  const int vz = (int)(std::floor(z * m.inv_z) * m.zs + m.z_half);
  std::vector<int> keys;
  const int init_x = (int)(vx - cells_xy * m.xy), end_x = (int)(vx + cells_xy * m.xy);
  const int init_y = (int)(vy - cells_xy * m.xy), end_y = (int)(vy + cells_xy * m.xy);
  for (int i = init_x; i <= end_x && keys.size() < 3 * 4096; i = (int)(i + m.xy))
    for (int j = init_y; j <= end_y && keys.size() < 3 * 4096; j = (int)(j + m.xy)) { keys.push_back(i); keys.push_back(j); keys.push_back(vz); }
  const int init_z = (int)(vz - cells_z * m.xy), end_z = (int)(vz + cells_z * m.xy);
  for (int i = init_z; i <= end_z && keys.size() < 3 * 4096; i = (int)(i + m.zs)) { keys.push_back(vx); keys.push_back(vy); keys.push_back(i); }
  const int nq = (int)keys.size() / 3;
  if (nq >= 4096) return mfail(c, LIODOM_E_CAPACITY, "getLocalMap asks for %d cells (limit 4095)", nq);
  MCK(cudaMemcpyAsync(c->q_keys, keys.data(), sizeof(int) * keys.size(), cudaMemcpyHostToDevice, c->stream));
  k_map_lookup<<<(nq + 127) / 128, 128, 0, c->stream>>>(m, c->q_keys, nq, c->q_off, c->q_cnt);
  MCK(cudaMemcpyAsync(c->q_pre, c->q_cnt, sizeof(int) * nq, cudaMemcpyDeviceToDevice, c->stream));
  k_scan1<<<1, 1024, 0, c->stream>>>(c->q_pre, nq, c->q_pre + nq);
  int total = 0;
  MCK(cudaMemcpyAsync(&total, c->q_pre + nq, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MCK(cudaStreamSynchronize(c->stream));
  c->launches += 2;
  if (n_points) *n_points = total;
  if (dev_out && total > 0) {
    if (cap < total) return mfail(c, LIODOM_E_CAPACITY, "output capacity %d < local map size %d", cap, total);
    int blocks = (total + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    k_map_gather<<<blocks, 256, 0, c->stream>>>(m.pool[c->cur], c->q_off, c->q_pre, nq, total, dev_out);
    c->launches += 1;
    MCK(cudaGetLastError());
    MCK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int liodom_map_get_local(liodom_map* c, const double* pose16, int cells_xy, int cells_z, float* xyzi, int cap, int* n_points) {
  if (!c || !pose16) return mfail(c, LIODOM_E_INVALID, "bad arguments");
  MCK(cudaSetDevice(c->device));
  int total = 0;
  int rc = map_local_to_device(c, pose16, cells_xy, cells_z, nullptr, 0, &total);
  if (rc) return rc;
  if (n_points) *n_points = total;
  if (xyzi && total > 0) {
    if (cap < total) return mfail(c, LIODOM_E_CAPACITY, "output capacity %d < local map size %d", cap, total);
    if (total > c->m.cap_points) return mfail(c, LIODOM_E_CAPACITY, "local map of %d points exceeds max_points (duplicated centre cell)", total);
    rc = map_local_to_device(c, pose16, cells_xy, cells_z, c->gather_out, c->m.cap_points, &total);
    if (rc) return rc;
    MCK(cudaMemcpyAsync(xyzi, c->gather_out, (size_t)total * 16, cudaMemcpyDeviceToHost, c->stream));
    MCK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

// Same, but the cloud stays on the device (`dev_xyzi`: device pointer, e.g. the buffer returned by
// liodom_received_map_buffer): the map -> odometry feedback of mapping=1 without a host round trip.
int liodom_map_get_local_device(liodom_map* c, const double* pose16, int cells_xy, int cells_z, void* dev_xyzi, int cap, int* n_points) {
  if (!c || !pose16 || !dev_xyzi) return mfail(c, LIODOM_E_INVALID, "bad arguments");
  MCK(cudaSetDevice(c->device));
  return map_local_to_device(c, pose16, cells_xy, cells_z, static_cast<float4*>(dev_xyzi), cap, n_points);
}

int liodom_map_cells(liodom_map* c, int32_t* keys3, int32_t* counts, int cap, int* n_cells) {
  if (!c) return LIODOM_E_INVALID;
  MCK(cudaSetDevice(c->device));
  if (n_cells) *n_cells = c->h_cells;
  if (c->h_cells > 0 && (keys3 || counts)) {
    if (cap < c->h_cells) return mfail(c, LIODOM_E_CAPACITY, "output capacity %d < %d cells", cap, c->h_cells);
    if (keys3) MCK(cudaMemcpyAsync(keys3, c->m.cell_key, sizeof(int) * 3 * c->h_cells, cudaMemcpyDeviceToHost, c->stream));
    if (counts) MCK(cudaMemcpyAsync(counts, c->m.cell_count, sizeof(int) * c->h_cells, cudaMemcpyDeviceToHost, c->stream));
    MCK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

}  // extern "C"

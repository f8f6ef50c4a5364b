// gpu_bvh.cuh — tree construction ON the device for big triangle meshes.  Product code.
//
// Role in the reference: Shape_tree.create (path_tracer/src/shape_tree.ml:252-263) as ganesha calls it on the PLY
// mesh, where it is the visible cost ("build time", ganesha/bin/main.ml:190-194).  The host builder (bvh.cpp,
// binned SAH) needs 0.45 s for 10^6 triangles on 16 cores.  Two device builders share steps 1, 6 and 7:
//   1. primitive boxes + centroid / box bounds    (one pass, ordered-int atomics)
//   SAH (default): binned SAH with the host builder's split rule, level by level (second half of this file);
//                  then a cub radix sort by path key.  39 ms for 10^6 triangles, tree as good as the host's.
//   LBVH (PTB_BUILDER=gpu-lbvh): 2. 63-bit Morton keys of the centroids, 3. cub radix sort, 4. binary radix tree
//                  (Karras 2012), 5. boxes bottom-up.  28 ms, but the tree traverses 1.4x slower.
//   6. collapse into the 4-wide nodes the traversal kernel reads, level by level from the root: a node takes its two
//      binary children and twice opens the one with the largest surface area
//   7. primitive tables gathered in sorted order
// Closest-hit results do not depend on the tree, so parity with the oracle is unchanged.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "device_types.cuh"
#include "presplit.hpp"

namespace ptb {
namespace gbvh {

constexpr int GLEAF = 4;  // primitives per leaf (= LEAF_MAX of the host builder)

struct Bounds6 {
  unsigned lo[3], hi[3];  // ordered-uint encodings
};
__device__ __forceinline__ unsigned enc_f(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __host__ __device__ __forceinline__ float dec_f(unsigned u) {
  const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

// 1. per-triangle box (float, conservative: rounded outward from the double vertices) + centroid bounds
__global__ void __launch_bounds__(256) k_tri_boxes(const double *__restrict__ vx, const double *__restrict__ vy,
                                                   const double *__restrict__ vz, const int32_t *__restrict__ idx, int n,
                                                   float4 *__restrict__ blo, float4 *__restrict__ bhi, Bounds6 *cb /* [0] centroids, [1] boxes */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float c[3] = {0, 0, 0}, bl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bh[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  const bool ok = i < n;
  if (ok) {
    const int a = idx[3 * i], b = idx[3 * i + 1], d = idx[3 * i + 2];
    const double x[3] = {vx[a], vx[b], vx[d]}, y[3] = {vy[a], vy[b], vy[d]}, z[3] = {vz[a], vz[b], vz[d]};
    const double lo[3] = {fmin(fmin(x[0], x[1]), x[2]), fmin(fmin(y[0], y[1]), y[2]), fmin(fmin(z[0], z[1]), z[2])};
    const double hi[3] = {fmax(fmax(x[0], x[1]), x[2]), fmax(fmax(y[0], y[1]), y[2]), fmax(fmax(z[0], z[1]), z[2])};
    blo[i] = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]), 0.f);
    bhi[i] = make_float4(__double2float_ru(hi[0]), __double2float_ru(hi[1]), __double2float_ru(hi[2]), 0.f);
    for (int k = 0; k < 3; ++k) c[k] = (float)(0.5 * (lo[k] + hi[k]));
    bl[0] = blo[i].x, bl[1] = blo[i].y, bl[2] = blo[i].z, bh[0] = bhi[i].x, bh[1] = bhi[i].y, bh[2] = bhi[i].z;
  }
  // warp reduce, then one atomic per warp and bound
  for (int k = 0; k < 3; ++k) {
    float mn = ok ? c[k] : 3.0e38f, mx = ok ? c[k] : -3.0e38f;
    for (int o = 16; o; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    float bmn = bl[k], bmx = bh[k];
    for (int o = 16; o; o >>= 1) {
      bmn = fminf(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
      bmx = fmaxf(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&cb[0].lo[k], enc_f(mn));
      atomicMax(&cb[0].hi[k], enc_f(mx));
      atomicMin(&cb[1].lo[k], enc_f(bmn));
      atomicMax(&cb[1].hi[k], enc_f(bmx));
    }
  }
}
// 1b. triangle pre-splitting (presplit.hpp): how many boxes contain a random point (the trigger), pieces per triangle
// for a cell size, and the pieces' boxes as the builder's primitives
__global__ void __launch_bounds__(256) k_box_volume_sum(const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n,
                                                        double cap, double *__restrict__ sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n) v = fmin(cap, (double)(bhi[i].x - blo[i].x) * (double)(bhi[i].y - blo[i].y) * (double)(bhi[i].z - blo[i].z));
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomicAdd(sum, v);
}
__device__ __forceinline__ void load_tri(const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                                         const int32_t *__restrict__ idx, int i, double v[3][3]) {
  for (int k = 0; k < 3; ++k) {
    const int a = idx[3 * i + k];
    v[k][0] = vx[a], v[k][1] = vy[a], v[k][2] = vz[a];
  }
}
__global__ void __launch_bounds__(128) k_presplit_count(const double *__restrict__ vx, const double *__restrict__ vy,
                                                        const double *__restrict__ vz, const int32_t *__restrict__ idx, int n,
                                                        double cell, double ox, double oy, double oz, int *__restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[3][3];
  load_tri(vx, vy, vz, idx, i, v);
  const double org[3] = {ox, oy, oz};
  counts[i] = presplit::pieces(v, cell, org, [](const double *, const double *) {});
}
__global__ void __launch_bounds__(128) k_presplit_emit(const double *__restrict__ vx, const double *__restrict__ vy,
                                                       const double *__restrict__ vz, const int32_t *__restrict__ idx, int n,
                                                       double cell, double ox, double oy, double oz, const int *__restrict__ offsets,
                                                       float4 *__restrict__ blo, float4 *__restrict__ bhi, int32_t *__restrict__ ref_tri) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[3][3];
  load_tri(vx, vy, vz, idx, i, v);
  const double org[3] = {ox, oy, oz};
  int r = offsets[i];
  presplit::pieces(v, cell, org, [&](const double *lo, const double *hi) {
    blo[r] = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]), 0.f);
    bhi[r] = make_float4(__double2float_ru(hi[0]), __double2float_ru(hi[1]), __double2float_ru(hi[2]), 0.f);
    ref_tri[r] = i;
    ++r;
  });
}
// centroid / box bounds of a list of boxes (k_tri_boxes does the same for the triangles' own boxes)
__global__ void __launch_bounds__(256) k_box_bounds(const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n, Bounds6 *cb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < n;
  float bl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bh[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (ok) bl[0] = blo[i].x, bl[1] = blo[i].y, bl[2] = blo[i].z, bh[0] = bhi[i].x, bh[1] = bhi[i].y, bh[2] = bhi[i].z;
  for (int k = 0; k < 3; ++k) {
    const float c = 0.5f * (bl[k] + bh[k]);
    float mn = ok ? c : 3.0e38f, mx = ok ? c : -3.0e38f, bmn = bl[k], bmx = bh[k];
    for (int o = 16; o; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      bmn = fminf(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
      bmx = fmaxf(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&cb[0].lo[k], enc_f(mn));
      atomicMax(&cb[0].hi[k], enc_f(mx));
      atomicMin(&cb[1].lo[k], enc_f(bmn));
      atomicMax(&cb[1].hi[k], enc_f(bmx));
    }
  }
}
__global__ void __launch_bounds__(256) k_iota(unsigned *v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (unsigned)i;
}

__device__ __forceinline__ unsigned long long split3(unsigned a) {
  unsigned long long x = a & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
// 2. Morton keys
__global__ void __launch_bounds__(256) k_morton(const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n,
                                                const Bounds6 *cb, unsigned long long *__restrict__ keys,
                                                unsigned *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lo[3] = {dec_f(cb->lo[0]), dec_f(cb->lo[1]), dec_f(cb->lo[2])};
  const float hi[3] = {dec_f(cb->hi[0]), dec_f(cb->hi[1]), dec_f(cb->hi[2])};
  const float c[3] = {0.5f * (blo[i].x + bhi[i].x), 0.5f * (blo[i].y + bhi[i].y), 0.5f * (blo[i].z + bhi[i].z)};
  unsigned q[3];
  for (int k = 0; k < 3; ++k) {
    const float ext = hi[k] - lo[k];
    float t = ext > 0.f ? (c[k] - lo[k]) / ext : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    q[k] = min((unsigned)(t * 2097152.0f), 2097151u);
  }
  keys[i] = (split3(q[0]) << 2) | (split3(q[1]) << 1) | split3(q[2]);
  vals[i] = (unsigned)i;
}

// common-prefix length of sorted keys i and j (index as tie-break so that equal keys still form a tree)
__device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
  return __clzll((long long)(a ^ b));
}
// 4. binary radix tree: internal node i in [0, n-2].  child refs: >= 0 internal node, < 0 leaf ~ref
__global__ void __launch_bounds__(256) k_radix_tree(const unsigned long long *__restrict__ keys, int n,
                                                    int *__restrict__ lch, int *__restrict__ rch, int *__restrict__ first,
                                                    int *__restrict__ count, int *__restrict__ par_int,
                                                    int *__restrict__ par_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) > 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int L = (lo == gamma) ? ~gamma : gamma, Rr = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  lch[i] = L, rch[i] = Rr, first[i] = lo, count[i] = hi - lo + 1;
  if (L >= 0) par_int[L] = i; else par_leaf[~L] = i;
  if (Rr >= 0) par_int[Rr] = i; else par_leaf[~Rr] = i;
  if (i == 0) par_int[0] = -1;
}

// 5. boxes bottom-up: the second thread to arrive at a node unions the children (the first one stops)
__global__ void __launch_bounds__(256) k_fit(const unsigned *__restrict__ vals, const float4 *__restrict__ blo,
                                             const float4 *__restrict__ bhi, int n, const int *__restrict__ lch,
                                             const int *__restrict__ rch, const int *__restrict__ par_int,
                                             const int *__restrict__ par_leaf, unsigned *__restrict__ visits,
                                             float4 *__restrict__ nlo, float4 *__restrict__ nhi) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int node = par_leaf[k];
  while (node >= 0) {
    if (atomicAdd(&visits[node], 1u) == 0u) return;  // first arrival: the sibling subtree is not done yet
    __threadfence();
    const int L = lch[node], Rr = rch[node];
    const float4 a0 = L >= 0 ? __ldcg(&nlo[L]) : blo[vals[~L]], a1 = L >= 0 ? __ldcg(&nhi[L]) : bhi[vals[~L]];
    const float4 b0 = Rr >= 0 ? __ldcg(&nlo[Rr]) : blo[vals[~Rr]], b1 = Rr >= 0 ? __ldcg(&nhi[Rr]) : bhi[vals[~Rr]];
    nlo[node] = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.f);
    nhi[node] = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.f);
    __threadfence();
    node = par_int[node];
  }
}

constexpr int LEAF_MARK = INT32_MIN;  // lch of a binary node that is a leaf (SAH build)
struct Frontier {
  int bin;   // binary internal node that becomes a wide node
  int wide;  // its index among the wide nodes
};
// conservative narrowing of a box plane, as render.cu box_lo / box_hi do for the host-built tree
__device__ __forceinline__ float pad_lo(float x, float ext) { return nextafterf(x - 4e-7f * (fabsf(x) + ext), -INFINITY); }
__device__ __forceinline__ float pad_hi(float x, float ext) { return nextafterf(x + 4e-7f * (fabsf(x) + ext), INFINITY); }

// 6. one level of the collapse: every wide node of this level takes its (up to four) grandchildren
__global__ void __launch_bounds__(128) k_collapse_level(const Frontier *__restrict__ cur, int n_cur, const int *__restrict__ lch,
                                                        const int *__restrict__ rch, const int *__restrict__ first,
                                                        const int *__restrict__ count, const unsigned *__restrict__ vals,
                                                        const float4 *__restrict__ blo, const float4 *__restrict__ bhi,
                                                        const float4 *__restrict__ nlo, const float4 *__restrict__ nhi,
                                                        Node4<float> *__restrict__ nodes, Frontier *__restrict__ next,
                                                        int *__restrict__ counters /* [0] wide nodes, [1] next frontier, [2] leaves */,
                                                        int leaf_thresh /* nodes of at most this many primitives are leaves;
                                                                           nodes marked LEAF_MARK in lch always are */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cur) return;
  const Frontier f = cur[t];
  // children: start from the two binary children and, twice, open the expandable one with the largest surface
  // area (the same greedy rule as the host collapse, bvh.cpp): big boxes are the ones worth splitting
  int ch[4] = {lch[f.bin], rch[f.bin], 0, 0}, nch = 2;
  for (int round = 0; round < 2; ++round) {
    int pick = -1;
    float best = -1.f;
    for (int k = 0; k < nch; ++k) {
      const int c = ch[k];
      if (c < 0 || count[c] <= leaf_thresh || lch[c] == LEAF_MARK) continue;
      const float4 a = nlo[c], b = nhi[c];
      const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
      const float area = ex * ey + ey * ez + ez * ex;
      if (area > best) best = area, pick = k;
    }
    if (pick < 0) break;
    const int c = ch[pick];
    ch[pick] = lch[c];
    ch[nch++] = rch[c];
  }
  Node4<float> out;
  for (int k = 0; k < 4; ++k) {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    int ref = INT32_MIN;  // EMPTY_CHILD
    if (k < nch) {
      const int c = ch[k];
      float4 a, b;
      if (c < 0) {  // single primitive
        a = blo[vals[~c]], b = bhi[vals[~c]];
        ref = ~(int)((unsigned)(~c) | (0u << 26) | (1u << 30));
        atomicAdd(&counters[2], 1);
      } else {
        a = nlo[c], b = nhi[c];
        if (count[c] <= leaf_thresh || lch[c] == LEAF_MARK) {  // leaf over its (contiguous) sorted range
          ref = ~(int)((unsigned)first[c] | ((unsigned)(count[c] - 1) << 26) | (1u << 30));
          atomicAdd(&counters[2], 1);
        } else {
          const int w = atomicAdd(&counters[0], 1);
          next[atomicAdd(&counters[1], 1)] = Frontier{c, w};
          ref = w;
        }
      }
      const float ext = fmaxf(fmaxf(b.x - a.x, b.y - a.y), b.z - a.z);
      lo[0] = pad_lo(a.x, ext), lo[1] = pad_lo(a.y, ext), lo[2] = pad_lo(a.z, ext);
      hi[0] = pad_hi(b.x, ext), hi[1] = pad_hi(b.y, ext), hi[2] = pad_hi(b.z, ext);
    }
    for (int a = 0; a < 3; ++a) out.lo[a][k] = lo[a], out.hi[a][k] = hi[a];
    out.child[k] = ref;
  }
  nodes[f.wide] = out;
}

// 7. primitive tables in sorted order (same arithmetic as ensure_tables<float> on the host: edges from the doubles)
__global__ void __launch_bounds__(256) k_emit_tris(const unsigned *__restrict__ vals, int n, const int32_t *__restrict__ ref_tri,
                                                   const double *__restrict__ vx,
                                                   const double *__restrict__ vy, const double *__restrict__ vz,
                                                   const int32_t *__restrict__ idx, const int32_t *__restrict__ tmat,
                                                   const double *__restrict__ tuv, const uint8_t *__restrict__ mat_kind,
                                                   Vec4<float> *__restrict__ tris, float *__restrict__ uv,
                                                   int32_t *__restrict__ tri_id, int32_t *__restrict__ tri_mat,
                                                   uint8_t *__restrict__ prim_kind) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = ref_tri ? ref_tri[vals[k]] : (int)vals[k];  // (pre-split meshes: primitive = a reference to a triangle)
  const int a = idx[3 * i], b = idx[3 * i + 1], c = idx[3 * i + 2];
  tris[3 * (size_t)k + 0] = {(float)vx[a], (float)vy[a], (float)vz[a], 0.f};
  tris[3 * (size_t)k + 1] = {(float)(vx[b] - vx[a]), (float)(vy[b] - vy[a]), (float)(vz[b] - vz[a]), 0.f};
  tris[3 * (size_t)k + 2] = {(float)(vx[c] - vx[a]), (float)(vy[c] - vy[a]), (float)(vz[c] - vz[a]), 0.f};
  for (int j = 0; j < 6; ++j) uv[6 * (size_t)k + j] = (float)tuv[6 * (size_t)i + j];
  tri_id[k] = i;
  const int m = tmat[i];
  tri_mat[k] = m;
  prim_kind[k] = mat_kind[m];
}

// =====================================================================================================================
// Binned-SAH build on the device: the same split rule as the host builder (32 bins per axis, cost = sum of
// count x area of the two sides), all nodes of a level at once.  Per level: (a) every primitive of an active node
// adds its box to the node's bins (3 axes), (b) one warp per node scans the bins and picks the split, creating the
// two children, (c) every primitive moves to its child and records the turn in a path key.  Sorting by path key
// makes every subtree's primitives contiguous.  Bins are over the node's box (centroids always lie inside it).
// =====================================================================================================================
constexpr int SB = 32;                // bins per axis = lanes of the warp that scans them
constexpr int BIN_WORDS = 7;          // lo[3], hi[3] (ordered-uint floats), count
constexpr int NODE_BINS = 3 * SB * BIN_WORDS;

struct SahTree {
  float4 *lo, *hi;        // node boxes
  int *cnt, *first, *l, *r;
  int *slot;              // index of the node's bins at its level, -1 = not active
  int *axis, *split;      // chosen split (axis -2: halve by ticket when all centroids coincide)
  int *level;
  unsigned *ticket;
};

__device__ __forceinline__ int sah_bin(float c, float lo, float hi) {
  const float ext = hi - lo;
  if (!(ext > 0.f)) return 0;
  return min(SB - 1, max(0, (int)((c - lo) * ((float)SB * 0.999999f) / ext)));
}

__global__ void __launch_bounds__(256) k_sah_clear(unsigned *bins, int n_active) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_active * 3 * SB) return;
  unsigned *b = bins + i * BIN_WORDS;
  b[0] = b[1] = b[2] = 0xffffffffu;
  b[3] = b[4] = b[5] = 0u;
  b[6] = 0u;
}

// (a) binning.  PRIV: the level has at most 4 active nodes, so each block accumulates in shared memory first
// (a million primitives hitting 96 global bins would serialise on the atomics)
template <bool PRIV>
__global__ void __launch_bounds__(256) k_sah_bin(const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n,
                                                 const int *__restrict__ node_id, SahTree T, unsigned *__restrict__ bins) {
  __shared__ unsigned sb[PRIV ? 4 * NODE_BINS : 1];
  if (PRIV) {
    for (int k = threadIdx.x; k < 4 * NODE_BINS; k += blockDim.x) {
      const int w = k % BIN_WORDS;
      sb[k] = w < 3 ? 0xffffffffu : 0u;
    }
    __syncthreads();
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int node = node_id[i];
    const int slot = T.slot[node];
    if (slot >= 0) {
      const float4 a = blo[i], b = bhi[i], nl = T.lo[node], nh = T.hi[node];
      const float c[3] = {0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z)};
      const float l3[3] = {nl.x, nl.y, nl.z}, h3[3] = {nh.x, nh.y, nh.z};
      const unsigned e[6] = {enc_f(a.x), enc_f(a.y), enc_f(a.z), enc_f(b.x), enc_f(b.y), enc_f(b.z)};
      for (int ax = 0; ax < 3; ++ax) {
        const int bi = sah_bin(c[ax], l3[ax], h3[ax]);
        unsigned *dst = (PRIV ? sb : bins) + ((size_t)slot * 3 + ax) * SB * BIN_WORDS + (size_t)bi * BIN_WORDS;
        atomicMin(dst + 0, e[0]), atomicMin(dst + 1, e[1]), atomicMin(dst + 2, e[2]);
        atomicMax(dst + 3, e[3]), atomicMax(dst + 4, e[4]), atomicMax(dst + 5, e[5]);
        atomicAdd(dst + 6, 1u);
      }
    }
  }
  if (PRIV) {
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * NODE_BINS; k += blockDim.x) {
      const int w = k % BIN_WORDS;
      const unsigned v = sb[k];
      if (w < 3) { if (v != 0xffffffffu) atomicMin(bins + k, v); }
      else if (w < 6) { if (v != 0u) atomicMax(bins + k, v); }
      else if (v) atomicAdd(bins + k, v);
    }
  }
}

struct WBox {
  float lx, ly, lz, hx, hy, hz;
  int c;
};
__device__ __forceinline__ WBox wbox_merge(WBox a, WBox b) {
  return {fminf(a.lx, b.lx), fminf(a.ly, b.ly), fminf(a.lz, b.lz), fmaxf(a.hx, b.hx), fmaxf(a.hy, b.hy), fmaxf(a.hz, b.hz), a.c + b.c};
}
__device__ __forceinline__ WBox wbox_shfl_up(WBox a, int d) {
  return {__shfl_up_sync(0xffffffffu, a.lx, d), __shfl_up_sync(0xffffffffu, a.ly, d), __shfl_up_sync(0xffffffffu, a.lz, d),
          __shfl_up_sync(0xffffffffu, a.hx, d), __shfl_up_sync(0xffffffffu, a.hy, d), __shfl_up_sync(0xffffffffu, a.hz, d),
          __shfl_up_sync(0xffffffffu, a.c, d)};
}
__device__ __forceinline__ WBox wbox_shfl_down(WBox a, int d) {
  return {__shfl_down_sync(0xffffffffu, a.lx, d), __shfl_down_sync(0xffffffffu, a.ly, d), __shfl_down_sync(0xffffffffu, a.lz, d),
          __shfl_down_sync(0xffffffffu, a.hx, d), __shfl_down_sync(0xffffffffu, a.hy, d), __shfl_down_sync(0xffffffffu, a.hz, d),
          __shfl_down_sync(0xffffffffu, a.c, d)};
}
__device__ __forceinline__ WBox wbox_shfl(WBox a, int src) {
  return {__shfl_sync(0xffffffffu, a.lx, src), __shfl_sync(0xffffffffu, a.ly, src), __shfl_sync(0xffffffffu, a.lz, src),
          __shfl_sync(0xffffffffu, a.hx, src), __shfl_sync(0xffffffffu, a.hy, src), __shfl_sync(0xffffffffu, a.hz, src),
          __shfl_sync(0xffffffffu, a.c, src)};
}
__device__ __forceinline__ float wbox_area(WBox a) {
  if (a.c == 0) return 0.f;
  const float x = a.hx - a.lx, y = a.hy - a.ly, z = a.hz - a.lz;
  return x * y + y * z + z * x;
}

// (b) one warp per active node: lane = bin.  Prefix (left side) and suffix (right side) unions by warp scans.
__global__ void __launch_bounds__(256) k_sah_select(const int *__restrict__ active, int n_active, const unsigned *__restrict__ bins,
                                                    SahTree T, int level, int *__restrict__ next_active,
                                                    int *__restrict__ counters /* [0] nodes, [1] next active */) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_active) return;
  const int node = active[w];
  const WBox EMPTY = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f, 0};
  float best_cost = 3e38f;
  int best_axis = -1, best_split = -1;
  WBox best_l = EMPTY, best_r = EMPTY;
  for (int ax = 0; ax < 3; ++ax) {
    const unsigned *b = bins + (((size_t)w * 3 + ax) * SB + lane) * BIN_WORDS;
    WBox me = EMPTY;
    const int c = (int)b[6];
    if (c) me = {dec_f(b[0]), dec_f(b[1]), dec_f(b[2]), dec_f(b[3]), dec_f(b[4]), dec_f(b[5]), c};
    WBox pre = me, suf = me;
    for (int d = 1; d < 32; d <<= 1) {
      const WBox u = wbox_shfl_up(pre, d), v = wbox_shfl_down(suf, d);
      if (lane >= d) pre = wbox_merge(pre, u);
      if (lane + d < 32) suf = wbox_merge(suf, v);
    }
    // candidate split after bin `lane`: left = bins [0, lane], right = bins [lane + 1, 31]
    const WBox right = wbox_shfl_down(suf, 1);
    float cost = 3e38f;
    if (lane < 31 && pre.c > 0 && right.c > 0) cost = wbox_area(pre) * (float)pre.c + wbox_area(right) * (float)right.c;
    float m = cost;
    int arg = lane;
    for (int d = 16; d; d >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, d);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
      if (om < m || (om == m && oa < arg)) m = om, arg = oa;
    }
    if (m < best_cost) {
      best_cost = m, best_axis = ax, best_split = arg;
      best_l = wbox_shfl(pre, arg), best_r = wbox_shfl(right, arg);
    }
  }
  if (lane != 0) return;
  T.slot[node] = -1;  // slots are per level: this node's primitives must not bin again
  const int n = T.cnt[node];
  int cl, cr;
  WBox bl, br;
  if (best_axis >= 0) {
    cl = best_l.c, cr = best_r.c, bl = best_l, br = best_r;
  } else {
    // every centroid in one bin on every axis: keep it as one (big) leaf if the leaf code can hold it, else halve
    if (n <= 16) {
      T.l[node] = LEAF_MARK, T.r[node] = LEAF_MARK;
      return;
    }
    const float4 a = T.lo[node], b = T.hi[node];
    cl = n / 2, cr = n - cl;
    bl = br = WBox{a.x, a.y, a.z, b.x, b.y, b.z, 0};
    best_axis = -2, best_split = cl;
  }
  const int kid = atomicAdd(&counters[0], 2);
  T.l[node] = kid, T.r[node] = kid + 1, T.axis[node] = best_axis, T.split[node] = best_split, T.level[node] = level;
  const int kc[2] = {cl, cr};
  const WBox kb[2] = {bl, br};
  for (int k = 0; k < 2; ++k) {
    const int c = kid + k;
    T.lo[c] = make_float4(kb[k].lx, kb[k].ly, kb[k].lz, 0.f), T.hi[c] = make_float4(kb[k].hx, kb[k].hy, kb[k].hz, 0.f);
    T.cnt[c] = kc[k], T.first[c] = INT32_MAX, T.ticket[c] = 0u, T.level[c] = -1;
    if (kc[k] > GLEAF) {
      const int s = atomicAdd(&counters[1], 1);
      T.slot[c] = s, next_active[s] = c;
      T.l[c] = -1, T.r[c] = -1;
    } else {
      T.slot[c] = -1, T.l[c] = LEAF_MARK, T.r[c] = LEAF_MARK;
    }
  }
}

// (c) every primitive of a node split at this level moves to its child; the turn goes into the path key
__global__ void __launch_bounds__(256) k_sah_assign(const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n,
                                                    int *__restrict__ node_id, unsigned long long *__restrict__ key, SahTree T,
                                                    int level) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int node = node_id[i];
  if (T.level[node] != level || T.l[node] < 0) return;
  const int ax = T.axis[node];
  int side;
  if (ax >= 0) {
    const float4 a = blo[i], b = bhi[i], nl = T.lo[node], nh = T.hi[node];
    const float c = ax == 0 ? 0.5f * (a.x + b.x) : (ax == 1 ? 0.5f * (a.y + b.y) : 0.5f * (a.z + b.z));
    const float lo = ax == 0 ? nl.x : (ax == 1 ? nl.y : nl.z), hi = ax == 0 ? nh.x : (ax == 1 ? nh.y : nh.z);
    side = sah_bin(c, lo, hi) <= T.split[node] ? 0 : 1;
  } else {
    side = atomicAdd(&T.ticket[node], 1u) < (unsigned)T.split[node] ? 0 : 1;
  }
  node_id[i] = side ? T.r[node] : T.l[node];
  if (side) key[i] |= 1ull << (63 - level);
}

// after the sort by path key: where each leaf's range starts
__global__ void __launch_bounds__(256) k_sah_first(const unsigned *__restrict__ vals, int n, const int *__restrict__ node_id,
                                                   int *__restrict__ first) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  atomicMin(&first[node_id[vals[p]]], p);
}

}  // namespace gbvh
}  // namespace ptb

// gpu_bvh.cuh — tree construction ON the device for big triangle meshes.  Product code.
//
// Role in the reference: Shape_tree.create (path_tracer/src/shape_tree.ml:252-263) as ganesha calls it on the PLY
// mesh, where it is the visible cost ("build time", ganesha/bin/main.ml:190-194).  The host builder (bvh.cpp,
// binned SAH) needs 0.5 s for 10^6 triangles on 16 cores; this one is a linear BVH built in a few milliseconds:
//   1. primitive boxes + centroid bounds          (one pass, ordered-int atomics)
//   2. 63-bit Morton keys of the centroids        (21 bits per axis)
//   3. radix sort of (key, primitive)             (cub::DeviceRadixSort — library plumbing)
//   4. binary radix tree over the sorted keys     (Karras 2012: every internal node finds its range and split
//                                                  independently from the common-prefix lengths of neighbouring keys)
//   5. boxes bottom-up                            (second arrival at a node unions its children)
//   6. collapse into the 4-wide nodes the traversal kernel reads, level by level from the root: a node takes its two
//      binary children and twice opens the one with the largest surface area; subtrees of <= 4 primitives become
//      leaves (sorted order makes their primitives contiguous)
//   7. primitive tables gathered in sorted order
// Closest-hit results do not depend on the tree, so parity with the oracle is unchanged; a linear BVH is a little
// slower to traverse than the SAH tree, which pays off as long as the build dominates (it is chosen by primitive
// count, PTB_BUILDER=host|gpu overrides).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "device_types.cuh"

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
__device__ __forceinline__ float dec_f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// 1. per-triangle box (float, conservative: rounded outward from the double vertices) + centroid bounds
__global__ void __launch_bounds__(256) k_tri_boxes(const double *__restrict__ vx, const double *__restrict__ vy,
                                                   const double *__restrict__ vz, const int32_t *__restrict__ idx, int n,
                                                   float4 *__restrict__ blo, float4 *__restrict__ bhi, Bounds6 *cb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float c[3] = {0, 0, 0};
  const bool ok = i < n;
  if (ok) {
    const int a = idx[3 * i], b = idx[3 * i + 1], d = idx[3 * i + 2];
    const double x[3] = {vx[a], vx[b], vx[d]}, y[3] = {vy[a], vy[b], vy[d]}, z[3] = {vz[a], vz[b], vz[d]};
    const double lo[3] = {fmin(fmin(x[0], x[1]), x[2]), fmin(fmin(y[0], y[1]), y[2]), fmin(fmin(z[0], z[1]), z[2])};
    const double hi[3] = {fmax(fmax(x[0], x[1]), x[2]), fmax(fmax(y[0], y[1]), y[2]), fmax(fmax(z[0], z[1]), z[2])};
    blo[i] = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]), 0.f);
    bhi[i] = make_float4(__double2float_ru(hi[0]), __double2float_ru(hi[1]), __double2float_ru(hi[2]), 0.f);
    for (int k = 0; k < 3; ++k) c[k] = (float)(0.5 * (lo[k] + hi[k]));
  }
  // warp reduce, then one atomic per warp and bound
  for (int k = 0; k < 3; ++k) {
    float mn = ok ? c[k] : 3.0e38f, mx = ok ? c[k] : -3.0e38f;
    for (int o = 16; o; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&cb->lo[k], enc_f(mn));
      atomicMax(&cb->hi[k], enc_f(mx));
    }
  }
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
                                                        int *__restrict__ counters /* [0] wide nodes, [1] next frontier, [2] leaves */) {
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
      if (c < 0 || count[c] <= GLEAF) continue;
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
        if (count[c] <= GLEAF) {  // small subtree -> leaf over its (contiguous) sorted range
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
__global__ void __launch_bounds__(256) k_emit_tris(const unsigned *__restrict__ vals, int n, const double *__restrict__ vx,
                                                   const double *__restrict__ vy, const double *__restrict__ vz,
                                                   const int32_t *__restrict__ idx, const int32_t *__restrict__ tmat,
                                                   const double *__restrict__ tuv, const uint8_t *__restrict__ mat_kind,
                                                   Vec4<float> *__restrict__ tris, float *__restrict__ uv,
                                                   int32_t *__restrict__ tri_id, int32_t *__restrict__ tri_mat,
                                                   uint8_t *__restrict__ prim_kind) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = (int)vals[k];
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

}  // namespace gbvh
}  // namespace ptb

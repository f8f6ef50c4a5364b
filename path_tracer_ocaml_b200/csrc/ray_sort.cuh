// Ray reordering for ptb_intersect_batch on scenes that stay in global memory.
//
// Caller-supplied rays arrive in no useful order: the 32 rays of a warp then walk 32 unrelated root-to-leaf paths, every
// node and triangle record is a separate L1 wavefront and mostly an L2 round trip ("global-memory
// scenes").  The reference's FFI contract (sphere-intersect-rs/src/lib.rs:53-76) only fixes that result i belongs to ray
// i, so the rays are put into the traversal queue in SPATIAL order instead — a counting sort on
//   key = Morton code of the cell (64^3 over the scene's box) where the ray enters the scene
// (optionally | direction octant: PTB_SORT_OCT=1, measured slower) — and every queue entry carries the caller's index of its ray (A.w), which is where k_trace<MODE 1> writes the
// result.  Five small kernels (keys + histogram, two for the scan of the bins, index scatter, gather into the queue),
// bound by memory traffic and together well over an order of magnitude cheaper than the traversal they speed up.
#pragma once
#include <cstdint>

#include "device_types.cuh"

namespace ptb {

constexpr int SORT_GRID_BITS = 6;                               // cells per axis = 64
constexpr unsigned SORT_BINS = 1u << (3 * SORT_GRID_BITS + 3);  // x 8 direction octants = 2^21 bins at most
constexpr unsigned SORT_SCAN_BLOCK = 256, SORT_SCAN_PER_THREAD = 8;
constexpr unsigned SORT_SCAN_TILE = SORT_SCAN_BLOCK * SORT_SCAN_PER_THREAD;  // bins per block of the scan
constexpr unsigned SORT_SCAN_BLOCKS = SORT_BINS / SORT_SCAN_TILE;            // 1024
static_assert(SORT_BINS % SORT_SCAN_TILE == 0 && SORT_SCAN_BLOCKS <= 1024, "scan shape");
// (with oct_bits = 0 the key is the cell alone: 2^18 bins, 128 blocks of the scan)

// bytes of scratch the sort of m rays needs: keys, permutation, bin counters, per-block totals of the scan
static inline size_t ray_sort_scratch_bytes(size_t m) { return 2 * m * 4 + (size_t)SORT_BINS * 4 + SORT_SCAN_BLOCKS * 4; }

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 6 bits -> every third bit
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// the box of the whole tree = the union of the root's child boxes (unused slots are skipped)
__device__ __forceinline__ void root_box(const Node4<float> *nodes, float lo[3], float hi[3]) {
  for (int a = 0; a < 3; ++a) lo[a] = 3.0e38f, hi[a] = -3.0e38f;
  for (int k = 0; k < 4; ++k) {
    if (nodes[0].child[k] == INT32_MIN) continue;  // EMPTY_CHILD
    for (int a = 0; a < 3; ++a) lo[a] = fminf(lo[a], nodes[0].lo[a][k]), hi[a] = fmaxf(hi[a], nodes[0].hi[a][k]);
  }
}

__global__ void __launch_bounds__(256) k_ray_keys(const float *__restrict__ o, const float *__restrict__ d, long long n,
                                                  const Node4<float> *__restrict__ nodes, float tmin, int oct_bits,
                                                  unsigned *__restrict__ keys, unsigned *__restrict__ bins) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lo[3], hi[3];
  root_box(nodes, lo, hi);
  const float ox[3] = {o[3 * i], o[3 * i + 1], o[3 * i + 2]}, dx[3] = {d[3 * i], d[3 * i + 1], d[3 * i + 2]};
  // where the ray enters the box (its origin if that is inside, or if it misses the box: any key will do then)
  float tn = fmaxf(tmin, 0.0f);
  for (int a = 0; a < 3; ++a) {
    const float id = 1.0f / dx[a];
    const float t0 = (lo[a] - ox[a]) * id, t1 = (hi[a] - ox[a]) * id;
    tn = fmaxf(tn, fminf(t0, t1));  // NaN (0 * inf) drops out: fminf / fmaxf return the other operand
  }
  if (!(tn < 3.0e38f)) tn = 0.0f;
  unsigned key = 0;
  for (int a = 0; a < 3; ++a) {
    const float ext = fmaxf(hi[a] - lo[a], 1e-30f);
    const float u = (fmaf(tn, dx[a], ox[a]) - lo[a]) / ext * (float)(1 << SORT_GRID_BITS);
    const unsigned c = (unsigned)fminf(fmaxf(u, 0.0f), (float)((1 << SORT_GRID_BITS) - 1));  // (NaN -> 0)
    key |= spread3(c) << a;
  }
  if (oct_bits) key = (key << 3) | (dx[0] < 0.0f ? 1u : 0u) | (dx[1] < 0.0f ? 2u : 0u) | (dx[2] < 0.0f ? 4u : 0u);
  keys[i] = key;
  atomicAdd(&bins[key], 1u);
}

// exclusive scan of the bin counters, in place: (1) inside tiles of 2048 bins, tile totals aside; (2) every tile adds
// the sum of the totals before it
__global__ void __launch_bounds__(SORT_SCAN_BLOCK) k_sort_scan_tiles(unsigned *__restrict__ bins, unsigned *__restrict__ totals) {
  __shared__ unsigned warp_sum[SORT_SCAN_BLOCK / 32];
  const unsigned t = threadIdx.x, lane = t & 31u, w = t >> 5;
  uint4 *p = reinterpret_cast<uint4 *>(bins + (size_t)blockIdx.x * SORT_SCAN_TILE + t * SORT_SCAN_PER_THREAD);
  uint4 a = p[0], b = p[1];
  const unsigned mine = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  unsigned incl = mine;
  for (int s = 1; s < 32; s <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, s);
    if ((int)lane >= s) incl += v;
  }
  if (lane == 31u) warp_sum[w] = incl;
  __syncthreads();
  unsigned before = 0;
  for (unsigned k = 0; k < w; ++k) before += warp_sum[k];
  unsigned run = before + incl - mine;
  uint4 ea, eb;
  ea.x = run, run += a.x, ea.y = run, run += a.y, ea.z = run, run += a.z, ea.w = run, run += a.w;
  eb.x = run, run += b.x, eb.y = run, run += b.y, eb.z = run, run += b.z, eb.w = run, run += b.w;
  p[0] = ea, p[1] = eb;
  if (t == SORT_SCAN_BLOCK - 1) totals[blockIdx.x] = run;
}
__global__ void __launch_bounds__(SORT_SCAN_BLOCK) k_sort_scan_add(unsigned *__restrict__ bins, const unsigned *__restrict__ totals) {
  __shared__ unsigned part[SORT_SCAN_BLOCK / 32];
  __shared__ unsigned base_s;
  const unsigned t = threadIdx.x;
  unsigned s = 0;
  for (unsigned k = t; k < blockIdx.x; k += SORT_SCAN_BLOCK) s += totals[k];
  for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((t & 31u) == 0u) part[t >> 5] = s;
  __syncthreads();
  if (t == 0) {
    unsigned b = 0;
    for (unsigned k = 0; k < SORT_SCAN_BLOCK / 32; ++k) b += part[k];
    base_s = b;
  }
  __syncthreads();
  const unsigned base = base_s;
  if (base == 0u) return;
  uint4 *p = reinterpret_cast<uint4 *>(bins + (size_t)blockIdx.x * SORT_SCAN_TILE + t * SORT_SCAN_PER_THREAD);
  uint4 a = p[0], b = p[1];
  a.x += base, a.y += base, a.z += base, a.w += base, b.x += base, b.y += base, b.z += base, b.w += base;
  p[0] = a, p[1] = b;
}

// The scatter moves only the ray's INDEX (4 bytes into a table that stays in the L2); the rays themselves are then
// gathered slot by slot, so that the queue is written in full, consecutive sectors.  (Scattering the 32 bytes of a ray
// as two 16-byte halves of 32-byte sectors was measured at 190 us per 4 Mi rays: every one of them a read-modify-write.)
__global__ void __launch_bounds__(256) k_sort_scatter(long long n, const unsigned *__restrict__ keys, unsigned *__restrict__ bins,
                                                      unsigned *__restrict__ perm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  perm[atomicAdd(&bins[keys[i]], 1u)] = (unsigned)i;
}
// user rays (3 floats each) -> ray queue: slot s takes ray perm[s]; A.w = the caller's index of that ray
template <class R>
__global__ void __launch_bounds__(256) k_pack_rays_sorted(const float *__restrict__ o, const float *__restrict__ d, long long n,
                                                          const unsigned *__restrict__ perm, Queue<R> q) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const size_t i = perm[s];
  Vec4<R> *e = q.A((unsigned)s);
  e[0] = {R(o[3 * i]), R(o[3 * i + 1]), R(o[3 * i + 2]), i2r((int)i, R())};
  e[SEG] = {R(d[3 * i]), R(d[3 * i + 1]), R(d[3 * i + 2]), R(0)};
  if (s % SEG == 0) q.seg_count[s / SEG] = (int32_t)(n - s < SEG ? n - s : SEG);  // dense
}

}  // namespace ptb

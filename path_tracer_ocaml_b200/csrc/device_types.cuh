// device_types.cuh — device-side data layout of libptb200 (sm_100a).  Product code.
//
// Everything is templated on the arithmetic type R: float is the production path, double is the
// validation mode (PTB_FLAG_F64) that runs the identical pipeline in the reference's precision.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptb {

constexpr int MAX_BOUNCES = 64;
constexpr int MAX_DIMS = 2 + 2 * MAX_BOUNCES;
constexpr int NUM_MAT_KINDS = 3;  // Lambertian, Metal, Dielectric (material.ml:3-9)

// 128-bit aligned 4-vector; for double it is two 128-bit halves.
template <class R>
struct alignas(16) Vec4 {
  R x, y, z, w;
};
template <class R>
struct V3 {
  R x, y, z;
};

// 4-wide BVH node, SoA over the four children: lo[axis][child], hi[axis][child], child[4].
// float: 112 B = 7 x 16 B.  The stride is an ODD multiple of 16 B on purpose: lanes that visit
// different nodes then spread over all eight 16-byte bank groups of shared memory instead of
// colliding on one (a 128 B stride would be an 8-way conflict on every field).
template <class R>
struct alignas(16) Node4 {
  R lo[3][4];
  R hi[3][4];
  int32_t child[4];
};
static_assert(sizeof(Node4<float>) == 112, "Node4<float> must be 7 x 16 bytes");
static_assert(sizeof(Node4<double>) == 208, "Node4<double> must be 13 x 16 bytes");

// Scenes that stay in GLOBAL memory, float: records sized for 256-bit loads.  The traversal of such a scene is bound
// by the L1 data pipe — one wavefront per lane per load instruction, whatever its width, because every lane reads its
// own node — so a node is read with 3 x 32 B + 16 B instead of 7 x 16 B and a triangle with 32 B + 4 B instead of
// 3 x 16 B.  A node holds, per axis, the low planes of its four children and then the high ones; near and far are
// told apart by min / max of the two plane distances (no address selection by the direction's sign).  An unused
// child slot has NaN planes: its far distance is NaN and the (unordered) miss test rejects it.
struct alignas(128) NodeG {
  float pl[3][8];
  int32_t child[4];
  int32_t pad[4];
};
struct alignas(64) TriG {
  float v[16];  // v0, e1, e2 (9 floats); the rest is padding
};
// The same node with 8-bit child boxes, 64 B = two 256-bit loads, for caller-supplied (incoherent) rays, whose traversal
// is bound by the bytes each lane gathers: plane = origin + q * s per axis, q in [0, 255], s a power of two, the low
// planes rounded down and the high planes rounded up with a margin that covers the decode arithmetic (render.cu /
// k_make_g_layout).  scale15 = s * 2^15: the kernel turns a byte into the float 1 + q * 2^-15 with one PRMT.
// An unused child slot has q_lo = 255 > q_hi = 0 on every axis: an inverted box, which every ray misses.
struct alignas(64) NodeQ {
  float origin[3];
  float scale15[3];
  uint32_t qlo[3];  // one byte per child
  uint32_t qhi[3];
  int32_t child[4];
};
static_assert(sizeof(NodeG) == 128 && sizeof(TriG) == 64 && sizeof(NodeQ) == 64, "256-bit load records");

struct alignas(16) DMat {
  int32_t kind;
  int32_t tex;
  double index;
  // the kind of texture `tex` and, when it is a solid one, its colour: the float path shades without the
  // dependent load of the texture record
  int32_t tex_kind;
  float rgb[3];
};
static_assert(sizeof(DMat) == 32, "DMat is read as two 16-byte words");
template <class R>
struct DTex {
  int32_t kind, w, h, even, odd, pad;
  R rgb[3];
};

template <class R>
struct DScene {
  const Node4<R> *nodes;
  const Vec4<R> *spheres;  // (cx, cy, cz, r) per slot
  const Vec4<R> *tris;     // 3 per slot: (v0,_), (e1,_), (e2,_)
  const NodeG *nodes_g;    // float scenes in global memory: the same tree and triangles as 256-bit load records
  const TriG *tris_g;
  const NodeQ *nodes_q;    // ... and with quantised child boxes (ptb_intersect_batch)
  const Vec4<R> *spheres_g;  // (cx, cy, cz, r^2), padded by 8 records
  const int32_t *sphere_id, *tri_id;    // slot -> caller index
  const int32_t *sphere_mat, *tri_mat;  // slot -> material row
  const uint8_t *prim_kind;             // slot -> material kind, spheres then triangles, padded to 16 B
  const R *tri_uv;                      // 6 per slot
  const DMat *mats;
  const DTex<R> *texs;
  int32_t n_nodes, n_spheres, n_tris;
  int32_t scene_in_smem;  // 1: nodes+spheres+tris are staged in shared memory by every block
  int32_t stack_cap;      // traversal stack entries per thread
  int32_t bg_kind;
  R bg0[3], bg1[3], bgd[3];  // bgd = bg1 - bg0
  // extension (ptb_scene_set_light_quad / PTB_MAT_EMISSIVE): diffuse_plus_light = Mix (Diffuse, Quad_light)
  int32_t has_light, has_emissive;
  R light_o[3], light_u[3], light_v[3];
};

// Wavefront queue entry = three Vec4 (48 B in float):
//   ray queue : A = (origin, pixel)      B = (direction, R2 offset)   C = (attenuation, -)
//   hit queue : A = (hit point, pixel)   B = (direction, R2 offset)   C = (attenuation, prim)
// All queues are SEGMENTED: slot = segment * SEG + i, segment s holds seg_count[s] <= SEG valid entries.
// A producer warp owns one open segment per output queue and touches the queue's global segment
// counter once per SEG entries (instead of one contended return-value atomic per warp flush); a
// consumer claims whole segments.  Only the last segment a warp opened can be partially filled.
constexpr int SEG = 128;
// Memory layout: SEGMENT-INTERLEAVED.  One allocation; segment s occupies 3 * SEG consecutive Vec4 as
// [A x SEG][B x SEG][C x SEG].  A warp that works through a segment therefore touches ONE region of memory (one open
// DRAM page stream in, one out, instead of three each), the three fields of an entry are at constant distances
// (one address computation + immediates), and 32 consecutive entries of a field are 512 contiguous bytes — the unit
// k_shade moves with bulk copies (cp.async.bulk).
template <class R>
struct Queue {
  Vec4<R> *base;
  int32_t *seg_count;
  // e/SEG*3*SEG + e%SEG = e + 2*SEG*(e/SEG), in 64 bits: a 512 Mi-path batch has 1.6 G entries in the joint hit queue
  static __host__ __device__ __forceinline__ size_t idx(unsigned e) { return (size_t)e + (size_t)(e >> 7) * (2u * SEG); }
  __host__ __device__ __forceinline__ Vec4<R> *A(unsigned e) const { return base + e + (size_t)(e >> 7) * (2u * SEG); }
  __host__ __device__ __forceinline__ Vec4<R> *B(unsigned e) const { return A(e) + SEG; }
  __host__ __device__ __forceinline__ Vec4<R> *C(unsigned e) const { return A(e) + 2 * SEG; }
};
static_assert(SEG == 128, "Queue::idx hard-codes SEG = 128");

struct Ctl {
  unsigned int nseg_rays[MAX_BOUNCES + 1];         // segments of the ray queue entering bounce b
  unsigned int nseg_mat[MAX_BOUNCES][4];           // segments of the per-material hit queues at bounce b
  unsigned int n_rays[MAX_BOUNCES + 1];            // rays traced at bounce b of the current batch
  unsigned int cursor[MAX_BOUNCES + 1];            // segment-claim cursor per launch ([MAX_BOUNCES] = ad-hoc)
  unsigned long long rays_by_bounce[MAX_BOUNCES];  // accumulated over batches
  unsigned long long total_rays;
};

struct RenderConst {
  int32_t W, H, spp, max_bounces;
  int32_t npix;  // pixels owned by this rank
  int32_t pad;
  double llx, lly, vx, vy;  // Camera.t (camera.ml:50-53)
  double widthf, heightf;   // 1 // width, 1 // height (integrator.ml:92-93)
  double alpha[MAX_DIMS];   // Low_discrepancy_sequence alpha table (host glibc pow)
};

}  // namespace ptb

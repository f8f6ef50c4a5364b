// kernels.cuh — the wavefront pipeline of libptb200: hand-written CUDA for sm_100a.  Product code.
//
// Stages (BASELINE.json north_star):
//   camera_sample  camera rays from the Roberts R2 sequence, evaluated on device in float64 with un-fused
//              multiplies/adds so the sample stream is bit-identical to the reference (integrator.ml:98-105,
//              low_discrepancy_sequence.ml:33-36, camera.ml:93-102); called by the bounce-0 k_trace, which
//              generates its rays in registers, and by k_raygen (test / first-hit entry points only)
//   k_trace    closest-hit traversal of the 4-wide BVH (nodes + primitives staged in shared memory,
//              per-thread stack in bank-conflict-free shared memory), ray-sphere and ray-triangle
//              tests (sphere.ml:35-54 / lib.rs:102-178, triangle.ml:74-98); misses are shaded with
//              the background and accumulated; hits are appended to per-material queues with
//              __ballot_sync/__popc warp aggregation
//   k_shade    per-material scatter (material.ml:22-57, shader_space.ml, pdf.ml); each warp works
//              on one material's queue, survivors are compacted into the next ray queue
//   k_resolve  3x3 binomial reconstruction filter + gamma (film_tile.ml:23-38, integrator.ml:114-128,
//              152-154) over the per-pixel sample sums
// No tensor cores: no stage is a dense contraction.
#pragma once
#include <cfloat>

#ifndef PTB_GLOBAL_BLOCKS
#define PTB_GLOBAL_BLOCKS 4  // resident 256-thread blocks per SM of the global-memory traversal kernel: 64 registers
                             // per thread, no spills (3 blocks / 70 registers: -13 % on the C5 triangle soup; 5 / 48: spills)
#endif
#ifndef PTB_STACK_HOT
#define PTB_STACK_HOT 12  // shared-memory levels of the hybrid traversal stack (global-memory scenes)
#endif


#include "device_types.cuh"

namespace ptb {

// ---------------------------------------------------------------------------------------------
// scalar helpers, overloaded on float/double
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float r_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double r_sqrt(double x) { return sqrt(x); }
// float: hardware approximations (MUFU.RCP / MUFU.SQRT, <= 2 ulp) — well inside the float32 error of the
// quantities they feed (t, 1/d); double: exact IEEE, the validation mode mirrors the reference
__device__ __forceinline__ float r_rcp(float x) {  // one MUFU.RCP
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double r_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float r_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double r_div(double a, double b) { return a / b; }
__device__ __forceinline__ float r_sqrt_fast(float x) {  // one MUFU.SQRT (no subnormal rescaling)
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ double r_sqrt_fast(double x) { return sqrt(x); }
__device__ __forceinline__ float r_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double r_abs(double x) { return fabs(x); }
__device__ __forceinline__ float r_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double r_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float r_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double r_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double r_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float r_copysign(float a, float b) { return copysignf(a, b); }
__device__ __forceinline__ double r_copysign(double a, double b) { return copysign(a, b); }
__device__ __forceinline__ void r_sincos2pi(float v, float *s, float *c) { sincospif(2.0f * v, s, c); }
__device__ __forceinline__ void r_sincos2pi(double v, double *s, double *c) {
  sincos(v * 2.0 * 3.141592653589793, s, c);  // shader_space.ml:59-61, same expression
}
__device__ __forceinline__ float r_pow5(float x) {
  float x2 = x * x;
  return x2 * x2 * x;
}
__device__ __forceinline__ double r_pow5(double x) { return pow(x, 5.0); }  // material.ml:19,37 `** 5.0`
// 1/|v|: reference normalizes with 1/hypot(x, hypot(y, z)) (affine.ml:65-68)
__device__ __forceinline__ float r_inv_len(float x, float y, float z) {
  return rsqrtf(fmaf(x, x, fmaf(y, y, z * z)));
}
__device__ __forceinline__ double r_inv_len(double x, double y, double z) {
  return 1.0 / hypot(x, hypot(y, z));
}
__device__ __forceinline__ float r_inv_len4(float r, float x, float y, float z) {
  return rsqrtf(fmaf(r, r, fmaf(x, x, fmaf(y, y, z * z))));
}
__device__ __forceinline__ double r_inv_len4(double r, double x, double y, double z) {
  return 1.0 / hypot(hypot(r, x), hypot(y, z));  // quaternion.ml:11-15
}
template <class R>
struct Lim;
template <>
struct Lim<float> {
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  static __device__ __forceinline__ float tmax() { return FLT_MAX; }
  static __device__ __forceinline__ float far_scale() { return 1.0000004f; }
  static __device__ __forceinline__ float tiny() { return 1e-30f; }
  static __device__ __forceinline__ float below_one() { return 0.99999994f; }
};
template <>
struct Lim<double> {
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  static __device__ __forceinline__ double tmax() { return DBL_MAX; }  // Float.max_finite_value (main.ml:274)
  static __device__ __forceinline__ double far_scale() { return 1.0 + 1e-15; }
  static __device__ __forceinline__ double tiny() { return 1e-300; }
  static __device__ __forceinline__ double below_one() { return 0.99999999999999989; }
};
__device__ __forceinline__ float i2r(int v, float) { return __int_as_float(v); }
__device__ __forceinline__ double i2r(int v, double) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ int r2i(float v) { return __float_as_int(v); }
__device__ __forceinline__ int r2i(double v) { return (int)__double_as_longlong(v); }

template <class R>
__device__ __forceinline__ V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R>
__device__ __forceinline__ V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R>
__device__ __forceinline__ V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <class R>
__device__ __forceinline__ V3<R> neg(V3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R>
__device__ __forceinline__ R dot(V3<R> a, V3<R> b) { return r_fma(a.x, b.x, r_fma(a.y, b.y, a.z * b.z)); }
template <class R>
__device__ __forceinline__ V3<R> cross(V3<R> p, V3<R> q) {
  return {r_fma(p.y, q.z, -(p.z * q.y)), r_fma(p.z, q.x, -(p.x * q.z)), r_fma(p.x, q.y, -(p.y * q.x))};
}
template <class R>
__device__ __forceinline__ V3<R> normalize(V3<R> v) { return v * r_inv_len(v.x, v.y, v.z); }

// ---------------------------------------------------------------------------------------------
// Roberts R2 sample (low_discrepancy_sequence.ml:19-20,33-36): frac(0.5 + alpha * float(1+offset)).
// float64, explicitly un-fused so no FMA contraction can change a bit.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double r2_sample(double alpha, int offset) {
  double x = __dadd_rn(0.5, __dmul_rn(alpha, (double)(1 + offset)));
  return __dsub_rn(x, trunc(x));
}

// ---------------------------------------------------------------------------------------------
// quaternion shading frame (shader_space.ml:11-32, quaternion.ml:25-42)
// ---------------------------------------------------------------------------------------------
template <class R>
struct Quat {
  R r;
  V3<R> v;
};
template <class R>
__device__ __forceinline__ Quat<R> quat_mul(Quat<R> a, Quat<R> b) {
  R r = a.r * b.r - dot(a.v, b.v);
  V3<R> v = (cross(a.v, b.v) + b.v * a.r) + a.v * b.r;
  return {r, v};
}
template <class R>
__device__ __forceinline__ V3<R> quat_transform(Quat<R> t, V3<R> v) {
  Quat<R> tc = {t.r, neg(t.v)};
  return quat_mul(quat_mul(t, Quat<R>{R(0), v}), tc).v;
}
// float path: the shading-frame quaternions always have v.z == 0 (frame_from_normal below), and q v q* for a unit
// q = (r; a) is v + 2 (r (a x v) + a x (a x v)): 14 instructions instead of the two Hamilton products (~56) the
// reference spells out (quaternion.ml:25-42).  Same rotation, float rounding only; the float64 validation mode
// keeps the reference's formula.
__device__ __forceinline__ V3<float> quat_transform(Quat<float> t, V3<float> v) {
  const float ax = t.v.x, ay = t.v.y;
  const float cx = ay * v.z, cy = -ax * v.z, cz = fmaf(ax, v.y, -(ay * v.x));   // c = a x v
  const float dx = ay * cz, dy = -ax * cz, dz = fmaf(ax, cy, -(ay * cx));        // a x c
  return {fmaf(2.0f, fmaf(t.r, cx, dx), v.x), fmaf(2.0f, fmaf(t.r, cy, dy), v.y), fmaf(2.0f, fmaf(t.r, cz, dz), v.z)};
}
template <class R>
__device__ __forceinline__ Quat<R> frame_from_normal(V3<R> n) {
  const R eps = sizeof(R) == 4 ? R(1e-7) : R(1e-9);  // shader_space.ml:9 uses 1e-9 (float64)
  if (n.z > R(1) - eps) return {R(1), {R(0), R(0), R(0)}};
  if (n.z < eps - R(1)) return {R(0), {R(0), R(1), R(0)}};
  R r = R(1) + n.z, x = n.y, y = -n.x;
  R s = r_inv_len4(r, x, y, R(0));
  return {r * s, {x * s, y * s, R(0)}};
}

// ---------------------------------------------------------------------------------------------
// primitive tests
// ---------------------------------------------------------------------------------------------
// Ray-sphere in the numerically robust form the reference uses (sphere.ml:35-54; lib.rs:130-166):
// discriminant from the perpendicular offset, q = b' + sign(b')*sqrt(a*disc), t = c>0 ? c/q : q/a.
// `s.w` holds r^2 here (SceneRef::sphere squares it or reads the pre-squared shared-memory copy).
// float: straight-line.  A negative discriminant makes sqrt.approx return NaN, NaN propagates into t and
// every comparison with NaN is false, so the miss needs no branch (the Rust kernel's NaN masking,
// lib.rs:160-166, is the same idea).
// UNIT: the ray direction has unit length (every ray of the render pipeline: Camera.ray normalizes, world_ray rotates
// unit vectors), so a = d.d = 1 drops out: three multiplies and two live registers less per test.
// TMIN0: t_min is 0, so "0 <= t <= tbest" is ONE unsigned comparison of the bit patterns (tbest is a non-negative
// number: every negative or NaN t has a larger pattern than any such tbest).
template <bool UNIT, bool TMIN0 = false>
__device__ __forceinline__ bool sphere_test_f(Vec4<float> s, V3<float> o, V3<float> d, float a, float inv_a,
                                              float tmin, float &tbest) {
  const V3<float> f = {s.x - o.x, s.y - o.y, s.z - o.z};
  const float bp = dot(f, d);
  const float boa = UNIT ? bp : bp * inv_a;
  const V3<float> w = {fmaf(d.x, boa, -f.x), fmaf(d.y, boa, -f.y), fmaf(d.z, boa, -f.z)};
  const float disc = fmaf(-w.x, w.x, fmaf(-w.y, w.y, fmaf(-w.z, w.z, s.w)));
  const float q = bp + copysignf(r_sqrt_fast(UNIT ? disc : a * disc), bp);
  const float c = fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, -s.w)));
  const float t = (c > 0.0f) ? c * r_rcp(q) : (UNIT ? q : q * inv_a);
  // `<=`: a later equal t wins (lib.rs:171-176)
  const bool ok = TMIN0 ? (__float_as_uint(t) <= __float_as_uint(tbest)) : ((t >= tmin) && (t <= tbest));
  tbest = ok ? t : tbest;
  return ok;
}
__device__ __forceinline__ void sphere_test(Vec4<float> s, V3<float> o, V3<float> d, float a, float inv_a,
                                            float tmin, float &tbest, int &best, int id) {
  const bool ok = sphere_test_f<false>(s, o, d, a, inv_a, tmin, tbest);
  best = ok ? id : best;
}
template <class R>
__device__ __forceinline__ void sphere_test(Vec4<R> s, V3<R> o, V3<R> d, R a, R inv_a, R tmin, R &tbest,
                                            int &best, int id) {
  V3<R> f = {s.x - o.x, s.y - o.y, s.z - o.z};
  R r2 = s.w;
  R bp = dot(f, d);
  R boa = bp * inv_a;
  V3<R> w = {r_fma(d.x, boa, -f.x), r_fma(d.y, boa, -f.y), r_fma(d.z, boa, -f.z)};
  R disc = r2 - dot(w, w);
  if (disc >= R(0)) {
    R q = bp + r_copysign(r_sqrt_fast(a * disc), bp);
    R c = dot(f, f) - r2;
    R t = (c > R(0)) ? r_div(c, q) : q * inv_a;
    if (t >= tmin && t <= tbest) {  // `<=`: a later equal t wins (lib.rs:171-176, shape_tree.ml:303-309)
      tbest = t;
      best = id;
    }
  }
}
// Moller-Trumbore, two-sided, |det| < 1e-6 rejected (triangle.ml:74-98)
template <class R>
__device__ __forceinline__ void tri_test(Vec4<R> v0, Vec4<R> e1v, Vec4<R> e2v, V3<R> o, V3<R> d, R tmin,
                                         R &tbest, int &best, int id) {
  V3<R> e1 = {e1v.x, e1v.y, e1v.z}, e2 = {e2v.x, e2v.y, e2v.z};
  V3<R> pvec = cross(d, e2);
  R det = dot(e1, pvec);
  if (r_abs(det) < R(1e-6)) return;
  R inv = r_rcp(det);  // float: one MUFU.RCP (|det| >= 1e-6, far from the flush-to-zero range); double: exact
  V3<R> tvec = {o.x - v0.x, o.y - v0.y, o.z - v0.z};
  R u = inv * dot(tvec, pvec);
  V3<R> qvec = cross(tvec, e1);
  R v = inv * dot(d, qvec);
  if (u >= R(0) && u <= R(1) && v >= R(0) && u + v <= R(1)) {
    R t = inv * dot(e2, qvec);
    if (t >= tmin && t <= tbest) {
      tbest = t;
      best = id;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// memory access helpers: explicit ld.shared / st.shared (32-bit shared addresses) so the hot loop
// never issues a generic load, and __ldg for the scenes that do not fit shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ Vec4<float> lds_vec4(unsigned addr, float) {
  Vec4<float> v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ Vec4<double> lds_vec4(unsigned addr, double) {
  Vec4<double> v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(addr + 16u));
  return v;
}
__device__ __forceinline__ void sts_vec4(unsigned addr, Vec4<float> v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void sts_vec4(unsigned addr, Vec4<double> v) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(v.x), "d"(v.y));
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr + 16u), "d"(v.z), "d"(v.w));
}
__device__ __forceinline__ int4 lds_int4(unsigned addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_i32(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_u8(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return (int)v;
}
__device__ __forceinline__ void sts_i32(unsigned addr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ float lds_r(unsigned addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_r(unsigned addr, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_r(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }
__device__ __forceinline__ void sts_r(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v)); }
// Packed FP32 (sm_100: one FFMA2 issues two FMAs on a 64-bit register pair).  The traversal kernel is bound by
// instruction issue, not by the FMA pipe, so halving the slab test's FMA instruction count pays even though
// the arithmetic rate is the same.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// asynchronous global -> shared copies (LDGSTS), used to keep the next work item of k_shade in flight
__device__ __forceinline__ void cp_async16(unsigned saddr, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_scalar(unsigned saddr, const float *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_scalar(unsigned saddr, const double *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
template <int OFF>
__device__ __forceinline__ void cp_async16_off(unsigned saddr, const void *g) {  // [g + OFF]: the offset is an immediate
  asm volatile("cp.async.cg.shared.global [%0], [%1+%2], 16;" ::"r"(saddr), "l"(g), "n"(OFF) : "memory");
}
// FIELD: which of the three fields of a segment-interleaved queue entry (0 = A, 1 = B, 2 = C), `g` = the entry's A
template <class R, int FIELD = 0>
__device__ __forceinline__ void cp_async_vec4(unsigned saddr, const Vec4<R> *g) {
  cp_async16_off<FIELD * SEG * (int)sizeof(Vec4<R>)>(saddr, g);
  if constexpr (sizeof(Vec4<R>) == 32) cp_async16_off<FIELD * SEG * (int)sizeof(Vec4<R>) + 16>(saddr + 16u, g);
}
// Bulk asynchronous copies (sm_90+/sm_100: the TMA unit's 1-D form, SASS UBLKCP) and the mbarrier that tracks them.
// One lane moves a whole contiguous run — 512 B of one field of 32 queue entries — with one instruction, instead of 32
// lanes moving 16 B each.
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(mbar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned sdst, const void *gsrc, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, unsigned ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// traversal-stack entry = (child ref, t_near): 8 B in float (one 64-bit shared access), 16 B in double
__device__ __forceinline__ void stk_store(unsigned addr, int ref, float t) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ref), "r"(__float_as_int(t)));
}
__device__ __forceinline__ void stk_store(unsigned addr, int ref, double t) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(ref));
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr + 8u), "d"(t));
}
__device__ __forceinline__ void stk_load(unsigned addr, int &ref, float &t) {
  int tb;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ref), "=r"(tb) : "r"(addr));
  t = __int_as_float(tb);
}
__device__ __forceinline__ void stk_load(unsigned addr, int &ref, double &t) {
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ref) : "r"(addr));
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(addr + 8u));
}
// Where the traversal stack lives.  Shared memory (stack[level][thread], see Lane) when the scene is staged in
// shared memory too; for big scenes, which run from global memory and want every warp the register file allows,
// the thread's LOCAL memory (the hardware interleaves it per thread, L1 keeps the few live levels).
// `sp` is a shared-window byte address in the first case and an entry index in the second.
constexpr int LOCAL_STACK_CAP = 97;
template <class R, bool LOCAL>
struct Stack;
template <class R>
struct Stack<R, false> {
  unsigned stride;
  __device__ __forceinline__ void store(unsigned sp, int ref, R t) { stk_store(sp, ref, t); }
  __device__ __forceinline__ void load(unsigned sp, int &ref, R &t) const { stk_load(sp, ref, t); }
};
template <class R>
struct Stack<R, true> {
  // hybrid: the first HOT levels (where almost all traffic is) sit in shared memory as stack[level][thread], the
  // rarely reached deep levels in local memory — 64 B of shared memory per thread instead of the tree's worst case
  static constexpr unsigned stride = 1u;
  static constexpr unsigned HOT = PTB_STACK_HOT;
  unsigned s_base, s_stride;  // shared-window address of this thread's level 0, bytes between levels
  int ref_[LOCAL_STACK_CAP - HOT];
  R t_[LOCAL_STACK_CAP - HOT];
  __device__ __forceinline__ void store(unsigned sp, int ref, R t) {
    if (sp < HOT) stk_store(s_base + sp * s_stride, ref, t);
    else ref_[sp - HOT] = ref, t_[sp - HOT] = t;
  }
  __device__ __forceinline__ void load(unsigned sp, int &ref, R &t) const {
    if (sp < HOT) stk_load(s_base + sp * s_stride, ref, t);
    else ref = ref_[sp - HOT], t = t_[sp - HOT];
  }
};
__device__ __forceinline__ Vec4<float> ldg_vec4(const Vec4<float> *p) {
  float4 v = __ldg(reinterpret_cast<const float4 *>(p));
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Vec4<double> ldg_vec4(const Vec4<double> *p) {
  double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return {a.x, a.y, b.x, b.y};
}

// 256-bit read-only global load (sm_100: LDG.E.256), 32-byte aligned
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ldg256(const void *p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}

// where the scene lives for this launch: shared memory (staged by the block) or global memory
template <class R, bool SMEM>
struct SceneRef {
  unsigned s_nodes, s_spheres, s_tris, s_kinds;  // shared-window byte addresses (SMEM)
  const char *g_nodes;
  const Vec4<R> *g_spheres, *g_tris;
  const char *g_nodes_g, *g_tris_g, *g_sph_g;  // float, global memory: NodeG / TriG records, spheres as (c, r^2)
  const char *g_nodes_q;                       // NodeQ records
  const uint8_t *g_kinds;
  static constexpr unsigned ROW = 4u * (unsigned)sizeof(R);  // one SoA row of a node: 4 children of one plane
  // material kind of a primitive; `best` = slot | type << 30; table = spheres then triangles
  __device__ __forceinline__ int kind(int best, int n_spheres) const {
    const int idx = (best & 0x3FFFFFFF) + (((best >> 30) & 1) ? n_spheres : 0);
    if (SMEM) return lds_u8(s_kinds + (unsigned)idx);
    return (int)__ldg(g_kinds + idx);
  }
  // row at byte offset `off` of node `node`.  SMEM: `node` IS the node's shared-window address — the block rewrites
  // the inner-node references of the staged tree once, so a node visit starts without an address multiply
  __device__ __forceinline__ Vec4<R> nrow(int node, unsigned off) const {
    if (SMEM) return lds_vec4((unsigned)node + off, R());
    return ldg_vec4(reinterpret_cast<const Vec4<R> *>(g_nodes + (size_t)node * sizeof(Node4<R>) + off));
  }
  __device__ __forceinline__ int4 children(int node) const {
    if (SMEM) return lds_int4((unsigned)node + 6u * ROW);
    return __ldg(reinterpret_cast<const int4 *>(g_nodes + (size_t)node * sizeof(Node4<R>) + 6 * ROW));
  }
  // (cx, cy, cz, r^2): the shared-memory copy is squared once when the block stages it
  __device__ __forceinline__ Vec4<R> sphere(int i) const {
    if (SMEM) return lds_vec4(s_spheres + (unsigned)i * (unsigned)sizeof(Vec4<R>), R());
    Vec4<R> v = ldg_vec4(g_spheres + i);
    v.w *= v.w;
    return v;
  }
  __device__ __forceinline__ Vec4<R> tri(int i, int k) const {
    if (SMEM) return lds_vec4(s_tris + (unsigned)(3 * i + k) * (unsigned)sizeof(Vec4<R>), R());
    return ldg_vec4(g_tris + 3 * (size_t)i + k);
  }
};

// ---------------------------------------------------------------------------------------------
// BVH4 traversal state of one lane.  The per-thread stack lives in shared memory as
// stack[level][thread] with one (ref, t_near) entry per level: consecutive lanes are consecutive
// entries at every level, so pushes and pops never bank-conflict whatever the per-lane depth.
// Entry 0 of every thread is a permanent sentinel (TRAV_DONE, -inf) written at kernel start: the pop
// loop needs no empty check, it simply pops the sentinel when the ray is finished.
// ---------------------------------------------------------------------------------------------
// `cur` doubles as the lane's state (no separate flags to juggle in the hot loop):
//   cur >= 0: inner node | TRAV_POP < cur < 0: leaf code | TRAV_POP: take the next stack entry
//   TRAV_DONE: ray finished, result not flushed yet (the stack's bottom entry) | TRAV_IDLE: no ray
// (leaf codes are ~(first | (count-1) << 26 | type << 30) >= 0xB0000000 as unsigned, far above the sentinels)
constexpr int TRAV_IDLE = INT32_MIN;
constexpr int TRAV_DONE = INT32_MIN + 1;
constexpr int TRAV_POP = INT32_MIN + 2;

// compare-exchange of (t, child) pairs: the distances go through min / max (they do not wait for the comparison),
// the child references through two selects
#define PTB_CSWAP(ta, ca, tb, cb)   \
  {                                 \
    const bool sw_ = tb < ta;       \
    const R lo_ = r_min(ta, tb);    \
    const R hi_ = r_max(ta, tb);    \
    const int cc_ = sw_ ? ca : cb;  \
    ca = sw_ ? cb : ca;             \
    ta = lo_;                       \
    tb = hi_;                       \
    cb = cc_;                       \
  }

template <class R>
struct Lane {
  V3<R> o, d, idir, oid;
  R a, inv_a, tmin, tbest;
  int best, cur;
  unsigned sp;             // shared-window byte address of the next free stack entry
  unsigned onx, ony, onz;  // byte offsets of the near x / y / z plane rows inside a node (by direction sign)
  // float path: (1/d, 1/d) and (-o/d, -o/d) per axis as packed pairs for FFMA2 (idir/oid above are then unused)
  f32x2_t ix2, iy2, iz2, nox2, noy2, noz2;
};

// UNIT: unit-length direction, a = 1 is not stored (float render pipeline).  root: the root's `cur` value (node
// index 0, or its shared-memory address when the staged tree holds addresses)
// QM: onx / ony / onz hold all-ones masks for negative direction components instead of row offsets (NodeQ traversal)
template <class R, bool UNIT = false, bool QM = false>
__device__ __forceinline__ void lane_init(Lane<R> &L, V3<R> o, V3<R> d, R tmin, R tmax, unsigned sp0, int root = 0) {
  constexpr unsigned ROW = 4u * (unsigned)sizeof(R);
  if constexpr (sizeof(R) == 4 && UNIT) {
    // Render pipeline, float: plain reciprocals.  A zero (or denormal) direction component gives +-inf here and
    // inf - inf = NaN in the slab arithmetic; min/max return their non-NaN operand, so such an axis simply drops out
    // of the slab test — a conservative answer (the ray runs parallel to those planes; camera and scattered directions
    // practically never have an exact zero), reached without the three compare-and-patch sequences a guarded
    // reciprocal costs per refilled ray.  The near / far rows are chosen by the SIGN BIT, so that -0.0 (whose
    // reciprocal is -inf) reads its planes in the order its reciprocal implies.  Caller-supplied rays
    // (intersect_batch), where axis-parallel directions are common, keep the guarded form below.
    L.idir = {r_rcp(d.x), r_rcp(d.y), r_rcp(d.z)};
    if constexpr (QM) {
      L.onx = (unsigned)(r2i(d.x) >> 31), L.ony = (unsigned)(r2i(d.y) >> 31), L.onz = (unsigned)(r2i(d.z) >> 31);
    } else {
      L.onx = r2i(d.x) < 0 ? 3u * ROW : 0u;  // rows: lo.x lo.y lo.z hi.x hi.y hi.z
      L.ony = r2i(d.y) < 0 ? 4u * ROW : ROW;
      L.onz = r2i(d.z) < 0 ? 5u * ROW : 2u * ROW;
    }
  } else {
    auto safe_rcp = [](R x) {
      return (r_abs(x) < Lim<R>::tiny()) ? r_copysign(R(1) / Lim<R>::tiny(), x) : r_rcp(x);
    };
    L.idir = {safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)};
    if constexpr (QM) {
      L.onx = d.x >= R(0) ? 0u : 0xffffffffu;
      L.ony = d.y >= R(0) ? 0u : 0xffffffffu;
      L.onz = d.z >= R(0) ? 0u : 0xffffffffu;
    } else {
      L.onx = d.x >= R(0) ? 0u : 3u * ROW;  // rows: lo.x lo.y lo.z hi.x hi.y hi.z
      L.ony = d.y >= R(0) ? ROW : 4u * ROW;
      L.onz = d.z >= R(0) ? 2u * ROW : 5u * ROW;
    }
  }
  L.oid = {o.x * L.idir.x, o.y * L.idir.y, o.z * L.idir.z};
  if constexpr (sizeof(R) == 4) {
    L.ix2 = pack2(L.idir.x, L.idir.x), L.iy2 = pack2(L.idir.y, L.idir.y), L.iz2 = pack2(L.idir.z, L.idir.z);
    L.nox2 = pack2(-L.oid.x, -L.oid.x), L.noy2 = pack2(-L.oid.y, -L.oid.y), L.noz2 = pack2(-L.oid.z, -L.oid.z);
  }
  L.o = o, L.d = d;
  if constexpr (!UNIT) {
    L.a = dot(d, d);
    L.inv_a = r_rcp(L.a);
  }
  L.tmin = tmin, L.tbest = tmax;
  L.best = -1, L.cur = root;
  L.sp = sp0;
}

// One traversal iteration of a warp = node phase, leaf phase, pop phase; each lane takes part in the phases
// its state calls for (see k_trace for how the warp decides when the leaf phase runs).  The ray is finished
// when cur == TRAV_DONE.
// TMIN0: t_min is the constant 0 (render pipeline).  FULLSORT: hit children are pushed strictly far-to-near
// (5-exchange network); otherwise only the nearest child is singled out (3 exchanges) and the others are
// pushed as they come — every pushed child is still visited unless its t_near exceeds the best hit, so the
// closest hit is unchanged (measured: +2 % node visits, -10 instructions per node).  CHECK: stack overflow
// guard (only when the tree's worst case exceeds the capacity).
template <class R, bool FULLSORT, bool CHECK, class STACK>
__device__ __forceinline__ void node_finish(Lane<R> &L, STACK &K, unsigned sp_limit, R tnx, R tny, R tnz, R tnw, int4 ch);

template <class R, bool SMEM, bool TMIN0, bool FULLSORT, bool CHECK>
__device__ __forceinline__ void node_phase(Lane<R> &L, const SceneRef<R, SMEM> &S, Stack<R, !SMEM> &K, unsigned sp_limit) {
  constexpr unsigned ROW = 4u * (unsigned)sizeof(R);
  const R INF = Lim<R>::inf();
  const R tmin = TMIN0 ? R(0) : L.tmin;
  // float boxes are padded outward on the host (render.cu box_lo/box_hi), which covers the rounding of the
  // plane distances; the unpadded double boxes get a relative slack on the far side instead
  R tnx, tny, tnz, tnw;
  int4 ch;
  if constexpr (sizeof(R) == 4) {
    const Vec4<R> bnx = S.nrow(L.cur, L.onx), bny = S.nrow(L.cur, L.ony), bnz = S.nrow(L.cur, L.onz);
    const Vec4<R> bfx = S.nrow(L.cur, 3u * ROW - L.onx), bfy = S.nrow(L.cur, 5u * ROW - L.ony),
                  bfz = S.nrow(L.cur, 7u * ROW - L.onz);
    ch = S.children(L.cur);
    // 12 FFMA2 instead of 24 FFMA: children (x,y) and (z,w) of a row are the register pairs of its LDS.128
#define PTB_ROW(row, i2, no2, a, b, c, d)                            \
  float a, b, c, d;                                                  \
  unpack2(ffma2(pack2(row.x, row.y), L.i2, L.no2), a, b);            \
  unpack2(ffma2(pack2(row.z, row.w), L.i2, L.no2), c, d);
    PTB_ROW(bnx, ix2, nox2, nx0, nx1, nx2, nx3)
    PTB_ROW(bny, iy2, noy2, ny0, ny1, ny2, ny3)
    PTB_ROW(bnz, iz2, noz2, nz0, nz1, nz2, nz3)
    PTB_ROW(bfx, ix2, nox2, fx0, fx1, fx2, fx3)
    PTB_ROW(bfy, iy2, noy2, fy0, fy1, fy2, fy3)
    PTB_ROW(bfz, iz2, noz2, fz0, fz1, fz2, fz3)
#undef PTB_ROW
  // (a miss becomes +inf by ADDING +inf under the miss predicate: a predicated FADD runs on the FMA pipe, which has
  // room, where the select it replaces would take a slot of the ALU pipe, which limits this kernel)
#define PTB_SLAB(k, i)                                                             \
  tn##k = fmaxf(fmaxf(nx##i, ny##i), fmaxf(nz##i, tmin));                          \
  {                                                                                \
    const float tf_ = fminf(fminf(fx##i, fy##i), fz##i);                           \
    asm("{\n.reg .pred p;\nsetp.gtu.f32 p, %0, %1;\n@p add.f32 %0, %0, 0f7F800000;\n}" : "+f"(tn##k) : "f"(tf_)); \
  }
    PTB_SLAB(x, 0)
    PTB_SLAB(y, 1)
    PTB_SLAB(z, 2)
    PTB_SLAB(w, 3)
#undef PTB_SLAB
  } else {
    const Vec4<R> bnx = S.nrow(L.cur, L.onx), bny = S.nrow(L.cur, L.ony), bnz = S.nrow(L.cur, L.onz);
    const Vec4<R> bfx = S.nrow(L.cur, 3u * ROW - L.onx), bfy = S.nrow(L.cur, 5u * ROW - L.ony),
                  bfz = S.nrow(L.cur, 7u * ROW - L.onz);
    ch = S.children(L.cur);
#define PTB_SLAB(k)                                                                                       \
  tn##k = r_max(r_max(r_fma(bnx.k, L.idir.x, -L.oid.x), r_fma(bny.k, L.idir.y, -L.oid.y)),                \
                r_max(r_fma(bnz.k, L.idir.z, -L.oid.z), tmin));                                           \
  {                                                                                                       \
    R tf_ = r_min(r_min(r_fma(bfx.k, L.idir.x, -L.oid.x), r_fma(bfy.k, L.idir.y, -L.oid.y)),              \
                  r_min(r_fma(bfz.k, L.idir.z, -L.oid.z), L.tbest));                                      \
    tf_ *= Lim<R>::far_scale();                                                                           \
    tn##k = (tn##k <= tf_) ? tn##k : INF;                                                                 \
  }
    PTB_SLAB(x)
    PTB_SLAB(y)
    PTB_SLAB(z)
    PTB_SLAB(w)
#undef PTB_SLAB
  }
  node_finish<R, FULLSORT, CHECK>(L, K, sp_limit, tnx, tny, tnz, tnw, ch);
}

// second half of a node visit: order the children that were hit, follow the nearest, push the others
template <class R, bool FULLSORT, bool CHECK, class STACK>
__device__ __forceinline__ void node_finish(Lane<R> &L, STACK &K, unsigned sp_limit, R tnx, R tny, R tnz, R tnw, int4 ch) {
  const R INF = Lim<R>::inf();
  int c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
  R t0 = tnx, t1 = tny, t2 = tnz, t3 = tnw;
  PTB_CSWAP(t0, c0, t1, c1)
  PTB_CSWAP(t2, c2, t3, c3)
  PTB_CSWAP(t0, c0, t2, c2)
  if (FULLSORT) {
    PTB_CSWAP(t1, c1, t3, c3)
    PTB_CSWAP(t1, c1, t2, c2)
  }
  // float path: the boxes are not clipped against the best hit (4 min instructions less per visit); the nearest child
  // is followed only if it starts before the best hit — the others are culled when they are popped
  L.cur = (sizeof(R) == 4 ? (t0 <= L.tbest) : (t0 < INF)) ? c0 : TRAV_POP;
  // (t0 == INF implies t1..t3 == INF: nothing is pushed)
  // float path: a child is pushed only if it starts before the best hit — the same comparison instruction as the
  // test against +inf, and it spares the store and the pop of every child the unclipped slab test let through
  const R push_below = sizeof(R) == 4 ? L.tbest : INF;
#define PTB_PUSHABLE(t) (sizeof(R) == 4 ? (t) <= push_below : (t) < push_below)
  if (PTB_PUSHABLE(t3) && (!CHECK || L.sp < sp_limit)) {
    K.store(L.sp, c3, t3);
    L.sp += K.stride;
  }
  if (PTB_PUSHABLE(t2) && (!CHECK || L.sp < sp_limit)) {
    K.store(L.sp, c2, t2);
    L.sp += K.stride;
  }
  if (PTB_PUSHABLE(t1) && (!CHECK || L.sp < sp_limit)) {
    K.store(L.sp, c1, t1);
    L.sp += K.stride;
  }
#undef PTB_PUSHABLE
}

// Sphere leaves of a tree staged in shared memory (float) are DIRECT: the block rewrites their child references to
// -(shared address of the first record | count - 1), so a leaf visit starts with two logic instructions instead of
// the unpacking of (first, count, type) and two address computations.  Triangle leaves keep the packed code.
constexpr int DIRECT_LEAF_MIN = -(1 << 24);  // direct codes are > this (shared addresses are < 2^18), packed ones far below
template <class R, bool SMEM, bool TMIN0, bool UNIT>
__device__ __forceinline__ void leaf_phase(Lane<R> &L, const SceneRef<R, SMEM> &S) {
  const R tmin = TMIN0 ? R(0) : L.tmin;
  if constexpr (SMEM && sizeof(R) == 4) {
    if (L.cur > DIRECT_LEAF_MIN) {
      // walk the leaf's records by shared-memory address; the hit is remembered as an address and turned into a
      // slot once per leaf (3 instructions of loop control per sphere instead of 6)
      const unsigned x = (unsigned)(-L.cur);
      unsigned addr = x & ~15u, hit = 0u;
      const unsigned end = addr + 16u + (x & 15u) * 16u;
#pragma unroll 1
      do {
        const bool ok = sphere_test_f<UNIT, TMIN0>(lds_vec4(addr, float()), L.o, L.d, UNIT ? 1.0f : L.a, UNIT ? 1.0f : L.inv_a, tmin, L.tbest);
        hit = ok ? addr : hit;
        addr += 16u;
      } while (addr < end);
      if (hit) L.best = (int)((hit - S.s_spheres) >> 4);
      L.cur = TRAV_POP;
      return;
    }
  }
  const unsigned code = ~(unsigned)L.cur;
  const int first = (int)(code & 0x3FFFFFFu);
  const int cnt = (int)((code >> 26) & 15u) + 1;
  if (((code >> 30) & 1u) == 0u) {
#pragma unroll 1
    for (int i = 0; i < cnt; ++i)
      sphere_test(S.sphere(first + i), L.o, L.d, UNIT ? R(1) : L.a, UNIT ? R(1) : L.inv_a, tmin, L.tbest, L.best, first + i);
  } else if constexpr (!SMEM && sizeof(R) == 4) {
    const char *tp = S.g_tris_g + (size_t)(unsigned)first * sizeof(TriG);  // 32 + 4 bytes per triangle instead of 3 x 16
#pragma unroll 1
    for (int i = 0; i < cnt; ++i, tp += sizeof(TriG)) {
      const F8 a = ldg256(tp);
      const float e2z = __ldg(reinterpret_cast<const float *>(tp + 32));
      tri_test<R>(Vec4<R>{a.v[0], a.v[1], a.v[2], 0.f}, Vec4<R>{a.v[3], a.v[4], a.v[5], 0.f}, Vec4<R>{a.v[6], a.v[7], e2z, 0.f},
                  L.o, L.d, tmin, L.tbest, L.best, (first + i) | (1 << 30));
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < cnt; ++i)
      tri_test<R>(S.tri(first + i, 0), S.tri(first + i, 1), S.tri(first + i, 2), L.o, L.d, tmin, L.tbest, L.best,
                  (first + i) | (1 << 30));
  }
  L.cur = TRAV_POP;
}

// pop, skipping subtrees that start beyond the current best hit; the sentinel (t = -inf) always stops the
// loop and leaves cur == TRAV_DONE: the ray is finished (lane_init resets sp)
template <class R, bool LOCAL>
__device__ __forceinline__ void pop_phase(Lane<R> &L, const Stack<R, LOCAL> &K) {
  R t;
  do {
    L.sp -= K.stride;
    K.load(L.sp, L.cur, t);
  } while (t > L.tbest);
}

// Node visit on NodeQ records (float, global memory, caller-supplied rays): two 256-bit loads instead of seven 128-bit
// ones.  Per axis the near / far byte words are picked by the direction's sign mask (one LOP3 each), every byte
// becomes the float 1 + q * 2^-15 with one PRMT, and its plane distance is one FFMA:
//   t = (origin + q s - o) / d = f * S + B,  S = s 2^15 / d,  B = origin / d - o / d - S.
template <bool TMIN0, bool CHECK>
__device__ __forceinline__ void node_q_math(Lane<float> &L, Stack<float, true> &K, unsigned sp_limit, const F8 &H0, const F8 &H1) {
  const unsigned lox = __float_as_uint(H0.v[6]), loy = __float_as_uint(H0.v[7]), loz = __float_as_uint(H1.v[0]);
  const unsigned hix = __float_as_uint(H1.v[1]), hiy = __float_as_uint(H1.v[2]), hiz = __float_as_uint(H1.v[3]);
  const int4 ch = {__float_as_int(H1.v[4]), __float_as_int(H1.v[5]), __float_as_int(H1.v[6]), __float_as_int(H1.v[7])};
  const float Sx = H0.v[3] * L.idir.x, Sy = H0.v[4] * L.idir.y, Sz = H0.v[5] * L.idir.z;
  const float Bx = fmaf(H0.v[0], L.idir.x, -L.oid.x) - Sx, By = fmaf(H0.v[1], L.idir.y, -L.oid.y) - Sy,
              Bz = fmaf(H0.v[2], L.idir.z, -L.oid.z) - Sz;
  const unsigned nwx = (lox & ~L.onx) | (hix & L.onx), fwx = (hix & ~L.onx) | (lox & L.onx);
  const unsigned nwy = (loy & ~L.ony) | (hiy & L.ony), fwy = (hiy & ~L.ony) | (loy & L.ony);
  const unsigned nwz = (loz & ~L.onz) | (hiz & L.onz), fwz = (hiz & ~L.onz) | (loz & L.onz);
  const float tmin = TMIN0 ? 0.0f : L.tmin;
  float tnx, tny, tnz, tnw;
#define PTB_QF(w, k) __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u | ((k) << 4)))
#define PTB_SLAB(o, k)                                                                                               \
  tn##o = fmaxf(fmaxf(fmaf(PTB_QF(nwx, k), Sx, Bx), fmaf(PTB_QF(nwy, k), Sy, By)), fmaxf(fmaf(PTB_QF(nwz, k), Sz, Bz), tmin)); \
  {                                                                                                                  \
    const float tf_ = fminf(fminf(fmaf(PTB_QF(fwx, k), Sx, Bx), fmaf(PTB_QF(fwy, k), Sy, By)), fmaf(PTB_QF(fwz, k), Sz, Bz)); \
    asm("{\n.reg .pred p;\nsetp.gtu.f32 p, %0, %1;\n@p add.f32 %0, %0, 0f7F800000;\n}" : "+f"(tn##o) : "f"(tf_));    \
  }
  PTB_SLAB(x, 0u)
  PTB_SLAB(y, 1u)
  PTB_SLAB(z, 2u)
  PTB_SLAB(w, 3u)
#undef PTB_SLAB
#undef PTB_QF
  node_finish<float, false, CHECK>(L, K, sp_limit, tnx, tny, tnz, tnw, ch);
}
template <bool CHECK>
__device__ __forceinline__ void node_phase_q(Lane<float> &L, const SceneRef<float, false> &S, Stack<float, true> &K,
                                             unsigned sp_limit) {
  const char *np = S.g_nodes_q + (size_t)(unsigned)L.cur * sizeof(NodeQ);
  const F8 H0 = ldg256(np), H1 = ldg256(np + 32);
  node_q_math<false, CHECK>(L, K, sp_limit, H0, H1);
}

// One traversal step per lane and iteration on 64-byte records (see fused_step_g below for the idea): a NodeQ node,
// ONE triangle (TriG), or the spheres of a leaf that lie in the 64 bytes fetched (3 or 4).
template <bool TMIN0, bool UNIT, bool CHECK>
__device__ __forceinline__ void fused_step_q(Lane<float> &L, const SceneRef<float, false> &S, Stack<float, true> &K,
                                             unsigned sp_limit) {
  const int cur = L.cur;
  if (cur <= TRAV_POP) return;  // (idle, finished)
  const unsigned code = ~(unsigned)cur;
  const bool is_tri = ((code >> 30) & 1u) != 0u;
  const unsigned first = code & 0x3FFFFFFu;
  const char *p = cur >= 0 ? S.g_nodes_q + (size_t)(unsigned)cur * sizeof(NodeQ)
                           : (is_tri ? S.g_tris_g + (size_t)first * sizeof(TriG)
                                     : S.g_sph_g + (size_t)(first & ~1u) * sizeof(Vec4<float>));
  const F8 H0 = ldg256(p), H1 = ldg256(p + 32);
  if (cur >= 0) {
    node_q_math<TMIN0, CHECK>(L, K, sp_limit, H0, H1);
    return;
  }
  const float tmin = TMIN0 ? 0.0f : L.tmin;
  const unsigned cnt = ((code >> 26) & 15u) + 1u;
  if (is_tri) {
    tri_test<float>(Vec4<float>{H0.v[0], H0.v[1], H0.v[2], 0.f}, Vec4<float>{H0.v[3], H0.v[4], H0.v[5], 0.f},
                    Vec4<float>{H0.v[6], H0.v[7], H1.v[0], 0.f}, L.o, L.d, tmin, L.tbest, L.best, (int)(first | (1u << 30)));
    L.cur = cnt <= 1u ? TRAV_POP : (int)~(code + 1u - (1u << 26));  // (first += 1, count -= 1)
  } else {
    // records (first & ~1) .. +3 are in; the leaf starts at the first or the second of them
    const bool odd = (first & 1u) != 0u;
    const unsigned here = min(cnt, odd ? 3u : 4u);
    const float a = UNIT ? 1.0f : L.a, inv_a = UNIT ? 1.0f : L.inv_a;
#define PTB_SPH(k, e0, e1, e2, e3, o0, o1, o2, o3)                                                              \
  if (here > k)                                                                                                   \
    sphere_test(odd ? Vec4<float>{o0, o1, o2, o3} : Vec4<float>{e0, e1, e2, e3}, L.o, L.d, a, inv_a, tmin, L.tbest, \
                L.best, (int)(first + k));
    PTB_SPH(0u, H0.v[0], H0.v[1], H0.v[2], H0.v[3], H0.v[4], H0.v[5], H0.v[6], H0.v[7])
    PTB_SPH(1u, H0.v[4], H0.v[5], H0.v[6], H0.v[7], H1.v[0], H1.v[1], H1.v[2], H1.v[3])
    PTB_SPH(2u, H1.v[0], H1.v[1], H1.v[2], H1.v[3], H1.v[4], H1.v[5], H1.v[6], H1.v[7])
    PTB_SPH(3u, H1.v[4], H1.v[5], H1.v[6], H1.v[7], H1.v[4], H1.v[5], H1.v[6], H1.v[7])
#undef PTB_SPH
    L.cur = cnt <= here ? TRAV_POP : (int)~(code + here - (here << 26));
  }
}

// ---------------------------------------------------------------------------------------------
// Global-memory scenes, float: ONE traversal step per lane and warp iteration, fed by ONE fetch.
// Traversal of a scene that does not fit shared memory is bound by memory latency (ncu: 9 of 10 warp cycles wait on
// a load), so what counts is how many dependent round trips a ray needs and how many lanes have a load in flight.
// Every traversing lane therefore reads 112 bytes at ITS address first — a node (NodeG), the next two triangles of
// its leaf (2 x TriG), or the (up to four) spheres of its leaf — with the same four load instructions, and only
// then the warp branches into the node step and the primitive tests: one round trip per iteration for the whole
// warp, where separate node and leaf phases (and a leaf loop that loads one triangle at a time) paid up to five.
// A triangle leaf with more than two triangles stays the lane's `cur`, advanced by two, for the next iteration.
// ---------------------------------------------------------------------------------------------
template <bool TMIN0, bool UNIT, bool CHECK>
__device__ __forceinline__ void fused_step_g(Lane<float> &L, const SceneRef<float, false> &S, Stack<float, true> &K,
                                             unsigned sp_limit) {
  const int cur = L.cur;
  if (cur <= TRAV_POP) return;  // (idle, finished)
  const unsigned code = ~(unsigned)cur;
  const bool is_tri = ((code >> 30) & 1u) != 0u;
  const unsigned first = code & 0x3FFFFFFu;
  const char *p = cur >= 0 ? S.g_nodes_g + (size_t)(unsigned)cur * sizeof(NodeG)
                           : (is_tri ? S.g_tris_g + (size_t)first * sizeof(TriG)
                                     : S.g_sph_g + (size_t)(first & ~1u) * sizeof(Vec4<float>));
  // (Loading the second half of the record only where it is needed — nodes, a second triangle, a fifth sphere —
  // was measured: 3 % slower on the soup, 4 % on the C3 mesh; profiles/README.md round 2.)
  const unsigned cnt = ((code >> 26) & 15u) + 1u;  // (leaves)
  const F8 X = ldg256(p), Y = ldg256(p + 32), Z = ldg256(p + 64);
  const int4 W = __ldg(reinterpret_cast<const int4 *>(p + 96));
  const float tmin = TMIN0 ? 0.0f : L.tmin;
  if (cur >= 0) {
    float tnx, tny, tnz, tnw;
#define PTB_AXIS(P, i2, no2, a)                                                   \
  float a##l0, a##l1, a##l2, a##l3, a##h0, a##h1, a##h2, a##h3;                    \
  unpack2(ffma2(pack2(P.v[0], P.v[1]), L.i2, L.no2), a##l0, a##l1);                \
  unpack2(ffma2(pack2(P.v[2], P.v[3]), L.i2, L.no2), a##l2, a##l3);                \
  unpack2(ffma2(pack2(P.v[4], P.v[5]), L.i2, L.no2), a##h0, a##h1);                \
  unpack2(ffma2(pack2(P.v[6], P.v[7]), L.i2, L.no2), a##h2, a##h3);
    PTB_AXIS(X, ix2, nox2, x)
    PTB_AXIS(Y, iy2, noy2, y)
    PTB_AXIS(Z, iz2, noz2, z)
#undef PTB_AXIS
    // near / far by min / max of the two plane distances; an unused slot (NaN planes) has a NaN far distance and
    // fails the unordered comparison
#define PTB_SLAB(k, i)                                                                                            \
  tn##k = fmaxf(fmaxf(fminf(xl##i, xh##i), fminf(yl##i, yh##i)), fmaxf(fminf(zl##i, zh##i), tmin));                \
  {                                                                                                               \
    const float tf_ = fminf(fminf(fmaxf(xl##i, xh##i), fmaxf(yl##i, yh##i)), fmaxf(zl##i, zh##i));                 \
    asm("{\n.reg .pred p;\nsetp.gtu.f32 p, %0, %1;\n@p add.f32 %0, %0, 0f7F800000;\n}" : "+f"(tn##k) : "f"(tf_)); \
  }
    PTB_SLAB(x, 0)
    PTB_SLAB(y, 1)
    PTB_SLAB(z, 2)
    PTB_SLAB(w, 3)
#undef PTB_SLAB
    node_finish<float, false, CHECK>(L, K, sp_limit, tnx, tny, tnz, tnw, W);
    return;
  }
  if (is_tri) {
    tri_test<float>(Vec4<float>{X.v[0], X.v[1], X.v[2], 0.f}, Vec4<float>{X.v[3], X.v[4], X.v[5], 0.f},
                    Vec4<float>{X.v[6], X.v[7], Y.v[0], 0.f}, L.o, L.d, tmin, L.tbest, L.best, (int)(first | (1u << 30)));
    if (cnt > 1u)
      tri_test<float>(Vec4<float>{Z.v[0], Z.v[1], Z.v[2], 0.f}, Vec4<float>{Z.v[3], Z.v[4], Z.v[5], 0.f},
                      Vec4<float>{Z.v[6], Z.v[7], __int_as_float(W.x), 0.f}, L.o, L.d, tmin, L.tbest, L.best,
                      (int)((first + 1u) | (1u << 30)));
    L.cur = cnt <= 2u ? TRAV_POP : (int)~(code + 2u - (2u << 26));  // (first += 2, count -= 2)
  } else {
    // records (first & ~1) .. +6 are in; the leaf starts at the first or the second of them
    const bool odd = (first & 1u) != 0u;
    const float a = UNIT ? 1.0f : L.a, inv_a = UNIT ? 1.0f : L.inv_a;
#define PTB_SPH(k, e0, e1, e2, e3, o0, o1, o2, o3)                                                              \
  if (cnt > k)                                                                                                    \
    sphere_test(odd ? Vec4<float>{o0, o1, o2, o3} : Vec4<float>{e0, e1, e2, e3}, L.o, L.d, a, inv_a, tmin, L.tbest, \
                L.best, (int)(first + k));
    PTB_SPH(0u, X.v[0], X.v[1], X.v[2], X.v[3], X.v[4], X.v[5], X.v[6], X.v[7])
    PTB_SPH(1u, X.v[4], X.v[5], X.v[6], X.v[7], Y.v[0], Y.v[1], Y.v[2], Y.v[3])
    PTB_SPH(2u, Y.v[0], Y.v[1], Y.v[2], Y.v[3], Y.v[4], Y.v[5], Y.v[6], Y.v[7])
    PTB_SPH(3u, Y.v[4], Y.v[5], Y.v[6], Y.v[7], Z.v[0], Z.v[1], Z.v[2], Z.v[3])
#undef PTB_SPH
    L.cur = cnt <= 4u ? TRAV_POP : (int)~(code + 4u - (4u << 26));
  }
}

// ---------------------------------------------------------------------------------------------
// textures (texture.ml:16-31) and backgrounds (shirley main.ml:104-110)
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ V3<R> tex_eval(const DTex<R> *__restrict__ texs, int t, R u, R v) {
  DTex<R> T = texs[t];
  if (T.kind == PTB_TEX_CHECKER) {
    R xp = u * R(T.w - 1), yp = v * R(T.h - 1);
    int px = ((int)xp) & 1, py = ((int)yp) & 1;  // Float.to_int truncates toward zero
    T = texs[px == py ? T.even : T.odd];
  }
  return {T.rgb[0], T.rgb[1], T.rgb[2]};
}
template <class R>
__device__ __forceinline__ V3<R> background(const DScene<R> &sc, V3<R> d) {
  if (sc.bg_kind == PTB_BG_CONSTANT) return {sc.bg0[0], sc.bg0[1], sc.bg0[2]};
  // (every ray of the pipeline has a unit direction — Camera.ray and world_ray produce them — so the float path
  // skips the reference's re-normalisation; float64 keeps it)
  V3<R> dn = sizeof(R) == 4 ? d : normalize(d);
  R t = R(0.5) * (dn.y + R(1));
  if constexpr (sizeof(R) == 4)  // c0 + t (c1 - c0): one FFMA per channel (float rounding only; float64 keeps lerp's form)
    return {r_fma(t, sc.bgd[0], sc.bg0[0]), r_fma(t, sc.bgd[1], sc.bg0[1]), r_fma(t, sc.bgd[2], sc.bg0[2])};
  R s = R(1) - t;
  return {sc.bg0[0] * s + sc.bg1[0] * t, sc.bg0[1] * s + sc.bg1[1] * t, sc.bg0[2] * s + sc.bg1[2] * t};
}

// ---------------------------------------------------------------------------------------------
// segmented-queue append (whole warp, convergent).  Lanes with `want` get a destination slot.
// seg_base/seg_fill are warp-uniform: the warp's open segment in this queue (seg_base = NO_SEG: none).
// ---------------------------------------------------------------------------------------------
constexpr unsigned NO_SEG = 0xffffffffu;
// A producer warp takes segments from the queue's global counter SEG_GROUP at a time (one return-value atomic per
// SEG_GROUP * SEG entries) and fills them one after the other; the ones it never opens are closed with count 0 when
// the warp retires (consumers skip empty segments).  An open segment whose index is not the last of its group has
// its successor for free.
constexpr unsigned SEG_GROUP = 4;
__device__ __forceinline__ bool seg_has_next(unsigned seg_base) {
  return seg_base != NO_SEG && ((seg_base / (unsigned)SEG) & (SEG_GROUP - 1u)) != SEG_GROUP - 1u;
}
__device__ __forceinline__ unsigned seg_append(bool want, unsigned &seg_base, unsigned &seg_fill,
                                               unsigned *__restrict__ nseg, int32_t *__restrict__ seg_count,
                                               unsigned lane, unsigned lt_mask) {
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (mask == 0u) return 0u;
  const unsigned cnt = (unsigned)__popc(mask), rank = (unsigned)__popc(mask & lt_mask);
  const unsigned room = (seg_base == NO_SEG) ? 0u : (unsigned)SEG - seg_fill;
  if (cnt <= room) {
    const unsigned dst = seg_base + seg_fill + rank;
    seg_fill += cnt;
    return dst;
  }
  unsigned nb;
  if (seg_has_next(seg_base)) {  // warp-uniform
    nb = seg_base + (unsigned)SEG;
    if (lane == 0) seg_count[seg_base / SEG] = SEG;
  } else {
    unsigned s = 0;
    if (lane == 0) {
      s = atomicAdd(nseg, SEG_GROUP);
      if (seg_base != NO_SEG) seg_count[seg_base / SEG] = SEG;  // the old segment is (or becomes) full
    }
    nb = __shfl_sync(0xffffffffu, s, 0) * (unsigned)SEG;
  }
  const unsigned dst = rank < room ? seg_base + seg_fill + rank : nb + (rank - room);
  seg_base = nb;
  seg_fill = cnt - room;
  return dst;
}
// the warp retires: its open segment gets its fill count, the unopened rest of the group is empty
__device__ __forceinline__ void seg_close(unsigned seg_base, unsigned seg_fill, int32_t *__restrict__ seg_count,
                                          unsigned lane) {
  if (seg_base == NO_SEG) return;
  const unsigned s = seg_base / (unsigned)SEG, last = s | (SEG_GROUP - 1u);
  if (s + lane <= last) seg_count[s + lane] = lane == 0 ? (int32_t)seg_fill : 0;
}

// =================================================================================================
// kernels
// =================================================================================================

// Camera ray of sample `k` of a batch (integrator.ml:96-105, camera.ml:93-102): the sample enumeration is
// pass-major over this rank's pixel list; R2 jitter, cx, cy and the un-normalized direction are float64 with
// un-fused operations (bit-exact against the reference).  Shared by k_raygen and by the bounce-0 traversal
// kernel, which generates its rays in registers instead of reading them from a queue.
struct GenConst {
  int32_t W, spp, npix, pass0, i0, pad;
  double llx, lly, vx, vy, widthf, heightf, alpha0, alpha1;
  double inv_npix, inv_W;  // (1/d) * (1 - 2^-40): quotient estimates that are never too big (udiv_by below)
  const int32_t *pixel_list;
};
// x / d for x < 2^32 without the ~20-instruction integer division sequence: a float64 estimate that is either
// the quotient or one less, then one fix-up.  r receives the remainder.
__device__ __forceinline__ unsigned udiv_by(unsigned x, unsigned d, double inv_d, unsigned &r) {
  unsigned q = (unsigned)__double2uint_rz((double)x * inv_d);
  r = x - q * d;
  if (r >= d) ++q, r -= d;
  return q;
}
// Sample `idx` of the rank's enumeration is pass idx / npix of pixel pixel_list[idx % npix].  The film position
// and un-normalized camera-space direction of that sample:
__device__ __forceinline__ void camera_dir(const GenConst &g, int pixel, int pass, int &offset, double &cx, double &cy,
                                           double &ddx, double &ddy) {
  unsigned rx;
  const int gy = (int)udiv_by((unsigned)pixel, (unsigned)g.W, g.inv_W, rx), gx = (int)rx;
  offset = pixel + (g.pass0 + pass) * g.spp;  // integrator.ml:98 (sic: pass * samples_per_pixel)
  const double dx = r2_sample(g.alpha0, offset), dy = r2_sample(g.alpha1, offset);
  cx = __dmul_rn(__dadd_rn((double)gx, dx), g.widthf);                    // integrator.ml:104
  cy = __dsub_rn(1.0, __dmul_rn(__dadd_rn((double)gy, dy), g.heightf));   // integrator.ml:105
  ddx = __dadd_rn(g.llx, __dmul_rn(g.vx, cx));                            // camera.ml:96-97
  ddy = __dadd_rn(g.lly, __dmul_rn(g.vy, cy));
}
__device__ __forceinline__ void camera_sample(const GenConst &g, unsigned k, int &pixel, int &offset, double &cx,
                                              double &cy, double &ddx, double &ddy) {
  const unsigned idx = (unsigned)g.i0 + k;  // < npix + batch size < 2^31 (checked on the host)
  unsigned ri;
  const unsigned q = udiv_by(idx, (unsigned)g.npix, g.inv_npix, ri);
  pixel = __ldg(g.pixel_list + (int)ri);
  camera_dir(g, pixel, (int)q, offset, cx, cy, ddx, ddy);
}

// Stage 1 — camera rays.  One thread per sample of the batch [first, first+n) of this rank's
// enumeration (pass-major over the rank's pixel list, which is in tile order).
template <class R>
__global__ void __launch_bounds__(256) k_raygen(GenConst g, unsigned n, Queue<R> out, double *__restrict__ dbg_cx,
                                                double *__restrict__ dbg_cy) {
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    int pixel, offset;
    double cx, cy, ddx, ddy;
    camera_sample(g, k, pixel, offset, cx, cy, ddx, ddy);
    if (dbg_cx) dbg_cx[k] = cx;
    if (dbg_cy) dbg_cy[k] = cy;
    V3<R> dir = normalize(V3<R>{R(ddx), R(ddy), R(-1)});
    Vec4<R> *e = out.A(k);
    e[0] = {R(0), R(0), R(0), i2r((int)k, R())};  // (w: where a MODE 1 traversal writes this ray's result)
    e[SEG] = {dir.x, dir.y, dir.z, i2r(offset, R())};
    e[2 * SEG] = {R(1), R(1), R(1), i2r(pixel, R())};
    if (k % SEG == 0) out.seg_count[k / SEG] = (int32_t)(n - k < (unsigned)SEG ? n - k : (unsigned)SEG);  // dense
  }
}

// Stage 2+3 — traversal and primitive tests.  Persistent warps: every warp keeps its 32 lanes busy by
// fetching new rays from the queue (chunked global cursor) whenever fewer than REFILL_BELOW lanes
// are still traversing; finished lanes park their result in registers until that refill point, where
// the whole warp runs the epilogue convergently:
//   miss -> background * attenuation added into the per-pixel sums (integrator.ml:36)
//   hit  -> appended to the queue of its material kind, warp-aggregated with __ballot_sync/__popc
// MODE 0: pipeline.  MODE 1: intersect only (t and caller primitive index per ray).
// The three per-material hit queues are equal slices of one allocation: queue k starts q_slots entries
// (q_slots / SEG segment counters) after queue k-1, `q0` is the first.
//
// Registers are the scarce resource (1024 threads per SM leave 64 each), so everything the traversal loop
// does not touch lives in shared memory: the ray's payload (attenuation, pixel, R2 offset) is parked in a
// per-thread slot at refill and read back at flush, and the warp-uniform bookkeeping (claimed input
// range, open output segments, statistics) sits in a per-warp record.
//
// Dynamic shared memory: [scene (SMEM)] [stack: stack_cap x threads x ENTRY] [payload: threads x (Vec4 + R + int)]
//                        [warp records: warps x WS_WORDS x 4 B] [staging rings: warps x RING x 2 Vec4]
//
// Incoming rays never wait on HBM/L2: each warp keeps the next RING entries (origin, direction) of its claimed
// chunk in flight into its staging ring (cp.async issued one refill ahead), and the payload of a ray goes
// straight from the queue into the lane's payload slot (cp.async, first read at the lane's flush).
#ifndef PTB_LEAF_MIN_G
#define PTB_LEAF_MIN_G 8  // the same threshold for scenes in global memory (phase loops)
#endif
constexpr int LEAF_MIN = 8;  // lanes holding a leaf before the warp runs the leaf phase (tuned: 6..12 equal, 1: -4 %;
                             // global-memory triangle scenes measured with 8 / 12 / 16 / 20 / 24: best at 8..12)
constexpr unsigned WS_NEXT = 0, WS_END = 1, WS_FETCHED = 2, WS_SEG = 3 /* base,fill x 3 kinds */, WS_STAGED = 9,
                   WS_REM_BASE = 10, WS_REM_CNT = 11 /* whole segments claimed but not opened yet */,
                   WS_SEEN = 12 /* the cursor at the last claim */, WS_IB = 13, WS_QB = 14 /* GEN: pixel-list index and
                   pass of entry WS_NEXT */, WS_KINDS = 15 /* shared-window address of the material-kind table */,
                   WS_ROOT = 16 /* `cur` of the root node */, WS_WORDS = 20;
constexpr unsigned RING = 32;  // entries of the per-warp staging ring for incoming rays

// dynamic shared memory per thread besides the staged scene: stack + payload slot + share of the warp record
template <class R>
__host__ __device__ constexpr size_t trace_smem_per_thread(int stack_cap, bool stack_in_smem) {
  return (size_t)(stack_in_smem ? stack_cap : PTB_STACK_HOT) * 2u * sizeof(R) + sizeof(Vec4<R>) + sizeof(R) + 4u /* sp0 */ +
         (WS_WORDS * 4u + 31u) / 32u + RING * 2u * sizeof(Vec4<R>) / 32u;
}

#ifndef PTB_GLOBAL_RENDER_LOOP
#define PTB_GLOBAL_RENDER_LOOP 1
#endif
#ifndef PTB_GLOBAL_BATCH_LOOP
#define PTB_GLOBAL_BATCH_LOOP 2
#endif

// GEN: bounce 0 — the rays are the camera samples [0, gen_n) of the batch, generated in registers
// (camera_sample) instead of being read from `rays`.
// BLK: the block size when it is known at compile time (the 1024-thread float launch of shared-memory scenes: the
// stack's level stride becomes an immediate), 0 = blockDim.x.
template <class R, int MODE, bool SMEM, bool GEN, int BLK = 0>
__global__ void __launch_bounds__(SMEM ? (sizeof(R) == 8 ? 512 : 1024) : 256, SMEM ? 1 : (sizeof(R) == 8 ? 1 : PTB_GLOBAL_BLOCKS))
    k_trace(DScene<R> sc, GenConst gen, unsigned gen_n, Queue<R> rays, const unsigned *__restrict__ nseg_ptr, unsigned nseg_imm,
            unsigned *__restrict__ cursor, int refill_below, Queue<R> q0, unsigned q_slots,
            unsigned *__restrict__ nseg_mat, unsigned *__restrict__ n_traced, int enqueue_hits, R *__restrict__ sums,
            R tmin_arg, R tmax_arg, R *__restrict__ out_t, int32_t *__restrict__ out_prim) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned nseg = nseg_ptr ? *nseg_ptr : nseg_imm;  // segments in the input ray queue
  const int tid = threadIdx.x;
  const unsigned nthreads = BLK ? (unsigned)BLK : blockDim.x;
  const unsigned lane = tid & 31;
  // claim granularity: a quarter, half or whole segment, so that a small launch (late bounces) still
  // spreads over every persistent warp and a big one pays one cursor atomic per 128 rays
  const unsigned per_warp = nseg * (unsigned)SEG / (gridDim.x * (nthreads >> 5));
  const unsigned claim_shift = per_warp >= 512u ? 0u : (per_warp >= 256u ? 1u : 2u);  // units per segment = 1 << shift
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
  SceneRef<R, SMEM> S;
  S.g_nodes = reinterpret_cast<const char *>(sc.nodes), S.g_spheres = sc.spheres, S.g_tris = sc.tris;
  S.g_nodes_g = reinterpret_cast<const char *>(sc.nodes_g), S.g_tris_g = reinterpret_cast<const char *>(sc.tris_g);
  S.g_sph_g = reinterpret_cast<const char *>(sc.spheres_g);
  S.g_nodes_q = reinterpret_cast<const char *>(sc.nodes_q);
  S.g_kinds = sc.prim_kind;
  unsigned scene_bytes = 0;
  if (SMEM) {
    const unsigned nb = (unsigned)sc.n_nodes * (unsigned)sizeof(Node4<R>);
    const unsigned sb = (unsigned)sc.n_spheres * (unsigned)sizeof(Vec4<R>);
    const unsigned tb = (unsigned)sc.n_tris * 3u * (unsigned)sizeof(Vec4<R>);
    const unsigned kb = ((unsigned)(sc.n_spheres + sc.n_tris) + 15u) & ~15u;  // prim_kind is padded to 16 B
    scene_bytes = nb + sb + tb + kb;
    int4 *dst = reinterpret_cast<int4 *>(smem);
    const int4 *s0 = reinterpret_cast<const int4 *>(sc.nodes);
    const int4 *s1 = reinterpret_cast<const int4 *>(sc.spheres);
    const int4 *s2 = reinterpret_cast<const int4 *>(sc.tris);
    const int4 *s3 = reinterpret_cast<const int4 *>(sc.prim_kind);
    for (unsigned i = tid; i < nb / 16; i += nthreads) dst[i] = s0[i];
    for (unsigned i = tid; i < sb / 16; i += nthreads) dst[nb / 16 + i] = s1[i];
    for (unsigned i = tid; i < tb / 16; i += nthreads) dst[(nb + sb) / 16 + i] = s2[i];
    for (unsigned i = tid; i < kb / 16; i += nthreads) dst[(nb + sb + tb) / 16 + i] = s3[i];
    S.s_nodes = smem_base, S.s_spheres = smem_base + nb, S.s_tris = smem_base + nb + sb;
    S.s_kinds = smem_base + nb + sb + tb;
    __syncthreads();
    // keep the two hot base addresses in registers (the compiler otherwise re-derives the shared window base
    // from %cluster_ctaid in every traversal iteration)
    asm volatile("" : "+r"(S.s_nodes), "+r"(S.s_spheres));
    // the shared-memory sphere records carry r^2 (what the intersection test needs)
    Vec4<R> *ssph = reinterpret_cast<Vec4<R> *>(smem + nb);
    for (unsigned i = tid; i < (unsigned)sc.n_spheres; i += nthreads) ssph[i].w *= ssph[i].w;
    // ... and the staged tree refers to inner nodes by their shared-window ADDRESS (SceneRef::nrow)
    Node4<R> *snodes = reinterpret_cast<Node4<R> *>(smem);
    for (unsigned i = tid; i < 4u * (unsigned)sc.n_nodes; i += nthreads) {
      const int c = snodes[i >> 2].child[i & 3u];
      if (c >= 0) {
        snodes[i >> 2].child[i & 3u] = (int)(smem_base + (unsigned)c * (unsigned)sizeof(Node4<R>));
      } else if (sizeof(R) == 4 && c != INT32_MIN) {  // (INT32_MIN = EMPTY_CHILD: an unused slot)
        const unsigned code = ~(unsigned)c;
        if (((code >> 30) & 1u) == 0u)  // sphere leaf -> direct (leaf_phase)
          snodes[i >> 2].child[i & 3u] = -(int)((S.s_spheres + (code & 0x3FFFFFFu) * 16u) | ((code >> 26) & 15u));
      }
    }
    __syncthreads();
  }
  // traversal stack: stack[level][thread], entry = (ref, t_near); level 0 is the sentinel
  constexpr unsigned ENTRY = 2u * (unsigned)sizeof(R);
  Stack<R, !SMEM> K;
  unsigned sp0, sp_limit, pay_base;
  if constexpr (SMEM) {
    K.stride = nthreads * ENTRY;
    const unsigned stk0 = smem_base + scene_bytes + (unsigned)tid * ENTRY;
    K.store(stk0, TRAV_DONE, -Lim<R>::inf());
    sp0 = stk0 + K.stride;
    sp_limit = stk0 + (unsigned)sc.stack_cap * K.stride;  // entries [1, stack_cap) hold pushes
    pay_base = smem_base + scene_bytes + (unsigned)sc.stack_cap * K.stride;
  } else {
    K.s_stride = nthreads * ENTRY;
    K.s_base = smem_base + (unsigned)tid * ENTRY;
    K.store(0u, TRAV_DONE, -Lim<R>::inf());
    sp0 = 1u;
    sp_limit = (unsigned)min(sc.stack_cap, LOCAL_STACK_CAP);
    pay_base = smem_base + Stack<R, true>::HOT * K.s_stride;
  }
  // per-thread payload slot and per-warp record
  const unsigned pay_v = pay_base + (unsigned)tid * (unsigned)sizeof(Vec4<R>);
  const unsigned pay_r = pay_base + nthreads * (unsigned)sizeof(Vec4<R>) + (unsigned)tid * (unsigned)sizeof(R);
  // this thread's first free stack entry, kept in shared memory: a refill reads it back with one LDS instead of
  // re-deriving it from the thread index and the scene size (registers are too scarce to hold it across the loop)
  // (float: the same 4-byte stride as the pay_r array right before it, so the slot is pay_r + a constant)
  const unsigned sp0_slot = sizeof(R) == 4 ? pay_r + nthreads * 4u
                                           : pay_base + nthreads * (unsigned)(sizeof(Vec4<R>) + sizeof(R)) + (unsigned)tid * 4u;
  sts_i32(sp0_slot, (int)sp0);
  const unsigned ws = pay_base + nthreads * (unsigned)(sizeof(Vec4<R>) + sizeof(R) + 4u) + ((unsigned)tid >> 5) * (WS_WORDS * 4u);
  if (lane < WS_WORDS)
    sts_i32(ws + lane * 4u, (lane >= WS_SEG && lane < WS_STAGED && ((lane - WS_SEG) & 1u) == 0u) ? (int)NO_SEG : 0);
  if (SMEM && lane == WS_KINDS) sts_i32(ws + WS_KINDS * 4u, (int)S.s_kinds);  // read back at every flush (one LDS)
  if (lane == WS_ROOT) sts_i32(ws + WS_ROOT * 4u, SMEM ? (int)smem_base : 0);     // read back at every refill
  __syncwarp();
  constexpr bool UNIT = MODE == 0 && sizeof(R) == 4;  // render pipeline, float: unit directions (sphere_test_f)
  // traversal loop of a float scene in global memory: 0 = phases on Node4 records, 1 = one 112-byte fetch per step
  // (NodeG), 2 = phases on quantised NodeQ records, 3 = one 64-byte fetch per step (NodeQ)
  constexpr int GLOOP = (SMEM || sizeof(R) != 4) ? 0 : (MODE == 0 ? PTB_GLOBAL_RENDER_LOOP : PTB_GLOBAL_BATCH_LOOP);
  constexpr bool QM = GLOOP >= 2;  // the lane keeps direction-sign masks instead of row offsets
  constexpr unsigned VB = (unsigned)sizeof(Vec4<R>);
  // staging ring of this warp: RING origins, then RING directions (entry i of the queue sits in slot i % RING)
  const unsigned ring_a = pay_base + nthreads * (unsigned)(sizeof(Vec4<R>) + sizeof(R) + 4u) + (nthreads >> 5) * (WS_WORDS * 4u) +
                          ((unsigned)tid >> 5) * (2u * RING * VB);
  const unsigned ring_b = ring_a + RING * VB;

  Lane<R> L;
  L.cur = TRAV_IDLE, L.sp = sp0, L.best = -1, L.tbest = R(0);
  L.onx = L.ony = L.onz = 0u;
  L.o = L.d = {R(0), R(0), R(1)};
  unsigned ray_i = 0;    // MODE 1: where the result goes
  bool more = true;      // warp-uniform: the queue may still hold segments

  for (;;) {
    // ---------------- flush finished lanes (whole warp, convergent; no global loads) ------------
    if (__ballot_sync(0xffffffffu, L.cur == TRAV_DONE)) {
      const bool done = L.cur == TRAV_DONE;
      if (MODE == 1) {
        if (done) {
          const int slot = L.best & 0x3FFFFFFF;
          out_t[ray_i] = L.best < 0 ? R(NAN) : L.tbest;
          out_prim[ray_i] = L.best < 0 ? -1 : ((L.best >> 30) & 1 ? sc.n_spheres + sc.tri_id[slot] : sc.sphere_id[slot]);
        }
      } else {
        const unsigned lt_mask = (1u << lane) - 1u;
        int kind = -1;
        Vec4<R> pv = {R(0), R(0), R(0), R(0)};  // (attenuation, pixel)
        if (!GEN) cp_async_wait_all();  // this lane's payload copy (issued at its refill, long since complete)
        if (done) {
          pv = lds_vec4(pay_v, R());
          if (L.best < 0) {
            // integrator.ml:36: emit0 + attn0 * background ray (emit0 == 0: Material.emit is black)
            const V3<R> bg = background(sc, L.d);
            const int pixel = r2i(pv.w);
            atomicAdd(&sums[3 * (size_t)pixel + 0], pv.x * bg.x);
            atomicAdd(&sums[3 * (size_t)pixel + 1], pv.y * bg.y);
            atomicAdd(&sums[3 * (size_t)pixel + 2], pv.z * bg.z);
          } else if (enqueue_hits) {
            if (SMEM)
              kind = lds_u8((unsigned)lds_i32(ws + WS_KINDS * 4u) + (unsigned)((L.best & 0x3FFFFFFF) + (((L.best >> 30) & 1) ? sc.n_spheres : 0)));
            else
              kind = S.kind(L.best, sc.n_spheres);
          }
        }
        // Per-material segmented queues: no global atomic unless a segment fills up.  Every hit lane reads the open
        // segment (base, fill) of ITS kind from the warp record (lanes of one kind read the same word: a broadcast),
        // ranks itself among the lanes of that kind, and the kind's first lane writes the new fill back.  Only when
        // some kind's segment overflows (once per 128 hits of that kind) does its first lane open a new one.
        const unsigned m0 = __ballot_sync(0xffffffffu, kind == 0), m1 = __ballot_sync(0xffffffffu, kind == 1),
                       m2 = __ballot_sync(0xffffffffu, kind == 2);
        if ((m0 | m1 | m2) != 0u) {
          const unsigned mk = kind == 0 ? m0 : (kind == 1 ? m1 : m2);
          const unsigned cnt = (unsigned)__popc(mk), rank = (unsigned)__popc(mk & lt_mask);
          const unsigned kk = (unsigned)max(kind, 0);
          const unsigned wsk = ws + (WS_SEG + 2u * kk) * 4u;
          const unsigned base = (unsigned)lds_i32(wsk), fill = (unsigned)lds_i32(wsk + 4u);
          const unsigned room = (base == NO_SEG) ? 0u : (unsigned)SEG - fill;
          const bool over = kind >= 0 && cnt > room;
          unsigned nb = 0;
          if (__any_sync(0xffffffffu, over)) {
            unsigned sn = 0;
            if (over && rank == 0u) {  // this kind's first lane: the old segment is (or becomes) full, open the next:
              if (base != NO_SEG) q0.seg_count[kk * (q_slots / SEG) + base / SEG] = SEG;
              // the successor inside the group it was taken with, or SEG_GROUP fresh ones (the only atomic)
              sn = seg_has_next(base) ? base / (unsigned)SEG + 1u : atomicAdd(&nseg_mat[kk], SEG_GROUP);
            }
            nb = __shfl_sync(0xffffffffu, sn, (__ffs((int)mk) - 1) & 31) * (unsigned)SEG;
          }
          __syncwarp();  // every lane has read (base, fill) before the first lanes update them
          if (kind >= 0) {
            if (rank == 0u) {
              if (!over) sts_i32(wsk + 4u, (int)(fill + cnt));
              else sts_i32(wsk, (int)nb), sts_i32(wsk + 4u, (int)(cnt - room));
            }
            const unsigned dst = rank < room ? base + fill + rank : nb + (rank - room);
            const unsigned e = kk * q_slots + dst;  // entry in the joint allocation
            Vec4<R> *qe = q0.A(e);  // one address, the fields at constant distances (segment-interleaved layout)
            qe[0] = {r_fma(L.tbest, L.d.x, L.o.x), r_fma(L.tbest, L.d.y, L.o.y), r_fma(L.tbest, L.d.z, L.o.z), pv.w};
            qe[SEG] = {L.d.x, L.d.y, L.d.z, lds_r(pay_r, R())};
            qe[2 * SEG] = {pv.x, pv.y, pv.z, i2r(L.best, R())};
          }
          __syncwarp();
        }
      }
      if (done) L.cur = TRAV_IDLE;
    }
    // ---------------- refill idle lanes --------------------------------------------------------
    // The warp works through a CHUNK [next, end) of contiguous queue entries.  Chunks come from claims of 1, 2 or 4
    // whole segments (guided: big claims while much of the queue is left, single segments towards its end; parts
    // of a segment when the launch is small); a claim's run of full segments is one chunk.  What the NEXT refill
    // will take is already in flight into the staging ring, and a new chunk is opened as soon as the old one runs
    // out, not when its entries are needed, so that a refill normally finds everything in shared memory.
    const unsigned idle = __ballot_sync(0xffffffffu, L.cur == TRAV_IDLE);
    if (more && idle) {
      unsigned next = (unsigned)lds_i32(ws + WS_NEXT * 4u), end = (unsigned)lds_i32(ws + WS_END * 4u);
      unsigned staged = (unsigned)lds_i32(ws + WS_STAGED * 4u);  // [next, staged) is in, or on its way into, the ring
      unsigned ib = 0, qb = 0;
      if (GEN) ib = (unsigned)lds_i32(ws + WS_IB * 4u), qb = (unsigned)lds_i32(ws + WS_QB * 4u);
      const int root = lds_i32(ws + WS_ROOT * 4u);
      const unsigned sp0r = (unsigned)lds_i32(sp0_slot);
      __syncwarp();
      unsigned idle_left = idle;
      for (;;) {
        if (next < end) {  // hand entries [next, next + take) to the idle lanes
          cp_async_wait_all();
          __syncwarp();
          const unsigned want = (unsigned)__popc(idle_left), avail = end - next;
          const unsigned take = want < avail ? want : avail;
          const unsigned rank = (unsigned)__popc(idle_left & ((1u << lane) - 1u));
          const bool mine = ((idle_left >> lane) & 1u) != 0u && rank < take;
          if (mine) {
            ray_i = next + rank;
            if (GEN) {
              unsigned i = ib + rank, q = qb;
              while (i >= (unsigned)gen.npix) i -= (unsigned)gen.npix, ++q;
              const int pixel = lds_i32(ring_a + (ray_i % RING) * 4u);
              int offset;
              double cx, cy, ddx, ddy;
              camera_dir(gen, pixel, (int)q, offset, cx, cy, ddx, ddy);
              const V3<R> dir = normalize(V3<R>{R(ddx), R(ddy), R(-1)});
              const Vec4<R> pv = {R(1), R(1), R(1), i2r(pixel, R())};
              sts_vec4(pay_v, pv);
              sts_r(pay_r, i2r(offset, R()));
              lane_init<R, UNIT, QM>(L, V3<R>{R(0), R(0), R(0)}, dir, R(0), Lim<R>::tmax(), sp0r, root);
            } else {
              const Vec4<R> A = lds_vec4(ring_a + (ray_i % RING) * VB, R()), B = lds_vec4(ring_b + (ray_i % RING) * VB, R());
              if (MODE == 0) {
                // the payload, (attenuation, pixel) and the R2 offset, goes from the queue straight into this lane's slot
                cp_async_vec4<R, 2>(pay_v, rays.A(ray_i));
                sts_r(pay_r, B.w);
              }
              if (MODE == 1) ray_i = (unsigned)r2i(A.w);  // the caller's index of this ray (the queue may be in sorted order)
              lane_init<R, UNIT, QM>(L, V3<R>{A.x, A.y, A.z}, V3<R>{B.x, B.y, B.z}, (MODE == 0) ? R(0) : tmin_arg,
                                 (MODE == 0) ? Lim<R>::tmax() : tmax_arg, sp0r, root);
            }
            // (Visiting the root right here, for all refilled lanes at once — broadcast shared-memory reads, no loop
            // iteration — was measured: +4.7 % trace time, profiles/README.md round 2.)
          }
          idle_left &= ~__ballot_sync(0xffffffffu, mine);  // (also orders the ring reads before the next copies into it)
          next += take;
          if (GEN) {
            ib += take;
            while (ib >= (unsigned)gen.npix) ib -= (unsigned)gen.npix, ++qb;
          }
        }
        if (next >= end) {  // open the next chunk
          unsigned rem_base = (unsigned)lds_i32(ws + WS_REM_BASE * 4u), rem_cnt = (unsigned)lds_i32(ws + WS_REM_CNT * 4u);
          __syncwarp();
          if (rem_cnt == 0u) {  // nothing left of the last claim: claim more of the input queue
            const unsigned nunits = nseg << claim_shift;
            unsigned g = 1u;
            if (claim_shift == 0u) {
              const unsigned seen = (unsigned)lds_i32(ws + WS_SEEN * 4u), nwarps = gridDim.x * (nthreads >> 5);
              const unsigned left = seen < nunits ? nunits - seen : 0u;
              g = left > 16u * nwarps ? 4u : (left > 8u * nwarps ? 2u : 1u);
            }
            unsigned u = 0;
            if (lane == 0) u = atomicAdd(cursor, g);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= nunits) {
              more = false;
              next = end = staged = 0;
              break;
            }
            if (claim_shift != 0u) {  // small launch: a unit is a part of one segment
              const unsigned claim = (unsigned)SEG >> claim_shift, sgm = u >> claim_shift;
              unsigned c = 0;
              if (GEN) {
                c = gen_n - sgm * (unsigned)SEG < (unsigned)SEG ? gen_n - sgm * (unsigned)SEG : (unsigned)SEG;
              } else {
                if (lane == 0) c = (unsigned)rays.seg_count[sgm];
                c = __shfl_sync(0xffffffffu, c, 0);
              }
              const unsigned off = (u & ((1u << claim_shift) - 1u)) * claim;  // offset of the unit inside its segment
              const unsigned lim = c < off + claim ? c : off + claim;         // valid entries end here
              next = sgm * SEG + off;
              end = lim > off ? sgm * SEG + lim : next;
            } else {
              rem_base = u, rem_cnt = nunits - u < g ? nunits - u : g;
              if (lane == 0) sts_i32(ws + WS_SEEN * 4u, (int)u);
            }
          }
          if (rem_cnt != 0u) {  // the run of full segments at the head of the remainder, plus the first partial one
            unsigned c = (unsigned)SEG;
            if (lane < rem_cnt) {
              if (GEN) {
                const unsigned s0 = (rem_base + lane) * (unsigned)SEG;
                c = gen_n - s0 < (unsigned)SEG ? gen_n - s0 : (unsigned)SEG;
              } else {
                c = (unsigned)rays.seg_count[rem_base + lane];
              }
            }
            const unsigned partial = __ballot_sync(0xffffffffu, c != (unsigned)SEG);
            const unsigned first_p = partial ? (unsigned)__ffs((int)partial) - 1u : rem_cnt;
            const unsigned nsegs = first_p < rem_cnt ? first_p + 1u : rem_cnt;
            const unsigned last_c = __shfl_sync(0xffffffffu, c, nsegs - 1u);
            next = rem_base * SEG;
            end = (rem_base + nsegs - 1u) * SEG + last_c;
            rem_base += nsegs, rem_cnt -= nsegs;
            if (lane == 0) sts_i32(ws + WS_REM_BASE * 4u, (int)rem_base), sts_i32(ws + WS_REM_CNT * 4u, (int)rem_cnt);
          }
          if (lane == 0 && MODE == 0) sts_i32(ws + WS_FETCHED * 4u, lds_i32(ws + WS_FETCHED * 4u) + (int)(end - next));
          staged = next;
          if (GEN) qb = udiv_by((unsigned)gen.i0 + next, (unsigned)gen.npix, gen.inv_npix, ib);
          if (next >= end) continue;  // an empty unit (the tail of a partly filled segment)
        }
        {  // keep the next RING entries of the chunk in flight
          const unsigned upto = end - next > RING ? next + RING : end;
          const unsigned e = staged + lane;
          if (e < upto) {
            if (GEN) {
              unsigned i = ib + (e - next);
              while (i >= (unsigned)gen.npix) i -= (unsigned)gen.npix;
              cp_async_scalar(ring_a + (e % RING) * 4u, reinterpret_cast<const float *>(gen.pixel_list + i));
            } else {
              const Vec4<R> *src = rays.A(e);
              cp_async_vec4<R, 0>(ring_a + (e % RING) * VB, src);
              cp_async_vec4<R, 1>(ring_b + (e % RING) * VB, src);
            }
          }
          if (staged < upto) staged = upto;
        }
        if (idle_left == 0u) break;
      }
      if (lane == 0) {
        sts_i32(ws + WS_NEXT * 4u, (int)next), sts_i32(ws + WS_END * 4u, (int)end);
        sts_i32(ws + WS_STAGED * 4u, (int)staged);
        if (GEN) sts_i32(ws + WS_IB * 4u, (int)ib), sts_i32(ws + WS_QB * 4u, (int)qb);
      }
      __syncwarp();
    }
    const unsigned active = __ballot_sync(0xffffffffu, L.cur > TRAV_DONE);
    if (active == 0u) {
      if (!more) break;
      continue;
    }
    // ---------------- traverse until too few lanes are left -------------------------------------
    const int keep = more ? refill_below : 1;
    // The leaf phase is POSTPONED until at least LEAF_MIN lanes hold a leaf (or every traversing lane does):
    // lanes waiting with a leaf sit out node phases, but the primitive tests then run with most of the warp
    // instead of a handful of lanes (warp-loop replay on the host, scripts/bvh_sim: -6 % warp instructions).
    unsigned act = active;
    int keep_r = keep;
    asm volatile("" : "+r"(keep_r));  // loop-invariant: keep it in a register instead of re-deriving it
    if constexpr (GLOOP == 1) {
      do {
        fused_step_g<MODE == 0, UNIT, true>(L, S, K, sp_limit);
        if (L.cur == TRAV_POP) pop_phase<R, true>(L, K);
        act = __ballot_sync(0xffffffffu, L.cur > TRAV_DONE);
      } while (__popc(act) >= keep_r);
    } else if constexpr (GLOOP == 3) {
      do {
        fused_step_q<MODE == 0, UNIT, true>(L, S, K, sp_limit);
        if (L.cur == TRAV_POP) pop_phase<R, true>(L, K);
        act = __ballot_sync(0xffffffffu, L.cur > TRAV_DONE);
      } while (__popc(act) >= keep_r);
    } else
    do {
      if constexpr (GLOOP == 2) {
        if (L.cur >= 0) node_phase_q<true>(L, S, K, sp_limit);
      } else {
        if (L.cur >= 0) node_phase<R, SMEM, MODE == 0, sizeof(R) == 8, !SMEM>(L, S, K, sp_limit);
      }
      const bool at_leaf = (unsigned)(L.cur - (TRAV_POP + 1)) < (unsigned)(0 - (TRAV_POP + 1));  // TRAV_POP < cur < 0
      const unsigned lm = __ballot_sync(0xffffffffu, at_leaf);
      if (__popc(lm) >= (SMEM ? LEAF_MIN : PTB_LEAF_MIN_G) || lm == act) {  // (lm == 0 never equals act inside the loop)
        if (at_leaf) leaf_phase<R, SMEM, MODE == 0, UNIT>(L, S);
      }
      if (L.cur == TRAV_POP) pop_phase<R, !SMEM>(L, K);
      act = __ballot_sync(0xffffffffu, L.cur > TRAV_DONE);
    } while (__popc(act) >= keep_r);
  }
  if (MODE == 0) {
    __syncwarp();
    const unsigned q_segs = q_slots / SEG;
    seg_close((unsigned)lds_i32(ws + (WS_SEG + 0u) * 4u), (unsigned)lds_i32(ws + (WS_SEG + 1u) * 4u), q0.seg_count, lane);
    seg_close((unsigned)lds_i32(ws + (WS_SEG + 2u) * 4u), (unsigned)lds_i32(ws + (WS_SEG + 3u) * 4u), q0.seg_count + q_segs, lane);
    seg_close((unsigned)lds_i32(ws + (WS_SEG + 4u) * 4u), (unsigned)lds_i32(ws + (WS_SEG + 5u) * 4u), q0.seg_count + 2u * q_segs, lane);
    const unsigned fetched = (unsigned)lds_i32(ws + WS_FETCHED * 4u);
    if (lane == 0 && fetched) atomicAdd(n_traced, fetched);
  }
}

// Stage 4 — material scatter.  Work item = one warp-sized chunk of ONE material's hit queue, so the
// material switch below is warp-uniform.
#ifndef PTB_SHADE_NBUF
#define PTB_SHADE_NBUF 3
#endif
constexpr unsigned SHADE_NBUF = PTB_SHADE_NBUF;  // staging buffers per thread: SHADE_NBUF - 1 items in flight while one is shaded
template <class R>
__host__ __device__ constexpr size_t shade_smem_bytes(unsigned block, bool bulk = false) {
  // BULK, per warp: SHADE_NBUF buffers of [A x 32][B x 32][C x 32], then the mbarriers (64 B)
  return bulk ? (size_t)(block / 32u) * (SHADE_NBUF * 96u * sizeof(Vec4<R>) + 64u)
              : (size_t)SHADE_NBUF * (3u * block * sizeof(Vec4<R>) + (block / 32u) * 4u);
}
// BULK: the entries of a work item come in with bulk asynchronous copies (the TMA unit's 1-D form, SASS UBLKCP): one
// lane issues three 512-byte copies — the item's A, B and C runs, contiguous in the segment-interleaved layout —
// tracked by an mbarrier, instead of 32 lanes issuing three 16-byte cp.async each (96 requests -> 3).  Measured on
// the 4K frame (profiles/README.md, round 2): 0.82 -> 0.91 of the HBM copy bandwidth, and the bounce-0 launch no
// longer wants fewer blocks than the others.  The outgoing rays stay per-lane 16-byte stores: staging them in shared
// memory for bulk stores was measured too and is slower (two more warp syncs and a proxy fence per item:
// profiles/experiments/r2_kshade_bulk_in_and_out.patch).
template <class R, bool BULK>
__global__ void __launch_bounds__(256)
    k_shade(DScene<R> sc, RenderConst rc, int bounce, Queue<R> q0, unsigned q_slots,
            unsigned *__restrict__ nseg_mat, Queue<R> out, unsigned *__restrict__ nseg_out, R *__restrict__ sums, int last) {
  // Software pipeline over the warp's work items (one item = 32 entries of one segment of one material's hit
  // queue): while item k is shaded, the 48 B entries of items k+1 .. k+SHADE_NBUF-1 and the fill counts of their
  // segments are in flight into shared memory (cp.async: no register and no scoreboard is tied up, each lane
  // copies and later reads only its own entry), enough bytes in flight per SM to keep HBM busy.  The copy does
  // not wait for the fill count: entries past it are allocated queue slots that are simply not used.
  extern __shared__ __align__(16) unsigned char shade_smem[];
  const unsigned lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned c0 = nseg_mat[0], c1 = nseg_mat[1], c2 = nseg_mat[2];
  const unsigned total = (c0 + c1 + c2) * (SEG / 32);
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  const int j = 2 + 2 * bounce;  // take_2d cursor (integrator.ml:20-28): every hit so far took two
  const double alpha_u = rc.alpha[j], alpha_v = rc.alpha[j + 1];
  unsigned ob = NO_SEG, of = 0;  // warp-uniform: this warp's open segment in the output ray queue
  constexpr unsigned VB = (unsigned)sizeof(Vec4<R>);
  // staging slots: [buffer][A,B,C][thread], then [warp][buffer] fill counts
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(shade_smem);
  const unsigned slot0 = sbase + threadIdx.x * VB, bufsz = 3u * blockDim.x * VB, arr = blockDim.x * VB;
  const unsigned cnt0 = sbase + SHADE_NBUF * bufsz + (threadIdx.x >> 5) * (SHADE_NBUF * 4u);
  // BULK layout of this warp's region
  constexpr unsigned IBUF = 96u * VB;  // one buffer: three fields x 32 entries
  const unsigned wbase = sbase + (threadIdx.x >> 5) * (SHADE_NBUF * IBUF + 64u);
  const unsigned mbar0 = wbase + SHADE_NBUF * IBUF;
  unsigned fill[SHADE_NBUF - 1];  // BULK, lane 0: fill counts of the segments of the items in flight
  unsigned parity = 0u;
  if constexpr (BULK) {
    if (lane == 0) {
#pragma unroll
      for (unsigned b = 0; b < SHADE_NBUF; ++b) mbar_init(mbar0 + 8u * b, 1u);
      mbar_fence_init();
    }
    __syncwarp();
  }
  // item -> (material, segment, first entry)
  auto decode = [&](unsigned w, int &m, unsigned &seg, unsigned &i0) {
    seg = w / (SEG / 32);
    i0 = (w % (SEG / 32)) * 32;
    if (seg < c0) m = 0;
    else if (seg < c0 + c1) m = 1, seg -= c0;
    else m = 2, seg -= c0 + c1;
  };
  auto prefetch = [&](unsigned w, unsigned buf) -> unsigned {
    int m;
    unsigned seg, i0;
    decode(w, m, seg, i0);
    if constexpr (BULK) {
      unsigned f = 0;
      if (lane == 0) {
        const Vec4<R> *src = q0.A((unsigned)m * q_slots + seg * SEG + i0);
        const unsigned dst = wbase + buf * IBUF, mb = mbar0 + 8u * buf;
        mbar_expect_tx(mb, IBUF);
        bulk_g2s(dst, src, 32u * VB, mb);
        bulk_g2s(dst + 32u * VB, src + SEG, 32u * VB, mb);
        bulk_g2s(dst + 64u * VB, src + 2 * SEG, 32u * VB, mb);
        f = (unsigned)__ldg(q0.seg_count + (unsigned)m * (q_slots / SEG) + seg);
      }
      return f;
    }
    const unsigned i = (unsigned)m * q_slots + seg * SEG + i0 + lane;  // entry in the joint allocation
    const unsigned dst = slot0 + buf * bufsz;
    const Vec4<R> *src = q0.A(i);
    cp_async_vec4<R, 0>(dst, src);
    cp_async_vec4<R, 1>(dst + arr, src);
    cp_async_vec4<R, 2>(dst + 2u * arr, src);
    if (lane == 0)
      cp_async_scalar(cnt0 + buf * 4u, reinterpret_cast<const float *>(q0.seg_count + (unsigned)m * (q_slots / SEG) + seg));
    return 0u;
  };
  // Items are handed out dynamically, a few at a time (nseg_mat[3] is the launch's cursor): the SMs do not all get
  // the same share of the memory system (GPCs differ in size), and with a static split the slowest ones set
  // the kernel's time.  The claim for the range after next is issued when a range is opened, so its latency is
  // covered by the items of a whole range.
  constexpr unsigned NONE = 0xffffffffu;
  unsigned *cursor = nseg_mat + 3;
  const unsigned per_claim = total / (warps * 4u) >= 8u ? 8u : (total / (warps * 4u) >= 1u ? total / (warps * 4u) : 1u);
  unsigned range_next = 0, range_end = 0, pending = 0;
  bool claims_left = true;
  if (lane == 0) pending = atomicAdd(cursor, per_claim);
  auto next_item = [&]() -> unsigned {  // warp-uniform
    if (range_next == range_end) {
      if (!claims_left) return NONE;
      const unsigned base = __shfl_sync(0xffffffffu, pending, 0);
      if (base >= total) {
        claims_left = false;
        return NONE;
      }
      range_next = base, range_end = total - base < per_claim ? total : base + per_claim;
      if (lane == 0) pending = atomicAdd(cursor, per_claim);
    }
    return range_next++;
  };
  unsigned ids[SHADE_NBUF - 1];  // the items in flight, ids[0] is shaded next
#pragma unroll
  for (unsigned k = 0; k + 1 < SHADE_NBUF; ++k) {
    ids[k] = next_item();
    fill[k] = 0u;
    if (ids[k] != NONE) fill[k] = prefetch(ids[k], k);
    if (!BULK) cp_async_commit();
  }
  unsigned buf = 0;
  while (ids[0] != NONE) {
    __syncwarp();  // every lane is done with the buffer of the item before, which is filled next
    const unsigned w = ids[0];
    const unsigned fw = fill[0];
#pragma unroll
    for (unsigned k = 0; k + 2 < SHADE_NBUF; ++k) ids[k] = ids[k + 1], fill[k] = fill[k + 1];
    {
      const unsigned wn = next_item();
      ids[SHADE_NBUF - 2] = wn;
      fill[SHADE_NBUF - 2] = 0u;
      if (wn != NONE) fill[SHADE_NBUF - 2] = prefetch(wn, buf == 0u ? SHADE_NBUF - 1u : buf - 1u);
    }
    int m;
    unsigned seg, i0;
    decode(w, m, seg, i0);
    const unsigned cbuf = buf;
    unsigned nm;
    if constexpr (BULK) {
      mbar_wait(mbar0 + 8u * cbuf, (parity >> cbuf) & 1u);  // item w has landed
      parity ^= 1u << cbuf;
      nm = __shfl_sync(0xffffffffu, fw, 0);
    } else {
      cp_async_commit();
      cp_async_wait<SHADE_NBUF - 1>();  // item w has landed
      __syncwarp();                     // (lane 0 copied the count for the warp)
      nm = (unsigned)lds_i32(cnt0 + cbuf * 4u);  // valid entries of the item's segment
    }
    buf = buf + 1u == SHADE_NBUF ? 0u : buf + 1u;
    if (i0 >= nm) continue;
   {
    const bool valid = i0 + lane < nm;
    bool alive = false;
    V3<R> no = {R(0), R(0), R(0)}, nd = {R(0), R(0), R(1)}, nattn = {R(0), R(0), R(0)};
    Vec4<R> A = {R(0), R(0), R(0), R(0)}, B = A;
    if (valid) {
      const unsigned src = BULK ? wbase + cbuf * IBUF + lane * VB : slot0 + cbuf * bufsz;
      const unsigned fstride = BULK ? 32u * VB : arr;
      A = lds_vec4(src, R());
      B = lds_vec4(src + fstride, R());
      const Vec4<R> C = lds_vec4(src + 2u * fstride, R());
      V3<R> p = {A.x, A.y, A.z};
      const V3<R> d = {B.x, B.y, B.z};
      const int offset = r2i(B.w);
      const int best = r2i(C.w);
      const int slot = best & 0x3FFFFFFF;
      const bool is_tri = (best >> 30) & 1;
      // table look-ups first (dependent loads: slot -> material row -> material), the sample arithmetic in between
      Vec4<R> v0 = {R(0), R(0), R(0), R(0)}, e1 = v0, e2 = v0, s = v0;
      int mrow;
      if (!is_tri) {
        s = ldg_vec4(sc.spheres + slot);
        mrow = __ldg(sc.sphere_mat + slot);
      } else {
        v0 = ldg_vec4(sc.tris + 3 * (size_t)slot), e1 = ldg_vec4(sc.tris + 3 * (size_t)slot + 1);
        e2 = ldg_vec4(sc.tris + 3 * (size_t)slot + 2);
        mrow = __ldg(sc.tri_mat + slot);
      }
      const double ud = r2_sample(alpha_u, offset), vd = r2_sample(alpha_v, offset);
      DMat mat;
      {
        const int4 m0 = __ldg(reinterpret_cast<const int4 *>(sc.mats + mrow)),
                   m1 = __ldg(reinterpret_cast<const int4 *>(sc.mats + mrow) + 1);
        mat.kind = m0.x, mat.tex = m0.y, mat.index = __hiloint2double(m0.w, m0.z);
        mat.tex_kind = m1.x, mat.rgb[0] = __int_as_float(m1.y), mat.rgb[1] = __int_as_float(m1.z), mat.rgb[2] = __int_as_float(m1.w);
      }
      // geometric normal (sphere.ml:21 / triangle.ml:18-23)
      V3<R> n;
      if (!is_tri) {
        n = normalize(V3<R>{p.x - s.x, p.y - s.y, p.z - s.z});
        if (sizeof(R) == 4) p = {r_fma(n.x, s.w, s.x), r_fma(n.y, s.w, s.y), r_fma(n.z, s.w, s.z)};
      } else {
        n = normalize(cross(V3<R>{e1.x, e1.y, e1.z}, V3<R>{e2.x, e2.y, e2.z}));
      }
      const bool front = dot(d, n) < R(0);  // sphere.ml:60 / triangle.ml:56
      if (!front) n = neg(n);
      const Quat<R> frame = frame_from_normal(n);              // Shader_space.create
      const V3<R> wi = quat_transform(frame, neg(d));           // Shader_space.omega_i
      const Quat<R> frame_inv = {frame.r, neg(frame.v)};
      V3<R> albedo = {R(1), R(1), R(1)};
      if (m != PTB_MAT_DIELECTRIC) {
        if (sizeof(R) == 4 && mat.tex_kind == PTB_TEX_SOLID) {
          albedo = {R(mat.rgb[0]), R(mat.rgb[1]), R(mat.rgb[2])};
        } else if (sizeof(R) == 8 && mat.tex_kind == PTB_TEX_SOLID) {
          const DTex<R> T = sc.texs[mat.tex];
          albedo = {T.rgb[0], T.rgb[1], T.rgb[2]};
        } else {
          R tu, tv;
          if (!is_tri) {  // Sphere.tex_coord (sphere.ml:25-33) on the (flipped) normal
            const R PI = R(3.141592653589793);
            R theta, phi;
            if (sizeof(R) == 4) {
              theta = atan2f(sqrtf(fmaf((float)n.x, (float)n.x, (float)n.z * (float)n.z)), -(float)n.y);
              phi = PI + atan2f(-(float)n.z, (float)n.x);
            } else {
              theta = acos(-(double)n.y);
              phi = PI + atan2(-(double)n.z, (double)n.x);
            }
            tu = phi * (R(1) / (R(2) * PI));
            tv = theta * (R(1) / PI);
          } else {  // barycentric mix of the three tex coords (triangle.ml:49-54)
            const V3<R> a1 = {e1.x, e1.y, e1.z}, a2 = {e2.x, e2.y, e2.z}, a3 = {p.x - v0.x, p.y - v0.y, p.z - v0.z};
            R d00 = dot(a1, a1), d01 = dot(a1, a2), d11 = dot(a2, a2), d20 = dot(a3, a1), d21 = dot(a3, a2);
            R inv = R(1) / (d00 * d11 - d01 * d01);
            R bu = (d11 * d20 - d01 * d21) * inv, bv = (d00 * d21 - d01 * d20) * inv, bw = R(1) - bu - bv;
            const R *uv = sc.tri_uv + 6 * (size_t)slot;
            tu = uv[0] * bw + uv[2] * bu + uv[4] * bv;
            tv = uv[1] * bw + uv[3] * bu + uv[5] * bv;
          }
          albedo = tex_eval(sc.texs, mat.tex, tu, tv);
        }
      }
      const V3<R> attn0 = {C.x, C.y, C.z};
      V3<R> dir_ss = {R(0), R(0), R(1)};
      if (m == PTB_MAT_LAMBERTIAN && mat.kind == PTB_MAT_EMISSIVE) {
        // extension: Material.emit = the texture, scatter = Absorb -> the path ends with emit0 + attn0 * emit
        // (integrator.ml:40,43; emit0 == 0 on every path because the only emitting kind absorbs).  Emissive hits
        // travel in the Lambertian queue.
        const int pixel = r2i(A.w);
        atomicAdd(&sums[3 * (size_t)pixel + 0], attn0.x * albedo.x);
        atomicAdd(&sums[3 * (size_t)pixel + 1], attn0.y * albedo.y);
        atomicAdd(&sums[3 * (size_t)pixel + 2], attn0.z * albedo.z);
      } else if (last) {
        // the path has used its max_bounces intersections: whatever it would scatter contributes black
        // (integrator.ml:31-32); this launch exists only to collect the emission above
      } else if (m == PTB_MAT_LAMBERTIAN) {
        // Scatter.Diffuse; Pdf.sample diffuse_plus_light (pdf.ml:5-9, shader_space.ml:56-64)
        if (!sc.has_light) {
          // diffuse_plus_light = Pdf.diffuse (render_command.ml:81): diffuse_pd / divisor == 1 exactly,
          // diffuse_pd = 0 iff z = 0 (integrator.ml:48-58)
          R u = r_min(R(ud), Lim<R>::below_one()), sn, cs;
          R rr = r_sqrt_fast(u);  // float: MUFU.SQRT (2 ulp); double: exact
          r_sincos2pi(R(vd), &sn, &cs);
          dir_ss = {rr * cs, rr * sn, r_sqrt_fast(R(1) - u)};
          alive = dir_ss.z > R(0);
          nattn = {albedo.x * attn0.x, albedo.y * attn0.y, albedo.z * attn0.z};
        } else {
          // extension: Pdf.Mix (Diffuse, Quad_light) — the first coordinate picks the component (ptb200.h)
          const V3<R> lo = {sc.light_o[0], sc.light_o[1], sc.light_o[2]}, lu = {sc.light_u[0], sc.light_u[1], sc.light_u[2]},
                      lv = {sc.light_v[0], sc.light_v[1], sc.light_v[2]};
          if (ud < 0.5) {
            R u = r_min(R(2.0 * ud), Lim<R>::below_one()), sn, cs;
            R rr = r_sqrt_fast(u);
            r_sincos2pi(R(vd), &sn, &cs);
            dir_ss = {rr * cs, rr * sn, r_sqrt_fast(R(1) - u)};
          } else {
            const R lu_ = R((2.0 * ud) - 1.0), lv_ = R(vd);
            const V3<R> q = (lo + lu * lu_) + lv * lv_;
            dir_ss = quat_transform(frame, normalize(q - p));
          }
          const R PI = R(3.141592653589793);
          const R diffuse_pd = dir_ss.z < R(0) ? R(0) : dir_ss.z / PI;  // Pdf.eval Pdf.diffuse (pdf.ml:11-15)
          // Pdf.eval Quad_light: solid-angle density of the direction on the parallelogram
          R light_pd = R(0);
          {
            const V3<R> w = quat_transform(frame_inv, dir_ss);
            const V3<R> N = cross(lu, lv);
            const R nn = dot(N, N), area = r_sqrt(nn), denom = dot(w, N) / area;
            if (r_abs(denom) > R(1e-9)) {
              const R t = (dot(lo - p, N) / area) / denom;
              if (t > R(1e-9)) {
                const V3<R> rel = (p + w * t) - lo;
                const R a = dot(N, cross(rel, lv)) / nn, b = dot(N, cross(lu, rel)) / nn;
                if (R(0) <= a && a <= R(1) && R(0) <= b && b <= R(1)) light_pd = (t * t) / (r_abs(denom) * area);
              }
            }
          }
          const R divisor = R(0.5) * (diffuse_pd + light_pd);
          const R pd = diffuse_pd / divisor;  // integrator.ml:55
          alive = diffuse_pd != R(0) && isfinite(pd);  // integrator.ml:51,56
          const V3<R> att = {pd * albedo.x, pd * albedo.y, pd * albedo.z};  // Color.scale attenuation pd (integrator.ml:60)
          nattn = {att.x * attn0.x, att.y * attn0.y, att.z * attn0.z};
        }
      } else if (m == PTB_MAT_METAL) {
        // material.ml:28-44: mirror, Schlick-tinted; omega_r.z <= 0 -> Absorb
        dir_ss = {-wi.x, -wi.y, wi.z};
        alive = wi.z > R(0);
        R s5 = r_pow5(R(1) - wi.z);
        V3<R> att = {albedo.x + (R(1) - albedo.x) * s5, albedo.y + (R(1) - albedo.y) * s5,
                     albedo.z + (R(1) - albedo.z) * s5};
        nattn = {att.x * attn0.x, att.y * attn0.y, att.z * attn0.z};
      } else {
        // material.ml:45-57: Dielectric; reflect on TIR or schlick > u, else refract
        R c = r_min(r_max(wi.z, R(0)), R(1));
        R s = r_sqrt_fast(R(1) - c * c);
        R ratio = front ? R(1.0 / mat.index) : R(mat.index);
        R q0_ = (R(1) - ratio) / (R(1) + ratio);
        R r0 = q0_ * q0_;
        R sch = r0 + (R(1) - r0) * r_pow5(R(1) - c);
        if (ratio * s > R(1) || sch > R(ud)) {
          dir_ss = {-wi.x, -wi.y, wi.z};
        } else {  // Shader_space.refract (shader_space.ml:41-49)
          R cc = r_min(wi.z, R(1));
          V3<R> perp = {(R(0) - wi.x) * ratio, (R(0) - wi.y) * ratio, (cc - wi.z) * ratio};
          dir_ss = {perp.x, perp.y, perp.z - r_sqrt_fast(r_abs(R(1) - dot(perp, perp)))};
        }
        alive = true;
        nattn = attn0;  // Color.white
      }
      // Shader_space.world_ray (shader_space.ml:51-54)
      nd = quat_transform(frame_inv, dir_ss);
      no = {r_fma(nd.x, R(1e-3), p.x), r_fma(nd.y, R(1e-3), p.y), r_fma(nd.z, R(1e-3), p.z)};
    }
    {
      const unsigned dst = seg_append(alive, ob, of, nseg_out, out.seg_count, lane, lt_mask);
      if (alive) {
        Vec4<R> *oe = out.A(dst);
        oe[0] = {no.x, no.y, no.z, R(0)};
        oe[SEG] = {nd.x, nd.y, nd.z, B.w};
        oe[2 * SEG] = {nattn.x, nattn.y, nattn.z, A.w};
      }
    }
   }
  }
  seg_close(ob, of, out.seg_count, lane);
}

// batch bookkeeping: fold the finished batch's ray counts into the totals, reset the counters and
// publish the next batch's ray-queue segment count.
__global__ void k_batch_ctl(Ctl *ctl, unsigned next_n, int max_bounces, unsigned long long *progress,
                            unsigned long long paths_done) {
  const int t = threadIdx.x;
  if (t == 0 && progress) {  // polled by ptb_render_progress: zero-copy host memory
    *reinterpret_cast<volatile unsigned long long *>(progress) = paths_done;
    __threadfence_system();
  }
  if (t < max_bounces) {  // only bounces that were actually traced (integrator.ml:31-32)
    unsigned v = ctl->n_rays[t];
    ctl->rays_by_bounce[t] += v;
    if (v) atomicAdd(&ctl->total_rays, (unsigned long long)v);
  }
  __syncthreads();
  if (t <= MAX_BOUNCES) {
    ctl->n_rays[t] = 0u;
    ctl->cursor[t] = 0u;
    ctl->nseg_rays[t] = (t == 0) ? (next_n + SEG - 1) / SEG : 0u;
  }
  if (t < MAX_BOUNCES) {
    ctl->nseg_mat[t][0] = ctl->nseg_mat[t][1] = ctl->nseg_mat[t][2] = ctl->nseg_mat[t][3] = 0u;
  }
}

// Stage 5 — reconstruction filter + gamma.  out[X,Y] = sum_d w(d) * S[X-dx, Y-dy] over sources inside
// the image (film_tile.ml:23-38 splats, integrator.ml:114-128 drops destinations outside the image),
// then sqrt(x * 1/spp) (integrator.ml:152-154).
template <class R, class OUT>
__global__ void __launch_bounds__(256) k_resolve(const R *__restrict__ sums, OUT *__restrict__ out, int W, int H,
                                                 double spp_inv, int flags, double w00, double w01, double w02,
                                                 double w10, double w11, double w12, double w20, double w21,
                                                 double w22) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= W * H) return;
  const int y = p / W, x = p - y * W;
  double acc[3] = {0, 0, 0};
  if (flags & PTB_FLAG_NO_FILTER) {
    for (int c = 0; c < 3; ++c) acc[c] = (double)sums[3 * (size_t)p + c];
  } else {
    const double wt[3][3] = {{w00, w01, w02}, {w10, w11, w12}, {w20, w21, w22}};
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int sx = x - dx, sy = y - dy;
        if (sx < 0 || sx >= W || sy < 0 || sy >= H) continue;
        const double wgt = wt[dy + 1][dx + 1];
        const R *s = sums + 3 * ((size_t)sy * W + sx);
        for (int c = 0; c < 3; ++c) acc[c] = fma(wgt, (double)s[c], acc[c]);
      }
  }
  const bool raw = flags & (PTB_FLAG_RAW_SUMS | PTB_FLAG_NO_FILTER);
  for (int c = 0; c < 3; ++c) out[3 * (size_t)p + c] = (OUT)(raw ? acc[c] : sqrt(acc[c] * spp_inv));
}

// Multi-GPU framebuffer reduce: device 0 adds the other devices' per-pixel sums into its own, reading them
// straight out of peer memory (NVLink loads, 128-bit, coalesced) — no staging copy, no second pass.
struct PeerPtrs {
  const float *src[8];
  int n;
};
// Commit of a float scene that stays in global memory: the tree and the triangles once more as NodeG / TriG records.
__global__ void __launch_bounds__(256) k_make_g_layout(const Node4<float> *__restrict__ nodes, int n_nodes,
                                                       const Vec4<float> *__restrict__ tris, int n_tris,
                                                       const Vec4<float> *__restrict__ spheres, int n_spheres,
                                                       NodeG *__restrict__ ng, TriG *__restrict__ tg,
                                                       Vec4<float> *__restrict__ sg, NodeQ *__restrict__ nq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_spheres + 8) {  // (c, r^2); the 8 records of padding keep the 112-byte fetch of the last leaf inside the array
    Vec4<float> v = {0.f, 0.f, 0.f, 0.f};
    if (i < n_spheres) v = spheres[i], v.w *= v.w;
    sg[i] = v;
  }
  if (i < n_nodes) {
    const Node4<float> n = nodes[i];
    NodeG g;
    for (int k = 0; k < 4; ++k) {
      const bool unused = n.child[k] == INT32_MIN;  // EMPTY_CHILD
      for (int a = 0; a < 3; ++a) g.pl[a][k] = unused ? NAN : n.lo[a][k], g.pl[a][4 + k] = unused ? NAN : n.hi[a][k];
      g.child[k] = n.child[k], g.pad[k] = 0;
    }
    ng[i] = g;
    // 8-bit child boxes relative to the node's own box.  All of this is exact in double (float origin, power-of-two
    // step, q <= 255); `margin` covers what the kernel's decode adds to the usual slab rounding: the roundings of
    // S = s 2^15 / d and of B (about 3e-3 of a step, i.e. < 2.4e-5 of the extent).
    NodeQ q;
    for (int a = 0; a < 3; ++a) {
      double mn = 1e300, mx = -1e300;
      for (int k = 0; k < 4; ++k)
        if (n.child[k] != INT32_MIN) mn = fmin(mn, (double)n.lo[a][k]), mx = fmax(mx, (double)n.hi[a][k]);
      if (mn > mx) mn = mx = 0.0;  // (no children at all: never visited)
      const double margin = 1e-4 * (mx - mn) + 1e-30;
      const float org = __double2float_rd(mn - 2.0 * margin);
      const double need = (mx + margin) - (double)org;
      int e = -60;
      if (need > 0.0) {
        frexp(need / 254.0, &e);  // need / 254 = m 2^e, m in [0.5, 1): 2^e >= need / 254
        e = max(e, -100);
      }
      const double st = ldexp(1.0, e);
      q.origin[a] = org;
      q.scale15[a] = (float)ldexp(1.0, e + 15);
      unsigned wl = 0, wh = 0;
      for (int k = 0; k < 4; ++k) {
        unsigned ql = 255u, qh = 0u;  // unused slot: inverted box
        if (n.child[k] != INT32_MIN) {
          const double lo = n.lo[a][k], hi = n.hi[a][k];
          double fl = floor((lo - (double)org) / st);
          if (lo - ((double)org + fl * st) < margin && fl > 0.0) fl -= 1.0;
          double ce = ceil((hi - (double)org) / st);
          if (((double)org + ce * st) - hi < margin) ce += 1.0;
          ql = (unsigned)fmin(fmax(fl, 0.0), 255.0), qh = (unsigned)fmin(fmax(ce, 0.0), 255.0);
        }
        wl |= ql << (8 * k), wh |= qh << (8 * k);
      }
      q.qlo[a] = wl, q.qhi[a] = wh;
    }
    for (int k = 0; k < 4; ++k) q.child[k] = n.child[k];
    nq[i] = q;
  }
  if (i < n_tris) {
    const Vec4<float> v0 = tris[3 * (size_t)i], e1 = tris[3 * (size_t)i + 1], e2 = tris[3 * (size_t)i + 2];
    TriG t;
    t.v[0] = v0.x, t.v[1] = v0.y, t.v[2] = v0.z, t.v[3] = e1.x, t.v[4] = e1.y, t.v[5] = e1.z;
    t.v[6] = e2.x, t.v[7] = e2.y, t.v[8] = e2.z;
    for (int k = 9; k < 16; ++k) t.v[k] = 0.f;
    tg[i] = t;
  }
}

__global__ void __launch_bounds__(256) k_reduce_peers(float *__restrict__ dst, PeerPtrs pp, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;  // grid covers ceil(n / 4) threads
  if (i >= n) return;
  if (i + 3 < n) {
    float4 a = *reinterpret_cast<const float4 *>(dst + i);
    for (int k = 0; k < pp.n; ++k) {
      const float4 b = *reinterpret_cast<const float4 *>(pp.src[k] + i);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    *reinterpret_cast<float4 *>(dst + i) = a;
  } else {
    for (size_t j = i; j < n; ++j) {
      float a = dst[j];
      for (int k = 0; k < pp.n; ++k) a += pp.src[k][j];
      dst[j] = a;
    }
  }
}

// R2 stream dump: out[i*D + d] = get(offsets[i], d), through the SAME device function as the pipeline
__global__ void k_r2_stream(RenderConst rc, int D, const int32_t *__restrict__ offsets, long long n,
                            double *__restrict__ out) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * D) return;
  long long i = k / D;
  int dim = (int)(k - i * D);
  out[k] = r2_sample(rc.alpha[dim], offsets[i]);
}

// FP32 issue-rate probe (roofline denominator): 16 independent FFMA chains per thread
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// user rays (3 floats each) -> ray queue
template <class R>
__global__ void k_pack_rays(const float *__restrict__ o, const float *__restrict__ d, long long n, Queue<R> q) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Vec4<R> *e = q.A((unsigned)i);
  e[0] = {R(o[3 * i]), R(o[3 * i + 1]), R(o[3 * i + 2]), i2r((int)i, R())};  // w: the result index
  e[SEG] = {R(d[3 * i]), R(d[3 * i + 1]), R(d[3 * i + 2]), R(0)};
  if (i % SEG == 0) q.seg_count[i / SEG] = (int32_t)(n - i < SEG ? n - i : SEG);  // dense
}

}  // namespace ptb

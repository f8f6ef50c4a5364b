// oracle.cpp — CPU ORACLE (float64).  TEST INFRASTRUCTURE ONLY — see oracle.h.
//
// A function-by-function restatement of the per-pixel render path of dalev/path-tracer-ocaml.  Every
// function cites the reference file:line it follows (paths relative to the reference repo).  Compile
// with -ffp-contract=off: ocamlopt never fuses a*b+c; FMAs appear only where the reference writes
// `Float.fma` (or the Rust kernel writes `_mm256_fmadd_pd`), and here they are explicit std::fma.
//
// PINNING STATUS (what the reference's own tests/fixtures hold for this path):
//   pinned by reference tests, checked in tests/test_oracle_known_answers.py:
//     Bbox.is_hit (path_tracer_test.ml:121-130), Tile.split/iter (…:34-70), Film_tile coords and
//     write_pixel locus (…:72-119), unit_square_to_hemisphere norm (…:132-142),
//     Low_discrepancy_sequence 1-D integrals (low_discrepancy_sequence_test.ml:28-57);
//   pinned by the reference's one golden artefact shirley-spheres.png (README.md:3,7), PIXEL BY PIXEL: the
//     600x300 / 32 spp / 8 bounce render of this oracle, quantised floor(255 v), equals the PNG in every byte
//     of every pixel (tests/test_golden_png.py; fixture tests/golden/shirley_png_rgb8.npz).  That one artefact
//     exercises and therefore pins: the scene generator (Base.Random.float over OCaml 5's LXM), Camera,
//     Mat4.look_at, the binned-SAH Shape_tree build and its ordered traversal, Bbox, the Rust AVX sphere kernel
//     (Simd_leaf), Sphere.hit / tex_coord, Texture (checker), Material.scatter for all three kinds,
//     Shader_space / Quaternion, the R2 stream, Integrator (dimension consumption, pass*spp offset), Film_tile
//     splat, stitch, gamma and the PNG quantisation;
//   PARITY UNPINNED by any reference test, artefact or runnable reference (no OCaml/Rust toolchain here, so no
//     oracle/_ref): Triangle.intersect / Triangle.Hit.to_hit and the scalar Sphere.intersect inside a mixed
//     tree (cornell / ganesha geometry — the reference renders those scenes with photon mapping, never through
//     Integrator).  For those this restatement itself is the oracle, cross-checked by analytic cases and, for the
//     scalar sphere test, by the `--no-simd` render matching the PNG to +-1 LSB on 99.5 % of the pixels;
//   EXTENSION, not in the reference (marked EXT below): an emissive material and an area-light
//     mixture pdf behind the hooks the reference already threads (Material.emit, Hit.emit, Pdf.t,
//     ~diffuse_plus_light).  Its only oracle is this file.
//   Third-party behaviour assumed: Base `Float.min/max` propagate NaN; Base `List.min_elt` keeps the
//     first minimum; `Num.float_of_num` yields the nearest double; glibc libm = OCaml's Float.*.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_HAVE_AVX2 1
#endif

namespace orc {

// ---------------------------------------------------------------------------------------------
// affine.ml — V3 / P3 (path_tracer/src/affine.ml:13-74)
// ---------------------------------------------------------------------------------------------
struct V3 {
  double x, y, z;
};
static inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // affine.ml:45
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // affine.ml:46
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // affine.ml:47
static inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }                             // affine.ml:49
// V3.fma u v w = map3 Float.fma (affine.ml:53)
static inline V3 vfma(V3 u, V3 v, V3 w) {
  return {std::fma(u.x, v.x, w.x), std::fma(u.y, v.y, w.y), std::fma(u.z, v.z, w.z)};
}
// V3.dot (affine.ml:60): fma v.x w.x (fma v.y w.y (v.z *. w.z))
static inline double dot(V3 v, V3 w) { return std::fma(v.x, w.x, std::fma(v.y, w.y, v.z * w.z)); }
static inline V3 scale(V3 v, double s) { return {s * v.x, s * v.y, s * v.z}; }  // affine.ml:61
static inline double quadrance(V3 v) { return dot(v, v); }                       // affine.ml:62
// V3.lerp t v w = scale v (1-t) + scale w t (affine.ml:63)
static inline V3 lerp(double t, V3 v, V3 w) { return scale(v, 1.0 - t) + scale(w, t); }
// V3.normalize (affine.ml:65-68)
static inline V3 normalize(V3 v) {
  double s = 1.0 / std::hypot(v.x, std::hypot(v.y, v.z));
  return scale(v, s);
}
// V3.cross (affine.ml:70-73): h w x y z = fma w x (-(y*z))
static inline V3 cross(V3 p, V3 q) {
  double a = p.x, b = p.y, c = p.z, d = q.x, e = q.y, f = q.z;
  auto h = [](double w, double x, double y, double z) { return std::fma(w, x, -(y * z)); };
  return {h(b, f, c, e), h(c, d, a, f), h(a, e, b, d)};
}

// Base Float.min / Float.max: NaN if either argument is NaN (assumed; only matters for 0*inf).
static inline double bmin(double x, double y) {
  if (std::isnan(x) || std::isnan(y)) return NAN;
  return x < y ? x : y;
}
static inline double bmax(double x, double y) {
  if (std::isnan(x) || std::isnan(y)) return NAN;
  return x > y ? x : y;
}
static inline double min_coord(V3 v) { return bmin(v.x, bmin(v.y, v.z)); }  // affine.ml:56
static inline double max_coord(V3 v) { return bmax(v.x, bmax(v.y, v.z)); }  // affine.ml:57
static inline double axis_of(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// ---------------------------------------------------------------------------------------------
// ray.ml (path_tracer/src/ray.ml:1-15)
// ---------------------------------------------------------------------------------------------
struct Ray {
  V3 o, d, dinv;
};
static inline Ray ray_create(V3 o, V3 d) { return {o, d, {1.0 / d.x, 1.0 / d.y, 1.0 / d.z}}; }
static inline V3 point_at(const Ray &r, double t) { return r.o + scale(r.d, t); }

// ---------------------------------------------------------------------------------------------
// low_discrepancy_sequence.ml (low_discrepancy_sequence/src/low_discrepancy_sequence.ml:1-36)
// ---------------------------------------------------------------------------------------------
static double phi_approx(int d) {  // :8-17
  double dp = 1.0 / ((double)d + 1.0);
  double x = 2.0;
  for (int it = 0; it < 100000; ++it) {  // reference loops until x = x'; cap guards a 2-cycle
    double xp = std::pow(1.0 + x, dp);
    if (x == xp) return x;
    x = xp;
  }
  return x;
}
static void lds_alpha(int dimension, double *alpha) {  // :22-25
  double phi = phi_approx(dimension);
  for (int i = 0; i < dimension; ++i) alpha[i] = 1.0 / std::pow(phi, (double)(i + 1));
}
static inline double lds_get(const double *alpha, int64_t offset, int dimension) {  // :19-20,33-36
  double x = 0.5 + alpha[dimension] * (double)(1 + offset);
  return x - std::trunc(x);
}

// ---------------------------------------------------------------------------------------------
// bbox.ml (path_tracer/src/bbox.ml:1-64)
// ---------------------------------------------------------------------------------------------
struct Bbox {
  V3 mn, mx;
};
static inline V3 bbox_center(const Bbox &b) { return scale(b.mn + b.mx, 0.5); }  // :12
static inline Bbox bbox_union(const Bbox &t, const Bbox &u) {                      // :14-18
  return {{bmin(t.mn.x, u.mn.x), bmin(t.mn.y, u.mn.y), bmin(t.mn.z, u.mn.z)},
          {bmax(t.mx.x, u.mx.x), bmax(t.mx.y, u.mx.y), bmax(t.mx.z, u.mx.z)}};
}
static inline double surface_area(const Bbox &b) {  // :33-38
  V3 e = b.mx - b.mn;
  double a = std::fma(e.x, e.y, std::fma(e.y, e.z, e.z * e.x));
  return 2.0 * a;
}
// Bbox.hit_range / is_hit (:40-56)
static inline bool bbox_is_hit(const Bbox &b, const Ray &r, double t_min, double t_max) {
  V3 t0 = (b.mn - r.o) * r.dinv;
  V3 t1 = (b.mx - r.o) * r.dinv;
  double a = max_coord({bmin(t0.x, t1.x), bmin(t0.y, t1.y), bmin(t0.z, t1.z)});
  double bb = min_coord({bmax(t0.x, t1.x), bmax(t0.y, t1.y), bmax(t0.z, t1.z)});
  double lo = bmax(t_min, a);
  double hi = bmin(t_max, bb);
  return lo <= hi;
}

// ---------------------------------------------------------------------------------------------
// quaternion.ml (path_tracer/src/quaternion.ml:1-42)
// ---------------------------------------------------------------------------------------------
struct Quat {
  double r;
  V3 v;
};
static inline Quat quat_normalize(Quat q) {  // :11-15
  double s = 1.0 / std::hypot(std::hypot(q.r, q.v.x), std::hypot(q.v.y, q.v.z));
  return {q.r * s, scale(q.v, s)};
}
static inline Quat quat_mul(Quat a, Quat b) {  // :25-32
  double r = (a.r * b.r) - dot(a.v, b.v);
  V3 v = (cross(a.v, b.v) + scale(b.v, a.r)) + scale(a.v, b.r);
  return {r, v};
}
static inline Quat quat_conj(Quat q) { return {q.r, neg(q.v)}; }  // :34-37
static inline V3 quat_transform(Quat t, V3 v) {                   // :39-42
  return quat_mul(quat_mul(t, Quat{0.0, v}), quat_conj(t)).v;
}

// ---------------------------------------------------------------------------------------------
// shader_space.ml (path_tracer/src/shader_space.ml:1-69)
// ---------------------------------------------------------------------------------------------
struct ShaderSpace {
  Quat rotation;
  V3 origin, normal;
};
static inline ShaderSpace ss_create(V3 n, V3 origin) {  // :11-23
  const double epsilon = 1e-9;
  Quat rot;
  if (n.z > 1.0 - epsilon)
    rot = {1.0, {0, 0, 0}};
  else if (n.z < epsilon - 1.0)
    rot = {0.0, {0.0, 1.0, 0.0}};
  else
    rot = quat_normalize({1.0 + n.z, {n.y, -n.x, 0.0}});
  return {rot, origin, n};
}
static inline V3 ss_rotate(const ShaderSpace &s, V3 v) { return quat_transform(s.rotation, v); }
static inline V3 ss_rotate_inv(const ShaderSpace &s, V3 v) {  // :29-32
  return quat_transform(quat_conj(s.rotation), v);
}
static inline V3 ss_reflect(V3 v) { return {-v.x, -v.y, v.z}; }  // :34-39
static inline V3 ss_refract(V3 wi, double index) {               // :41-49
  double c = bmin(wi.z, 1.0);
  V3 perp = scale(v3(0.0, 0.0, c) - wi, index);
  V3 para = {0.0, 0.0, -std::sqrt(std::fabs(1.0 - quadrance(perp)))};
  return perp + para;
}
static inline Ray ss_world_ray(const ShaderSpace &s, V3 dir_ss) {  // :51-54
  V3 dir = ss_rotate_inv(s, dir_ss);
  return ray_create(s.origin + scale(dir, 1e-3), dir);
}
static inline V3 unit_square_to_hemisphere(double u, double v) {  // :56-64
  double r = std::sqrt(u);
  double theta = v * 2.0 * M_PI;
  double x = r * std::cos(theta);
  double y = r * std::sin(theta);
  double z = std::sqrt(1.0 - u);
  return {x, y, z};
}
static inline V3 ss_omega_i(const ShaderSpace &s, const Ray &ray) {  // :66-69
  return ss_rotate(s, neg(ray.d));
}

// ---------------------------------------------------------------------------------------------
// texture.ml / material.ml / scatter.ml / pdf.ml
// ---------------------------------------------------------------------------------------------
struct Scene;
struct TexCoord {
  double u, v;
};
enum ScatterKind { ABSORB = 0, SPECULAR = 1, DIFFUSE = 2 };  // scatter.ml:1-4
struct Scatter {
  ScatterKind kind;
  Ray ray;  // Specular
  V3 attenuation;
};

static inline int float_to_int_parity(double a) {  // texture.ml:24  `Float.to_int a land 1`
  long long i = (long long)a;                      // truncation toward zero
  return (int)(i & 1);
}

// ---------------------------------------------------------------------------------------------
// Scene tables + primitives
// ---------------------------------------------------------------------------------------------
struct Sphere {
  V3 c;
  double r;
  int mat;
};
struct Tri {
  int a, b, c;
  int mat;
  TexCoord ta, tb, tc;
};
struct EltHit {
  double t;
  int prim;  // set-order id: sphere i -> i ; triangle j -> nS + j
  double u, v;
};

struct Counters {
  orc_counters c;
  Counters() { std::memset(&c, 0, sizeof c); }
};

struct Leaf {
  std::vector<int> prims;                // set-order ids, in leaf order
  std::vector<double> xs, ys, zs, rs;    // Simd_leaf coords, NaN padded to x4 (main.ml:177-193)
};
struct Node {
  Bbox box;
  int leaf = -1;
  int axis = 0;
  int lhs = -1, rhs = -1;
};

struct Scene {
  std::vector<ptb_texture> tex;
  std::vector<ptb_material> mat;
  std::vector<Sphere> spheres;
  std::vector<V3> verts;
  std::vector<Tri> tris;
  int bg_kind = PTB_BG_GRADIENT_Y;
  V3 bg0{1, 1, 1}, bg1{0.5, 0.7, 1.0};
  // EXT: ~diffuse_plus_light = Pdf.Mix (Diffuse, Quad_light {origin; u; v}) when set, Pdf.diffuse otherwise
  bool has_light = false;
  V3 light_o{0, 0, 0}, light_u{1, 0, 0}, light_v{0, 0, 1};
  // tree
  int leaf_kind = ORC_LEAF_SIMD;
  int length_cutoff = 16;
  std::vector<Node> nodes;
  std::vector<Leaf> leaves;
  int root = -1;
  bool committed = false;
};

// Texture.eval (texture.ml:16-31)
static V3 texture_eval(const Scene &s, int t, TexCoord coord) {
  const ptb_texture &T = s.tex[t];
  if (T.kind == PTB_TEX_SOLID) return {T.rgb[0], T.rgb[1], T.rgb[2]};
  double width = (double)(T.width - 1), height = (double)(T.height - 1);
  double xp = coord.u * width, yp = coord.v * height;
  int sel = (float_to_int_parity(xp) == float_to_int_parity(yp)) ? T.even : T.odd;
  return texture_eval(s, sel, coord);
}

// Material.schlick_reflectance (material.ml:16-20)
static inline double schlick_reflectance(double cos_theta, double index) {
  double q = (1.0 - index) / (1.0 + index);
  double r0 = q * q;
  return r0 + ((1.0 - r0) * std::pow(1.0 - cos_theta, 5.0));
}
// Float.clamp_exn ~min ~max (Base): if t < min then min else if max < t then max else t
static inline double clamp_exn(double t, double lo, double hi) {
  if (t < lo) return lo;
  if (hi < t) return hi;
  return t;
}

// Material.scatter (material.ml:22-57), applied to u (Hit.scatter, hit.ml:9)
static Scatter material_scatter(const Scene &s, int m, const ShaderSpace &ss, TexCoord tc, V3 omega_i,
                                bool hit_front, double u) {
  const ptb_material &M = s.mat[m];
  Scatter out;
  out.kind = ABSORB;
  out.attenuation = {0, 0, 0};
  out.ray = Ray{};
  if (M.kind == PTB_MAT_LAMBERTIAN) {
    out.kind = DIFFUSE;
    out.attenuation = texture_eval(s, M.texture, tc);
  } else if (M.kind == PTB_MAT_METAL) {
    V3 omega_r = ss_reflect(omega_i);
    double z = omega_r.z;
    if (z <= 0.0) {
      out.kind = ABSORB;
    } else {
      V3 a = texture_eval(s, M.texture, tc);
      double sch = std::pow(1.0 - omega_i.z, 5.0);
      V3 c = scale(v3(1, 1, 1) - a, sch);
      out.kind = SPECULAR;
      out.attenuation = a + c;
      out.ray = ss_world_ray(ss, omega_r);
    }
  } else if (M.kind == PTB_MAT_EMISSIVE) {
    out.kind = ABSORB;  // EXT: a light only emits (Material.emit below)
  } else {
    double index = M.index, index_inv = 1.0 / M.index;  // material.ml:13
    double wi_z = omega_i.z;
    double c = clamp_exn(wi_z, 0.0, 1.0);
    double sn = std::sqrt(1.0 - c * c);
    double ratio = hit_front ? index_inv : index;
    V3 wo;
    if (ratio * sn > 1.0 || schlick_reflectance(c, ratio) > u)
      wo = ss_reflect(omega_i);
    else
      wo = ss_refract(omega_i, ratio);
    out.kind = SPECULAR;
    out.attenuation = {1, 1, 1};
    out.ray = ss_world_ray(ss, wo);
  }
  return out;
}

// Material.emit (material.ml:59): black for every kind of the reference.  EXT: Emissive tex -> Texture.eval tex.
static V3 material_emit(const Scene &s, int m, TexCoord tc) {
  const ptb_material &M = s.mat[m];
  if (M.kind == PTB_MAT_EMISSIVE) return texture_eval(s, M.texture, tc);
  return {0, 0, 0};
}

// Pdf.eval Diffuse (pdf.ml:11-15)
static inline double pdf_eval_diffuse(V3 dir) { return (dir.z < 0.0) ? 0.0 : dir.z / M_PI; }

// EXT — Pdf.Quad_light {origin; u; v}: directions towards a parallelogram light, density with respect to solid
// angle.  `sample` and `eval` have the signatures of pdf.ml:5-15 (shader space in, shader-space direction out / in).
static V3 pdf_sample_quad_light(const Scene &sc, const ShaderSpace &ss, double u, double v) {
  V3 q = (sc.light_o + scale(sc.light_u, u)) + scale(sc.light_v, v);
  return ss_rotate(ss, normalize(q - ss.origin));
}
static double pdf_eval_quad_light(const Scene &sc, V3 dir, const ShaderSpace &ss) {
  V3 w = ss_rotate_inv(ss, dir);
  V3 N = cross(sc.light_u, sc.light_v);
  double nn = dot(N, N);
  double area = std::sqrt(nn);
  double denom = dot(w, N) / area;  // cosine at the light
  if (!(std::fabs(denom) > 1e-9)) return 0.0;
  double t = (dot(sc.light_o - ss.origin, N) / area) / denom;
  if (!(t > 1e-9)) return 0.0;
  V3 rel = (ss.origin + scale(w, t)) - sc.light_o;
  double a = dot(N, cross(rel, sc.light_v)) / nn, b = dot(N, cross(sc.light_u, rel)) / nn;
  if (!(0.0 <= a && a <= 1.0 && 0.0 <= b && b <= 1.0)) return 0.0;
  return (t * t) / (std::fabs(denom) * area);
}
// EXT — Pdf.Mix (Diffuse, Quad_light), equal weights; the first coordinate of the 2-D sample picks the component
// and is stretched back to [0,1)
static V3 pdf_sample_mix(const Scene &sc, const ShaderSpace &ss, double u, double v) {
  if (u < 0.5) return unit_square_to_hemisphere(2.0 * u, v);
  return pdf_sample_quad_light(sc, ss, (2.0 * u) - 1.0, v);
}
static double pdf_eval_mix(const Scene &sc, V3 dir, const ShaderSpace &ss) {
  return 0.5 * (pdf_eval_diffuse(dir) + pdf_eval_quad_light(sc, dir, ss));
}

// ---------------------------------------------------------------------------------------------
// sphere.ml (sphere/src/sphere.ml:1-69)
// ---------------------------------------------------------------------------------------------
static inline Bbox sphere_bbox(const Sphere &s) {  // :16-19
  V3 r = {s.r, s.r, s.r};
  return {s.c + neg(r), s.c + r};
}
// Sphere.intersect (scalar OCaml) :35-54
static inline bool sphere_intersect_scalar(V3 center, double radius, const Ray &ray, double t_min,
                                           double t_max, double *t_out) {
  V3 d = ray.d;
  double r2 = radius * radius;
  V3 f = center - ray.o;
  double bp = dot(f, d);
  double a = quadrance(d);
  double discrim = r2 - quadrance(scale(d, bp / a) - f);
  if (discrim < 0.0) return false;
  double sign_bp = (bp >= 0.0) ? 1.0 : -1.0;
  double q = std::fma(sign_bp, std::sqrt(a * discrim), bp);
  double c = quadrance(f) - r2;
  double t_hit = (c > 0.0) ? c / q : q / a;
  if (t_min <= t_hit && t_hit <= t_max) {
    *t_out = t_hit;
    return true;
  }
  return false;
}

// spheres_intersect_aux, AVX2 flavour (sphere-intersect-rs/src/lib.rs:102-178), emulated lane by
// lane with the same IEEE operations (mul/add/fma/div/sqrt are exactly rounded in both).
static const int LEAF_SIZE = 16;  // lib.rs:13
static inline int spheres_intersect_simd_emul(const double *xs, const double *ys, const double *zs,
                                              const double *rs, int len, V3 o, V3 d, double t_min,
                                              double t_max, double *t_found_out) {
  double t_hits[LEAF_SIZE];
  for (int i = 0; i < LEAF_SIZE; ++i) t_hits[i] = 0.0;  // lib.rs:114
  double a = d.x * d.x + d.y * d.y + d.z * d.z;          // lib.rs:38-40,115 (no fma)
  double one_over_a = 1.0 / a;
  int chunks = std::min(len / 4, LEAF_SIZE / 4);
  for (int i = 0; i < chunks * 4; ++i) {
    double fx = xs[i] - o.x, fy = ys[i] - o.y, fz = zs[i] - o.z;
    double r2 = rs[i] * rs[i];
    double c = std::fma(fx, fx, std::fma(fy, fy, fz * fz)) - r2;
    double bp = std::fma(fx, d.x, std::fma(fy, d.y, fz * d.z));
    double bp_over_a = bp * one_over_a;
    double wx = std::fma(d.x, bp_over_a, -fx);
    double wy = std::fma(d.y, bp_over_a, -fy);
    double wz = std::fma(d.z, bp_over_a, -fz);
    double wq = std::fma(wx, wx, std::fma(wy, wy, wz * wz));
    double disc = r2 - wq;
    double q_rhs = std::sqrt(a * disc);
    double q = std::signbit(bp) ? (bp - q_rhs) : (bp + q_rhs);  // blendv on sign bit of bp
    double c_div_q = c / q;
    double q_div_a = q * one_over_a;
    double t_hit = std::signbit(c) ? q_div_a : c_div_q;  // blendv on sign bit of c
    bool outside = (t_hit < t_min) || (t_hit > t_max);   // _CMP_LT_OQ / _CMP_GT_OQ
    if (std::signbit(disc) || outside) t_hit = NAN;       // blendv on (disc | outside)
    t_hits[i] = t_hit;
  }
  double t_found = t_max;
  int found = -1;
  int n = std::min(len, LEAF_SIZE);
  for (int i = 0; i < n; ++i) {  // lib.rs:169-176: `<=` so a later equal t wins
    if (t_hits[i] <= t_found) {
      t_found = t_hits[i];
      found = i;
    }
  }
  *t_found_out = t_found;
  return found;
}
#ifdef ORC_HAVE_AVX2
// The same kernel with the intrinsics the Rust crate uses; this is what the timed CPU baseline runs.
static inline int spheres_intersect_simd_avx(const double *xs, const double *ys, const double *zs,
                                             const double *rs, int len, V3 o, V3 d, double t_min,
                                             double t_max, double *t_found_out) {
  alignas(32) double t_hits[LEAF_SIZE] = {0};
  double dq = d.x * d.x + d.y * d.y + d.z * d.z;
  __m256d a = _mm256_set1_pd(dq);
  __m256d one_over_a = _mm256_div_pd(_mm256_set1_pd(1.0), a);
  __m256d ox = _mm256_set1_pd(o.x), oy = _mm256_set1_pd(o.y), oz = _mm256_set1_pd(o.z);
  __m256d dx = _mm256_set1_pd(d.x), dy = _mm256_set1_pd(d.y), dz = _mm256_set1_pd(d.z);
  int chunks = std::min(len / 4, LEAF_SIZE / 4);
  for (int k = 0; k < chunks; ++k) {
    __m256d x = _mm256_loadu_pd(xs + 4 * k), y = _mm256_loadu_pd(ys + 4 * k);
    __m256d z = _mm256_loadu_pd(zs + 4 * k), r = _mm256_loadu_pd(rs + 4 * k);
    __m256d fx = _mm256_sub_pd(x, ox), fy = _mm256_sub_pd(y, oy), fz = _mm256_sub_pd(z, oz);
    __m256d r2 = _mm256_mul_pd(r, r);
    auto dot4 = [](__m256d vx, __m256d vy, __m256d vz, __m256d wx, __m256d wy, __m256d wz) {
      return _mm256_fmadd_pd(vx, wx, _mm256_fmadd_pd(vy, wy, _mm256_mul_pd(vz, wz)));
    };
    __m256d c = _mm256_sub_pd(dot4(fx, fy, fz, fx, fy, fz), r2);
    __m256d bp = dot4(fx, fy, fz, dx, dy, dz);
    __m256d bp_over_a = _mm256_mul_pd(bp, one_over_a);
    __m256d wx = _mm256_fmsub_pd(dx, bp_over_a, fx);
    __m256d wy = _mm256_fmsub_pd(dy, bp_over_a, fy);
    __m256d wz = _mm256_fmsub_pd(dz, bp_over_a, fz);
    __m256d wq = dot4(wx, wy, wz, wx, wy, wz);
    __m256d disc = _mm256_sub_pd(r2, wq);
    __m256d q_rhs = _mm256_sqrt_pd(_mm256_mul_pd(a, disc));
    __m256d q = _mm256_blendv_pd(_mm256_add_pd(bp, q_rhs), _mm256_sub_pd(bp, q_rhs), bp);
    __m256d c_div_q = _mm256_div_pd(c, q);
    __m256d q_div_a = _mm256_mul_pd(q, one_over_a);
    __m256d t_hit = _mm256_blendv_pd(c_div_q, q_div_a, c);
    __m256d outside = _mm256_or_pd(_mm256_cmp_pd(t_hit, _mm256_set1_pd(t_min), _CMP_LT_OQ),
                                   _mm256_cmp_pd(t_hit, _mm256_set1_pd(t_max), _CMP_GT_OQ));
    t_hit = _mm256_blendv_pd(t_hit, _mm256_set1_pd(NAN), _mm256_or_pd(disc, outside));
    _mm256_store_pd(t_hits + 4 * k, t_hit);
  }
  double t_found = t_max;
  int found = -1;
  int n = std::min(len, LEAF_SIZE);
  for (int i = 0; i < n; ++i) {
    if (t_hits[i] <= t_found) {
      t_found = t_hits[i];
      found = i;
    }
  }
  *t_found_out = t_found;
  return found;
}
#endif
static inline int spheres_intersect_simd(const double *xs, const double *ys, const double *zs,
                                         const double *rs, int len, V3 o, V3 d, double t_min,
                                         double t_max, double *t_found_out) {
#ifdef ORC_HAVE_AVX2
  return spheres_intersect_simd_avx(xs, ys, zs, rs, len, o, d, t_min, t_max, t_found_out);
#else
  return spheres_intersect_simd_emul(xs, ys, zs, rs, len, o, d, t_min, t_max, t_found_out);
#endif
}

// ---------------------------------------------------------------------------------------------
// triangle.ml (triangle/triangle.ml:1-99)
// ---------------------------------------------------------------------------------------------
static inline Bbox tri_bbox(V3 a, V3 b, V3 c) {  // :66-72
  auto lo = [](V3 p, V3 q) { return V3{bmin(p.x, q.x), bmin(p.y, q.y), bmin(p.z, q.z)}; };
  auto hi = [](V3 p, V3 q) { return V3{bmax(p.x, q.x), bmax(p.y, q.y), bmax(p.z, q.z)}; };
  return {lo(lo(a, b), c), hi(hi(a, b), c)};
}
// Triangle.intersect :74-98
static inline bool tri_intersect(V3 a, V3 b, V3 c, const Ray &r, double t_min, double t_max,
                                 double *t_out, double *u_out, double *v_out) {
  const double epsilon = 1e-6;
  V3 e1 = b - a;
  V3 e2 = c - a;
  V3 dir = r.d;
  V3 pvec = cross(dir, e2);
  double det = dot(e1, pvec);
  if (std::fabs(det) < epsilon) return false;
  double det_inv = 1.0 / det;
  V3 tvec = r.o - a;
  double u = det_inv * dot(tvec, pvec);
  V3 qvec = cross(tvec, e1);
  double v = det_inv * dot(dir, qvec);
  if (0.0 <= u && u <= 1.0 && 0.0 <= v && u + v <= 1.0) {
    double t_hit = det_inv * dot(e2, qvec);
    if (t_min <= t_hit && t_hit <= t_max) {
      *t_out = t_hit;
      *u_out = u;
      *v_out = v;
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// hit.ml — Hit.t built by Sphere.hit (sphere.ml:56-69) / Triangle.Hit.to_hit (triangle.ml:43-64)
// ---------------------------------------------------------------------------------------------
struct Hit {
  ShaderSpace ss;
  V3 emit;
  int mat;
  TexCoord tc;
  V3 omega_i;
  bool hit_front;
};
static inline TexCoord sphere_tex_coord(V3 n) {  // sphere.ml:22-33
  const double one_over_pi = 1.0 / M_PI;
  const double one_over_two_pi = 1.0 / (2.0 * M_PI);
  double theta = std::acos(-n.y);
  double phi = M_PI + std::atan2(-n.z, n.x);
  return {phi * one_over_two_pi, theta * one_over_pi};
}
static Hit sphere_hit(const Scene &sc, const Sphere &s, double t_hit, const Ray &ray) {  // sphere.ml:56-69
  V3 point = point_at(ray, t_hit);
  V3 normal = normalize(point - s.c);  // sphere.ml:21
  bool hit_front = dot(ray.d, normal) < 0.0;
  if (!hit_front) normal = neg(normal);
  Hit h;
  h.tc = sphere_tex_coord(normal);
  h.ss = ss_create(normal, point);
  h.mat = s.mat;
  h.emit = material_emit(sc, s.mat, h.tc);  // sphere.ml:65 Material.emit (material.ml:59: black; EXT: Emissive)
  h.omega_i = ss_omega_i(h.ss, ray);
  h.hit_front = hit_front;
  return h;
}
static Hit tri_hit(const Scene &sc, const Tri &t, double u, double v, const Ray &r) {
  V3 a = sc.verts[t.a], b = sc.verts[t.b], c = sc.verts[t.c];
  V3 g_normal = normalize(cross(b - a, c - a));  // triangle.ml:18-23
  double w = 1.0 - u - v;                        // triangle.ml:25-32
  V3 pt = (scale(a, w) + scale(b, u)) + scale(c, v);
  double w2 = 1.0 - u - v;  // triangle.ml:49-54
  TexCoord tc;
  tc.u = ((t.ta.u * w2) + (t.tb.u * u)) + (t.tc.u * v);
  tc.v = ((t.ta.v * w2) + (t.tb.v * u)) + (t.tc.v * v);
  bool hit_front = dot(r.d, g_normal) < 0.0;
  V3 normal = hit_front ? g_normal : neg(g_normal);
  Hit h;
  h.ss = ss_create(normal, pt);
  h.omega_i = ss_omega_i(h.ss, r);
  h.mat = t.mat;
  h.tc = tc;
  h.emit = material_emit(sc, t.mat, tc);  // triangle.ml:63 writes Color.black; EXT: routed through Material.emit like sphere.ml:65
  h.hit_front = hit_front;
  return h;
}

// ---------------------------------------------------------------------------------------------
// shape_tree.ml (path_tracer/src/shape_tree.ml:1-312) + slice.ml:67-80
// ---------------------------------------------------------------------------------------------
struct Bshape {  // :4-25
  int prim;
  Bbox bbox;
  V3 centroid;
};
struct OptBox {
  bool some = false;
  Bbox b;
};
static inline OptBox union_opt(const OptBox &o, const OptBox &p) {  // bbox.ml:20-24
  if (!o.some) return p;
  if (!p.some) return o;
  return {true, bbox_union(o.b, p.b)};
}
struct Proposal {  // :72-80
  double cost;
  int split_index;
  int axis;
  double scale, cb_min;  // together = the `on_lhs` closure: to_bin b <= split_index
  Bbox lhs_box, rhs_box;
};
// OCaml Float.compare: nan equals nan and is below every other float
static inline int ocaml_float_compare(double a, double b) {
  bool na = std::isnan(a), nb = std::isnan(b);
  if (na || nb) return na && nb ? 0 : (na ? -1 : 1);
  return a < b ? -1 : (a > b ? 1 : 0);
}

struct Builder {
  Scene &sc;
  std::vector<Bshape> shapes;
  int num_bins = 32;
  explicit Builder(Scene &s) : sc(s) {}

  Bbox elt_bbox(int prim) const {
    int nS = (int)sc.spheres.size();
    if (prim < nS) return sphere_bbox(sc.spheres[prim]);
    const Tri &t = sc.tris[prim - nS];
    return tri_bbox(sc.verts[t.a], sc.verts[t.b], sc.verts[t.c]);
  }
  static int to_bin(const Proposal &p, const Bshape &b) {  // :133
    return (int)(p.scale * (axis_of(b.centroid, p.axis) - p.cb_min));
  }
  // Proposal.propose_split_one_axis (:123-139) + candidates (:91-119)
  bool propose_axis(int lo, int len, int axis, const Bbox &cbbox, Proposal *out) const {
    const double epsilon = 1e-6;
    double cb_min = axis_of(cbbox.mn, axis), cb_max = axis_of(cbbox.mx, axis);
    double scale = (double)num_bins * (1.0 - epsilon) / (cb_max - cb_min);
    if (!std::isfinite(scale)) return false;
    struct Bin {
      int count = 0;
      OptBox bounds, l, r;
    };
    std::vector<Bin> bins(num_bins);
    for (int i = lo; i < lo + len; ++i) {  // Bin.insert :41-51
      const Bshape &s = shapes[i];
      int b = (int)(scale * (axis_of(s.centroid, axis) - cb_min));
      Bin &bin = bins[b];
      bin.bounds = bin.bounds.some ? OptBox{true, bbox_union(bin.bounds.b, s.bbox)}
                                   : OptBox{true, s.bbox};
      bin.count++;
    }
    bins[num_bins - 1].r = bins[num_bins - 1].bounds;  // populate_bbox_r :53-60
    for (int j = num_bins - 2; j >= 0; --j) bins[j].r = union_opt(bins[j].bounds, bins[j + 1].r);
    bins[0].l = bins[0].bounds;  // populate_bbox_l :62-69
    for (int j = 1; j < num_bins; ++j) bins[j].l = union_opt(bins[j].bounds, bins[j - 1].l);
    // candidates
    const double costI = 1.0, costT = 0.25;
    double total_area = surface_area(bins[num_bins - 1].l.b);
    int total_count = 0;
    for (auto &b : bins) total_count += b.count;
    // The reference conses candidates (so the list runs from the highest p down) and List.min_elt
    // keeps the FIRST minimum: on equal cost the highest split index wins.
    bool have = false;
    Proposal best{};
    int n_left = 0;
    std::vector<Proposal> cands;
    for (int p = 0; p < num_bins - 1; ++p) {
      int lhs_count = n_left + bins[p].count;
      int rhs_count = total_count - lhs_count;
      n_left = lhs_count;
      if (!bins[p].l.some || !bins[p + 1].r.some) continue;
      double lhs_area = (double)lhs_count * surface_area(bins[p].l.b);
      double rhs_area = (double)rhs_count * surface_area(bins[p + 1].r.b);
      double cost = costT + ((lhs_area + rhs_area) * costI / total_area);
      cands.push_back({cost, p, axis, scale, cb_min, bins[p].l.b, bins[p + 1].r.b});
    }
    for (int k = (int)cands.size() - 1; k >= 0; --k) {
      if (!have || ocaml_float_compare(best.cost, cands[k].cost) > 0) {
        best = cands[k];
        have = true;
      }
    }
    if (have) *out = best;
    return have;
  }
  // Proposal.create (:141-146)
  bool propose(int lo, int len, Proposal *out) const {
    Bbox cb{shapes[lo].centroid, shapes[lo].centroid};  // Bshape.centroid_bbox :21-24
    for (int i = lo + 1; i < lo + len; ++i)
      cb = bbox_union(cb, Bbox{shapes[i].centroid, shapes[i].centroid});
    bool have = false;
    Proposal best{};
    for (int axis = 0; axis < 3; ++axis) {  // Axis.all = [X;Y;Z]; first minimum wins
      Proposal p{};
      if (!propose_axis(lo, len, axis, cb, &p)) continue;
      if (!have || ocaml_float_compare(best.cost, p.cost) > 0) {
        best = p;
        have = true;
      }
    }
    if (have) *out = best;
    return have;
  }
  int make_leaf(const Bbox &bbox, int lo, int len) {  // :173-175
    Leaf L;
    for (int i = lo; i < lo + len; ++i) L.prims.push_back(shapes[i].prim);
    if (sc.leaf_kind == ORC_LEAF_SIMD) {  // Simd_leaf.of_elts (main.ml:177-193)
      int rem = len % 4;
      int padded = len + (rem == 0 ? 0 : 4 - rem);
      L.xs.assign(padded, NAN);
      L.ys.assign(padded, NAN);
      L.zs.assign(padded, NAN);
      L.rs.assign(padded, NAN);
      for (int i = 0; i < len; ++i) {
        const Sphere &s = sc.spheres[L.prims[i]];
        L.xs[i] = s.c.x;
        L.ys[i] = s.c.y;
        L.zs[i] = s.c.z;
        L.rs[i] = s.r;
      }
    }
    sc.leaves.push_back(std::move(L));
    Node n;
    n.box = bbox;
    n.leaf = (int)sc.leaves.size() - 1;
    sc.nodes.push_back(n);
    return (int)sc.nodes.size() - 1;
  }
  // Slice.partition_in_place (slice.ml:67-80); returns the split position
  int partition(int lo, int len, const Proposal &p) {
    auto on_lhs = [&](int i) { return to_bin(p, shapes[lo + i]) <= p.split_index; };
    int i = 0, j = len - 1;
    while (i < j) {
      while (on_lhs(i) && i < j) ++i;
      while (j >= 0 && !on_lhs(j)) --j;  // reference evaluates on_lhs first; j>=0 always here
      if (i < j) std::swap(shapes[lo + i], shapes[lo + j]);
    }
    return i;
  }
  // Tree.create (:177-196)
  int build(const Bbox &bbox, int lo, int len) {
    Proposal p{};
    if (!propose(lo, len, &p)) return make_leaf(bbox, lo, len);
    double leaf_cost = 1.0 * (double)len;
    if ((p.cost >= leaf_cost && len <= sc.length_cutoff) || len <= 4) return make_leaf(bbox, lo, len);
    int split = partition(lo, len, p);
    int lhs = build(p.lhs_box, lo, split);
    int rhs = build(p.rhs_box, lo + split, len - split);
    Node n;
    n.box = bbox;
    n.axis = p.axis;
    n.lhs = lhs;
    n.rhs = rhs;
    sc.nodes.push_back(n);
    return (int)sc.nodes.size() - 1;
  }
};

// Leaf.intersect
static inline bool leaf_intersect(const Scene &sc, const Leaf &L, const Ray &ray, double t_min,
                                  double t_max, EltHit *out, orc_counters *cn) {
  cn->leaf_visits++;
  if (sc.leaf_kind == ORC_LEAF_SIMD) {  // Simd_leaf.intersect (main.ml:206-217)
    double t;
    cn->sphere_tests += L.prims.size();
    int idx = spheres_intersect_simd(L.xs.data(), L.ys.data(), L.zs.data(), L.rs.data(),
                                     (int)L.xs.size(), ray.o, ray.d, t_min, t_max, &t);
    if (idx < 0) return false;
    *out = {t, L.prims[idx], 0, 0};
    return true;
  }
  // Array_leaf.intersect (shape_tree.ml:299-311): shrinking t_max, later element wins ties
  bool found = false;
  int nS = (int)sc.spheres.size();
  for (int prim : L.prims) {
    if (prim < nS) {
      cn->sphere_tests++;
      double t;
      const Sphere &s = sc.spheres[prim];
      if (sphere_intersect_scalar(s.c, s.r, ray, t_min, t_max, &t)) {
        *out = {t, prim, 0, 0};
        t_max = t;
        found = true;
      }
    } else {
      cn->tri_tests++;
      const Tri &tr = sc.tris[prim - nS];
      double t, u, v;
      if (tri_intersect(sc.verts[tr.a], sc.verts[tr.b], sc.verts[tr.c], ray, t_min, t_max, &t, &u,
                        &v)) {
        *out = {t, prim, u, v};
        t_max = t;
        found = true;
      }
    }
  }
  return found;
}

// Tree.intersect (shape_tree.ml:198-220)
static bool tree_intersect_rec(const Scene &sc, int node, const Ray &ray, const bool dirs[3],
                               double t_min, double t_max, EltHit *out, orc_counters *cn) {
  const Node &n = sc.nodes[node];
  cn->box_tests++;
  if (!bbox_is_hit(n.box, ray, t_min, t_max)) return false;
  if (n.leaf >= 0) return leaf_intersect(sc, sc.leaves[n.leaf], ray, t_min, t_max, out, cn);
  int t1 = dirs[n.axis] ? n.lhs : n.rhs;
  int t2 = dirs[n.axis] ? n.rhs : n.lhs;
  EltHit h1;
  if (!tree_intersect_rec(sc, t1, ray, dirs, t_min, t_max, &h1, cn))
    return tree_intersect_rec(sc, t2, ray, dirs, t_min, t_max, out, cn);
  EltHit h2;
  if (tree_intersect_rec(sc, t2, ray, dirs, t_min, h1.t, &h2, cn)) {
    *out = h2;
    return true;
  }
  *out = h1;
  return true;
}
static inline bool tree_intersect(const Scene &sc, const Ray &ray, double t_min, double t_max,
                                  EltHit *out, orc_counters *cn) {
  bool dirs[3] = {ray.d.x >= 0.0, ray.d.y >= 0.0, ray.d.z >= 0.0};
  return tree_intersect_rec(sc, sc.root, ray, dirs, t_min, t_max, out, cn);
}

// Scene.intersect (shirley main.ml:273-277; cornell main.ml:230-234)
static inline bool scene_intersect(const Scene &sc, const Ray &ray, Hit *hit, orc_counters *cn) {
  EltHit e;
  cn->rays++;
  if (!tree_intersect(sc, ray, 0.0, DBL_MAX, &e, cn)) return false;
  cn->hits++;
  int nS = (int)sc.spheres.size();
  if (e.prim < nS)
    *hit = sphere_hit(sc, sc.spheres[e.prim], e.t, ray);
  else
    *hit = tri_hit(sc, sc.tris[e.prim - nS], e.u, e.v, ray);
  return true;
}
// Scene.background (shirley main.ml:104-110)
static inline V3 background(const Scene &sc, const Ray &ray) {
  if (sc.bg_kind == PTB_BG_CONSTANT) return sc.bg0;
  V3 d = normalize(ray.d);
  double t = 0.5 * (dot(d, v3(0, 1, 0)) + 1.0);
  return lerp(t, sc.bg0, sc.bg1);
}

// ---------------------------------------------------------------------------------------------
// camera.ml (path_tracer/src/camera.ml:1-102)
// ---------------------------------------------------------------------------------------------
struct Camera {
  double llx, lly, vx, vy;
  double look_at[16];
};
static Camera camera_create(V3 eye, V3 target, V3 up, double aspect, double vfov_deg) {  // :58-83
  Camera c;
  double rad = vfov_deg * M_PI / 180.0;  // to_radians :56
  double half_height = std::tan(0.5 * rad);
  double half_width = aspect * half_height;
  c.llx = -half_width;
  c.lly = -half_height;
  c.vx = 2.0 * half_width;
  c.vy = 2.0 * half_height;
  // Mat4.look_at :14-27
  V3 zp = normalize(target - eye);
  V3 xp = normalize(cross(zp, normalize(up)));
  V3 yp = normalize(cross(xp, zp));
  auto e = [&](V3 v, double *row) {
    row[0] = v.x, row[1] = v.y, row[2] = v.z, row[3] = -(dot(eye, v));
  };
  auto ep = [&](V3 v, double *row) {
    row[0] = -v.x, row[1] = -v.y, row[2] = -v.z, row[3] = dot(eye, v);
  };
  e(xp, c.look_at + 0);
  e(yp, c.look_at + 4);
  ep(zp, c.look_at + 8);
  c.look_at[12] = 0, c.look_at[13] = 0, c.look_at[14] = 0, c.look_at[15] = 1;
  return c;
}
static inline V3 camera_transform(const double *m, V3 p) {  // Mat4.transform :39-43 ; dot4 :9-12
  auto dot4 = [&](const double *r) { return (((p.x * r[0]) + (p.y * r[1])) + (p.z * r[2])) + (1.0 * r[3]); };
  double x = dot4(m + 0), y = dot4(m + 4), z = dot4(m + 8), w = dot4(m + 12);
  return scale(v3(x, y, z), 1.0 / w);
}
static inline Ray camera_ray(double llx, double lly, double vx, double vy, double dx, double dy) {
  V3 dir = normalize(v3(llx + (vx * dx), lly + (vy * dy), -1.0));  // :93-102
  return ray_create(v3(0, 0, 0), dir);
}

// ---------------------------------------------------------------------------------------------
// tile.ml / filter_kernel.ml / film_tile.ml
// ---------------------------------------------------------------------------------------------
struct Tile {
  int row, col, width, height;
};
static void tile_split(Tile t, int max_area, std::vector<Tile> &out) {  // tile.ml:14-39
  if (t.width * t.height <= max_area) {
    out.push_back(t);
    return;
  }
  Tile lhs = t, rhs = t;
  if (t.width > t.height) {
    int half = t.width / 2;
    lhs.width = half;
    rhs.col = t.col + half;
    rhs.width = t.width - half;
  } else {
    int half = t.height / 2;
    lhs.height = half;
    rhs.row = t.row + half;
    rhs.height = t.height - half;
  }
  tile_split(lhs, max_area, out);
  tile_split(rhs, max_area, out);
}

struct Rat {  // the exact rationals of `num` that Binomial.create needs
  long long n, d;
  static long long g(long long a, long long b) { return b == 0 ? (a < 0 ? -a : a) : g(b, a % b); }
  Rat(long long nn = 0, long long dd = 1) {
    long long k = g(nn, dd);
    if (k == 0) k = 1;
    if (dd < 0) k = -k;
    n = nn / k;
    d = dd / k;
  }
  Rat operator+(Rat o) const { return Rat(n * o.d + o.n * d, d * o.d); }
  Rat operator-(Rat o) const { return Rat(n * o.d - o.n * d, d * o.d); }
  Rat operator*(Rat o) const { return Rat(n * o.n, d * o.d); }
  long long floor() const { return n >= 0 ? n / d : -((-n + d - 1) / d); }
  long long ceil() const { return -Rat(-n, d).floor(); }
  Rat frac() const { return *this - Rat(floor(), 1); }  // mod_num n 1
  double to_double() const { return (double)n / (double)d; }
};
static std::vector<double> filter_binomial(int order, int pixel_radius) {  // filter_kernel.ml:49-85
  auto pow_falling = [](long long n, long long k) {
    long long r = 1;
    for (long long i = 0; i < k; ++i) r *= (n - i);
    return r;
  };
  auto binomial = [&](int n, int k) { return pow_falling(n, k) / pow_falling(k, k); };
  int f_width = 1 + 2 * pixel_radius;
  Rat ratio(order, f_width);
  std::vector<long long> coeffs(order);
  for (int k = 0; k < order; ++k) coeffs[k] = binomial(order - 1, k);
  std::vector<double> w(f_width);
  for (int i = 0; i < f_width; ++i) {
    Rat ip = Rat(i, 1) * ratio;
    Rat jp = ip + ratio;
    long long beg = ip.floor();
    long long end_ = jp.ceil();
    int len = (int)(end_ - beg);
    Rat sum(0, 1);
    for (int k = 0; k < len; ++k) {
      Rat weight = (k == 0) ? (Rat(1, 1) - ip.frac())
                            : (k == len - 1 ? (Rat(1, 1) - (Rat(end_, 1) - jp)) : Rat(1, 1));
      sum = sum + weight * Rat(coeffs[k + beg], 1);
    }
    w[i] = sum.to_double();
  }
  double total = 0.0;
  for (double x : w) total = total + x;
  for (double &x : w) x = x / total;
  std::vector<double> data((size_t)f_width * f_width);
  for (int j = 0; j < f_width * f_width; ++j) data[j] = w[j / f_width] * w[j % f_width];
  return data;
}

struct FilmTile {  // film_tile.ml:6-21
  Tile tile;
  int border, w, h;
  std::vector<double> px;  // (w*h*3), interleaved rgb
  const std::vector<double> *fk;
  FilmTile(Tile t, const std::vector<double> *k, int radius)
      : tile(t), border(radius), w(t.width + 2 * radius), h(t.height + 2 * radius),
        px((size_t)w * h * 3, 0.0), fk(k) {}
  void write_pixel(int x, int y, V3 color) {  // :23-38
    x += border;
    y += border;
    int i = 0;
    for (int dy = -border; dy <= border; ++dy)  // Filter_kernel.iter (filter_kernel.ml:14-24)
      for (int dx = -border; dx <= border; ++dx) {
        double weight = (*fk)[i++];
        double *p = &px[((size_t)(y + dy) * w + (x + dx)) * 3];
        p[0] = std::fma(weight, color.x, p[0]);
        p[1] = std::fma(weight, color.y, p[1]);
        p[2] = std::fma(weight, color.z, p[2]);
      }
  }
  void write_sample(double x, double y, V3 color) {  // :40-45
    write_pixel((int)x, (int)y, color);
  }
};

// ---------------------------------------------------------------------------------------------
// integrator.ml (path_tracer/src/integrator.ml:1-156)
// ---------------------------------------------------------------------------------------------
struct Integrator {
  const Scene &sc;
  ptb_params p;
  std::vector<double> alpha;  // create_sampler :89
  Integrator(const Scene &s, const ptb_params &pp) : sc(s), p(pp) {
    int D = 2 + 2 * p.max_bounces;
    alpha.resize(D);
    lds_alpha(D, alpha.data());
  }
  // path_tracer (:16-69)
  V3 trace_path(double cx, double cy, int64_t offset, orc_counters *cn) const {
    Ray ray = camera_ray(p.lower_left_x, p.lower_left_y, p.view_x, p.view_y, cx, cy);
    int samples_index = 2;
    int max_bounces = p.max_bounces;
    V3 emit0 = {0, 0, 0}, attn0 = {1, 1, 1};
    const V3 black = {0, 0, 0};
    auto add_mul = [](V3 a, V3 b, V3 c) { return vfma(b, c, a); };  // :29
    int bounce = 0;
    for (;;) {
      if (max_bounces <= 0) {
        cn->exhausted++;
        return add_mul(emit0, attn0, black);
      }
      max_bounces -= 1;
      Hit h;
      if (bounce < 64) cn->rays_by_bounce[bounce]++;
      bounce++;
      if (!scene_intersect(sc, ray, &h, cn)) {
        cn->missed++;
        return add_mul(emit0, attn0, background(sc, ray));
      }
      V3 emit = h.emit;
      int j = samples_index;  // take_2d :20-28
      double u = lds_get(alpha.data(), offset, j);
      double v = lds_get(alpha.data(), offset, j + 1);
      samples_index = j + 2;
      Scatter s = material_scatter(sc, h.mat, h.ss, h.tc, h.omega_i, h.hit_front, u);
      if (s.kind == ABSORB) {
        cn->absorbed++;
        return add_mul(emit0, attn0, emit);
      } else if (s.kind == SPECULAR) {
        if (sc.mat[h.mat].kind == PTB_MAT_METAL) cn->scatter_metal++; else cn->scatter_dielectric++;
        ray = s.ray;
        emit0 = add_mul(emit, s.attenuation, emit0);
        attn0 = s.attenuation * attn0;
      } else {
        cn->scatter_lambert++;
        // Pdf.sample diffuse_plus_light ss u v (pdf.ml:5-9); diffuse_plus_light = Pdf.diffuse (render_command.ml:81)
        // unless the scene carries a light (EXT: Pdf.Mix)
        V3 dir = sc.has_light ? pdf_sample_mix(sc, h.ss, u, v) : unit_square_to_hemisphere(u, v);
        double diffuse_pd = pdf_eval_diffuse(dir);
        if (diffuse_pd == 0.0) {
          cn->absorbed++;
          return add_mul(emit0, attn0, emit);
        }
        double divisor = sc.has_light ? pdf_eval_mix(sc, dir, h.ss) : pdf_eval_diffuse(dir);
        double pd = diffuse_pd / divisor;
        if (!std::isfinite(pd)) {
          cn->absorbed++;
          return add_mul(emit0, attn0, emit);
        }
        Ray scattered = ss_world_ray(h.ss, dir);
        V3 attenuation = scale(s.attenuation, pd);
        ray = scattered;
        emit0 = add_mul(emit, attenuation, emit0);
        attn0 = attenuation * attn0;
      }
    }
  }
  // the body of render_tile's Tile.iter closure (:97-109), minus the film write
  V3 sample(int gx, int gy, int pass, double *dx_out, double *dy_out, orc_counters *cn) const {
    double widthf = 1.0 / (double)p.width, heightf = 1.0 / (double)p.height;
    int64_t offset = ((int64_t)gy * p.width) + gx + ((int64_t)pass * p.samples_per_pixel);
    double xf = (double)gx, yf = (double)gy;
    double dx = lds_get(alpha.data(), offset, 0), dy = lds_get(alpha.data(), offset, 1);
    double cx = (xf + dx) * widthf;
    double cy = 1.0 - ((yf + dy) * heightf);
    cn->paths++;
    *dx_out = dx;
    *dy_out = dy;
    return trace_path(cx, cy, offset, cn);
  }
  // render_tile (:91-112)
  std::unique_ptr<FilmTile> render_tile(Tile tile, const std::vector<double> *fk, int passes,
                                        orc_counters *cn) const {
    auto ft = std::make_unique<FilmTile>(tile, fk, 1);
    for (int pass = 0; pass < passes; ++pass)
      for (int ly = 0; ly < tile.height; ++ly)
        for (int lx = 0; lx < tile.width; ++lx) {
          double dx, dy;
          V3 color = sample(lx + tile.col, ly + tile.row, pass, &dx, &dy, cn);
          ft->write_sample((double)lx + dx, (double)ly + dy, color);
        }
    return ft;
  }
};
// stitch_tile (:114-128)
static void stitch_tile(double *img, int W, int H, const FilmTile &ft) {
  for (int ly = 0; ly < ft.h; ++ly) {
    int gy = ly + ft.tile.row - ft.border;
    for (int lx = 0; lx < ft.w; ++lx) {
      int gx = lx + ft.tile.col - ft.border;
      if (0 <= gx && gx < W && 0 <= gy && gy < H) {
        const double *s = &ft.px[((size_t)ly * ft.w + lx) * 3];
        double *d = &img[((size_t)gy * W + gx) * 3];
        d[0] = s[0] + d[0];
        d[1] = s[1] + d[1];
        d[2] = s[2] + d[2];
      }
    }
  }
}
static void add_counters(orc_counters *dst, const orc_counters &src) {
  const uint64_t *s = reinterpret_cast<const uint64_t *>(&src);
  uint64_t *d = reinterpret_cast<uint64_t *>(dst);
  for (size_t i = 0; i < sizeof(orc_counters) / sizeof(uint64_t); ++i) d[i] += s[i];
}

// Integrator.render (:130-156)
static int render(Scene &sc, const ptb_params &p, double *image, orc_counters *out_cn, int n_threads,
                  int pass_limit) {
  Integrator integ(sc, p);
  const int W = p.width, H = p.height;
  std::vector<Tile> all_tiles, tiles;
  tile_split({0, 0, W, H}, 32 * 32, all_tiles);
  int world = p.tile_world > 0 ? p.tile_world : 1;
  for (size_t i = 0; i < all_tiles.size(); ++i)
    if ((int)(i % world) == p.tile_rank) tiles.push_back(all_tiles[i]);
  std::vector<double> fk = filter_binomial(5, 1);
  const bool no_filter = (p.flags & PTB_FLAG_NO_FILTER) != 0;
  std::vector<double> delta = {0, 0, 0, 0, 1, 0, 0, 0, 0};
  const std::vector<double> *kernel = no_filter ? &delta : &fk;
  int passes = pass_limit > 0 ? std::min(pass_limit, p.samples_per_pixel) : p.samples_per_pixel;
  std::fill(image, image + (size_t)W * H * 3, 0.0);
  orc_counters total;
  std::memset(&total, 0, sizeof total);
  if (n_threads <= 1) {
    for (const Tile &t : tiles) {
      auto ft = integ.render_tile(t, kernel, passes, &total);
      stitch_tile(image, W, H, *ft);
    }
  } else {
    // N-1 worker domains + the main domain stitching through a channel (:136-151)
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::unique_ptr<FilmTile>> chan;
    std::atomic<size_t> next{0};
    int workers = std::max(1, n_threads - 1);
    std::vector<std::thread> pool;
    std::vector<orc_counters> cns(workers);
    for (int w = 0; w < workers; ++w) {
      std::memset(&cns[w], 0, sizeof(orc_counters));
      pool.emplace_back([&, w] {
        for (;;) {
          size_t i = next.fetch_add(1);
          if (i >= tiles.size()) break;
          auto ft = integ.render_tile(tiles[i], kernel, passes, &cns[w]);
          {
            std::lock_guard<std::mutex> lk(mu);
            chan.push_back(std::move(ft));
          }
          cv.notify_one();
        }
      });
    }
    for (size_t i = 0; i < tiles.size(); ++i) {
      std::unique_ptr<FilmTile> ft;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !chan.empty(); });
        ft = std::move(chan.front());
        chan.pop_front();
      }
      stitch_tile(image, W, H, *ft);
    }
    for (auto &t : pool) t.join();
    for (auto &c : cns) add_counters(&total, c);
  }
  if (!(p.flags & (PTB_FLAG_RAW_SUMS | PTB_FLAG_NO_FILTER))) {
    double spp_inv = 1.0 / (double)p.samples_per_pixel;  // :152-154
    size_t n = (size_t)W * H * 3;
    for (size_t i = 0; i < n; ++i) image[i] = std::sqrt(image[i] * spp_inv);
  }
  if (out_cn) *out_cn = total;
  return 0;
}

}  // namespace orc

// =================================================================================================
// C API
// =================================================================================================
using namespace orc;
struct orc_scene {
  Scene s;
};
static inline V3 V(const double *p) { return {p[0], p[1], p[2]}; }

extern "C" {

double orc_lds_phi(int dimension) { return phi_approx(dimension); }
void orc_lds_alpha(int dimension, double *alpha) { lds_alpha(dimension, alpha); }
double orc_lds_get(const double *alpha, int64_t offset, int dimension) {
  return lds_get(alpha, offset, dimension);
}
void orc_filter_binomial(int order, int pixel_radius, double *weights) {
  auto w = filter_binomial(order, pixel_radius);
  std::copy(w.begin(), w.end(), weights);
}
int orc_tile_split(int width, int height, int max_area, int32_t *row, int32_t *col, int32_t *w,
                   int32_t *h, int cap) {
  std::vector<Tile> t;
  tile_split({0, 0, width, height}, max_area, t);
  for (int i = 0; i < (int)t.size() && i < cap; ++i) {
    row[i] = t[i].row, col[i] = t[i].col, w[i] = t[i].width, h[i] = t[i].height;
  }
  return (int)t.size();
}
int orc_bbox_is_hit(const double bmin_[3], const double bmax_[3], const double o[3],
                    const double d[3], double t_min, double t_max) {
  return bbox_is_hit(Bbox{V(bmin_), V(bmax_)}, ray_create(V(o), V(d)), t_min, t_max) ? 1 : 0;
}
void orc_camera_create(const double eye[3], const double target[3], const double up[3],
                       double aspect, double vfov_deg, double out[20]) {
  Camera c = camera_create(V(eye), V(target), V(up), aspect, vfov_deg);
  out[0] = c.llx, out[1] = c.lly, out[2] = c.vx, out[3] = c.vy;
  std::copy(c.look_at, c.look_at + 16, out + 4);
}
void orc_camera_transform(const double look_at[16], double *xs, double *ys, double *zs, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    V3 p = camera_transform(look_at, v3(xs[i], ys[i], zs[i]));
    xs[i] = p.x, ys[i] = p.y, zs[i] = p.z;
  }
}
void orc_camera_ray(const double cam4[4], double cx, double cy, double dir[3]) {
  Ray r = camera_ray(cam4[0], cam4[1], cam4[2], cam4[3], cx, cy);
  dir[0] = r.d.x, dir[1] = r.d.y, dir[2] = r.d.z;
}
void orc_unit_square_to_hemisphere(double u, double v, double out[3]) {
  V3 w = unit_square_to_hemisphere(u, v);
  out[0] = w.x, out[1] = w.y, out[2] = w.z;
}
void orc_film_tile_write_pixel(int tile_w, int tile_h, int x, int y, const double rgb[3],
                               double *out) {
  auto fk = filter_binomial(5, 1);
  FilmTile ft({0, 0, tile_w, tile_h}, &fk, 1);
  ft.write_pixel(x, y, V(rgb));
  std::copy(ft.px.begin(), ft.px.end(), out);
}
int orc_sphere_intersect_scalar(const double c[3], double r, const double o[3], const double d[3],
                                double t_min, double t_max, double *t) {
  return sphere_intersect_scalar(V(c), r, ray_create(V(o), V(d)), t_min, t_max, t) ? 1 : 0;
}
int orc_spheres_intersect_simd(const double *xs, const double *ys, const double *zs,
                               const double *rs, int len, const double o[3], const double d[3],
                               double t_min, double t_max, double *t) {
  double t1, t2;
  int i1 = spheres_intersect_simd_emul(xs, ys, zs, rs, len, V(o), V(d), t_min, t_max, &t1);
  int i2 = spheres_intersect_simd(xs, ys, zs, rs, len, V(o), V(d), t_min, t_max, &t2);
  // the intrinsics build and the lane-by-lane emulation must agree bit for bit
  if (i1 != i2 || std::memcmp(&t1, &t2, sizeof t1) != 0) return -2;
  *t = t1;
  return i1;
}
int orc_triangle_intersect(const double a[3], const double b[3], const double c[3],
                           const double o[3], const double d[3], double t_min, double t_max,
                           double *t, double *u, double *v) {
  return tri_intersect(V(a), V(b), V(c), ray_create(V(o), V(d)), t_min, t_max, t, u, v) ? 1 : 0;
}
void orc_shader_space_rotate(const double n[3], const double v[3], int inverse, double out[3]) {
  ShaderSpace ss = ss_create(V(n), v3(0, 0, 0));
  V3 r = inverse ? ss_rotate_inv(ss, V(v)) : ss_rotate(ss, V(v));
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}

orc_scene *orc_scene_create(void) { return new orc_scene(); }
void orc_scene_destroy(orc_scene *s) { delete s; }
void orc_scene_set_textures(orc_scene *s, const ptb_texture *t, int n) {
  s->s.tex.assign(t, t + n);
  s->s.committed = false;
}
void orc_scene_set_materials(orc_scene *s, const ptb_material *m, int n) {
  s->s.mat.assign(m, m + n);
  s->s.committed = false;
}
void orc_scene_set_spheres(orc_scene *s, const double *xs, const double *ys, const double *zs,
                           const double *rs, const int32_t *material, int64_t n) {
  s->s.spheres.clear();
  for (int64_t i = 0; i < n; ++i)
    s->s.spheres.push_back({v3(xs[i], ys[i], zs[i]), rs[i], material ? material[i] : 0});
  s->s.committed = false;
}
void orc_scene_set_triangles(orc_scene *s, const double *vx, const double *vy, const double *vz,
                             int64_t nv, const int32_t *idx, const int32_t *material,
                             const double *uv, int64_t nt) {
  s->s.verts.clear();
  s->s.tris.clear();
  for (int64_t i = 0; i < nv; ++i) s->s.verts.push_back(v3(vx[i], vy[i], vz[i]));
  for (int64_t i = 0; i < nt; ++i) {
    Tri t;
    t.a = idx[3 * i], t.b = idx[3 * i + 1], t.c = idx[3 * i + 2];
    t.mat = material ? material[i] : 0;
    if (uv) {
      t.ta = {uv[6 * i], uv[6 * i + 1]};
      t.tb = {uv[6 * i + 2], uv[6 * i + 3]};
      t.tc = {uv[6 * i + 4], uv[6 * i + 5]};
    } else {  // ganesha: (t00, t01, t11) (ganesha/bin/main.ml:111)
      t.ta = {0, 0}, t.tb = {0, 1}, t.tc = {1, 1};
    }
    s->s.tris.push_back(t);
  }
  s->s.committed = false;
}
void orc_scene_set_light_quad(orc_scene *s, const double origin[3], const double u[3], const double v[3]) {
  s->s.has_light = origin != nullptr;
  if (origin) s->s.light_o = V(origin), s->s.light_u = V(u), s->s.light_v = V(v);
}
void orc_scene_set_background(orc_scene *s, int kind, const double c0[3], const double c1[3]) {
  s->s.bg_kind = kind;
  s->s.bg0 = V(c0);
  if (c1) s->s.bg1 = V(c1);
}
int orc_scene_commit(orc_scene *s, int leaf_kind, int length_cutoff, const int32_t *prim_order,
                     int64_t n_order) {
  Scene &sc = s->s;
  int nS = (int)sc.spheres.size(), nT = (int)sc.tris.size();
  if (nS + nT == 0) return -1;  // Shape_tree.create: expected non-empty list (shape_tree.ml:254-255)
  if (leaf_kind == ORC_LEAF_SIMD && nT > 0) return -2;
  sc.leaf_kind = leaf_kind;
  sc.length_cutoff = length_cutoff;
  sc.nodes.clear();
  sc.leaves.clear();
  Builder b(sc);
  auto push = [&](int prim) {  // Bshape.create :14-18
    Bbox bb = b.elt_bbox(prim);
    b.shapes.push_back({prim, bb, bbox_center(bb)});
  };
  if (prim_order) {
    for (int64_t i = 0; i < n_order; ++i) push(prim_order[i] >= 0 ? prim_order[i] : nS + (~prim_order[i]));
  } else {
    for (int i = 0; i < nS + nT; ++i) push(i);
  }
  Bbox root = b.shapes[0].bbox;  // slice_bbox :258-259
  for (size_t i = 1; i < b.shapes.size(); ++i) root = bbox_union(root, b.shapes[i].bbox);
  sc.root = b.build(root, 0, (int)b.shapes.size());
  sc.committed = true;
  return 0;
}
static int depth_rec(const Scene &sc, int n) {  // Make.depth :239 with L.depth = 0 (Simd) or 1+0 (Array_leaf)
  const Node &nd = sc.nodes[n];
  if (nd.leaf >= 0) return sc.leaf_kind == ORC_LEAF_SIMD ? 0 : 1;
  return 1 + std::max(depth_rec(sc, nd.lhs), depth_rec(sc, nd.rhs));
}
int orc_scene_tree_depth(const orc_scene *s) { return depth_rec(s->s, s->s.root); }
int64_t orc_scene_node_count(const orc_scene *s) { return (int64_t)s->s.nodes.size(); }
int orc_scene_leaf_histogram(const orc_scene *s, int32_t *sizes, int32_t *counts, int cap) {
  std::map<int, int> h;
  for (const Leaf &L : s->s.leaves) {
    // Simd_leaf.length is the PADDED length (main.ml:192,196)
    int len = s->s.leaf_kind == ORC_LEAF_SIMD ? (int)L.xs.size() : (int)L.prims.size();
    h[len]++;
  }
  int i = 0;
  for (auto &kv : h) {
    if (i < cap) sizes[i] = kv.first, counts[i] = kv.second;
    ++i;
  }
  return i;
}

int orc_render(orc_scene *s, const ptb_params *p, double *image, orc_counters *cn, int n_threads,
               int pass_limit) {
  if (!s->s.committed) return -1;
  return render(s->s, *p, image, cn, n_threads, pass_limit);
}
void orc_trace_sample(orc_scene *s, const ptb_params *p, int gx, int gy, int pass, double rgb[3],
                      orc_counters *cn) {
  Integrator integ(s->s, *p);
  orc_counters local;
  std::memset(&local, 0, sizeof local);
  double dx, dy;
  V3 c = integ.sample(gx, gy, pass, &dx, &dy, cn ? cn : &local);
  rgb[0] = c.x, rgb[1] = c.y, rgb[2] = c.z;
}
void orc_intersect_batch(orc_scene *s, const double *o, const double *d, double t_min, double t_max,
                         int64_t n, double *t_hit, int32_t *prim, orc_counters *cn, int n_threads) {
  const Scene &sc = s->s;
  int T = std::max(1, n_threads);
  std::vector<orc_counters> cns(T);
  auto work = [&](int w) {
    std::memset(&cns[w], 0, sizeof(orc_counters));
    int64_t lo = n * w / T, hi = n * (w + 1) / T;
    for (int64_t i = lo; i < hi; ++i) {
      Ray r = ray_create(V(o + 3 * i), V(d + 3 * i));
      EltHit e;
      cns[w].rays++;
      if (tree_intersect(sc, r, t_min, t_max, &e, &cns[w])) {
        t_hit[i] = e.t;
        prim[i] = e.prim;
      } else {
        t_hit[i] = NAN;
        prim[i] = -1;
      }
    }
  };
  if (T == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < T; ++w) pool.emplace_back(work, w);
    for (auto &t : pool) t.join();
  }
  if (cn) {
    std::memset(cn, 0, sizeof *cn);
    for (auto &c : cns) add_counters(cn, c);
  }
}
// The same rays against ONE Array_leaf holding every primitive in set order (Array_leaf.intersect,
// shape_tree.ml:299-311: linear scan, shrinking t_max): the closest hit without any tree — an independent check
// for scenes whose reference tree is too expensive to build in a test (10^7 triangles).  Needs no commit.
void orc_intersect_batch_linear(orc_scene *s, const double *o, const double *d, double t_min, double t_max,
                                int64_t n, double *t_hit, int32_t *prim, int n_threads) {
  Scene &sc = s->s;
  const int saved_kind = sc.leaf_kind;
  sc.leaf_kind = ORC_LEAF_ARRAY;
  Leaf all;
  const int total = (int)(sc.spheres.size() + sc.tris.size());
  all.prims.resize(total);
  std::iota(all.prims.begin(), all.prims.end(), 0);
  int T = std::max(1, n_threads);
  auto work = [&](int w) {
    orc_counters cn;
    std::memset(&cn, 0, sizeof cn);
    for (int64_t i = n * w / T; i < n * (w + 1) / T; ++i) {
      Ray r = ray_create(V(o + 3 * i), V(d + 3 * i));
      EltHit e;
      if (leaf_intersect(sc, all, r, t_min, t_max, &e, &cn)) {
        t_hit[i] = e.t;
        prim[i] = e.prim;
      } else {
        t_hit[i] = NAN;
        prim[i] = -1;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int w = 0; w < T; ++w) pool.emplace_back(work, w);
  for (auto &t : pool) t.join();
  sc.leaf_kind = saved_kind;
}
void orc_first_hit(orc_scene *s, const ptb_params *p, double *t_hit, int32_t *prim, double *cx_out,
                   double *cy_out) {
  const Scene &sc = s->s;
  Integrator integ(sc, *p);
  orc_counters cn;
  std::memset(&cn, 0, sizeof cn);
  double widthf = 1.0 / (double)p->width, heightf = 1.0 / (double)p->height;
  for (int gy = 0; gy < p->height; ++gy)
    for (int gx = 0; gx < p->width; ++gx) {
      int64_t offset = (int64_t)gy * p->width + gx;
      double dx = lds_get(integ.alpha.data(), offset, 0), dy = lds_get(integ.alpha.data(), offset, 1);
      double cx = ((double)gx + dx) * widthf;
      double cy = 1.0 - (((double)gy + dy) * heightf);
      Ray r = camera_ray(p->lower_left_x, p->lower_left_y, p->view_x, p->view_y, cx, cy);
      EltHit e;
      size_t i = (size_t)gy * p->width + gx;
      if (cx_out) cx_out[i] = cx;
      if (cy_out) cy_out[i] = cy;
      if (tree_intersect(sc, r, 0.0, DBL_MAX, &e, &cn)) {
        t_hit[i] = e.t;
        prim[i] = e.prim;
      } else {
        t_hit[i] = NAN;
        prim[i] = -1;
      }
    }
}

}  // extern "C"

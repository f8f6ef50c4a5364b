// scenes.cpp — scene tables: the ptb_scene setters/getters and the data of the reference's three
// scene binaries (see include/ptb200_scenes.h).  Host-only, float64.  Product code.
#include <cmath>
#include <cstring>

#include "../../include/ptb200_scenes.h"
#include "handle.hpp"

using namespace ptb;

// ---------------------------------------------------------------------------------------------
// setters
// ---------------------------------------------------------------------------------------------
extern "C" {

ptb_scene *ptb_scene_create(void) { return new (std::nothrow) ptb_scene(); }
void ptb_scene_destroy(ptb_scene *s) {
  if (!s) return;
  for (ptb_scene *r : s->replicas) ptb_scene_destroy(r);
  if (s->dev) destroy_device_state(s->dev);
  delete s;
}

int ptb_scene_set_textures(ptb_scene *s, const ptb_texture *t, int32_t n) {
  if (!s || (n > 0 && !t) || n < 0) return fail(PTB_E_INVALID, "set_textures: bad args");
  for (int i = 0; i < n; ++i) {
    if (t[i].kind == PTB_TEX_CHECKER) {
      if (t[i].even < 0 || t[i].even >= n || t[i].odd < 0 || t[i].odd >= n ||
          t[t[i].even].kind != PTB_TEX_SOLID || t[t[i].odd].kind != PTB_TEX_SOLID)
        return fail(PTB_E_INVALID, "set_textures: checker rows must reference solid rows");
    } else if (t[i].kind != PTB_TEX_SOLID) {
      return fail(PTB_E_INVALID, "set_textures: unknown texture kind");
    }
  }
  s->host.tex.assign(t, t + n);
  s->committed = false;
  return PTB_OK;
}

int ptb_scene_set_materials(ptb_scene *s, const ptb_material *m, int32_t n) {
  if (!s || (n > 0 && !m) || n < 0) return fail(PTB_E_INVALID, "set_materials: bad args");
  for (int i = 0; i < n; ++i)
    if (m[i].kind < PTB_MAT_LAMBERTIAN || m[i].kind > PTB_MAT_EMISSIVE)
      return fail(PTB_E_INVALID, "set_materials: unknown material kind");
  s->host.mat.assign(m, m + n);
  s->committed = false;
  return PTB_OK;
}

int ptb_scene_set_spheres(ptb_scene *s, const double *xs, const double *ys, const double *zs,
                          const double *rs, const int32_t *material, int64_t n) {
  if (!s || n < 0 || (n > 0 && (!xs || !ys || !zs || !rs)))
    return fail(PTB_E_INVALID, "set_spheres: bad args");
  HostScene &h = s->host;
  h.sx.assign(xs, xs + n);
  h.sy.assign(ys, ys + n);
  h.sz.assign(zs, zs + n);
  h.sr.assign(rs, rs + n);
  if (material)
    h.smat.assign(material, material + n);
  else
    h.smat.assign((size_t)n, 0);
  s->committed = false;
  return PTB_OK;
}

int ptb_scene_set_triangles(ptb_scene *s, const double *vx, const double *vy, const double *vz,
                            int64_t nv, const int32_t *indices, const int32_t *material,
                            const double *uv, int64_t nt) {
  if (!s || nv < 0 || nt < 0 || (nt > 0 && (!vx || !vy || !vz || !indices)))
    return fail(PTB_E_INVALID, "set_triangles: bad args");
  for (int64_t i = 0; i < 3 * nt; ++i)  // ganesha/bin/main.ml:82-84 asserts the same
    if (indices[i] < 0 || indices[i] >= nv) return fail(PTB_E_INVALID, "set_triangles: vertex index out of bounds");
  HostScene &h = s->host;
  h.vx.assign(vx, vx + nv);
  h.vy.assign(vy, vy + nv);
  h.vz.assign(vz, vz + nv);
  h.tidx.assign(indices, indices + 3 * nt);
  if (material)
    h.tmat.assign(material, material + nt);
  else
    h.tmat.assign((size_t)nt, 0);
  h.tuv.resize((size_t)6 * nt);
  for (int64_t i = 0; i < nt; ++i) {
    if (uv) {
      std::memcpy(&h.tuv[6 * i], uv + 6 * i, 6 * sizeof(double));
    } else {  // (t00, t01, t11), ganesha/bin/main.ml:111
      const double d[6] = {0, 0, 0, 1, 1, 1};
      std::memcpy(&h.tuv[6 * i], d, sizeof d);
    }
  }
  s->committed = false;
  return PTB_OK;
}

int ptb_scene_set_background(ptb_scene *s, int32_t kind, const double c0[3], const double c1[3]) {
  if (!s || !c0 || (kind != PTB_BG_CONSTANT && kind != PTB_BG_GRADIENT_Y) ||
      (kind == PTB_BG_GRADIENT_Y && !c1))
    return fail(PTB_E_INVALID, "set_background: bad args");
  s->host.bg_kind = kind;
  for (int i = 0; i < 3; ++i) s->host.bg0[i] = c0[i], s->host.bg1[i] = c1 ? c1[i] : c0[i];
  s->committed = false;
  return PTB_OK;
}

int ptb_scene_set_light_quad(ptb_scene *s, const double origin[3], const double u[3], const double v[3]) {
  if (!s || (origin && (!u || !v))) return fail(PTB_E_INVALID, "set_light_quad: bad args");
  HostScene &h = s->host;
  h.has_light = origin != nullptr;
  if (origin) {
    const double nx = u[1] * v[2] - u[2] * v[1], ny = u[2] * v[0] - u[0] * v[2], nz = u[0] * v[1] - u[1] * v[0];
    if (!(nx * nx + ny * ny + nz * nz > 0.0)) return fail(PTB_E_INVALID, "set_light_quad: degenerate quad");
    for (int i = 0; i < 3; ++i) h.light_o[i] = origin[i], h.light_u[i] = u[i], h.light_v[i] = v[i];
  }
  s->committed = false;
  return PTB_OK;
}
int ptb_scene_get_light_quad(const ptb_scene *s, int32_t *has_light, double origin[3], double u[3], double v[3]) {
  if (!s) return fail(PTB_E_INVALID, "get_light_quad: null scene");
  if (has_light) *has_light = s->host.has_light ? 1 : 0;
  for (int i = 0; i < 3; ++i) {
    if (origin) origin[i] = s->host.light_o[i];
    if (u) u[i] = s->host.light_u[i];
    if (v) v[i] = s->host.light_v[i];
  }
  return PTB_OK;
}

int64_t ptb_scene_primitive_count(const ptb_scene *s) {
  return s ? s->host.n_spheres() + s->host.n_tris() : 0;
}

// ---------------------------------------------------------------------------------------------
// getters
// ---------------------------------------------------------------------------------------------
int ptb_scene_counts(const ptb_scene *s, int64_t *ns, int64_t *nv, int64_t *nt, int32_t *nm,
                     int32_t *ntex) {
  if (!s) return fail(PTB_E_INVALID, "scene_counts: null scene");
  if (ns) *ns = s->host.n_spheres();
  if (nv) *nv = (int64_t)s->host.vx.size();
  if (nt) *nt = s->host.n_tris();
  if (nm) *nm = (int32_t)s->host.mat.size();
  if (ntex) *ntex = (int32_t)s->host.tex.size();
  return PTB_OK;
}
int ptb_scene_get_spheres(const ptb_scene *s, double *xs, double *ys, double *zs, double *rs,
                          int32_t *material) {
  if (!s) return fail(PTB_E_INVALID, "get_spheres: null scene");
  const HostScene &h = s->host;
  size_t n = h.sr.size();
  if (xs) std::memcpy(xs, h.sx.data(), n * 8);
  if (ys) std::memcpy(ys, h.sy.data(), n * 8);
  if (zs) std::memcpy(zs, h.sz.data(), n * 8);
  if (rs) std::memcpy(rs, h.sr.data(), n * 8);
  if (material) std::memcpy(material, h.smat.data(), n * 4);
  return PTB_OK;
}
int ptb_scene_get_triangles(const ptb_scene *s, double *vx, double *vy, double *vz, int32_t *indices,
                            int32_t *material, double *uv) {
  if (!s) return fail(PTB_E_INVALID, "get_triangles: null scene");
  const HostScene &h = s->host;
  if (vx) std::memcpy(vx, h.vx.data(), h.vx.size() * 8);
  if (vy) std::memcpy(vy, h.vy.data(), h.vy.size() * 8);
  if (vz) std::memcpy(vz, h.vz.data(), h.vz.size() * 8);
  if (indices) std::memcpy(indices, h.tidx.data(), h.tidx.size() * 4);
  if (material) std::memcpy(material, h.tmat.data(), h.tmat.size() * 4);
  if (uv) std::memcpy(uv, h.tuv.data(), h.tuv.size() * 8);
  return PTB_OK;
}
int ptb_scene_get_materials(const ptb_scene *s, ptb_material *m, ptb_texture *t) {
  if (!s) return fail(PTB_E_INVALID, "get_materials: null scene");
  if (m) std::memcpy(m, s->host.mat.data(), s->host.mat.size() * sizeof(ptb_material));
  if (t) std::memcpy(t, s->host.tex.data(), s->host.tex.size() * sizeof(ptb_texture));
  return PTB_OK;
}
int ptb_scene_get_background(const ptb_scene *s, int32_t *kind, double c0[3], double c1[3]) {
  if (!s) return fail(PTB_E_INVALID, "get_background: null scene");
  if (kind) *kind = s->host.bg_kind;
  for (int i = 0; i < 3; ++i) {
    if (c0) c0[i] = s->host.bg0[i];
    if (c1) c1[i] = s->host.bg1[i];
  }
  return PTB_OK;
}
int ptb_scene_get_prim_order(const ptb_scene *s, int32_t *order, int64_t cap) {
  if (!s) return fail(PTB_E_INVALID, "get_prim_order: null scene");
  int64_t n = (int64_t)s->ref_order.size();
  if (order)
    for (int64_t i = 0; i < n && i < cap; ++i) order[i] = s->ref_order[i];
  return (int)n;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// OCaml 5 Stdlib.Random (LXM L64X128, MD5-seeded) under Base.Random.init / Base.Random.float
// (shirley_spheres/bin/main.ml:56,251).  Restated from the OCaml runtime's and Base's published algorithms; no live
// OCaml in this image, but the resulting sphere field is pinned by the golden image shirley-spheres.png.
// ---------------------------------------------------------------------------------------------
namespace {

struct Md5 {
  uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  static uint32_t rotl(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
  void block(const uint8_t *p) {
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                              5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                              4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                              6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t K[64];
    for (int i = 0; i < 64; ++i) K[i] = (uint32_t)std::floor(std::fabs(std::sin((double)(i + 1))) * 4294967296.0);
    uint32_t M[16];
    for (int i = 0; i < 16; ++i)
      M[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
             ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      if (i < 16) f = (b & c) | (~b & d), g = i;
      else if (i < 32) f = (d & b) | (~d & c), g = (5 * i + 1) % 16;
      else if (i < 48) f = b ^ c ^ d, g = (3 * i + 5) % 16;
      else f = c ^ (b | ~d), g = (7 * i) % 16;
      uint32_t t = d;
      d = c;
      c = b;
      b = b + rotl(a + f + K[i] + M[g], S[i]);
      a = t;
    }
    h[0] += a, h[1] += b, h[2] += c, h[3] += d;
  }
  void digest(const uint8_t *msg, size_t len, uint8_t out[16]) {
    std::vector<uint8_t> buf(msg, msg + len);
    buf.push_back(0x80);
    while (buf.size() % 64 != 56) buf.push_back(0);
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) buf.push_back((uint8_t)(bits >> (8 * i)));
    for (size_t i = 0; i < buf.size(); i += 64) block(&buf[i]);
    for (int i = 0; i < 4; ++i)
      for (int k = 0; k < 4; ++k) out[4 * i + k] = (uint8_t)(h[i] >> (8 * k));
  }
};

struct Lxm {
  uint64_t a, s, x0, x1;
  static uint64_t le64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
  }
  explicit Lxm(int64_t seed) {  // Random.State.make [| seed |]
    uint8_t buf[9];
    for (int i = 0; i < 8; ++i) buf[i] = (uint8_t)((uint64_t)seed >> (8 * i));
    uint8_t d1[16], d2[16];
    buf[8] = 1;
    Md5().digest(buf, 9, d1);
    buf[8] = 2;
    Md5().digest(buf, 9, d2);
    a = le64(d1) | 1;
    s = le64(d1 + 8);
    x0 = le64(d2);
    x1 = le64(d2 + 8);
    if (x0 == 0) x0 = 1;
    if (x1 == 0) x1 = 2;
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t z = s + x0;
    z = (z ^ (z >> 32)) * 0xdaba0b6eb09322e3ull;
    z = (z ^ (z >> 32)) * 0xdaba0b6eb09322e3ull;
    z = z ^ (z >> 32);
    s = s * 0xd1342543de82ef95ull + a;
    uint64_t q0 = x0, q1 = x1;
    q1 ^= q0;
    q0 = rotl(q0, 24);
    q0 = q0 ^ q1 ^ (q1 << 16);
    q1 = rotl(q1, 37);
    x0 = q0, x1 = q1;
    return z;
  }
  // Random.State.bits: 30 bits of one draw (OCaml 5 stdlib random.ml)
  uint32_t bits30() { return (uint32_t)(next() & 0x3FFFFFFFull); }
  // Base.Random.float 1.0 — shirley_spheres/bin/main.ml does `open! Base`, so `Random.float` is Base's, not
  // Stdlib's: two 30-bit draws, ((r1·2^-30) + r2)·2^-30, redrawn when the sum rounds up to 1 (Base random.ml,
  // third-party; restated from its published source).  Pinned by shirley-spheres.png: the projected small spheres'
  // albedos agree with the PNG's colours (tests/test_golden_png.py); Stdlib's 53-bit rule does not.
  double float1() {
    for (;;) {
      double r1 = (double)bits30();
      double r2 = (double)bits30();
      double x = ((r1 * 0x1.0p-30) + r2) * 0x1.0p-30;
      if (x < 1.0) return x * 1.0;
    }
  }
};

struct SceneWriter {
  HostScene &h;
  int solid(double r, double g, double b) {
    ptb_texture t{};
    t.kind = PTB_TEX_SOLID;
    t.rgb[0] = r, t.rgb[1] = g, t.rgb[2] = b;
    h.tex.push_back(t);
    return (int)h.tex.size() - 1;
  }
  int checker(int w, int hh, int even, int odd) {
    ptb_texture t{};
    t.kind = PTB_TEX_CHECKER;
    t.width = w, t.height = hh, t.even = even, t.odd = odd;
    h.tex.push_back(t);
    return (int)h.tex.size() - 1;
  }
  int material(int kind, int tex, double index) {
    ptb_material m{};
    m.kind = kind, m.texture = tex, m.index = index;
    h.mat.push_back(m);
    return (int)h.mat.size() - 1;
  }
  void sphere(double x, double y, double z, double r, int mat) {
    h.sx.push_back(x), h.sy.push_back(y), h.sz.push_back(z), h.sr.push_back(r), h.smat.push_back(mat);
  }
  // one triangle with its own three vertices; uv = (ua,va, ub,vb, uc,vc)
  void triangle(D3 a, D3 b, D3 c, int mat, const double uv[6]) {
    int base = (int)h.vx.size();
    for (D3 p : {a, b, c}) h.vx.push_back(p.x), h.vy.push_back(p.y), h.vz.push_back(p.z);
    h.tidx.push_back(base), h.tidx.push_back(base + 1), h.tidx.push_back(base + 2);
    h.tmat.push_back(mat);
    for (int i = 0; i < 6; ++i) h.tuv.push_back(uv[i]);
  }
};

void clear_scene(ptb_scene *s) {
  s->host = HostScene();
  s->ref_order.clear();
  s->committed = false;
}

void transform_all(ptb_scene *s, const double cam[20]) {
  HostScene &h = s->host;
  if (!h.sx.empty()) ptb_camera_transform(cam + 4, h.sx.data(), h.sy.data(), h.sz.data(), (int64_t)h.sx.size());
  if (!h.vx.empty()) ptb_camera_transform(cam + 4, h.vx.data(), h.vy.data(), h.vz.data(), (int64_t)h.vx.size());
}

}  // namespace

extern "C" {

int ptb_scene_load_shirley(ptb_scene *s, double aspect, int32_t seed, double cam[20]) {
  if (!s || !cam) return fail(PTB_E_INVALID, "load_shirley: null argument");
  clear_scene(s);
  SceneWriter w{s->host};
  // main.ml:26-31
  const double eye[3] = {13.0, 2.0, 4.5}, target[3] = {0, 0, 0}, up[3] = {0, 1, 0};
  ptb_camera_create(eye, target, up, aspect, 20.0, cam);
  // ground (main.ml:38-43)
  int ta = w.solid(0.2, 0.3, 0.1), tb = w.solid(0.9, 0.9, 0.9);
  int checks = w.material(PTB_MAT_LAMBERTIAN, w.checker(1000, 2000, ta, tb), 0.0);
  w.sphere(0.0, -1000.0, 0.0, 1000.0, checks);
  // big spheres (main.ml:45-54)
  int glass = w.material(PTB_MAT_DIELECTRIC, 0, 1.5);
  int metal = w.material(PTB_MAT_METAL, w.solid(0.7, 0.6, 0.5), 0.0);
  int blue = w.material(PTB_MAT_LAMBERTIAN, w.solid(0.1, 0.1, 0.7), 0.0);
  w.sphere(-4.0, 1.0, 0.0, 1.0, glass);
  w.sphere(0.0, 1.0, 0.0, 1.0, metal);
  w.sphere(4.0, 1.0, 0.0, 1.0, blue);
  // small spheres (main.ml:56-101), Random.init seed just before (main.ml:251)
  Lxm rng(seed);
  for (int a = -11; a <= 11; ++a)
    for (int b = -11; b <= 11; ++b) {
      double x = (double)a + (0.9 * rng.float1());
      double z = (double)b + (0.9 * rng.float1());
      const double radius = 0.2;
      double dx = 4.0 - x, dy = radius - radius, dz = 0.0 - z;
      double quadrance = std::fma(dx, dx, std::fma(dy, dy, dz * dz));
      if (!(quadrance > 0.81)) continue;
      double roll = rng.float1();
      int m;
      if (roll < 0.8) {
        // V3.Infix.(random_v3 () * random_v3 ()): component-wise product, order-insensitive
        double p[3], q[3];
        for (double &v : p) v = rng.float1();
        for (double &v : q) v = rng.float1();
        m = w.material(PTB_MAT_LAMBERTIAN, w.solid(p[0] * q[0], p[1] * q[1], p[2] * q[2]), 0.0);
      } else if (roll < 0.95) {
        double zc = (0.5 * rng.float1()) + 0.5;
        m = w.material(PTB_MAT_METAL, w.solid(zc, zc, zc), 0.0);
      } else {
        m = glass;
      }
      w.sphere(x, radius, z, radius, m);
    }
  // background (main.ml:104-110)
  s->host.bg_kind = PTB_BG_GRADIENT_Y;
  const double c0[3] = {1, 1, 1}, c1[3] = {0.5, 0.7, 1.0};
  std::memcpy(s->host.bg0, c0, sizeof c0);
  std::memcpy(s->host.bg1, c1, sizeof c1);
  transform_all(s, cam);  // main.ml:258-260
  for (int i = 0; i < (int)s->host.sr.size(); ++i) s->ref_order.push_back(i);
  return PTB_OK;
}

}  // extern "C"
static int load_cornell_impl(ptb_scene *s, double aspect, int32_t background_kind, const double c0[3],
                             const double c1[3], const double *radiance, double cam[20]) {
  clear_scene(s);
  SceneWriter w{s->host};
  // camera (main.ml:172-182)
  const double eye[3] = {0.5, 0.5, -1.0}, target[3] = {0.5, 0.5, 0.0}, up[3] = {0, 1, 0};
  double fov = (2.0 * std::atan(0.5)) * 180.0 / M_PI;
  ptb_camera_create(eye, target, up, aspect, fov, cam);
  // quad ~material a u v (main.ml:30-48): fan (a,t00)(b,t10)(c,t11)(d,t01) -> [(a,c,d); (a,b,c)]
  auto quad = [&](int mat, D3 a, D3 u, D3 v, std::vector<int> *ids) {
    D3 b = a + v, c = b + u, d = a + u;
    const double uv1[6] = {0, 0, 1, 1, 0, 1};  // a:t00 c:t11 d:t01
    const double uv2[6] = {0, 0, 1, 0, 1, 1};  // a:t00 b:t10 c:t11
    ids->push_back((int)s->host.tmat.size());
    w.triangle(a, c, d, mat, uv1);
    ids->push_back((int)s->host.tmat.size());
    w.triangle(a, b, c, mat, uv2);
  };
  // List.concat_no_order folds with rev_append: the result lists the LAST group first, each reversed
  auto concat_no_order = [](const std::vector<std::vector<int>> &groups) {
    std::vector<int> acc;
    for (const auto &g : groups) {
      std::vector<int> r(g.rbegin(), g.rend());
      r.insert(r.end(), acc.begin(), acc.end());
      acc.swap(r);
    }
    return acc;
  };
  const D3 O{0, 0, 0}, X{1, 0, 0}, Y{0, 1, 0}, Z{0, 0, 1};
  // light enclosure (main.ml:190-210)
  int lm = w.material(PTB_MAT_METAL, w.solid(0.30, 0.999, 0.30), 0.0);
  const double r = 0.05;
  D3 rx = X * r, ry = Y * r, rz = Z * r, lc{0.5, 0.82, 0.5};
  D3 pa = ((lc - rx) - ry) - rz, pb = ((lc + rx) - ry) + rz;
  std::vector<int> qr, qf, ql, qb;
  quad(lm, pa, rz * 2.0, ry * 2.0, &qr);
  quad(lm, pa, ry * 2.0, rx * 2.0, &qf);
  quad(lm, pb, (O - rz) * 2.0, ry * 2.0, &ql);
  quad(lm, pb, rx * 2.0, ry * 2.0, &qb);
  std::vector<int> enclosure = concat_no_order({qr, qf, ql, qb});
  // empty box (main.ml:52-68)
  int red = w.material(PTB_MAT_LAMBERTIAN, w.solid(0.7, 0.0, 0.0), 0.0);
  int bluem = w.material(PTB_MAT_LAMBERTIAN, w.solid(0.0, 0.0, 0.7), 0.0);
  int grey = w.material(PTB_MAT_LAMBERTIAN, w.solid(0.7, 0.7, 0.7), 0.0);
  int ca = w.solid(0.2, 0.3, 0.1), cb = w.solid(0.9, 0.9, 0.9);
  int chk = w.material(PTB_MAT_LAMBERTIAN, w.checker(10, 10, ca, cb), 0.0);
  std::vector<int> right, left, floor_, ceil_, rear;
  quad(red, O, Z, Y, &right);
  quad(bluem, X, Z, Y, &left);
  quad(chk, O, X, Z, &floor_);
  quad(grey, Y, X, Z, &ceil_);
  quad(grey, Z, X, Y, &rear);
  std::vector<int> box = concat_no_order({right, left, floor_, ceil_, rear});
  // spheres (main.ml:70-91)
  const double radius = 0.20;
  int wm = w.material(PTB_MAT_METAL, w.solid(1.0, 1.0, 1.0), 0.0);
  int gl = w.material(PTB_MAT_DIELECTRIC, 0, 1.5);
  int bc = w.material(PTB_MAT_LAMBERTIAN, w.solid(0.75, 0.75, 0.75), 0.0);
  w.sphere(1.0 - 0.1 - radius, radius, 1.0 - 0.2 - radius, radius, wm);
  w.sphere(0.1 + radius, 0.1 + radius, 0.2 + radius, radius, gl);
  w.sphere(0.5, 0.5, -2.0 - 10.0, 10.0, bc);
  // list order (main.ml:213-217): enclosure @ box @ spheres
  for (int t : enclosure) s->ref_order.push_back(~t);
  for (int t : box) s->ref_order.push_back(~t);
  for (int i = 0; i < 3; ++i) s->ref_order.push_back(i);
  s->host.bg_kind = background_kind;
  for (int i = 0; i < 3; ++i) s->host.bg0[i] = c0[i], s->host.bg1[i] = c1 ? c1[i] : c0[i];
  if (radiance) {
    // EXTENSION (BASELINE.json configs[1] "diffuse+light sampling"): the reference's point
    // light `light_pos` (main.ml:184-188; photons only) becomes a horizontal emissive square of the enclosure's
    // cross-section at that height, and diffuse_plus_light mixes it with the cosine lobe
    int em = w.material(PTB_MAT_EMISSIVE, w.solid(radiance[0], radiance[1], radiance[2]), 0.0);
    std::vector<int> lq;
    const D3 la = (lc - rx) - rz;
    quad(em, la, rx * 2.0, rz * 2.0, &lq);
    for (int t : lq) s->ref_order.push_back(~t);
    // the light quad moves to camera space with the scene: origin as a point, edges as point differences
    double px[3] = {la.x, la.x + 2 * r, la.x}, py[3] = {la.y, la.y, la.y}, pz[3] = {la.z, la.z, la.z + 2 * r};
    ptb_camera_transform(cam + 4, px, py, pz, 3);
    HostScene &h = s->host;
    h.has_light = true;
    h.light_o[0] = px[0], h.light_o[1] = py[0], h.light_o[2] = pz[0];
    h.light_u[0] = px[1] - px[0], h.light_u[1] = py[1] - py[0], h.light_u[2] = pz[1] - pz[0];
    h.light_v[0] = px[2] - px[0], h.light_v[1] = py[2] - py[0], h.light_v[2] = pz[2] - pz[0];
  }
  transform_all(s, cam);  // main.ml:211-218
  return PTB_OK;
}
extern "C" {
int ptb_scene_load_cornell(ptb_scene *s, double aspect, int32_t background_kind, const double c0[3],
                           const double c1[3], double cam[20]) {
  if (!s || !cam || !c0) return fail(PTB_E_INVALID, "load_cornell: null argument");
  return load_cornell_impl(s, aspect, background_kind, c0, c1, nullptr, cam);
}
int ptb_scene_load_cornell_lit(ptb_scene *s, double aspect, const double radiance[3], double cam[20]) {
  if (!s || !cam || !radiance) return fail(PTB_E_INVALID, "load_cornell_lit: null argument");
  const double black[3] = {0, 0, 0};
  return load_cornell_impl(s, aspect, PTB_BG_CONSTANT, black, black, radiance, cam);
}

int ptb_scene_load_mesh(ptb_scene *s, const float *xyz, int64_t nv, const int32_t *faces, int64_t nf,
                        double aspect, double cam[20]) {
  if (!s || !cam || !xyz || !faces || nv <= 0 || nf <= 0) return fail(PTB_E_INVALID, "load_mesh: bad args");
  for (int64_t i = 0; i < 3 * nf; ++i)
    if (faces[i] < 0 || faces[i] >= nv) return fail(PTB_E_INVALID, "load_mesh: vertex index out of bounds");
  clear_scene(s);
  HostScene &h = s->host;
  SceneWriter w{h};
  // camera (ganesha/bin/main.ml:30-35)
  const double eye[3] = {328.0, 70.282, 345.0}, target[3] = {328.0, 10.0, 0.0};
  const double up[3] = {-0.00212272, 0.998201, -0.0599264};
  ptb_camera_create(eye, target, up, aspect, 30.0, cam);
  // Mesh.create (main.ml:45-86): float columns widened to double, moved to camera space
  h.vx.resize((size_t)nv), h.vy.resize((size_t)nv), h.vz.resize((size_t)nv);
  for (int64_t i = 0; i < nv; ++i) h.vx[i] = xyz[3 * i], h.vy[i] = xyz[3 * i + 1], h.vz[i] = xyz[3 * i + 2];
  ptb_camera_transform(cam + 4, h.vx.data(), h.vy.data(), h.vz.data(), nv);
  int green = w.material(PTB_MAT_LAMBERTIAN, w.solid(0.1, 0.7, 0.2), 0.0);  // main.ml:113-115
  h.tidx.assign(faces, faces + 3 * nf);
  h.tmat.assign((size_t)nf, green);
  h.tuv.resize((size_t)6 * nf);
  for (int64_t i = 0; i < nf; ++i) {
    const double d[6] = {0, 0, 0, 1, 1, 1};
    std::memcpy(&h.tuv[6 * i], d, sizeof d);
  }
  // Floor (main.ml:205-228), in camera space, from the mesh tree's bbox
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = 0; i < 3 * nf; ++i) {
    int v = faces[i];
    double p[3] = {h.vx[v], h.vy[v], h.vz[v]};
    for (int a = 0; a < 3; ++a) mn[a] = std::min(mn[a], p[a]), mx[a] = std::max(mx[a], p[a]);
  }
  D3 center{(mn[0] + mx[0]) * 0.5, mn[1], (mn[2] + mx[2]) * 0.5};
  D3 xp{5000.0, 0, 0}, zp{0, 0, 5000.0};
  int fa = w.solid(0.2, 0.3, 0.1), fb = w.solid(0.9, 0.9, 0.9);
  int fm = w.material(PTB_MAT_LAMBERTIAN, w.checker(500, 500, fa, fb), 0.0);
  D3 A = center - (xp + zp), B = A + xp * 2.0, C = B + zp * 2.0, Dd = A + zp * 2.0;
  const double uv1[6] = {0, 0, 0, 1, 1, 1};  // a:t00 b:t01 c:t11
  const double uv2[6] = {0, 0, 1, 1, 1, 0};  // a:t00 c:t11 d:t10
  w.triangle(A, B, C, fm, uv1);
  w.triangle(A, C, Dd, fm, uv2);
  for (int64_t i = 0; i < nf + 2; ++i) s->ref_order.push_back(~(int32_t)i);
  h.bg_kind = PTB_BG_GRADIENT_Y;
  return PTB_OK;
}

int ptb_mesh_synthetic(int64_t target_faces, uint32_t seed, float *xyz, int64_t cap_v, int32_t *faces,
                       int64_t cap_f, int64_t *nv_out, int64_t *nf_out) {
  if (target_faces < 8) return fail(PTB_E_INVALID, "mesh_synthetic: too few faces");
  // lat-long sphere with nu x nv quads -> 2*nu*nv triangles, radius displaced by seeded waves
  int64_t nvq = (int64_t)std::max(2.0, std::floor(std::sqrt((double)target_faces / 4.0)));
  int64_t nu = 2 * nvq;
  int64_t n_vertices = (nu + 1) * (nvq + 1), n_faces = 2 * nu * nvq;
  if (nv_out) *nv_out = n_vertices;
  if (nf_out) *nf_out = n_faces;
  if (!xyz || !faces) return PTB_OK;
  if (cap_v < n_vertices || cap_f < n_faces) return fail(PTB_E_INVALID, "mesh_synthetic: buffers too small");
  auto hash01 = [&](uint32_t k) {
    uint32_t x = k * 0x9E3779B9u + seed;
    x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
    return (double)x / 4294967296.0;
  };
  double ph[8], fr[8];
  for (int k = 0; k < 8; ++k) ph[k] = 6.283185307179586 * hash01(k), fr[k] = 2.0 + std::floor(9.0 * hash01(100 + k));
  const double cx = 328.0, cy = 45.0, cz = 0.0, R = 35.0;
  for (int64_t j = 0; j <= nvq; ++j)
    for (int64_t i = 0; i <= nu; ++i) {
      double u = (double)(i % nu) / (double)nu, v = (double)j / (double)nvq;
      double phi = 6.283185307179586 * u, th = 3.141592653589793 * v;
      double disp = 0.0;
      for (int k = 0; k < 8; ++k) disp += std::sin(fr[k] * phi * (k % 2 ? 1.0 : 0.0) + fr[7 - k] * th + ph[k]) / (2.0 + k);
      double rr = R * (1.0 + 0.12 * disp * std::sin(th));
      int64_t id = j * (nu + 1) + i;
      xyz[3 * id] = (float)(cx + rr * std::sin(th) * std::cos(phi));
      xyz[3 * id + 1] = (float)(cy + rr * std::cos(th));
      xyz[3 * id + 2] = (float)(cz + rr * std::sin(th) * std::sin(phi));
    }
  int64_t f = 0;
  for (int64_t j = 0; j < nvq; ++j)
    for (int64_t i = 0; i < nu; ++i) {
      int32_t a = (int32_t)(j * (nu + 1) + i), b = a + 1, c = (int32_t)(a + nu + 1), d = c + 1;
      faces[3 * f] = a, faces[3 * f + 1] = c, faces[3 * f + 2] = b, ++f;
      faces[3 * f] = b, faces[3 * f + 1] = c, faces[3 * f + 2] = d, ++f;
    }
  return PTB_OK;
}

}  // extern "C"

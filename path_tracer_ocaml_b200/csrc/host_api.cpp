// host_api.cpp — host-only entry points of libptb200: error state and the small reference
// functions a caller needs around the render call (sample-sequence constants, tile list, filter
// weights, camera).  All float64, as in the reference.  Product code: nothing from oracle/.
#include <cmath>
#include <cstring>
#include <string>

#include "presplit.hpp"
#include "scene.hpp"

namespace ptb {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}

// Low_discrepancy_sequence.phi_approx / alpha (low_discrepancy_sequence.ml:8-25).  The constants
// must come from the same libm `pow` the OCaml runtime calls, iterated to the same fixed point,
// because the sample stream is required to be bit-exact.  volatile keeps gcc from folding pow.
void lds_alpha(int dimension, double *alpha) {
  volatile double inv = 1.0 / ((double)dimension + 1.0);
  double phi = 2.0;
  for (int guard = 0; guard < 100000; ++guard) {
    double next = std::pow(1.0 + phi, inv);
    if (next == phi) break;
    phi = next;
  }
  for (int i = 0; i < dimension; ++i) alpha[i] = 1.0 / std::pow(phi, (double)(i + 1));
}

// Filter_kernel.Binomial.create (filter_kernel.ml:49-85): the order-`order` binomial row is box-
// resampled onto 2r+1 taps.  The reference does this in exact rationals; with every quantity a
// multiple of 1/(2r+1) the same sums are exact in integers scaled by (2r+1).
void filter_binomial(int order, int pixel_radius, std::vector<double> *out) {
  const int taps = 2 * pixel_radius + 1;
  std::vector<long long> row(order, 1);  // C(order-1, k)
  for (int k = 1; k < order; ++k) row[k] = row[k - 1] * (order - k) / k;
  // tap i covers [i*order/taps, (i+1)*order/taps) of the source row; in units of 1/taps the source
  // cell k spans [k*taps, (k+1)*taps) and the tap spans [i*order, (i+1)*order).
  std::vector<double> w1(taps);
  for (int i = 0; i < taps; ++i) {
    long long lo = (long long)i * order, hi = (long long)(i + 1) * order, acc = 0;
    for (int k = 0; k < order; ++k) {
      long long a = std::max(lo, (long long)k * taps), b = std::min(hi, (long long)(k + 1) * taps);
      if (b > a) acc += (b - a) * row[k];
    }
    w1[i] = (double)acc / (double)taps;  // = float_of_num of the exact rational
  }
  double total = 0.0;
  for (double x : w1) total += x;
  for (double &x : w1) x /= total;
  out->assign((size_t)taps * taps, 0.0);
  for (int r = 0; r < taps; ++r)
    for (int c = 0; c < taps; ++c) (*out)[(size_t)r * taps + c] = w1[r] * w1[c];
}

// Tile.split (tile.ml:14-39): halve the longer side (width only if strictly longer) until the area
// fits; left/top half gets floor(n/2); output order = depth-first, left before right.
void tile_split(int width, int height, int max_area, std::vector<TileRect> *out) {
  std::vector<TileRect> stack{{0, 0, width, height}};
  while (!stack.empty()) {
    TileRect t = stack.back();
    stack.pop_back();
    if ((long long)t.w * t.h <= max_area) {
      out->push_back(t);
      continue;
    }
    TileRect a = t, b = t;
    if (t.w > t.h) {
      a.w = t.w / 2;
      b.col = t.col + a.w;
      b.w = t.w - a.w;
    } else {
      a.h = t.h / 2;
      b.row = t.row + a.h;
      b.h = t.h - a.h;
    }
    stack.push_back(b);  // right/bottom half is visited after the whole left/top subtree
    stack.push_back(a);
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" {

const char *ptb_last_error(void) { return g_error.c_str(); }
const char *ptb_version(void) { return "ptb200 0.1 (sm_100a)"; }
int ptb_leaf_size(void) { return LEAF_MAX; }

int ptb_lds_alpha(int32_t dimension, double *alpha) {
  // Low_discrepancy_sequence.create: `failwith "expected dimension >= 1"` (…ml:27-31)
  if (dimension < 1 || !alpha) return fail(PTB_E_INVALID, "ptb_lds_alpha: expected dimension >= 1");
  lds_alpha(dimension, alpha);
  return PTB_OK;
}

int ptb_tile_split(int32_t width, int32_t height, int32_t max_area, int32_t *row, int32_t *col,
                   int32_t *w, int32_t *h, int32_t cap) {
  if (width <= 0 || height <= 0 || max_area <= 0) return fail(PTB_E_INVALID, "ptb_tile_split: bad size");
  std::vector<TileRect> t;
  tile_split(width, height, max_area, &t);
  if ((int)t.size() > cap) return -1000 - (int)t.size();
  for (size_t i = 0; i < t.size(); ++i) row[i] = t[i].row, col[i] = t[i].col, w[i] = t[i].w, h[i] = t[i].h;
  return (int)t.size();
}

int ptb_filter_binomial(int32_t order, int32_t pixel_radius, double *weights) {
  if (order < 1 || pixel_radius < 0 || !weights) return fail(PTB_E_INVALID, "ptb_filter_binomial: bad args");
  std::vector<double> w;
  filter_binomial(order, pixel_radius, &w);
  std::memcpy(weights, w.data(), w.size() * sizeof(double));
  return PTB_OK;
}

int ptb_presplit_boxes(const double *v, double cell, const double origin[3], double *boxes, int32_t cap) {
  if (!v || !origin || (!boxes && cap > 0) || !(cell > 0.0)) return fail(PTB_E_INVALID, "ptb_presplit_boxes: bad args");
  double tri[3][3];
  for (int k = 0; k < 3; ++k)
    for (int a = 0; a < 3; ++a) tri[k][a] = v[3 * k + a];
  int n = 0;
  return presplit::pieces(tri, cell, origin, [&](const double lo[3], const double hi[3]) {
    if (n < cap)
      for (int a = 0; a < 3; ++a) boxes[6 * n + a] = lo[a], boxes[6 * n + 3 + a] = hi[a];
    ++n;
  });
}

// Camera.create (camera.ml:58-83) keeps only what Camera.ray and Camera.transform read.
int ptb_camera_create(const double eye[3], const double target[3], const double up[3], double aspect,
                      double vertical_fov_deg, double out[20]) {
  if (!eye || !target || !up || !out) return fail(PTB_E_INVALID, "ptb_camera_create: null argument");
  const double half_h = std::tan(0.5 * (vertical_fov_deg * M_PI / 180.0));
  const double half_w = aspect * half_h;
  out[0] = -half_w;
  out[1] = -half_h;
  out[2] = 2.0 * half_w;
  out[3] = 2.0 * half_h;
  // Mat4.look_at (camera.ml:14-27).  The reference's V3.normalize is 1/hypot(x, hypot(y, z)) then a
  // scale, V3.cross/dot use Float.fma (affine.ml:60-73); the scene is moved by this matrix before
  // anything else happens, so follow the same operation order.
  auto nrm = [](const double v[3], double o[3]) {
    double s = 1.0 / std::hypot(v[0], std::hypot(v[1], v[2]));
    o[0] = s * v[0], o[1] = s * v[1], o[2] = s * v[2];
  };
  auto crs = [](const double p[3], const double q[3], double o[3]) {
    o[0] = std::fma(p[1], q[2], -(p[2] * q[1]));
    o[1] = std::fma(p[2], q[0], -(p[0] * q[2]));
    o[2] = std::fma(p[0], q[1], -(p[1] * q[0]));
  };
  auto dt = [](const double p[3], const double q[3]) {
    return std::fma(p[0], q[0], std::fma(p[1], q[1], p[2] * q[2]));
  };
  double fwd[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
  double zc[3], upn[3], xc[3], yc[3], t[3];
  nrm(fwd, zc);
  nrm(up, upn);
  crs(zc, upn, t);
  nrm(t, xc);
  crs(xc, zc, t);
  nrm(t, yc);
  double *m = out + 4;
  for (int i = 0; i < 3; ++i) m[i] = xc[i], m[4 + i] = yc[i], m[8 + i] = -zc[i];
  m[3] = -dt(eye, xc);
  m[7] = -dt(eye, yc);
  m[11] = dt(eye, zc);
  m[12] = m[13] = m[14] = 0.0;
  m[15] = 1.0;
  return PTB_OK;
}

// Mat4.transform (camera.ml:29-43): four plain (unfused) 4-term dot products, then divide by w.
int ptb_camera_transform(const double look_at[16], double *xs, double *ys, double *zs, int64_t n) {
  if (!look_at || !xs || !ys || !zs || n < 0) return fail(PTB_E_INVALID, "ptb_camera_transform: bad args");
  for (int64_t i = 0; i < n; ++i) {
    double v[4] = {xs[i], ys[i], zs[i], 1.0}, r[4];
    for (int k = 0; k < 4; ++k) {
      const double *row = look_at + 4 * k;
      volatile double p0 = v[0] * row[0], p1 = v[1] * row[1], p2 = v[2] * row[2], p3 = v[3] * row[3];
      r[k] = ((p0 + p1) + p2) + p3;
    }
    double s = 1.0 / r[3];
    xs[i] = s * r[0], ys[i] = s * r[1], zs[i] = s * r[2];
  }
  return PTB_OK;
}

}  // extern "C"

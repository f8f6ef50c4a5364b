// scene.hpp — host-side scene container and wide-BVH description shared by the .cpp/.cu files of
// libptb200.  Product code: must not include anything from oracle/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ptb200.h"

namespace ptb {

struct D3 {
  double x, y, z;
};
static inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline D3 cross(D3 a, D3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Box {
  double mn[3], mx[3];
  void reset() {
    for (int i = 0; i < 3; ++i) mn[i] = 1e300, mx[i] = -1e300;
  }
  void grow(const Box &o) {
    for (int i = 0; i < 3; ++i) {
      if (o.mn[i] < mn[i]) mn[i] = o.mn[i];
      if (o.mx[i] > mx[i]) mx[i] = o.mx[i];
    }
  }
  double area() const {
    double x = mx[0] - mn[0], y = mx[1] - mn[1], z = mx[2] - mn[2];
    if (x < 0 || y < 0 || z < 0) return 0.0;
    return 2.0 * (x * y + y * z + z * x);
  }
};

// 4-wide BVH node, double precision master copy.  child[k] >= 0: inner node index;
// child[k] < 0: leaf, ~child[k] = first | (count-1) << 26 | type << 30; EMPTY_CHILD: unused slot
// (its box is inverted so the slab test can never pass).
constexpr int32_t EMPTY_CHILD = INT32_MIN;
constexpr int LEAF_MAX = 4;
struct WideNode {
  double mn[3][4], mx[3][4];
  int32_t child[4];
};
static inline int32_t leaf_code(int type, int first, int count) {
  return ~(int32_t)((uint32_t)first | ((uint32_t)(count - 1) << 26) | ((uint32_t)type << 30));
}

struct WideBVH {
  std::vector<WideNode> nodes;     // BFS order, node 0 = root
  std::vector<int32_t> sphere_order;  // device slot -> sphere index as set by the caller
  std::vector<int32_t> tri_order;     // device slot -> triangle index as set by the caller
  int depth = 0;                   // inner levels
  int max_stack = 0;               // worst-case traversal stack entries
  bool presplit = false;           // triangle slots are REFERENCES (presplit.hpp): a triangle may occupy several slots
  // tree built on the device (gpu_bvh.cuh): `nodes` / `tri_order` stay empty, only the counts are known here
  bool device_built = false;
  int64_t dev_nodes = 0, dev_tris = 0, dev_leaves = 0;
  int64_t node_count() const { return device_built ? dev_nodes : (int64_t)nodes.size(); }
  int64_t tri_count() const { return device_built ? dev_tris : (int64_t)tri_order.size(); }
};

struct HostScene {
  std::vector<ptb_texture> tex;
  std::vector<ptb_material> mat;
  std::vector<double> sx, sy, sz, sr;
  std::vector<int32_t> smat;
  std::vector<double> vx, vy, vz;
  std::vector<int32_t> tidx, tmat;
  std::vector<double> tuv;  // 6 per triangle
  int bg_kind = PTB_BG_GRADIENT_Y;
  double bg0[3] = {1, 1, 1}, bg1[3] = {0.5, 0.7, 1.0};
  // extension (ptb_scene_set_light_quad): diffuse_plus_light = Mix (Diffuse, Quad_light)
  bool has_light = false;
  double light_o[3] = {0, 0, 0}, light_u[3] = {1, 0, 0}, light_v[3] = {0, 0, 1};
  bool has_emissive() const {
    for (const ptb_material &m : mat)
      if (m.kind == PTB_MAT_EMISSIVE) return true;
    return false;
  }
  int64_t n_spheres() const { return (int64_t)sr.size(); }
  int64_t n_tris() const { return (int64_t)tidx.size() / 3; }
};

// bvh.cpp
void build_wide_bvh(const HostScene &s, WideBVH *out);

// errors (host_api.cpp)
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

// host helpers (host_api.cpp)
void lds_alpha(int dimension, double *alpha);
void filter_binomial(int order, int pixel_radius, std::vector<double> *w);
struct TileRect {
  int row, col, w, h;
};
void tile_split(int width, int height, int max_area, std::vector<TileRect> *out);

}  // namespace ptb

// soup_sim.cpp — development tool (CPU, not product, not oracle): per-ray traversal work of the host builder's 4-wide
// tree on the C5 triangle soup (random triangles in [-10,10]^3, edge ~0.3), for incoherent rays: node visits, leaf
// visits and triangle tests per ray, with the builder's knobs (PTB_BVH_PRESPLIT, PTB_BVH_LEAF, PTB_BVH_CT) taken
// from the environment.
// Build: g++ -O2 -std=c++17 -pthread -I../../path_tracer_ocaml_b200/csrc soup_sim.cpp ../../path_tracer_ocaml_b200/csrc/bvh.cpp -o /tmp/soup_sim
// usage: /tmp/soup_sim [triangles=1000000] [rays=20000]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "scene.hpp"

namespace ptb {
void set_error(const std::string &) {}
int fail(int c, const std::string &) { return c; }
}  // namespace ptb
using namespace ptb;

static HostScene H;
static WideBVH B;

static bool tri_hit(int id, const double o[3], const double d[3], double &tbest) {
  const int *ix = &H.tidx[3 * id];
  D3 v0{H.vx[ix[0]], H.vy[ix[0]], H.vz[ix[0]]}, v1{H.vx[ix[1]], H.vy[ix[1]], H.vz[ix[1]]}, v2{H.vx[ix[2]], H.vy[ix[2]], H.vz[ix[2]]};
  D3 e1 = v1 - v0, e2 = v2 - v0, D{d[0], d[1], d[2]}, O{o[0], o[1], o[2]};
  D3 p = cross(D, e2);
  double det = dot(e1, p);
  if (std::fabs(det) < 1e-6) return false;
  double inv = 1.0 / det;
  D3 tv = O - v0;
  double u = dot(tv, p) * inv;
  if (u < 0 || u > 1) return false;
  D3 q = cross(tv, e1);
  double v = dot(D, q) * inv;
  if (v < 0 || u + v > 1) return false;
  double t = dot(e2, q) * inv;
  if (t >= 0 && t <= tbest) {
    tbest = t;
    return true;
  }
  return false;
}

int main(int argc, char **argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 1000000, nr = argc > 2 ? atoi(argv[2]) : 20000;
  std::mt19937_64 rng(0xB200);
  std::uniform_real_distribution<double> U(-10, 10);
  std::normal_distribution<double> N(0.0, 0.3), N1(0.0, 1.0);
  H.vx.resize(3 * (size_t)m), H.vy.resize(3 * (size_t)m), H.vz.resize(3 * (size_t)m), H.tidx.resize(3 * (size_t)m);
  H.tmat.assign(m, 0);
  for (int i = 0; i < m; ++i) {
    double a[3] = {U(rng), U(rng), U(rng)};
    for (int k = 0; k < 3; ++k) {
      H.vx[3 * i + k] = a[0] + (k ? N(rng) : 0), H.vy[3 * i + k] = a[1] + (k ? N(rng) : 0), H.vz[3 * i + k] = a[2] + (k ? N(rng) : 0);
      H.tidx[3 * i + k] = 3 * i + k;
    }
  }
  build_wide_bvh(H, &B);
  long leaves = 0, leaf_prims = 0;
  for (auto &nd : B.nodes)
    for (int k = 0; k < 4; ++k)
      if (nd.child[k] < 0 && nd.child[k] != EMPTY_CHILD) ++leaves, leaf_prims += ((~(unsigned)nd.child[k] >> 26) & 15) + 1;
  printf("tree: %zu nodes, %zu refs (x%.2f), depth %d, max_stack %d, %ld leaves (%.2f refs/leaf)\n", B.nodes.size(), B.tri_order.size(),
         (double)B.tri_order.size() / m, B.depth, B.max_stack, leaves, (double)leaf_prims / leaves);
  double nodes = 0, lvs = 0, tris = 0, hits = 0, maxsp = 0;
  for (int r = 0; r < nr; ++r) {
    double o[3] = {U(rng), U(rng), U(rng)}, d[3] = {N1(rng), N1(rng), N1(rng)};
    double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int a = 0; a < 3; ++a) d[a] /= l;
    double id[3] = {1 / d[0], 1 / d[1], 1 / d[2]};
    double tbest = 3e38;
    int best = -1;
    struct E {
      int ref;
      double t;
    };
    std::vector<E> st;
    int cur = 0;
    for (;;) {
      if (cur >= 0) {
        ++nodes;
        const WideNode &n = B.nodes[cur];
        E h[4];
        int nh = 0;
        for (int k = 0; k < 4; ++k) {
          if (n.child[k] == EMPTY_CHILD) continue;
          double tn = 0, tf = 3e38;
          for (int a = 0; a < 3; ++a) {
            double t0 = (n.mn[a][k] - o[a]) * id[a], t1 = (n.mx[a][k] - o[a]) * id[a];
            tn = std::max(tn, std::min(t0, t1)), tf = std::min(tf, std::max(t0, t1));
          }
          if (tn <= tf && tn < tbest) h[nh++] = {n.child[k], tn};
        }
        std::sort(h, h + nh, [](const E &a, const E &b) { return a.t < b.t; });
        for (int k = nh - 1; k >= 1; --k) st.push_back(h[k]);
        maxsp = std::max(maxsp, (double)st.size());
        if (nh) {
          cur = h[0].ref;
          continue;
        }
      } else {
        ++lvs;
        unsigned code = ~(unsigned)cur;
        int first = code & 0x3FFFFFF, cnt = ((code >> 26) & 15) + 1;
        for (int i = 0; i < cnt; ++i) {
          ++tris;
          if (tri_hit(B.tri_order[first + i], o, d, tbest)) best = B.tri_order[first + i];
        }
      }
      bool got = false;
      while (!st.empty()) {
        E e = st.back();
        st.pop_back();
        if (e.t <= tbest) {
          cur = e.ref, got = true;
          break;
        }
      }
      if (!got) break;
    }
    hits += best >= 0;
  }
  printf("per ray: %.1f node visits, %.1f leaf visits, %.1f triangle tests; hit fraction %.3f; deepest stack %g\n", nodes / nr, lvs / nr,
         tris / nr, hits / nr, maxsp);
  return 0;
}

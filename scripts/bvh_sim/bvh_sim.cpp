// bvh_sim.cpp — development tool (CPU, not product, not oracle): replays the k_trace warp loop on the host
// to count, for a given tree builder / traversal policy, the lane-level work (node visits, sphere tests,
// pops) and the WARP-level instruction cost of the if-if loop (cost model calibrated against ncu).
// Build: g++ -O2 -std=c++17 -I../../path_tracer_ocaml_b200/csrc bvh_sim.cpp ../../path_tracer_ocaml_b200/csrc/bvh.cpp -o /tmp/bvh_sim
// Input: /tmp/shirley_scene.bin (scripts/bvh_sim/dump_scene.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "scene.hpp"

namespace ptb {
void set_error(const std::string &) {}
int fail(int c, const std::string &) { return c; }
}  // namespace ptb
using namespace ptb;

struct Ray {
  float o[3], d[3];
};
struct F3 {
  float x, y, z;
};
static inline F3 sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline F3 mul(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline F3 add(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline F3 norm(F3 a) { return mul(a, 1.0f / std::sqrt(dot(a, a))); }

static HostScene H;
static std::vector<int> kind;
static WideBVH B;

static bool sphere_hit(int i, const Ray &r, float &tbest) {
  F3 o = {r.o[0], r.o[1], r.o[2]}, d = {r.d[0], r.d[1], r.d[2]};
  F3 c = {(float)H.sx[i], (float)H.sy[i], (float)H.sz[i]};
  float rad = (float)H.sr[i];
  F3 f = sub(c, o);
  float a = dot(d, d), bp = dot(f, d), boa = bp / a;
  F3 w = sub(mul(d, boa), f);
  float disc = rad * rad - dot(w, w);
  if (disc < 0) return false;
  float q = bp + std::copysign(std::sqrt(a * disc), bp);
  float cc = dot(f, f) - rad * rad;
  float t = cc > 0 ? cc / q : q / a;
  if (t >= 0 && t <= tbest) {
    tbest = t;
    return true;
  }
  return false;
}

// ---- policy knobs -------------------------------------------------------------------------------
static int SORT = 2;   // 0: nearest only, rest in slot order; 1: 3-exchange; 2: full sort
static int KEEP = 14;  // refill threshold
static int FULLSORT = 0;  // 1: upper bound — every bounce's rays globally sorted by (direction octant, origin Morton cell)
static int OCT = 0;    // 1: k_shade bins outgoing rays by direction octant (8 open segments per producer warp)
static int SPEC = 0;   // 1: speculative traversal: a lane parks ONE leaf and keeps traversing
static int DEFER = 0;  // 1: flush work runs on full warps of parked results (cost C_FLUSH per 32 rays) + C_PARK per event
static int C_PARK = 25;
static int ONE = 0;  // 1: the leaf phase tests ONE sphere per iteration (multi-sphere leaves stay pending)
static int DOUBLE = 0;  // 1: two node steps per loop iteration
static int LEAF_T = 1; // leaf phase runs when >= LEAF_T lanes hold a leaf, or no lane can do anything else
static int C_NODE = 93, C_SPH = 36, C_LEAF0 = 12, C_POP0 = 8, C_POPIT = 6, C_LOOP = 14, C_FLUSH = 90, C_REFILL = 85;

struct Lane {
  Ray r;
  float idir[3], tbest;
  int best, cur;
  int pending = 0;  // SPEC: parked leaf code (0 = none)
  std::vector<std::pair<int, float>> stk;
  bool active = false;
};
struct Counts {
  double node = 0, sph = 0, pop = 0, rays = 0;
  double w_iters = 0, w_node = 0, w_leaf_sph = 0, w_leaf = 0, w_pop = 0, w_popit = 0, w_flush = 0, cost = 0;
  double lanes_node = 0, lanes_leaf = 0;
};
constexpr int POP = INT32_MIN + 2, DONE = INT32_MIN + 1;

static void init_lane(Lane &L, const Ray &r) {
  L.r = r;
  for (int a = 0; a < 3; ++a) L.idir[a] = 1.0f / (std::fabs(r.d[a]) < 1e-30f ? std::copysign(1e-30f, r.d[a]) : r.d[a]);
  L.tbest = 3.4e38f;
  L.best = -1;
  L.cur = 0;
  L.pending = 0;
  L.stk.clear();
  L.active = true;
}

// node step of every lane that is at an inner node; returns the number of lanes that took it
static int node_step(Lane *W, Counts &C) {
  bool any_node = false;
  int nl_node = 0;
  // node phase
  for (int l = 0; l < 32; ++l) {
    Lane &L = W[l];
    if (!L.active || L.cur < 0) continue;
    any_node = true;
    ++nl_node;
    C.node++;
    const WideNode &n = B.nodes[L.cur];
    float tn[4];
    int ch[4];
    for (int k = 0; k < 4; ++k) {
      float t0 = 0, t1 = L.tbest;
      for (int a = 0; a < 3; ++a) {
        float lo = ((float)n.mn[a][k] - L.r.o[a]) * L.idir[a], hi = ((float)n.mx[a][k] - L.r.o[a]) * L.idir[a];
        if (lo > hi) std::swap(lo, hi);
        t0 = std::max(t0, lo), t1 = std::min(t1, hi);
      }
      tn[k] = (t0 <= t1 && n.child[k] != EMPTY_CHILD) ? t0 : INFINITY;
      ch[k] = n.child[k];
    }
    int order[4] = {0, 1, 2, 3};
    if (SORT == 2) {
      std::stable_sort(order, order + 4, [&](int a, int b) { return tn[a] < tn[b]; });
    } else {
      int m = 0;
      for (int k = 1; k < 4; ++k)
        if (tn[k] < tn[m]) m = k;
      std::swap(order[0], order[m]);
      if (SORT == 1) {  // emulate: rest in "some" order; use slot order as well
      }
    }
    if (tn[order[0]] == INFINITY) {
      L.cur = POP;
    } else {
      L.cur = ch[order[0]];
      for (int k = 3; k >= 1; --k)
        if (tn[order[k]] < INFINITY) L.stk.push_back({ch[order[k]], tn[order[k]]});
    }
  }
  return nl_node;
}
// pop step of every lane in POP state; returns the longest pop loop (0: no lane popped)
static int pop_step(Lane *W, Counts &C) {
  int max_pop = 0;
  for (int l = 0; l < 32; ++l) {
    Lane &L = W[l];
    if (!L.active || L.cur != POP) continue;
    int it = 0;
    for (;;) {
      ++it;
      if (L.stk.empty()) {
        L.cur = DONE;
        L.active = false;
        break;
      }
      auto e = L.stk.back();
      L.stk.pop_back();
      C.pop++;
      if (e.second <= L.tbest) {
        L.cur = e.first;
        break;
      }
    }
    max_pop = std::max(max_pop, it);
  }
  return max_pop;
}

// one if-if iteration for a warp; returns cost
static void warp_iter(Lane *W, Counts &C) {
  bool any_node = false, any_leaf = false, any_pop = false;
  int nl_node = 0, nl_leaf = 0;
  nl_node = node_step(W, C);
  any_node = nl_node > 0;
  double extra = 0;
  if (DOUBLE && !SPEC) {  // a second node step in the same iteration (lanes that missed pop in between)
    const int mp = pop_step(W, C);
    if (mp) extra += C_POP0 + C_POPIT * mp + 4, C.w_pop++, C.w_popit += mp;
    const int n2 = node_step(W, C);
    if (n2) extra += C_NODE + 3, C.w_node++, C.lanes_node += n2;
  }
  if (SPEC) {
    // park leaves: a lane whose cur is a leaf and whose slot is free parks it and pops on
    int pend = 0, blocked = 0, other = 0;
    for (int l = 0; l < 32; ++l) {
      Lane &L = W[l];
      if (!L.active) continue;
      const bool at_leaf = L.cur < 0 && L.cur > POP;
      if (at_leaf && L.pending == 0) L.pending = L.cur, L.cur = POP;
    }
    // pop phase (speculative: tbest does not include the parked leaf)
    int max_pop = 0;
    for (int l = 0; l < 32; ++l) {
      Lane &L = W[l];
      if (!L.active || L.cur != POP) continue;
      int it = 0;
      for (;;) {
        ++it;
        if (L.stk.empty()) { L.cur = DONE; break; }
        auto e = L.stk.back();
        L.stk.pop_back();
        C.pop++;
        if (e.second <= L.tbest) { L.cur = e.first; break; }
      }
      any_pop = true;
      max_pop = std::max(max_pop, it);
    }
    for (int l = 0; l < 32; ++l) {
      Lane &L = W[l];
      if (!L.active) continue;
      const bool at_leaf = L.cur < 0 && L.cur > POP;
      if (L.pending) ++pend;
      if ((at_leaf && L.pending) || (L.cur == DONE && L.pending)) ++blocked;
      else if (L.cur != DONE) ++other;
    }
    int max_cnt = 0;
    const bool do_leaf = pend >= LEAF_T || (pend > 0 && other == 0) || blocked >= LEAF_T / 2 + 1;
    if (do_leaf) {
      for (int l = 0; l < 32; ++l) {
        Lane &L = W[l];
        if (!L.active || !L.pending) continue;
        any_leaf = true;
        ++nl_leaf;
        unsigned code = ~(unsigned)L.pending;
        int first = code & 0x3FFFFFF, cnt = ((code >> 26) & 15) + 1;
        max_cnt = std::max(max_cnt, cnt);
        for (int i = 0; i < cnt; ++i) {
          C.sph++;
          if (sphere_hit(B.sphere_order[first + i], L.r, L.tbest)) L.best = first + i;
        }
        L.pending = 0;
      }
    }
    for (int l = 0; l < 32; ++l) {
      Lane &L = W[l];
      if (L.active && L.cur == DONE && !L.pending) L.active = false;
    }
    C.w_iters++;
    double c = C_LOOP + 6;
    if (any_node) c += C_NODE, C.w_node++, C.lanes_node += nl_node;
    if (any_leaf) c += C_LEAF0 + C_SPH * max_cnt, C.w_leaf++, C.w_leaf_sph += max_cnt, C.lanes_leaf += nl_leaf;
    if (any_pop) c += C_POP0 + C_POPIT * max_pop, C.w_pop++, C.w_popit += max_pop;
    C.cost += c;
    return;
  }
  // leaf phase
  int max_cnt = 0;
  int pend = 0, other = 0;
  for (int l = 0; l < 32; ++l) {
    Lane &L = W[l];
    if (!L.active) continue;
    if (L.cur < 0 && L.cur > POP) ++pend; else ++other;
  }
  const bool do_leaf = pend >= LEAF_T || other == 0;
  for (int l = 0; l < 32 && do_leaf; ++l) {
    Lane &L = W[l];
    if (!L.active || !(L.cur < 0 && L.cur > POP)) continue;
    any_leaf = true;
    ++nl_leaf;
    unsigned code = ~(unsigned)L.cur;
    int first = code & 0x3FFFFFF, cnt = ((code >> 26) & 15) + 1;
    if (ONE) {
      max_cnt = 1;
      C.sph++;
      if (sphere_hit(B.sphere_order[first], L.r, L.tbest)) L.best = first;
      if (cnt > 1) L.cur = ~(int)((unsigned)(first + 1) | ((unsigned)(cnt - 2) << 26)); else L.cur = POP;
      continue;
    }
    max_cnt = std::max(max_cnt, cnt);
    for (int i = 0; i < cnt; ++i) {
      C.sph++;
      if (sphere_hit(B.sphere_order[first + i], L.r, L.tbest)) L.best = first + i;
    }
    L.cur = POP;
  }
  // pop phase
  int max_pop = 0;
  for (int l = 0; l < 32; ++l) {
    Lane &L = W[l];
    if (!L.active || L.cur != POP) continue;
    any_pop = true;
    int it = 0;
    for (;;) {
      ++it;
      if (L.stk.empty()) {
        L.cur = DONE;
        L.active = false;
        break;
      }
      auto e = L.stk.back();
      L.stk.pop_back();
      C.pop++;
      if (e.second <= L.tbest) {
        L.cur = e.first;
        break;
      }
    }
    max_pop = std::max(max_pop, it);
  }
  C.w_iters++;
  double c = C_LOOP + extra;
  if (any_node) c += C_NODE, C.w_node++, C.lanes_node += nl_node;
  if (any_leaf) c += C_LEAF0 + C_SPH * max_cnt, C.w_leaf++, C.w_leaf_sph += max_cnt, C.lanes_leaf += nl_leaf;
  if (any_pop) c += C_POP0 + C_POPIT * max_pop, C.w_pop++, C.w_popit += max_pop;
  C.cost += c;
}

static Counts simulate(const std::vector<Ray> &rays) {
  Counts C;
  Lane W[32];
  size_t next = 0;
  C.rays = (double)rays.size();
  for (;;) {
    int act = 0;
    for (int l = 0; l < 32; ++l) act += W[l].active;
    bool more = next < rays.size();
    if (more && act < 32) {
      // flush + refill event
      C.w_flush++;
      C.cost += (DEFER ? C_PARK : C_FLUSH) + C_REFILL;
      for (int l = 0; l < 32 && next < rays.size(); ++l)
        if (!W[l].active) init_lane(W[l], rays[next++]), ++act;
    }
    if (act == 0) break;
    more = next < rays.size();
    int keep = more ? KEEP : 1;
    do {
      warp_iter(W, C);
      act = 0;
      for (int l = 0; l < 32; ++l) act += W[l].active;
    } while (act >= keep);
    if (!more && act == 0) {
      C.w_flush++, C.cost += DEFER ? C_PARK : C_FLUSH;
      break;
    }
  }
  if (DEFER) C.cost += C.rays / 32.0 * C_FLUSH;
  return C;
}

// ---- brute-force path tracer to produce per-bounce ray lists -------------------------------------
static int closest(const Ray &r, float &t) {
  t = 3.4e38f;
  int best = -1;
  for (int i = 0; i < (int)H.sr.size(); ++i)
    if (sphere_hit(i, r, t)) best = i;
  return best;
}

int main(int argc, char **argv) {
  int W = 480, Hh = 270;
  for (int i = 1; i < argc; ++i) {
    if (!strncmp(argv[i], "sort=", 5)) SORT = atoi(argv[i] + 5);
    if (!strncmp(argv[i], "keep=", 5)) KEEP = atoi(argv[i] + 5);
    if (!strncmp(argv[i], "leaft=", 6)) LEAF_T = atoi(argv[i] + 6);
    if (!strncmp(argv[i], "one=", 4)) ONE = atoi(argv[i] + 4);
    if (!strncmp(argv[i], "double=", 7)) DOUBLE = atoi(argv[i] + 7);
    if (!strncmp(argv[i], "cloop=", 6)) C_LOOP = atoi(argv[i] + 6);
    if (!strncmp(argv[i], "cnode=", 6)) C_NODE = atoi(argv[i] + 6);
    if (!strncmp(argv[i], "cflush=", 7)) C_FLUSH = atoi(argv[i] + 7);
    if (!strncmp(argv[i], "crefill=", 8)) C_REFILL = atoi(argv[i] + 8);
    if (!strncmp(argv[i], "csph=", 5)) C_SPH = atoi(argv[i] + 5);
    if (!strncmp(argv[i], "defer=", 6)) DEFER = atoi(argv[i] + 6);
    if (!strncmp(argv[i], "spec=", 5)) SPEC = atoi(argv[i] + 5);
    if (!strncmp(argv[i], "oct=", 4)) OCT = atoi(argv[i] + 4);
    if (!strncmp(argv[i], "fullsort=", 9)) FULLSORT = atoi(argv[i] + 9);
    if (!strncmp(argv[i], "w=", 2)) W = atoi(argv[i] + 2), Hh = W * 9 / 16;
  }
  FILE *f = fopen("/tmp/shirley_scene.bin", "rb");
  if (!f) return 1;
  int64_t n;
  double cam[4];
  if (fread(&n, 8, 1, f) != 1 || fread(cam, 8, 4, f) != 4) return 1;
  H.sx.resize(n), H.sy.resize(n), H.sz.resize(n), H.sr.resize(n), kind.resize(n);
  if (fread(H.sx.data(), 8, n, f) != (size_t)n || fread(H.sy.data(), 8, n, f) != (size_t)n ||
      fread(H.sz.data(), 8, n, f) != (size_t)n || fread(H.sr.data(), 8, n, f) != (size_t)n ||
      fread(kind.data(), 4, n, f) != (size_t)n)
    return 1;
  fclose(f);
  H.smat.assign(n, 0);
  build_wide_bvh(H, &B);
  int leaves = 0, leaf_prims = 0;
  for (auto &nd : B.nodes)
    for (int k = 0; k < 4; ++k)
      if (nd.child[k] < 0 && nd.child[k] != EMPTY_CHILD) ++leaves, leaf_prims += ((~(unsigned)nd.child[k] >> 26) & 15) + 1;
  printf("tree: %zu nodes depth %d max_stack %d leaves %d (%.2f prims/leaf)  sort=%d keep=%d\n", B.nodes.size(), B.depth,
         B.max_stack, leaves, (double)leaf_prims / leaves, SORT, KEEP);
  // camera rays in 32x32-ish tile order (row-major inside tiles of 32 wide)
  std::vector<std::vector<Ray>> per_bounce(8);
  std::mt19937 rng(1234);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  struct Path {
    Ray r;
  };
  std::vector<Ray> cur;
  for (int ty = 0; ty < Hh; ty += 32)
    for (int tx = 0; tx < W; tx += 32)
      for (int y = ty; y < std::min(ty + 32, Hh); ++y)
        for (int x = tx; x < std::min(tx + 32, W); ++x) {
          float cx = (x + U(rng)) / W, cy = 1.0f - (y + U(rng)) / Hh;
          F3 d = norm(F3{(float)(cam[0] + cam[2] * cx), (float)(cam[1] + cam[3] * cy), -1.0f});
          cur.push_back(Ray{{0, 0, 0}, {d.x, d.y, d.z}});
        }
  for (int b = 0; b < 8; ++b) {
    per_bounce[b] = cur;
    std::vector<Ray> nxt;
    for (const Ray &r : cur) {
      float t;
      int i = closest(r, t);
      if (i < 0) continue;
      F3 o = {r.o[0], r.o[1], r.o[2]}, d = {r.d[0], r.d[1], r.d[2]};
      F3 p = add(o, mul(d, t));
      F3 nrm = norm(sub(p, F3{(float)H.sx[i], (float)H.sy[i], (float)H.sz[i]}));
      bool front = dot(d, nrm) < 0;
      if (!front) nrm = mul(nrm, -1.f);
      F3 nd;
      if (kind[i] == 0) {  // lambert: cosine hemisphere around nrm
        float u = U(rng), v = U(rng), rr = std::sqrt(u), ph = 6.2831853f * v;
        F3 a = std::fabs(nrm.x) > 0.9f ? F3{0, 1, 0} : F3{1, 0, 0};
        F3 tx = norm(sub(a, mul(nrm, dot(a, nrm))));
        F3 ty = {nrm.y * tx.z - nrm.z * tx.y, nrm.z * tx.x - nrm.x * tx.z, nrm.x * tx.y - nrm.y * tx.x};
        nd = add(add(mul(tx, rr * std::cos(ph)), mul(ty, rr * std::sin(ph))), mul(nrm, std::sqrt(1 - u)));
      } else if (kind[i] == 1) {
        nd = sub(d, mul(nrm, 2 * dot(d, nrm)));
      } else {
        float eta = front ? 1.f / 1.5f : 1.5f, c = std::min(-dot(d, nrm), 1.f), s = std::sqrt(std::max(0.f, 1 - c * c));
        float r0 = (1 - eta) / (1 + eta);
        r0 *= r0;
        float sch = r0 + (1 - r0) * std::pow(1 - c, 5.f);
        if (eta * s > 1 || sch > U(rng))
          nd = sub(d, mul(nrm, 2 * dot(d, nrm)));
        else {
          F3 perp = mul(add(d, mul(nrm, c)), eta);
          nd = sub(perp, mul(nrm, std::sqrt(std::fabs(1 - dot(perp, perp)))));
        }
      }
      nd = norm(nd);
      F3 no = add(p, mul(nd, 1e-3f));
      nxt.push_back(Ray{{no.x, no.y, no.z}, {nd.x, nd.y, nd.z}});
    }
    cur.swap(nxt);
  }
  Counts T;
  printf("bounce  rays   node/ray sph/ray pop/ray | warp: cost/ray  iters/ray*32 lanes@node lanes@leaf sph/leafstep flush/ray*32\n");
  for (int b = 0; b < 8; ++b) {
    if (OCT && b > 0) {
      // emulate the producer: NPROD warps take 32-ray items round-robin; each keeps 8 open 128-ray segments
      const int NPROD = 64;
      std::vector<std::vector<Ray>> open((size_t)NPROD * 8);
      std::vector<Ray> out;
      const std::vector<Ray> &in = per_bounce[b];
      for (size_t i = 0; i < in.size(); ++i) {
        const int w = (int)((i / 32) % NPROD);
        const Ray &r = in[i];
        const int o = (r.d[0] < 0) | ((r.d[1] < 0) << 1) | ((r.d[2] < 0) << 2);
        auto &seg = open[(size_t)w * 8 + o];
        seg.push_back(r);
        if (seg.size() == 128) out.insert(out.end(), seg.begin(), seg.end()), seg.clear();
      }
      for (auto &seg : open) out.insert(out.end(), seg.begin(), seg.end());
      per_bounce[b] = out;
    }
    if (FULLSORT && b > 0) {
      auto key = [](const Ray &r) -> unsigned long long {
        auto q = [](float v, float lo, float hi) { float t = (v - lo) / (hi - lo); t = t < 0 ? 0 : (t > 0.999f ? 0.999f : t); return (unsigned)(t * 1024); };
        const unsigned x = q(r.o[0], -15, 15), y = q(r.o[1], -3, 3), z = q(r.o[2], -25, 5);
        unsigned long long m = 0;
        for (int b2 = 9; b2 >= 0; --b2) m = (m << 3) | (((x >> b2) & 1) << 2) | (((y >> b2) & 1) << 1) | ((z >> b2) & 1);
        const unsigned long long o = (r.d[0] < 0) | ((r.d[1] < 0) << 1) | ((r.d[2] < 0) << 2);
        return FULLSORT == 2 ? ((m >> 12) << 3 | o) : (o << 30 | m);  // 2: coarse cell first, then octant
      };
      std::stable_sort(per_bounce[b].begin(), per_bounce[b].end(), [&](const Ray &a, const Ray &c) { return key(a) < key(c); });
    }
    Counts C = simulate(per_bounce[b]);
    printf("%d %8.0f  %6.2f  %6.2f  %6.2f | %8.2f  %6.2f  %5.1f  %5.1f  %5.2f  %5.2f\n", b, C.rays, C.node / C.rays,
           C.sph / C.rays, C.pop / C.rays, C.cost / C.rays, C.w_iters / C.rays * 32, C.lanes_node / std::max(1.0, C.w_node),
           C.lanes_leaf / std::max(1.0, C.w_leaf), C.w_leaf_sph / std::max(1.0, C.w_leaf), C.w_flush / C.rays * 32);
    T.rays += C.rays, T.node += C.node, T.sph += C.sph, T.pop += C.pop, T.cost += C.cost;
  }
  printf("all %8.0f  %6.2f  %6.2f  %6.2f | %8.2f warp-instr/ray\n", T.rays, T.node / T.rays, T.sph / T.rays, T.pop / T.rays,
         T.cost / T.rays);
  return 0;
}

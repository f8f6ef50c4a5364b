// bvh.cpp — host-side tree construction for the device.
//
// Role in the reference: Shape_tree.Make(L).create (path_tracer/src/shape_tree.ml:252-263), a binary
// binned-SAH tree with 32 bins whose leaves hold <=16 SoA spheres (Simd_leaf) or small arrays.
// This is NOT that tree: the device wants few, fat, 128-bit-aligned nodes, so we build a binary
// binned-SAH tree (our own cost model and termination) and collapse it into a 4-wide BVH laid out
// breadth-first, so that the first K nodes are the top levels and can be staged in shared memory.
// Spheres and triangles get separate subtrees joined under the root, which keeps every leaf
// homogeneous (the reference's cornell `Shape` sum type, cornell-box/bin/main.ml:93-155, becomes a
// type bit in the leaf code).  Closest-hit results do not depend on the tree shape.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <thread>

#include "presplit.hpp"
#include "scene.hpp"

namespace ptb {
namespace {

struct PrimRef {
  Box box;
  double c[3];
  int32_t id;
};

struct BinNode {
  Box box;
  int lhs = -1, rhs = -1;  // children (binary inner)
  int first = 0, count = 0, type = 0;  // leaf
  bool leaf() const { return lhs < 0; }
};

constexpr int NBINS = 32;

// builder tuning (overridable for experiments: PTB_BVH_LEAF, PTB_BVH_CT, PTB_BVH_MINLEAF)
struct Tune {
  int leaf_max = LEAF_MAX;  // primitives per leaf (<= 16: the leaf code has 4 count bits)
  double ct = 1.2;          // cost of visiting a 4-wide node, in primitive tests
  int min_leaf = 2;         // ranges this small always become a leaf
  Tune() {
    if (const char *e = std::getenv("PTB_BVH_LEAF")) leaf_max = std::min(16, std::max(1, std::atoi(e)));
    if (const char *e = std::getenv("PTB_BVH_CT")) ct = std::atof(e);
    if (const char *e = std::getenv("PTB_BVH_MINLEAF")) min_leaf = std::max(1, std::atoi(e));
  }
};
static const Tune &tune() {
  static Tune t;
  return t;
}

// A range handed to a worker thread: its subtree is built into a private node vector and stitched in afterwards.
struct SubtreeTask {
  int node;    // placeholder in the shared node vector
  int lo, hi;  // primitive range
  std::vector<BinNode> local;
};

struct BinaryBuilder {
  std::vector<PrimRef> &prims;
  std::vector<BinNode> &nodes;
  int type;
  int base;  // offset of this type's slot range
  std::vector<SubtreeTask> *defer = nullptr;  // non-null: ranges of at most `defer_below` primitives become tasks
  int defer_below = 0;

  int build(int lo, int hi) {
    if (defer && hi - lo <= defer_below && hi - lo > 64) {
      int me = (int)nodes.size();
      nodes.emplace_back();
      defer->push_back(SubtreeTask{me, lo, hi, {}});
      return me;
    }
    const int n = hi - lo;
    // big ranges of the (serial) top phase: bounds and bins are reduced over chunks by a few threads
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());  // (a syscall: ask once)
    const int nchunk = (defer && n >= 65536) ? (int)std::min(hw, 16u) : 1;
    auto for_chunks = [&](auto &&f) {
      if (nchunk == 1) {
        f(lo, hi, 0);
        return;
      }
      std::vector<std::thread> th;
      for (int c = 1; c < nchunk; ++c)
        th.emplace_back([&, c] { f(lo + (int)((long long)n * c / nchunk), lo + (int)((long long)n * (c + 1) / nchunk), c); });
      f(lo, lo + n / nchunk, 0);
      for (auto &t : th) t.join();
    };
    struct Bounds {
      Box box, cbox;
    };
    Bounds pb1;
    std::vector<Bounds> pbv;
    if (nchunk > 1) pbv.resize((size_t)nchunk);
    Bounds *pb = nchunk > 1 ? pbv.data() : &pb1;
    for_chunks([&](int clo, int chi, int slot) {
      Bounds &B = pb[slot];
      B.box.reset(), B.cbox.reset();
      for (int i = clo; i < chi; ++i) {
        B.box.grow(prims[i].box);
        for (int a = 0; a < 3; ++a) {
          B.cbox.mn[a] = std::min(B.cbox.mn[a], prims[i].c[a]);
          B.cbox.mx[a] = std::max(B.cbox.mx[a], prims[i].c[a]);
        }
      }
    });
    Box box = pb[0].box, cbox = pb[0].cbox;
    for (int c = 1; c < nchunk; ++c) box.grow(pb[c].box), cbox.grow(pb[c].cbox);
    int me = (int)nodes.size();
    nodes.emplace_back();
    nodes[me].box = box;
    auto make_leaf = [&]() {
      nodes[me].first = lo;
      nodes[me].count = n;
      nodes[me].type = type;
      return me;
    };
    if (n <= tune().leaf_max && n <= tune().min_leaf) return make_leaf();
    // binned SAH over the three axes: one pass fills the bins of all three
    struct Bins {
      Box bb[3][NBINS];
      int cnt[3][NBINS];
    };
    double kk[3];
    bool use[3];
    for (int a = 0; a < 3; ++a) {
      const double ext = cbox.mx[a] - cbox.mn[a];
      use[a] = ext > 0.0;
      kk[a] = use[a] ? NBINS * (1.0 - 1e-9) / ext : 0.0;
    }
    Bins bins1;
    std::vector<Bins> binsv;
    if (nchunk > 1) binsv.resize((size_t)nchunk);
    Bins *bins = nchunk > 1 ? binsv.data() : &bins1;
    for_chunks([&](int clo, int chi, int slot) {
      Bins &B = bins[slot];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < NBINS; ++b) B.bb[a][b].reset(), B.cnt[a][b] = 0;
      for (int i = clo; i < chi; ++i)
        for (int a = 0; a < 3; ++a) {
          if (!use[a]) continue;
          int b = (int)(kk[a] * (prims[i].c[a] - cbox.mn[a]));
          b = std::min(std::max(b, 0), NBINS - 1);
          B.bb[a][b].grow(prims[i].box);
          B.cnt[a][b]++;
        }
    });
    for (int c = 1; c < nchunk; ++c)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < NBINS; ++b) bins[0].bb[a][b].grow(bins[c].bb[a][b]), bins[0].cnt[a][b] += bins[c].cnt[a][b];
    double best_cost = 1e300;
    int best_axis = -1, best_split = -1;
    for (int a = 0; a < 3; ++a) {
      if (!use[a]) continue;
      const Box *bb = bins[0].bb[a];
      const int *cnt = bins[0].cnt[a];
      double right_area[NBINS];
      int right_cnt[NBINS];
      Box acc;
      acc.reset();
      int c = 0;
      for (int b = NBINS - 1; b > 0; --b) {
        acc.grow(bb[b]);
        c += cnt[b];
        right_area[b] = acc.area();
        right_cnt[b] = c;
      }
      acc.reset();
      c = 0;
      for (int b = 0; b < NBINS - 1; ++b) {
        acc.grow(bb[b]);
        c += cnt[b];
        if (c == 0 || right_cnt[b + 1] == 0) continue;
        double cost = acc.area() * c + right_area[b + 1] * right_cnt[b + 1];
        if (cost < best_cost) best_cost = cost, best_axis = a, best_split = b;
      }
    }
    if (best_axis < 0) {
      // all centroids coincide: split by index while the leaf is too big
      if (n <= tune().leaf_max) return make_leaf();
      int mid = lo + n / 2;
      int l = build(lo, mid);
      int r = build(mid, hi);
      nodes[me].lhs = l, nodes[me].rhs = r;
      return me;
    }
    if (n <= tune().leaf_max) {
      // leaf cost n * Ci vs split cost Ct + sum; Ci = 1, Ct = 1.2 (a 4-wide node visit costs a few
      // primitive tests on the device)
      double split_cost = tune().ct + best_cost / std::max(box.area(), 1e-300);
      if (split_cost >= (double)n) return make_leaf();
    }
    double ext = cbox.mx[best_axis] - cbox.mn[best_axis];
    double k = NBINS * (1.0 - 1e-9) / ext;
    auto mid_it = std::partition(prims.begin() + lo, prims.begin() + hi, [&](const PrimRef &p) {
      int b = (int)(k * (p.c[best_axis] - cbox.mn[best_axis]));
      b = std::min(std::max(b, 0), NBINS - 1);
      return b <= best_split;
    });
    int mid = (int)(mid_it - prims.begin());
    if (mid == lo || mid == hi) mid = lo + n / 2;
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes[me].lhs = l, nodes[me].rhs = r;
    return me;
  }
};

// ---- triangle pre-splitting (presplit.hpp) ---------------------------------------------------------------------
// Replaces `tri` (one reference per triangle) by at most `max_factor` x as many references, each with the box of one
// piece of its triangle.  The cell size starts at the mean spacing of the triangles and grows until the budget holds.
static double presplit_triangles(const HostScene &s, std::vector<PrimRef> &tri, const Box &all, double max_factor) {
  const size_t n = tri.size();
  double vol = 1.0;
  for (int a = 0; a < 3; ++a) vol *= std::max(all.mx[a] - all.mn[a], 1e-12);
  double h = std::cbrt(vol / (double)n);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  auto run = [&](double cell, std::vector<PrimRef> *dst) -> size_t {
    const unsigned nt = (unsigned)std::min<size_t>(hw, std::max<size_t>(1, n / 4096));
    std::vector<std::vector<PrimRef>> parts(nt);
    std::vector<size_t> counts(nt, 0);
    std::vector<std::thread> th;
    auto work = [&](unsigned t) {
      size_t c = 0;
      for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) {
        double v[3][3];
        for (int k = 0; k < 3; ++k) {
          const int x = s.tidx[3 * (size_t)tri[i].id + k];
          v[k][0] = s.vx[x], v[k][1] = s.vy[x], v[k][2] = s.vz[x];
        }
        const int32_t id = tri[i].id;
        c += (size_t)presplit::pieces(v, cell, all.mn, [&](const double lo[3], const double hi[3]) {
          if (!dst) return;
          PrimRef r;
          for (int a = 0; a < 3; ++a) r.box.mn[a] = lo[a], r.box.mx[a] = hi[a], r.c[a] = 0.5 * (lo[a] + hi[a]);
          r.id = id;
          parts[t].push_back(r);
        });
      }
      counts[t] = c;
    };
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    size_t total = 0;
    for (unsigned t = 0; t < nt; ++t) total += counts[t];
    if (dst) {
      dst->clear();
      dst->reserve(total);
      for (auto &v : parts) dst->insert(dst->end(), v.begin(), v.end());
    }
    return total;
  };
  for (int it = 0; it < 40 && (double)run(h, nullptr) > max_factor * (double)n; ++it) h *= 1.1225;
  std::vector<PrimRef> refs;
  run(h, &refs);
  if (refs.size() < ((size_t)1 << 26)) tri.swap(refs);  // (a leaf code addresses 2^26 slots; beyond that the triangles stay whole)
  return h;
}

}  // namespace

void build_wide_bvh(const HostScene &s, WideBVH *out) {
  const bool timing = std::getenv("PTB_BVH_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto t_start = now();
  auto lap = [&](const char *what) {
    if (timing) std::fprintf(stderr, "[bvh] %-10s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now() - t_start).count());
    t_start = now();
  };
  out->nodes.clear();
  out->sphere_order.clear();
  out->tri_order.clear();
  out->presplit = false;
  std::vector<BinNode> bn;
  std::vector<PrimRef> sph((size_t)s.n_spheres()), tri((size_t)s.n_tris());
  for (size_t i = 0; i < sph.size(); ++i) {
    PrimRef &p = sph[i];
    double c[3] = {s.sx[i], s.sy[i], s.sz[i]};
    for (int a = 0; a < 3; ++a) p.box.mn[a] = c[a] - s.sr[i], p.box.mx[a] = c[a] + s.sr[i], p.c[a] = c[a];
    p.id = (int32_t)i;
  }
  for (size_t i = 0; i < tri.size(); ++i) {
    PrimRef &p = tri[i];
    p.box.reset();
    for (int k = 0; k < 3; ++k) {
      int v = s.tidx[3 * i + k];
      double c[3] = {s.vx[v], s.vy[v], s.vz[v]};
      for (int a = 0; a < 3; ++a) {
        p.box.mn[a] = std::min(p.box.mn[a], c[a]);
        p.box.mx[a] = std::max(p.box.mx[a], c[a]);
      }
    }
    for (int a = 0; a < 3; ++a) p.c[a] = 0.5 * (p.box.mn[a] + p.box.mx[a]);
    p.id = (int32_t)i;
  }
  if (tri.size() >= 2) {
    // how many triangle boxes contain a random point of the scene (presplit.hpp: far below 1 for surface meshes)
    Box all;
    all.reset();
    for (const PrimRef &p : tri) all.grow(p.box);
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) vol *= std::max(all.mx[a] - all.mn[a], 1e-12);
    // (one box counts for at most 64 mean cells: a tilted ground plane under a mesh must not pass for a soup)
    const double vcap = presplit::OVERLAP_BOX_CAP * vol / (double)tri.size();
    double vsum = 0.0;
    for (const PrimRef &p : tri)
      vsum += std::min(vcap, (p.box.mx[0] - p.box.mn[0]) * (p.box.mx[1] - p.box.mn[1]) * (p.box.mx[2] - p.box.mn[2]));
    double f = presplit::budget_factor(vsum / vol);
    if (const char *e = std::getenv("PTB_BVH_PRESPLIT")) f = std::atof(e);  // 0 / 1: off; > 1: references per triangle allowed
    // (small scenes are traversed out of shared memory and must stay small)
    if (f > 1.0 && tri.size() >= 4096 && (double)tri.size() * f < (double)(1 << 26)) {
      const size_t before = tri.size();
      const double h = presplit_triangles(s, tri, all, f);
      out->presplit = tri.size() > before;
      if (timing) std::fprintf(stderr, "[bvh] presplit: overlap %.2f, %zu -> %zu references, cell %.4g\n", vsum / vol, before, tri.size(), h);
    }
  }
  // Big inputs: the top of the tree is built on this thread down to ranges of about n / (8 x threads) primitives;
  // those subtrees are independent (disjoint primitive ranges, sorted in place) and are built by a pool of
  // threads into private node vectors, then appended with their indices shifted.
  auto build_all = [&](std::vector<PrimRef> &prims, int type) -> int {
    const int n = (int)prims.size();
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (n < 65536 || hw < 2) {
      BinaryBuilder b{prims, bn, type, 0};
      return b.build(0, n);
    }
    std::vector<SubtreeTask> tasks;
    BinaryBuilder top{prims, bn, type, 0, &tasks, std::max(4096, n / (int)(8 * hw))};
    const int root = top.build(0, n);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      for (size_t i; (i = next.fetch_add(1)) < tasks.size();) {
        BinaryBuilder b{prims, tasks[i].local, type, 0};
        b.build(tasks[i].lo, tasks[i].hi);
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < std::min<size_t>(hw, tasks.size()); ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    for (SubtreeTask &t : tasks) {
      // local[0] replaces the placeholder, local[1..] go to the end; child indices move accordingly
      const int shift = (int)bn.size() - 1;
      auto fix = [&](BinNode nd) {
        if (!nd.leaf()) nd.lhs += shift, nd.rhs += shift;
        return nd;
      };
      bn[t.node] = fix(t.local[0]);
      for (size_t j = 1; j < t.local.size(); ++j) bn.push_back(fix(t.local[j]));
    }
    return root;
  };
  lap("prims");
  int sroot = -1, troot = -1;
  if (!sph.empty()) sroot = build_all(sph, 0);
  if (!tri.empty()) troot = build_all(tri, 1);
  lap("binary");
  for (auto &p : sph) out->sphere_order.push_back(p.id);
  for (auto &p : tri) out->tri_order.push_back(p.id);

  // children of the (virtual) root
  std::vector<int> top;
  if (sroot >= 0) top.push_back(sroot);
  if (troot >= 0) top.push_back(troot);

  // collapse: BFS; each wide node gathers up to 4 binary descendants, always opening the inner
  // candidate with the largest surface area.
  struct Pending {
    std::vector<int> seeds;  // binary nodes whose union this wide node covers
    int depth;
  };
  std::queue<Pending> q;
  q.push({top, 1});
  int max_depth = 0;
  // wide node index is assigned in BFS order = push order
  std::vector<std::vector<int>> child_sets;
  std::vector<int> node_depth;
  while (!q.empty()) {
    Pending cur = q.front();
    q.pop();
    std::vector<int> ch = cur.seeds;
    // a single inner seed is opened unconditionally so that a wide node never has one inner child
    for (;;) {
      int pick = -1;
      double pa = -1.0;
      for (size_t i = 0; i < ch.size(); ++i)
        if (!bn[ch[i]].leaf() && bn[ch[i]].box.area() > pa) pa = bn[ch[i]].box.area(), pick = (int)i;
      if (pick < 0 || ch.size() >= 4) break;
      int n = ch[pick];
      ch[pick] = bn[n].lhs;
      ch.push_back(bn[n].rhs);
    }
    child_sets.push_back(ch);
    node_depth.push_back(cur.depth);
    max_depth = std::max(max_depth, cur.depth);
    for (int c : ch)
      if (!bn[c].leaf()) q.push({{bn[c].lhs, bn[c].rhs}, cur.depth + 1});
  }
  lap("collapse");
  // second pass: emit nodes; inner children get indices in the same BFS order
  out->nodes.resize(child_sets.size());
  int next_inner = 1;
  for (size_t i = 0; i < child_sets.size(); ++i) {
    WideNode &w = out->nodes[i];
    for (int k = 0; k < 4; ++k) {
      for (int a = 0; a < 3; ++a) w.mn[a][k] = 1e30, w.mx[a][k] = -1e30;
      w.child[k] = EMPTY_CHILD;
    }
    const auto &ch = child_sets[i];
    for (size_t k = 0; k < ch.size(); ++k) {
      const BinNode &b = bn[ch[k]];
      for (int a = 0; a < 3; ++a) w.mn[a][k] = b.box.mn[a], w.mx[a][k] = b.box.mx[a];
      if (b.leaf())
        w.child[k] = leaf_code(b.type, b.first, b.count);
      else
        w.child[k] = next_inner++;
    }
  }
  out->depth = max_depth;
  // exact worst case of the traversal stack: at a node with m children we push at most m-1 entries and
  // descend into one child, whose subtree then needs its own worst case on top (children have larger
  // indices than their parent, so one reverse sweep suffices)
  std::vector<int> need(out->nodes.size(), 0);
  for (int i = (int)out->nodes.size() - 1; i >= 0; --i) {
    int m = 0, deepest = 0;
    for (int k = 0; k < 4; ++k) {
      int c = out->nodes[i].child[k];
      if (c == EMPTY_CHILD) continue;
      ++m;
      if (c >= 0) deepest = std::max(deepest, need[c]);
    }
    need[i] = std::max(m - 1, 0) + deepest;
  }
  out->max_stack = need.empty() ? 1 : std::max(need[0], 1);
  lap("emit");
}

}  // namespace ptb

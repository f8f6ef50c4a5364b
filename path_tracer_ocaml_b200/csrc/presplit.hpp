// presplit.hpp — triangle pre-splitting ("early split clipping") for both tree builders (bvh.cpp on the host,
// gpu_bvh.cuh on the device).  Product code: must not include anything from oracle/.
//
// A triangle soup whose triangles are larger than the spacing between them gives every builder the same problem: the
// boxes of neighbouring leaves overlap many times over, and a ray has to enter every box that contains its origin
// before it can find its first hit (10^6 triangles of edge ~0.3 in [-10,10]^3: 16 boxes contain a random point; 49
// node visits and 40 triangle tests per ray).  Handing the builder several REFERENCES per triangle, each with the box
// of one piece of the triangle, removes most of that overlap: the triangle is cut — exactly, by polygon clipping —
// at planes of one global power-of-two grid hierarchy (so neighbouring triangles are cut at the same places) until no
// piece is longer than the cell size h on any axis.  Leaves then hold references to whole triangles: the closest hit
// is unchanged (a triangle may be tested twice, with the same result), the work per ray roughly halves at 3-4
// references per triangle (24 node visits, 14 triangle tests on the soup above).
//
// The pieces of one triangle form a binary tree (piece -> its two halves).  They are enumerated WITHOUT a stack: a
// piece is addressed by the path of left / right turns that leads to it, and is re-derived from the triangle along
// that path — a few clips more per piece, no per-thread stack of polygons on the device.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define PTB_HD __host__ __device__ __forceinline__
#else
#define PTB_HD inline
#endif

namespace ptb {
namespace presplit {

constexpr int MAX_VERTS = 10;  // a triangle cut by axis-aligned half-spaces: at most 3 + 6 corners
constexpr int MAX_DEPTH = 24;  // cuts along one path (2^24 cells across the largest triangle is far beyond any budget)

struct Poly {
  int n;
  double v[MAX_VERTS][3];
};

// the part of a convex polygon with x[axis] <= pos (side 0) or >= pos (side 1)
PTB_HD void clip(const Poly &in, int axis, double pos, int side, Poly *out) {
  int m = 0;
  for (int i = 0; i < in.n; ++i) {
    const double *a = in.v[i], *b = in.v[i + 1 == in.n ? 0 : i + 1];
    const double da = side ? a[axis] - pos : pos - a[axis], db = side ? b[axis] - pos : pos - b[axis];
    if (da >= 0 && m < MAX_VERTS) {
      out->v[m][0] = a[0], out->v[m][1] = a[1], out->v[m][2] = a[2];
      ++m;
    }
    if (((da > 0 && db < 0) || (da < 0 && db > 0)) && m < MAX_VERTS) {
      const double t = da / (da - db);
      for (int k = 0; k < 3; ++k) out->v[m][k] = a[k] + t * (b[k] - a[k]);
      out->v[m][axis] = pos;
      ++m;
    }
  }
  out->n = m;
}

// box of a polygon, clipped to `lo/hi` (the cell the polygon was cut to)
PTB_HD void box_of(const Poly &p, double lo[3], double hi[3]) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < p.n; ++i)
    for (int a = 0; a < 3; ++a) mn[a] = fmin(mn[a], p.v[i][a]), mx[a] = fmax(mx[a], p.v[i][a]);
  for (int a = 0; a < 3; ++a) lo[a] = fmax(lo[a], mn[a]), hi[a] = fmin(hi[a], mx[a]);
}

// Calls emit(lo, hi) for the box of every piece of the triangle (v[corner][axis]); returns the number of pieces (>= 1).
// h: cell size; org: origin of the grid hierarchy (any fixed point, e.g. the scene's lower corner).
template <class Emit>
PTB_HD int pieces(const double v[3][3], double h, const double org[3], Emit &&emit) {
  // One triangle never becomes more than a few dozen pieces: a triangle much longer than the cell (a ground plane under
  // a soup) is cut on a coarser level of the same grid hierarchy, at most 8 cells along its longest side.
  {
    double e = 0.0;
    for (int a = 0; a < 3; ++a) e = fmax(e, fmax(fmax(v[0][a], v[1][a]), v[2][a]) - fmin(fmin(v[0][a], v[1][a]), v[2][a]));
    for (int k = 0; k < 60 && e > 8.0 * h; ++k) h *= 2.0;
  }
  const double finest = h * (1.0 / 64.0);
  int count = 0;
  uint32_t path = 0;  // bit d: the turn taken at depth d (0 = lower side)
  int depth = 0;      // turns of `path` that are fixed; below them the enumeration keeps to the lower side
  for (;;) {
    Poly p, q;
    p.n = 3;
    double lo[3] = {-1e300, -1e300, -1e300}, hi[3] = {1e300, 1e300, 1e300};
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) p.v[k][a] = v[k][a];
    box_of(p, lo, hi);
    int d = 0;
    bool empty = false;
    for (;; ++d) {
      int ax = 0;
      double ext = hi[0] - lo[0];
      for (int a = 1; a < 3; ++a)
        if (hi[a] - lo[a] > ext) ext = hi[a] - lo[a], ax = a;
      if (!(ext > h) || d >= MAX_DEPTH) break;  // a piece
      // the coarsest plane of the grid hierarchy strictly inside the box (the midpoint if rounding hides them all)
      double pos = 0.5 * (lo[ax] + hi[ax]);
      for (double s = h * 1048576.0; s >= finest; s *= 0.5) {
        double c = (floor((lo[ax] - org[ax]) / s) + 1.0) * s + org[ax];  // the first plane above the lower face
        if (!(c > lo[ax] + 1e-9 * s)) c += s;
        if (c < hi[ax] - 1e-9 * s) {
          pos = c;
          break;
        }
      }
      const int side = d < depth ? (int)((path >> d) & 1u) : 0;
      if (d >= depth) path &= ~(1u << d);
      clip(p, ax, pos, side, &q);
      if (q.n < 3) {  // nothing of the triangle on this side (it only touches the plane)
        empty = true;
        ++d;
        break;
      }
      p = q;
      if (side) lo[ax] = pos;
      else hi[ax] = pos;
      box_of(p, lo, hi);
    }
    if (!empty) {
      emit(lo, hi);
      ++count;
    }
    // next piece: the deepest turn to the lower side becomes a turn to the upper side
    int j = d - 1;
    while (j >= 0 && ((path >> j) & 1u)) --j;
    if (j < 0) break;
    path = (path | (1u << j)) & ((2u << j) - 1u);
    depth = j + 1;
  }
  if (count == 0) {  // numerically degenerate (a sliver lost between two clips): keep the triangle whole
    Poly p;
    p.n = 3;
    double lo[3] = {-1e300, -1e300, -1e300}, hi[3] = {1e300, 1e300, 1e300};
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) p.v[k][a] = v[k][a];
    box_of(p, lo, hi);
    emit(lo, hi);
    count = 1;
  }
  return count;
}

// Pre-split when the boxes of the triangles overlap: `overlap` = sum of the box volumes / volume of the scene's box =
// how many boxes contain a random point.  Surface meshes are far below 1, soups of big triangles far above.
// A single box enters the sum with at most OVERLAP_BOX_CAP mean cells (scene volume / triangles), so that a few huge
// triangles — a tilted ground plane under a mesh — do not make a surface look like a soup.
// Returns the references per triangle the builder may spend (1 = leave the triangles whole).
constexpr double OVERLAP_BOX_CAP = 64.0;
PTB_HD double budget_factor(double overlap) { return overlap <= 0.5 ? 1.0 : (overlap >= 5.0 ? 6.0 : 1.0 + overlap); }

}  // namespace presplit
}  // namespace ptb

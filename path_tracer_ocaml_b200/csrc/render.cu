// render.cu — device state, launch orchestration and the compute entry points of the C ABI.
// Product code: nothing from oracle/, no CPU fallback (every entry point needs a CUDA device).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "handle.hpp"
#include "kernels.cuh"
#include "gpu_bvh.cuh"
#include "ray_sort.cuh"

namespace ptb {

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(PTB_E_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e_));     \
  } while (0)

template <class R>
struct Tables {
  Node4<R> *nodes = nullptr;
  Vec4<R> *spheres = nullptr, *tris = nullptr;
  R *tri_uv = nullptr;
  DTex<R> *texs = nullptr;
  NodeG *nodes_g = nullptr;  // float scenes that stay in global memory (make_dscene): 256-bit load records
  TriG *tris_g = nullptr;
  Vec4<R> *spheres_g = nullptr;
  NodeQ *nodes_q = nullptr;
  bool ready = false;
  bool in_blob = false;  // the pointers are slices of DeviceState::blob (freed with it)
};
// producer warps that may each leave one partially filled segment behind in a queue (upper bound over
// every launch shape used here), i.e. the slack a segmented queue needs on top of its dense capacity
constexpr size_t MAX_PRODUCER_WARPS = 8192;
template <class R>
struct Work {
  Queue<R> rays{nullptr, nullptr};
  Queue<R> mq{nullptr, nullptr};  // the NUM_MAT_KINDS hit queues: equal slices of one allocation
  size_t slots = 0;                                   // entries per queue (a multiple of SEG)
  size_t cap = 0;                                     // rays per batch the queues were sized for
};

// Per-device work buffers (wavefront queues, control block, pixel list).  They belong to the device,
// not to a scene: committing another scene must not reallocate gigabytes of queue memory.
struct DevicePool {
  // Serialises every entry point that uses this device's work buffers (ptb200.h "Threading"): the OCaml stubs release
  // the runtime lock and ctypes drops the GIL, so two host threads can be inside the library at once.  Recursive:
  // ptb_intersect_batch calls ptb_intersect_batch_device, a commit of a replica happens inside commit_multi.
  std::recursive_mutex mu;
  Work<float> wf;
  Work<double> wd;
  Ctl *ctl = nullptr;
  // polled progress (ptb_render_progress): a zero-copy host word the batch-control kernel stores into
  unsigned long long *progress_host = nullptr, *progress_dev = nullptr;
  unsigned long long progress_total = 0;
  // cached device properties (cudaGetDeviceProperties costs milliseconds; a re-commit must not pay it again)
  int sm_count = 0;
  size_t smem_optin = 0;
  // pinned staging: one block for the scene tables of a commit, two chunk buffers for ptb_intersect_batch
  void *stage_host = nullptr;
  size_t stage_cap = 0;
  // ptb_render's device image buffers (per-pixel sums and the resolved image), grown on demand and kept:
  // allocating and freeing ~300 MB per call costs tens to hundreds of ms of host time on a busy allocator
  void *sums_buf = nullptr, *img_buf = nullptr;
  size_t sums_cap = 0, img_cap = 0;
  void *batch_buf[4] = {nullptr, nullptr, nullptr, nullptr};  // ptb_intersect_batch: origins, directions, t, prim
  size_t batch_cap[4] = {0, 0, 0, 0};
  void *sort_buf = nullptr;  // ptb_intersect_batch on global-memory scenes: keys + bins of the ray sort (ray_sort.cuh)
  size_t sort_cap = 0;
  int32_t *pixel_list = nullptr;
  std::vector<int32_t> pixel_list_host;
  int pl_W = 0, pl_H = 0, pl_rank = -1, pl_world = 0, npix = 0;
  template <class R>
  Work<R> &work();
};
template <>
Work<float> &DevicePool::work<float>() { return wf; }
template <>
Work<double> &DevicePool::work<double>() { return wd; }
static DevicePool *pool_for(int device) {
  static DevicePool pools[64];
  return &pools[device & 63];
}
using PoolLock = std::lock_guard<std::recursive_mutex>;

struct DeviceState {
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
  int32_t *sphere_id = nullptr, *tri_id = nullptr, *sphere_mat = nullptr, *tri_mat = nullptr;
  DMat *mats = nullptr;
  uint8_t *prim_kind = nullptr;
  // host-built trees: every float-path table above and in `tf` is a slice of ONE allocation filled by ONE copy
  void *blob = nullptr;
  bool ids_in_blob = false;
  Tables<float> tf;
  Tables<double> td;
  DevicePool *pool = nullptr;
  template <class R>
  Tables<R> &tables();
};
template <>
Tables<float> &DeviceState::tables<float>() { return tf; }
template <>
Tables<double> &DeviceState::tables<double>() { return td; }

// Scene tables come from the device's stream-ordered memory pool, told to keep what is freed: a re-commit then
// costs microseconds.  cudaMalloc/cudaFree are device-wide synchronisations with page-table updates, and with
// tens of GB of wavefront queues in the context the dozen of them a commit needs took 250-290 ms.
static cudaError_t tbl_alloc(void **p, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  static bool configured[64] = {};
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    configured[dev] = true;
  }
  return cudaMallocAsync(p, std::max<size_t>(bytes, 16), 0);
}
static void tbl_free(void *p) {
  if (p) cudaFreeAsync(p, 0);
}
template <class R>
static void free_tables(Tables<R> &t) {
  if (!t.in_blob) tbl_free(t.nodes), tbl_free(t.spheres), tbl_free(t.tris), tbl_free(t.tri_uv), tbl_free(t.texs);
  tbl_free(t.nodes_g), tbl_free(t.tris_g), tbl_free(t.spheres_g), tbl_free(t.nodes_q);
  t = Tables<R>();
}
template <class R>
static void free_queue(Queue<R> &q) {
  cudaFree(q.base), cudaFree(q.seg_count);
  q = Queue<R>{nullptr, nullptr};
}
template <class R>
static void free_work(Work<R> &w) {
  free_queue(w.rays);
  free_queue(w.mq);
  w.cap = 0, w.slots = 0;
}
// releases every table of `d` (stream-ordered frees: cheap, the pool keeps the memory) but keeps the state object
static void release_device_tables(DeviceState *d) {
  if (!d->ids_in_blob) {
    tbl_free(d->sphere_id), tbl_free(d->tri_id), tbl_free(d->sphere_mat), tbl_free(d->tri_mat);
    tbl_free(d->mats), tbl_free(d->prim_kind);
  }
  d->sphere_id = d->tri_id = d->sphere_mat = d->tri_mat = nullptr, d->mats = nullptr, d->prim_kind = nullptr;
  free_tables(d->tf), free_tables(d->td);
  tbl_free(d->blob);
  d->blob = nullptr, d->ids_in_blob = false;
}
void destroy_device_state(DeviceState *d) {
  if (!d) return;
  if (d->device >= 0) cudaSetDevice(d->device);
  PoolLock lk(pool_for(d->device)->mu);
  cudaDeviceSynchronize();  // nothing on any stream may still be reading the tables
  release_device_tables(d);
  delete d;
}

static int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(PTB_E_NO_DEVICE, std::string("no CUDA device: libptb200 has no CPU path (") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + ")");
  if (device < 0 || device >= n) return fail(PTB_E_INVALID, "device ordinal out of range");
  CK(cudaSetDevice(device));
  return PTB_OK;
}

template <class T>
static int upload(T **dst, const std::vector<T> &src) {
  *dst = nullptr;
  size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
  CK(tbl_alloc((void **)dst, bytes));
  if (!src.empty()) CK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return PTB_OK;
}

// conservative narrowing of box planes: the float box must contain the float primitives
template <class R>
static R box_lo(double x, double ext);
template <class R>
static R box_hi(double x, double ext);
template <>
double box_lo<double>(double x, double) { return x; }
template <>
double box_hi<double>(double x, double) { return x; }
template <>
float box_lo<float>(double x, double ext) {
  if (x > 1e29) return 1e30f;
  float f = (float)(x - 4e-7 * (std::fabs(x) + ext));
  return std::nextafterf(f, -INFINITY);
}
template <>
float box_hi<float>(double x, double ext) {
  if (x < -1e29) return -1e30f;
  float f = (float)(x + 4e-7 * (std::fabs(x) + ext));
  return std::nextafterf(f, INFINITY);
}

template <class R>
struct HostTables {
  std::vector<Node4<R>> nodes;
  std::vector<Vec4<R>> sph, tri;
  std::vector<R> uv;
  std::vector<DTex<R>> texs;
};
template <class R>
static void make_host_tables(const ptb_scene *s, HostTables<R> *out) {
  const HostScene &h = s->host;
  const WideBVH &b = s->bvh;
  std::vector<Node4<R>> &nodes = out->nodes;
  nodes.resize(b.nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) {
    for (int k = 0; k < 4; ++k) {
      double ext = 0;
      for (int a = 0; a < 3; ++a) ext = std::max(ext, b.nodes[i].mx[a][k] - b.nodes[i].mn[a][k]);
      for (int a = 0; a < 3; ++a) {
        nodes[i].lo[a][k] = box_lo<R>(b.nodes[i].mn[a][k], ext);
        nodes[i].hi[a][k] = box_hi<R>(b.nodes[i].mx[a][k], ext);
      }
      nodes[i].child[k] = b.nodes[i].child[k];
    }
  }
  std::vector<Vec4<R>> &sph = out->sph;
  sph.resize(b.sphere_order.size());
  for (size_t k = 0; k < sph.size(); ++k) {
    int i = b.sphere_order[k];
    sph[k] = {(R)h.sx[i], (R)h.sy[i], (R)h.sz[i], (R)h.sr[i]};
  }
  std::vector<Vec4<R>> &tri = out->tri;
  std::vector<R> &uv = out->uv;
  tri.resize(3 * b.tri_order.size());
  uv.resize(6 * b.tri_order.size());
  for (size_t k = 0; k < b.tri_order.size(); ++k) {
    int i = b.tri_order[k];
    int ia = h.tidx[3 * i], ib = h.tidx[3 * i + 1], ic = h.tidx[3 * i + 2];
    tri[3 * k + 0] = {(R)h.vx[ia], (R)h.vy[ia], (R)h.vz[ia], R(0)};
    tri[3 * k + 1] = {(R)(h.vx[ib] - h.vx[ia]), (R)(h.vy[ib] - h.vy[ia]), (R)(h.vz[ib] - h.vz[ia]), R(0)};
    tri[3 * k + 2] = {(R)(h.vx[ic] - h.vx[ia]), (R)(h.vy[ic] - h.vy[ia]), (R)(h.vz[ic] - h.vz[ia]), R(0)};
    for (int c = 0; c < 6; ++c) uv[6 * k + c] = (R)h.tuv[6 * i + c];
  }
  std::vector<DTex<R>> &texs = out->texs;
  texs.resize(h.tex.size());
  for (size_t i = 0; i < texs.size(); ++i) {
    const ptb_texture &x = h.tex[i];
    texs[i] = {x.kind, x.width, x.height, x.even, x.odd, 0, {(R)x.rgb[0], (R)x.rgb[1], (R)x.rgb[2]}};
  }
}

// The float tables are uploaded by the commit (upload_scene: one blob, one copy).  The float64 tables of the
// validation mode are made on first use, table by table.
template <class R>
static int ensure_tables(ptb_scene *s) {
  DeviceState *d = s->dev;
  Tables<R> &t = d->tables<R>();
  if (t.ready) return PTB_OK;
  if (s->bvh.device_built)
    return fail(PTB_E_UNSUPPORTED, "this scene's tree was built on the device (float32 tables only): use PTB_BUILDER=host for PTB_FLAG_F64");
  HostTables<R> ht;
  make_host_tables<R>(s, &ht);
  int rc;
  if ((rc = upload(&t.nodes, ht.nodes))) return rc;
  if ((rc = upload(&t.spheres, ht.sph))) return rc;
  if ((rc = upload(&t.tris, ht.tri))) return rc;
  if ((rc = upload(&t.tri_uv, ht.uv))) return rc;
  if ((rc = upload(&t.texs, ht.texs))) return rc;
  t.ready = true;
  return PTB_OK;
}

template <class R>
static int ensure_work(DevicePool *d, size_t cap) {
  Work<R> &w = d->work<R>();
  if (w.cap >= cap) return PTB_OK;
  free_work(w);
  // (a producer warp may leave one partly filled and SEG_GROUP - 1 unopened segments behind per queue)
  const size_t segs = (cap + SEG - 1) / SEG + SEG_GROUP * MAX_PRODUCER_WARPS, slots = segs * SEG;
  auto alloc_q = [&](Queue<R> &q, size_t n) -> int {
    CK(cudaMalloc((void **)&q.base, 3 * n * slots * sizeof(Vec4<R>)));  // segment-interleaved A, B, C (device_types.cuh)
    CK(cudaMalloc((void **)&q.seg_count, n * segs * sizeof(int32_t)));
    return PTB_OK;
  };
  int rc;
  if ((rc = alloc_q(w.rays, 1))) return rc;
  if ((rc = alloc_q(w.mq, NUM_MAT_KINDS))) return rc;
  w.cap = cap, w.slots = slots;
  return PTB_OK;
}

// this rank's pixels, in the reference's tile order (Tile.split ~max_area:1024, integrator.ml:132-133;
// tile t belongs to rank t mod world), row-major inside a tile like Tile.iter (tile.ml:71-79)
static int ensure_pixel_list(DevicePool *d, int W, int H, int rank, int world) {
  if (d->pixel_list && d->pl_W == W && d->pl_H == H && d->pl_rank == rank && d->pl_world == world) return PTB_OK;
  std::vector<TileRect> tiles;
  tile_split(W, H, 32 * 32, &tiles);
  std::vector<int32_t> &pl = d->pixel_list_host;
  pl.clear();
  for (size_t t = 0; t < tiles.size(); ++t) {
    if ((int)(t % (size_t)world) != rank) continue;
    const TileRect &r = tiles[t];
    for (int y = 0; y < r.h; ++y)
      for (int x = 0; x < r.w; ++x) pl.push_back((r.row + y) * W + (r.col + x));
  }
  cudaFree(d->pixel_list);
  d->pixel_list = nullptr;
  int rc = upload(&d->pixel_list, pl);
  if (rc) return rc;
  d->pl_W = W, d->pl_H = H, d->pl_rank = rank, d->pl_world = world, d->npix = (int)pl.size();
  return PTB_OK;
}

// passes rendered by a call (ptb_params.pass_first / pass_count; {0, 0} = all)
static int passes_of(const ptb_params &p) { return p.pass_count > 0 ? p.pass_count : p.samples_per_pixel - p.pass_first; }

static int fill_render_const(const ptb_params &p, int npix, RenderConst *rc) {
  if (p.width <= 0 || p.height <= 0 || p.samples_per_pixel <= 0)
    return fail(PTB_E_INVALID, "render: width, height and samples_per_pixel must be positive");
  if (p.max_bounces < 0 || p.max_bounces > MAX_BOUNCES)
    return fail(PTB_E_INVALID, "render: max_bounces must be in [0, 64]");
  if ((long long)p.width * p.height + (long long)p.samples_per_pixel * p.samples_per_pixel >= (1LL << 31))
    return fail(PTB_E_INVALID, "render: sample offset would overflow int32");
  if (p.pass_first < 0 || p.pass_count < 0 || p.pass_first >= p.samples_per_pixel ||
      (long long)p.pass_first + p.pass_count > p.samples_per_pixel)
    return fail(PTB_E_INVALID, "render: pass range outside [0, samples_per_pixel)");
  std::memset(rc, 0, sizeof *rc);
  rc->W = p.width, rc->H = p.height, rc->spp = p.samples_per_pixel, rc->max_bounces = p.max_bounces;
  rc->npix = npix;
  rc->llx = p.lower_left_x, rc->lly = p.lower_left_y, rc->vx = p.view_x, rc->vy = p.view_y;
  rc->widthf = 1.0 / (double)p.width;
  rc->heightf = 1.0 / (double)p.height;
  lds_alpha(2 + 2 * p.max_bounces, rc->alpha);  // Integrator.create_sampler (integrator.ml:89)
  return PTB_OK;
}

static GenConst make_gen(const RenderConst &rc, const int32_t *pixel_list, int pass0, int i0) {
  GenConst g;
  std::memset(&g, 0, sizeof g);
  g.W = rc.W, g.spp = rc.spp, g.npix = rc.npix, g.pass0 = pass0, g.i0 = i0;
  g.llx = rc.llx, g.lly = rc.lly, g.vx = rc.vx, g.vy = rc.vy, g.widthf = rc.widthf, g.heightf = rc.heightf;
  g.alpha0 = rc.alpha[0], g.alpha1 = rc.alpha[1];
  g.inv_npix = (1.0 / (double)rc.npix) * (1.0 - 0x1p-40), g.inv_W = (1.0 / (double)rc.W) * (1.0 - 0x1p-40);
  g.pixel_list = pixel_list;
  return g;
}

// want_records: the caller runs the render pipeline, whose global-memory traversal reads the 256-bit load records
template <class R>
static DScene<R> make_dscene(ptb_scene *s, size_t *scene_bytes_out, bool want_records = false) {
  DeviceState *d = s->dev;
  Tables<R> &t = d->tables<R>();
  DScene<R> sc;
  std::memset(&sc, 0, sizeof sc);
  sc.nodes = t.nodes, sc.spheres = t.spheres, sc.tris = t.tris;
  sc.sphere_id = d->sphere_id, sc.tri_id = d->tri_id, sc.sphere_mat = d->sphere_mat, sc.tri_mat = d->tri_mat;
  sc.tri_uv = t.tri_uv, sc.mats = d->mats, sc.texs = t.texs;
  sc.prim_kind = d->prim_kind;
  sc.n_nodes = (int)s->bvh.node_count();
  sc.n_spheres = (int)s->bvh.sphere_order.size();
  sc.n_tris = (int)s->bvh.tri_count();
  size_t bytes = (size_t)sc.n_nodes * sizeof(Node4<R>) + (size_t)sc.n_spheres * sizeof(Vec4<R>) +
                 (size_t)sc.n_tris * 3 * sizeof(Vec4<R>) + (((size_t)sc.n_spheres + sc.n_tris + 15) / 16) * 16;
  // entry 0 of the traversal stack is the sentinel.  A tree whose worst case fits gets an unchecked stack
  // (shared-memory scenes); deeper trees run the global-memory variant, which guards every push.
  sc.stack_cap = std::min(std::max(s->bvh.max_stack + 1, 4), 97);
  sc.scene_in_smem = (bytes <= 100 * 1024 && s->bvh.max_stack + 1 <= sc.stack_cap) ? 1 : 0;
  sc.bg_kind = s->host.bg_kind;
  for (int i = 0; i < 3; ++i)
    sc.bg0[i] = (R)s->host.bg0[i], sc.bg1[i] = (R)s->host.bg1[i], sc.bgd[i] = (R)(s->host.bg1[i] - s->host.bg0[i]);
  sc.has_light = s->host.has_light ? 1 : 0;
  sc.has_emissive = s->host.has_emissive() ? 1 : 0;
  for (int i = 0; i < 3; ++i)
    sc.light_o[i] = (R)s->host.light_o[i], sc.light_u[i] = (R)s->host.light_u[i], sc.light_v[i] = (R)s->host.light_v[i];
  *scene_bytes_out = sc.scene_in_smem ? bytes : 0;
  if constexpr (sizeof(R) == 4) {
    if (!sc.scene_in_smem) {  // first use of a committed global-memory scene: its 256-bit load records
      // (+2 triangle records: the fetch of a leaf's last triangle reads the record after it as well)
      if (!t.nodes_g && tbl_alloc((void **)&t.nodes_g, (size_t)sc.n_nodes * sizeof(NodeG)) == cudaSuccess &&
          tbl_alloc((void **)&t.tris_g, ((size_t)sc.n_tris + 2) * sizeof(TriG)) == cudaSuccess &&
          tbl_alloc((void **)&t.spheres_g, ((size_t)sc.n_spheres + 8) * sizeof(Vec4<R>)) == cudaSuccess &&
          tbl_alloc((void **)&t.nodes_q, (size_t)sc.n_nodes * sizeof(NodeQ)) == cudaSuccess) {
        const int n = std::max(std::max(sc.n_nodes, sc.n_tris), sc.n_spheres + 8);
        k_make_g_layout<<<(unsigned)((n + 255) / 256), 256>>>(t.nodes, sc.n_nodes, t.tris, sc.n_tris, t.spheres, sc.n_spheres,
                                                              t.nodes_g, t.tris_g, t.spheres_g, t.nodes_q);
        if (cudaStreamSynchronize(0) != cudaSuccess)
          tbl_free(t.nodes_g), tbl_free(t.tris_g), tbl_free(t.spheres_g), tbl_free(t.nodes_q), t.nodes_g = nullptr,
              t.tris_g = nullptr, t.spheres_g = nullptr, t.nodes_q = nullptr;
      }
      const bool ok = t.nodes_g && t.tris_g && t.spheres_g && t.nodes_q;  // (null: trace_config reports the failure)
      sc.nodes_g = ok ? t.nodes_g : nullptr, sc.tris_g = t.tris_g, sc.spheres_g = t.spheres_g, sc.nodes_q = t.nodes_q;
    }
  }
  return sc;
}

struct TraceLaunch {
  int block = 256, grid = 0;
  size_t smem = 0;
  bool scene_smem = false;
};
// Launch shape of the traversal kernel.  Scene in shared memory: ONE block per SM, as many warps as the
// per-thread stacks leave room for (up to 32), so the scene is staged once per SM.  Otherwise 256-thread
// blocks at whatever occupancy the stack allows.
template <class R, int MODE>
static int trace_config(DeviceState *d, const DScene<R> &sc, size_t scene_bytes, TraceLaunch *tl) {
  const size_t per_thread = trace_smem_per_thread<R>(sc.stack_cap, sc.scene_in_smem != 0);  // (stack +) payload + warp record
  tl->scene_smem = sc.scene_in_smem != 0;
  if (sizeof(R) == 4 && !tl->scene_smem && !sc.nodes_g)
    return fail(PTB_E_NOMEM, "trace: could not build the global-memory scene records (NodeG / TriG)");
  if (tl->scene_smem) {
    tl->block = sizeof(R) == 8 ? 512 : 1024;  // = the kernel's __launch_bounds__
    if (const char *e = std::getenv("PTB_TRACE_BLOCK")) tl->block = std::min(tl->block, std::max(128, std::atoi(e) / 32 * 32));
    while (tl->block > 128 && scene_bytes + per_thread * tl->block > d->smem_optin) tl->block -= 128;
    tl->smem = scene_bytes + per_thread * tl->block;
    if (tl->smem > d->smem_optin) return fail(PTB_E_NOMEM, "trace: scene + stack do not fit shared memory");
    CK(cudaFuncSetAttribute(k_trace<R, MODE, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl->smem));
    if (MODE == 0)
      CK(cudaFuncSetAttribute(k_trace<R, MODE, true, MODE == 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl->smem));
    if (sizeof(R) == 4 && tl->block == 1024) {  // the block size is a compile-time constant in this instantiation
      CK(cudaFuncSetAttribute(k_trace<R, MODE, true, false, sizeof(R) == 4 ? 1024 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl->smem));
      if (MODE == 0)
        CK(cudaFuncSetAttribute(k_trace<R, MODE, true, MODE == 0, sizeof(R) == 4 ? 1024 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl->smem));
    }
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<R, MODE, true, false>, tl->block, tl->smem));
    if (per_sm < 1) return fail(PTB_E_CUDA, "trace: kernel cannot be resident");
    tl->grid = d->sm_count * per_sm;
  } else {
    tl->block = 256;
    while (tl->block > 64 && per_thread * tl->block > d->smem_optin / 2) tl->block /= 2;
    tl->smem = per_thread * tl->block;
    CK(cudaFuncSetAttribute(k_trace<R, MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl->smem));
    if (MODE == 0)
      CK(cudaFuncSetAttribute(k_trace<R, MODE, false, MODE == 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl->smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<R, MODE, false, false>, tl->block, tl->smem));
    if (per_sm < 1) return fail(PTB_E_CUDA, "trace: kernel cannot be resident");
    tl->grid = d->sm_count * per_sm;
  }
  return PTB_OK;
}
// A warp flushes its finished lanes and refills them when fewer than this many lanes are still traversing.
static int refill_below(bool scene_smem, bool batch_mode) {
  static int env = -2;
  if (env == -2) {
    env = -1;
    if (const char *e = std::getenv("PTB_REFILL")) env = std::min(32, std::max(1, std::atoi(e)));
  }
  if (env > 0) return env;
  // Render pipeline and scenes in shared memory: measured flat between 8 and 14, slower above (a flush + refill of the
  // render pipeline is ~200 instructions per warp; the C3 mesh renders lose 5-7 % at 26).  ptb_intersect_batch on a
  // scene in global memory: every traversal step waits on a gather, the flush is two stores, idle lanes are what costs:
  // refilling below 26 lanes gives +16 % / +21 % on 10^5 / 10^4 triangles, +13-18 % on 4096 spheres and +3 % on the
  // pre-split 10^6 soup (incoherent rays).
  return (!scene_smem && batch_mode) ? 26 : 10;
}
// gen != nullptr: bounce 0, the kernel generates the camera rays [0, gen_n) of the batch itself
template <class R, int MODE>
static void launch_trace(const TraceLaunch &tl, cudaStream_t st, const DScene<R> &sc, const GenConst *gen, unsigned gen_n,
                         Queue<R> rays, const unsigned *nseg_ptr, unsigned nseg_imm, unsigned *cursor, Queue<R> mq,
                         unsigned mq_slots, unsigned *nseg_mat, unsigned *n_traced, int enqueue_hits, R *sums, R tmin, R tmax, R *out_t,
                         int32_t *out_prim) {
  GenConst g;
  std::memset(&g, 0, sizeof g);
  if (gen) g = *gen;
#define PTB_LAUNCH(SM, GN, BK)                                                                                         \
  k_trace<R, MODE, SM, GN, BK><<<tl.grid, tl.block, tl.smem, st>>>(sc, g, gen_n, rays, nseg_ptr, nseg_imm, cursor,    \
                                                                   refill_below(tl.scene_smem, MODE == 1), mq, mq_slots, nseg_mat,                \
                                                                   n_traced, enqueue_hits, sums, tmin, tmax, out_t, out_prim)
  constexpr int FIXED = sizeof(R) == 4 ? 1024 : 0;  // float, shared-memory scene, full block: compile-time block size
  const bool fixed = FIXED != 0 && tl.scene_smem && tl.block == 1024;
  if (MODE == 0 && gen) {
    if (fixed)
      PTB_LAUNCH(true, MODE == 0, FIXED);
    else if (tl.scene_smem)
      PTB_LAUNCH(true, MODE == 0, 0);
    else
      PTB_LAUNCH(false, MODE == 0, 0);
  } else {
    if (fixed)
      PTB_LAUNCH(true, false, FIXED);
    else if (tl.scene_smem)
      PTB_LAUNCH(true, false, 0);
    else
      PTB_LAUNCH(false, false, 0);
  }
#undef PTB_LAUNCH
}

// Paths per wavefront batch.  Bigger batches mean fewer, longer launches (less tail and launch overhead per ray:
// 32 Mi -> 256 Mi paths is -7 % trace time); the queues of a 256 Mi batch take 52 GB of the 180 GB, and the size
// is halved until they fit in 60 % of the memory that is free.
static size_t batch_capacity(size_t have /* capacity of the queues the device pool already holds */,
                             size_t vec4_bytes /* sizeof(Vec4<R>) of the pipeline that will run */) {
  size_t nb = (size_t)1 << 29;
  if (const char *e = std::getenv("PTB_BATCH")) {
    long long v = std::atoll(e);
    if (v >= 1024) nb = (size_t)v;
  }
  if (have >= nb) return nb;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    const size_t per_path = (size_t)(1 + NUM_MAT_KINDS) * 3 * vec4_bytes + 8;  // ray + hit queues
    free_b += have * per_path;  // growing frees the old queues first
    while (nb > have && nb > ((size_t)1 << 20) && nb * per_path > free_b / 10 * 6) nb >>= 1;
  }
  return std::max(nb, std::min(have, (size_t)1 << 29));
}

// The wavefront loop: raygen, then per bounce trace -> shade, batch after batch, all asynchronous on
// `st`.  Adds into d_sums (R[3*W*H]).
template <class R>
static int render_impl(ptb_scene *s, const ptb_params &p, R *d_sums, cudaStream_t st, ptb_stats *stats) {
  DeviceState *d = s->dev;
  int rc;
  if ((rc = ensure_tables<R>(s))) return rc;
  int world = p.tile_world > 0 ? p.tile_world : 1;
  if (p.tile_rank < 0 || p.tile_rank >= world) return fail(PTB_E_INVALID, "render: tile_rank out of range");
  DevicePool *pl = d->pool;
  if ((rc = ensure_pixel_list(pl, p.width, p.height, p.tile_rank, world))) return rc;
  RenderConst rcst;
  if ((rc = fill_render_const(p, pl->npix, &rcst))) return rc;
  // samples [base, base + total) of the pass-major enumeration of this rank's pixels
  const long long base = (long long)pl->npix * p.pass_first, total = (long long)pl->npix * passes_of(p);
  const size_t NB = std::min<size_t>(batch_capacity(pl->work<R>().cap, sizeof(Vec4<R>)), (size_t)std::max<long long>(total, 1));
  if ((rc = ensure_work<R>(pl, NB))) return rc;
  Work<R> &w = pl->work<R>();
  Ctl *ctl = pl->ctl;
  size_t scene_bytes = 0;
  DScene<R> sc = make_dscene<R>(s, &scene_bytes, true);
  TraceLaunch tl;
  if ((rc = trace_config<R, 0>(d, sc, scene_bytes, &tl))) return rc;
  const bool profile = (p.flags & PTB_FLAG_PROFILE) != 0;
  // every event of this call, destroyed on every way out (the CK early returns included)
  struct Events {
    std::vector<cudaEvent_t> all;
    ~Events() {
      for (cudaEvent_t e : all) cudaEventDestroy(e);
    }
    int make(cudaEvent_t *e) {
      CK(cudaEventCreate(e));
      all.push_back(*e);
      return PTB_OK;
    }
  } events;
  std::vector<cudaEvent_t> tev;
  cudaEvent_t e0, e1;
  if ((rc = events.make(&e0)) || (rc = events.make(&e1))) return rc;
  // PTB_FLAG_PROFILE: one pair of events per traversal launch, created before the timed region starts
  const size_t n_batches = (size_t)((total + (long long)NB - 1) / (long long)NB);
  if (profile) {
    tev.resize(2 * n_batches * (size_t)std::max(p.max_bounces, 0));
    for (cudaEvent_t &e : tev)
      if ((rc = events.make(&e))) return rc;
  }
  size_t tev_next = 0;
  pl->progress_total = (unsigned long long)total;
  *pl->progress_host = 0ull;
  // a scene with emitters collects their emission in k_shade, also after the last allowed intersection
  const bool emissive = sc.has_emissive != 0;
  CK(cudaMemsetAsync(ctl, 0, sizeof(Ctl), st));
  CK(cudaEventRecord(e0, st));
  uint64_t launches = 0;
  // queue entries move with bulk asynchronous copies (PTB_SHADE_BULK=0: the per-lane cp.async path, kept for A/B runs)
  static const bool shade_bulk = !(std::getenv("PTB_SHADE_BULK") && std::atoi(std::getenv("PTB_SHADE_BULK")) == 0);
  const size_t shade_smem = shade_smem_bytes<R>(256, shade_bulk);  // staging buffers of (A, B, C) + fill counts / mbarriers
  CK(cudaFuncSetAttribute(k_shade<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade_smem_bytes<R>(256, true)));
  CK(cudaFuncSetAttribute(k_shade<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade_smem_bytes<R>(256, false)));
  int shade_per_sm = 0;
  if (shade_bulk) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shade_per_sm, k_shade<R, true>, 256, shade_smem));
  else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shade_per_sm, k_shade<R, false>, 256, shade_smem));
  if (const char *e = std::getenv("PTB_SHADE_BLOCKS")) shade_per_sm = std::min(shade_per_sm, std::max(1, std::atoi(e)));
  const int shade_grid = d->sm_count * std::max(shade_per_sm, 1);
  // With per-lane cp.async input (PTB_SHADE_BULK=0) the bounce-0 launch of a big batch wants fewer blocks: its hits
  // are coherent, the table look-ups cost nothing, and the memory system delivers LESS when more warps stream 16-byte
  // requests through it (5.48 / 4.44 / 3.89 / 6.41 ms with 4 / 3 / 2 / 1 blocks per SM), while the later, incoherent
  // launches need every warp to hide their look-ups.  With bulk copies (3 requests of 512 B per item instead of 96
  // of 16 B) that congestion is gone and every launch runs best with all the blocks that fit.
  int shade_per_sm0 = shade_bulk ? shade_per_sm : std::min(shade_per_sm, 2);
  const int shade_grid0 = d->sm_count * std::max(shade_per_sm0, 1);
  for (long long first = 0; first < total; first += (long long)NB) {
    const unsigned n = (unsigned)std::min<long long>((long long)NB, total - first);
    k_batch_ctl<<<1, 128, 0, st>>>(ctl, n, p.max_bounces, pl->progress_dev, (unsigned long long)first);
    const int pass0 = (int)((base + first) / pl->npix), i0 = (int)((base + first) % pl->npix);
    const GenConst gen = make_gen(rcst, pl->pixel_list, pass0, i0);  // bounce 0 generates its own camera rays
    launches += 1;
    for (int b = 0; b < p.max_bounces; ++b) {
      const bool last = (b == p.max_bounces - 1);
      if (profile) {
        CK(cudaEventRecord(tev[tev_next], st));
      }
      launch_trace<R, 0>(tl, st, sc, b == 0 ? &gen : nullptr, n, w.rays, &ctl->nseg_rays[b], 0u, &ctl->cursor[b], w.mq,
                         (unsigned)w.slots, &ctl->nseg_mat[b][0], &ctl->n_rays[b], (last && !emissive) ? 0 : 1, d_sums, R(0), R(0), nullptr, nullptr);
      if (profile) {
        CK(cudaEventRecord(tev[tev_next + 1], st));
        tev_next += 2;
      }
      ++launches;
      if (!last || emissive) {
        // a path that is still alive after the last allowed bounce contributes black (integrator.ml:31-32), so the
        // last bounce needs no scatter — unless the scene has emitters, whose emission k_shade collects (last = 1)
        const int sg = (b == 0 && n >= (1u << 24)) ? shade_grid0 : shade_grid;
        if (shade_bulk)
          k_shade<R, true><<<sg, 256, shade_smem, st>>>(sc, rcst, b, w.mq, (unsigned)w.slots, &ctl->nseg_mat[b][0], w.rays,
                                                          &ctl->nseg_rays[b + 1], d_sums, last ? 1 : 0);
        else
          k_shade<R, false><<<sg, 256, shade_smem, st>>>(sc, rcst, b, w.mq, (unsigned)w.slots, &ctl->nseg_mat[b][0], w.rays,
                                                           &ctl->nseg_rays[b + 1], d_sums, last ? 1 : 0);
        ++launches;
      }
    }
  }
  k_batch_ctl<<<1, 128, 0, st>>>(ctl, 0u, p.max_bounces, pl->progress_dev, (unsigned long long)total);
  ++launches;
  CK(cudaEventRecord(e1, st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  if (stats) {
    Ctl host;
    CK(cudaMemcpy(&host, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    stats->paths = (uint64_t)total;
    stats->rays = host.total_rays;
    for (int b = 0; b < MAX_BOUNCES; ++b) stats->rays_by_bounce[b] = host.rays_by_bounce[b];
    stats->kernel_launches += launches;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    stats->ms_device = ms;
    double tr = 0;
    for (size_t i = 0; i + 1 < tev.size(); i += 2) {
      float m = 0;
      CK(cudaEventElapsedTime(&m, tev[i], tev[i + 1]));
      tr += m;
    }
    stats->ms_trace = tr;
    if (std::getenv("PTB_TIMING")) {  // per-launch device times: trace of bounce b, and the shade that follows it
      const size_t per_batch = 2 * (size_t)p.max_bounces;
      for (size_t i = 0; i + 1 < tev.size(); i += 2) {
        float t = 0, sh = 0;
        cudaEventElapsedTime(&t, tev[i], tev[i + 1]);
        if (i + 2 < tev.size() && (i + 2) % per_batch != 0) cudaEventElapsedTime(&sh, tev[i + 1], tev[i + 2]);
        std::fprintf(stderr, "[ptb timing] batch %zu bounce %zu: trace %.3f ms, shade %.3f ms (rays at this bounce, all batches: %llu)\n",
                     i / per_batch, (i % per_batch) / 2, t, sh, (unsigned long long)host.rays_by_bounce[(i % per_batch) / 2]);
      }
    }
  }
  return PTB_OK;
}

template <class R, class OUT>
static int resolve_impl(const R *d_sums, OUT *d_out, int W, int H, int spp, int flags, cudaStream_t st) {
  std::vector<double> w;
  filter_binomial(5, 1, &w);  // integrator.ml:134-135
  const int n = W * H;
  k_resolve<R, OUT><<<(n + 255) / 256, 256, 0, st>>>(d_sums, d_out, W, H, 1.0 / (double)spp, flags, w[0], w[1],
                                                      w[2], w[3], w[4], w[5], w[6], w[7], w[8]);
  CK(cudaGetLastError());
  return PTB_OK;
}

template <class R>
static int render_host_impl(ptb_scene *s, const ptb_params &p, double *image, ptb_stats *stats) {
  using clk = std::chrono::steady_clock;
  const size_t n3 = (size_t)p.width * p.height * 3;
  DevicePool *pl = s->dev->pool;
  auto grow = [](void **buf, size_t *cap, size_t bytes) -> int {
    if (*cap >= bytes) return PTB_OK;
    cudaFree(*buf);
    *buf = nullptr, *cap = 0;
    CK(cudaMalloc(buf, bytes));
    *cap = bytes;
    return PTB_OK;
  };
  int rc0;
  if ((rc0 = grow(&pl->sums_buf, &pl->sums_cap, n3 * sizeof(R)))) return rc0;
  if ((rc0 = grow(&pl->img_buf, &pl->img_cap, n3 * sizeof(double)))) return rc0;
  R *d_sums = (R *)pl->sums_buf;
  double *d_img = (double *)pl->img_buf;
  CK(cudaMemsetAsync(d_sums, 0, n3 * sizeof(R), 0));
  int rc = render_impl<R>(s, p, d_sums, 0, stats);
  if (!rc) rc = resolve_impl<R, double>(d_sums, d_img, p.width, p.height, passes_of(p), p.flags, 0);
  if (!rc) {
    auto t0 = clk::now();
    cudaError_t e = cudaMemcpy(image, d_img, n3 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(PTB_E_CUDA, std::string("image copy failed: ") + cudaGetErrorString(e));
    if (stats) {
      stats->ms_d2h = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
      stats->d2h_bytes = n3 * sizeof(double);
      stats->kernel_launches += 1;
    }
  }
  return rc;
}

// which of the per-material hit queues a hit on this kind of material goes to: emissive hits (extension) travel
// in the Lambertian queue, k_shade tells them apart by the material record it loads anyway
static uint8_t queue_kind(int kind) { return (uint8_t)(kind == PTB_MAT_EMISSIVE ? PTB_MAT_LAMBERTIAN : kind); }

static DMat make_dmat(const HostScene &h, size_t i) {
  DMat m;
  m.kind = h.mat[i].kind, m.tex = h.mat[i].texture, m.index = h.mat[i].index;
  m.tex_kind = PTB_TEX_SOLID, m.rgb[0] = m.rgb[1] = m.rgb[2] = 1.0f;
  if (m.tex >= 0 && (size_t)m.tex < h.tex.size()) {
    const ptb_texture &x = h.tex[m.tex];
    m.tex_kind = x.kind, m.rgb[0] = (float)x.rgb[0], m.rgb[1] = (float)x.rgb[1], m.rgb[2] = (float)x.rgb[2];
  }
  return m;
}

// per-device one-time setup: cached properties, the control block, the progress word
static int ensure_pool(DevicePool *pl, int device) {
  if (pl->sm_count == 0) {
    int sms = 0, optin = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    pl->sm_count = sms, pl->smem_optin = (size_t)optin;
  }
  if (!pl->ctl) {
    CK(cudaMalloc((void **)&pl->ctl, sizeof(Ctl)));
    CK(cudaMemset(pl->ctl, 0, sizeof(Ctl)));
  }
  if (!pl->progress_host) {
    CK(cudaHostAlloc((void **)&pl->progress_host, sizeof(unsigned long long), cudaHostAllocMapped));
    *pl->progress_host = 0;
    CK(cudaHostGetDevicePointer((void **)&pl->progress_dev, pl->progress_host, 0));
  }
  return PTB_OK;
}
// a DeviceState for `s` on `device`: the one it already has there (tables released, object kept) or a new one
static int new_device_state(ptb_scene *s, int device) {
  int rc = check_device(device);
  if (rc) return rc;
  DevicePool *pl = pool_for(device);
  if ((rc = ensure_pool(pl, device))) return rc;
  if (s->dev && s->dev->device == device) {
    release_device_tables(s->dev);
  } else {
    if (s->dev) destroy_device_state(s->dev);
    s->dev = new DeviceState();
  }
  DeviceState *d = s->dev;
  d->device = device;
  d->sm_count = pl->sm_count;
  d->smem_optin = pl->smem_optin;
  d->pool = pl;
  return PTB_OK;
}

// ---- device-side tree build for big triangle meshes (gpu_bvh.cuh) ------------------------------------------------
static bool want_gpu_builder(const HostScene &h) {
  if (h.n_spheres() != 0 || h.n_tris() < 2) return false;  // pure triangle meshes only
  if (const char *e = std::getenv("PTB_BUILDER")) {
    if (!std::strcmp(e, "host")) return false;
    if (!std::strcmp(e, "gpu") || !std::strcmp(e, "gpu-lbvh") || !std::strcmp(e, "gpu-sah")) return true;
  }
  return h.n_tris() >= 200000;  // below that the host SAH build is a few tens of ms and its tree is better
}

// Builds tree + float32 tables of `s` on s->dev->device.  Returns PTB_OK, or a positive value when the mesh
// cannot be handled here (tree too deep for the traversal stack) and the host builder should take over.
static bool gpu_builder_sah() {  // PTB_BUILDER=gpu-lbvh selects the linear builder, anything else the binned-SAH one
  const char *e = std::getenv("PTB_BUILDER");
  return !(e && !std::strcmp(e, "gpu-lbvh"));
}
static int gpu_build_mesh(ptb_scene *s) {
  using namespace gbvh;
  DeviceState *d = s->dev;
  const HostScene &h = s->host;
  const int nt = (int)h.n_tris();  // triangles; `n` below counts the builder's primitives = references to triangles
  int n = nt;
  const size_t nv = h.vx.size();
  double *vx = nullptr, *vy = nullptr, *vz = nullptr, *tuv = nullptr;
  int32_t *idx = nullptr, *tmat = nullptr;
  uint8_t *mkind = nullptr;
  float4 *blo = nullptr, *bhi = nullptr, *nlo = nullptr, *nhi = nullptr;
  Bounds6 *cb = nullptr;
  unsigned long long *k0 = nullptr, *k1 = nullptr;
  unsigned *v0 = nullptr, *v1 = nullptr, *visits = nullptr;
  int *lch = nullptr, *rch = nullptr, *first = nullptr, *count = nullptr, *pari = nullptr, *parl = nullptr, *counters = nullptr;
  Frontier *fa = nullptr, *fb = nullptr;
  Node4<float> *tmp_nodes = nullptr;
  void *cub_tmp = nullptr;
  std::vector<void *> to_free;
  // scratch comes from the device's stream-ordered pool, which keeps it cached between commits (cudaMalloc /
  // cudaFree of ~1 GB costs anything from 10 to 150 ms per build)
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, d->device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  auto A = [&](auto **p, size_t count_) -> int {
    CK(cudaMallocAsync((void **)p, std::max<size_t>(count_, 1) * sizeof(**p), 0));
    to_free.push_back((void *)*p);
    return PTB_OK;
  };
  auto cleanup = [&]() {
    for (void *p : to_free) cudaFreeAsync(p, 0);
    cudaStreamSynchronize(0);
  };
  int rc = PTB_OK;
#define G(x)            \
  if ((rc = (x))) {     \
    cleanup();          \
    return rc;          \
  }
  G(A(&vx, nv)) G(A(&vy, nv)) G(A(&vz, nv)) G(A(&idx, 3 * (size_t)nt)) G(A(&tmat, (size_t)nt)) G(A(&tuv, 6 * (size_t)nt))
  G(A(&mkind, h.mat.size())) G(A(&blo, (size_t)nt)) G(A(&bhi, (size_t)nt)) G(A(&cb, 2))
  auto H2D = [&](void *dst, const void *src, size_t bytes) -> int {
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return PTB_OK;
  };
  std::vector<uint8_t> mk(h.mat.size());
  for (size_t i = 0; i < mk.size(); ++i) mk[i] = queue_kind(h.mat[i].kind);
  G(H2D(vx, h.vx.data(), nv * 8)) G(H2D(vy, h.vy.data(), nv * 8)) G(H2D(vz, h.vz.data(), nv * 8))
  G(H2D(idx, h.tidx.data(), 3 * (size_t)nt * 4)) G(H2D(tmat, h.tmat.data(), (size_t)nt * 4)) G(H2D(tuv, h.tuv.data(), 6 * (size_t)nt * 8))
  G(H2D(mkind, mk.data(), mk.size()))
  Bounds6 cb0[2];
  for (int q = 0; q < 2; ++q)
    for (int k = 0; k < 3; ++k) cb0[q].lo[k] = 0xffffffffu, cb0[q].hi[k] = 0u;
  G(H2D(cb, cb0, sizeof cb0))
  k_tri_boxes<<<(unsigned)((nt + 255) / 256), 256>>>(vx, vy, vz, idx, nt, blo, bhi, cb);
  // Triangle pre-splitting (presplit.hpp): where the triangles' boxes overlap several times over — a soup, not a surface —
  // the builder gets several references per triangle, each with the box of one piece of it.
  int32_t *ref_tri = nullptr;  // reference -> triangle (null: the identity)
  {
    double *vsum = nullptr;
    G(A(&vsum, 1))
    if (cudaMemset(vsum, 0, 8) != cudaSuccess) {
      cleanup();
      return fail(PTB_E_CUDA, "gpu build: memset failed");
    }
    Bounds6 hb[2];
    double hv = 0.0;
    if (cudaMemcpy(hb, cb, sizeof hb, cudaMemcpyDeviceToHost) != cudaSuccess) {
      cleanup();
      return fail(PTB_E_CUDA, "gpu build: box bounds");
    }
    double org[3], vol = 1.0;
    for (int k = 0; k < 3; ++k) {
      org[k] = (double)dec_f(hb[1].lo[k]);
      vol *= std::max((double)dec_f(hb[1].hi[k]) - org[k], 1e-12);
    }
    k_box_volume_sum<<<(unsigned)((nt + 255) / 256), 256>>>(blo, bhi, nt, presplit::OVERLAP_BOX_CAP * vol / (double)nt, vsum);
    if (cudaMemcpy(&hv, vsum, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
      cleanup();
      return fail(PTB_E_CUDA, "gpu build: box volumes");
    }
    double f = presplit::budget_factor(hv / vol);
    if (const char *e = std::getenv("PTB_BVH_PRESPLIT")) f = std::atof(e);  // 0 / 1: off; > 1: references per triangle allowed
    if (f > 1.0 && (double)nt * f < (double)(1 << 26)) {
      int *cnt = nullptr, *off = nullptr;
      void *scan_tmp = nullptr;
      size_t scan_bytes = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt, off, nt);
      G(A(&cnt, (size_t)nt)) G(A(&off, (size_t)nt)) G(A((char **)&scan_tmp, scan_bytes))
      double cell = std::cbrt(vol / (double)nt), next_cell = cell;
      long long total = 0;
      for (int it = 0; it < 40; ++it, next_cell *= 1.1225) {
        cell = next_cell;  // (the cell size the counts and offsets below belong to: k_presplit_emit must use the same)
        k_presplit_count<<<(unsigned)((nt + 127) / 128), 128>>>(vx, vy, vz, idx, nt, cell, org[0], org[1], org[2], cnt);
        cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, cnt, off, nt);
        int last_off = 0, last_cnt = 0;
        if (cudaMemcpy(&last_off, off + nt - 1, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(&last_cnt, cnt + nt - 1, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
          cleanup();
          return fail(PTB_E_CUDA, "gpu build: pre-split count");
        }
        total = (long long)last_off + last_cnt;
        if ((double)total <= f * (double)nt) break;
      }
      if (total > nt && total < (1 << 26)) {
        float4 *rlo = nullptr, *rhi = nullptr;
        G(A(&rlo, (size_t)total)) G(A(&rhi, (size_t)total)) G(A(&ref_tri, (size_t)total))
        k_presplit_emit<<<(unsigned)((nt + 127) / 128), 128>>>(vx, vy, vz, idx, nt, cell, org[0], org[1], org[2], off, rlo, rhi, ref_tri);
        G(H2D(cb, cb0, sizeof cb0))
        k_box_bounds<<<(unsigned)((total + 255) / 256), 256>>>(rlo, rhi, (int)total, cb);
        blo = rlo, bhi = rhi, n = (int)total;
        if (std::getenv("PTB_BVH_TIMING"))
          std::fprintf(stderr, "[gpu bvh] presplit: overlap %.2f, %d -> %d references, cell %.4g\n", hv / vol, nt, n, cell);
      }
    }
  }
  const bool sah = gpu_builder_sah();
  const size_t nn = sah ? 2 * (size_t)n + 2 : (size_t)n;  // binary nodes
  G(A(&nlo, nn)) G(A(&nhi, nn))
  G(A(&k0, (size_t)n)) G(A(&k1, (size_t)n)) G(A(&v0, (size_t)n)) G(A(&v1, (size_t)n)) G(A(&visits, nn))
  G(A(&lch, nn)) G(A(&rch, nn)) G(A(&first, nn)) G(A(&count, nn)) G(A(&pari, nn)) G(A(&parl, nn))
  G(A(&counters, 4)) G(A(&fa, (size_t)n)) G(A(&fb, (size_t)n)) G(A(&tmp_nodes, (size_t)n))
  (void)visits;
  const unsigned gb = (unsigned)((n + 255) / 256);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k0, k1, v0, v1, n, 0, 64);
  G(A((char **)&cub_tmp, cub_bytes))
  if (cudaMemset(visits, 0, nn * 4) != cudaSuccess || cudaMemset(counters, 0, 16) != cudaSuccess) {
    cleanup();
    return fail(PTB_E_CUDA, "gpu build: memset failed");
  }
  int leaf_thresh = GLEAF;
  if (!sah) {
    k_morton<<<gb, 256>>>(blo, bhi, n, cb, k0, v0);
    cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k0, k1, v0, v1, n, 0, 63);
    k_radix_tree<<<gb, 256>>>(k1, n, lch, rch, first, count, pari, parl);
    k_fit<<<gb, 256>>>(v1, blo, bhi, n, lch, rch, pari, parl, visits, nlo, nhi);
  } else {
    // binned SAH, level by level (gpu_bvh.cuh).  pari = slot, parl = axis, rch.. map onto the SahTree fields
    int *node_id = nullptr, *act_a = nullptr, *act_b = nullptr, *split = nullptr, *level_of = nullptr;
    unsigned *bins = nullptr;
    const size_t max_active = (size_t)n / (GLEAF + 1) + 4;
    G(A(&node_id, (size_t)n)) G(A(&act_a, max_active)) G(A(&act_b, max_active)) G(A(&split, nn)) G(A(&level_of, nn))
    G(A(&bins, max_active * NODE_BINS))
    SahTree T{nlo, nhi, count, first, lch, rch, pari, parl, split, level_of, visits};
    Bounds6 hb[2];
    if (cudaMemcpy(hb, cb, sizeof hb, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemset(node_id, 0, (size_t)n * 4) != cudaSuccess ||
        cudaMemset(k0, 0, (size_t)n * 8) != cudaSuccess) {
      cleanup();
      return fail(PTB_E_CUDA, "gpu build: setup failed");
    }
    const float4 rlo = make_float4(dec_f(hb[1].lo[0]), dec_f(hb[1].lo[1]), dec_f(hb[1].lo[2]), 0.f);
    const float4 rhi = make_float4(dec_f(hb[1].hi[0]), dec_f(hb[1].hi[1]), dec_f(hb[1].hi[2]), 0.f);
    const int i_n = n, i_zero = 0, i_m1 = -1, i_max = INT32_MAX;
    G(H2D(nlo, &rlo, 16)) G(H2D(nhi, &rhi, 16)) G(H2D(count, &i_n, 4)) G(H2D(first, &i_max, 4)) G(H2D(lch, &i_m1, 4))
    G(H2D(rch, &i_m1, 4)) G(H2D(pari, &i_zero, 4)) G(H2D(level_of, &i_m1, 4)) G(H2D(act_a, &i_zero, 4))
    int hcs[4] = {1, 0, 0, 0};
    G(H2D(counters, hcs, sizeof hcs))
    int n_active = 1;
    for (int level = 0; n_active > 0; ++level) {
      if (level >= 63) {  // path key exhausted: not a mesh for this builder
        cleanup();
        return 1;
      }
      k_sah_clear<<<(unsigned)(((size_t)n_active * 3 * SB + 255) / 256), 256>>>(bins, n_active);
      if (n_active <= 4) k_sah_bin<true><<<gb, 256>>>(blo, bhi, n, node_id, T, bins);
      else k_sah_bin<false><<<gb, 256>>>(blo, bhi, n, node_id, T, bins);
      k_sah_select<<<(unsigned)(((size_t)n_active * 32 + 255) / 256), 256>>>(act_a, n_active, bins, T, level, act_b, counters);
      k_sah_assign<<<gb, 256>>>(blo, bhi, n, node_id, k0, T, level);
      if (cudaMemcpy(hcs, counters, sizeof hcs, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        cleanup();
        return fail(PTB_E_CUDA, std::string("gpu build (sah): ") + cudaGetErrorString(e));
      }
      n_active = hcs[1];
      G(H2D(counters + 1, &i_zero, 4))
      std::swap(act_a, act_b);
    }
    k_iota<<<gb, 256>>>(v0, n);
    cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k0, k1, v0, v1, n, 0, 64);
    k_sah_first<<<gb, 256>>>(v1, n, node_id, first);
    leaf_thresh = 0;
    if (cudaMemset(counters, 0, 16) != cudaSuccess) {
      cleanup();
      return fail(PTB_E_CUDA, "gpu build: memset failed");
    }
  }
  // collapse, level by level from the root (binary node 0 = wide node 0)
  Frontier root{0, 0};
  G(H2D(fa, &root, sizeof root))
  int hc[4] = {1, 0, 0, 0};
  G(H2D(counters, hc, sizeof hc))
  int n_cur = 1, levels = 0;
  while (n_cur > 0) {
    ++levels;
    k_collapse_level<<<(unsigned)((n_cur + 127) / 128), 128>>>(fa, n_cur, lch, rch, first, count, v1, blo, bhi, nlo, nhi, tmp_nodes, fb,
                                                             counters, leaf_thresh);
    if (cudaMemcpy(hc, counters, sizeof hc, cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaError_t e = cudaGetLastError();
      cleanup();
      return fail(PTB_E_CUDA, std::string("gpu build: ") + cudaGetErrorString(e));
    }
    n_cur = hc[1];
    const int zero = 0;
    G(H2D(counters + 1, &zero, 4))
    std::swap(fa, fb);
    if (3 * levels + 1 > 97) {  // deeper than the traversal stack can hold: let the host builder do this mesh
      cleanup();
      return 1;
    }
  }
  const int n_wide = hc[0];
  // outputs
  Tables<float> &t = d->tf;
  release_device_tables(d);
  auto OUT = [&](auto **p, size_t count_) -> int {
    CK(tbl_alloc((void **)p, std::max<size_t>(count_, 1) * sizeof(**p)));
    return PTB_OK;
  };
  G(OUT(&t.nodes, (size_t)n_wide)) G(OUT(&t.spheres, 1)) G(OUT(&t.tris, 3 * (size_t)n)) G(OUT(&t.tri_uv, 6 * (size_t)n))
  G(OUT(&d->sphere_id, 1)) G(OUT(&d->sphere_mat, 1)) G(OUT(&d->tri_id, (size_t)n)) G(OUT(&d->tri_mat, (size_t)n))
  G(OUT(&d->prim_kind, ((size_t)n + 15) / 16 * 16 + 16))
  if (cudaMemcpy(t.nodes, tmp_nodes, (size_t)n_wide * sizeof(Node4<float>), cudaMemcpyDeviceToDevice) != cudaSuccess ||
      cudaMemset(d->prim_kind, 0, ((size_t)n + 15) / 16 * 16 + 16) != cudaSuccess) {
    cleanup();
    return fail(PTB_E_CUDA, "gpu build: copy failed");
  }
  k_emit_tris<<<gb, 256>>>(v1, n, ref_tri, vx, vy, vz, idx, tmat, tuv, mkind, t.tris, t.tri_uv, d->tri_id, d->tri_mat, d->prim_kind);
  std::vector<DTex<float>> texs(h.tex.size());
  for (size_t i = 0; i < texs.size(); ++i) {
    const ptb_texture &x = h.tex[i];
    texs[i] = {x.kind, x.width, x.height, x.even, x.odd, 0, {(float)x.rgb[0], (float)x.rgb[1], (float)x.rgb[2]}};
  }
  std::vector<DMat> mats(h.mat.size());
  for (size_t i = 0; i < mats.size(); ++i) mats[i] = make_dmat(h, i);
  G(upload(&t.texs, texs)) G(upload(&d->mats, mats))
  cudaError_t e = cudaDeviceSynchronize();
  cleanup();
  if (e != cudaSuccess) return fail(PTB_E_CUDA, std::string("gpu build: ") + cudaGetErrorString(e));
#undef G
  t.ready = true;
  WideBVH &b = s->bvh;
  b.nodes.clear(), b.sphere_order.clear(), b.tri_order.clear();
  b.device_built = true, b.dev_nodes = n_wide, b.dev_tris = n, b.dev_leaves = hc[2];
  b.presplit = ref_tri != nullptr;
  b.depth = levels, b.max_stack = 3 * levels;
  return PTB_OK;
}

static int require_committed(ptb_scene *s, int device) {
  if (!s) return fail(PTB_E_INVALID, "null scene");
  if (!s->committed || !s->dev) return fail(PTB_E_STATE, "scene is not committed (call ptb_scene_commit)");
  if (s->dev->device != device) return fail(PTB_E_STATE, "scene was committed on a different device");
  return check_device(device);
}

}  // namespace ptb

using namespace ptb;

extern "C" {

static int upload_scene(ptb_scene *s, int32_t device);

int ptb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ptb_scene_commit(ptb_scene *s, int32_t device, double *ms) {
  using clk = std::chrono::steady_clock;
  if (!s) return fail(PTB_E_INVALID, "commit: null scene");
  s->committed = false;  // until every table is on the device (a failed re-commit must not leave a usable-looking scene)
  const HostScene &h = s->host;
  // Shape_tree.create: `failwith "expected non-empty list of shapes"` (shape_tree.ml:254-255)
  if (h.n_spheres() + h.n_tris() == 0) return fail(PTB_E_INVALID, "commit: expected non-empty list of shapes");
  if (h.n_spheres() >= (1 << 26) || h.n_tris() >= (1 << 26)) return fail(PTB_E_INVALID, "commit: too many primitives");
  auto bad_mat = [&](const std::vector<int32_t> &m) {
    for (int32_t v : m)
      if (v < 0 || v >= (int32_t)h.mat.size()) return true;
    return false;
  };
  if (bad_mat(h.smat) || bad_mat(h.tmat)) return fail(PTB_E_INVALID, "commit: material row out of range");
  for (const ptb_material &m : h.mat)
    if (m.kind != PTB_MAT_DIELECTRIC && (m.texture < 0 || m.texture >= (int32_t)h.tex.size()))
      return fail(PTB_E_INVALID, "commit: texture row out of range");
  int rc = check_device(device);
  if (rc) return rc;
  PoolLock lk(pool_for(device)->mu);
  auto t0 = clk::now();
  for (ptb_scene *r : s->replicas) ptb_scene_destroy(r);  // replicas of an older commit
  s->replicas.clear();
  s->bvh = WideBVH();
  auto drop_state = [&]() {  // a commit that failed half way leaves no device state behind
    if (s->dev) destroy_device_state(s->dev);
    s->dev = nullptr;
  };
  if (want_gpu_builder(h)) {  // big pure-triangle mesh: build the tree on the device
    if ((rc = new_device_state(s, device))) return rc;
    rc = gpu_build_mesh(s);
    if (rc < 0) {
      drop_state();
      return rc;
    }
    if (rc == PTB_OK) {
      s->committed = true;
      if (ms) *ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
      return PTB_OK;
    }
  }
  build_wide_bvh(h, &s->bvh);
  // the traversal stack holds the tree's exact worst case + the sentinel; refuse trees it cannot hold rather
  // than dropping pushes on the device
  if (s->bvh.max_stack + 1 > 97) return fail(PTB_E_INVALID, "commit: tree too deep for the device traversal stack");
  if ((rc = upload_scene(s, device))) {
    drop_state();
    return rc;
  }
  if (ms) *ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  return PTB_OK;
}

/* The device half of a commit: ids, materials, tree and float tables of `s` (already built) onto `device` as ONE
 * block — written into pinned staging, copied with one cudaMemcpyAsync into one allocation.  (A dozen blocking
 * cudaMemcpy's of pageable vectors cost 3.5 - 230 ms per commit of a 54 KB scene next to 50 GB of queues.) */
static int upload_scene(ptb_scene *s, int32_t device) {
  const HostScene &h = s->host;
  s->committed = false;
  int rc = new_device_state(s, device);
  if (rc) return rc;
  DeviceState *d = s->dev;
  DevicePool *pl = d->pool;
  PoolLock lk(pl->mu);
  const size_t nS = s->bvh.sphere_order.size(), nT = s->bvh.tri_order.size();
  std::vector<int32_t> smat(nS), tmat(nT);
  for (size_t k = 0; k < nS; ++k) smat[k] = h.smat[s->bvh.sphere_order[k]];
  for (size_t k = 0; k < nT; ++k) tmat[k] = h.tmat[s->bvh.tri_order[k]];
  std::vector<DMat> mats(h.mat.size());
  for (size_t i = 0; i < mats.size(); ++i) mats[i] = make_dmat(h, i);
  std::vector<uint8_t> kinds(((nS + nT + 15) / 16) * 16 + 16, 0);
  for (size_t k = 0; k < nS; ++k) kinds[k] = queue_kind(h.mat[smat[k]].kind);
  for (size_t k = 0; k < nT; ++k) kinds[nS + k] = queue_kind(h.mat[tmat[k]].kind);
  HostTables<float> ht;
  make_host_tables<float>(s, &ht);
  // layout of the block: 256-byte aligned slices
  struct Slice {
    const void *src;
    size_t bytes, off;
  };
  Slice sl[11] = {{s->bvh.sphere_order.data(), nS * 4, 0}, {s->bvh.tri_order.data(), nT * 4, 0},
                  {smat.data(), nS * 4, 0},               {tmat.data(), nT * 4, 0},
                  {mats.data(), mats.size() * sizeof(DMat), 0}, {kinds.data(), kinds.size(), 0},
                  {ht.nodes.data(), ht.nodes.size() * sizeof(Node4<float>), 0},
                  {ht.sph.data(), ht.sph.size() * sizeof(Vec4<float>), 0},
                  {ht.tri.data(), ht.tri.size() * sizeof(Vec4<float>), 0},
                  {ht.uv.data(), ht.uv.size() * sizeof(float), 0},
                  {ht.texs.data(), ht.texs.size() * sizeof(DTex<float>), 0}};
  size_t total = 0;
  for (Slice &x : sl) {
    x.off = total;
    total += (std::max<size_t>(x.bytes, 16) + 255) / 256 * 256;
  }
  if (pl->stage_cap < total) {
    if (pl->stage_host) cudaFreeHost(pl->stage_host);
    pl->stage_host = nullptr, pl->stage_cap = 0;
    CK(cudaHostAlloc(&pl->stage_host, total + total / 2, cudaHostAllocDefault));
    pl->stage_cap = total + total / 2;
  }
  char *hb = (char *)pl->stage_host;
  for (const Slice &x : sl)
    if (x.bytes) std::memcpy(hb + x.off, x.src, x.bytes);
  CK(tbl_alloc(&d->blob, total));
  CK(cudaMemcpyAsync(d->blob, hb, total, cudaMemcpyHostToDevice, 0));
  char *db = (char *)d->blob;
  d->sphere_id = (int32_t *)(db + sl[0].off), d->tri_id = (int32_t *)(db + sl[1].off);
  d->sphere_mat = (int32_t *)(db + sl[2].off), d->tri_mat = (int32_t *)(db + sl[3].off);
  d->mats = (DMat *)(db + sl[4].off), d->prim_kind = (uint8_t *)(db + sl[5].off);
  d->ids_in_blob = true;
  Tables<float> &t = d->tf;
  t.nodes = (Node4<float> *)(db + sl[6].off), t.spheres = (Vec4<float> *)(db + sl[7].off);
  t.tris = (Vec4<float> *)(db + sl[8].off), t.tri_uv = (float *)(db + sl[9].off), t.texs = (DTex<float> *)(db + sl[10].off);
  t.in_blob = true, t.ready = true;
  CK(cudaStreamSynchronize(0));  // the staging block may be reused by the next commit
  s->committed = true;
  return PTB_OK;
}

/* Single-process multi-GPU: the scene is replicated on devices 0..n-1 (tree built once). */
int ptb_scene_commit_multi(ptb_scene *s, int32_t n_devices, double *ms) {
  using clk = std::chrono::steady_clock;
  if (!s || n_devices < 1) return fail(PTB_E_INVALID, "commit_multi: bad args");
  if (n_devices > ptb_device_count()) return fail(PTB_E_NO_DEVICE, "commit_multi: fewer CUDA devices than requested");
  auto t0 = clk::now();
  int rc = ptb_scene_commit(s, 0, nullptr);
  if (rc) return rc;
  for (int32_t i = 1; i < n_devices; ++i) {
    ptb_scene *r = new ptb_scene();
    r->host = s->host, r->bvh = s->bvh, r->ref_order = s->ref_order;
    s->replicas.push_back(r);
    if (s->bvh.device_built) {  // the device builder is deterministic: every replica builds the same tree
      r->bvh = WideBVH();
      if ((rc = new_device_state(r, i))) return rc;
      rc = gpu_build_mesh(r);
      if (rc > 0) rc = fail(PTB_E_STATE, "commit_multi: device build refused on a replica");
      if (rc) return rc;
      r->committed = true;
    } else if ((rc = upload_scene(r, i))) {
      return rc;
    }
  }
  check_device(0);
  if (ms) *ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  return PTB_OK;
}

/* Integrator.render over n_devices GPUs of this process: device i renders the tiles t = i (mod n) of the
 * reference's tile list into its own per-pixel sums (one host thread per device), device 0 then adds the other
 * devices' sums straight out of their memory over NVLink (peer loads) and runs the filter + gamma resolve. */
int ptb_render_multi(ptb_scene *s, const ptb_params *p, int32_t n_devices, double *image, ptb_stats *stats) {
  using clk = std::chrono::steady_clock;
  if (!s || !p || !image || n_devices < 1) return fail(PTB_E_INVALID, "render_multi: bad args");
  if (p->flags & PTB_FLAG_F64) return fail(PTB_E_INVALID, "render_multi: float32 pipeline only");
  if (!s->committed || !s->dev || s->dev->device != 0 || (int32_t)s->replicas.size() + 1 < n_devices)
    return fail(PTB_E_STATE, "render_multi: scene is not committed on that many devices (call ptb_scene_commit_multi)");
  if (n_devices > 8) return fail(PTB_E_INVALID, "render_multi: at most 8 devices");
  auto t0 = clk::now();
  const size_t n3 = (size_t)p->width * p->height * 3;
  std::vector<int> rcs((size_t)n_devices, 0);
  std::vector<std::string> msgs((size_t)n_devices);
  std::vector<ptb_stats> st((size_t)n_devices);
  std::vector<float *> sums((size_t)n_devices, nullptr);
  auto one = [&](int i) {
    ptb_scene *sc = i == 0 ? s : s->replicas[(size_t)i - 1];
    auto body = [&]() -> int {
      int rc = check_device(i);
      if (rc) return rc;
      DevicePool *pl = sc->dev->pool;
      PoolLock lk(pl->mu);
      if (pl->sums_cap < n3 * sizeof(float)) {
        cudaFree(pl->sums_buf);
        pl->sums_buf = nullptr, pl->sums_cap = 0;
        CK(cudaMalloc(&pl->sums_buf, n3 * sizeof(float)));
        pl->sums_cap = n3 * sizeof(float);
      }
      sums[(size_t)i] = (float *)pl->sums_buf;
      CK(cudaMemsetAsync(sums[(size_t)i], 0, n3 * sizeof(float), 0));
      ptb_params q = *p;
      q.device = i, q.tile_rank = i, q.tile_world = n_devices;
      std::memset(&st[(size_t)i], 0, sizeof(ptb_stats));
      return render_impl<float>(sc, q, sums[(size_t)i], 0, &st[(size_t)i]);  // synchronises its stream
    };
    rcs[(size_t)i] = body();
    if (rcs[(size_t)i]) msgs[(size_t)i] = ptb_last_error();
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n_devices; ++i) th.emplace_back(one, i);
  one(0);
  for (auto &t : th) t.join();
  for (int i = 0; i < n_devices; ++i)
    if (rcs[(size_t)i]) return fail(rcs[(size_t)i], "render_multi: device " + std::to_string(i) + ": " + msgs[(size_t)i]);
  // ---- framebuffer reduce on device 0 over peer memory, then resolve -------------------------------------
  int rc = check_device(0);
  if (rc) return rc;
  DevicePool *pl0 = s->dev->pool;
  PoolLock lk0(pl0->mu);
  PeerPtrs pp;
  std::memset(&pp, 0, sizeof pp);
  float *stage = nullptr;
  for (int i = 1; i < n_devices; ++i) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, i));
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(i, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(PTB_E_CUDA, std::string("peer access: ") + cudaGetErrorString(e));
      cudaGetLastError();
      pp.src[pp.n++] = sums[(size_t)i];
    } else {  // no NVLink/PCIe peer path: stage the shard through a copy and add it
      if (!stage) CK(cudaMalloc((void **)&stage, n3 * sizeof(float)));
      CK(cudaMemcpyPeer(stage, 0, sums[(size_t)i], i, n3 * sizeof(float)));
      PeerPtrs one_src;
      std::memset(&one_src, 0, sizeof one_src);
      one_src.src[0] = stage, one_src.n = 1;
      k_reduce_peers<<<(unsigned)(((n3 + 3) / 4 + 255) / 256), 256>>>(sums[0], one_src, n3);
      CK(cudaDeviceSynchronize());
    }
  }
  if (pp.n) k_reduce_peers<<<(unsigned)(((n3 + 3) / 4 + 255) / 256), 256>>>(sums[0], pp, n3);
  CK(cudaGetLastError());
  if (stage) cudaFree(stage);
  if (pl0->img_cap < n3 * sizeof(double)) {
    cudaFree(pl0->img_buf);
    pl0->img_buf = nullptr, pl0->img_cap = 0;
    CK(cudaMalloc(&pl0->img_buf, n3 * sizeof(double)));
    pl0->img_cap = n3 * sizeof(double);
  }
  rc = resolve_impl<float, double>(sums[0], (double *)pl0->img_buf, p->width, p->height, passes_of(*p), p->flags, 0);
  if (rc) return rc;
  CK(cudaMemcpy(image, pl0->img_buf, n3 * sizeof(double), cudaMemcpyDeviceToHost));
  if (stats) {
    std::memset(stats, 0, sizeof *stats);
    for (int i = 0; i < n_devices; ++i) {
      const ptb_stats &a = st[(size_t)i];
      stats->paths += a.paths, stats->rays += a.rays, stats->kernel_launches += a.kernel_launches;
      for (int b = 0; b < MAX_BOUNCES; ++b) stats->rays_by_bounce[b] += a.rays_by_bounce[b];
      stats->ms_device = std::max(stats->ms_device, a.ms_device);
      stats->ms_trace = std::max(stats->ms_trace, a.ms_trace);
    }
    stats->kernel_launches += 2;
    stats->d2h_bytes = n3 * sizeof(double);
    stats->ms_total = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  }
  return PTB_OK;
}

int ptb_scene_tree_stats(const ptb_scene *s, int32_t out[8]) {
  if (!s || !out) return fail(PTB_E_INVALID, "tree_stats: null argument");
  if (!s->committed) return fail(PTB_E_STATE, "scene is not committed (call ptb_scene_commit)");
  std::memset(out, 0, 8 * sizeof(int32_t));
  out[0] = (int32_t)s->bvh.node_count();
  out[1] = s->bvh.depth;
  out[2] = s->bvh.max_stack;
  out[3] = (int32_t)s->bvh.sphere_order.size();
  out[4] = (int32_t)s->bvh.tri_count();
  int leaves = (int)s->bvh.dev_leaves;
  for (const WideNode &n : s->bvh.nodes)
    for (int k = 0; k < 4; ++k)
      if (n.child[k] < 0 && n.child[k] != EMPTY_CHILD) ++leaves;
  out[5] = leaves;
  out[6] = LEAF_MAX;
  return PTB_OK;
}

int ptb_render(ptb_scene *s, const ptb_params *p, double *image, ptb_stats *stats) {
  using clk = std::chrono::steady_clock;
  if (!p || !image) return fail(PTB_E_INVALID, "render: null argument");
  int rc = require_committed(s, p->device);
  if (rc) return rc;
  PoolLock lk(s->dev->pool->mu);
  if (stats) std::memset(stats, 0, sizeof *stats);
  auto t0 = clk::now();
  rc = (p->flags & PTB_FLAG_F64) ? render_host_impl<double>(s, *p, image, stats)
                                 : render_host_impl<float>(s, *p, image, stats);
  if (stats) stats->ms_total = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  return rc;
}

int ptb_render_progress(int32_t device, uint64_t *paths_done, uint64_t *paths_total) {
  if (device < 0 || device >= 64) return fail(PTB_E_INVALID, "render_progress: device ordinal out of range");
  const DevicePool *pl = pool_for(device);  // no lock: this is what another thread polls while a render holds it
  const volatile unsigned long long *w = pl->progress_host;
  if (paths_done) *paths_done = w ? (uint64_t)*w : 0u;
  if (paths_total) *paths_total = (uint64_t)pl->progress_total;
  return PTB_OK;
}

void *ptb_host_alloc(uint64_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, (size_t)std::max<uint64_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    fail(PTB_E_CUDA, std::string("host_alloc: ") + cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
void ptb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int ptb_render_device(ptb_scene *s, const ptb_params *p, float *d_sums, void *stream, ptb_stats *stats) {
  using clk = std::chrono::steady_clock;
  if (!p || !d_sums) return fail(PTB_E_INVALID, "render_device: null argument");
  if (p->flags & PTB_FLAG_F64) return fail(PTB_E_INVALID, "render_device: float32 sums only");
  int rc = require_committed(s, p->device);
  if (rc) return rc;
  PoolLock lk(s->dev->pool->mu);
  if (stats) std::memset(stats, 0, sizeof *stats);
  auto t0 = clk::now();
  rc = render_impl<float>(s, *p, d_sums, (cudaStream_t)stream, stats);
  if (stats) stats->ms_total = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  return rc;
}

int ptb_resolve_device(const float *d_sums, float *d_image, int32_t width, int32_t height, int32_t spp,
                       int32_t flags, int32_t device, void *stream) {
  if (!d_sums || !d_image || width <= 0 || height <= 0 || spp <= 0) return fail(PTB_E_INVALID, "resolve: bad args");
  int rc = check_device(device);
  if (rc) return rc;
  return resolve_impl<float, float>(d_sums, d_image, width, height, spp, flags, (cudaStream_t)stream);
}

// Caller-supplied rays -> the ray queue.  Scenes that are traversed in global memory get their rays in spatial order
// (ray_sort.cuh); a scene staged in shared memory gains nothing from that (no memory system to be kind to) and keeps
// the caller's order.  PTB_BATCH_SORT=0/1 overrides the choice (A/B measurements).
static int pack_user_rays(DevicePool *pl, const ptb_scene *s, const DScene<float> &sc, const float *d_o, const float *d_d, long long m,
                          float t_min, Queue<float> q, cudaStream_t st, uint64_t *launches) {
  const char *e = std::getenv("PTB_BATCH_SORT");
  const int env = e ? (std::atoi(e) != 0 ? 1 : 0) : 2;
  // The sort costs about as much as tracing the batch against a few thousand primitives, so it is kept for the scenes
  // whose traversal is expensive and far-flung in memory: soups the builder pre-split, beyond 48 MB of records (measured:
  // +10-17 % there, +30 % on 10^7 triangles; -7 to -21 % on 10^4-10^5 triangles, sphere sets and meshes most rays miss).
  const bool heavy = s->bvh.presplit && (size_t)sc.n_tris * sizeof(TriG) + (size_t)sc.n_nodes * sizeof(NodeQ) > ((size_t)48 << 20);
  const bool sorted = env == 2 ? (sc.scene_in_smem == 0 && heavy && m >= (1 << 15)) : env == 1;
  const unsigned blocks = (unsigned)((m + 255) / 256);
  if (!sorted || m <= 0) {
    if (m > 0) k_pack_rays<float><<<blocks, 256, 0, st>>>(d_o, d_d, m, q);
    *launches += 1;
    return PTB_OK;
  }
  const size_t need = ray_sort_scratch_bytes((size_t)m);
  if (pl->sort_cap < need) {
    cudaFree(pl->sort_buf);
    pl->sort_buf = nullptr, pl->sort_cap = 0;
    CK(cudaMalloc(&pl->sort_buf, need));
    pl->sort_cap = need;
  }
  unsigned *bins = (unsigned *)pl->sort_buf, *totals = bins + SORT_BINS, *keys = totals + SORT_SCAN_BLOCKS;
  int oct_bits = 0;  // the direction octant in the key: measured slower than the cell alone (more bins, shorter runs)
  if (const char *o = std::getenv("PTB_SORT_OCT")) oct_bits = std::atoi(o) ? 3 : 0;
  const unsigned nbins = SORT_BINS >> (3 - oct_bits), scan_blocks = nbins / SORT_SCAN_TILE;
  CK(cudaMemsetAsync(bins, 0, (size_t)nbins * 4, st));
  k_ray_keys<<<blocks, 256, 0, st>>>(d_o, d_d, m, sc.nodes, t_min, oct_bits, keys, bins);
  k_sort_scan_tiles<<<scan_blocks, SORT_SCAN_BLOCK, 0, st>>>(bins, totals);
  k_sort_scan_add<<<scan_blocks, SORT_SCAN_BLOCK, 0, st>>>(bins, totals);
  unsigned *perm = keys + m;
  k_sort_scatter<<<blocks, 256, 0, st>>>(m, keys, bins, perm);
  k_pack_rays_sorted<float><<<blocks, 256, 0, st>>>(d_o, d_d, m, perm, q);
  *launches += 5;
  return PTB_OK;
}

int ptb_intersect_batch_device(ptb_scene *s, const float *d_o, const float *d_d, float t_min, float t_max,
                               int64_t n, float *d_t, int32_t *d_prim, int32_t device, void *stream,
                               ptb_stats *stats) {
  if (!d_o || !d_d || !d_t || !d_prim || n < 0) return fail(PTB_E_INVALID, "intersect_batch: bad args");
  int rc = require_committed(s, device);
  if (rc) return rc;
  DeviceState *d = s->dev;
  DevicePool *pl = d->pool;
  PoolLock lk(pl->mu);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cap = std::min<size_t>(batch_capacity(pl->work<float>().cap, sizeof(Vec4<float>)), (size_t)std::max<int64_t>(n, 1));
  if ((rc = ensure_work<float>(pl, cap))) return rc;
  Work<float> &w = pl->work<float>();
  size_t scene_bytes = 0;
  DScene<float> sc = make_dscene<float>(s, &scene_bytes);
  TraceLaunch tl;
  if ((rc = trace_config<float, 1>(d, sc, scene_bytes, &tl))) return rc;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  uint64_t launches = 0;
  for (int64_t first = 0; first < n; first += (int64_t)cap) {
    const long long m = std::min<int64_t>((int64_t)cap, n - first);
    if ((rc = pack_user_rays(pl, s, sc, d_o + 3 * first, d_d + 3 * first, m, t_min, w.rays, st, &launches))) return rc;
    CK(cudaMemsetAsync(&pl->ctl->cursor[MAX_BOUNCES], 0, sizeof(unsigned), st));
    launch_trace<float, 1>(tl, st, sc, nullptr, 0u, w.rays, nullptr, (unsigned)((m + SEG - 1) / SEG),
                           &pl->ctl->cursor[MAX_BOUNCES], w.mq, (unsigned)w.slots, nullptr, nullptr, 0, nullptr, t_min, t_max, d_t + first,
                           d_prim + first);
    launches += 1;
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  if (stats) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    stats->ms_device = ms;
    stats->rays = (uint64_t)n;
    stats->kernel_launches += launches;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return PTB_OK;
}

/* Host buffers.  The rays are cut into chunks; chunk k+1 is copied to the device and chunk k-1's results come back
 * while chunk k is traced (two streams, two sets of device staging buffers).  That only overlaps when the copies are
 * real DMA transfers, i.e. from page-locked memory: buffers from ptb_host_alloc (or any registered memory) go
 * straight through; pageable buffers of a big call are pinned for its duration (cudaHostRegister), small calls
 * take the driver's staged copy. */
int ptb_intersect_batch(ptb_scene *s, const float *o, const float *dd, float t_min, float t_max, int64_t n,
                        float *t_hit, int32_t *prim, int32_t device, ptb_stats *stats) {
  using clk = std::chrono::steady_clock;
  if (!o || !dd || !t_hit || !prim || n < 0) return fail(PTB_E_INVALID, "intersect_batch: bad args");
  int rc = require_committed(s, device);
  if (rc) return rc;
  DeviceState *d = s->dev;
  DevicePool *pl = d->pool;
  PoolLock lk(pl->mu);
  if (stats) std::memset(stats, 0, sizeof *stats);
  auto t0 = clk::now();
  size_t chunk = (size_t)1 << 22;  // rays per chunk: 96 MB in, 32 MB out
  if (const char *e = std::getenv("PTB_BATCH_CHUNK")) chunk = (size_t)std::max<long long>(1024, std::atoll(e));
  chunk = std::min<size_t>(chunk, (size_t)std::max<int64_t>(n, 1));
  const int nbuf = (size_t)n > chunk ? 2 : 1;
  // device staging: [buffer][origins, directions, t, prim], kept between calls (grow-only)
  const size_t want[4] = {nbuf * chunk * 12, nbuf * chunk * 12, nbuf * chunk * 4, nbuf * chunk * 4};
  for (int k = 0; k < 4; ++k)
    if (pl->batch_cap[k] < want[k]) {
      cudaFree(pl->batch_buf[k]);
      pl->batch_buf[k] = nullptr, pl->batch_cap[k] = 0;
      CK(cudaMalloc(&pl->batch_buf[k], want[k]));
      pl->batch_cap[k] = want[k];
    }
  if ((rc = ensure_work<float>(pl, chunk))) return rc;
  Work<float> &w = pl->work<float>();
  size_t scene_bytes = 0;
  DScene<float> sc = make_dscene<float>(s, &scene_bytes);
  TraceLaunch tl;
  if ((rc = trace_config<float, 1>(d, sc, scene_bytes, &tl))) return rc;
  // pin pageable caller buffers for the duration of a big call
  struct Pin {
    void *p = nullptr;
    ~Pin() {
      if (p) cudaHostUnregister(p);
    }
  } pins[4];
  const void *hostp[4] = {o, dd, t_hit, prim};
  const size_t hostb[4] = {(size_t)n * 12, (size_t)n * 12, (size_t)n * 4, (size_t)n * 4};
  bool pin_env = true;
  if (const char *e = std::getenv("PTB_BATCH_PIN")) pin_env = std::atoi(e) != 0;
  if (nbuf == 2 && pin_env)
    for (int k = 0; k < 4; ++k) {
      cudaPointerAttributes at;
      const bool known = cudaPointerGetAttributes(&at, hostp[k]) == cudaSuccess && at.type != cudaMemoryTypeUnregistered;
      cudaGetLastError();
      if (known) continue;  // already page-locked (ptb_host_alloc, torch pinned memory, ...)
      if (cudaHostRegister(const_cast<void *>(hostp[k]), hostb[k], k < 2 ? cudaHostRegisterReadOnly : cudaHostRegisterDefault) == cudaSuccess)
        pins[k].p = const_cast<void *>(hostp[k]);
      else if (cudaGetLastError(), cudaHostRegister(const_cast<void *>(hostp[k]), hostb[k], cudaHostRegisterDefault) == cudaSuccess)
        pins[k].p = const_cast<void *>(hostp[k]);
      else
        cudaGetLastError();  // not fatal: that array takes the staged copy
    }
  struct Streams {
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t traced[2] = {nullptr, nullptr}, e0 = nullptr, e1 = nullptr;
    ~Streams() {
      for (int i = 0; i < 2; ++i) {
        if (st[i]) cudaStreamDestroy(st[i]);
        if (traced[i]) cudaEventDestroy(traced[i]);
      }
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    }
  } S;
  for (int i = 0; i < nbuf; ++i) {
    CK(cudaStreamCreateWithFlags(&S.st[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&S.traced[i], cudaEventDisableTiming));
  }
  CK(cudaEventCreate(&S.e0));
  CK(cudaEventCreate(&S.e1));
  CK(cudaDeviceSynchronize());  // the non-blocking streams below do not order against earlier default-stream work
  CK(cudaEventRecord(S.e0, S.st[0]));
  uint64_t launches = 0;
  int64_t k = 0;
  for (int64_t first = 0; first < n; first += (int64_t)chunk, ++k) {
    const int b = (int)(k % nbuf);
    const long long m = std::min<int64_t>((int64_t)chunk, n - first);
    cudaStream_t st = S.st[b];
    float *d_o = (float *)pl->batch_buf[0] + (size_t)b * chunk * 3, *d_d = (float *)pl->batch_buf[1] + (size_t)b * chunk * 3;
    float *d_t = (float *)pl->batch_buf[2] + (size_t)b * chunk;
    int32_t *d_p = (int32_t *)pl->batch_buf[3] + (size_t)b * chunk;
    CK(cudaMemcpyAsync(d_o, o + 3 * first, (size_t)m * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_d, dd + 3 * first, (size_t)m * 12, cudaMemcpyHostToDevice, st));
    // the ray queue and the claim cursor are shared by the two streams: the kernels of chunk k wait for chunk k-1's
    if (k > 0) CK(cudaStreamWaitEvent(st, S.traced[(k - 1) % nbuf], 0));
    if ((rc = pack_user_rays(pl, s, sc, d_o, d_d, m, t_min, w.rays, st, &launches))) return rc;
    CK(cudaMemsetAsync(&pl->ctl->cursor[MAX_BOUNCES], 0, sizeof(unsigned), st));
    launch_trace<float, 1>(tl, st, sc, nullptr, 0u, w.rays, nullptr, (unsigned)((m + SEG - 1) / SEG),
                           &pl->ctl->cursor[MAX_BOUNCES], w.mq, (unsigned)w.slots, nullptr, nullptr, 0, nullptr, t_min, t_max, d_t, d_p);
    CK(cudaEventRecord(S.traced[b], st));
    CK(cudaMemcpyAsync(t_hit + first, d_t, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(prim + first, d_p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    launches += 1;
  }
  CK(cudaGetLastError());
  for (int i = 0; i < nbuf; ++i) CK(cudaStreamSynchronize(S.st[i]));
  if (stats) {
    stats->rays = (uint64_t)n;
    stats->kernel_launches += launches;
    stats->h2d_bytes = (uint64_t)n * 24;
    stats->d2h_bytes = (uint64_t)n * 8;
    stats->ms_total = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
    stats->ms_device = stats->ms_total;  // copies and kernels overlap: the pipeline's time is the call's
  }
  return rc;
}

// FP32 issue-rate micro-benchmark: 16 independent FFMA chains per thread, enough resident warps to
// saturate every SM sub-partition.  Returns 1e12 FFMA lane-operations per second (1 FFMA = 1 lane-op).
int ptb_fp32_peak(int32_t device, double *tera_lane_ops, double *ms_out) {
  if (!tera_lane_ops) return fail(PTB_E_INVALID, "fp32_peak: null argument");
  int rc = check_device(device);
  if (rc) return rc;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  float *d_out = nullptr;
  CK(cudaMalloc((void **)&d_out, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    k_fma_peak<<<blocks, threads>>>(d_out, iters, 1.0001f, 0.9999f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, (double)ms);
  }
  CK(cudaGetLastError());
  cudaFree(d_out);
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  const double ops = (double)blocks * threads * (double)iters * 16.0;
  *tera_lane_ops = ops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return PTB_OK;
}

int ptb_r2_stream(int32_t max_bounces, const int32_t *offsets, int64_t n, double *out, int32_t device) {
  if (!offsets || !out || n < 0 || max_bounces < 0 || max_bounces > MAX_BOUNCES)
    return fail(PTB_E_INVALID, "r2_stream: bad args");
  int rc = check_device(device);
  if (rc) return rc;
  RenderConst rcst;
  std::memset(&rcst, 0, sizeof rcst);
  const int D = 2 + 2 * max_bounces;
  lds_alpha(D, rcst.alpha);
  int32_t *d_off = nullptr;
  double *d_out = nullptr;
  const size_t nn = (size_t)std::max<int64_t>(n, 1);
  CK(cudaMalloc((void **)&d_off, nn * 4));
  CK(cudaMalloc((void **)&d_out, nn * D * 8));
  CK(cudaMemcpy(d_off, offsets, (size_t)n * 4, cudaMemcpyHostToDevice));
  const long long total = (long long)n * D;
  if (total > 0) k_r2_stream<<<(unsigned)((total + 255) / 256), 256>>>(rcst, D, d_off, n, d_out);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d_out, (size_t)n * D * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_off), cudaFree(d_out);
  return PTB_OK;
}

int ptb_raygen(const ptb_params *p, int64_t first, int64_t n, int32_t *pixel, int32_t *offset, double *cx,
               double *cy, float *dir_xyz) {
  if (!p || n < 0 || first < 0) return fail(PTB_E_INVALID, "raygen: bad args");
  int rc = check_device(p->device);
  if (rc) return rc;
  DevicePool &tmp = *pool_for(p->device);  // raygen needs no scene
  PoolLock lk(tmp.mu);
  int world = p->tile_world > 0 ? p->tile_world : 1;
  if ((rc = ensure_pixel_list(&tmp, p->width, p->height, p->tile_rank, world))) return rc;
  RenderConst rcst;
  if ((rc = fill_render_const(*p, tmp.npix, &rcst))) return rc;
  const long long total = (long long)tmp.npix * p->samples_per_pixel;
  if (first + n > total) return fail(PTB_E_INVALID, "raygen: sample range exceeds W*H*spp of this rank");
  const size_t nn = (size_t)std::max<int64_t>(n, 1);
  Queue<float> q;
  double *d_cx = nullptr, *d_cy = nullptr;
  const size_t nseg_q = nn / SEG + 1;
  CK(cudaMalloc((void **)&q.base, 3 * nseg_q * SEG * 16));
  CK(cudaMalloc((void **)&q.seg_count, (nn / SEG + 1) * sizeof(int32_t)));
  CK(cudaMalloc((void **)&d_cx, nn * 8));
  CK(cudaMalloc((void **)&d_cy, nn * 8));
  if (n > 0) {
    const int pass0 = (int)(first / tmp.npix), i0 = (int)(first % tmp.npix);
    k_raygen<float><<<(unsigned)((n + 255) / 256), 256>>>(make_gen(rcst, tmp.pixel_list, pass0, i0), (unsigned)n, q, d_cx, d_cy);
  }
  CK(cudaGetLastError());
  std::vector<Vec4<float>> all(3 * nseg_q * SEG), A(nn), B(nn);  // A: the entry's C (attenuation, pixel)
  CK(cudaMemcpy(all.data(), q.base, all.size() * 16, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) {
    const size_t at = Queue<float>::idx((unsigned)i);
    A[i] = all[at + 2 * SEG], B[i] = all[at + SEG];
  }
  if (cx) CK(cudaMemcpy(cx, d_cx, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (cy) CK(cudaMemcpy(cy, d_cy, (size_t)n * 8, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) {
    int32_t pi, oi;
    std::memcpy(&pi, &A[i].w, 4);
    std::memcpy(&oi, &B[i].w, 4);
    if (pixel) pixel[i] = pi;
    if (offset) offset[i] = oi;
    if (dir_xyz) dir_xyz[3 * i] = B[i].x, dir_xyz[3 * i + 1] = B[i].y, dir_xyz[3 * i + 2] = B[i].z;
  }
  cudaFree(q.base), cudaFree(q.seg_count), cudaFree(d_cx), cudaFree(d_cy);
  return PTB_OK;
}

int ptb_first_hit(ptb_scene *s, const ptb_params *p, float *t_hit, int32_t *prim) {
  if (!p || !t_hit || !prim) return fail(PTB_E_INVALID, "first_hit: null argument");
  int rc = require_committed(s, p->device);
  if (rc) return rc;
  DeviceState *d = s->dev;
  DevicePool *pl = d->pool;
  PoolLock lk(pl->mu);
  if ((rc = ensure_pixel_list(pl, p->width, p->height, 0, 1))) return rc;
  RenderConst rcst;
  if ((rc = fill_render_const(*p, pl->npix, &rcst))) return rc;
  const size_t n = (size_t)pl->npix;
  if ((rc = ensure_work<float>(pl, n))) return rc;
  Work<float> &w = pl->work<float>();
  size_t scene_bytes = 0;
  DScene<float> sc = make_dscene<float>(s, &scene_bytes);
  TraceLaunch tl;
  if ((rc = trace_config<float, 1>(d, sc, scene_bytes, &tl))) return rc;
  float *d_t = nullptr;
  int32_t *d_p = nullptr;
  CK(cudaMalloc((void **)&d_t, n * 4));
  CK(cudaMalloc((void **)&d_p, n * 4));
  k_raygen<float><<<(unsigned)((n + 255) / 256), 256>>>(make_gen(rcst, pl->pixel_list, 0, 0), (unsigned)n, w.rays, nullptr, nullptr);
  CK(cudaMemsetAsync(&pl->ctl->cursor[MAX_BOUNCES], 0, sizeof(unsigned), 0));
  launch_trace<float, 1>(tl, 0, sc, nullptr, 0u, w.rays, nullptr, (unsigned)((n + SEG - 1) / SEG),
                         &pl->ctl->cursor[MAX_BOUNCES], w.mq, (unsigned)w.slots, nullptr, nullptr, 0, nullptr, 0.0f, FLT_MAX, d_t, d_p);
  CK(cudaGetLastError());
  std::vector<float> ht(n);
  std::vector<int32_t> hp(n);
  CK(cudaMemcpy(ht.data(), d_t, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hp.data(), d_p, n * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_t), cudaFree(d_p);
  for (size_t k = 0; k < n; ++k) {
    int px = pl->pixel_list_host[k];
    t_hit[px] = ht[k];
    prim[px] = hp[k];
  }
  return PTB_OK;
}

}  // extern "C"

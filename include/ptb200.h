/* ptb200.h — C ABI of libptb200, the B200-native backend for the per-pixel render loop of
 * dalev/path-tracer-ocaml.
 *
 * Where this boundary sits in the reference (paths relative to the reference repo):
 *   - It replaces the whole-image render call `Integrator.create` + `Integrator.render`
 *     (path_tracer/src/integrator.mli:4-16, integrator.ml:71-156) as invoked by
 *     `Render_command.Make(Scene).run` (render_command/src/render_command.ml:64-109).
 *   - It generalises the only native boundary the reference has today,
 *       external spheres_intersect_native : coords -> float -> float -> Ray.t -> float_ref -> int
 *       external leaf_size : unit -> int
 *     (shirley_spheres/bin/main.ml:162-172; sphere-intersect-rs/src/lib.rs:15-18,53-76), from one
 *     ray against one <=16-sphere leaf to a batch of rays against a whole scene
 *     (`ptb_intersect_batch`).
 *   The reference passes closures (`intersect : Ray.t -> Hit.t option`, `background : Ray.t ->
 *   Color.t`, render_command/src/render_command.mli:18-24); closures cannot cross to a GPU, so the
 *   scene crosses as data tables that mirror the reference's variants one to one.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * PTB_E_* code, with a thread-local message behind ptb_last_error(); the caller owns all host
 * buffers; the library owns device memory behind the opaque handle; all scene coordinates arrive
 * ALREADY IN CAMERA SPACE, exactly as the reference moves its scene into camera space before building
 * the tree (shirley_spheres/bin/main.ml:258-260; cornell-box/bin/main.ml:211-218;
 * ganesha/bin/main.ml:74-80).  There is no CPU fallback: every compute entry point fails with
 * PTB_E_NO_DEVICE when no CUDA device is usable.
 *
 * Threading: every entry point may be called from any thread (the OCaml stubs release the runtime lock, ctypes
 * drops the GIL).  The wavefront queues and image buffers belong to the DEVICE, so calls that use the same device —
 * commit, render, intersect_batch, first_hit, raygen — are serialised by a per-device lock inside the library: two
 * domains rendering on one GPU take turns, on different GPUs they run concurrently.  A ptb_scene handle itself
 * must not be modified (set_*, commit, destroy) by one thread while another renders it.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ---------------------------------------------------------------------------- */
enum {
  PTB_OK = 0,
  PTB_E_INVALID = -1,   /* bad argument (the reference's `failwith` / `assert` sites) */
  PTB_E_NO_DEVICE = -2, /* no CUDA device / wrong architecture */
  PTB_E_CUDA = -3,      /* a CUDA runtime call failed; message carries cudaGetErrorString */
  PTB_E_STATE = -4,     /* scene not committed, or modified after commit */
  PTB_E_NOMEM = -5,
  PTB_E_UNSUPPORTED = -6 /* the reference's own "to do" branches (e.g. ascii / big-endian PLY) */
};

/* ---- scene tables ---------------------------------------------------------------------------- */
/* Texture.t (path_tracer/src/texture.ml:16-31): `solid c` | `checker ~width ~height even odd`.
 * The reference's textures are closures; here they are rows.  `even`/`odd` index other rows. */
enum { PTB_TEX_SOLID = 0, PTB_TEX_CHECKER = 1 };
typedef struct ptb_texture {
  int32_t kind;
  int32_t width, height; /* checker only: the ~width/~height arguments, NOT minus one */
  int32_t even, odd;     /* checker only: texture rows (must be SOLID rows) */
  int32_t _pad;
  double rgb[3];         /* solid only */
} ptb_texture;

/* Material.t (path_tracer/src/material.ml:3-9): Lambertian tex | Metal tex | Dielectric {index}.
 * PTB_MAT_EMISSIVE is an EXTENSION beyond the reference: `Material.emit` is black for every
 * kind there (material.ml:59); an Emissive tex material returns `Texture.eval tex coord` from that hook and its
 * `scatter` is `Absorb`, so a path that reaches it ends with `emit0 + attn0 * emit` (integrator.ml:40,43). */
enum { PTB_MAT_LAMBERTIAN = 0, PTB_MAT_METAL = 1, PTB_MAT_DIELECTRIC = 2, PTB_MAT_EMISSIVE = 3 };
typedef struct ptb_material {
  int32_t kind;
  int32_t texture; /* Lambertian / Metal: texture row */
  double index;    /* Dielectric: refractive index (Material.glass = 1.5, material.ml:14) */
} ptb_material;

/* Scene `background : Ray.t -> Color.t`.  GRADIENT_Y is shirley's sky
 * (shirley_spheres/bin/main.ml:104-110): t = 0.5*(normalize(d).y + 1); lerp t c0 c1, evaluated on
 * the camera-space direction.  CONSTANT returns c0. */
enum { PTB_BG_CONSTANT = 0, PTB_BG_GRADIENT_Y = 1 };

/* ---- render parameters ------------------------------------------------------------------------ */
enum {
  PTB_FLAG_F64 = 1,       /* run the whole device pipeline in float64 (validation mode) */
  PTB_FLAG_RAW_SUMS = 2,  /* return filtered sums before `sqrt(x/spp)` (integrator.ml:152-154) */
  PTB_FLAG_NO_FILTER = 4, /* return per-pixel sample sums (no 3x3 splat, no gamma) */
  PTB_FLAG_PROFILE = 8    /* bracket every traversal launch with CUDA events -> stats.ms_trace */
};

typedef struct ptb_params {
  int32_t width, height;     /* Render_command.Args --dimension (render_command.ml:16-25) */
  int32_t samples_per_pixel; /* --samples-per-pixel (render_command.ml:32-35) */
  int32_t max_bounces;       /* --max-ray-bounces  (render_command.ml:40-43) */
  /* Camera.t fields used by Camera.ray (path_tracer/src/camera.ml:50-53,93-102) */
  double lower_left_x, lower_left_y, view_x, view_y;
  /* image-tile sharding (integrator.ml:132-146): this call renders the tiles t of
   * Tile.split ~max_area:1024 with t mod tile_world == tile_rank.  {0,1} = whole image. */
  int32_t tile_rank, tile_world;
  int32_t flags;
  int32_t device;            /* CUDA device ordinal */
  /* pass range: this call renders the sample passes [pass_first, pass_first + pass_count) of the
   * samples_per_pixel passes of integrator.ml:95 (`for pass = 0 to spp - 1`); the R2 offsets stay those of the
   * full render (pixel + pass * spp), so consecutive ranges add up to exactly the full render's samples.
   * {0, 0} = all passes.  ptb_render / ptb_render_multi normalise their image by the passes of the call (a
   * progressive preview); ptb_render_device only adds sums. */
  int32_t pass_first, pass_count;
} ptb_params;

typedef struct ptb_stats {
  uint64_t paths;              /* W*H*spp restricted to this rank's tiles */
  uint64_t rays;               /* number of `intersect` invocations (camera + scattered) */
  uint64_t rays_by_bounce[64];
  uint64_t kernel_launches;    /* kernels of this library launched by the call */
  double ms_total;             /* host wall time of the call */
  double ms_device;            /* CUDA-event time of the render pipeline on its stream */
  double ms_trace;             /* CUDA-event time spent in the traversal kernel (if measured) */
  double ms_h2d, ms_d2h;
  uint64_t h2d_bytes, d2h_bytes;
} ptb_stats;

typedef struct ptb_scene ptb_scene;

/* ---- entry points ----------------------------------------------------------------------------- */
int ptb_device_count(void);
const char *ptb_last_error(void);
const char *ptb_version(void);

/* replaces `leaf_size : unit -> int` (sphere-intersect-rs/src/lib.rs:15-18): the widest leaf the
 * device tree stores. */
int ptb_leaf_size(void);

ptb_scene *ptb_scene_create(void);
void ptb_scene_destroy(ptb_scene *);
int ptb_scene_set_textures(ptb_scene *, const ptb_texture *, int32_t n);
int ptb_scene_set_materials(ptb_scene *, const ptb_material *, int32_t n);
/* Sphere.t list (sphere/src/sphere.ml:4-8) as the SoA `coords` record of Simd_leaf
 * (shirley_spheres/bin/main.ml:137-142): four flat double arrays + a material row per sphere. */
int ptb_scene_set_spheres(ptb_scene *, const double *xs, const double *ys, const double *zs,
                          const double *rs, const int32_t *material, int64_t n);
/* Indexed triangles as ganesha's Mesh SoA (ganesha/bin/main.ml:37-43,99-110): vertex columns +
 * 3 indices per face; `material` per face (NULL = row 0); `uv` = 6 doubles per face
 * (ua,va,ub,vb,uc,vc; Triangle Face.tex_coords, triangle/triangle.mli) or NULL for
 * (t00,t01,t11) as ganesha uses (ganesha/bin/main.ml:111). */
int ptb_scene_set_triangles(ptb_scene *, const double *vx, const double *vy, const double *vz,
                            int64_t n_vertices, const int32_t *indices, const int32_t *material,
                            const double *uv, int64_t n_triangles);
int ptb_scene_set_background(ptb_scene *, int32_t kind, const double c0[3], const double c1[3]);
/* EXTENSION beyond the reference, behind the hook the reference already threads:
 * `Integrator.create ~diffuse_plus_light` (integrator.mli:13; always `Pdf.diffuse` today, render_command.ml:81;
 * `Pdf.t` has the single constructor `Diffuse`, pdf.ml:3).  With a light quad set, diffuse scattering samples
 * `Pdf.Mix (Diffuse, Quad_light {origin; u; v})` — equal weights; the first sample coordinate picks the component —
 * and weighs the path by `Pdf.eval diffuse / Pdf.eval mix` exactly as integrator.ml:48-66 spells out.  The quad is
 * the parallelogram origin + a*u + b*v, a, b in [0,1], in camera space; the emitting geometry itself is ordinary
 * triangles with a PTB_MAT_EMISSIVE material.  NULL origin removes the light (back to `Pdf.diffuse`). */
int ptb_scene_set_light_quad(ptb_scene *, const double origin[3], const double u[3], const double v[3]);
/* Shape_tree.create (path_tracer/src/shape_tree.ml:252-263): builds the tree (on the host, or on the device for
 * pure triangle meshes of >= 200 k triangles) and uploads it to `device`.  Returns build+upload milliseconds through *ms if non-NULL. */
int ptb_scene_commit(ptb_scene *, int32_t device, double *ms);
int64_t ptb_scene_primitive_count(const ptb_scene *);
/* Shape of the committed device tree (the reference prints depth and a leaf-length histogram,
 * shirley_spheres/bin/main.ml:263-267): out = {wide nodes, depth, worst-case stack entries, spheres,
 * triangle slots, leaves, max leaf size, 0}.  Triangle slots = triangles, except for soups of overlapping triangles,
 * which the builders pre-split (csrc/presplit.hpp): then a triangle occupies one slot per box reference;
 * ptb_scene_primitive_count still counts the caller's shapes, and hits always name the caller's triangle. */
int ptb_scene_tree_stats(const ptb_scene *, int32_t out[8]);

/* Integrator.render (integrator.ml:130-156) with HOST image: `image_rgb` is 3*W*H doubles laid out
 * (y*W + x)*3 + c, row 0 = top — the Bimage f64 rgb layout the reference allocates
 * (render_command.ml:65).  Includes the host<->device copies.  The work buffers (wavefront queues, device
 * image) belong to the DEVICE and are reused between calls: one render or intersect_batch at a time per
 * device and process — like the reference, which renders one image per process. */
int ptb_render(ptb_scene *, const ptb_params *, double *image_rgb, ptb_stats *);
/* `update_progress : int -> unit` (integrator.ml:130,150; the Progress bar of render_command.ml:86-104) as a POLLED
 * counter, because the library never calls back into the host: paths finished / paths in total of the render that
 * is running (or ran last) on `device`, readable from any other thread while ptb_render / ptb_render_device /
 * ptb_render_multi is in flight.  The counter advances once per wavefront batch. */
int ptb_render_progress(int32_t device, uint64_t *paths_done, uint64_t *paths_total);

/* Single-process multi-GPU render (BASELINE.json configs[3]): the scene is replicated on devices
 * 0..n_devices-1 (the tree is built once), device i renders the tiles t = i (mod n_devices) of the reference's tile
 * list (Tile.split ~max_area:1024, integrator.ml:132-133) on its own host thread, then device 0 adds the other
 * devices' per-pixel sums straight out of their memory (peer loads over NVLink) and resolves.  Same image layout
 * as ptb_render; float32 pipeline only.  This is what the CLI twin's --gpus N calls; one process per GPU with an
 * NCCL reduce (ptb_render_device + ptb_resolve_device) is the other way to shard. */
int ptb_scene_commit_multi(ptb_scene *, int32_t n_devices, double *ms);
int ptb_render_multi(ptb_scene *, const ptb_params *, int32_t n_devices, double *image_rgb, ptb_stats *);

/* Same pipeline, device-resident output: adds this rank's per-pixel sample sums into `d_sums`
 * (float32[3*W*H], same layout, device memory on params->device), enqueued on `stream`
 * (a cudaStream_t; NULL = default stream).  The caller zeroes d_sums, reduces it across ranks, and
 * calls ptb_resolve_device once. */
int ptb_render_device(ptb_scene *, const ptb_params *, float *d_sums, void *stream, ptb_stats *);
/* Filter_kernel splat (film_tile.ml:23-38; integrator.ml:114-128) + gamma (integrator.ml:152-154):
 * d_image[p] = sqrt( (sum_d w(d) * d_sums[p-d]) / spp ), taps outside the image dropped. */
int ptb_resolve_device(const float *d_sums, float *d_image, int32_t width, int32_t height,
                       int32_t samples_per_pixel, int32_t flags, int32_t device, void *stream);

/* Batched generalisation of spheres_intersect_native (lib.rs:53-76): n rays against the committed
 * scene; origins/directions are 3*n floats (x,y,z interleaved); writes nearest t (NaN on miss, as
 * the reference initialises t_hit_ref, main.ml:207) and primitive index (-1 on miss, lib.rs:74;
 * spheres first then triangles, in the order they were set).  The library may trace the rays in another
 * order than the caller's (csrc/ray_sort.cuh: spatial order on big triangle soups); result i always belongs to ray i
 * and does not depend on that order. */
int ptb_intersect_batch(ptb_scene *, const float *origins, const float *directions, float t_min,
                        float t_max, int64_t n, float *t_hit, int32_t *prim, int32_t device,
                        ptb_stats *);
int ptb_intersect_batch_device(ptb_scene *, const float *d_origins, const float *d_directions,
                               float t_min, float t_max, int64_t n, float *d_t_hit, int32_t *d_prim,
                               int32_t device, void *stream, ptb_stats *);

/* Page-locked host memory for callers that want ptb_intersect_batch / ptb_render to copy at full PCIe speed
 * (the OCaml side would allocate its Bigarrays once through these).  Plain malloc'ed buffers work too; they are
 * pinned for the duration of a big call (cudaHostRegister) or staged. */
void *ptb_host_alloc(uint64_t bytes);
void ptb_host_free(void *);

/* Low_discrepancy_sequence.get (low_discrepancy_sequence.ml:33-36) evaluated ON THE DEVICE by the
 * same device function the ray generator uses: out[i*D + d] for offsets[i], d < D = 2+2*max_bounces.
 * Used to prove the sample stream is bit-exact. */
int ptb_r2_stream(int32_t max_bounces, const int32_t *offsets, int64_t n, double *out,
                  int32_t device);
/* The camera-ray generator alone (integrator.ml:98-105 + camera.ml:93-102) for `n` samples starting
 * at linear sample index `first` of this rank's enumeration: writes pixel index, R2 offset, cx, cy
 * (float64, bit-exact contract) and the normalized direction. Any output may be NULL. */
int ptb_raygen(const ptb_params *, int64_t first, int64_t n, int32_t *pixel, int32_t *offset,
               double *cx, double *cy, float *dir_xyz);
/* First-hit map of the camera rays of pass 0 (one ray per pixel, through the production raygen and
 * traversal kernels): t (NaN on miss) and primitive index per pixel. */
int ptb_first_hit(ptb_scene *, const ptb_params *, float *t_hit, int32_t *prim);

/* Measures this device's FP32 FFMA issue rate (the roofline denominator SURVEY.md §8d asks for):
 * 1e12 lane-operations/s, one FFMA on one lane = one lane-op.  Diagnostic, not on the render path. */
int ptb_fp32_peak(int32_t device, double *tera_lane_ops, double *ms);

/* Host-side helpers that mirror small reference functions the callers need. */
/* Low_discrepancy_sequence.create (low_discrepancy_sequence.ml:8-31): alpha[0..dimension). */
int ptb_lds_alpha(int32_t dimension, double *alpha);
/* Tile.split ~max_area (path_tracer/src/tile.ml:30-39): fills row/col/width/height; returns count
 * (or the needed capacity, negated minus 1000, if cap is too small). */
int ptb_tile_split(int32_t width, int32_t height, int32_t max_area, int32_t *row, int32_t *col,
                   int32_t *w, int32_t *h, int32_t cap);
/* Filter_kernel.Binomial.create ~order ~pixel_radius (filter_kernel.ml:49-85): (2r+1)^2 weights. */
int ptb_filter_binomial(int32_t order, int32_t pixel_radius, double *weights);
/* Camera.create + Mat4.look_at (camera.ml:14-27,58-83): out[0..3] = lower_left_x, lower_left_y,
 * view_x, view_y; out[4..19] = look_at rows (row-major 4x4). */
int ptb_camera_create(const double eye[3], const double target[3], const double up[3],
                      double aspect, double vertical_fov_deg, double out[20]);
/* Camera.transform (camera.ml:39-43,91) applied in place to n points. */
int ptb_camera_transform(const double look_at[16], double *xs, double *ys, double *zs, int64_t n);

/* Triangle pre-splitting as both tree builders apply it to soups of overlapping triangles (csrc/presplit.hpp; the
 * reference's Shape_tree.create, shape_tree.ml:252-263, has one box per shape): the boxes of the pieces of ONE
 * triangle (v = 9 doubles: corner-major x, y, z) for the cell size `cell` on the grid anchored at `origin`.  Writes
 * min(count, cap) boxes as lo[3], hi[3] (6 doubles each) and returns the count.  Host only; diagnostic / tests. */
int ptb_presplit_boxes(const double *v, double cell, const double origin[3], double *boxes, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */

/* oracle.h — C API of the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A float64 restatement of the reference's CPU path (dalev/path-tracer-ocaml), function by function,
 * used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
 * the CHECKER.  Nothing in the product (libptb200, path_tracer_ocaml_b200/) may include, link, import
 * or execute anything in this directory.
 *
 * kind = "port": the reference is OCaml 5 + Rust and neither toolchain exists in this image, so
 * oracle/_ref cannot be built.  What pins this restatement, and what does not, is
 * listed at the top of oracle.cpp.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include "../include/ptb200.h" /* POD table types only (ptb_texture, ptb_material, ptb_params) */

#ifdef __cplusplus
extern "C" {
#endif

/* leaf flavours of Shape_tree.Make(L) */
enum {
  ORC_LEAF_SIMD = 0,  /* Simd_leaf: spheres only, <=16 SoA, Rust/AVX kernel (main.ml:132-218) */
  ORC_LEAF_ARRAY = 1  /* Shape_tree.Array_leaf with scalar Sphere.intersect / Triangle.intersect */
};

typedef struct orc_scene orc_scene;

typedef struct orc_counters {
  uint64_t paths, rays, box_tests, sphere_tests, tri_tests, leaf_visits, hits;
  uint64_t scatter_lambert, scatter_metal, scatter_dielectric, absorbed, missed, exhausted;
  uint64_t rays_by_bounce[64];
} orc_counters;

/* --- small functions, for known-answer tests ---------------------------------------------------- */
double orc_lds_phi(int dimension);
void orc_lds_alpha(int dimension, double *alpha);
double orc_lds_get(const double *alpha, int64_t offset, int dimension);
void orc_filter_binomial(int order, int pixel_radius, double *weights);
int orc_tile_split(int width, int height, int max_area, int32_t *row, int32_t *col, int32_t *w,
                   int32_t *h, int cap);
int orc_bbox_is_hit(const double bmin[3], const double bmax[3], const double o[3],
                    const double d[3], double t_min, double t_max);
void orc_camera_create(const double eye[3], const double target[3], const double up[3],
                       double aspect, double vfov_deg, double out[20]);
void orc_camera_transform(const double look_at[16], double *xs, double *ys, double *zs, int64_t n);
void orc_camera_ray(const double cam4[4], double cx, double cy, double dir[3]);
void orc_unit_square_to_hemisphere(double u, double v, double out[3]);
/* Film_tile.create + write_pixel ~x ~y color, then Film_tile.iter: out is (w+2)*(h+2)*3 */
void orc_film_tile_write_pixel(int tile_w, int tile_h, int x, int y, const double rgb[3],
                               double *out);
/* Sphere.intersect (scalar OCaml, sphere.ml:35-54): returns 1 and *t on hit */
int orc_sphere_intersect_scalar(const double c[3], double r, const double o[3], const double d[3],
                                double t_min, double t_max, double *t);
/* spheres_intersect_aux (lib.rs:102-178) on padded SoA arrays of length len (multiple of 4) */
int orc_spheres_intersect_simd(const double *xs, const double *ys, const double *zs,
                               const double *rs, int len, const double o[3], const double d[3],
                               double t_min, double t_max, double *t);
/* Triangle.intersect (triangle.ml:74-98): returns 1 and t,u,v */
int orc_triangle_intersect(const double a[3], const double b[3], const double c[3],
                           const double o[3], const double d[3], double t_min, double t_max,
                           double *t, double *u, double *v);
/* Shader_space.create n p; then rotate / rotate_inv of v (shader_space.ml:11-32) */
void orc_shader_space_rotate(const double n[3], const double v[3], int inverse, double out[3]);

/* --- scenes --------------------------------------------------------------------------------------- */
orc_scene *orc_scene_create(void);
void orc_scene_destroy(orc_scene *);
void orc_scene_set_textures(orc_scene *, const ptb_texture *, int n);
void orc_scene_set_materials(orc_scene *, const ptb_material *, int n);
void orc_scene_set_spheres(orc_scene *, const double *xs, const double *ys, const double *zs,
                           const double *rs, const int32_t *material, int64_t n);
void orc_scene_set_triangles(orc_scene *, const double *vx, const double *vy, const double *vz,
                             int64_t nv, const int32_t *idx, const int32_t *material,
                             const double *uv, int64_t nt);
void orc_scene_set_background(orc_scene *, int kind, const double c0[3], const double c1[3]);
/* EXTENSION (not in the reference): ~diffuse_plus_light = Pdf.Mix (Diffuse, Quad_light {origin; u; v}); NULL origin =
 * back to Pdf.diffuse.  See ptb_scene_set_light_quad in ptb200.h. */
void orc_scene_set_light_quad(orc_scene *, const double origin[3], const double u[3], const double v[3]);
/* prim_order: list order handed to Shape_tree.create; entry >= 0 = sphere i, < 0 = triangle ~entry.
 * NULL = spheres then triangles. */
int orc_scene_commit(orc_scene *, int leaf_kind, int length_cutoff, const int32_t *prim_order,
                     int64_t n_order);
int orc_scene_tree_depth(const orc_scene *);
int orc_scene_leaf_histogram(const orc_scene *, int32_t *sizes, int32_t *counts, int cap);
int64_t orc_scene_node_count(const orc_scene *);

/* Integrator.create + render (integrator.ml:71-156).  image: 3*W*H doubles, (y*W+x)*3+c.
 * n_threads <= 1: tiles rendered and stitched in list order (deterministic).  n_threads > 1:
 * n_threads-1 workers + this thread stitching, as integrator.ml:137-151.
 * flags: PTB_FLAG_RAW_SUMS / PTB_FLAG_NO_FILTER as in ptb200.h.  pass_limit > 0 renders only the
 * first pass_limit passes but keeps spp in the offset formula and the gamma (bounded CPU samples). */
int orc_render(orc_scene *, const ptb_params *, double *image, orc_counters *, int n_threads,
               int pass_limit);
/* radiance of one sample: pixel (gx,gy), pass — Integrator.render_tile's inner body */
void orc_trace_sample(orc_scene *, const ptb_params *, int gx, int gy, int pass, double rgb[3],
                      orc_counters *);
/* Scene.intersect on explicit rays (float64): t (NaN on miss) and primitive id in set order
 * (spheres first, then triangles), -1 on miss */
void orc_intersect_batch(orc_scene *, const double *o, const double *d, double t_min, double t_max,
                         int64_t n, double *t_hit, int32_t *prim, orc_counters *, int n_threads);
/* the same against one Array_leaf holding every primitive (shape_tree.ml:299-311): no tree, no commit needed */
void orc_intersect_batch_linear(orc_scene *, const double *o, const double *d, double t_min, double t_max,
                                int64_t n, double *t_hit, int32_t *prim, int n_threads);
/* camera rays of pass 0 for every pixel */
void orc_first_hit(orc_scene *, const ptb_params *, double *t_hit, int32_t *prim, double *cx,
                   double *cy);

#ifdef __cplusplus
}
#endif
#endif

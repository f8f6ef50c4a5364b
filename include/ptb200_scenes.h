/* ptb200_scenes.h — the scene DATA of the reference's three scene binaries, as C ABI loaders.
 *
 * In the reference a scene is OCaml code that builds closures (shirley_spheres/bin/main.ml:26-110,
 * cornell-box/bin/main.ml:43-91,170-219, ganesha/bin/main.ml:30-119,205-260).  Here the same
 * geometry/material tables are written into a ptb_scene (already in camera space) so the C++ CLI
 * twins and the Python harness render exactly the same input; the getters let a checker read the
 * tables back and feed them to another implementation.
 */
#ifndef PTB200_SCENES_H
#define PTB200_SCENES_H
#include "ptb200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* shirley_spheres (main.ml:26-110,250-260): ground r=1000 checker, 3 big spheres, 23x23 jittered
 * small spheres drawn from OCaml 5's Random (LXM) seeded with `seed` (the reference uses 42).
 * cam[20] as ptb_camera_create.  `Random.float` is Base's (main.ml opens Base): two 30-bit draws per float.  With
 * seed 42 this is the reference's sphere field — the oracle's render of it equals shirley-spheres.png pixel by pixel
 * (tests/test_golden_png.py). */
int ptb_scene_load_shirley(ptb_scene *, double aspect, int32_t seed, double cam[20]);

/* cornell-box geometry (main.ml:43-91,170-219): 18 triangles + 3 spheres, camera eye (.5,.5,-1).
 * The reference renders it with photon mapping only and has no background; `background_kind` /
 * c0 / c1 are the caller's choice. */
int ptb_scene_load_cornell(ptb_scene *, double aspect, int32_t background_kind, const double c0[3],
                           const double c1[3], double cam[20]);

/* EXTENSION (BASELINE.json configs[1] "diffuse+light sampling"; beyond the reference): the
 * same cornell geometry in a closed, black-background box, lit by an emissive square (PTB_MAT_EMISSIVE, solid
 * `radiance`) of side 0.1 at the reference's `light_pos` (0.5, 0.82, 0.5) inside the mirror enclosure
 * (main.ml:184-210), with ptb_scene_set_light_quad on that square. */
int ptb_scene_load_cornell_lit(ptb_scene *, double aspect, const double radiance[3], double cam[20]);

/* ganesha assembly (ganesha/bin/main.ml:30-119,205-260) from an indexed mesh in WORLD space:
 * camera eye (328,70.282,345)->(328,10,0) fov 30, all faces lambert (0.1,0.7,0.2) with tex
 * (t00,t01,t11), plus the 10000x10000 checker floor (2 triangles) at the mesh's camera-space
 * bbox-min y.  `xyz` = 3*nv floats (x,y,z interleaved, as a PLY vertex element stores them). */
int ptb_scene_load_mesh(ptb_scene *, const float *xyz, int64_t n_vertices, const int32_t *faces,
                        int64_t n_faces, double aspect, double cam[20]);

/* Seeded synthetic stand-in for ganesha.ply (not in the repo): a displaced icosphere
 * with about `target_faces` triangles placed where the ganesha mesh sits.  Writes up to the given
 * capacities; returns counts through nv/nf. */
int ptb_mesh_synthetic(int64_t target_faces, uint32_t seed, float *xyz, int64_t cap_vertices,
                       int32_t *faces, int64_t cap_faces, int64_t *nv, int64_t *nf);

/* The subset of PLY that Ply.of_bigstring reads (ply_format/src/ply.ml:340-352: "ply\n" magic, header up to
 * end_header, binary_little_endian 1.0 only, elements that are all-atomic or exactly one list property), reduced
 * to what ganesha takes from it (ganesha/bin/main.ml:50-60,182-185): the x,y,z columns of element "vertex"
 * (float or double, returned as float32 x,y,z interleaved) and the rows of the list property "vertex_indices"
 * (exactly 3 per row).  On success *xyz and *faces are malloc'ed by the library: release them with ptb_free. */
int ptb_ply_read_mesh(const char *path, float **xyz, int64_t *n_vertices, int32_t **faces, int64_t *n_faces);
int ptb_ply_parse_mesh(const void *bytes, int64_t len, float **xyz, int64_t *n_vertices, int32_t **faces,
                       int64_t *n_faces);
void ptb_free(void *);

/* table getters (sizes first with NULL buffers) */
int ptb_scene_counts(const ptb_scene *, int64_t *n_spheres, int64_t *n_vertices,
                     int64_t *n_triangles, int32_t *n_materials, int32_t *n_textures);
int ptb_scene_get_spheres(const ptb_scene *, double *xs, double *ys, double *zs, double *rs,
                          int32_t *material);
int ptb_scene_get_triangles(const ptb_scene *, double *vx, double *vy, double *vz, int32_t *indices,
                            int32_t *material, double *uv);
int ptb_scene_get_materials(const ptb_scene *, ptb_material *, ptb_texture *);
int ptb_scene_get_background(const ptb_scene *, int32_t *kind, double c0[3], double c1[3]);
int ptb_scene_get_light_quad(const ptb_scene *, int32_t *has_light, double origin[3], double u[3], double v[3]);
/* reference list order handed to Shape_tree.create: entry >= 0 sphere i, < 0 triangle ~entry */
int ptb_scene_get_prim_order(const ptb_scene *, int32_t *order, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif

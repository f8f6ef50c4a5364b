/* ptb_stubs.c — C stubs between OCaml 5 and libptb200 (include/ptb200.h); see ptb.ml.
 *
 * STATUS: no OCaml toolchain exists in this image, so this file is compiled (-Wall -Werror), linked and RUN against
 * a mock of the OCaml C runtime (tests/mock_caml/, tests/test_ocaml_stubs.py): every setter's marshalling, the
 * Failure paths, and — on a GPU box — a render and an intersect_batch whose results equal the ctypes path's.  Not
 * checked: the real runtime's GC interaction (CAMLparam rooting) and ptb.ml's typing.  Same route as sphere-intersect-rs
 * (sphere-intersect-rs/src/lib.rs:53-76) one level up: one call renders the image.  Unlike that [@@noalloc]
 * per-leaf call, these release the runtime lock around the device work and raise Failure on error
 * (the reference's own style: `failwith`, shape_tree.ml:254-255). */
#include <string.h>
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>
#include "ptb200.h"

#define Scene_val(v) (*((ptb_scene **)Data_custom_val(v)))
static void finalize_scene(value v) { ptb_scene_destroy(Scene_val(v)); }
static struct custom_operations scene_ops = {"ptb.scene", finalize_scene, custom_compare_default, custom_hash_default,
                                             custom_serialize_default, custom_deserialize_default,
                                             custom_compare_ext_default, custom_fixed_length_default};
static void check(int rc) { if (rc < 0) caml_failwith(ptb_last_error()); }

CAMLprim value ptb_ml_device_count(value unit) { (void)unit; return Val_int(ptb_device_count()); }

CAMLprim value ptb_ml_scene_create(value unit) {
  CAMLparam1(unit);
  CAMLlocal1(v);
  v = caml_alloc_custom(&scene_ops, sizeof(ptb_scene *), 0, 1);
  Scene_val(v) = ptb_scene_create();
  if (!Scene_val(v)) caml_failwith("ptb_scene_create failed");
  CAMLreturn(v);
}

/* texture_row: Solid of float*float*float (tag 0) | Checker of {width;height;even;odd} (tag 1) */
CAMLprim value ptb_ml_set_textures(value s, value rows) {
  CAMLparam2(s, rows);
  mlsize_t n = Wosize_val(rows);
  ptb_texture *t = (ptb_texture *)caml_stat_alloc(sizeof(ptb_texture) * (n ? n : 1));
  for (mlsize_t i = 0; i < n; ++i) {
    value r = Field(rows, i);
    memset(&t[i], 0, sizeof t[i]);
    if (Tag_val(r) == 0) {
      t[i].kind = PTB_TEX_SOLID;
      for (int c = 0; c < 3; ++c) t[i].rgb[c] = Double_val(Field(r, c));
    } else {
      t[i].kind = PTB_TEX_CHECKER;
      t[i].width = Int_val(Field(r, 0)), t[i].height = Int_val(Field(r, 1));
      t[i].even = Int_val(Field(r, 2)), t[i].odd = Int_val(Field(r, 3));
    }
  }
  int rc = ptb_scene_set_textures(Scene_val(s), t, (int32_t)n);
  caml_stat_free(t);
  check(rc);
  CAMLreturn(Val_unit);
}

/* material_row: Lambertian of int (tag 0) | Metal of int (tag 1) | Dielectric of float (tag 2) | Emissive of int (tag 3,
   the extension of ptb200.h: Material.emit = the texture, scatter = Absorb) */
CAMLprim value ptb_ml_set_materials(value s, value rows) {
  CAMLparam2(s, rows);
  mlsize_t n = Wosize_val(rows);
  ptb_material *m = (ptb_material *)caml_stat_alloc(sizeof(ptb_material) * (n ? n : 1));
  for (mlsize_t i = 0; i < n; ++i) {
    value r = Field(rows, i);
    memset(&m[i], 0, sizeof m[i]);
    switch (Tag_val(r)) {
      case 0: m[i].kind = PTB_MAT_LAMBERTIAN, m[i].texture = Int_val(Field(r, 0)), m[i].index = 1.0; break;
      case 1: m[i].kind = PTB_MAT_METAL, m[i].texture = Int_val(Field(r, 0)), m[i].index = 1.0; break;
      case 3: m[i].kind = PTB_MAT_EMISSIVE, m[i].texture = Int_val(Field(r, 0)), m[i].index = 1.0; break;
      default: m[i].kind = PTB_MAT_DIELECTRIC, m[i].texture = -1, m[i].index = Double_val(Field(r, 0)); break;
    }
  }
  int rc = ptb_scene_set_materials(Scene_val(s), m, (int32_t)n);
  caml_stat_free(m);
  check(rc);
  CAMLreturn(Val_unit);
}

CAMLprim value ptb_ml_set_spheres(value s, value xs, value ys, value zs, value rs, value mat) {
  CAMLparam5(s, xs, ys, zs, rs);
  CAMLxparam1(mat);
  check(ptb_scene_set_spheres(Scene_val(s), (const double *)Caml_ba_data_val(xs), (const double *)Caml_ba_data_val(ys),
                              (const double *)Caml_ba_data_val(zs), (const double *)Caml_ba_data_val(rs),
                              (const int32_t *)Caml_ba_data_val(mat), Caml_ba_array_val(rs)->dim[0]));
  CAMLreturn(Val_unit);
}
CAMLprim value ptb_ml_set_spheres_bc(value *a, int n) { (void)n; return ptb_ml_set_spheres(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value ptb_ml_set_triangles(value s, value vx, value vy, value vz, value idx, value mat) {
  CAMLparam5(s, vx, vy, vz, idx);
  CAMLxparam1(mat);
  check(ptb_scene_set_triangles(Scene_val(s), (const double *)Caml_ba_data_val(vx), (const double *)Caml_ba_data_val(vy),
                                (const double *)Caml_ba_data_val(vz), Caml_ba_array_val(vx)->dim[0],
                                (const int32_t *)Caml_ba_data_val(idx), (const int32_t *)Caml_ba_data_val(mat), NULL,
                                Caml_ba_array_val(idx)->dim[0] / 3));
  CAMLreturn(Val_unit);
}
CAMLprim value ptb_ml_set_triangles_bc(value *a, int n) { (void)n; return ptb_ml_set_triangles(a[0], a[1], a[2], a[3], a[4], a[5]); }

/* background: Constant of float*float*float (tag 0) | Gradient_y of (f*f*f)*(f*f*f) (tag 1) */
CAMLprim value ptb_ml_set_background(value s, value bg) {
  CAMLparam2(s, bg);
  double c0[3], c1[3] = {0, 0, 0};
  if (Tag_val(bg) == 0) {
    for (int c = 0; c < 3; ++c) c0[c] = Double_val(Field(bg, c));
    check(ptb_scene_set_background(Scene_val(s), PTB_BG_CONSTANT, c0, NULL));
  } else {
    for (int c = 0; c < 3; ++c) c0[c] = Double_val(Field(Field(bg, 0), c)), c1[c] = Double_val(Field(Field(bg, 1), c));
    check(ptb_scene_set_background(Scene_val(s), PTB_BG_GRADIENT_Y, c0, c1));
  }
  CAMLreturn(Val_unit);
}

/* light quad: (origin, u, v), three float triples (extension: ~diffuse_plus_light = Mix (Diffuse, Quad_light)) */
CAMLprim value ptb_ml_set_light_quad(value s, value quad) {
  CAMLparam2(s, quad);
  double o[3], u[3], v[3];
  for (int c = 0; c < 3; ++c)
    o[c] = Double_val(Field(Field(quad, 0), c)), u[c] = Double_val(Field(Field(quad, 1), c)), v[c] = Double_val(Field(Field(quad, 2), c));
  check(ptb_scene_set_light_quad(Scene_val(s), o, u, v));
  CAMLreturn(Val_unit);
}

/* update_progress (integrator.ml:130,150) as a poll: (paths done, paths in total) of the render in flight on `device`.
   Cheap and lock-free: meant to be called from another domain while ptb_ml_render has the runtime lock released. */
CAMLprim value ptb_ml_render_progress(value device) {
  CAMLparam1(device);
  CAMLlocal1(pair);
  uint64_t done = 0, total = 0;
  check(ptb_render_progress(Int_val(device), &done, &total));
  pair = caml_alloc(2, 0);
  Store_field(pair, 0, Val_long((intnat)done));
  Store_field(pair, 1, Val_long((intnat)total));
  CAMLreturn(pair);
}

CAMLprim value ptb_ml_commit(value s, value device) {
  CAMLparam2(s, device);
  double ms = 0;
  ptb_scene *sc = Scene_val(s);
  int dev = Int_val(device);
  caml_release_runtime_system();  /* the tree build takes up to seconds for big meshes */
  int rc = ptb_scene_commit(sc, dev, &ms);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(caml_copy_double(ms));
}

/* NOT [@@noalloc]: a render takes seconds; the runtime lock is released so other domains keep running.
   The image bigarray's data is malloc'ed, so the GC does not move it meanwhile. */
CAMLprim value ptb_ml_render(value s, value w, value h, value spp, value mb, value cam, value device, value img) {
  CAMLparam5(s, w, h, spp, mb);
  CAMLxparam3(cam, device, img);
  ptb_params p;
  memset(&p, 0, sizeof p);
  p.width = Int_val(w), p.height = Int_val(h), p.samples_per_pixel = Int_val(spp), p.max_bounces = Int_val(mb);
  p.lower_left_x = Double_val(Field(cam, 0)), p.lower_left_y = Double_val(Field(cam, 1)); /* camera.ml:50-53 */
  p.view_x = Double_val(Field(cam, 2)), p.view_y = Double_val(Field(cam, 3));
  p.tile_rank = 0, p.tile_world = 1, p.flags = 0, p.device = Int_val(device);
  if (Caml_ba_array_val(img)->dim[0] < (intnat)3 * p.width * p.height) caml_invalid_argument("Ptb.render: image too small");
  ptb_scene *sc = Scene_val(s);
  double *image = (double *)Caml_ba_data_val(img);
  ptb_stats st;
  caml_release_runtime_system();
  int rc = ptb_render(sc, &p, image, &st);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(caml_copy_double(st.ms_device));
}
CAMLprim value ptb_ml_render_bc(value *a, int n) { (void)n; return ptb_ml_render(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); }

CAMLprim value ptb_ml_intersect_batch(value s, value o, value d, value tmin, value tmax, value device, value t_hit, value prim) {
  CAMLparam5(s, o, d, tmin, tmax);
  CAMLxparam3(device, t_hit, prim);
  ptb_scene *sc = Scene_val(s);
  const float *po = (const float *)Caml_ba_data_val(o), *pd = (const float *)Caml_ba_data_val(d);
  float *pt = (float *)Caml_ba_data_val(t_hit);
  int32_t *pp = (int32_t *)Caml_ba_data_val(prim);
  int64_t n = Caml_ba_array_val(t_hit)->dim[0];
  float t0 = (float)Double_val(tmin), t1 = (float)Double_val(tmax);
  int dev = Int_val(device);
  caml_release_runtime_system();
  int rc = ptb_intersect_batch(sc, po, pd, t0, t1, n, pt, pp, dev, NULL);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value ptb_ml_intersect_batch_bc(value *a, int n) { (void)n; return ptb_ml_intersect_batch(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); }

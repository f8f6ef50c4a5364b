/* Drives integration/ocaml/ptb_stubs.c the way OCaml code compiled from ptb.ml would — same value shapes
 * (constructor blocks, boxed floats, tuples, bigarrays, a custom block for the scene) — on top of the mock runtime
 * in tests/mock_caml/caml/.  Test infrastructure: it proves the reference-side binding compiles against
 * include/ptb200.h, marshals every argument correctly and (on a GPU box) renders the same image as the ctypes path.
 *
 *   stub_driver nogpu           -> builds the scene, expects commit to raise Failure (no CUDA device)
 *   stub_driver render OUT.f64  -> builds, commits on device 0, renders 64x32 @ 16 spp, 6 bounces into OUT.f64
 * The scene is the one tests/test_ocaml_stubs.py builds through the Python wrapper. */
#include <math.h>
#include <stdio.h>
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>
#include "ptb200.h"

jmp_buf mock_caml_handler;
char mock_caml_exn[512];
int mock_caml_runtime_released = 0;

value ptb_ml_device_count(value);
value ptb_ml_scene_create(value);
value ptb_ml_set_textures(value, value);
value ptb_ml_set_materials(value, value);
value ptb_ml_set_spheres_bc(value *, int);
value ptb_ml_set_triangles_bc(value *, int);
value ptb_ml_set_background(value, value);
value ptb_ml_commit(value, value);
value ptb_ml_render_bc(value *, int);
value ptb_ml_intersect_batch_bc(value *, int);
value ptb_ml_set_light_quad(value, value);
value ptb_ml_render_progress(value);

static value boxed3(int tag, double a, double b, double c) {  /* C of float * float * float, or a float triple */
  value v = caml_alloc(3, tag);
  Field(v, 0) = caml_copy_double(a), Field(v, 1) = caml_copy_double(b), Field(v, 2) = caml_copy_double(c);
  return v;
}
static value ints4(int tag, int a, int b, int c, int d) {
  value v = caml_alloc(4, tag);
  Field(v, 0) = Val_int(a), Field(v, 1) = Val_int(b), Field(v, 2) = Val_int(c), Field(v, 3) = Val_int(d);
  return v;
}
static value one_int(int tag, int a) {
  value v = caml_alloc(1, tag);
  Field(v, 0) = Val_int(a);
  return v;
}
static value f64_array(const double *src, int n) {
  double *d = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1));
  if (src) memcpy(d, src, sizeof(double) * (size_t)n);
  return mock_caml_ba_alloc_1d(d, n);
}
static value i32_array(const int32_t *src, int n) {
  int32_t *d = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
  if (src) memcpy(d, src, sizeof(int32_t) * (size_t)n);
  return mock_caml_ba_alloc_1d(d, n);
}
static value f32_array(const float *src, int n) {
  float *d = (float *)malloc(sizeof(float) * (size_t)(n ? n : 1));
  if (src) memcpy(d, src, sizeof(float) * (size_t)n);
  return mock_caml_ba_alloc_1d(d, n);
}

int main(int argc, char **argv) {
  const char *mode = argc > 1 ? argv[1] : "nogpu";
  if (setjmp(mock_caml_handler)) {
    printf("EXCEPTION %s\n", mock_caml_exn);
    return strcmp(mode, "nogpu") == 0 && strstr(mock_caml_exn, "no CUDA device") ? 0 : 3;
  }
  printf("devices %d\n", Int_val(ptb_ml_device_count(Val_unit)));
  value scene = ptb_ml_scene_create(Val_unit);

  /* [| Solid (0.8,0.3,0.3); Solid (0.9,0.9,0.9); Solid (0.2,0.2,0.2); Checker {width=10; height=20; even=1; odd=2} |] */
  value tex = caml_alloc(4, 0);
  Field(tex, 0) = boxed3(0, 0.8, 0.3, 0.3), Field(tex, 1) = boxed3(0, 0.9, 0.9, 0.9), Field(tex, 2) = boxed3(0, 0.2, 0.2, 0.2);
  Field(tex, 3) = ints4(1, 10, 20, 1, 2);
  ptb_ml_set_textures(scene, tex);
  /* [| Lambertian 3; Lambertian 0; Metal 1; Dielectric 1.5 |] */
  value mat = caml_alloc(4, 0);
  Field(mat, 0) = one_int(0, 3), Field(mat, 1) = one_int(0, 0), Field(mat, 2) = one_int(1, 1);
  value di = caml_alloc(1, 2);
  Field(di, 0) = caml_copy_double(1.5);
  Field(mat, 3) = di;
  ptb_ml_set_materials(scene, mat);

  const double xs[4] = {0.0, 0.0, -1.0, 1.0}, ys[4] = {-100.5, 0.0, 0.0, 0.0}, zs[4] = {-1.0, -1.2, -1.0, -1.0},
               rs[4] = {100.0, 0.5, 0.5, 0.5};
  const int32_t sm[4] = {0, 1, 3, 2};
  value a6[6] = {scene, f64_array(xs, 4), f64_array(ys, 4), f64_array(zs, 4), f64_array(rs, 4), i32_array(sm, 4)};
  ptb_ml_set_spheres_bc(a6, 6);
  /* Gradient_y ((1,1,1), (0.5,0.7,1.0)) */
  value bg = caml_alloc(2, 1);
  Field(bg, 0) = boxed3(0, 1.0, 1.0, 1.0), Field(bg, 1) = boxed3(0, 0.5, 0.7, 1.0);
  ptb_ml_set_background(scene, bg);

  /* extension stubs: a light quad is accepted and removed again (NULL through the C ABI, not reachable from OCaml:
     the scene is rebuilt instead), and the progress poll marshals a pair of ints */
  {
    value quad = caml_alloc(3, 0);
    Field(quad, 0) = boxed3(0, 0.0, 1.0, -1.0), Field(quad, 1) = boxed3(0, 0.5, 0.0, 0.0), Field(quad, 2) = boxed3(0, 0.0, 0.0, 0.5);
    ptb_ml_set_light_quad(scene, quad);
    ptb_scene_set_light_quad(*((ptb_scene **)Data_custom_val(scene)), NULL, NULL, NULL);
    value pr0 = ptb_ml_render_progress(Val_int(0));
    printf("progress %ld of %ld\n", (long)Long_val(Field(pr0, 0)), (long)Long_val(Field(pr0, 1)));
  }

  /* error path of the stubs: a material row that does not exist must surface as Failure, not crash */
  if (strcmp(mode, "badrow") == 0) {
    const int32_t bad[4] = {0, 1, 7, 2};
    value b6[6] = {scene, a6[1], a6[2], a6[3], a6[4], i32_array(bad, 4)};
    ptb_ml_set_spheres_bc(b6, 6);
    ptb_ml_commit(scene, Val_int(0));
    printf("NO EXCEPTION\n");
    return 4;
  }

  value ms = ptb_ml_commit(scene, Val_int(0)); /* raises Failure without a CUDA device */
  printf("commit %.3f ms\n", Double_val(ms));

  const int W = 64, H = 32, SPP = 16, MB = 6;
  value cam = caml_alloc(4, 0);
  Field(cam, 0) = caml_copy_double(-2.0), Field(cam, 1) = caml_copy_double(-1.0);
  Field(cam, 2) = caml_copy_double(4.0), Field(cam, 3) = caml_copy_double(2.0);
  value img = f64_array(NULL, 3 * W * H);
  value a8[8] = {scene, Val_int(W), Val_int(H), Val_int(SPP), Val_int(MB), cam, Val_int(0), img};
  value dev_ms = ptb_ml_render_bc(a8, 8);
  printf("render %.3f ms on the device\n", Double_val(dev_ms));

  /* batched intersection through the stubs: one ray at the middle sphere, one at the sky */
  const float o[6] = {0, 0, 0, 0, 0, 0}, d[6] = {0, 0, -1, 0, 1, 0};
  value t_hit = f32_array(NULL, 2), prim = i32_array(NULL, 2);
  value b8[8] = {scene, f32_array(o, 6), f32_array(d, 6), caml_copy_double(0.0), caml_copy_double(3.0e38), Val_int(0), t_hit, prim};
  ptb_ml_intersect_batch_bc(b8, 8);
  const float *th = (const float *)Caml_ba_data_val(t_hit);
  const int32_t *pr = (const int32_t *)Caml_ba_data_val(prim);
  printf("intersect t0 %.6f prim0 %d miss1 %d prim1 %d\n", th[0], pr[0], isnan(th[1]) ? 1 : 0, pr[1]);

  if (mock_caml_runtime_released != 0) {
    printf("runtime lock not re-acquired\n");
    return 5;
  }
  if (argc > 2) {
    FILE *f = fopen(argv[2], "wb");
    if (!f) return 6;
    fwrite(Caml_ba_data_val(img), sizeof(double), (size_t)3 * W * H, f);
    fclose(f);
  }
  mock_caml_finalize(scene); /* the GC's finaliser: ptb_scene_destroy */
  printf("OK\n");
  return 0;
}

(* ptb.ml — OCaml bindings to libptb200 (include/ptb200.h), the whole-render replacement of the per-leaf FFI
   `spheres_intersect_native` (shirley_spheres/bin/main.ml:162-172).

   STATUS: written against the OCaml 5 C API; this image has no ocaml / dune / opam, so this file is not
   type-checked.  The C stubs it binds (ptb_stubs.c) are compiled and run on a mock of the OCaml runtime with the
   value shapes these declarations imply (tests/test_ocaml_stubs.py), and the C ABI below them is the one the
   Python/ctypes tests and the C++ CLI twin exercise. *)

type scene (* custom block holding a ptb_scene* *)

type f64array = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type i32array = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array1.t

(* rows mirror Texture.t / Material.t (path_tracer/src/texture.ml:19-31, material.ml:3-9) *)
type texture_row =
  | Solid of float * float * float
  | Checker of { width : int; height : int; even : int; odd : int } (* rows of the two sub-textures *)

type material_row =
  | Lambertian of int (* texture row *)
  | Metal of int
  | Dielectric of float (* refractive index *)
  | Emissive of int (* extension: Material.emit = this texture row, scatter = Absorb (ptb200.h) *)

type background =
  | Constant of float * float * float
  | Gradient_y of (float * float * float) * (float * float * float) (* shirley main.ml:104-110 *)

external device_count : unit -> int = "ptb_ml_device_count"
external scene_create : unit -> scene = "ptb_ml_scene_create"
external set_textures : scene -> texture_row array -> unit = "ptb_ml_set_textures"
external set_materials : scene -> material_row array -> unit = "ptb_ml_set_materials"

(* coords SoA exactly as Simd_leaf.of_elts packs them (shirley main.ml:177-193) + a material row per sphere *)
external set_spheres : scene -> f64array -> f64array -> f64array -> f64array -> i32array -> unit
  = "ptb_ml_set_spheres_bc" "ptb_ml_set_spheres"

(* ganesha's Mesh SoA (ganesha/bin/main.ml:37-43): vertex columns, 3 indices per face, material per face *)
external set_triangles : scene -> f64array -> f64array -> f64array -> i32array -> i32array -> unit
  = "ptb_ml_set_triangles_bc" "ptb_ml_set_triangles"

external set_background : scene -> background -> unit = "ptb_ml_set_background"

(* extension: ~diffuse_plus_light = Mix (Diffuse, Quad_light {origin; u; v}) instead of Pdf.diffuse (render_command.ml:81) *)
external set_light_quad :
  scene -> (float * float * float) * (float * float * float) * (float * float * float) -> unit = "ptb_ml_set_light_quad"

(* Shape_tree.create + upload; returns milliseconds, printed like "build time" (shirley main.ml:264) *)
external commit : scene -> device:int -> float = "ptb_ml_commit"

(* Integrator.render into the Bimage f64 rgb buffer behind Image.data (render_command.ml:65).
   camera = (lower_left_x, lower_left_y, view_x, view_y) (camera.ml:50-53). *)
external render :
  scene -> width:int -> height:int -> spp:int -> max_bounces:int -> camera:float * float * float * float
  -> device:int -> f64array -> float (* device ms *) = "ptb_ml_render_bc" "ptb_ml_render"

(* `update_progress` (integrator.ml:130,150; render_command.ml:86-104) as a poll: (paths done, paths in total) of the
   render in flight on [device].  Call it from another domain while [render] runs — the library serialises calls
   that use the same device and never calls back into OCaml:
     let d = Domain.spawn (fun () -> Ptb.render scene ... image) in
     while not (finished ()) do let done_, total = Ptb.render_progress ~device:0 in report (done_ - !seen); ... done *)
external render_progress : device:int -> int * int = "ptb_ml_render_progress"

(* batched generalisation of spheres_intersect_native: t (nan = miss) and primitive index (-1 = miss) per ray *)
external intersect_batch :
  scene -> origins:(float, Bigarray.float32_elt, Bigarray.c_layout) Bigarray.Array1.t
  -> directions:(float, Bigarray.float32_elt, Bigarray.c_layout) Bigarray.Array1.t -> t_min:float -> t_max:float
  -> device:int -> t_hit:(float, Bigarray.float32_elt, Bigarray.c_layout) Bigarray.Array1.t -> prim:i32array -> unit
  = "ptb_ml_intersect_batch_bc" "ptb_ml_intersect_batch"

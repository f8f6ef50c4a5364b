// cli.cpp — C++ twin of the reference's scene binaries over libptb200's C ABI.
//
// The reference ships three executables built on Render_command.Args / Render_command.Make(Scene).run
// (render_command/src/render_command.ml:16-47,64-109; shirley_spheres/bin/main.ml:220-292).  OCaml cannot be
// built in this image, so this file is the host program that is actually exercised: same flags, same
// printed lines, plus the device flag the port adds.  One binary, dispatching on its name or first argument:
//
//   shirley_spheres --dimension=600,300 --samples-per-pixel=32 --max-ray-bounces=8 [-o out.png] [--no-simd]
//   cornell_box     -d 1024,1024 --samples-per-pixel=256 --max-ray-bounces=16 [--background white|sky|light]
//                   (light: closed black box lit by an emissive square, sampled through diffuse_plus_light — the
//                    extension behind BASELINE.json configs[1])
//   ganesha         -d 1920,1080 --samples-per-pixel=256 (--ganesha-ply FILE | --synthetic-faces N)
//   common: [--no-progress] [--device cuda[:N]] [--gpus N] [--f64]   (output *.ppm writes a PPM instead of a PNG)
//
// cornell_box and ganesha render through the path integrator here (the reference's own binaries use the
// progressive photon mapper, which is out of scope: SURVEY.md D1/D2).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ptb200.h"
#include "../../include/ptb200_scenes.h"

namespace {

struct Args {  // Render_command.Args.t (render_command.ml:6-14) + what the port adds
  int width = 0, height = 0, samples_per_pixel = 1, max_bounces = 8, device = 0, gpus = 1;
  int preview_every = 0;  // > 0: render in batches of this many sample passes and rewrite the output after each
  std::string output = "output.png";
  bool no_progress = false, no_simd = false, f64 = false;
  std::string ply, background = "white";
  long long synthetic_faces = 0;
};

[[noreturn]] void die(const std::string &m) {
  std::fprintf(stderr, "%s\n", m.c_str());
  std::exit(124);  // Cmdliner's exit code for command line parse errors
}
void check(int rc, const char *what) {
  if (rc) die(std::string(what) + ": " + ptb_last_error());
}

// ---- PNG (stored deflate blocks: no compression library needed) -----------------------------------
uint32_t crc_table[256];
void crc_init() {
  for (uint32_t n = 0; n < 256; ++n) {
    uint32_t c = n;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
}
uint32_t crc(const unsigned char *p, size_t n, uint32_t c = 0xffffffffu) {
  for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return c;
}
void put32(std::vector<unsigned char> &v, uint32_t x) {
  for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s));
}
void chunk(FILE *f, const char *type, const std::vector<unsigned char> &data) {
  std::vector<unsigned char> b;
  put32(b, (uint32_t)data.size());
  std::fwrite(b.data(), 1, 4, f);
  std::vector<unsigned char> td(type, type + 4);
  td.insert(td.end(), data.begin(), data.end());
  std::fwrite(td.data(), 1, td.size(), f);
  b.clear();
  put32(b, crc(td.data(), td.size()) ^ 0xffffffffu);
  std::fwrite(b.data(), 1, 4, f);
}
bool write_png(const std::string &path, const std::vector<unsigned char> &rgb, int w, int h) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  crc_init();
  const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  std::fwrite(sig, 1, 8, f);
  std::vector<unsigned char> ihdr;
  put32(ihdr, (uint32_t)w), put32(ihdr, (uint32_t)h);
  ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
  chunk(f, "IHDR", ihdr);
  std::vector<unsigned char> raw;
  raw.reserve((size_t)h * (3 * w + 1));
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);
    raw.insert(raw.end(), rgb.begin() + (size_t)y * 3 * w, rgb.begin() + (size_t)(y + 1) * 3 * w);
  }
  std::vector<unsigned char> z = {0x78, 0x01};
  uint32_t a = 1, b = 0;
  for (unsigned char c : raw) a = (a + c) % 65521u, b = (b + a) % 65521u;
  for (size_t pos = 0; pos < raw.size() || pos == 0;) {
    size_t n = std::min<size_t>(65535, raw.size() - pos);
    z.push_back(pos + n >= raw.size() ? 1 : 0);
    z.push_back((unsigned char)(n & 255)), z.push_back((unsigned char)(n >> 8));
    z.push_back((unsigned char)(~n & 255)), z.push_back((unsigned char)((~n >> 8) & 255));
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    pos += n;
    if (n == 0) break;
  }
  put32(z, (b << 16) | a);
  chunk(f, "IDAT", z);
  chunk(f, "IEND", {});
  return std::fclose(f) == 0;
}

// ---- Cmdliner-style parsing: --name=value, --name value, -d value ------------------------------------
Args parse(int argc, char **argv, int first) {
  Args a;
  bool have_dim = false;
  for (int i = first; i < argc; ++i) {
    std::string s = argv[i], val;
    bool has_val = false;
    size_t eq = s.find('=');
    if (s.rfind("--", 0) == 0 && eq != std::string::npos) val = s.substr(eq + 1), s = s.substr(0, eq), has_val = true;
    auto need = [&]() -> std::string {
      if (has_val) return val;
      if (i + 1 >= argc) die("option '" + s + "' needs an argument");
      return argv[++i];
    };
    if (s == "-d" || s == "--dimension") {
      std::string v = need();
      if (std::sscanf(v.c_str(), "%d,%d", &a.width, &a.height) != 2) die("option '--dimension': invalid value '" + v + "', expected WIDTH,HEIGHT");
      have_dim = true;
    } else if (s == "--samples-per-pixel") a.samples_per_pixel = std::atoi(need().c_str());
    else if (s == "--max-ray-bounces") a.max_bounces = std::atoi(need().c_str());
    else if (s == "-o" || s == "--output") a.output = need();
    else if (s == "--no-progress") a.no_progress = true;
    else if (s == "--no-simd") a.no_simd = true;  // Simd_leaf vs Array_leaf is a CPU-side choice; one device kernel
    else if (s == "--f64") a.f64 = true;
    else if (s == "--device") {
      std::string v = need();
      if (v == "cpu") die("--device=cpu: this backend has no CPU path (use the reference renderer)");
      size_t c = v.find(':');
      a.device = c == std::string::npos ? (v == "cuda" ? 0 : std::atoi(v.c_str())) : std::atoi(v.c_str() + c + 1);
    } else if (s == "--gpus") a.gpus = std::atoi(need().c_str());
    else if (s == "--ganesha-ply" || s == "-ganesha-ply") a.ply = need();
    else if (s == "--synthetic-faces") a.synthetic_faces = std::atoll(need().c_str());
    else if (s == "--background") a.background = need();
    else if (s == "--preview-every") a.preview_every = std::atoi(need().c_str());
    else die("unknown option '" + s + "'");
  }
  if (!have_dim) die("required option --dimension is missing");  // Arg.required (render_command.ml:21-25)
  return a;
}

}  // namespace

int main(int argc, char **argv) {
  using clk = std::chrono::steady_clock;
  std::string name = argv[0];
  size_t sl = name.find_last_of('/');
  if (sl != std::string::npos) name = name.substr(sl + 1);
  int first = 1;
  if (name != "shirley_spheres" && name != "cornell_box" && name != "ganesha") {
    if (argc < 2) die("usage: " + name + " (shirley_spheres|cornell_box|ganesha) --dimension=W,H [options]");
    name = argv[1], first = 2;
  }
  Args a = parse(argc, argv, first);
  if (ptb_device_count() == 0) die("no CUDA device: libptb200 has no CPU path");
  ptb_scene *sc = ptb_scene_create();
  double cam[20];
  const double aspect = (double)a.width / (double)a.height;
  if (name == "shirley_spheres") {
    check(ptb_scene_load_shirley(sc, aspect, 42, cam), "shirley scene");  // Random.init 42 (main.ml:251)
  } else if (name == "cornell_box") {
    const double white[3] = {1, 1, 1}, sky0[3] = {1, 1, 1}, sky1[3] = {0.5, 0.7, 1.0};
    const double radiance[3] = {32, 32, 32};
    if (a.background == "light") check(ptb_scene_load_cornell_lit(sc, aspect, radiance, cam), "cornell scene");
    else if (a.background == "sky") check(ptb_scene_load_cornell(sc, aspect, PTB_BG_GRADIENT_Y, sky0, sky1, cam), "cornell scene");
    else check(ptb_scene_load_cornell(sc, aspect, PTB_BG_CONSTANT, white, nullptr, cam), "cornell scene");
  } else if (name == "ganesha") {
    float *xyz = nullptr;
    int32_t *faces = nullptr;
    int64_t nv = 0, nf = 0;
    std::vector<float> sx;
    std::vector<int32_t> sf;
    if (!a.ply.empty()) {
      check(ptb_ply_read_mesh(a.ply.c_str(), &xyz, &nv, &faces, &nf), "ply");
    } else {
      if (a.synthetic_faces <= 0) die("ganesha: give --ganesha-ply FILE (pbrt-v3-scenes ganesha.ply) or --synthetic-faces N");
      check(ptb_mesh_synthetic(a.synthetic_faces, 0xB200u, nullptr, 0, nullptr, 0, &nv, &nf), "synthetic mesh");
      sx.resize(3 * nv), sf.resize(3 * nf);
      check(ptb_mesh_synthetic(a.synthetic_faces, 0xB200u, sx.data(), nv, sf.data(), nf, &nv, &nf), "synthetic mesh");
      xyz = sx.data(), faces = sf.data();
    }
    std::printf("#vertices = %lld\n#faces = %lld\n", (long long)nv, (long long)nf);
    check(ptb_scene_load_mesh(sc, xyz, nv, faces, nf, aspect, cam), "mesh scene");
    if (!a.ply.empty()) ptb_free(xyz), ptb_free(faces);
  } else {
    die("unknown scene '" + name + "'");
  }
  std::printf("dim = %d x %d;\n", a.width, a.height);  // shirley main.ml:254-255
  int64_t ns = 0, nvv = 0, nt = 0;
  int32_t nm = 0, ntex = 0;
  ptb_scene_counts(sc, &ns, &nvv, &nt, &nm, &ntex);
  if (ns) std::printf("#spheres = %lld\n", (long long)ns);
  if (nt) std::printf("#triangles = %lld\n", (long long)nt);
  double build_ms = 0;
  if (a.gpus > 1) check(ptb_scene_commit_multi(sc, a.gpus, &build_ms), "commit");
  else check(ptb_scene_commit(sc, a.device, &build_ms), "commit");
  int32_t ts[8];
  ptb_scene_tree_stats(sc, ts);
  std::printf("tree depth = %d\n", ts[1]);          // main.ml:263
  std::printf("build time = %.3f ms\n", build_ms);  // main.ml:264 (here: build + upload)
  ptb_params p;
  std::memset(&p, 0, sizeof p);
  p.width = a.width, p.height = a.height, p.samples_per_pixel = a.samples_per_pixel, p.max_bounces = a.max_bounces;
  p.lower_left_x = cam[0], p.lower_left_y = cam[1], p.view_x = cam[2], p.view_y = cam[3];
  p.tile_rank = 0, p.tile_world = 1, p.flags = a.f64 ? PTB_FLAG_F64 : 0, p.device = a.device;
  std::vector<double> image((size_t)3 * a.width * a.height);
  ptb_stats st;
  // Progress bar (render_command.ml:86-104: `Progress.with_reporter` fed by update_progress with tile areas): a thread
  // polls the library's path counter while the render call blocks this one; --no-progress switches it off
  std::atomic<bool> rendering{true};
  std::thread bar;
  if (!a.no_progress)
    bar = std::thread([&] {
      uint64_t done = 0, total = 0, shown = ~0ull;
      while (rendering.load()) {
        if (ptb_render_progress(a.device, &done, &total) == 0 && total && done != shown) {
          const int w = 40, fill = (int)((double)done / (double)total * w);
          std::fprintf(stderr, "\rRendering [%.*s%*s] %3d%%", fill, "########################################", w - fill, "",
                       (int)(100.0 * (double)done / (double)total));
          std::fflush(stderr);
          shown = done;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
      }
      if (shown != ~0ull) std::fprintf(stderr, "\rRendering [########################################] 100%%\n");
    });
  // Bimage_unix.Stb.write of an f64 image: 8-bit, truncating (pinned by the sky rows of the golden PNG,
  // tests/golden/shirley_png_facts.json), clamped to [0, 255]
  auto write_image = [&](const std::vector<double> &img) {
    std::vector<unsigned char> rgb(img.size());
    for (size_t i = 0; i < img.size(); ++i) {
      double v = img[i] * 255.0;
      rgb[i] = (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : (int)v);
    }
    bool ok;
    if (a.output.size() > 4 && a.output.substr(a.output.size() - 4) == ".ppm") {
      FILE *f = std::fopen(a.output.c_str(), "wb");
      ok = f != nullptr;
      if (ok) {
        std::fprintf(f, "P6\n%d %d\n255\n", a.width, a.height);
        std::fwrite(rgb.data(), 1, rgb.size(), f);
        ok = std::fclose(f) == 0;
      }
    } else {
      ok = write_png(a.output, rgb, a.width, a.height);
    }
    if (!ok) die("cannot write " + a.output);
  };
  auto t0 = clk::now();
  int rrc = 0;
  if (a.preview_every > 0 && a.preview_every < a.samples_per_pixel) {
    // Progressive output (what the reference's photon-map binary does per iteration, progressive_photon_map.ml:447-449):
    // batches of sample passes (ptb_params.pass_first / pass_count), the filtered sums added up on the host, the
    // output file rewritten with the image of the passes finished so far.  Same samples as the one-call render.
    std::vector<double> acc(image.size(), 0.0), part(image.size());
    ptb_stats total;
    std::memset(&total, 0, sizeof total);
    for (int first_pass = 0; first_pass < a.samples_per_pixel && rrc == 0; first_pass += a.preview_every) {
      ptb_params q = p;
      q.pass_first = first_pass, q.pass_count = std::min(a.preview_every, a.samples_per_pixel - first_pass);
      q.flags |= PTB_FLAG_RAW_SUMS;
      rrc = a.gpus > 1 ? ptb_render_multi(sc, &q, a.gpus, part.data(), &st) : ptb_render(sc, &q, part.data(), &st);
      if (rrc) break;
      total.paths += st.paths, total.rays += st.rays, total.kernel_launches += st.kernel_launches, total.ms_device += st.ms_device;
      const double inv = 1.0 / (double)(first_pass + q.pass_count);
      for (size_t i = 0; i < acc.size(); ++i) acc[i] += part[i], image[i] = std::sqrt(acc[i] * inv);  // integrator.ml:152-154
      write_image(image);
      if (!a.no_progress) std::fprintf(stderr, "\rpreview: %d of %d passes written to %s\n", first_pass + q.pass_count, a.samples_per_pixel, a.output.c_str());
    }
    st = total;
  } else {
    rrc = a.gpus > 1 ? ptb_render_multi(sc, &p, a.gpus, image.data(), &st) : ptb_render(sc, &p, image.data(), &st);
  }
  const double ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  rendering.store(false);
  if (bar.joinable()) bar.join();
  check(rrc, "render");
  if (!(a.preview_every > 0 && a.preview_every < a.samples_per_pixel)) write_image(image);
  std::printf("rendered in: %.3f ms\n", ms);  // render_command.ml:108
  std::printf("device: %.3f ms, %.1f Mpaths/s, %.1f Mrays/s, %llu kernel launches\n", st.ms_device,
              (double)st.paths / st.ms_device / 1e3, (double)st.rays / st.ms_device / 1e3,
              (unsigned long long)st.kernel_launches);
  ptb_scene_destroy(sc);
  return 0;
}

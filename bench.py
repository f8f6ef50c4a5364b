#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native path-tracing backend.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one full render of the workload.  Workload (config.workload): BASELINE.json configs[3],
`shirley_spheres 3840x2160, 1024 spp, 8 bounces` — the configuration the metric is quoted on; it fits
one GPU.  Metric: Mpaths/s = W*H*spp / t_render (Mrays/s is reported beside it), t_render = ray
generation -> traversal/shading -> (N>1: reduce-scatter of the per-pixel sums by row bands) -> filter+gamma resolve.
Scene generation and BVH build/upload are outside `value` (the reference prints them separately,
shirley_spheres/bin/main.ml:264) and inside `e2e`.

`--impl reference`: the reference renderer is OCaml 5 + Rust and cannot be built in this image, so this
arm times oracle/ (kind "port": the C++ float64 restatement of the OCaml-domains + AVX path, AVX2
intrinsics leaf kernel, N-1 worker threads + 1 stitcher like integrator.ml:137-151) on all host cores,
on a bounded sample (the first passes) of the same workload.  That arm is self-contained: it loads only
oracle/ (liboracle.so + the committed scene file oracle/scenes/shirley_spheres.npz), never the product.

N>1 (torchrun, one rank per GPU): tiles dealt round-robin; the only exchange is a reduce-scatter of the sums
by row bands + a one-row halo swap, every rank resolves its own band, the bands are gathered on rank 0
(path_tracer_ocaml_b200/distributed.py).  `ms_reduce` / `ms_resolve` / `ms_gather` time those phases.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP, MB = 3840, 2160, 1024, 8
WORKLOAD = f"shirley_spheres {W}x{H}, {SPP} spp, {MB} bounces (BASELINE.json configs[3])"
# algorithmic lane-ops per unit, counted from the reference's own formulas
C_BOX, C_SPH, C_TRI, C_HIT, C_FILM = 25.0, 34.0, 46.0, 200.0, 27.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def config_dict(world, spp):
    """The `config` of BOTH arms (the reference arm runs a bounded sample of it, described in cpu_baseline.sample)."""
    return {"workload": WORKLOAD if spp == SPP else f"DEV OVERRIDE spp={spp}: " + WORKLOAD,
            "sharding": f"reference tile list (Tile.split 1024 px), tile t -> rank t mod {world}",
            "l2": "no flush needed: each wavefront batch (up to 512 Mi paths) streams tens of GB of queue state (>> 126 MB L2)",
            "paths_per_step": W * H * spp}


def oracle_sample(spp, pass_limit, threads):
    """Times the oracle (CPU baseline, kind 'port') on the first `pass_limit` passes of the workload and
    returns throughput + the reference-faithful per-ray test counts that define the algorithmic work.
    Loads oracle/ only: the scene comes from oracle/scenes/shirley_spheres.npz."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    osc = O.OracleScene(O.load_scene_file("shirley_spheres"))
    params = O.make_params(O.shirley_camera(W / H), W, H, spp, MB)
    t0 = time.perf_counter()
    _, cn = osc.render(params, n_threads=threads, pass_limit=pass_limit)
    dt = time.perf_counter() - t0
    rays = max(cn.rays, 1)
    per_ray = {"box": cn.box_tests / rays, "sphere": cn.sphere_tests / rays, "tri": cn.tri_tests / rays,
               "rays_per_path": cn.rays / cn.paths, "hits_per_path": cn.hits / cn.paths}
    return {"seconds": dt, "paths": int(cn.paths), "rays": int(cn.rays), "mpaths": cn.paths / dt / 1e6,
            "mrays": cn.rays / dt / 1e6, "per_ray": per_ray, "tree": osc.tree_stats()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pass_limit = 4  # 33 M paths per step: a bounded sample of the 8.49 G-path workload
    vals = []
    for i in range(args.warmup + args.steps):
        r = oracle_sample(SPP, pass_limit, cores)
        log(f"[reference] step {i}: {r['seconds']:.2f}s {r['mpaths']:.2f} Mpaths/s")
        if i >= args.warmup:
            vals.append(r)
    sec = sum(v["seconds"] for v in vals)
    paths = sum(v["paths"] for v in vals)
    rays = sum(v["rays"] for v in vals)
    value = paths / sec / 1e6
    sample = (f"first {pass_limit} of {SPP} passes of the workload per step ({paths // len(vals)} paths), "
              f"throughput is spp-independent")
    line = {"impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "mrays_per_s": rays / sec / 1e6,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(vals),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus, SPP),
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample,
                             "what": "C++ float64 restatement of the OCaml-domains+AVX path (oracle/), not the OCaml binary"},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


def main():
    # exactly ONE line on stdout (the JSON): libraries print there too (NCCL's version banner), so everything
    # else goes to stderr and the JSON line is written to the saved descriptor at the end
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    globals()["_emit"] = lambda line: (real_stdout.write(line + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--spp", type=int, default=SPP, help="override spp (development only; invalidates the metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import path_tracer_ocaml_b200 as P
    from path_tracer_ocaml_b200 import capi
    from path_tracer_ocaml_b200.distributed import SharedHostImage, alloc_padded_sums, render_sharded_bands

    spp = args.spp
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}")
    if P.lib().ptb_device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; libptb200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    scene = P.shirley_spheres(W, H)
    integ = P.Integrator(scene, W, H, spp, MB, device=local, tile_rank=rank, tile_world=world)
    sums = alloc_padded_sums(H, W, world, dev)  # the first H rows are the frame
    launches = [0]
    acc = {"rays": 0, "ms_trace": 0.0, "ms_device": 0.0}
    phase_ms = {"reduce": 0.0, "resolve": 0.0, "gather": 0.0}
    out = {}

    def step(profile, gather_to=0):
        """One render: this rank's tiles -> sums; reduce-scatter by row bands + halo swap; every rank resolves its
        band; (gather_to = 0) the bands are gathered on rank 0's device."""
        sums.zero_()
        integ.render_device(sums, flags=capi.PTB_FLAG_PROFILE if profile else 0)
        launches[0] += integ.stats.kernel_launches + 1
        acc["rays"] += integ.stats.rays
        acc["ms_trace"] += integ.stats.ms_trace
        acc["ms_device"] += integ.stats.ms_device
        timers = {} if profile else None

        def resolve_rows(loc):
            launches[0] += 1
            return integ.resolve_rows_device(loc)

        out["band"], out["image"] = render_sharded_bands(sums, H, resolve_rows, gather_to=gather_to, timers=timers)
        return timers

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    # ---- timed region: exactly K steps, device events, barrier + synchronize on both sides ----------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    acc.update(rays=0, ms_trace=0.0, ms_device=0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    all_timers = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        all_timers.append(step(True))
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    for tm in all_timers:  # (events are read after the closing synchronize)
        for k, (a, z) in tm.items():
            phase_ms[k] += a.elapsed_time(z)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    stat = torch.tensor([float(acc["rays"]), float(launches[0])], device=dev, dtype=torch.float64)
    tr = torch.tensor([acc["ms_trace"]], device=dev, dtype=torch.float64)
    ph = torch.tensor([phase_ms["reduce"], phase_ms["resolve"], phase_ms["gather"], acc["ms_device"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stat, op=dist.ReduceOp.SUM)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    total_ms = ms.item()
    total_rays, total_launches, ms_trace = stat[0].item(), int(stat[1].item()), tr.item()
    timed = dict(acc)  # rank 0's counters of the timed region only (the e2e leg below keeps accumulating)
    paths_per_step = W * H * spp
    value = paths_per_step * args.steps / (total_ms * 1e-3) / 1e6
    mrays = total_rays / (total_ms * 1e-3) / 1e6

    # ---- e2e: through the public API with HOST buffers, wall clock, every step ---------------------------------
    # N=1: the reference-facing C-ABI calls: ptb_scene_commit (BVH build + host->device copy of the scene tables) +
    # ptb_render (render, resolve, device->host copy of the float64 image the reference's Bimage holds).
    # N>1: commit + sharded render + band exchange + per-rank resolve, then every rank copies ITS band of the image
    # into one page-locked host image shared by the ranks (each over its own PCIe link); the step ends when the whole
    # image is in host memory (barrier).
    t = scene.tables()
    h2d = int(t["n_spheres"] * 5 * 8 + t["n_materials"] * 16 + t["n_textures"] * 48 + 256)
    shared = None
    if world == 1:
        host_img = torch.empty(H, W, 3, dtype=torch.float64).pin_memory()
        host_np = host_img.numpy()
        d2h = int(W * H * 3 * 8)
        e2e_what = ("ptb_scene_commit (BVH build + table upload) + ptb_render into a pinned host float64 image "
                    "(render + resolve + device->host copy), wall clock")
    else:
        shared = SharedHostImage(H, W, world, rank)
        d2h = int(W * H * 3 * 4)
        e2e_what = ("scene commit (BVH build + table upload) + sharded render + reduce-scatter by row bands + per-rank "
                    "resolve + every rank's band copied into one shared page-locked host float32 image, wall clock "
                    f"(d2h_bytes_per_step is the whole image; each rank moves 1/{world} of it)")

    def e2e_step():
        tc = time.perf_counter()
        scene.commit(local)  # host->device copy of the step's inputs (the scene tables), incl. BVH build
        if world == 1:
            tr = time.perf_counter()
            integ.render(image=host_np)
            launches[0] += integ.stats.kernel_launches
            log(f"e2e step: commit {1e3 * (tr - tc):.1f} ms, ptb_render {1e3 * (time.perf_counter() - tr):.1f} ms "
                f"(inside the library {integ.stats.ms_total:.1f} ms: device {integ.stats.ms_device:.1f} ms, d2h {integ.stats.ms_d2h:.1f} ms)")
        else:
            step(False, gather_to=None)
            shared.put_band(out["band"])
            barrier()  # the image is complete in host memory when every rank's copy has landed

    n_launch_timed = launches[0]
    e2e_step()  # one untimed pass: the first host-buffer call grows the library's pooled device image buffers
    launches[0] = n_launch_timed
    barrier()
    # the interpreter's cyclic collector is kept out of the timed calls (a generation-2 pass over a process that has
    # imported torch takes tens of ms and would be charged to whichever ptb_render call it interrupts)
    import gc
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    e2e_val = paths_per_step * args.steps / e2e_s / 1e6
    if shared is not None:
        shared.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_trace) and CPU baseline ------------------------------------
    peak = C.c_double(0)
    capi.check(P.lib().ptb_fp32_peak(local, C.byref(peak), None))
    cpu = None
    per_ray = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        r = oracle_sample(spp, 4 if spp >= 4 else spp, cores)
        per_ray = r["per_ray"]
        cpu = {"value": r["mpaths"], "unit": "Mpaths/s", "mrays_per_s": r["mrays"], "cores": cores, "kind": "port",
               "sample": f"first 4 of {spp} passes of the workload ({r['paths']} paths, {r['seconds']:.1f} s)",
               "what": "C++ float64 restatement of the OCaml-domains+AVX path (oracle/), not the OCaml binary"}
    if per_ray is None:  # N>1 (no CPU leg): the committed oracle counts of the same config — deterministic numbers
        per_ray = json.load(open(os.path.join(ROOT, "profiles", "algorithmic_work_c4.json")))["per_ray"]
    a_trace = per_ray["box"] * C_BOX + per_ray["sphere"] * C_SPH + per_ray["tri"] * C_TRI
    c_gen = 40.0 + 5.0 * (2.0 + 2.0 * per_ray["hits_per_path"])
    a_ray = a_trace + C_HIT + (c_gen + C_FILM) / per_ray["rays_per_path"]
    rays_rank0 = timed["rays"]
    trace_tlops = rays_rank0 * a_trace / (max(timed["ms_trace"], 1e-9) * 1e-3) / 1e12  # rank 0's launches
    pipe_tlops = total_rays * a_ray / (total_ms * 1e-3) / 1e12 / world               # per GPU
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/): bytes per
    # ray there x the rays of an average launch here
    traffic = traffic_note = None
    try:
        tk = json.load(open(os.path.join(ROOT, "profiles", "trace_kernel_ncu.json")))
        batch = int(os.environ.get("PTB_BATCH", 1 << 28))                        # paths per wavefront batch
        n_trace_launches = args.steps * MB * -(-(paths_per_step // world) // batch)  # one k_trace per bounce per batch
        traffic = tk["dram_bytes_per_ray"] * rays_rank0 / n_trace_launches
        traffic_note = ("STATIC, not measured in this run: DRAM bytes per ray of the committed `ncu --set full` capture "
                        "(profiles/trace_kernel_ncu.json) x the rays of an average launch of this run. " + (tk.get("note") or ""))
    except Exception:
        pass
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    # algorithmic HBM bytes per ray: the scene is resident in shared memory, so only wavefront state moves —
    # 48 B entries: rays of bounce >= 1 are written by k_shade and read by k_trace, hits are written by k_trace
    # and read by k_shade; plus 12 B of film accumulation per path
    rpp, hpp = per_ray["rays_per_path"], per_ray["hits_per_path"]
    b_ray = (96.0 * (rpp - 1.0) + 96.0 * hpp + 12.0) / rpp
    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "mrays_per_s": mrays,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (R2 sample stream and cx,cy in f64, bit-exact)", "data": "synthetic",
        "config": config_dict(world, spp),
        "roofline": {"bound": "fp32_issue", "kernel": "k_trace<float,0>", "achieved": trace_tlops, "peak": peak.value,
                     "unit": "Tlane-op/s", "frac": trace_tlops / peak.value, "traffic": traffic,
                     "traffic_note": traffic_note,
                     "algorithmic_lane_ops_per_ray": a_trace,
                     "peak_source": "measured live: ptb_fp32_peak FFMA micro-benchmark on this GPU (MEASURED_PEAKS.json has no FP32 figure)",
                     "per_ray_counts": per_ray, "trace_ms_per_step": timed["ms_trace"] / args.steps,
                     "trace_share_of_step": timed["ms_trace"] / max(timed["ms_device"], 1e-9)},
        "roofline_pipeline": {"bound": "fp32_issue", "achieved": pipe_tlops, "peak": peak.value, "unit": "Tlane-op/s",
                              "frac": pipe_tlops / peak.value, "algorithmic_lane_ops_per_ray": a_ray,
                              "hbm": {"algorithmic_bytes_per_ray": b_ray,
                                      "achieved_gbs": total_rays * b_ray / (total_ms * 1e-3) / 1e9 / world,
                                      "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"}},
        # the second kernel: everything of the device time that is not k_trace is k_shade (k_resolve and the batch
        # bookkeeping are < 0.3 %); it moves 48 B in and 48 B out per scattered hit
        "roofline_shade": {"bound": "hbm", "kernel": "k_shade<float>",
                           "achieved": (rays_rank0 - paths_per_step * args.steps / world) * 96.0 /
                                       (max(timed["ms_device"] - timed["ms_trace"], 1e-9) * 1e-3) / 1e9,
                           "peak": hbm_peak, "unit": "GB/s",
                           "frac": (rays_rank0 - paths_per_step * args.steps / world) * 96.0 /
                                   (max(timed["ms_device"] - timed["ms_trace"], 1e-9) * 1e-3) / 1e9 / hbm_peak,
                           "algorithmic_bytes_per_hit": 96.0,
                           "ms_per_step": (timed["ms_device"] - timed["ms_trace"]) / args.steps,
                           "note": "scattered hits = rays of bounce >= 1 = rays - paths (rank 0)"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps, "what": e2e_what},
        "gpu_launches": total_launches,
        "clocks": clocks,
        # the exchange after the render, max over ranks, per step (device events on the launching stream): reduce =
        # reduce-scatter by row bands + halo rows, resolve = filter + gamma of the own band, gather = bands to rank 0;
        # render = the wavefront pipeline itself
        "exchange": {"ms_reduce": ph[0].item() / args.steps, "ms_resolve": ph[1].item() / args.steps,
                     "ms_gather": ph[2].item() / args.steps, "ms_render_max_rank": ph[3].item() / args.steps},
    }
    _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

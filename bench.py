#!/usr/bin/env python
"""bench.py - Mrays/s (primary + bounce + shadow) of the B200 path-tracing device on
BASELINE.json config 2 (1M-triangle displaced mesh, diffuse, 1080p, 256 spp), with
the reference's CPU Cycles timed beside it, plus sub-records for the other BASELINE
configs.

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm

A "step" is one pass of the hot path over one batch of synthetic input: the whole
frame of the workload at its spp through DeviceTask::RENDER (b200_render).

N > 1 (one process per GPU, launched by torch.distributed.run): the samples of the
frame are SPLIT over the ranks (sample split; rank r renders a contiguous
share of the sample indices) and the films are summed with one NCCL all-reduce per
step: total work is fixed, "scaling": "strong".  `--scaling weak` gives every rank the
config's full spp instead.

The JSON line (rank 0) carries
  value / ms_per_step / roofline / e2e / cpu_baseline   the headline, config 2
  configs: {cube, cornell, instanced}                   configs 1, 3, 4 at their own
                                                        size, each with value,
                                                        ms_per_step and its own roofline
                                                        (N = 1 only)
  config5: {...}                                        config 5: the instanced scene at
                                                        3840x2160, 1024 spp split over the
                                                        N ranks, ms_per_frame, the
                                                        all-reduce timed by its own events
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic_ncu.json")
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback

# algorithmic bytes per unit (SURVEY.md 8d / DESIGN.md)
RAY_IN, HIT_OUT_CLOSEST, HIT_OUT_SHADOW = 32, 16, 4
NODE_BYTES, PRIM_BYTES, INST_BYTES = 80, 48, 56


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="terrain",
                    choices=["terrain", "instanced", "cube", "cornell"])
    ap.add_argument("--materials", default=None,
                    help="scenes.cornell(materials=...) / scenes.default_cube(material=...)")
    ap.add_argument("--distribution", default="Multiscatter GGX",
                    help="Principled distribution of the cube / cornell workloads "
                         "(the node's default, or GGX)")
    ap.add_argument("--width", type=int, default=0, help="0 = the config's")
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0, help="samples per step (0 = the config's)")
    ap.add_argument("--cpu-spp", type=int, default=0,
                    help="samples of the bounded CPU sample (0 = auto)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong: the config's spp are split over the ranks; weak: every rank "
                         "renders the config's spp")
    ap.add_argument("--configs", default="auto",
                    help="sub-records: auto (N=1: cube,cornell,instanced,config5; N>1: config5), "
                         "none, or a comma list of cube,cornell,instanced,config5")
    ap.add_argument("--config-steps", type=int, default=2, help="timed steps of a sub-record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[],
                    help="device tunable name=value (b200_set_option), repeatable")
    return ap.parse_args()


def make_desc(workload, materials=None, width=0, height=0, spp=0,
              distribution="Multiscatter GGX"):
    from raytracingproject_b200 import scenes
    kw = {}
    if width:
        kw["width"] = width
    if height:
        kw["height"] = height
    if workload == "terrain":
        d = scenes.terrain(**kw)
    elif workload == "instanced":
        d = scenes.instanced(**kw)
    elif workload == "cube":
        # config 1: Blender's default material = a default Principled BSDF, whose
        # distribution is Multiscatter GGX (render/nodes.cpp:2728-2730)
        d = scenes.default_cube(material=materials or "principled",
                                distribution=distribution, **kw)
    else:
        d = scenes.cornell(materials=materials or "principled",
                           distribution=distribution, **kw)
    if spp:
        d.spp = spp
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def hbm_peak():
    try:
        with open(PEAKS_FILE) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per k_intersect_closest launch from the committed `ncu --set full`
    capture of this workload (profiles/traffic_ncu.json, written by
    profiles/summarize_ncu.py with the capture's provenance) - a PROFILER number taken in
    a separate run of the same build, not a measurement of this run; None when the
    workload has no capture."""
    try:
        rec = json.load(open(TRAFFIC_FILE)).get(workload)
        if rec:
            return rec.get("dram_bytes_per_launch"), rec
    except Exception:
        pass
    return None, None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_measure(rs, desc, cpu_spp, start_sample=0):
    """Time the reference CPUDevice (AVX2 kernels, all host threads) on a bounded
    sample of the workload and convert to Mrays/s with the reference's own ray
    census of exactly those samples."""
    _, sec = rs.render(start_sample, cpu_spp, tile_size=64)
    counts = rs.count_rays(start_sample, cpu_spp)
    rays = sum(counts)
    return {
        "value": rays / sec / 1e6, "unit": "Mrays/s", "cores": rs.num_threads(),
        "kind": "reference",
        "sample": "%dx%d, %d spp of the workload through the reference CPUDevice "
                  "(AVX2 kernel, BVH2, 64x64 tiles), rays from the reference's own census"
                  % (desc.width, desc.height, cpu_spp),
        "seconds": sec, "rays": {"camera": counts[0], "bounce": counts[1], "shadow": counts[2]},
        "spp_per_s": cpu_spp / sec,
    }


def auto_cpu_spp(rs, desc):
    """Pick a sample count that keeps the CPU leg near 10-20 s."""
    t0 = time.perf_counter()
    rs.render(0, 1, tile_size=64)
    one = max(time.perf_counter() - t0, 1e-3)
    return int(min(64, max(1, round(12.0 / one))))


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import cycles_ref
    desc = make_desc(args.workload, args.materials, args.width, args.height, args.spp,
                     args.distribution)
    rs = cycles_ref.build_scene(desc, kernel=cycles_ref.RefScene.AVX2)
    cpu_spp = args.cpu_spp or auto_cpu_spp(rs, desc)
    for _ in range(min(args.warmup, 1)):
        rs.render(0, 1, tile_size=64)
    secs = 0.0
    counts = rs.count_rays(0, cpu_spp)
    rays = sum(counts)
    for _ in range(args.steps):
        _, sec = rs.render(0, cpu_spp, tile_size=64)
        secs += sec
    value = rays * args.steps / secs / 1e6
    cpu = {"value": value, "unit": "Mrays/s", "cores": rs.num_threads(), "kind": "reference",
           "sample": "%dx%d, %d spp per step through the reference CPUDevice (AVX2 kernel, BVH2, "
                     "64x64 tiles, all host threads)" % (desc.width, desc.height, cpu_spp)}
    out = {
        "impl": "reference", "metric": "Mrays/s (primary+bounce+shadow)", "value": value,
        "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(desc), "width": desc.width, "height": desc.height,
                   "spp_per_step": cpu_spp, "triangles": desc.num_triangles,
                   "device": "CPU (reference Cycles, BVH2)"},
        "spp_per_s": cpu_spp * args.steps / secs,
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_name(desc):
    return "%s %dx%d %d spp (%s, %d tris)" % (desc.name, desc.width, desc.height, desc.spp,
                                              desc.notes, desc.num_triangles)


STAT_KEYS = ("primary_rays", "bounce_rays", "shadow_rays", "kernel_launches",
             "closest_launches", "shadow_launches", "closest_ms", "shadow_ms", "device_ms",
             "shade_ms", "batches", "iterations", "host_syncs", "host_waits")


def control_of(agg, steps):
    """Wavefront control of the timed steps (this rank): where the device time went and how
    often the host stopped the stream.  Phase times are CUDA-event sums; `other` is what is
    left of the call (init_from_camera, film, iteration roll-over, idle gaps)."""
    dev_ms = max(agg["device_ms"], 1e-9)
    return {
        "batches_per_step": agg["batches"] / steps, "iterations_per_step": agg["iterations"] / steps,
        "launches_per_step": agg["kernel_launches"] / steps,
        "stream_syncs_per_step": agg["host_syncs"] / steps,
        "lagged_counter_reads_per_step": agg["host_waits"] / steps,
        # block shape of the lean multiscatter / full shading kernel the device settled on for
        # this scene (-1 still probing, 0 = two blocks of 256 threads per SM, 1 = one of 512,
        # 2 = one of 1024)
        "shade_wide": agg.get("shade_wide", 0),
        "share": {"intersect_closest": agg["closest_ms"] / dev_ms,
                  "sort_and_shade": agg["shade_ms"] / dev_ms,
                  "intersect_shadow": agg["shadow_ms"] / dev_ms,
                  "other": 1.0 - (agg["closest_ms"] + agg["shade_ms"] + agg["shadow_ms"]) / dev_ms},
    }


class Arm:
    """One workload bound to this rank's B200 device: the scene flattened by the
    reference's own host code (Scene::device_update - its role in the reference), the
    arrays pushed through Device::mem_copy_to / const_copy_to, the BVH8 built."""

    def __init__(self, desc, local, stream, opts):
        from oracle import cycles_ref  # scene front-end (reference host code) + cpu_baseline leg
        from raytracingproject_b200.device import B200Device
        self.desc = desc
        t0 = time.perf_counter()
        self.rs = cycles_ref.build_scene(desc, kernel=cycles_ref.RefScene.AVX2)
        arrays = self.rs.device_arrays()
        self.t_scene = time.perf_counter() - t0
        self.ps = self.rs.pass_stride
        self.dev = B200Device(local)
        self.dev.set_stream(stream.cuda_stream)
        for o in opts:
            k, v = o.split("=")
            self.dev.set_option(k, int(v))
        t0 = time.perf_counter()
        self.dev.upload_scene(arrays)
        self.bvh = self.dev.build_bvh()
        self.t_upload = time.perf_counter() - t0
        self.film_numel = desc.width * desc.height * self.ps

    def close(self):
        self.dev.close()
        self.rs.close()


def timed_steps(arm, reducer, stream, steps, warmup, start_sample, my_spp, rank, world, local,
                sample_clocks=False):
    """W untimed + K timed frames, barrier + synchronize on both sides, CUDA events on the
    render stream, max over ranks.  Returns (ms_max, per-rank stat sums, clocks, reduce_ms,
    pool_bytes)."""
    import torch
    import torch.distributed as dist
    dev, w, h = arm.dev, arm.desc.width, arm.desc.height

    def step():
        film = reducer.begin_frame()
        dev.render_tile(film.data_ptr(), 0, 0, w, h, start_sample, my_spp, 0, w)
        st = dev.stats()
        reducer.end_frame()  # NCCL all-reduce over NVLink when world > 1, on a side stream
        return st

    mem0 = dev.mem_used()
    for _ in range(max(warmup, 0)):
        step()
    reducer.finish()
    pool_bytes = dev.mem_used() - mem0  # the path pool is allocated by the first render
    torch.cuda.synchronize()
    reducer.reduce_ms()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    agg = {k: 0 for k in STAT_KEYS}
    for _ in range(steps):
        st = step()
        for k in agg:
            agg[k] += st[k]
    agg["shade_wide"] = st.get("shade_wide", 0)  # the last step's (decided in the warm-up)
    reducer.finish()
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    reduce_ms = reducer.reduce_ms()
    if world > 1:
        t = torch.tensor([ms, reduce_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, reduce_ms = float(t[0].item()), float(t[1].item())
    return ms, agg, clocks, reduce_ms, pool_bytes


def all_sum(values, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return [float(v) for v in values]
    r = torch.tensor(list(values), dtype=torch.float64, device="cuda")
    dist.all_reduce(r)
    return [float(x) for x in r.tolist()]


def roofline_of(arm, film, agg, workload, spp):
    """Bytes-per-ray roofline of k_intersect_closest: per-ray node / triangle
    / instance averages from a counter build of the same kernel on the same rays, times the
    rays the timed launches traced, over their CUDA-event time."""
    dev, w, h = arm.dev, arm.desc.width, arm.desc.height
    dev.set_option("count_traversal", 1)
    count_spp = min(spp, 4)
    film.zero_()
    dev.render_tile(film.data_ptr(), 0, 0, w, h, 0, count_spp, 0, w)
    cs = dev.stats()
    dev.set_option("count_traversal", 0)
    n_closest = cs["primary_rays"] + cs["bounce_rays"]
    nodes_per_ray = cs["closest_nodes"] / max(n_closest, 1)
    tris_per_ray = cs["closest_tris"] / max(n_closest, 1)
    inst_per_ray = cs["closest_instances"] / max(n_closest, 1)
    bytes_per_ray = (RAY_IN + HIT_OUT_CLOSEST + nodes_per_ray * NODE_BYTES +
                     tris_per_ray * PRIM_BYTES + inst_per_ray * INST_BYTES)
    sh_nodes = cs["shadow_nodes"] / max(cs["shadow_rays"], 1)
    sh_tris = cs["shadow_tris"] / max(cs["shadow_rays"], 1)
    sh_inst = cs["shadow_instances"] / max(cs["shadow_rays"], 1)
    bytes_per_shadow_ray = (RAY_IN + HIT_OUT_SHADOW + sh_nodes * NODE_BYTES +
                            sh_tris * PRIM_BYTES + sh_inst * INST_BYTES)
    closest_rays = agg["primary_rays"] + agg["bounce_rays"]
    closest_s = agg["closest_ms"] * 1e-3
    achieved = closest_rays * bytes_per_ray / max(closest_s, 1e-12) / 1e9
    peak, peak_src = hbm_peak()
    traffic, traffic_rec = ncu_traffic(workload)
    return {
        "bound": "hbm", "kernel": "k_intersect_closest", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": traffic_rec, "peak_source": peak_src,
        "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
        "tris_per_ray": tris_per_ray, "instances_per_ray": inst_per_ray,
        "rays_per_launch": closest_rays / max(agg["closest_launches"], 1),
        "avg_launch_ms": agg["closest_ms"] / max(agg["closest_launches"], 1),
        "launches": agg["closest_launches"],
        "share_of_step": agg["closest_ms"] / max(agg["device_ms"], 1e-9),
        "grays_per_s": closest_rays / max(closest_s, 1e-12) / 1e9,
        "shadow": {"bytes_per_ray": bytes_per_shadow_ray, "nodes_per_ray": sh_nodes,
                   "tris_per_ray": sh_tris, "instances_per_ray": sh_inst,
                   "achieved": agg["shadow_rays"] * bytes_per_shadow_ray /
                   max(agg["shadow_ms"] * 1e-3, 1e-12) / 1e9,
                   "share_of_step": agg["shadow_ms"] / max(agg["device_ms"], 1e-9)},
    }


def sub_record(name, desc, args, stream, rank, world, local, scaling="strong", steps=None,
               warmup=3, with_roofline=True):
    """One BASELINE config as a sub-record of the JSON line."""
    from raytracingproject_b200 import multigpu
    arm = Arm(desc, local, stream, args.opt)
    try:
        spp = desc.spp
        if scaling == "strong":
            start_sample, my_spp = multigpu.strong_range(rank, world, spp)
            total_spp = spp
        else:
            start_sample, my_spp = multigpu.weak_range(rank, spp)
            total_spp = spp * world
        reducer = multigpu.FilmReducer(arm.film_numel, "cuda")
        steps = steps or args.config_steps
        ms, agg, _, reduce_ms, pool_bytes = timed_steps(
            arm, reducer, stream, steps, warmup, start_sample, my_spp, rank, world, local)
        rays_rank = agg["primary_rays"] + agg["bounce_rays"] + agg["shadow_rays"]
        rays_total, launches, prim, bnc, shd = all_sum(
            [rays_rank, agg["kernel_launches"], agg["primary_rays"], agg["bounce_rays"],
             agg["shadow_rays"]], world)
        rec = {
            "workload": workload_name(desc), "value": rays_total / (ms * 1e-3) / 1e6,
            "unit": "Mrays/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "n_gpus": world, "scaling": scaling, "spp_per_step": total_spp,
            "spp_per_rank": my_spp, "spp_per_s": total_spp * steps / (ms * 1e-3),
            "rays": {"primary": prim, "bounce": bnc, "shadow": shd},
            "gpu_launches": int(launches), "allreduce_ms": reduce_ms if world > 1 else None,
            "bvh8": arm.bvh, "host_bvh_s": host_bvh_seconds(arm),
            "path_pool_mb": pool_bytes / 1e6, "control": control_of(agg, steps),
        }
        if with_roofline and rank == 0:
            rec["roofline"] = roofline_of(arm, reducer.films[0], agg, name, spp)
        return rec, arm, reducer
    except Exception:
        arm.close()
        raise


def host_bvh_seconds(arm):
    """Everything the host does between meshes and a traversable BVH8: the reference's
    scene update (SAH BVH2 build + pack + the other managers; an upper bound of its BVH
    share) plus the BVH8 collapse and its upload."""
    return {"scene_update_s": arm.t_scene, "bvh8_collapse_s": arm.bvh["build_ms"] * 1e-3,
            "upload_and_bvh8_s": arm.t_upload,
            "total_s": arm.t_scene + arm.t_upload}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from oracle import cycles_ref
    from raytracingproject_b200 import multigpu
    from raytracingproject_b200.device import DeviceMemory

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    # ------------------------------------------------------------ headline
    desc = make_desc(args.workload, args.materials, args.width, args.height, args.spp,
                     args.distribution)
    spp = desc.spp
    w, h = desc.width, desc.height
    arm = Arm(desc, local, stream, args.opt)
    dev, rs, ps = arm.dev, arm.rs, arm.ps
    if args.scaling == "strong":
        start_sample, my_spp = multigpu.strong_range(rank, world, spp)
    else:
        start_sample, my_spp = multigpu.weak_range(rank, spp)
    total_spp = spp if args.scaling == "strong" else spp * world
    reducer = multigpu.FilmReducer(arm.film_numel, "cuda")
    film = reducer.films[0]

    ms_max, agg, clocks, reduce_ms, pool_bytes = timed_steps(
        arm, reducer, stream, args.steps, args.warmup, start_sample, my_spp, rank, world, local,
        sample_clocks=True)
    rays_rank = agg["primary_rays"] + agg["bounce_rays"] + agg["shadow_rays"]
    rays_total, launches_total = all_sum([rays_rank, agg["kernel_launches"]], world)
    launches_total = int(launches_total)
    value = rays_total / (ms_max * 1e-3) / 1e6
    spp_per_s = total_spp * args.steps / (ms_max * 1e-3)

    # ---- e2e: the Device call a host makes, host buffers, copies inside the timing.
    # Every rank uploads its RenderBuffers from pinned host memory, renders its share of the
    # samples, joins the film reduction and reads the film back; wall clock between
    # barriers, max over ranks. ----
    e2e_all = None
    host_film = None
    if not args.no_e2e:
        # the step's input: an empty film in pinned host memory (never written, so it needs
        # no clearing between steps); the step's result comes back into `host_film`
        empty_film = torch.zeros(h * w * ps, dtype=torch.float32).pin_memory()
        host_film = torch.zeros(h * w * ps, dtype=torch.float32).pin_memory()
        mem = DeviceMemory("RenderBuffers", host_film.numpy())
        dev.mem_alloc(mem)
        nbytes = int(host_film.numel() * 4)
        e2e_steps = max(1, min(args.steps, 3))

        def upload_empty_film():  # mem_copy_to with the input buffer as the source
            dev._check(dev._L.b200_h2d(dev._ctx, mem.device_pointer, empty_film.data_ptr(), 0,
                                       nbytes), "mem_copy_to(RenderBuffers)")

        upload_empty_film()
        dev.render_tile(mem.device_pointer, 0, 0, w, h, start_sample, my_spp, 0, w)  # warm
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e_rays = 0
        for _ in range(e2e_steps):
            upload_empty_film()           # H2D of the step's film (RenderBuffers), every rank
            dev.render_tile(mem.device_pointer, 0, 0, w, h, start_sample, my_spp, 0, w)
            s_ = dev.stats()
            e_rays += s_["primary_rays"] + s_["bounce_rays"] + s_["shadow_rays"]
            if world > 1:                 # the sample split's one exchange: sum onto rank 0
                multigpu.reduce_film(mem_as_tensor(mem, film), dst=0)
            if rank == 0:
                dev.mem_copy_from(mem)    # D2H of the result: the one film of the job
        torch.cuda.synchronize()
        e_sec = time.perf_counter() - t0
        dev.mem_free(mem)
        if world > 1:
            t = torch.tensor([e_sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_sec = float(t.item())
        e_rays = all_sum([e_rays], world)[0]
        e2e_all = {"value": e_rays / e_sec / 1e6, "unit": "Mrays/s",
                   "h2d_bytes_per_step": nbytes * world,
                   "d2h_bytes_per_step": nbytes,
                   "ms_per_step": 1e3 * e_sec / e2e_steps, "steps": e2e_steps,
                   "api": "B200Device.mem_copy_to (an empty film, every rank) / render_tile "
                          "(DeviceTask::RENDER) / film sum onto rank 0 / mem_copy_from (rank 0) "
                          "over the C ABI"}

    out = None
    if rank == 0:
        roofline = roofline_of(arm, film, agg, args.workload, spp)

        # ---- e2e (measured on every rank above); rank 0 adds the reference-driven flow ----
        e2e = e2e_all
        if e2e is not None and world == 1:
            e2e_steps = e2e["steps"]
            # the same step through the C++ `B200Device : ccl::Device`, driven by the
            # reference's own Scene::device_update + DeviceTask::RENDER + RenderBuffers
            # readback (mem_zero on the device, D2H of the film)
            try:
                from raytracingproject_b200.device import B200HostDevice
                host = B200HostDevice(local)
                t_host = time.perf_counter()
                rs_gpu = cycles_ref.build_scene(desc, external_device=host.ptr)
                t_host = time.perf_counter() - t_host
                rs_gpu.render(0, spp, tile_size=0)
                bvh_dev = host.bvh_info()
                _, pack_s, _ = host.host_bvh8_report()
                t0 = time.perf_counter()
                p_rays = 0
                for _ in range(e2e_steps):
                    rs_gpu.render(0, spp, tile_size=0)
                    s = host.stats()
                    p_rays += s["primary_rays"] + s["bounce_rays"] + s["shadow_rays"]
                p_sec = time.perf_counter() - t0
                e2e["reference_flow"] = {
                    "value": p_rays / p_sec / 1e6, "unit": "Mrays/s",
                    "ms_per_step": 1e3 * p_sec / e2e_steps, "h2d_bytes_per_step": 64,
                    "d2h_bytes_per_step": int(host_film.numel() * 4),
                    "api": "reference Scene + DeviceTask::RENDER -> C++ B200Device -> C ABI",
                    # BVH as a host layout of the reference: Scene::device_update builds the
                    # SAH binary tree, BVH8::pack_nodes (csrc/bvh8_host.cpp) collapses it and
                    # hands the device ITS arrays - the whole host BVH cost is inside
                    # scene_update_s, nothing is built on the device (host_packed = 1)
                    "host_bvh": {"layout": "BVH_LAYOUT_BVH8" if bvh_dev["host_packed"]
                                 else "BVH_LAYOUT_BVH2 + device-side collapse",
                                 "scene_update_s": t_host,
                                 "bvh8_pack_nodes_s": pack_s if bvh_dev["host_packed"] else None,
                                 "device_build_ms": bvh_dev["build_ms"],
                                 "nodes": bvh_dev["num_nodes"]}}
                rs_gpu.close()
                host.close()
            except Exception as exc:  # the shim needs the reference headers to be built
                e2e["reference_flow"] = {"unavailable": str(exc)[:200]}

        # ---- CPU baseline on the box's host cores (bounded sample) ----
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu_spp = args.cpu_spp or auto_cpu_spp(rs, desc)
            cpu = cpu_measure(rs, desc, cpu_spp)

        out = {
            "metric": "Mrays/s (primary+bounce+shadow)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(desc), "width": w, "height": h,
                "spp_per_step": total_spp, "spp_per_rank": my_spp,
                "triangles": desc.num_triangles,
                "parallelism": "sample-split x%d (%s)" % (world, args.scaling),
                "l2": "inputs larger than L2: %.0f MB of BVH8 + %.0f MB of path state per batch"
                      % ((arm.bvh["node_bytes"] + arm.bvh["tri_bytes"]) / 1e6, pool_bytes / 1e6),
                "bvh8": arm.bvh, "host_bvh_s": host_bvh_seconds(arm),
            },
            "spp_per_s": spp_per_s,
            "rays": {"primary": agg["primary_rays"], "bounce": agg["bounce_rays"],
                     "shadow": agg["shadow_rays"], "per_rank_per_run": rays_rank},
            "gpu_launches": launches_total,
            "control": control_of(agg, args.steps),
            "allreduce_ms": reduce_ms if world > 1 else None,
            "clocks": clocks,
            "roofline": roofline,
            "e2e": e2e,
            "cpu_baseline": cpu,
        }
    arm.close()
    del reducer, film
    torch.cuda.empty_cache()

    # --------------------------------------------- the other BASELINE configs
    wanted = args.configs
    if wanted == "auto":
        wanted = "cube,cornell,instanced,config5" if world == 1 else "config5"
    wanted = [] if wanted == "none" else [x for x in wanted.split(",") if x]
    configs = {}
    config5 = None
    for name in wanted:
        try:
            if name == "config5":
                config5 = run_config5(args, stream, rank, world, local)
                continue
            d = make_desc(name)
            rec, a, red = sub_record(name, d, args, stream, rank, world, local)
            a.close()
            del red
            torch.cuda.empty_cache()
            configs[name] = rec
        except Exception as exc:  # a sub-record never takes the headline down with it
            configs[name] = {"error": str(exc)[:300]}
    if rank == 0:
        if configs:
            out["configs"] = configs
        if config5 is not None:
            out["config5"] = config5
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def run_config5(args, stream, rank, world, local):
    """BASELINE config 5: the instanced scene at 3840x2160, 1024 spp split as
    [k*1024/N, (k+1)*1024/N) over the N ranks, films summed with one NCCL all-reduce per
    frame (timed by its own events on the reduce stream).  At N > 1 rank 0 then renders the
    whole 1024 spp alone once, on the same GPU in the same job: the N = 1 time the claimed
    speed-up is taken against."""
    import torch
    import torch.distributed as dist
    from raytracingproject_b200 import multigpu
    d = make_desc("instanced", width=3840, height=2160, spp=1024)
    w, h = d.width, d.height
    steps = 1 if world == 1 else 2
    arm = Arm(d, local, stream, args.opt)
    try:
        start_sample, my_spp = multigpu.strong_range(rank, world, d.spp)
        reducer = multigpu.FilmReducer(arm.film_numel, "cuda")
        # warm-up: a 16 s frame is not repeated three times - 64 spp of it page the kernels
        # in and size the path pool, then one untimed all-reduce
        film = reducer.begin_frame()
        arm.dev.render_tile(film.data_ptr(), 0, 0, w, h, start_sample, min(my_spp, 64), 0, w)
        reducer.end_frame()
        reducer.finish()
        torch.cuda.synchronize()
        ms, agg, _, reduce_ms, pool_bytes = timed_steps(
            arm, reducer, stream, steps, 0, start_sample, my_spp, rank, world, local)
        rays_rank = agg["primary_rays"] + agg["bounce_rays"] + agg["shadow_rays"]
        rays_total, launches = all_sum([rays_rank, agg["kernel_launches"]], world)
        rec = {
            "workload": workload_name(d), "value": rays_total / (ms * 1e-3) / 1e6,
            "unit": "Mrays/s", "ms_per_frame": ms / steps, "steps": steps,
            "warmup": "64 spp + one all-reduce, untimed", "n_gpus": world, "scaling": "strong",
            "spp_per_frame": d.spp, "spp_per_rank": my_spp,
            "spp_per_s": d.spp * steps / (ms * 1e-3), "gpu_launches": int(launches),
            "allreduce_ms": reduce_ms if world > 1 else None,
            "film_bytes": arm.film_numel * 4, "path_pool_mb": pool_bytes / 1e6,
        }
        if world > 1:
            n1_ms = None
            if rank == 0:
                f = reducer.films[0]
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                f.zero_()
                e0.record(stream)
                arm.dev.render_tile(f.data_ptr(), 0, 0, w, h, 0, d.spp, 0, w)
                e1.record(stream)
                torch.cuda.synchronize()
                n1_ms = e0.elapsed_time(e1)
            dist.barrier()
            if rank == 0:
                rec["n1_ms_per_frame_same_job"] = n1_ms
                rec["speedup_vs_n1_claimed"] = n1_ms / rec["ms_per_frame"]
        return rec
    finally:
        arm.close()


_REAL_STDOUT = None


class _RawCuda:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 2}


def mem_as_tensor(mem, like):
    """View a C-ABI device allocation as a torch tensor (for the NCCL film reduce)."""
    import torch
    return torch.as_tensor(_RawCuda(mem.device_pointer, like.numel()), device=like.device)


def emit(obj):
    """The ONE JSON line, on the real stdout (libraries such as NCCL print their
    banners on fd 1; those are diverted to stderr in main())."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - Mrays/s (primary + bounce + shadow) of the B200 path-tracing device on
BASELINE.json config 2 (1M-triangle displaced mesh, diffuse, 1080p, 256 spp), with
the reference's CPU Cycles timed beside it.

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm

A "step" is one pass of the hot path over one batch of synthetic input: the whole
1080p frame at the workload's spp through DeviceTask::RENDER (b200_render).
N > 1 (one process per GPU, launched by torch.distributed.run): rank r renders the
disjoint sample range [r*spp, (r+1)*spp) and the films are summed with one NCCL
all-reduce per step (sample split) - per-GPU work is fixed: weak
scaling.  Prints ONE JSON line on rank 0.
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
    ap.add_argument("--materials", default="diffuse",
                    help="cornell only: scenes.cornell(materials=...), e.g. principled, textured")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=0, help="samples per step (0 = the config's)")
    ap.add_argument("--cpu-spp", type=int, default=0,
                    help="samples of the bounded CPU sample (0 = auto)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank renders the config's spp; strong: the spp are split")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[],
                    help="device tunable name=value (b200_set_option), repeatable")
    return ap.parse_args()


def make_desc(args):
    from raytracingproject_b200 import scenes
    kw = dict(width=args.width, height=args.height)
    if args.workload == "terrain":
        d = scenes.terrain(**kw)
    elif args.workload == "instanced":
        d = scenes.instanced(**kw)
    elif args.workload == "cube":
        d = scenes.default_cube(material="diffuse", **kw)
    else:
        d = scenes.cornell(materials=args.materials, **kw)
    if args.spp:
        d.spp = args.spp
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
    desc = make_desc(args)
    rs = cycles_ref.build_scene(desc, kernel=cycles_ref.RefScene.AVX2)
    cpu_spp = args.cpu_spp or auto_cpu_spp(rs, desc)
    for _ in range(min(args.warmup, 1)):
        rs.render(0, 1, tile_size=64)
    vals, secs = [], 0.0
    counts = rs.count_rays(0, cpu_spp)
    rays = sum(counts)
    for _ in range(args.steps):
        _, sec = rs.render(0, cpu_spp, tile_size=64)
        secs += sec
        vals.append(rays / sec / 1e6)
    value = rays * args.steps / secs / 1e6
    cpu = {"value": value, "unit": "Mrays/s", "cores": rs.num_threads(), "kind": "reference",
           "sample": "%dx%d, %d spp per step through the reference CPUDevice (AVX2 kernel, BVH2, "
                     "64x64 tiles, all host threads)" % (desc.width, desc.height, cpu_spp)}
    out = {
        "impl": "reference", "metric": "Mrays/s (primary+bounce+shadow)", "value": value,
        "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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


def run_b200(args):
    import torch
    import torch.distributed as dist
    from oracle import cycles_ref  # scene front-end (reference host code) + cpu_baseline leg
    from raytracingproject_b200 import multigpu
    from raytracingproject_b200.device import B200Device, DeviceMemory

    rank, world, local = dist_env()
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)

    desc = make_desc(args)
    spp = desc.spp
    w, h = desc.width, desc.height

    # Scene flattening by the reference's own host code (Scene::device_update), then the
    # arrays go through Device::mem_copy_to / const_copy_to of the B200 device.
    t0 = time.perf_counter()
    rs = cycles_ref.build_scene(desc, kernel=cycles_ref.RefScene.AVX2)
    arrays = rs.device_arrays()
    t_scene = time.perf_counter() - t0
    ps = rs.pass_stride

    dev = B200Device(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    for o in args.opt:
        k, v = o.split("=")
        dev.set_option(k, int(v))
    t0 = time.perf_counter()
    dev.upload_scene(arrays)
    bvh = dev.build_bvh()
    t_upload = time.perf_counter() - t0

    film = torch.zeros(h * w * ps, dtype=torch.float32, device="cuda")
    if args.scaling == "strong":
        start_sample, my_spp = multigpu.strong_range(rank, world, spp)
    else:
        start_sample, my_spp = multigpu.weak_range(rank, spp)
    total_spp = spp if args.scaling == "strong" else spp * world

    def step():
        film.zero_()
        dev.render_tile(film.data_ptr(), 0, 0, w, h, start_sample, my_spp, 0, w)
        multigpu.reduce_film(film)  # NCCL all-reduce over NVLink when world > 1
        return dev.stats()

    mem0 = dev.mem_used()
    for _ in range(max(args.warmup, 0)):
        step()
    pool_bytes = dev.mem_used() - mem0   # the path pool is allocated by the first render

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    agg = {"primary_rays": 0, "bounce_rays": 0, "shadow_rays": 0, "kernel_launches": 0,
           "closest_launches": 0, "shadow_launches": 0, "closest_ms": 0.0, "shadow_ms": 0.0,
           "device_ms": 0.0}
    for _ in range(args.steps):
        st = step()
        for k in agg:
            agg[k] += st[k]
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    rays_rank = agg["primary_rays"] + agg["bounce_rays"] + agg["shadow_rays"]

    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        r = torch.tensor([rays_rank, agg["kernel_launches"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(r)
        rays_total, launches_total = float(r[0].item()), int(r[1].item())
    else:
        ms_max, rays_total, launches_total = ms, float(rays_rank), agg["kernel_launches"]

    value = rays_total / (ms_max * 1e-3) / 1e6
    spp_per_s = total_spp * args.steps / (ms_max * 1e-3)

    # ---- e2e: the Device call a host makes, host buffers, copies inside the timing.
    # Every rank uploads its RenderBuffers from pinned host memory, renders its share of the
    # samples, joins the film reduction and reads the film back; wall clock between
    # barriers, max over ranks. ----
    e2e_all = None
    if not args.no_e2e:
        host_film = torch.zeros(h * w * ps, dtype=torch.float32).pin_memory()
        mem = DeviceMemory("RenderBuffers", host_film.numpy())
        dev.mem_alloc(mem)
        e2e_steps = max(1, min(args.steps, 3))
        dev.mem_copy_to(mem)
        dev.render_tile(mem.device_pointer, 0, 0, w, h, start_sample, my_spp, 0, w)  # warm
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e_rays = 0
        for _ in range(e2e_steps):
            host_film.zero_()
            dev.mem_copy_to(mem)          # H2D of the step's film (RenderBuffers)
            dev.render_tile(mem.device_pointer, 0, 0, w, h, start_sample, my_spp, 0, w)
            s_ = dev.stats()
            e_rays += s_["primary_rays"] + s_["bounce_rays"] + s_["shadow_rays"]
            if world > 1:
                multigpu.reduce_film(mem_as_tensor(mem, film))
            dev.mem_copy_from(mem)        # D2H of the result
        torch.cuda.synchronize()
        e_sec = time.perf_counter() - t0
        dev.mem_free(mem)
        if world > 1:
            t = torch.tensor([e_sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_sec = float(t.item())
            r = torch.tensor([e_rays], dtype=torch.float64, device="cuda")
            dist.all_reduce(r)
            e_rays = float(r.item())
        e2e_all = {"value": e_rays / e_sec / 1e6, "unit": "Mrays/s",
                   "h2d_bytes_per_step": int(host_film.numel() * 4) * world,
                   "d2h_bytes_per_step": int(host_film.numel() * 4) * world,
                   "ms_per_step": 1e3 * e_sec / e2e_steps, "steps": e2e_steps,
                   "api": "B200Device.mem_copy_to / render_tile (DeviceTask::RENDER) / "
                          "mem_copy_from over the C ABI, every rank"}

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (intersect_closest) ----
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
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic_intersect_closest.json")
        if os.path.exists(tfile):
            try:
                traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "bound": "hbm", "kernel": "k_intersect_closest", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
            "tris_per_ray": tris_per_ray, "instances_per_ray": inst_per_ray,
            "rays_per_launch": closest_rays / max(agg["closest_launches"], 1),
            "avg_launch_ms": agg["closest_ms"] / max(agg["closest_launches"], 1),
            "launches": agg["closest_launches"],
            "share_of_step": agg["closest_ms"] / max(agg["device_ms"], 1e-9),
            "grays_per_s": closest_rays / max(closest_s, 1e-12) / 1e9,
            "shadow": {"bytes_per_ray": bytes_per_shadow_ray, "nodes_per_ray": sh_nodes,
                       "tris_per_ray": sh_tris,
                       "achieved": agg["shadow_rays"] * bytes_per_shadow_ray /
                       max(agg["shadow_ms"] * 1e-3, 1e-12) / 1e9,
                       "share_of_step": agg["shadow_ms"] / max(agg["device_ms"], 1e-9)},
        }

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
                rs_gpu = cycles_ref.build_scene(desc, external_device=host.ptr)
                rs_gpu.render(0, spp, tile_size=0)
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
                    "api": "reference Scene + DeviceTask::RENDER -> C++ B200Device -> C ABI"}
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
                "workload": workload_name(desc), "width": w, "height": h, "spp_per_step": total_spp,
                "spp_per_rank": my_spp, "triangles": desc.num_triangles, "parallelism": "sample-split x%d" % world,
                "l2": "inputs larger than L2: %.0f MB of BVH8 + %.0f MB of path state per batch"
                      % ((bvh["node_bytes"] + bvh["tri_bytes"]) / 1e6, pool_bytes / 1e6),
                "bvh8": bvh, "scene_build_s": t_scene, "upload_and_bvh8_s": t_upload,
            },
            "spp_per_s": spp_per_s,
            "rays": {"primary": agg["primary_rays"], "bounce": agg["bounce_rays"],
                     "shadow": agg["shadow_rays"], "per_rank_per_run": rays_rank},
            "gpu_launches": launches_total,
            "clocks": clocks,
            "roofline": roofline,
            "e2e": e2e,
            "cpu_baseline": cpu,
        }
    dev.close()
    rs.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


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

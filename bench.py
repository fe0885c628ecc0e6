#!/usr/bin/env python
"""bench.py -- path-segments/s and ms/frame of the hot path on N B200s of one node.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the RTOW
book-1 random-spheres scene, every sphere an instance of a tessellated unit sphere with
the reference's subdivision mix 9/6/3/8/6/3 (optx/rtwo.cxx:138-234: ~8.8 M instanced
triangles), 1200x800, 500 spp, depth 50, defocus blur on.  One "step" = one frame.

  value      whole-job path-segments/s: segments of a frame / device time of a frame, scene
             resident in HBM; at N>1 the frame's samples are split over the ranks (strong
             scaling) and the fixed-point accumulation buffers are summed by one NCCL reduce
             inside the timed region.
  e2e        the same through the C ABI a host program calls per frame (rtx_render* +
             rtx_postproc + rtx_read of the 8-bit image into host memory).
  roofline   dominant kernel k_render: algorithmic bytes per segment of SURVEY.md 8(d)
             (64*ceil(log2 N) + 48 + 128) x segments per launch / its CUDA-event duration.
  roofline.counted  work per segment from the kernel's own node-visit / primitive-test counters
             (instrumented build librtx_count.so, one short frame in a child process, untimed).
  cpu_baseline  the oracle's double-precision port of rtow.cxx (the reference's CPU path:
             analytic spheres, exhaustive scan) on a bounded sample, all host threads.

`--impl reference` times that CPU implementation alone, on the same image and scene.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "path-segments/s"
UNIT = "segments/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=1200)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--spp", type=int, default=500)
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--mode", default="mesh", choices=["mesh", "analytic"])
    ap.add_argument("--ndiv", type=int, default=None, help="one subdivision count for every sphere (default: reference mix)")
    ap.add_argument("--seed", type=int, default=4711)
    ap.add_argument("--scene", default="book1", help="book1 (default) or grid<N>: N x N field of small spheres (stress scene, BASELINE configs[4])")
    ap.add_argument("--cpu-spp", type=int, default=2, help="samples per pixel of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-count", action="store_true", help="skip the counted-work leg (instrumented library)")
    ap.add_argument("--count-spp", type=int, default=4, help="samples per pixel of the counted-work frame")
    ap.add_argument("--count-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (long configurations: the e2e fields then repeat the device-timed ones and say so)")
    ap.add_argument("--no-rtow", action="store_true", help="skip the run of the unmodified reference binary oracle/_ref/rtow (~80 s on one core)")
    return ap.parse_args()


def workload_name(a):
    mix = "tessellated 9/6/3/8/6/3" if a.scene == "book1" else "tessellated 9 (ground) / 6"
    geo = "analytic spheres" if a.mode == "analytic" else ("tessellated ndiv %d" % a.ndiv if a.ndiv is not None else mix)
    name = "RTOW book-1 random spheres" if a.scene == "book1" else "%s field of spheres" % a.scene
    return "%s, %s, %dx%d, %d spp, depth %d, defocus on" % (name, geo, a.width, a.height, a.spp, a.depth)


def make_scene(a):
    from rtxplay_b200 import scenes
    if a.scene == "book1":
        return scenes.book1(seed=1)
    if a.scene.startswith("grid"):
        return scenes.grid_field(int(a.scene[4:]), seed=2, ndiv=6 if a.ndiv is None else a.ndiv)
    raise SystemExit("unknown --scene " + a.scene)


# ------------------------------------------------------------------------- CPU legs
def cpu_leg(a, spp, steps=1, warmup=0):
    """The reference's CPU path (oracle<double> port of rtow.cxx, analytic spheres, exhaustive
    scan) on `spp` samples per pixel of the same image; returns (segments/s, cores, description)."""
    import oracle as orc
    from rtxplay_b200 import scenes
    spheres = make_scene(a)
    tab, _ = scenes.table(spheres, "analytic")
    cam = orc.camera_f32((13, 2, 3), (0, 0, 0), (0, 1, 0), 20., a.width / a.height, .1, 10.)
    cores = os.cpu_count() or 1
    times, segs = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = orc.render(orc.F64_PCG, tab, cam, a.width, a.height, spp, a.depth, seed=a.seed, sample0=it * spp, threads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            segs += int(out["rpp"].sum())
    total = sum(times)
    desc = "%dx%d, %d of %d spp per step, analytic spheres (rtow.cxx geometry), exhaustive scan, %d threads" % (
        a.width, a.height, spp, a.spp, cores)
    return segs / total, cores, desc, 1e3 * total / max(steps, 1)


def cpu_worker(a):
    """Child process of the GPU arm: the oracle is loaded here, never in the process that drives the GPU."""
    v, cores, desc, ms = cpu_leg(a, a.cpu_spp)
    print("CPULEG " + json.dumps({"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}), flush=True)


def cpu_leg_child(a):
    argv = [sys.executable, os.path.abspath(__file__), "--cpu-worker", "--width", str(a.width), "--height", str(a.height), "--depth", str(a.depth),
            "--spp", str(a.spp), "--seed", str(a.seed), "--scene", a.scene, "--cpu-spp", str(a.cpu_spp)]
    try:
        out = subprocess.run(argv, capture_output=True, text=True, timeout=600)
        return json.loads([l for l in out.stdout.splitlines() if l.startswith("CPULEG ")][-1][7:])
    except Exception as e:
        return {"unavailable": "cpu baseline run failed: %s" % e}


RTOW_PATHS = 1280 * 720 * 10            # rtow.cxx:85, 96-99: hard-coded image and samples
RTOW_SEG_PER_PATH = 1625532 / 600000.   # instrumented-reference path statistic (tests/golden/rtow_pin.npz stat_300x200)
RTOW_MD5 = "2c912270982463c81cf15fc57be2a8d6"


class RtowRun:
    """The UNMODIFIED reference program (oracle/_ref/rtow = g++ -O2 /root/reference/rtow.cxx, built by
    oracle/Makefile where the reference tree exists) timed on one host core beside the GPU run
    (/root/reference/rtow.cxx:82-122: single-threaded, no options: 1280x720, 10 spp, depth 50,
    its own rand()-built scene).  Started before the GPU legs, collected after them."""

    def __init__(self):
        self.exe = os.path.join(ROOT, "oracle", "_ref", "rtow")
        self.proc, self.t0 = None, None
        if os.path.exists(self.exe):
            self.t0 = time.perf_counter()
            self.proc = subprocess.Popen([self.exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)

    def result(self):
        if self.proc is None:
            return {"unavailable": "oracle/_ref/rtow not built (needs /root/reference at build time)"}
        import hashlib
        try:
            out, _ = self.proc.communicate(timeout=900)
        except Exception as e:
            self.proc.kill()
            return {"unavailable": "rtow run failed: %s" % e}
        wall = time.perf_counter() - self.t0
        md5 = hashlib.md5(out).hexdigest()
        return {"kind": "reference", "binary": "oracle/_ref/rtow (g++ -O2, unmodified /root/reference/rtow.cxx)", "cores": 1,
                "sample": "the program as shipped: 1280x720, 10 spp, depth 50, its own 487-sphere scene (no options exist)",
                "wall_s": wall, "paths_per_s": RTOW_PATHS / wall, "value": RTOW_PATHS * RTOW_SEG_PER_PATH / wall, "unit": UNIT,
                "segments_per_path": RTOW_SEG_PER_PATH, "md5": md5, "md5_is_the_pinned_one": md5 == RTOW_MD5,
                "note": "ran concurrently with the GPU legs of this bench (one of the host's cores)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, cores, desc, ms = cpu_leg(a, a.cpu_spp, steps=max(a.steps, 1), warmup=min(a.warmup, 1))
    rtow = None if a.no_rtow else RtowRun()                   # after the port's steps: alone on the box
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "step": "bounded sample: " + desc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if rtow is not None:
        line["cpu_baseline"]["reference_binary"] = rtow.result()
        line["cpu_baseline"]["reference_binary"]["note"] = "ran alone, after the timed steps of the port"
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- counted work
def count_worker(a):
    """Runs in a child process with RTX_LIB = librtx_count.so (the instrumented build): one frame
    of the same scene at --count-spp, traversal events from the kernel's own counters."""
    from rtxplay_b200 import api, scenes
    ctx = api.Context(int(os.environ.get("LOCAL_RANK", "0")))
    scenes.load(ctx, make_scene(a), a.mode, a.ndiv if a.scene == "book1" else None)
    ctx.resize(a.width, a.height)
    ctx.counters(reset=True)
    ctx.render(ctx.params(api.camera(aspratio=a.width / a.height), a.count_spp, a.depth, seed=a.seed))
    c = ctx.counters()
    c["segments"] = ctx.stats()["segments"]
    fs = ctx.frame_stats()
    c["lanes_per_step"], c["steps"], c["live_paths"], c["kernel"] = fs.get("lanes_per_step"), fs.get("steps"), fs.get("live_paths"), fs.get("kernel")
    ctx.close()
    print("COUNTED " + json.dumps(c), flush=True)


def counted_leg(a):
    """work per segment from the kernel's own node-visit / primitive-test counters,
    in the FLOP and byte conventions of the model (box test 23 flop, triangle 56, sphere 29, shading
    60; a node record 128 B, a triangle record 64 B, a thing's traversal record 128 + 16 B)."""
    lib = os.path.join(ROOT, "rtxplay_b200", "librtx_count.so")
    if not os.path.exists(lib):
        return {"unavailable": "librtx_count.so not built"}
    argv = [sys.executable, os.path.abspath(__file__), "--count-worker", "--width", str(a.width), "--height", str(a.height), "--depth", str(a.depth),
            "--mode", a.mode, "--seed", str(a.seed), "--scene", a.scene, "--count-spp", str(a.count_spp)] + (["--ndiv", str(a.ndiv)] if a.ndiv is not None else [])
    try:
        out = subprocess.run(argv, env=dict(os.environ, RTX_LIB=lib), capture_output=True, text=True, timeout=300)
        c = json.loads([l for l in out.stdout.splitlines() if l.startswith("COUNTED ")][-1][8:])
    except Exception as e:                                   # an instrument: its failure must not cost the bench line
        return {"unavailable": "counted-work run failed: %s" % e}
    n = float(max(c["segments"], 1))
    nodes, leaves, tris, things = c["nodes"] / n, c["leaves"] / n, c["tris"] / n, c["things"] / n
    other = c["culled_or_sphere_tests"] / n
    if a.mode == "mesh":
        flop = 92 * nodes + 56 * tris + 25 * things + 60
        byts = 128 * nodes + 64 * tris + 16 * things + 128 * (things - other) + 128
    else:
        flop = 92 * nodes + 29 * other + 60
        byts = 128 * nodes + 32 * things + 128
    return {"sample": "%dx%d, %d spp, instrumented build (one global atomic per event)" % (a.width, a.height, a.count_spp),
            "segments": c["segments"], "node_steps_per_segment": nodes, "leaf_steps_per_segment": leaves, "triangle_tests_per_segment": tris,
            "thing_visits_per_segment": things, ("culled_visits_per_segment" if a.mode == "mesh" else "sphere_tests_per_segment"): other,
            "flop_per_segment": flop, "bytes_per_segment": byts,
            "lanes_per_step": c.get("lanes_per_step"), "warp_steps_per_segment": {k: v / n for k, v in (c.get("steps") or {}).items()},
            "live_paths_per_bounce": c.get("live_paths")}


# ------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------- GPU leg
class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def run_b200(a):
    import torch
    import torch.distributed as dist
    from rtxplay_b200 import api, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # scene + acceleration structures (outside the timed window, like optx/rtwo.cxx:133-250)
    ctx = api.Context(local)
    spheres = make_scene(a)
    scenes.load(ctx, spheres, a.mode, a.ndiv if a.scene == "book1" else None)
    ctx.resize(a.width, a.height)
    cam = api.camera(aspratio=a.width / a.height)
    st0 = ctx.stats()
    mesh_stages, top_stages = ctx.build_stages()
    # L2 read peak of this box: all SMs read a 32 MiB buffer 1000 times past L1
    l2_peak = ctx.probe_read(32 << 20, 1000) if rank == 0 else None
    n_prims = st0["n_triangles_instanced"] if a.mode == "mesh" else st0["n_things"]

    # this rank's share of the frame's samples: global indices rank, rank+world, ...
    spp_local = a.spp // world + (1 if rank < a.spp % world else 0)
    ptr, nbytes = ctx.device_ptr(api.BUF_ACCUM)
    accum = torch.as_tensor(_DevArr(ptr, nbytes // 8), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    image_host = torch.empty((a.height, a.width, 4), dtype=torch.uint8).pin_memory()
    L, c = ctx._L, ctx._c

    def frame(e2e):
        p = ctx.params(cam, spp_local, a.depth, seed=a.seed, sample0=rank, sample_stride=world)
        if world == 1 and not e2e:
            ctx.render(p)                       # k_render + k_resolve
            return
        ctx.render_accumulate(p)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().synchronize()
        if rank == 0:
            ctx.resolve(a.spp)
            if e2e:
                ctx.postproc(api.PP_SRGB)
                ctx._ck(L.rtx_read(c, ctypes.c_int(api.BUF_IMAGE), ctypes.c_void_p(image_host.data_ptr()), ctypes.c_size_t(image_host.numel())))

    def timed(steps, e2e):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        t0 = time.perf_counter()
        kern_ms = 0.0
        for _ in range(steps):
            flush.fill_(1)                      # L2 flush: 256 MiB write, larger than the 126 MB L2
            frame(e2e)
            kern_ms += ctx.last_render_ms()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms, wall * 1e3, kern_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    rtow = RtowRun() if (rank == 0 and world == 1 and not a.no_cpu and not a.no_rtow) else None
    for _ in range(max(a.warmup, 0)):
        flush.fill_(1)
        frame(False)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.stats()["launches"]
    ms_dev, ms_wall, kern_ms = timed(a.steps, False)
    st = ctx.stats()
    launches = st["launches"] - launches0 - 1          # minus the k_sum_segments of the stats() call itself
    segs_local = st["segments"] if world == 1 else None
    # segments of the whole frame: after the reduce rank 0 holds the summed counters
    seg_t = torch.tensor([st["segments"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(seg_t, src=0)
    segments = int(seg_t.item())
    e2e_dev, e2e_wall, _ = timed(a.steps, True) if not a.no_e2e else (ms_dev, ms_wall, 0.)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank == 0:
        ms_step = ms_dev / a.steps
        value = segments / (ms_step * 1e-3)
        e2e_ms = max(e2e_dev, e2e_wall) / a.steps
        # roofline of k_render: per segment 64*D + B_prim + 128 bytes, 46*D + P + 60 flop
        D = max(1, math.ceil(math.log2(max(n_prims, 2))))
        b_prim, p_flop = (48, 56) if a.mode == "mesh" else (16, 29)
        bytes_seg, flop_seg = 64 * D + b_prim + 128, 46 * D + p_flop + 60
        k_ms = kern_ms / a.steps
        segs_launch = segments / world                   # each rank's launch traces its share
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = segs_launch * bytes_seg / (k_ms * 1e-3) / 1e9
        # committed ncu captures of this same configuration: DRAM bytes of one frame (proof that the working set is
        # cache resident) and the kernel's executed warp instructions / lanes per instruction per segment
        traffic, traffic_src, issue = None, None, None
        std_cfg = world == 1 and a.scene == "book1" and a.mode == "mesh" and a.ndiv is None and (a.width, a.height, a.spp) == (1200, 800, 500)
        for rnd in ("r02", "r01"):
            tpath = os.path.join(ROOT, "profiles", rnd + "_k_render_traffic.json")
            if traffic is None and os.path.exists(tpath) and std_cfg:
                t = json.load(open(tpath))
                traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], "profiles/%s_k_render_traffic.json (%s)" % (rnd, t["source"])
            ipath = os.path.join(ROOT, "profiles", rnd + "_k_render_issue.json")
            if issue is None and os.path.exists(ipath) and a.scene == "book1" and a.mode == "mesh" and a.ndiv is None:
                issue = json.load(open(ipath))
                issue["file"] = "profiles/%s_k_render_issue.json" % rnd
        clocks = sampler.summary() if sampler else {}
        sm_mhz = clocks.get("sm_mhz") or 1965
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        # The bound the kernel has (profiles/): instruction issue x SIMD efficiency.  achieved = warp
        # instructions issued per second = (warp instructions per segment of the committed ncu capture) x segments of
        # the launch / its CUDA-event duration measured here; peak = 4 schedulers x 148 SMs x the SM clock under load.
        issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9
        if issue is not None:
            issue_ach = issue["warp_instructions_per_segment"] * segs_launch / (k_ms * 1e-3) / 1e9
            roof = {"bound": "issue", "kernel": issue.get("kernel", "k_render"), "achieved": issue_ach, "peak": issue_peak, "unit": "Gwarp-inst/s",
                    "frac": issue_ach / issue_peak, "peak_source": "4 warp schedulers x 148 SMs x %d MHz (median SM clock sampled during the timed region)" % sm_mhz,
                    "simd": {"lanes_per_instruction": issue["lanes_per_instruction"], "frac": issue["lanes_per_instruction"] / 32.,
                             "lane_instructions_per_segment": issue["warp_instructions_per_segment"] * issue["lanes_per_instruction"]},
                    "useful_frac": issue_ach / issue_peak * issue["lanes_per_instruction"] / 32.,
                    "ncu": {k: issue.get(k) for k in ("issue_active_pct", "l1_data_pipe_pct", "l1_hit_pct", "l2_hit_pct", "registers", "warps_active_pct", "file", "command")},
                    "note": "issue slots used x lanes per instruction: the kernel is bound by instruction issue and SIMD efficiency with the L1 data pipe as co-limiter (DESIGN.md 4); "
                            "`traffic` = DRAM bytes of one frame from the committed ncu pass, against %.1f TB of algorithmic node + primitive bytes -- the working set is L1/L2 resident, "
                            "so HBM is not the bound (the contract's hbm figure is kept under `hbm_model`)" % (segs_launch * bytes_seg / 1e12)}
        else:
            roof = {"bound": "issue", "kernel": "k_render", "achieved": None, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": None,
                    "note": "no committed ncu capture for this configuration (profiles/rNN_k_render_issue.json): issue-slot figures unavailable; see hbm_model / l2 / fp32"}
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel_ms": k_ms, "bytes_per_segment": bytes_seg,
                     "hbm_model": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                                   "what": "algorithmic bytes per segment of SURVEY.md 8(d), 64*ceil(log2 N) + 48 + 128, x segments / kernel time: a cache-bandwidth figure quoted against the HBM peak"},
                     "l2": {"achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak if l2_peak else None,
                            "peak_source": "measured in this run: rtx_probe_read, 32 MiB buffer read 1000x by all SMs with ld.global.cg.v4"},
                     "fp32": {"flop_per_segment": flop_seg, "achieved_tflops": segs_launch * flop_seg / (k_ms * 1e-3) / 1e12,
                              "peak_tflops": fp32_peak, "frac": segs_launch * flop_seg / (k_ms * 1e-3) / 1e12 / fp32_peak}})
        fs = ctx.frame_stats()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "segments_per_frame": segments, "paths_per_frame": a.width * a.height * a.spp,
                       "build_ms": {"meshes": st["ms_build_blas"], "top_level": st["ms_build_tlas"], "mesh_stages": mesh_stages, "top_level_stages": top_stages}, "device_mb": st["bytes_device"] / 1048576.,
                       "things": st["n_things"], "triangles_instanced": st["n_triangles_instanced"], "triangles_stored": st["n_triangles"],
                       "parallelism": "spp split over %d rank(s) + NCCL reduce of the u64 accumulation buffer" % world if world > 1 else "1 GPU",
                       "l2": "256 MiB device write between steps (inside the timed region)",
                       "kernel": "k_render_q (compacting ray pool)" if fs["kernel"] else "k_render (one ray per lane)",
                       "stage_ms_last_frame": {k: fs[k] for k in ("ms_frame", "ms_trace", "ms_reduce_resolve", "ms_postproc")},
                       "ms_per_frame": ms_step, "ms_per_frame_e2e": e2e_ms, "wall_ms_per_step": ms_wall / a.steps},
            "roofline": roof,
            "e2e": {"value": segments / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ctypes.sizeof(ctx.params(cam, 1)),
                    "d2h_bytes_per_step": int(image_host.numel()), "ms_per_frame": e2e_ms, **({"skipped": "--no-e2e: these repeat the device-timed frame"} if a.no_e2e else {})},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not a.no_count and world == 1:
            cnt = counted_leg(a)
            if "flop_per_segment" in cnt:
                cnt["achieved_tflops"] = segs_launch * cnt["flop_per_segment"] / (k_ms * 1e-3) / 1e12
                cnt["frac_fp32"] = cnt["achieved_tflops"] / fp32_peak
                cnt["achieved_gbs"] = segs_launch * cnt["bytes_per_segment"] / (k_ms * 1e-3) / 1e9
                cnt["frac_l2"] = cnt["achieved_gbs"] / l2_peak if l2_peak else None
            line["roofline"]["counted"] = cnt
        if not a.no_cpu and world == 1:
            ref_bin = rtow.result() if rtow is not None else None      # (collect the one-core run first: the port then has every core)
            line["cpu_baseline"] = cpu_leg_child(a)
            if ref_bin is not None:
                line["cpu_baseline"]["reference_binary"] = ref_bin
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.count_worker:
        count_worker(a)
    elif a.cpu_worker:
        cpu_worker(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()

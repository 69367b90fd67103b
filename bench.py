#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 spectral path tracer.

    python bench.py --gpus N --steps K --warmup W            # our arm (libsrt.so through the C-ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU path

Metric (BASELINE.json): spectral path samples/s.  One "step" = one full render of the workload.
Workload at every N: BASELINE.json configs[1] -- reference Cornell scene (id 0), 1920x1080, 64 spp,
depth 10 -- split over the ranks by interleaved image tiles (strong scaling), per-rank XYZ films
summed with one NCCL reduce to rank 0 which tonemaps.  L2 is flushed before every timed step.

  value      samples / device time, scene + state resident in HBM (CUDA events inside libsrt around
             the render kernels; at N>1 plus the NCCL film reduce; max over ranks)
  e2e        the same metric through the public C-ABI with HOST buffers: every step uploads the scene
             (host triangle build + H2D + LBVH build), renders, tonemaps and copies the film back into
             caller-owned host planes; wall clock around the calls, barrier on both sides
  roofline   k_wavefront (the dominant kernel): algorithmic FLOPs (oracle op counters x the constants of
             SURVEY.md 8d) / its CUDA-event duration against the FP32 issue peak measured in this run
  cpu_baseline  the reference's own host-compiled code (oracle/_ref) or the C port (oracle/), timed on the
             box's host cores on a bounded sample of the same workload
  c5, soup   extra, same N GPUs: BASELINE configs[4] (Cornell 3840x2160, 1024 spp) and whole renders of the
             1M-triangle soup of configs[3] (1920x1080, 8 spp); 1 warm-up + 2 timed steps each
  reference_cuda / strict_fp / lbvh   (N=1) the reference's CUDA renderer on this GPU, our -fmad=false kernels,
             the 1M-triangle LBVH build and soup ray queries with their rooflines
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene id, width, height, spp, depth)
    "c1": (0, 400, 225, 8, 10),
    "c2": (0, 1920, 1080, 64, 10),
    "c3": (1, 1920, 1080, 256, 10),
    "c5": (0, 3840, 2160, 1024, 10),
}
METRIC = "spectral path samples/s"
UNIT = "samples/s"
# FLOP constants of SURVEY.md 8(d): counted from the reference source, div/sqrt/pow = 1
FLOPS = dict(box_tests=24, tri_tests=50, scatter_lm=60, rejection_iter=12, scatter_dielectric=90, interp=8, xyz_per_sample=210, tonemap_per_pixel=40)


def algorithmic_flops_per_sample(cnt, spp):
    s = float(cnt["samples"])
    f = (cnt["box_tests"] * FLOPS["box_tests"] + cnt["tri_tests"] * FLOPS["tri_tests"]
         + (cnt["scatter_lambert"] + cnt["scatter_metal"]) * FLOPS["scatter_lm"] + cnt["rejection_iters"] * FLOPS["rejection_iter"]
         + cnt["scatter_dielectric"] * FLOPS["scatter_dielectric"] + cnt["interps"] * FLOPS["interp"]) / s
    return f + FLOPS["xyz_per_sample"] + FLOPS["tonemap_per_pixel"] / spp


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = [float(s[1]) for s in self.samples if len(s) > 8 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 8 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active") and not v.lower().startswith("not"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_reference_render(w, h, scene, spp, depth, threads=0):
    """One bounded CPU render with the reference's own code when oracle/_ref is there, else the C port.
    Returns (seconds, kind, counters-or-None)."""
    import refhost
    import oracle

    if refhost.available():
        R = refhost.RefHost()
        # silence the reference's own chatter on stdout (it prints buffer sizes from init_device_params)
        sys.stdout.flush()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            R.open("-s", scene, "-xr", w, "-ar", "%d/%d" % (w, h), "-ns", spp, "-bl", depth, "--no-show")
            assert (R.W, R.H) == (w, h), (R.W, R.H)
            t0 = time.perf_counter()
            R.render()
            dt = time.perf_counter() - t0
            R.close()
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
        return dt, "reference"
    S = oracle.Scene(scene)
    cam = oracle.camera(w, h)
    t0 = time.perf_counter()
    oracle.render(S, cam, spp, depth, nthreads=threads)
    return time.perf_counter() - t0, "port"


def run_reference_arm(a, wl_name, wl):
    scene, w, h, spp, depth = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # other ranks exit without work
    sample_spp = max(1, min(spp, a.ref_spp))
    cores = os.cpu_count() or 1
    times = []
    kind = "port"
    for i in range(a.warmup + a.steps):
        dt, kind = cpu_reference_render(w, h, scene, sample_spp, depth)
        if i >= a.warmup:
            times.append(dt)
    t = float(np.mean(times))
    value = w * h * sample_spp / t
    sample = "%dx%d at %d spp of the workload's %d spp per step (rate is spp-independent), depth %d" % (w, h, sample_spp, spp, depth)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": wl_name, "scene": scene, "width": w, "height": h, "spp": spp, "depth": depth},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def reference_cuda_baseline(wl):
    """The reference's own CUDA renderer compiled for sm_100a (baseline/_ref/ref_cuda_render), if it was built."""
    exe = ROOT / "baseline" / "_ref" / "ref_cuda_render"
    if not exe.exists():
        return None
    scene, w, h, spp, depth = wl
    try:
        out = subprocess.run([str(exe), "-s", str(scene), "-xr", str(w), "-ar", "%d/%d" % (w, h), "-ns", str(spp), "-bl", str(depth), "--no-show", "--repeat", "3"],
                             capture_output=True, text=True, timeout=600, cwd=str(exe.parent)).stdout
        for ln in out.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                return {"value": d["samples_per_s"], "unit": UNIT, "median_kernel_ms": d["median_ms"], "what": "reference spectral_render_kernel recompiled for sm_100a, CUDA-event kernel time"}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-spp", type=int, default=4, help="spp of the bounded CPU sample")
    ap.add_argument("--strict", action="store_true", help="strict FP mode (-fmad=false kernels)")
    ap.add_argument("--tile-w", type=int, default=0, help="0 = library picks by pixels per rank")
    ap.add_argument("--tile-h", type=int, default=0)
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the extra BASELINE configs[4] measurement (4K, 1024 spp)")
    a = ap.parse_args()
    # torchrun exports OMP_NUM_THREADS=1; the CPU reference legs must use every host core (the OpenMP
    # runtime reads the variable when the oracle library is loaded, which happens later)
    os.environ.pop("OMP_NUM_THREADS", None)
    wl_name = a.workload
    wl = WORKLOADS[wl_name]
    scene_id, w, h, spp, depth = wl
    if a.impl == "reference":
        return run_reference_arm(a, wl_name, wl)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import srt_b200 as S

    if S.lib().srt_device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- libsrt has no CPU fallback")
    S.lib().srt_set_device(local_rank)
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    launches0 = S.kernel_launch_count()

    # ---- resident-state arm: scene + render manager built once, every step re-renders the image
    sc = S.Scene(scene_id)
    cam = sc.camera(w, h)
    fb = S.FrameBuffer(w, h)
    rm = S.RenderManager(sc, cam, fb)
    rm.init_renderer(depth, spp)
    rm.set_option(S.OPT_FP_MODE, 1 if a.strict else 0)
    rm.set_option(S.OPT_KERNEL_TIMING, 1)
    rm.set_option(S.OPT_TILE_W, a.tile_w); rm.set_option(S.OPT_TILE_H, a.tile_h)
    if world > 1:
        rm.set_option(S.OPT_RANK, rank); rm.set_option(S.OPT_WORLD, world)
    rm.init_device_params(0, 0)
    film = None
    if world > 1:
        class _Film:  # zero-copy view of libsrt's device film for torch.distributed
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        film = torch.as_tensor(_Film(rm.device_film(), 3 * w * h), device="cuda")

    def one_step():
        """returns device milliseconds of this rank (render kernels [+ film reduce])"""
        S.lib().srt_measure_copy_gbs(256)  # untimed 256 MB device copy (2x + 2x the 126 MB L2): evicts the previous step's film / RNG / path state
        rm.restart()
        while rm.step():
            pass
        st = rm.stats()
        ms = st["render_ms"]
        if world > 1:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, st

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    dev_ms, kern_ms = [], []
    st = None
    launches_timed0 = S.kernel_launch_count()
    for _ in range(a.steps):
        ms, st = one_step()
        dev_ms.append(ms)
        kern_ms.append(st["wavefront_ms"] if st["wavefront_ms"] > 0 else st["render_ms"])
    launches = S.kernel_launch_count() - launches_timed0
    barrier()
    clocks = sampler.summary() if sampler else None
    step_ms = float(np.mean(dev_ms))
    if world > 1:
        tmax = torch.tensor([step_ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        step_ms = float(tmax.item())
    if rank == 0:
        rm.resolve_film()
    total_samples = w * h * spp
    value = total_samples / (step_ms * 1e-3)

    # ---- BASELINE configs[4] next to the headline: Cornell 3840x2160, 1024 spp on the same N GPUs (device time incl.
    # the film reduce, max over ranks).  Its per-pixel chains are 16x longer than C2's, so it shows how the tile split
    # scales when the end-of-kernel drain is amortised; 1 warm-up + 2 timed steps keep it to seconds.
    def resident_leg(scene_obj, w_, h_, spp_, d_, warm=1, steps=2):
        """device ms per step (render [+ film reduce], max over ranks) of one more workload on the same N GPUs"""
        fb_ = S.FrameBuffer(w_, h_)
        rm_ = S.RenderManager(scene_obj, scene_obj.camera(w_, h_), fb_)
        rm_.init_renderer(d_, spp_)
        rm_.set_option(S.OPT_TILE_W, a.tile_w); rm_.set_option(S.OPT_TILE_H, a.tile_h)
        if world > 1:
            rm_.set_option(S.OPT_RANK, rank); rm_.set_option(S.OPT_WORLD, world)
        rm_.init_device_params(0, 0)
        film_ = torch.as_tensor(_Film(rm_.device_film(), 3 * w_ * h_), device="cuda") if world > 1 else None
        ms_ = []
        for i in range(warm + steps):
            S.lib().srt_measure_copy_gbs(256)
            rm_.restart()
            while rm_.step():
                pass
            ms = rm_.stats()["render_ms"]
            if world > 1:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                dist.reduce(film_, dst=0, op=dist.ReduceOp.SUM)
                e1.record()
                e1.synchronize()
                ms += e0.elapsed_time(e1)
            if i >= warm:
                ms_.append(ms)
        m = float(np.mean(ms_))
        rays_per_sample = rm_.stats()["rays"] / max(1, rm_.stats()["samples"])
        if world > 1:
            tmax = torch.tensor([m], device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            m = float(tmax.item())
        del rm_, film_, fb_
        return m, rays_per_sample

    c5_info = None
    soup_info = None
    if wl_name == "c2" and not a.no_c5 and not a.strict:
        s5, w5, h5, spp5, d5 = WORKLOADS["c5"]
        m5, _ = resident_leg(S.Scene(s5), w5, h5, spp5, d5)
        c5_info = {"workload": "c5", "width": w5, "height": h5, "spp": spp5, "depth": d5, "value": w5 * h5 * spp5 / (m5 * 1e-3), "unit": UNIT,
                   "ms_per_step": m5, "steps": 2, "warmup": 1, "what": "BASELINE configs[4] on the same GPUs, device time incl. the film reduce, max over ranks"}
        # BASELINE configs[3] as a render: the seeded 1M-triangle soup (LBVH walk from global memory), 1920x1080, 8 spp
        soup_sc = S.Scene(soup=1 << 20, seed=1984)
        ms_s, rps = resident_leg(soup_sc, 1920, 1080, 8, 10)
        soup_info = {"workload": "c4-render", "n_tris": 1 << 20, "generator": "SplitMix64 seed 1984 (srt_scene_create_soup)", "width": 1920, "height": 1080, "spp": 8,
                     "depth": 10, "value": 1920 * 1080 * 8 / (ms_s * 1e-3), "unit": UNIT, "rays_per_s": 1920 * 1080 * 8 * rps / (ms_s * 1e-3),
                     "ms_per_step": ms_s, "steps": 2, "warmup": 1, "what": "whole renders of the 1M-triangle soup on the same GPUs (every rank holds the full LBVH)"}
        del soup_sc

    # ---- the same resident-state measurement in strict FP mode (-fmad=false kernels: the mode whose film is
    # bit-identical to the reference's host-compiled image); reported next to the headline, single GPU only
    strict_info = None
    if world == 1 and not a.strict:
        fb_s = S.FrameBuffer(w, h)
        rm_s = S.RenderManager(sc, cam, fb_s)
        rm_s.init_renderer(depth, spp)
        rm_s.set_option(S.OPT_FP_MODE, 1)
        rm_s.set_option(S.OPT_TILE_W, a.tile_w); rm_s.set_option(S.OPT_TILE_H, a.tile_h)
        rm_s.init_device_params(0, 0)
        ms_s = []
        for i in range(2 + min(a.steps, 3)):
            rm_s.restart()
            while rm_s.step():
                pass
            if i >= 2:
                ms_s.append(rm_s.stats()["render_ms"])
        strict_info = {"value": total_samples / (float(np.mean(ms_s)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ms_s)),
                       "what": "same workload, kernels built with -fmad=false: film bit-identical to the reference's host build (tests/test_gpu_parity.py)"}
        del rm_s

    # ---- end-to-end arm: host buffers in, host film out, every step
    def e2e_step():
        sc2 = S.Scene(scene_id)  # host triangle/material build + H2D + device LBVH
        fb2 = S.FrameBuffer(w, h)
        rm2 = S.RenderManager(sc2, sc2.camera(w, h), fb2)
        rm2.init_renderer(depth, spp)
        rm2.set_option(S.OPT_FP_MODE, 1 if a.strict else 0)
        rm2.set_option(S.OPT_TILE_W, a.tile_w); rm2.set_option(S.OPT_TILE_H, a.tile_h)
        if world > 1:
            rm2.set_option(S.OPT_RANK, rank); rm2.set_option(S.OPT_WORLD, world)
        rm2.init_device_params(0, 0)
        if world > 1:
            while rm2.step():
                pass
            f2 = torch.as_tensor(_Film(rm2.device_film(), 3 * w * h), device="cuda")
            dist.reduce(f2, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                rm2.resolve_film()
        else:
            rm2.render_all()  # worker thread renders, caller thread resolves + copies D2H into fb2
        return float(fb2.r.sum())

    e2e_warm = min(a.warmup, 2)
    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / a.steps
    if world > 1:
        tmax = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    h2d = sc.ntris * (9 * 4 + 48) + sc.nmats * 416 + 4 * 95 * 4
    d2h = 3 * w * h  # the film crosses PCIe as one byte per channel (values are 0..255); the host widens it to float planes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    import oracle
    fp32_peak = S.lib().srt_measure_fp32_tflops()
    cw, chh, cspp = 480, 270, 4  # bounded, deterministic counter sample of the same scene/camera
    _, _, cnt = oracle.render(oracle.Scene(scene_id), oracle.camera(cw, chh), cspp, depth, counters=True)
    fl_sample = algorithmic_flops_per_sample(cnt, spp)
    k_ms = float(np.mean(kern_ms))
    samples_rank = st["samples"]
    achieved = fl_sample * samples_rank / (k_ms * 1e-3) / 1e12
    traffic = None
    tfile = ROOT / "profiles" / "r1_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(wl_name, {}).get("k_wavefront_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "kernel": "k_wavefront", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                "traffic": traffic, "peak_source": "measured in this run (srt_measure_fp32_tflops: FFMA chains on all SMs); MEASURED_PEAKS.json has no FP32 figure",
                "algorithmic_flops_per_sample": fl_sample, "kernel_ms": k_ms, "rays_per_sample": st["rays"] / max(1, st["samples"])}

    # ---- LBVH build (second half of the BASELINE metric): 1M-triangle soup, device resident
    del rm  # the renderer's persisting-L2 set-aside goes back to the normal cache before other kernels are timed
    lb = None
    if world == 1:
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            hbm = float(peaks["hbm_gbs"]); hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm = 6650.0; hbm_src = "fallback (B200_PROFILING.md)"
        soup = S.Scene(soup=1 << 20, seed=1984)
        soup.rebuild_lbvh(3)
        ms = [soup.rebuild_lbvh(1)["total"] for _ in range(10)]
        t = float(np.median(ms))
        ach = 256.0 * (1 << 20) / (t * 1e-3) / 1e9
        phases = soup.rebuild_lbvh(1)
        cub = None
        exe = ROOT / "baseline" / "_ref" / "cub_sort_bench"
        if exe.exists():  # library sort of the same pairs, timing context only
            try:
                out = subprocess.run([str(exe), str(1 << 20), "20"], capture_output=True, text=True, timeout=120).stdout
                cub = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
            except Exception as e:  # pragma: no cover
                cub = {"error": str(e)}
        lb = {"n_tris": 1 << 20, "build_ms": t, "phases_ms": phases, "sort_baseline_cub": cub, "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                                                              "peak_source": hbm_src, "algorithmic_bytes_per_tri": 256}}
        # BASELINE.json configs[3]: closest-hit rays/s through the device LBVH of the same soup
        # (primary = the reference camera at 1920x1080, secondary = one random bounce off the primary hits)
        cam_a = soup.camera(1920, 1080).as_array()
        ys, xs = np.mgrid[0:1080, 0:1920]
        d = (cam_a[8:11][None, :] + xs.reshape(-1, 1) * cam_a[2:5][None, :] + ys.reshape(-1, 1) * cam_a[5:8][None, :] - cam_a[12:15][None, :]).astype(np.float32)
        o = np.tile(cam_a[12:15], (d.shape[0], 1)).astype(np.float32)
        tt, tri, ms1, v1 = soup.trace_rays(o, d, counted=True)
        hit = tri >= 0
        rs = np.random.RandomState(1)
        d2 = rs.randn(int(hit.sum()), 3).astype(np.float32)
        o2 = (o[hit] + tt[hit, None] * d[hit] + 1e-3 * d2).astype(np.float32)
        _, _, ms2, v2 = soup.trace_rays(o2, d2, counted=True)
        def walk_roofline(visits, ms):  # 32 B per node visit + 48 B per triangle test
            ach = (32.0 * visits[0] + 48.0 * visits[1]) / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "peak_source": hbm_src}
        lb["trace"] = {"primary_rays_per_s": d.shape[0] / (ms1 * 1e-3), "secondary_rays_per_s": d2.shape[0] / (ms2 * 1e-3), "primary_hit_fraction": float(hit.mean()),
                       "primary_nodes_per_ray": v1[0] / d.shape[0], "primary_tris_per_ray": v1[1] / d.shape[0],
                       "secondary_nodes_per_ray": v2[0] / max(1, d2.shape[0]), "secondary_tris_per_ray": v2[1] / max(1, d2.shape[0]),
                       "primary_roofline": walk_roofline(v1, ms1), "secondary_roofline": walk_roofline(v2, ms2)}
        del soup

    cpu_base = None
    if not a.no_cpu_baseline and world == 1:
        dt, kind = cpu_reference_render(w, h, scene_id, a.ref_spp, depth)
        cpu_base = {"value": w * h * a.ref_spp / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                    "sample": "%dx%d at %d spp of %d, depth %d, one render" % (w, h, a.ref_spp, spp, depth)}
    ref_cuda = None if (a.no_ref_cuda or world > 1) else reference_cuda_baseline(wl)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "scene": scene_id, "width": w, "height": h, "spp": spp, "depth": depth,
                   "fp_mode": "strict(-fmad=false)" if a.strict else "fast(fma, as the reference's nvcc build)",
                   "parallelism": "image tiles (auto size) interleaved over %d ranks + nccl film reduce" % world if world > 1 else "single gpu",
                   "l2": "flushed before every step (untimed 256 MB device-to-device copy); film, RNG and path state are re-initialised every step"},
        "clocks": clocks,
        "e2e": {"value": total_samples / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_base, "c5": c5_info, "soup": soup_info,
        "reference_cuda": ref_cuda,
        "strict_fp": strict_info,
        "lbvh": lb,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

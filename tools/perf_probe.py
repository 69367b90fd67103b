#!/usr/bin/env python3
"""Quick performance probe on the GPU box (not a bench): renders a config several ways and prints stats."""
import argparse, json, sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument("--scene", type=int, default=0)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--variants", default="wave,wave_t,mega")
a = ap.parse_args()
print("fp32 peak TFLOP/s:", S.lib().srt_measure_fp32_tflops(), " copy GB/s:", S.lib().srt_measure_copy_gbs(1024))
sc = S.Scene(a.scene)
def run(name, **kw):
    best = None
    for rep in range(3):
        rgb, xyz, st = S.render(scene=sc, w=a.w, h=a.h, spp=a.spp, bounce=10, **kw)
        if best is None or st["render_ms"] < best["render_ms"]:
            best = st
    st = best
    print("%-28s %8.2f ms  %7.3f Gsamples/s  %6.3f Grays/s  rounds %d launches %5d  wavefront %.2f mega %.2f order %.2f other %.2f drain %.2f" % (
        name, st["render_ms"], st["samples"] / st["render_ms"] / 1e6, st["rays"] / st["render_ms"] / 1e6,
        st["rounds"], st["kernel_launches"], st["wavefront_ms"], st["megakernel_ms"], st["order_ms"], st["other_ms"], st["drain_ms"]), flush=True)
for v in a.variants.split(","):
    if v == "wave": run("wavefront", pipeline=0)
    elif v == "wave_t": run("wavefront+timing", pipeline=0, kernel_timing=True)
    elif v == "mega": run("megakernel", pipeline=1)
    elif v == "strict": run("wavefront strict", pipeline=0, strict=True)
    elif v == "pass1": run("wavefront pass scheduler, 1 round", pipeline=0, sched_flags=0, rounds=1)
    elif v == "bvh": run("wavefront lbvh-walk smem", pipeline=0, traversal=1)
    elif v == "bvhg": run("wavefront lbvh-walk global", pipeline=0, traversal=3)
    elif v.startswith("rounds"): run("wavefront %s rounds" % v[6:], pipeline=0, rounds=int(v[6:]), kernel_timing=True)
    elif v.startswith("bs"): run("wavefront %s paths per block" % v[2:], pipeline=0, block_slots=int(v[2:]))
    elif v.startswith("tile"):
        tw, th = v[4:].split("x"); run("wavefront tile %sx%s" % (tw, th), pipeline=0, tiles=(int(tw), int(th), 0, 1))

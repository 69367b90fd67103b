"""Ablation of the wavefront scheduler features (SRT_OPT_SCHED_FLAGS) on rank 0's share of an N-rank split of C2."""
import sys, pathlib, itertools
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4,8").split(",")]
flags = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,4,3,7").split(",")]
rounds = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2").split(",")]
w, h, spp = 1920, 1080, 64
for world in worlds:
    for fl, rd in itertools.product(flags, rounds):
        best = None
        for rep in range(3):
            rgb, xyz, st = S.render(scene=sc, w=w, h=h, spp=spp, bounce=10, tiles=(0, 0, 0, world), rounds=rd, sched_flags=fl, kernel_timing=True)
            if best is None or st["render_ms"] < best["render_ms"]:
                best = st
        print("world %d flags %d rounds %d: render %7.2f ms  wavefront %7.2f order %.3f drain %6.2f  (ideal %.2f)" % (world, fl, rd, best["render_ms"], best["wavefront_ms"], best["order_ms"], best["drain_ms"], 37.0 / world), flush=True)

"""Sweep of paths per block / threads per block / rounds on one GPU rendering rank 0's share of an N-rank split (C2)."""
import sys, pathlib, itertools
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8").split(",")]
slots = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,64,128,256,512").split(",")]
threads = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,128").split(",")]
rounds = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1,2").split(",")]
flags = [int(x) for x in (sys.argv[5] if len(sys.argv) > 5 else "0").split(",")]
w, h, spp = 1920, 1080, 64
for world in worlds:
    for bs, bt, rd, fl in itertools.product(slots, threads, rounds, flags):
        best = None
        for rep in range(3):
            kw = {}
            if bs: kw["block_slots"] = bs
            if bt: kw["block_threads"] = bt
            rgb, xyz, st = S.render(scene=sc, w=w, h=h, spp=spp, bounce=10, tiles=(0, 0, 0, world), rounds=rd, sched_flags=fl, kernel_timing=True, **kw)
            if best is None or st["render_ms"] < best["render_ms"]:
                best = st
        print("world %d slots %4d threads %3d rounds %d flags %d: render %7.2f ms  drain %6.2f  (ideal %.2f)" % (world, bs, bt, rd, fl, best["render_ms"], best["drain_ms"], 37.0 / world), flush=True)

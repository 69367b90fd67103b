"""End-of-launch drain of k_wavefront vs the number of rounds, on one GPU rendering rank 0's share of an N-rank split.
Prints render ms, the summed drain (device globaltimer, see srt_stats.drain_ms) and checks that the film does not depend on the rounds."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
cfgs = ([] if "--only-c5" in sys.argv else [("C2", 1920, 1080, 64)]) + ([("C5/4", 3840, 2160, 256)] if "--c5" in sys.argv or "--only-c5" in sys.argv else [])
for name, w, h, spp in cfgs:
    for world in (1, 2, 4, 8):
        ref = None
        for rounds in (1, 2, 3, 4):
            best = None
            for rep in range(3):
                rgb, xyz, st = S.render(scene=sc, w=w, h=h, spp=spp, bounce=10, tiles=(0, 0, 0, world), rounds=rounds, kernel_timing=True)
                if best is None or st["render_ms"] < best["render_ms"]:
                    best = st
            same = True if ref is None else bool(np.array_equal(ref.view(np.uint32), xyz.view(np.uint32)))
            if ref is None:
                ref = xyz
            print("%s world %d rounds %d: render %7.2f ms (wavefront %7.2f, order %.3f) drain %6.2f ms  ideal %.2f  film identical to rounds=1: %s" % (
                name, world, rounds, best["render_ms"], best["wavefront_ms"], best["order_ms"], best["drain_ms"], 0, same), flush=True)

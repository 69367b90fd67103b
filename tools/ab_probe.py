"""A/B timing of two builds of libsrt.so on the same GPU box: SRT_LIB=<path> python tools/ab_probe.py  (C2, rank 0 of world N)"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
scene = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = S.Scene(scene)
for world in worlds:
    ts = []
    for rep in range(5):
        rgb, xyz, st = S.render(scene=sc, w=1920, h=1080, spp=64, bounce=10, tiles=(0, 0, 0, world))
        ts.append(st["render_ms"])
    print("%s scene %d world %d: best %.2f median %.2f ms" % (S.LIB_PATH.name, scene, world, min(ts), sorted(ts)[2]), flush=True)

"""Single-GPU proxy of the per-rank work of an N-GPU run: render only rank r's tiles of world N."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
def run(name, w, h, spp, world, rank, tw=0, th=0, **kw):
    best = 1e9
    for rep in range(3):
        rgb, xyz, st = S.render(scene=sc, w=w, h=h, spp=spp, bounce=10, tiles=(tw, th, rank, world), **kw)
        best = min(best, st["render_ms"])
    print("%-4s world %d rank %d tile %2dx%-2d %s: %8.2f ms  -> aggregate %.2f Gsamples/s" % (name, world, rank, tw, th, kw, best, w * h * spp / best / 1e6), flush=True)
    return best
run("C2", 1920, 1080, 64, 1, 0)
for world in (2, 4, 8):
    ts = [run("C2", 1920, 1080, 64, world, r) for r in range(world)]
    print("   world %d: max %.2f ms, mean %.2f ms" % (world, max(ts), sum(ts) / len(ts)))
for tw, th in ((32, 32), (32, 16), (16, 16), (8, 8)):
    run("C2", 1920, 1080, 64, 8, 0, tw, th)

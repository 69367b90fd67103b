"""Single-GPU proxy of the per-rank work of an N-GPU run: render only rank 0's tiles of world N."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
def run(name, w, h, spp, world, tw, th, threads):
    best = 1e9
    for rep in range(3):
        rgb, xyz, st = S.render(scene=sc, w=w, h=h, spp=spp, bounce=10, tiles=(tw, th, 0, world), block_threads=threads)
        best = min(best, st["render_ms"])
    print("%-10s world %d tile %2dx%-2d threads %3d : %8.2f ms  -> aggregate %.2f Gsamples/s" % (name, world, tw, th, threads, best, w * h * spp / best / 1e6), flush=True)
for world in (1, 8):
    for (tw, th, thr) in ((32, 16, 256), (16, 16, 256), (16, 16, 128), (16, 8, 128), (32, 8, 128), (8, 8, 64), (16, 8, 64)):
        run("C2", 1920, 1080, 64, world, tw, th, thr)

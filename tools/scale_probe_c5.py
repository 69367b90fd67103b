import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
def run(w, h, spp, world, tw, th, rank=0):
    rgb, xyz, st = S.render(scene=sc, w=w, h=h, spp=spp, bounce=10, tiles=(tw, th, rank, world))
    print("C5 world %d rank %d tile %2dx%-2d : %8.2f ms  -> aggregate %.2f Gsamples/s (samples this rank %.3g)" % (world, rank, tw, th, st["render_ms"], w * h * spp / st["render_ms"] / 1e6, st["samples"]), flush=True)
for (tw, th) in ((32, 16), (16, 16), (32, 32), (16, 8)):
    run(3840, 2160, 1024, 8, tw, th)
run(3840, 2160, 1024, 8, 32, 16, rank=3)
run(3840, 2160, 1024, 8, 16, 16, rank=5)

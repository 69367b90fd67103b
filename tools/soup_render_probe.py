"""Whole render of the 1M-triangle soup through the wavefront (LBVH walk from global memory): the ncu target for k_wavefront<false,false,*>."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = S.Scene(soup=n, seed=1984)
for rep in range(2):
    rgb, xyz, st = S.render(scene=sc, w=1920, h=1080, spp=spp, bounce=10)
    print("render 1080p %d spp depth 10: %.1f ms, %.3f Gsamples/s, %.2f Grays/s" % (spp, st["render_ms"], st["samples"] / st["render_ms"] / 1e6, st["rays"] / st["render_ms"] / 1e6), flush=True)

import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
import numpy as np
for scene in (0, 1, 2):
    for strict in (False, True):
        rgb, xyz, st = S.render(scene_id=scene, w=100, h=57, spp=3, bounce=10, strict=strict)
rgb, xyz, st = S.render(scene_id=0, w=96, h=54, spp=2, bounce=10, chunk=(48, 27))
rgb, xyz, st = S.render(scene_id=0, w=96, h=54, spp=2, bounce=10, tiles=(16, 16, 1, 3))
rgb, xyz, st = S.render(scene_id=0, w=64, h=36, spp=2, bounce=10, pipeline=1)
sg = S.Scene(soup=5000, seed=3)
rgb, xyz, st = S.render(scene=sg, w=64, h=36, spp=2, bounce=5)
rgb, xyz, st = S.render(scene=sg, w=64, h=36, spp=2, bounce=5, traversal=3)
o = (np.random.rand(1000, 3) * 555).astype(np.float32); d = (np.random.rand(1000, 3) - 0.5).astype(np.float32)
sg.trace_rays(o, d)
S.Scene(soup=70000, seed=5).lbvh()
print("sanitize probe done")

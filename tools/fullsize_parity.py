#!/usr/bin/env python3
"""One-off evidence run (GPU box): the FULL bench workload (Cornell 1920x1080, 64 spp, depth 10) rendered by the
strict-FP GPU build and by the CPU oracle (all host cores), films compared bit for bit.  ~435 M rays go through the
conservative wide-leaf pre-test; a single wrongly rejected candidate would show up as a differing pixel."""
import sys, time, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200")); sys.path.insert(0, str(ROOT / "tests"))
import srt_b200 as S
import oracle
scene = int(sys.argv[1]) if len(sys.argv) > 1 else 0
w, h, spp = 1920, 1080, int(sys.argv[2]) if len(sys.argv) > 2 else 64
t0 = time.time()
rgb, xyz, st = S.render(scene_id=scene, w=w, h=h, spp=spp, bounce=10, strict=True)
t1 = time.time()
orgb, oxyz = oracle.render(oracle.Scene(scene), oracle.camera(w, h), spp, 10)
t2 = time.time()
diff = (xyz.view(np.uint32) != oxyz.view(np.uint32)).any(axis=0)
print("scene %d %dx%d %d spp: gpu %.2f s (kernel %.1f ms, %d rays), oracle %.1f s; pixels whose XYZ bits differ: %d of %d; max |dXYZ| %.3g; sRGB differing pixels: %d"
      % (scene, w, h, spp, t1 - t0, st["render_ms"], st["rays"], t2 - t1, int(diff.sum()), w * h, float(np.abs(xyz - oxyz).max()), int((rgb != orgb).any(axis=0).sum())))

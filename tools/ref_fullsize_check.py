"""Authoring-container check (needs oracle/_ref): the CPU oracle against the REAL reference host build on the full bench frame
(1920x1080, 64 spp).  Measured: 0 of 2 073 600 pixels differ (scene 0; ~2 minutes on 8 cores).
usage: ref_fullsize_check.py <scene> <pixel x> <pixel y>   (the pixel is printed from both)"""
import sys, time
import pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1] / "tests"))
import numpy as np, refhost, oracle
scene = int(sys.argv[1]); px, py = int(sys.argv[2]), int(sys.argv[3])
w, h, spp = 1920, 1080, 64
t0 = time.time()
o = oracle.render(oracle.Scene(scene), oracle.camera(w, h), spp, 10)[1]
t1 = time.time()
R = refhost.RefHost("ltr")
R.open("-s", scene, "-xr", w, "-ar", "16/9", "-ns", spp, "-bl", 10, "--no-show")
rgb, xyz = R.render()
t2 = time.time()
d = (xyz.view(np.uint32) != o.view(np.uint32)).any(axis=0)
print("oracle %.0f s, reference host build %.0f s; pixels where oracle != reference: %d" % (t1 - t0, t2 - t1, int(d.sum())), np.argwhere(d)[:5].tolist())
print("pixel", (px, py), "oracle", o[:, py, px], "reference", xyz[:, py, px])
np.save("/tmp/ref_full_scene%d.npy" % scene, xyz)

"""Where a k_wavefront task spends its cycles (probe build: make B=build_prof LIB=libsrt_prof.so EXTRA=-DSRT_PHASE_CLOCKS libsrt_prof.so;
run with SRT_LIB=.../libsrt_prof.so).  Warp 0 of blocks 0..3: SM cycles of the first task of every pass in four phases."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
h = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
spp = 64
sc = S.Scene(0)
cam = sc.camera(w, h); fb = S.FrameBuffer(w, h)
rm = S.RenderManager(sc, cam, fb); rm.init_renderer(10, spp)
rm.set_option(S.OPT_RANK, 0); rm.set_option(S.OPT_WORLD, world); rm.set_option(S.OPT_ROUNDS, 1); rm.set_option(S.OPT_PASS_LOG, 1)
if bs: rm.set_option(S.OPT_BLOCK_SLOTS, bs)
rm.init_device_params(0, 0)
for rep in range(2):
    rm.restart()
    while rm.step(): pass
print("world %d block_slots %d image %dx%d: render_ms %.3f" % (world, bs, w, h, rm.stats()["render_ms"]))
log = rm.pass_log()
names = ["regen", "lambert", "metal", "dielectric"]
for b in range(2):
    L = log[b]; C = log[b + 4]
    n = int((L[:, 0] != 0).sum())
    t = L[:n, 0].astype(np.int64); dt = (np.diff(t) & 0xFFFFFFFF) * 1e-3
    live = (L[:n, 1] + L[:n, 2] + (L[:n, 3] & 0xFFFF) + (L[:n, 3] >> 16)).astype(np.int64)
    kind = (C[:n, 3] >> 28).astype(np.int64)
    ph = np.stack([C[:n, 0], C[:n, 1], C[:n, 2], C[:n, 3] & 0x0FFFFFFF], 1).astype(np.float64)
    print("block %d: %d passes" % (b, n))
    for lo, hi in ((256, 257), (192, 256), (128, 192), (64, 128), (16, 64), (1, 16)):
        m = (live[:-1] >= lo) & (live[:-1] < hi)
        if not m.any(): continue
        print("  live in [%3d,%3d): %4d passes, period %5.2f us = %6.0f cycles" % (lo, hi, m.sum(), dt[m].mean(), dt[m].mean() * 1965))
        for k in range(4):
            mk = m & (kind[:-1] == k)
            if not mk.any(): continue
            p = ph[:-1][mk].mean(0)
            print("      first task %-10s (%4d): fetch %5.0f  regen/scatter %5.0f  closest hit %5.0f  store+push %5.0f  = %6.0f cycles" % (names[k], mk.sum(), p[0], p[1], p[2], p[3], p.sum()))

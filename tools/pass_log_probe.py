"""Per-pass timing of k_wavefront blocks: pass duration vs the number of live paths in the block (debug option SRT_OPT_PASS_LOG)."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sc = S.Scene(0)
w, h, spp = 1920, 1080, 64
cam = sc.camera(w, h); fb = S.FrameBuffer(w, h)
rm = S.RenderManager(sc, cam, fb); rm.init_renderer(10, spp)
rm.set_option(S.OPT_RANK, 0); rm.set_option(S.OPT_WORLD, world); rm.set_option(S.OPT_ROUNDS, rounds); rm.set_option(S.OPT_PASS_LOG, 1)
if bs: rm.set_option(S.OPT_BLOCK_SLOTS, bs)
rm.init_device_params(0, 0)
for rep in range(2):
    rm.restart()
    while rm.step(): pass
print("render_ms", rm.stats()["render_ms"])
log = rm.pass_log()
for b in range(2):
    L = log[b]; n = int((L[:, 0] != 0).sum())
    t = L[:n, 0].astype(np.int64); dt = np.diff(t) & 0xFFFFFFFF
    live = (L[:n, 1] + L[:n, 2] + (L[:n, 3] & 0xFFFF) + (L[:n, 3] >> 16)).astype(np.int64)
    q = [L[:n, 1].astype(np.int64), L[:n, 2].astype(np.int64), (L[:n, 3] & 0xFFFF).astype(np.int64), (L[:n, 3] >> 16).astype(np.int64)]
    tasks = sum(x // 32 for x in q) + (sum(x % 32 for x in q) + 31) // 32  # full warps per queue + the pooled remainders
    print("block %d: %d passes, total %.3f ms" % (b, n, dt.sum() * 1e-6))
    step = max(1, n // 40)
    for i in range(0, n - 1, step):
        j = min(n - 1, i + step)
        print("  passes %4d-%4d: live %6.1f (R %5.1f L %5.1f M %4.1f D %4.1f) warp-tasks %5.1f  us/pass %6.2f" % (i, j, live[i:j].mean(), L[i:j, 1].mean(), L[i:j, 2].mean(),
              (L[i:j, 3] & 0xFFFF).mean(), (L[i:j, 3] >> 16).mean(), tasks[i:j].mean(), dt[i:j].mean() * 1e-3))

import sys, pathlib, time
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
for n in ([int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else (1 << 20, 10 * (1 << 20))):
    t0 = time.time(); sc = S.Scene(soup=n, seed=1984); t1 = time.time()
    sc.rebuild_lbvh(3)
    rs = [sc.rebuild_lbvh(1) for _ in range(10)]
    med = {k: float(np.median([r[k] for r in rs])) for k in rs[0]}
    print("n=%d host+upload %.2fs  build ms:" % (n, t1 - t0), {k: round(v, 4) for k, v in med.items()}, " GB/s (256B/tri): %.0f" % (256.0 * n / med["total"] / 1e6), flush=True)
    # traversal: primary rays from the reference camera at 1920x1080 + one random bounce
    cam = sc.camera(1920, 1080).as_array()
    ys, xs = np.mgrid[0:1080, 0:1920]
    d = (cam[8:11][None, :] + xs.reshape(-1, 1) * cam[2:5][None, :] + ys.reshape(-1, 1) * cam[5:8][None, :] - cam[12:15][None, :]).astype(np.float32)
    o = np.tile(cam[12:15], (d.shape[0], 1)).astype(np.float32)
    t, tri, ms = sc.trace_rays(o, d)
    hit = tri >= 0
    print("   primary: %.2f ms, %.2f Grays/s, hit fraction %.3f" % (ms, d.shape[0] / ms / 1e6, hit.mean()), flush=True)
    rs_ = np.random.RandomState(1)
    p = o[hit] + t[hit, None] * d[hit]
    d2 = rs_.randn(p.shape[0], 3).astype(np.float32)
    t2, tri2, ms2 = sc.trace_rays(p + 1e-3 * d2, d2)
    print("   secondary (incoherent): %.2f ms, %.2f Grays/s, hit fraction %.3f" % (ms2, p.shape[0] / ms2 / 1e6, (tri2 >= 0).mean()), flush=True)
    if n <= (1 << 20):
        rgb, xyz, st = S.render(scene=sc, w=1920, h=1080, spp=4, bounce=10)
        print('   render 1080p 4spp depth 10 through the wavefront: %.1f ms, %.3f Gsamples/s, %.2f Grays/s, rays/sample %.2f' % (st['render_ms'], st['samples']/st['render_ms']/1e6, st['rays']/st['render_ms']/1e6, st['rays']/st['samples']), flush=True)
    del sc

"""How much of a closest-hit query is the end-of-kernel drain: the same incoherent ray distribution at 1/4 M ... 8 M rays.
Without a drain the time is proportional to the ray count; the intercept of the fit is the drain."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(soup=1 << 20, seed=1984)
rs = np.random.RandomState(1)
N = 1 << 23
o = (rs.rand(N, 3) * 555).astype(np.float32)
d = rs.randn(N, 3).astype(np.float32)
pts = []
for n in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    t, tri, ms = sc.trace_rays(o[:n], d[:n])
    pts.append((n, ms))
    print("%8d rays: %.3f ms, %.2f Grays/s" % (n, ms, n / ms / 1e6), flush=True)
a, b = np.polyfit([p[0] for p in pts], [p[1] for p in pts], 1)
print("fit: %.3f ms + n / (%.2f Grays/s)" % (b, 1e-6 / a))

import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(soup=1 << 20, seed=1984)
rs = np.random.RandomState(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
o = (rs.rand(n, 3) * 555).astype(np.float32)
d = rs.randn(n, 3).astype(np.float32)
t, tri, ms = sc.trace_rays(o, d)
print("incoherent rays from inside the soup: %.2f ms, %.2f Grays/s, hit %.3f" % (ms, n / ms / 1e6, (tri >= 0).mean()))

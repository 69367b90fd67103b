import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
sc = S.Scene(soup=n, seed=1984)
print(sc.rebuild_lbvh(3))

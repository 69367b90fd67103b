for L in "$@"; do echo "== $L"; for n in 1048576 10485760; do SRT_LIB=$PWD/cuda-spectral-ray-tracer_b200/$L python - $n <<'PY'
import sys, pathlib, numpy as np
sys.path.insert(0, "cuda-spectral-ray-tracer_b200")
import srt_b200 as S
n=int(sys.argv[1]); sc=S.Scene(soup=n, seed=1984); sc.rebuild_lbvh(3)
rs=[sc.rebuild_lbvh(1) for _ in range(12)]
print(n, {k: round(float(np.median([r[k] for r in rs])),4) for k in rs[0]})
PY
done; done

"""Which closest-hit path is right when the wide leaf and the LBVH walk disagree?  The oracle's brute force over all triangles decides."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200")); sys.path.insert(0, str(ROOT / "tests"))
import srt_b200 as S, oracle
import test_gpu_parity as T
rs = np.random.RandomState(11)
for scene in (0, 1, 2):
    sc = S.Scene(scene); osc = oracle.Scene(scene)
    O, D, kind = T.adversarial_rays(sc, rs)
    S.lib().srt_set_query_fp_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    t_walk, tri_walk, _ = sc.trace_rays(O, D)
    t_flat, tri_flat = sc.trace_rays_flat(O, D)
    bad = np.nonzero((tri_walk != tri_flat) | (t_walk.view(np.uint32) != t_flat.view(np.uint32)))[0]
    print("scene", scene, "rays", len(O), "differ", len(bad))
    for i in bad[:20]:
        h, out, ti = osc.brute_hit(O[i], D[i])
        print("  ray %d kind %s o %s d %s: walk tri %d t %.9g | flat tri %d t %.9g | brute hit %d tri %d t %.9g" % (i, kind[i], O[i], D[i], tri_walk[i], t_walk[i], tri_flat[i], t_flat[i], h, ti, out[1]))

#!/usr/bin/env python3
"""Golden vectors for the opt-in stratified pixel sampler, from the REAL reference host build
(oracle/_ref/libsrt_ref_ltr.so): renderer::get_ray_stratified_sample (rendering/rendering.cu:89-118) exists in the
reference but its kernel never calls it, so this known-answer set is the only thing the reference can say about it.
Runs only in the authoring container.  Output: tests/golden/ref_stratified_kat.npz"""
import pathlib, sys
import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import refhost  # noqa: E402


def main():
    R = refhost.RefHost("ltr")
    rs = np.random.RandomState(2024)
    R.open("-s", 0, "-xr", 400, "-ar", "16/9", "-ns", 16, "-bl", 10, "--no-show")
    ij, sxy, recip, rng_in, rng_out, out = [], [], [], [], [], []
    for k in range(96):
        n = int(rs.choice([1, 2, 3, 4, 8, 16]))
        i, j = int(rs.randint(0, 400)), int(rs.randint(0, 225))
        sx, sy = int(rs.randint(0, n)), int(rs.randint(0, n))
        rng = np.array([11 + k, 7 + k * 77, 3 + k * 13, 9 + k * 1001, 5 + k, 6 + k * 31], np.uint32)
        r = np.float32(1.0) / np.float32(n)
        o, ro = R.get_ray_stratified(i, j, sx, sy, float(r), rng)
        ij.append((i, j)); sxy.append((sx, sy)); recip.append(r); rng_in.append(rng); rng_out.append(ro); out.append(o)
    R.close()
    np.savez_compressed(HERE / "ref_stratified_kat.npz", ij=np.array(ij, np.int32), sxy=np.array(sxy, np.int32),
                        recip=np.array(recip, np.float32), rng_in=np.array(rng_in), rng_out=np.array(rng_out), out=np.array(out))
    print("wrote ref_stratified_kat.npz:", len(out), "rays")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Golden fixture for the physically meant Sellmeier mode (the library's srt_set_ref_compat(0), `srt_cli --physical`).

Source of truth: the REAL reference host-compiled with ONE extra token patched -- materials/material.cuh:67
`sellmeier_C[i] = b[i]` -> `= c[i]` (oracle/ref_host/build_ref.sh --draw-order ltr --physical ->
oracle/_ref/libsrt_ref_ltr_physical.so).  Runs only in the authoring container.  Output: tests/golden/ref_physical.npz
  scene{1,2}_rgb / _xyz   C1-sized renders (400x225, 8 spp, depth 10) of the Prism and the Different-Materials scene
  scene{1,2}_mats_f       material dumps (B and C now differ)
  silica_n / silica_lambda  sellmeier_index() of the reference for the fused-silica table (refraction/sellmeier.cuh:9-10),
                          which no reference scene uses, at 64 wavelengths
"""
import pathlib, sys
import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import refhost  # noqa: E402


def main():
    R = refhost.RefHost("ltr_physical")
    out = {}
    for scene in (1, 2):
        R.open("-s", scene, "-xr", 400, "-ar", "16/9", "-ns", 8, "-bl", 10, "--no-show")
        mf, mi = R.materials()
        rgb, xyz = R.render()
        out["scene%d_rgb" % scene] = rgb.astype(np.uint8)
        out["scene%d_xyz" % scene] = xyz
        out["scene%d_mats_f" % scene] = mf
        R.close()
    b = np.array([0.6961663, 0.4079426, 0.8974794], np.float32)
    c = np.array([0.0684043, 0.1162414, 9.896161], np.float32)
    lam = np.linspace(360.0, 830.0, 64).astype(np.float32)
    out["silica_lambda"] = lam
    out["silica_n"] = np.array([R.lib.srt_ref_sellmeier(b.ctypes.data, c.ctypes.data, float(l)) for l in lam], np.float32)
    np.savez_compressed(HERE / "ref_physical.npz", **out)
    print("wrote", HERE / "ref_physical.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generate the committed golden fixtures from the REAL reference, host-compiled
(oracle/_ref/libsrt_ref_ltr.so, built by oracle/ref_host/build_ref.sh from /root/reference).

Runs only in the authoring container (needs oracle/_ref).  Output: tests/golden/*.npz
  ref_scene{0,1,2}.npz   triangles (reference array order), materials, camera, BVH preorder,
                         C1 image (400x225, 8 spp, depth 10): sRGB uint8 + XYZ float32,
                         a 2x2-chunk render of a 96x54 image (RNG state carried across chunks)
  ref_kat.npz            XORWOW outputs, per-ray bvh::hit results, material::scatter results,
                         get_ray results, sellmeier / interp / XYZ / tonemap values
"""
import pathlib, sys
import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import refhost  # noqa: E402


def main():
    R = refhost.RefHost("ltr")
    rs = np.random.RandomState(1984)
    kat = {}
    for scene in (0, 1, 2):
        R.open("-s", scene, "-xr", 400, "-ar", "16/9", "-ns", 8, "-bl", 10, "--no-show")
        tf, ti = R.tris()
        mf, mi = R.materials()
        rgb, xyz = R.render()
        d = dict(tris_f=tf, tris_i=ti, mats_f=mf, mats_i=mi, camera=R.camera(), bvh_preorder=R.bvh_preorder(),
                 c1_rgb=rgb.astype(np.uint8), c1_xyz=xyz)
        # per-ray closest-hit KATs through the reference's own bvh::hit
        cam = R.camera()
        n = 400
        o = np.tile(cam[12:15], (n, 1)).astype(np.float32)
        dirs = np.zeros((n, 3), np.float32)
        px = rs.randint(0, 400, n); py = rs.randint(0, 225, n)
        for k in range(n):
            dirs[k] = cam[8:11] + px[k] * cam[2:5] + py[k] * cam[5:8] - cam[12:15]
        # secondary-like rays from inside the box
        o2 = (rs.rand(n, 3) * 500 + 25).astype(np.float32)
        d2 = (rs.rand(n, 3) * 2 - 1).astype(np.float32)
        O = np.concatenate([o, o2]); D = np.concatenate([dirs, d2])
        hits = np.zeros((2 * n, 10), np.float32)
        for k in range(2 * n):
            _, hits[k] = R.bvh_hit(O[k], D[k])
        d.update(kat_ray_o=O, kat_ray_d=D, kat_hits=hits)
        # scatter KATs: feed every hit to its material with a fresh RNG state
        sc_in, sc_out, sc_did, sc_rng_in, sc_rng_out, sc_mat, sc_rec = [], [], [], [], [], [], []
        for k in range(2 * n):
            if hits[k, 0] == 0:
                continue
            wl0 = 360.0 + 470.0 * rs.rand()
            wl = [wl0]
            for _ in range(6):
                l = np.float32(wl[-1]) + np.float32(470.0 / 7.0)
                if l > 830.0:
                    l = np.float32(360.0) + (l - np.float32(830.0))
                wl.append(float(l))
            ray_io = np.array(list(O[k]) + list(D[k]) + [7] + wl + list(rs.rand(7)), np.float32)
            rec = np.array(list(hits[k, 2:5]) + list(hits[k, 5:8]) + [hits[k, 1], hits[k, 8]], np.float32)
            raw, _ = R.xorwow(5000 + k, 1)
            rng = np.array([6615241 + k, 123456789 ^ k, 362436069 + 3 * k, 521288629 ^ (k << 3), 88675123 + k, 5783321 + 7 * k],
                           np.uint64).astype(np.uint32)
            did, ray_out, rng_out = R.scatter(int(hits[k, 9]), ray_io, rec, rng)
            sc_in.append(ray_io); sc_out.append(ray_out); sc_did.append(did); sc_rng_in.append(rng); sc_rng_out.append(rng_out)
            sc_mat.append(int(hits[k, 9])); sc_rec.append(rec)
        d.update(sc_in=np.array(sc_in), sc_out=np.array(sc_out), sc_did=np.array(sc_did, np.int32),
                 sc_rng_in=np.array(sc_rng_in), sc_rng_out=np.array(sc_rng_out), sc_mat=np.array(sc_mat, np.int32),
                 sc_rec=np.array(sc_rec))
        # camera rays
        gr_rng, gr_out, gr_rng_out, gr_ij = [], [], [], []
        for k in range(64):
            i, j = int(rs.randint(0, 400)), int(rs.randint(0, 225))
            rng = np.array([1 + k, 2 + k * 77, 3 + k * 13, 4 + k * 1001, 5 + k, 6 + k * 31], np.uint32)
            out, rng_out = R.get_ray(i, j, rng)
            gr_rng.append(rng); gr_out.append(out); gr_rng_out.append(rng_out); gr_ij.append((i, j))
        d.update(gr_rng=np.array(gr_rng), gr_out=np.array(gr_out), gr_rng_out=np.array(gr_rng_out), gr_ij=np.array(gr_ij, np.int32))
        R.close()
        # chunked small render: 96x54, 4 spp, 2x2 chunks of 48x27
        R.open("-s", scene, "-xr", 96, "-ar", "16/9", "-ns", 4, "-bl", 10, "-xc", 48, "-yc", 27, "--no-show")
        rgb_c, xyz_c = R.render()
        d.update(chunk_rgb=rgb_c.astype(np.uint8), chunk_xyz=xyz_c)
        R.close()
        # small single-chunk render used by the GPU smoke test
        R.open("-s", scene, "-xr", 96, "-ar", "16/9", "-ns", 4, "-bl", 10, "--no-show")
        rgb_s, xyz_s = R.render()
        d.update(small_rgb=rgb_s.astype(np.uint8), small_xyz=xyz_s)
        R.close()
        np.savez_compressed(HERE / ("ref_scene%d.npz" % scene), **d)
        print("scene", scene, "done; lit pixels", int((rgb.sum(0) > 0).sum()))

    # scalar KATs
    R.open("-s", 0, "-xr", 64, "-ar", "16/9", "-ns", 1, "-bl", 2, "--no-show")
    raw, uni = [], []
    for seed in (1984, 1985, 1984 + 448 * 15, 4242):
        r, u = R.xorwow(seed, 16)
        raw.append(r); uni.append(u)
    kat["xorwow_seeds"] = np.array([1984, 1985, 1984 + 448 * 15, 4242], np.uint32)
    kat["xorwow_raw"] = np.array(raw); kat["xorwow_uni"] = np.array(uni)
    lam = np.linspace(355, 835, 97).astype(np.float32)
    flint_b = np.array([1.34533359, 0.209073176, 0.937357162], np.float32)
    bk7_b = np.array([1.03961212, 0.231792344, 1.01046945], np.float32)
    bk7_c = np.array([6.00069867e-3, 2.00179144e-2, 1.03560653e2], np.float32)
    kat["lam"] = lam
    kat["sell_flint_bug"] = np.array([R.lib.srt_ref_sellmeier(flint_b.ctypes.data, flint_b.ctypes.data, float(l)) for l in lam], np.float32)
    kat["sell_bk7_true"] = np.array([R.lib.srt_ref_sellmeier(bk7_b.ctypes.data, bk7_c.ctypes.data, float(l)) for l in lam], np.float32)
    mf, _ = R.materials()
    red = np.ascontiguousarray(mf[0, 11:106])
    kat["interp_red"] = np.array([R.lib.srt_ref_spectrum_interp(red.ctypes.data, float(l)) for l in lam], np.float32)
    wl = (360 + 470 * rs.rand(32, 7)).astype(np.float32); pw = rs.rand(32, 7).astype(np.float32)
    nv = rs.randint(0, 8, 32).astype(np.int32)
    xyz = np.zeros((32, 3), np.float32); tm = np.zeros((32, 3), np.float32)
    xyz_in = (rs.rand(32, 3) * np.array([1.2, 1.2, 1.2]) - 0.05).astype(np.float32)
    for k in range(32):
        R.lib.srt_ref_spectrum_to_xyz(wl[k].ctypes.data, pw[k].ctypes.data, int(nv[k]), xyz[k].ctypes.data)
        R.lib.srt_ref_tonemap(xyz_in[k].ctypes.data, tm[k].ctypes.data)
    kat.update(xyz_wl=wl, xyz_pw=pw, xyz_nv=nv, xyz_out=xyz, tm_in=xyz_in, tm_out=tm)
    R.close()
    np.savez_compressed(HERE / "ref_kat.npz", **kat)
    print("kat done")


if __name__ == "__main__":
    main()

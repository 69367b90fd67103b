"""ctypes loader for oracle/_ref/libsrt_ref_<order>.so -- the REAL reference, host-compiled
(oracle/ref_host/).  Test infrastructure only."""
import ctypes as C
import os
import pathlib
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]


def lib_path(order="ltr"):
    """order: ltr | rtl (RNG draw order baked in), or ltr_physical (materials/material.cuh:67 fixed as well)"""
    return ROOT / "oracle" / "_ref" / ("libsrt_ref_%s.so" % order)


def available(order="ltr"):
    return lib_path(order).exists()


class RefHost:
    def __init__(self, order="ltr"):
        self.lib = C.CDLL(str(lib_path(order)))
        L = self.lib
        L.srt_ref_open.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.srt_ref_open.restype = C.c_int
        L.srt_ref_render.restype = C.c_int
        L.srt_ref_sellmeier.restype = C.c_float
        L.srt_ref_sellmeier.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.srt_ref_spectrum_interp.restype = C.c_float
        L.srt_ref_spectrum_interp.argtypes = [C.c_void_p, C.c_float]
        L.srt_ref_spectrum_to_xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.srt_ref_tonemap.argtypes = [C.c_void_p, C.c_void_p]
        L.srt_ref_color_spectrum.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
        L.srt_ref_prepare_color.argtypes = [C.c_float, C.c_float, C.c_float]

    def open(self, *args):
        argv = (C.c_char_p * len(args))(*[str(a).encode() for a in args])
        rc = self.lib.srt_ref_open(len(args), argv)
        if rc != 0:
            raise RuntimeError("srt_ref_open failed: %d" % rc)
        self.W, self.H = self.lib.srt_ref_width(), self.lib.srt_ref_height()

    def close(self):
        self.lib.srt_ref_close()

    def render(self):
        n = self.W * self.H
        r, g, b = (np.zeros(n, np.float32) for _ in range(3))
        xyz = np.zeros(3 * n, np.float32)
        rc = self.lib.srt_ref_render(r.ctypes, g.ctypes, b.ctypes, xyz.ctypes)
        assert rc == 0
        rgb = np.stack([r, g, b]).reshape(3, self.H, self.W)
        return rgb, xyz.reshape(3, self.H, self.W)

    def tris(self):
        n = self.lib.srt_ref_num_tris()
        f = np.zeros((n, 22), np.float32)
        iv = np.zeros((n, 3), np.int32)
        self.lib.srt_ref_get_tris(f.ctypes, iv.ctypes)
        return f, iv

    def materials(self):
        n = self.lib.srt_ref_num_materials()
        f = np.zeros((n, 108), np.float32)
        iv = np.zeros(n, np.int32)
        self.lib.srt_ref_get_materials(f.ctypes, iv.ctypes)
        return f, iv

    def camera(self):
        out = np.zeros(22, np.float32)
        self.lib.srt_ref_get_camera(out.ctypes)
        return out

    def bvh_preorder(self):
        n = self.lib.srt_ref_num_tris()
        out = np.zeros(2 * n + 2, np.int32)
        k = self.lib.srt_ref_bvh_preorder(out.ctypes)
        return out[:k].copy()

    def xorwow(self, seed, n):
        raw = np.zeros(n, np.uint32)
        uni = np.zeros(n, np.float32)
        self.lib.srt_ref_xorwow(C.c_uint(seed), n, raw.ctypes, uni.ctypes)
        return raw, uni

    def bvh_hit(self, o, d):
        out = np.zeros(10, np.float32)
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32)
        h = self.lib.srt_ref_bvh_hit(o.ctypes, d.ctypes, out.ctypes)
        return h, out

    def scatter(self, mat, ray_io, rec, rng):
        ray_io = np.array(ray_io, np.float32); rec = np.asarray(rec, np.float32); rng = np.array(rng, np.uint32)
        did = self.lib.srt_ref_scatter(mat, ray_io.ctypes, rec.ctypes, rng.ctypes)
        return did, ray_io, rng

    def get_ray(self, i, j, rng):
        rng = np.array(rng, np.uint32); out = np.zeros(13, np.float32)
        self.lib.srt_ref_get_ray(C.c_uint(i), C.c_uint(j), rng.ctypes, out.ctypes)
        return out, rng

    def get_ray_stratified(self, i, j, sx, sy, recip, rng):
        rng = np.array(rng, np.uint32); out = np.zeros(13, np.float32)
        self.lib.srt_ref_get_ray_stratified(C.c_uint(i), C.c_uint(j), C.c_uint(sx), C.c_uint(sy), C.c_float(recip), rng.ctypes, out.ctypes)
        return out, rng


def write_ppm(path, rgb):
    h, w = rgb.shape[1:]
    arr = np.clip(rgb, 0, 255).astype(np.uint8).transpose(1, 2, 0)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(arr.tobytes())

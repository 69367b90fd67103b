"""ctypes loader for oracle/libsrt_oracle.so (the C restatement).  Test infrastructure only."""
import ctypes as C
import pathlib
import subprocess
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = ROOT / "oracle" / "libsrt_oracle.so"


class OV3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class OTri(C.Structure):
    _fields_ = [("v", OV3 * 3), ("clockwise", C.c_int), ("aa_plane", C.c_int), ("mat", C.c_uint32),
                ("bb", C.c_float * 6), ("n", OV3), ("D", C.c_float)]


class OMat(C.Structure):
    _fields_ = [("type", C.c_int), ("col", OV3), ("fuzz", C.c_float), ("power", C.c_float),
                ("B", C.c_float * 3), ("C", C.c_float * 3), ("spec", C.c_float * 95)]


class OCam(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("du", OV3), ("dv", OV3), ("p00", OV3),
                ("defocus_angle", C.c_float), ("center", OV3), ("disk_u", OV3), ("disk_v", OV3),
                ("background", OV3)]


class OCounters(C.Structure):
    _names = ["samples", "rays", "box_tests", "tri_tests", "rng_draws", "interps", "scatter_lambert",
              "scatter_metal", "scatter_dielectric", "rejection_iters", "end_miss", "end_limit",
              "end_emissive", "end_absorbed", "nan_rays"]
    _fields_ = [(n, C.c_uint64) for n in _names]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in self._names}


def build():
    if not LIB.exists() or any(p.stat().st_mtime > LIB.stat().st_mtime
                               for p in (ROOT / "oracle").glob("*.[ch]")):
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "libsrt_oracle.so"],
                              stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.srt_oracle_scene_create.restype = C.c_void_p
        L.srt_oracle_scene_create.argtypes = [C.c_int]
        L.srt_oracle_scene_soup.restype = C.c_void_p
        L.srt_oracle_scene_soup.argtypes = [C.c_int, C.c_uint64]
        L.srt_oracle_scene_destroy.argtypes = [C.c_void_p]
        L.srt_oracle_scene_ntris.argtypes = [C.c_void_p]
        L.srt_oracle_scene_nmats.argtypes = [C.c_void_p]
        L.srt_oracle_scene_tris.restype = C.POINTER(OTri)
        L.srt_oracle_scene_tris.argtypes = [C.c_void_p]
        L.srt_oracle_scene_mats.restype = C.POINTER(OMat)
        L.srt_oracle_scene_mats.argtypes = [C.c_void_p]
        L.srt_oracle_scene_refbvh_preorder.argtypes = [C.c_void_p, C.c_void_p]
        L.srt_oracle_camera_default.argtypes = [C.c_int, C.c_int, C.POINTER(OCam)]
        L.srt_oracle_yres.argtypes = [C.c_int, C.c_float]
        L.srt_oracle_render.argtypes = [C.c_void_p, C.POINTER(OCam), C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.POINTER(OCounters), C.c_int]
        L.srt_oracle_render_tiles.argtypes = [C.c_void_p, C.POINTER(OCam), C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_void_p]
        L.srt_oracle_bvh_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.srt_oracle_brute_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.srt_oracle_scatter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.srt_oracle_get_ray.argtypes = [C.POINTER(OCam), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.srt_oracle_get_ray_stratified.argtypes = [C.POINTER(OCam), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
        L.srt_oracle_render_opts.argtypes = [C.c_void_p, C.POINTER(OCam)] + [C.c_int] * 5 + [C.c_void_p] * 3 + [C.c_int]
        L.srt_oracle_debug_pixel.argtypes = [C.c_void_p, C.POINTER(OCam), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.srt_oracle_set_physical_sellmeier.argtypes = [C.c_int]
        L.srt_oracle_glass.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.srt_oracle_sellmeier.restype = C.c_float
        L.srt_oracle_sellmeier.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.srt_oracle_spectrum_interp.restype = C.c_float
        L.srt_oracle_spectrum_interp.argtypes = [C.c_void_p, C.c_float]
        L.srt_oracle_rng_uniform.restype = C.c_float
        L.srt_oracle_spectrum_to_xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.srt_oracle_tonemap.argtypes = [C.c_void_p, C.c_void_p]
        L.srt_oracle_color_spectrum.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
        L.srt_oracle_morton30.restype = C.c_uint32
        L.srt_oracle_morton30.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.srt_oracle_lbvh_build.argtypes = [C.c_int] + [C.c_void_p] * 9
        L.srt_oracle_rgb2spec_cell.argtypes = [C.c_int] * 5 + [C.c_void_p]
        L.srt_oracle_rgb2spec_eval.argtypes = [C.c_void_p, C.c_void_p]
        L.srt_oracle_rgb2spec_scale.restype = C.c_float
        _lib = L
    return _lib


class Scene:
    def __init__(self, scene_id=None, soup=None, seed=1984, physical=False):
        """physical=True: dielectrics with the one-token fix of materials/material.cuh:67 (C[i] = c[i])"""
        L = lib()
        L.srt_oracle_set_physical_sellmeier(1 if physical else 0)
        try:
            self.h = L.srt_oracle_scene_soup(soup, seed) if soup is not None else L.srt_oracle_scene_create(scene_id)
        finally:
            L.srt_oracle_set_physical_sellmeier(0)
        self.ntris = L.srt_oracle_scene_ntris(self.h)
        self.nmats = L.srt_oracle_scene_nmats(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().srt_oracle_scene_destroy(self.h)
            self.h = None

    def tris(self):
        """(n,22) float32 + (n,3) int32 in the same layout as refhost.RefHost.tris()"""
        p = lib().srt_oracle_scene_tris(self.h)
        f = np.zeros((self.ntris, 22), np.float32)
        iv = np.zeros((self.ntris, 3), np.int32)
        for t in range(self.ntris):
            T = p[t]
            for k in range(3):
                f[t, 3 * k:3 * k + 3] = (T.v[k].x, T.v[k].y, T.v[k].z)
            f[t, 9:12] = (T.n.x, T.n.y, T.n.z)
            f[t, 12] = T.D
            f[t, 13:19] = list(T.bb)
            iv[t] = (T.clockwise, T.aa_plane, T.mat)
        return f, iv

    def materials(self):
        p = lib().srt_oracle_scene_mats(self.h)
        f = np.zeros((self.nmats, 108), np.float32)
        iv = np.zeros(self.nmats, np.int32)
        for m in range(self.nmats):
            M = p[m]
            f[m, 0:3] = (M.col.x, M.col.y, M.col.z)
            f[m, 3] = M.fuzz
            f[m, 4] = M.power
            f[m, 5:8] = list(M.B)
            f[m, 8:11] = list(M.C)
            f[m, 11:106] = list(M.spec)
            iv[m] = M.type
        return f, iv

    def refbvh_preorder(self):
        out = np.zeros(2 * self.ntris + 2, np.int32)
        k = lib().srt_oracle_scene_refbvh_preorder(self.h, out.ctypes.data)
        return out[:k].copy()

    def reforder(self):
        out = np.zeros(self.ntris, np.int32)
        lib().srt_oracle_scene_reforder(C.c_void_p(self.h), out.ctypes.data)
        return out

    def bvh_hit(self, o, d):
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32); out = np.zeros(10, np.float32)
        h = lib().srt_oracle_bvh_hit(self.h, o.ctypes.data, d.ctypes.data, out.ctypes.data)
        return h, out

    def brute_hit(self, o, d):
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32); out = np.zeros(10, np.float32)
        idx = C.c_int(-1)
        h = lib().srt_oracle_brute_hit(self.h, o.ctypes.data, d.ctypes.data, out.ctypes.data, C.byref(idx))
        return h, out, idx.value

    def scatter(self, mat, ray_io, rec, rng):
        ray_io = np.array(ray_io, np.float32); rec = np.asarray(rec, np.float32); rng = np.array(rng, np.uint32)
        did = lib().srt_oracle_scatter(self.h, mat, ray_io.ctypes.data, rec.ctypes.data, rng.ctypes.data)
        return did, ray_io, rng


def camera(w, h):
    cam = OCam()
    lib().srt_oracle_camera_default(w, h, C.byref(cam))
    return cam


def camera_make(w, h, vfov, lookfrom, lookat, vup, defocus_angle, focus_dist, background=(0, 0, 0)):
    cam = OCam()
    L = lib()
    L.srt_oracle_camera_make.argtypes = [C.c_int, C.c_int, C.c_float, OV3, OV3, OV3, C.c_float, C.c_float, OV3, C.POINTER(OCam)]
    L.srt_oracle_camera_make(w, h, vfov, OV3(*lookfrom), OV3(*lookat), OV3(*vup), defocus_angle, focus_dist, OV3(*background), C.byref(cam))
    return cam


def camera_array(cam):
    v = lambda a: [a.x, a.y, a.z]
    return np.array([cam.w, cam.h] + v(cam.du) + v(cam.dv) + v(cam.p00) + [cam.defocus_angle] + v(cam.center)
                    + v(cam.disk_u) + v(cam.disk_v), np.float32)


def render(scene, cam, spp, bounce=10, chunk_w=0, chunk_h=0, counters=False, nthreads=0, stratified=False):
    n = cam.w * cam.h
    rgb = np.zeros(3 * n, np.float32)
    xyz = np.zeros(3 * n, np.float32)
    cnt = OCounters()
    rc = lib().srt_oracle_render_opts(scene.h, C.byref(cam), spp, bounce, chunk_w, chunk_h, 1 if stratified else 0, rgb.ctypes.data,
                                      xyz.ctypes.data, C.cast(C.byref(cnt), C.c_void_p) if counters else None, nthreads)
    if rc != 0:
        raise ValueError("srt_oracle_render_opts failed: %d" % rc)
    out = (rgb.reshape(3, cam.h, cam.w), xyz.reshape(3, cam.h, cam.w))
    return out + (cnt.as_dict(),) if counters else out


def render_tiles(scene, cam, spp, bounce, tile_w, tile_h, rank, world):
    n = cam.w * cam.h
    xyz = np.zeros(3 * n, np.float32)
    lib().srt_oracle_render_tiles(scene.h, C.byref(cam), spp, bounce, tile_w, tile_h, rank, world, xyz.ctypes.data)
    return xyz.reshape(3, cam.h, cam.w)


def lbvh_build(leaf_boxes, centroids):
    n = leaf_boxes.shape[0]
    lb = np.ascontiguousarray(leaf_boxes, np.float32); ce = np.ascontiguousarray(centroids, np.float32)
    sb = np.zeros(6, np.float32); codes = np.zeros(n, np.uint32); sidx = np.zeros(n, np.uint32)
    left = np.zeros(max(n - 1, 1), np.int32); right = np.zeros(max(n - 1, 1), np.int32)
    parent = np.zeros(2 * n - 1, np.int32); nb = np.zeros((2 * n - 1, 6), np.float32)
    lib().srt_oracle_lbvh_build(n, lb.ctypes.data, ce.ctypes.data, sb.ctypes.data, codes.ctypes.data, sidx.ctypes.data,
                                left.ctypes.data, right.ctypes.data, parent.ctypes.data, nb.ctypes.data)
    return dict(scene_box=sb, codes=codes, sorted_idx=sidx, left=left[:n - 1], right=right[:n - 1], parent=parent,
                node_boxes=nb)

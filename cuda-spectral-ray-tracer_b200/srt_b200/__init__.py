"""ctypes binding of libsrt.so (include/srt.h) -- used by tests/, bench.py and __graft_entry__.py.

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
scene is created, this raises.  The Python classes mirror the reference's objects one to one
(param_manager, camera_builder, scene_manager, render_manager, frame_buffer)."""
import ctypes as C
import os
import pathlib

import numpy as np

PKG_DIR = pathlib.Path(__file__).resolve().parents[1]
LIB_PATH = pathlib.Path(os.environ.get("SRT_LIB", PKG_DIR / "libsrt.so"))  # SRT_LIB: A/B runs against another build of the library


class SrtError(RuntimeError):
    pass


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pixel_delta_u", Vec3), ("pixel_delta_v", Vec3),
                ("pixel00_loc", Vec3), ("defocus_angle", C.c_float), ("camera_center", Vec3), ("defocus_disk_u", Vec3),
                ("defocus_disk_v", Vec3), ("background", Vec3)]

    def as_array(self):
        v = lambda a: [a.x, a.y, a.z]
        return np.array([self.width, self.height] + v(self.pixel_delta_u) + v(self.pixel_delta_v) + v(self.pixel00_loc)
                        + [self.defocus_angle] + v(self.camera_center) + v(self.defocus_disk_u) + v(self.defocus_disk_v), np.float32)


class MaterialDesc(C.Structure):
    _fields_ = [("type", C.c_uint32), ("color", C.c_float * 3), ("fuzz", C.c_float), ("emission_power", C.c_float),
                ("sellmeier_b", C.c_float * 3), ("sellmeier_c", C.c_float * 3)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("wavefront_launches", C.c_uint64), ("rounds", C.c_uint64), ("render_ms", C.c_double), ("lbvh_ms", C.c_double),
                ("wavefront_ms", C.c_double), ("megakernel_ms", C.c_double), ("order_ms", C.c_double), ("other_ms", C.c_double),
                ("drain_ms", C.c_double), ("exchange_ms", C.c_double), ("film_out_ms", C.c_double)]


OPT_FP_MODE, OPT_PIPELINE, OPT_TILE_W, OPT_TILE_H, OPT_RANK, OPT_WORLD, OPT_KERNEL_TIMING, OPT_TRAVERSAL, OPT_BLOCK_SLOTS, OPT_BLOCK_THREADS, OPT_STRATIFIED, OPT_ROUNDS, OPT_L2_PERSIST, OPT_PASS_LOG, OPT_SCHED_FLAGS = 1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16, 17
MAT_LAMBERTIAN, MAT_METALLIC, MAT_DIELECTRIC, MAT_EMISSIVE = 0, 1, 2, 4

# every symbol include/srt.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGS = [
    ("srt_last_error", C.c_char_p, []),
    ("srt_kernel_launch_count", C.c_uint64, []),
    ("srt_device_count", C.c_int, []),
    ("srt_set_device", C.c_int, [C.c_int]),
    ("srt_params_instance", _P, []),
    ("srt_params_reset", None, []),
    ("srt_params_parse", None, [_P, C.c_int, C.POINTER(C.c_char_p)]),
    ("srt_params_scene_id", C.c_uint, [_P]), ("srt_params_xres", C.c_uint, [_P]), ("srt_params_yres", C.c_uint, [_P]),
    ("srt_params_ar", C.c_float, [_P]), ("srt_params_xcsize", C.c_uint, [_P]), ("srt_params_ycsize", C.c_uint, [_P]),
    ("srt_params_nsamples", C.c_uint, [_P]), ("srt_params_bounce_limit", C.c_uint, [_P]),
    ("srt_params_log_active", C.c_int, [_P]), ("srt_params_do_save", C.c_int, [_P]), ("srt_params_show_render", C.c_int, [_P]),
    ("srt_params_img_title", C.c_char_p, [_P]), ("srt_params_log_subdir", C.c_char_p, [_P]),
    ("srt_camera_builder_create", _P, []), ("srt_camera_builder_destroy", None, [_P]),
    ("srt_camera_builder_set_vfov", None, [_P, C.c_float]),
    ("srt_camera_builder_set_lookfrom", None, [_P, C.c_float, C.c_float, C.c_float]),
    ("srt_camera_builder_set_lookat", None, [_P, C.c_float, C.c_float, C.c_float]),
    ("srt_camera_builder_set_vup", None, [_P, C.c_float, C.c_float, C.c_float]),
    ("srt_camera_builder_set_defocus_angle", None, [_P, C.c_float]),
    ("srt_camera_builder_set_focus_dist", None, [_P, C.c_float]),
    ("srt_camera_builder_set_background", None, [_P, C.c_float, C.c_float, C.c_float]),
    ("srt_camera_builder_get_camera", C.c_int, [_P, C.POINTER(Camera)]),
    ("srt_camera_builder_get_camera_res", C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(Camera)]),
    ("srt_scene_create", _P, [C.c_uint]),
    ("srt_scene_create_soup", _P, [C.c_uint32, C.c_uint64]),
    ("srt_scene_create_mesh", _P, [_P, _P, C.c_uint32, C.POINTER(MaterialDesc), C.c_uint32]),
    ("srt_scene_create_obj", _P, [C.c_char_p, C.POINTER(MaterialDesc), C.c_uint32]),
    ("srt_scene_create_ply", _P, [C.c_char_p, C.POINTER(MaterialDesc), C.c_uint32]),
    ("srt_scene_destroy", None, [_P]),
    ("srt_scene_result", C.c_int, [_P, C.POINTER(C.c_char_p)]),
    ("srt_scene_camera", C.c_int, [_P, C.POINTER(Camera)]),
    ("srt_scene_camera_res", C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(Camera)]),
    ("srt_scene_num_tris", C.c_uint32, [_P]), ("srt_scene_num_materials", C.c_uint32, [_P]), ("srt_scene_num_units", C.c_uint32, [_P]),
    ("srt_set_ref_compat", None, [C.c_int]),
    ("srt_glass_coefficients", C.c_int, [C.c_int, _P, _P]),
    ("srt_scene_get_tris", C.c_int, [_P, _P, _P]), ("srt_scene_get_materials", C.c_int, [_P, _P, _P]),
    ("srt_scene_get_lbvh", C.c_int, [_P] * 8),
    ("srt_scene_rebuild_lbvh", C.c_int, [_P, C.c_int, _P]),
    ("srt_scene_trace_rays", C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, _P]),
    ("srt_set_query_fp_mode", None, [C.c_int]),
    ("srt_scene_trace_rays_flat", C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    ("srt_scene_trace_rays_counted", C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    ("srt_render_manager_create", _P, [_P, C.POINTER(Camera), _P, _P, _P]),
    ("srt_render_manager_destroy", None, [_P]),
    ("srt_rm_init_renderer", C.c_int, [_P, C.c_uint, C.c_uint]),
    ("srt_rm_init_device_params", C.c_int, [_P, C.c_uint, C.c_uint]),
    ("srt_rm_is_ready_to_render", C.c_int, [_P]), ("srt_rm_is_done", C.c_int, [_P]),
    ("srt_rm_im_width", C.c_uint, [_P]), ("srt_rm_im_height", C.c_uint, [_P]),
    ("srt_rm_step", C.c_int, [_P]), ("srt_rm_update_fb", C.c_int, [_P]),
    ("srt_rm_render_cycle", C.c_int, [_P]), ("srt_rm_end_render", C.c_int, [_P]), ("srt_rm_render_all", C.c_int, [_P]),
    ("srt_rm_set_option", C.c_int, [_P, C.c_int, C.c_int]),
    ("srt_rm_get_xyz", C.c_int, [_P, _P]),
    ("srt_rm_device_film", _P, [_P]),
    ("srt_rm_resolve_film", C.c_int, [_P]),
    ("srt_rm_restart", C.c_int, [_P]),
    ("srt_rm_get_stats", C.c_int, [_P, C.POINTER(Stats)]),
    ("srt_rm_get_pass_log", C.c_int, [_P, _P]),
    ("srt_comm_get_unique_id", C.c_int, [_P]),
    ("srt_comm_create", _P, [_P, C.c_int, C.c_int]),
    ("srt_comm_destroy", None, [_P]),
    ("srt_comm_rank", C.c_int, [_P]), ("srt_comm_world", C.c_int, [_P]),
    ("srt_comm_max_double", C.c_int, [_P, C.POINTER(C.c_double)]),
    ("srt_nccl_version", C.c_int, []),
    ("srt_rm_set_comm", C.c_int, [_P, _P]),
    ("srt_rm_exchange_film", C.c_int, [_P]),
    ("srt_rm_film_checksum", C.c_int, [_P, C.POINTER(C.c_uint64)]),
    ("srt_trim_caches", None, []),
    ("srt_measure_fp32_tflops", C.c_double, []),
    ("srt_measure_copy_gbs", C.c_double, [C.c_uint32]),
    ("srt_measure_l2_read_gbs", C.c_double, []),
    ("srt_write_ppm", C.c_int, [C.c_char_p, _P, _P, _P, C.c_uint32, C.c_uint32]),
    ("srt_write_bmp", C.c_int, [C.c_char_p, _P, _P, _P, C.c_uint32, C.c_uint32]),
]
SYMBOLS = [s[0] for s in _SIGS]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SrtError("libsrt.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(str(LIB_PATH))
        for name, res, args in _SIGS:
            if "SRT_LIB" in os.environ and not hasattr(L, name):
                continue  # A/B runs against an older build
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise SrtError(lib().srt_last_error().decode() or ("srt error %d" % rc))


def kernel_launch_count():
    return int(lib().srt_kernel_launch_count())


class Params:
    """param_manager singleton (reference io/params.h:226-315)."""

    def __init__(self, *argv, reset=True):
        L = lib()
        if reset:
            L.srt_params_reset()
        self.h = L.srt_params_instance()
        if argv:
            self.parse(*argv)

    def parse(self, *argv):
        args = [b"srt"] + [str(a).encode() for a in argv]
        arr = (C.c_char_p * len(args))(*args)
        lib().srt_params_parse(self.h, len(args), arr)

    def __getattr__(self, name):
        f = getattr(lib(), "srt_params_" + name)
        v = f(self.h)
        return v.decode() if isinstance(v, bytes) else v


class CameraBuilder:
    """camera_builder (reference rendering/camera_builder.cuh)."""

    def __init__(self):
        self.h = lib().srt_camera_builder_create()

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().srt_camera_builder_destroy(self.h)
            self.h = None

    def _set(self, name, *v):
        getattr(lib(), "srt_camera_builder_set_" + name)(self.h, *v)
        return self

    def setVfov(self, v): return self._set("vfov", v)
    def setLookfrom(self, x, y, z): return self._set("lookfrom", x, y, z)
    def setLookat(self, x, y, z): return self._set("lookat", x, y, z)
    def setVup(self, x, y, z): return self._set("vup", x, y, z)
    def setDefocusAngle(self, v): return self._set("defocus_angle", v)
    def setFocusDist(self, v): return self._set("focus_dist", v)
    def setBackground(self, r, g, b): return self._set("background", r, g, b)

    def getCamera(self, w=None, h=None):
        cam = Camera()
        if w is None:
            _check(lib().srt_camera_builder_get_camera(self.h, C.byref(cam)))
        else:
            _check(lib().srt_camera_builder_get_camera_res(self.h, w, h, C.byref(cam)))
        return cam


class Scene:
    """scene_manager (reference scene/scene.cuh:103-176): builds triangles, materials and the device LBVH."""

    def __init__(self, scene_id=None, soup=None, seed=1984, mesh=None, obj=None, ply=None, host_only=False):
        """host_only=True keeps a scene whose device upload failed (no GPU): only the host-side
        dumps (tris(), materials(), camera()) work on it; rendering raises."""
        L = lib()
        if soup is not None:
            self.h = L.srt_scene_create_soup(soup, seed)
        elif mesh is not None:
            verts, mat_idx, mats = mesh
            verts = np.ascontiguousarray(verts, np.float32)
            mat_idx = np.ascontiguousarray(mat_idx, np.uint32)
            arr = (MaterialDesc * len(mats))(*mats)
            self.h = L.srt_scene_create_mesh(verts.ctypes.data, mat_idx.ctypes.data, len(mat_idx), arr, len(mats))
        elif obj is not None:
            path, mats = obj
            arr = (MaterialDesc * len(mats))(*mats)
            self.h = L.srt_scene_create_obj(str(path).encode(), arr, len(mats))
        elif ply is not None:
            path, mats = ply
            arr = (MaterialDesc * len(mats))(*mats)
            self.h = L.srt_scene_create_ply(str(path).encode(), arr, len(mats))
        else:
            self.h = L.srt_scene_create(scene_id)
        if not self.h:
            raise SrtError(L.srt_last_error().decode())
        msg = C.c_char_p()
        self.ok = bool(L.srt_scene_result(self.h, C.byref(msg)))
        self.msg = (msg.value or b"").decode()
        if not self.ok and not host_only:
            raise SrtError(self.msg)
        self.ntris = L.srt_scene_num_tris(self.h)
        self.nmats = L.srt_scene_num_materials(self.h)
        self.nunits = L.srt_scene_num_units(self.h) if hasattr(L, "srt_scene_num_units") else -1

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:  # `lib` is gone when the interpreter shuts down
            lib().srt_scene_destroy(self.h)
            self.h = None

    def camera(self, w=None, h=None):
        cam = Camera()
        if w is None:
            _check(lib().srt_scene_camera(self.h, C.byref(cam)))
        else:
            _check(lib().srt_scene_camera_res(self.h, w, h, C.byref(cam)))
        return cam

    def tris(self):
        f = np.zeros((self.ntris, 22), np.float32); iv = np.zeros((self.ntris, 3), np.int32)
        _check(lib().srt_scene_get_tris(self.h, f.ctypes.data, iv.ctypes.data))
        return f, iv

    def materials(self):
        f = np.zeros((self.nmats, 108), np.float32); iv = np.zeros(self.nmats, np.int32)
        _check(lib().srt_scene_get_materials(self.h, f.ctypes.data, iv.ctypes.data))
        return f, iv

    def lbvh(self):
        n = self.ntris
        d = dict(scene_box=np.zeros(6, np.float32), codes=np.zeros(n, np.uint32), sorted_idx=np.zeros(n, np.uint32),
                 left=np.zeros(max(n - 1, 1), np.int32), right=np.zeros(max(n - 1, 1), np.int32),
                 parent=np.zeros(max(2 * n - 1, 1), np.int32), node_boxes=np.zeros((max(2 * n - 1, 1), 6), np.float32))
        _check(lib().srt_scene_get_lbvh(self.h, d["codes"].ctypes.data, d["sorted_idx"].ctypes.data, d["left"].ctypes.data,
                                        d["right"].ctypes.data, d["parent"].ctypes.data, d["node_boxes"].ctypes.data,
                                        d["scene_box"].ctypes.data))
        d["left"] = d["left"][:max(n - 1, 0)]; d["right"] = d["right"][:max(n - 1, 0)]
        return d

    def rebuild_lbvh(self, repeats=1):
        ms = np.zeros(5, np.float32)
        _check(lib().srt_scene_rebuild_lbvh(self.h, repeats, ms.ctypes.data))
        return dict(total=float(ms[0]), bounds_morton=float(ms[1]), sort=float(ms[2]), tree=float(ms[3]), collapse=float(ms[4]))

    def trace_rays(self, o, d, counted=False):
        o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        t = np.zeros(n, np.float32); tri = np.zeros(n, np.int32); ms = C.c_float(0)
        if counted:
            v = np.zeros(2, np.uint64)
            _check(lib().srt_scene_trace_rays_counted(self.h, n, o.ctypes.data, d.ctypes.data, t.ctypes.data, tri.ctypes.data, C.byref(ms), v.ctypes.data))
            return t, tri, ms.value, (int(v[0]), int(v[1]))
        _check(lib().srt_scene_trace_rays(self.h, n, o.ctypes.data, d.ctypes.data, t.ctypes.data, tri.ctypes.data, C.byref(ms)))
        return t, tri, ms.value


def _trace_rays_flat(self, o, d):
    o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
    n = o.shape[0]
    t = np.zeros(n, np.float32); tri = np.zeros(n, np.int32)
    _check(lib().srt_scene_trace_rays_flat(self.h, n, o.ctypes.data, d.ctypes.data, t.ctypes.data, tri.ctypes.data))
    return t, tri


Scene.trace_rays_flat = _trace_rays_flat


class Comm:
    """NCCL communicator owned by libsrt (one process per GPU).  Rank 0 calls Comm.unique_id() and hands the 128 bytes
    to the other ranks out of band; then every rank calls Comm(id, rank, world) after srt_set_device."""

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        _check(lib().srt_comm_get_unique_id(buf))
        return bytes(buf)

    def __init__(self, uid, rank, world):
        buf = (C.c_ubyte * 128)(*uid)
        self.h = lib().srt_comm_create(buf, rank, world)
        if not self.h:
            raise SrtError(lib().srt_last_error().decode())
        self.rank, self.world = rank, world

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib().srt_comm_destroy(self.h)
            self.h = None

    def max(self, v):
        d = C.c_double(v)
        _check(lib().srt_comm_max_double(self.h, C.byref(d)))
        return d.value

    def barrier(self):
        self.max(0.0)


class FrameBuffer:
    """frame_buffer (reference rendering/frame_buffer.cuh): planar float32 R, G, B, raster order, 0..255."""

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.r = np.zeros(w * h, np.float32); self.g = np.zeros(w * h, np.float32); self.b = np.zeros(w * h, np.float32)

    def rgb(self):
        return np.stack([self.r, self.g, self.b]).reshape(3, self.h, self.w)


class RenderManager:
    """render_manager (reference rendering/render_manager.cuh:37-173)."""

    def __init__(self, scene, cam, fb):
        self.scene, self.cam, self.fb = scene, cam, fb  # borrowed, keep alive
        self.h = lib().srt_render_manager_create(scene.h, C.byref(cam), fb.r.ctypes.data, fb.g.ctypes.data, fb.b.ctypes.data)
        if not self.h:
            raise SrtError(lib().srt_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().srt_render_manager_destroy(self.h)
            self.h = None

    def init_renderer(self, bounce_limit, spp): _check(lib().srt_rm_init_renderer(self.h, bounce_limit, spp))
    def init_device_params(self, cw=0, ch=0): _check(lib().srt_rm_init_device_params(self.h, cw, ch))
    def set_option(self, opt, value): _check(lib().srt_rm_set_option(self.h, opt, value))
    def isReadyToRender(self): return bool(lib().srt_rm_is_ready_to_render(self.h))
    def isDone(self): return bool(lib().srt_rm_is_done(self.h))
    def getImWidth(self): return lib().srt_rm_im_width(self.h)
    def getImHeight(self): return lib().srt_rm_im_height(self.h)

    def step(self):
        rc = lib().srt_rm_step(self.h)
        if rc < 0:
            raise SrtError(lib().srt_last_error().decode())
        return bool(rc)

    def update_fb(self):
        rc = lib().srt_rm_update_fb(self.h)
        if rc < 0:
            raise SrtError(lib().srt_last_error().decode())
        return bool(rc)

    def render_cycle(self): _check(lib().srt_rm_render_cycle(self.h))
    def end_render(self): _check(lib().srt_rm_end_render(self.h))
    def render_all(self): _check(lib().srt_rm_render_all(self.h))
    def device_film(self): return lib().srt_rm_device_film(self.h)

    def set_comm(self, comm):
        self.comm = comm  # borrowed, keep alive
        _check(lib().srt_rm_set_comm(self.h, comm.h if comm is not None else None))

    def exchange_film(self): _check(lib().srt_rm_exchange_film(self.h))

    def film_checksum(self):
        v = C.c_uint64(0)
        _check(lib().srt_rm_film_checksum(self.h, C.byref(v)))
        return int(v.value)

    def resolve_film(self): _check(lib().srt_rm_resolve_film(self.h))
    def restart(self): _check(lib().srt_rm_restart(self.h))

    def xyz(self):
        out = np.zeros(3 * self.cam.width * self.cam.height, np.float32)
        _check(lib().srt_rm_get_xyz(self.h, out.ctypes.data))
        return out.reshape(3, self.cam.height, self.cam.width)

    def pass_log(self):
        out = np.zeros((8, 8192, 4), np.uint32)
        _check(lib().srt_rm_get_pass_log(self.h, out.ctypes.data))
        return out

    def stats(self):
        s = Stats()
        _check(lib().srt_rm_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}


def render(scene_id=0, w=400, h=225, spp=8, bounce=10, chunk=(0, 0), strict=False, pipeline=0, scene=None, tiles=None,
           kernel_timing=False, traversal=None, block_slots=None, block_threads=None, stratified=False, rounds=None, sched_flags=None):
    """One-call helper: returns (rgb[3,h,w] float32 0..255, xyz[3,h,w] float32, stats dict)."""
    sc = scene if scene is not None else Scene(scene_id)
    cam = sc.camera(w, h)
    fb = FrameBuffer(w, h)
    rm = RenderManager(sc, cam, fb)
    rm.init_renderer(bounce, spp)
    rm.set_option(OPT_FP_MODE, 1 if strict else 0)
    rm.set_option(OPT_PIPELINE, pipeline)
    if kernel_timing:
        rm.set_option(OPT_KERNEL_TIMING, 1)
    if traversal is not None:
        rm.set_option(OPT_TRAVERSAL, traversal)
    if block_threads is not None:
        rm.set_option(OPT_BLOCK_THREADS, block_threads)
    if stratified:
        rm.set_option(OPT_STRATIFIED, 1)
    if rounds is not None:
        rm.set_option(OPT_ROUNDS, rounds)
    if sched_flags is not None:
        rm.set_option(OPT_SCHED_FLAGS, sched_flags)
    if tiles is not None:
        tw, th, rank, world = tiles
        rm.set_option(OPT_TILE_W, tw); rm.set_option(OPT_TILE_H, th); rm.set_option(OPT_RANK, rank); rm.set_option(OPT_WORLD, world)
    if block_slots is not None:  # paths in flight per wavefront block (power of two)
        rm.set_option(OPT_BLOCK_SLOTS, block_slots)
    rm.init_device_params(*chunk)
    rm.render_all()
    return fb.rgb().copy(), rm.xyz(), rm.stats()

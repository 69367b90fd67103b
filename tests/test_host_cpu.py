"""CPU-only tests (no GPU): the C-ABI library loads and exports every symbol of include/srt.h, the
host-side mirror of the reference objects (params, camera_builder, scene construction, material
spectra) is bit-exact against the golden vectors produced by the real reference, and the multi-rank
tile split is a partition (world_size-2 gloo)."""
import ctypes as C
import os
import pathlib
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def srt():
    import srt_b200

    return srt_b200


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_every_declared_symbol_is_exported(srt):
    header = (ROOT / "include" / "srt.h").read_text()
    declared = set(re.findall(r"\b(srt_[a-z0-9_]+)\s*\(", header))
    declared -= {"srt_material_desc", "srt_stats", "srt_vec3", "srt_camera"}
    L = C.CDLL(str(srt.LIB_PATH))
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(srt.SYMBOLS) == declared  # the Python binding covers the whole header


def test_no_cpu_fallback(srt):
    """without a GPU the product must fail loudly, not fall back to anything"""
    if srt.lib().srt_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(srt.SrtError, match="no CUDA device"):
        srt.Scene(0)
    # and nothing under the package imports or links the oracle
    for p in (ROOT / "cuda-spectral-ray-tracer_b200").rglob("*"):
        if p.suffix in (".py", ".cpp", ".cu", ".cuh", ".h", ".hpp") or p.name == "Makefile":
            txt = p.read_text(errors="ignore")
            uses = re.search(r'#include\s*[<"][^>"]*oracle|libsrt_oracle|lsrt_oracle|import\s+oracle|from\s+oracle|srt_oracle_\w+\s*\(', txt)
            assert not uses, (p, uses.group(0))


def test_params_mirror_reference_semantics(srt):
    p = srt.Params()  # defaults: io/params.h:197-222
    assert (p.scene_id, p.xres, p.yres, p.nsamples, p.bounce_limit) == (0, 600, 600, 500, 10)
    assert p.xcsize == 600 and p.ycsize == 600 and p.show_render == 1 and p.do_save == 0 and p.log_active == 0
    assert p.img_title == "Cornell Box"
    p = srt.Params("-s", 1, "-xr", 400, "-ar", "16/9", "-ns", 8, "-bl", 7, "--no-show", "--save", "--do-log", "-t", "my title", "-lsub", "x")
    assert (p.scene_id, p.xres, p.yres, p.nsamples, p.bounce_limit) == (1, 400, 225, 8, 7)
    assert p.show_render == 0 and p.do_save == 1 and p.log_active == 1 and p.img_title == "my title" and p.log_subdir == "x"
    # float32 uint(W / (16.f/9.f)) gives exactly 225 / 1080 / 2160
    for w, h in ((400, 225), (1920, 1080), (3840, 2160)):
        assert srt.Params("-xr", w, "-ar", "16/9").yres == h
    # chunk defaults follow each other (params.h:53-63)
    p = srt.Params("-xr", 100, "-xc", 32)
    assert p.xcsize == 32 and p.ycsize == 32
    p = srt.Params("-xr", 100, "-yc", 20)
    assert p.xcsize == 20 and p.ycsize == 20
    # bad values keep the previous value; a flag without a value and unknown flags are reported and skipped
    p = srt.Params("-ns", "abc", "-xr", "12x", "--bogus", "-bl")
    assert p.nsamples == 500 and p.xres == 12 and p.bounce_limit == 10
    # the singleton persists between parses unless reset (param_manager::getInstance)
    srt.Params("-ns", 3)
    assert srt.Params(reset=False).nsamples == 3


@pytest.mark.parametrize("scene", [0, 1, 2])
def test_host_scene_matches_reference(srt, golden, scene):
    g = golden["ref_scene%d" % scene]
    import oracle

    sc = srt.Scene(scene, host_only=True)
    f, iv = sc.tris()
    order = oracle.Scene(scene).reforder()  # the reference permutes its triangle array while building its own BVH
    assert np.array_equal(bits(g["tris_f"]), bits(f[order]))
    assert np.array_equal(g["tris_i"], iv[order])
    mf, mi = sc.materials()
    assert np.array_equal(bits(g["mats_f"]), bits(mf))  # includes the on-demand Jakob-Hanika cells and the C := B quirk
    assert np.array_equal(g["mats_i"], mi)
    cam = sc.camera(400, 225)
    assert np.array_equal(bits(g["camera"][:21]), bits(cam.as_array()))
    srt.Params("-xr", 400, "-ar", "16/9")
    assert np.array_equal(bits(sc.camera().as_array()), bits(cam.as_array()))  # getCamPtr(): resolution from the params singleton


def test_camera_builder(srt):
    import oracle

    b = srt.CameraBuilder().setVfov(40).setLookfrom(278, 278, -800).setLookat(278, 278, 0).setVup(0, 1, 0).setDefocusAngle(0).setFocusDist(10).setBackground(0, 0, 0)
    cam = b.getCamera(1920, 1080)
    assert np.array_equal(bits(cam.as_array()), bits(oracle.camera_array(oracle.camera(1920, 1080))))
    d = srt.CameraBuilder().getCamera(64, 64)  # defaults of camera_builder.cuh:64-70
    assert d.camera_center.z == -1.0 and d.width == 64


def test_physical_sellmeier_switch(srt):
    L = srt.lib()
    try:
        L.srt_set_ref_compat(0)
        mf, _ = srt.Scene(1, host_only=True).materials()
        assert np.allclose(mf[2, 8:11], [0.00997743871, 0.0470450767, 111.886764])
    finally:
        L.srt_set_ref_compat(1)
    mf, _ = srt.Scene(1, host_only=True).materials()
    assert np.array_equal(mf[2, 8:11], mf[2, 5:8])  # reference quirk F4: C := B


def test_physical_mode_materials_match_the_patched_reference(srt):
    """srt_set_ref_compat(0): host materials equal those of the real reference built with materials/material.cuh:67 fixed
    (tests/golden/ref_physical.npz), and srt_glass_coefficients carries the three tables of refraction/sellmeier.cuh:6-13"""
    import oracle

    g = np.load(ROOT / "tests" / "golden" / "ref_physical.npz")
    L = srt.lib()
    try:
        L.srt_set_ref_compat(0)
        for scene in (1, 2):
            mf, _ = srt.Scene(scene, host_only=True).materials()
            assert np.array_equal(bits(mf), bits(g["scene%d_mats_f" % scene]))
    finally:
        L.srt_set_ref_compat(1)
    for which in (0, 1, 2):
        b = np.zeros(3, np.float32); c = np.zeros(3, np.float32); ob = np.zeros(3, np.float32); oc = np.zeros(3, np.float32)
        assert L.srt_glass_coefficients(which, b.ctypes.data, c.ctypes.data) == 0
        oracle.lib().srt_oracle_glass(which, ob.ctypes.data, oc.ctypes.data)
        assert np.array_equal(b, ob) and np.array_equal(c, oc)
    assert L.srt_glass_coefficients(7, b.ctypes.data, c.ctypes.data) != 0


def build_shim_driver(out_dir):
    """compiles tests/shim/shim_main.cpp -- the reference's main.cpp:16-133 call sequence on the class names of include/srt_shim.h --
    and links it against libsrt.so"""
    exe = pathlib.Path(out_dir) / "shim_main"
    pkg = ROOT / "cuda-spectral-ray-tracer_b200"
    out = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + str(ROOT / "include"), str(ROOT / "tests" / "shim" / "shim_main.cpp"), "-o", str(exe),
                          "-L" + str(pkg), "-lsrt", "-Wl,-rpath," + str(pkg)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    return exe


def test_integration_shim_compiles_and_links(srt, tmp_path):
    """INTEGRATION.md section 2 is real code: include/srt_shim.h + a main.cpp-shaped driver build warning-free against the C-ABI;
    without a GPU the driver stops at the scene with the library's no-device error (no CPU fallback)"""
    exe = build_shim_driver(tmp_path)
    if srt.lib().srt_device_count() == 0:
        out = subprocess.run([str(exe), "-s", "1", "-xr", "32", "--no-show"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 1 and "no CUDA device" in out.stderr


def test_soup_and_mesh_host_side(srt):
    import oracle

    a, ai = srt.Scene(soup=500, seed=99, host_only=True).tris()
    b, bi = oracle.Scene(soup=500, seed=99).tris()
    assert np.array_equal(bits(a), bits(b)) and np.array_equal(ai, bi)
    v = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    m = srt.MaterialDesc(type=srt.MAT_LAMBERTIAN, color=(0.5, 0.5, 0.5), fuzz=1.0)
    sc = srt.Scene(mesh=(v, np.zeros(1, np.uint32), [m]), host_only=True)
    f, iv = sc.tris()
    assert list(f[0, 9:12]) == [0, 0, 1] and iv[0, 1] == 1  # normal +z, plane XY


def test_obj_loader_and_image_writers(srt, tmp_path):
    obj = tmp_path / "q.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\nf -4//1 -3//1 -2//1\n")
    m = srt.MaterialDesc(type=srt.MAT_LAMBERTIAN, color=(0.5, 0.5, 0.5), fuzz=1.0)
    sc = srt.Scene(obj=(obj, [m]), host_only=True)
    assert sc.ntris == 3
    r = np.arange(6, dtype=np.float32) * 40; g = r[::-1].copy(); b = np.full(6, 300, np.float32)
    assert srt.lib().srt_write_ppm(str(tmp_path / "a.ppm").encode(), r.ctypes.data, g.ctypes.data, b.ctypes.data, 3, 2) == 0
    data = (tmp_path / "a.ppm").read_bytes()
    assert data.startswith(b"P6\n3 2\n255\n") and data[11:14] == bytes([0, 200, 255])
    assert srt.lib().srt_write_bmp(str(tmp_path / "a.bmp").encode(), r.ctypes.data, g.ctypes.data, b.ctypes.data, 3, 2) == 0
    bmp = (tmp_path / "a.bmp").read_bytes()
    assert bmp[:2] == b"BM" and len(bmp) == 54 + 12 * 2


def test_ply_loader(srt, tmp_path):
    """ascii and binary_little_endian PLY give the same triangles as the equivalent OBJ (fan triangulation, extra
    vertex properties and double coordinates skipped / converted); broken files are refused with a message"""
    import struct

    m = srt.MaterialDesc(type=srt.MAT_LAMBERTIAN, color=(0.5, 0.5, 0.5), fuzz=1.0)
    obj = tmp_path / "q.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0.5\nv 0 1 0\nf 1 2 3 4\nf 2 3 4\n")
    want = srt.Scene(obj=(obj, [m]), host_only=True).tris()[0]
    a = tmp_path / "a.ply"
    a.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nelement face 2\nproperty list uchar int vertex_indices\nend_header\n"
                 "0 0 0 255\n1 0 0 255\n1 1 0.5 0\n0 1 0 7\n4 0 1 2 3\n3 1 2 3\n")
    sa = srt.Scene(ply=(a, [m]), host_only=True)
    assert sa.ntris == 3 and np.array_equal(sa.tris()[0], want)
    b = tmp_path / "b.ply"
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex 4\nproperty double x\nproperty double y\nproperty double z\n"
           "element face 2\nproperty list uchar uint vertex_index\nend_header\n").encode()
    body = b"".join(struct.pack("<3d", *v) for v in [(0, 0, 0), (1, 0, 0), (1, 1, 0.5), (0, 1, 0)])
    body += struct.pack("<B4I", 4, 0, 1, 2, 3) + struct.pack("<B3I", 3, 1, 2, 3)
    b.write_bytes(hdr + body)
    sb = srt.Scene(ply=(b, [m]), host_only=True)
    assert sb.ntris == 3 and np.array_equal(sb.tris()[0], want)
    bad = tmp_path / "bad.ply"
    bad.write_bytes(hdr + body[:40])
    with pytest.raises(srt.SrtError):
        srt.Scene(ply=(bad, [m]), host_only=True)
    bad.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nelement face 1\n"
                   "property list uchar int vertex_indices\nend_header\n0 0 0\n3 0 1 2\n")
    with pytest.raises(srt.SrtError):
        srt.Scene(ply=(bad, [m]), host_only=True)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver times next to ours) runs on the host alone and prints ONE JSON line
    with the contract's keys; under a multi-rank launch only rank 0 prints"""
    import json
    import subprocess

    env = dict(os.environ)
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-spp", "1", "--workload", "c1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in j, key
    assert j["impl"] == "reference" and j["value"] > 0 and j["unit"] == "samples/s" and j["config"]["workload"] == "c1"
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
S = oracle.Scene(1); cam = oracle.camera(64, 36)
part = oracle.render_tiles(S, cam, 2, 10, 16, 16, rank, world)      # this rank's tiles only (XYZ sums)
film = torch.from_numpy(part.copy())
owned = torch.from_numpy((part.sum(0) != 0).astype(np.int32))
dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)                      # the NCCL film reduce, on gloo
dist.all_reduce(owned, op=dist.ReduceOp.SUM)
if rank == 0:
    full = oracle.render(S, cam, 2, 10)[1] * 2                      # mean -> sum
    assert owned.max().item() <= 1, "tiles overlap between ranks"
    assert np.allclose(film.numpy(), full, rtol=0, atol=1e-6), "sum of rank films != single-process film"
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_two_rank_tile_split_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": str(ROOT)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]

"""A/B timing of two builds of libsrt.so on the same GPU box: SRT_LIB=<path> python tools/ab_probe.py <worlds> <scene> <sched flags,...>  (C2 frame, rank 0 of world N);
prints a checksum of the XYZ film bits so that builds / schedulers can be compared for identity"""
import sys, pathlib, zlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
scene = int(sys.argv[2]) if len(sys.argv) > 2 else 0
flags = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
sc = S.Scene(scene)
for world in worlds:
    for fl in flags:
        ts = []
        for rep in range(5):
            rgb, xyz, st = S.render(scene=sc, w=1920, h=1080, spp=64, bounce=10, tiles=(0, 0, 0, world), sched_flags=fl)
            ts.append(st["render_ms"])
        print("%s scene %d world %d flags %d: best %.2f median %.2f ms  film crc %08x" % (S.LIB_PATH.name, scene, world, fl, min(ts), sorted(ts)[2], zlib.crc32(xyz.tobytes())), flush=True)

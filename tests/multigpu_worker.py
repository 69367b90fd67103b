"""One rank of a multi-GPU render through libsrt's own NCCL communicator (spawned by tests/test_gpu_parity.py and usable by hand):
    python tests/multigpu_worker.py RANK WORLD RENDEZVOUS_DIR [scene w h spp strict]
Rank 0 writes the NCCL unique id into RENDEZVOUS_DIR/id.bin, the others wait for it.  Every rank renders its tiles, all exchange the
film (srt_rm_exchange_film), and each writes rank<k>.npz with the film checksum; rank 0 adds its frame buffer and the reduced XYZ."""
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S  # noqa: E402


def main():
    rank, world, rdv = int(sys.argv[1]), int(sys.argv[2]), pathlib.Path(sys.argv[3])
    scene, w, h, spp, strict = ([int(x) for x in sys.argv[4:9]] + [0, 320, 180, 16, 1][len(sys.argv) - 4:])[:5]
    L = S.lib()
    ndev = L.srt_device_count()
    if L.srt_set_device(rank % ndev) != 0:
        raise SystemExit(L.srt_last_error().decode())
    idf = rdv / "id.bin"
    if rank == 0:
        uid = S.Comm.unique_id()
        tmp = rdv / "id.tmp"
        tmp.write_bytes(uid)
        os.replace(tmp, idf)
    else:
        t0 = time.time()
        while not idf.exists():
            if time.time() - t0 > 120:
                raise SystemExit("rendezvous timed out")
            time.sleep(0.02)
        uid = idf.read_bytes()
    comm = S.Comm(uid, rank, world)
    sc = S.Scene(scene)
    fb = S.FrameBuffer(w, h)
    rm = S.RenderManager(sc, sc.camera(w, h), fb)
    rm.init_renderer(10, spp)
    rm.set_option(S.OPT_FP_MODE, strict)
    rm.set_comm(comm)
    rm.init_device_params(0, 0)
    while rm.step():
        pass
    rm.exchange_film()
    crc = rm.film_checksum()
    xyz = rm.xyz()  # collective after an exchange
    st = rm.stats()
    out = dict(crc=np.uint64(crc), samples=np.uint64(st["samples"]), exchange_ms=st["exchange_ms"])
    if rank == 0:
        out.update(rgb=fb.rgb(), xyz=xyz)
    np.savez(rdv / ("rank%d.npz" % rank), **out)
    comm.barrier()
    del rm
    comm.close()


if __name__ == "__main__":
    main()

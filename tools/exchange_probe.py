"""Phases of srt_rm_exchange_film and of the renderer set-up (run under torchrun with SRT_TRACE=1 for the host-side trace)."""
import os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
import srt_b200 as S
S.lib().srt_set_device(lr)
comm = None
if world > 1:
    box = [S.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = S.Comm(box[0], rank, world)
w, h, spp = 1920, 1080, 64
for rep in range(4):
    t0 = time.perf_counter()
    sc = S.Scene(0); fb = S.FrameBuffer(w, h); rm = S.RenderManager(sc, sc.camera(w, h), fb); rm.init_renderer(10, spp)
    if comm: rm.set_comm(comm)
    t1 = time.perf_counter()
    rm.init_device_params(0, 0)
    t2 = time.perf_counter()
    while rm.step(): pass
    t3 = time.perf_counter()
    if comm: rm.exchange_film()
    else: rm.resolve_film()
    t4 = time.perf_counter()
    st = rm.stats()
    print("rank %d rep %d: scene+rm %.2f  init_device_params %.2f  render %.2f (device %.2f)  film out %.2f (device exchange %.2f) ms" % (
        rank, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["render_ms"], (t4 - t3) * 1e3, st["exchange_ms"]), flush=True)
    del rm

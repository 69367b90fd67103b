"""N-GPU probe (torchrun): how long do the candidate film exchanges take when all ranks enter together?
Times dist.reduce, dist.all_reduce and reduce_scatter on float32 tensors of the 1080p (24.9 MB) and 4K (99.5 MB) film."""
import os, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def timeit(name, fn, n_iter=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ts = []
    for _ in range(n_iter):
        dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-34s median over iterations, max over ranks: %.3f ms" % (name, t.item()), flush=True)
for label, n in (("1080p film 24.9 MB", 3 * 1920 * 1080), ("4K film 99.5 MB", 3 * 3840 * 2160)):
    n -= n % world
    x = torch.ones(n, device="cuda")
    out = torch.empty(n // world, device="cuda")
    timeit(label + " reduce->0", lambda: dist.reduce(x, dst=0, op=dist.ReduceOp.SUM))
    timeit(label + " all_reduce", lambda: dist.all_reduce(x, op=dist.ReduceOp.SUM))
    timeit(label + " reduce_scatter", lambda: dist.reduce_scatter_tensor(out, x, op=dist.ReduceOp.SUM))
dist.destroy_process_group()

"""NCCL transport / bandwidth probe: which transport the ranks use (P2P over NVLink or shared host memory) and what a
large send/recv and all-to-all achieve.  torchrun --nproc-per-node N tools/nccl_probe.py"""
import os, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 27  # 1 GiB of doubles
a = torch.ones(n, dtype=torch.float64, device="cuda")
b = torch.empty_like(a)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def sr():
    ops = [dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
    for r in dist.batch_isend_irecv(ops): r.wait()
t = timed(sr)
if rank == 0: print(f"ring send/recv 1 GiB: {n*8/t/1e9:.1f} GB/s per direction", flush=True)
t = timed(lambda: dist.all_to_all_single(b, a))
if rank == 0: print(f"all_to_all 1 GiB per rank: {n*8*(world-1)/world/t/1e9:.1f} GB/s out per rank", flush=True)
t = timed(lambda: dist.all_reduce(a))
if rank == 0: print(f"all_reduce 1 GiB: busbw {2*(world-1)/world*n*8/t/1e9:.1f} GB/s", flush=True)
dist.destroy_process_group()

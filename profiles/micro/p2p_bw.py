"""NVLink peer-memory micro-benchmark (torchrun, 2+ GPUs): bandwidth of torch copies into / out of a peer's symmetric
buffer.  Usage: python -m torch.distributed.run --nproc-per-node 2 profiles/micro/p2p_bw.py"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = dist.group.WORLD.group_name
try: symm.enable_symm_mem_for_group(name)
except Exception: pass
n = 16 * 1024 * 1024   # floats = 64 MB
buf = symm.empty((n,), dtype=torch.float32, device=dev)
hdl = symm.rendezvous(buf, name)
src = torch.rand(n, device=dev)
peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
dist.barrier()
t_push = timeit(lambda: peer.copy_(src))
dist.barrier()
t_pull = timeit(lambda: src.copy_(peer))
dist.barrier()
t_loc = timeit(lambda: buf.copy_(src))
idx = torch.randperm(n // 16, device=dev)[: n // 32]
rows_src = src.view(-1, 16)
def scatter_rows(): peer.view(-1, 16)[idx] = rows_src[idx]
t_sc = timeit(scatter_rows)
print(f"rank {rank}: push 64MB {t_push:.3f} ms = {64e-3*1.048576/t_push:.0f} GB/s | pull {t_pull:.3f} ms = {64e-3*1.048576/t_pull:.0f} GB/s | "
      f"local {t_loc:.3f} ms | scatter 0.5M random 64B rows to peer {t_sc:.3f} ms = {0.5*1.048576*64e-3/t_sc:.0f} GB/s", flush=True)
import subprocess
if rank == 0: print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
dist.destroy_process_group()

"""Probe: does this box give torch symmetric memory with NVLS multicast?  torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: world {h.world_size} multicast_ptr {h.multicast_ptr:#x} buffer_ptrs {[hex(p) for p in h.buffer_ptrs]} backend {getattr(h, 'get_backend', lambda: '?')() if False else ''}", flush=True)
t.fill_(rank + 1.0)
h.barrier()
peer = h.get_buffer((rank + 1) % h.world_size, (4,), torch.float32)
print(f"rank {rank}: peer value {peer[:2].tolist()}", flush=True)
h.barrier()
dist.destroy_process_group()

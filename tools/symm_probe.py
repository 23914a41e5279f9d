"""Development probe: does torch's symmetric memory (peer pointers, NVSwitch multicast) work on this box?
torchrun --nproc-per-node N tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

def _mc_support(idx):
    try:
        from torch._C._autograd import DeviceType
        return symm_mem._SymmetricMemory.has_multicast_support(DeviceType.CUDA, idx)
    except Exception as e:
        return repr(e)


rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm_mem.empty(64 << 20, dtype=torch.float32, device=dev)      # 256 MB
    t.fill_(rank + 1)
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; world", h.world_size, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:4],
          "multicast support", _mc_support(rank), "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None,
          "signal pad", h.signal_pad_size, flush=True)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", peer[:2].tolist(), flush=True)
    # a slice of a symmetric tensor: offset within the allocation
    v = t[1024:2048]
    print(rank, "view ptr - base", v.data_ptr() - t.data_ptr(), flush=True)
    h.barrier()
    # timing of NCCL all-reduce of 538 MB for comparison
    x = torch.ones(134_544_384, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    print(rank, "NCCL all_reduce 538 MB:", (time.perf_counter() - t0) / 5 * 1e3, "ms", flush=True)
except Exception as e:
    import traceback
    traceback.print_exc()
    print(rank, "FAILED", repr(e), flush=True)
dist.destroy_process_group()

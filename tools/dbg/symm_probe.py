"""Probe: does this box give torch symmetric memory with an NVLS multicast mapping? (run under torchrun)"""
import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    print(f"[rank {rank}] buffer_ptrs={[hex(p) for p in h.buffer_ptrs]} multicast_ptr={hex(h.multicast_ptr)} "
          f"signal_pads={len(h.signal_pad_ptrs)} data_ptr={hex(t.data_ptr())}", flush=True)
    t.fill_(rank + 1.0)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (8,), torch.float32)
    print(f"[rank {rank}] peer value {peer[:2].tolist()}", flush=True)
    h.barrier()
except Exception as e:
    print(f"[rank {rank}] symmetric memory failed: {type(e).__name__}: {e}", flush=True)
dist.destroy_process_group()

"""Probe (2+ GPUs, torchrun): does torch symmetric memory expose an NVLS multicast pointer on this box?"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1.0)
h = symm.rendezvous(t, dist.group.WORLD)
print("rank", rank, "world", h.world_size, "multicast_ptr", hex(int(getattr(h, "multicast_ptr", 0) or 0)), "buffer_ptrs", [hex(int(p)) for p in h.buffer_ptrs][:3],
      "has_multicast_support", getattr(symm, "has_multicast_support", lambda *a: "n/a")("cuda", local) if hasattr(symm, "has_multicast_support") else "n/a", flush=True)
try:
    from torch._C._distributed_c10d import _SymmetricMemory
    print("rank", rank, "_SymmetricMemory.has_multicast_support:", _SymmetricMemory.has_multicast_support(torch.device("cuda").type if False else dist.distributed_c10d.DeviceType.CUDA if hasattr(dist.distributed_c10d, "DeviceType") else 0, local), flush=True)
except Exception as e:
    print("rank", rank, "has_multicast_support query failed:", repr(e)[:200], flush=True)
dist.barrier(); torch.cuda.synchronize()
if int(getattr(h, "multicast_ptr", 0) or 0):
    # built-in op that uses multimem.ld_reduce: one-shot all-reduce over the multicast address
    try:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print("rank", rank, "multimem_all_reduce_ ->", float(t[0]), flush=True)
    except Exception as e:
        print("rank", rank, "multimem_all_reduce_ failed:", repr(e)[:300], flush=True)
dist.destroy_process_group()

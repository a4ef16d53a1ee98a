"""torchrun probe: is NVLS multicast memory available through torch's symmetric memory on this box?"""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    try:
        sup = type(hdl).has_multicast_support(torch._C._autograd.DeviceType.CUDA, local)
    except Exception as e:
        sup = "?(%r)" % (e,)
    print("rank", rank, "multicast support", sup, "mc_ptr", hex(hdl.multicast_ptr), "local", hex(t.data_ptr()), flush=True)
    t.zero_()
    hdl.barrier()
    # every rank adds (rank + 1) through plain peer pointers as a sanity check of the P2P mapping
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    peer.add_(rank + 1.0)
    hdl.barrier()
    print("rank", rank, "value after peer add", float(t[0]), "expected", float(((rank - 1) % world) + 1), flush=True)
except Exception as e:
    print("rank", rank, "ERR", repr(e)[:500], flush=True)
dist.destroy_process_group()

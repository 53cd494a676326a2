"""Experiment: does torch symmetric memory (P2P-mapped peer buffers) work on this box?  torchrun --nproc-per-node 2."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
try:
    import torch.distributed._symmetric_memory as symm_mem
    N = 1_000_000
    t = symm_mem.empty(N * 3, dtype=torch.float32, device=f"cuda:{rank}")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (N, 3), torch.float32)
    torch.cuda.synchronize()
    print(rank, "peer buffer mean", float(pb.mean()), "ptr", hex(pb.data_ptr()), "world", hdl.world_size, flush=True)
    # bandwidth of a plain peer read
    out = torch.empty_like(pb)
    for _ in range(3): out.copy_(pb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out.copy_(pb)
    e1.record(); torch.cuda.synchronize()
    print(rank, "peer read GB/s", 10 * N * 12 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
    hdl.barrier()
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", repr(e), flush=True)
dist.destroy_process_group()

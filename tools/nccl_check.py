"""All-reduce bandwidth of the data-parallel gradient exchange (364 MB of bf16 gradients, in place, 64 MB chunks)."""
import os, sys, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(182_000_000, dtype=torch.bfloat16, device="cuda")
for chunk in (182_000_000, 32 << 20):
    for it in range(4):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in range(0, x.numel(), chunk):
            dist.all_reduce(x[o:o + chunk], op=dist.ReduceOp.AVG)
        e1.record(); torch.cuda.synchronize()
        if dist.get_rank() == 0 and it == 3:
            ms = e0.elapsed_time(e1)
            print("all-reduce 364 MB in chunks of %d elements: %.2f ms  (%.1f GB/s algorithmic)" % (chunk, ms, 0.364 / ms * 1e3), file=sys.stderr)
dist.destroy_process_group()

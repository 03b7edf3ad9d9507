"""All-reduce time of the data-parallel gradient exchange (364 MB of bf16 gradients, in place) by chunk size and reduce op:
    torchrun --nproc-per-node 8 tools/nccl_check.py"""
import os, sys, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(182_000_000, dtype=torch.bfloat16, device="cuda")
for op_name, op in (("AVG", dist.ReduceOp.AVG), ("SUM", dist.ReduceOp.SUM)):
    for chunk in (182_000_000, 64 << 20, 32 << 20, 8 << 20):
        for asyn in (False, True):
            best = 1e9
            for it in range(5):
                x.fill_(1.0)
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                works = []
                for o in range(0, x.numel(), chunk):
                    w = dist.all_reduce(x[o:o + chunk], op=op, async_op=asyn)
                    if asyn:
                        works.append(w)
                for w in works:
                    w.wait()
                e1.record(); torch.cuda.synchronize()
                if it:
                    best = min(best, e0.elapsed_time(e1))
            if dist.get_rank() == 0:
                print("all-reduce 364 MB bf16 %s, chunks of %9d elements, async_op=%d: %.2f ms  (%.0f GB/s algorithmic)" % (
                    op_name, chunk, asyn, best, 0.364 / best * 1e3), flush=True)
dist.destroy_process_group()

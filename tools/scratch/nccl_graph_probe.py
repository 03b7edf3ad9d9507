import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.full((1 << 20,), float(dist.get_rank() + 1), device="cuda", dtype=torch.bfloat16)
y = torch.zeros(1 << 20, device="cuda")
dist.all_reduce(x, op=dist.ReduceOp.AVG)        # communicator up before capture
torch.cuda.synchronize()
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    x.mul_(2.0)
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    y.add_(1.0)                                   # overlaps with the collective
    cur.wait_stream(side)
    y.add_(x.float())
x.fill_(float(dist.get_rank() + 1)); y.zero_()
g.replay(); torch.cuda.synchronize()
print("rank", dist.get_rank(), "x", float(x[0]), "y", float(y[0]), "(expect x 3.0, y 4.0 with 2 ranks)", flush=True)
x.fill_(float(dist.get_rank() + 3)); y.zero_()
g.replay(); torch.cuda.synchronize()
print("rank", dist.get_rank(), "x", float(x[0]), "y", float(y[0]), "(expect x 7.0, y 8.0)", flush=True)
dist.destroy_process_group()

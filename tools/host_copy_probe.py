"""How much host<->device bandwidth do N ranks of one box get TOGETHER?  (what bounds the e2e leg of bench.py at 8 ranks)
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_copy_probe.py
Every rank copies 2 GiB pinned host -> device and 2 GiB device -> pinned host at the same time (two streams), first
rank 0 alone (the others idle), then all ranks together."""
import json
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = 2 << 30
h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
d_dn = torch.zeros(n, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(active, both=True, reps=3):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        s_up.wait_event(e0); s_dn.wait_event(e0)
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
            if both:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return n * reps / (float(t.item()) * 1e-3) / 1e9


run(True)   # warm-up
solo_up = run(rank == 0, both=False)
solo = run(rank == 0)
every_up = run(True, both=False)
every = run(True)
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_copy": n,
                      "one_rank_alone_h2d_only_gbs": solo_up, "one_rank_alone_each_way_gbs": solo,
                      "all_ranks_h2d_only_gbs_per_rank": every_up, "all_ranks_h2d_only_gbs_total": every_up * world,
                      "all_ranks_each_way_gbs_per_rank": every, "all_ranks_each_way_gbs_total": every * world}))
dist.destroy_process_group()

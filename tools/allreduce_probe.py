#!/usr/bin/env python
"""Times the per-minibatch collective of the training path in isolation: all-reduce of lambda_len+4 doubles (cfg4: 891 214) over NCCL.
Run under torchrun."""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n, dt in [(891214, torch.float64), (891214, torch.float32), (10191, torch.float64)]:
    g = torch.ones(n, dtype=dt, device="cuda")
    for _ in range(5): dist.all_reduce(g)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_reduce(g)
    e1.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0: print(f"all_reduce {n} x {dt}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
dist.destroy_process_group()

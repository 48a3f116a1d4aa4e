"""all-reduce timing of the mapping step's gradient buffer (47 MB f32) under torchrun"""
import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (86256, 12_000_000):
    x = torch.ones(n, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): dist.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0:
        ms = a.elapsed_time(b) / 20
        print(f"all_reduce {n*4/1e6:.1f} MB x{dist.get_world_size()}: {ms:.3f} ms  algbw {n*4/ms/1e6:.1f} GB/s", flush=True)
dist.destroy_process_group()

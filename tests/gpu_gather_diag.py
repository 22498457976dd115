"""torchrun target (N ranks, NCCL): what does the per-step all-gather of the output latents cost, and why?
Times, with CUDA events on the compute stream and MAX over ranks: (a) K engine steps alone, (b) K x (step + synchronous all-gather),
(c) K all-gathers alone (no compute), (d) K x (step + all-gather) with an explicit barrier-free idle gap.  Run it under different NCCL
settings (NCCL_MAX_NCHANNELS, NCCL_PROTO, ...) from the shell; prints one JSON line on rank 0.  Evidence for DESIGN section 5."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "img2img-turbo_b200"))
import bench  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K, B, S, dt = 20, 8, 512, torch.bfloat16
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    from _host import build_text_stack
    text_stack = build_text_stack(1024)
    wl = bench.Workload("pix2pix", False, dt, B, S, rank, text_stack)
gathered = torch.empty(world * B, 4, S // 8, S // 8, device="cuda", dtype=dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def loop(fn, k=K, do_flush=True):
    for _ in range(3):
        if do_flush:
            flush.zero_()
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        if do_flush:
            flush.zero_()
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / k], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(t.item(), 4)


def step_gather():
    wl.step()
    dist.all_gather_into_tensor(gathered, wl.lat)


def per_step(fn, k=40):
    """per-step durations (ms) of k steps on this rank, CUDA events around every step (flush excluded)"""
    for _ in range(3):
        flush.zero_(); fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    for a_, b_ in ev:
        flush.zero_()
        a_.record(); fn(); b_.record()
    torch.cuda.synchronize()
    t = torch.tensor([a_.elapsed_time(b_) for a_, b_ in ev], device="cuda")
    allt = torch.empty(world, k, device="cuda")
    dist.all_gather_into_tensor(allt, t)
    return allt.cpu()


res = {"world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_") and k != "NCCL_DEBUG"}}
res["ms_step_only"] = loop(wl.step)
res["ms_step_plus_sync_gather"] = loop(step_gather)
res["ms_gather_only_back_to_back"] = loop(lambda: dist.all_gather_into_tensor(gathered, wl.lat), k=200, do_flush=False)
res["ms_flush_only"] = loop(lambda: None, k=50)
res["ms_step_only_again"] = loop(wl.step)
if os.environ.get("I2IT_DIAG_PER_STEP"):
    free = per_step(wl.step)                      # ranks run freely: what would lock-step cost if only the jitter mattered?
    lock = per_step(step_gather)
    res["per_step_free"] = {"mean_per_rank": [round(v, 3) for v in free.mean(1).tolist()], "std_per_rank": [round(v, 3) for v in free.std(1).tolist()],
                            "min": round(free.min().item(), 3), "max": round(free.max().item(), 3),
                            "mean_of_max_over_ranks": round(free.max(0).values.mean().item(), 3)}
    res["per_step_lockstep"] = {"mean_per_rank": [round(v, 3) for v in lock.mean(1).tolist()], "std_per_rank": [round(v, 3) for v in lock.std(1).tolist()],
                                "min": round(lock.min().item(), 3), "max": round(lock.max().item(), 3)}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()

#!/usr/bin/env python
"""bench.py — 512x512 images/s of the one-step image-translation path on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own (CPU) implementation of the path, host cores

One "step" = one forward of the hot path (VAE.encode -> UNet(t=999) -> DDPM x0 -> VAE.decode) over one per-GPU
batch of synthetic 512x512 inputs (pix2pix-turbo edge_to_image, bf16, batch 8 per GPU,
random-init weights, LoRA folded).  N > 1 shards images across ranks (weak scaling, no data-path collective)
plus ONE NCCL all-gather of the output latents per step.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "img2img-turbo_b200"))

FLOPS_PER_IMAGE = 4467.6e9          # 2*MAC over every conv / linear / attention at 512x512
CPU_THREADS = os.cpu_count()
WORKLOAD = "pix2pix-turbo edge_to_image bf16 batch=8/GPU 512x512 (BASELINE config #2)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1407.3), d.get("hbm_gbs", 6576.4), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def synthetic_inputs(B, size, cross_dim, dtype, device, seed_offset=0):
    """BASELINE.md config #1/#2 inputs: canny-like {0,1} control image, randn text embedding, randn posterior eps."""
    g = torch.Generator().manual_seed(1 + seed_offset)
    c_t = (torch.rand(B, 1, size, size, generator=g) < 0.08).float().expand(-1, 3, -1, -1).contiguous()
    text = torch.randn(1, 77, cross_dim, generator=torch.Generator().manual_seed(2))
    eps = torch.randn(B, 4, size // 8, size // 8, generator=torch.Generator().manual_seed(3 + seed_offset))
    return (c_t.to(dtype).to(device) if device else c_t.to(dtype), text.to(dtype).to(device) if device else text.to(dtype),
            eps.to(dtype).to(device) if device else eps.to(dtype))


def cpu_oracle_images_per_s(sd, size, steps, warmup, cross_dim, cfg):
    """The reference's algorithm on the host cores: oracle/ (fp32 restatement of the diffusers path), B=1 per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    c_t, text, eps = synthetic_inputs(1, size, cross_dim, torch.float32, None)
    # pick the host thread count that is actually fastest (oversubscribed many-core hosts are slower with every thread)
    cs, ts, es = synthetic_inputs(1, 128, cross_dim, torch.float32, None)
    best = (1e30, os.cpu_count())
    with torch.no_grad():
        for nt in sorted({os.cpu_count(), 64, 32, 16}):
            if nt > os.cpu_count():
                continue
            torch.set_num_threads(nt)
            O.pix2pix_forward(sd, cs, ts, es, cfg)
            t0 = time.time()
            O.pix2pix_forward(sd, cs, ts, es, cfg)
            best = min(best, (time.time() - t0, nt))
    torch.set_num_threads(best[1])
    global CPU_THREADS
    CPU_THREADS = best[1]
    with torch.no_grad():
        for _ in range(warmup):
            O.pix2pix_forward(sd, c_t, text, eps, cfg)
        t0 = time.time()
        for _ in range(steps):
            O.pix2pix_forward(sd, c_t, text, eps, cfg)
        dt = (time.time() - t0) / steps
    return 1.0 / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  diffusers/peft are not installable here
    (not in /opt/wheelhouse, no network), so this times oracle/ — the restatement of the same algorithm —
    with every host thread.  One step = one 512x512 image (a bounded sample of the batch-8 workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import weights as W
    sd = W.make_state_dict("pix2pix", W.SD_TURBO, seed=0)
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    ips, sec = cpu_oracle_images_per_s(sd, args.size, steps, warmup, 1024, W.SD_TURBO)
    cores = CPU_THREADS
    line = {"impl": "reference", "metric": "512x512 images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: fp32, batch 1 per step (bounded sample), all host threads"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} x one {args.size}x{args.size} image, oracle fp32 (steps/warmup capped at 3/1)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp16"])
    ap.add_argument("--model", default="pix2pix", choices=["pix2pix", "cyclegan"],
                    help="pix2pix = BASELINE config #2 (default, the headline); cyclegan = config #3 (day_to_night a2b, fp16, batch 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="", help="write the per-launch timing table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import i2it
    import weights as W
    from pix2pix_turbo import Pix2Pix_Turbo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.dtype is None:
        args.dtype = "bf16" if args.model == "pix2pix" else "fp16"
    if args.model == "cyclegan" and args.batch == 8:
        args.batch = 16
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    B, S, K, Wm = args.batch, args.size, args.steps, max(args.warmup, 3)
    workload = WORKLOAD if args.model == "pix2pix" else "cyclegan-turbo day_to_night (a2b) fp16 batch=16/GPU 512x512 (BASELINE config #3)"

    # ---- model through the public (reference-compatible) API: random init, LoRA folded at load ----
    torch.manual_seed(0)
    prompt = "a synthetic benchmark prompt"
    c_t, text, eps = synthetic_inputs(B, S, 1024, dt, "cuda", seed_offset=rank)
    if args.model == "pix2pix":
        model = Pix2Pix_Turbo()                  # pretrained_name=None, pretrained_path=None -> random init (reference :131)
        model.set_eval()
        model.to(dt)
        call = lambda x, **kw: model(x, prompt, **kw)
    else:
        from cyclegan_turbo import CycleGAN_Turbo
        model = CycleGAN_Turbo(synthetic_caption="driving in the night", synthetic_direction="a2b")
        model.eval()
        model.to(dt)
        prompt = model.caption
        c_t = (c_t.float() * 0 + torch.rand(c_t.shape, generator=torch.Generator().manual_seed(1 + rank)).to("cuda") * 2 - 1).to(dt)
        call = lambda x, **kw: model(x, **kw)
    with torch.no_grad():
        out = call(c_t, eps=eps)                 # builds engine + plan, caches the prompt embedding
    eng = model._get_engine()
    text_emb = model._encode_text(prompt)
    torch.cuda.synchronize()

    lat = torch.empty(B, 4, S // 8, S // 8, device="cuda", dtype=dt)
    gathered = torch.empty(world * B, 4, S // 8, S // 8, device="cuda", dtype=dt) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step():
        flush.zero_()                                                  # L2 flush between iterations
        eng.forward(c_t, text_emb, eps, out=out, out_latent=lat)
        if world > 1:
            dist.all_gather_into_tensor(gathered, lat)                 # the single collective: output latents over NVLink

    for _ in range(Wm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t = torch.tensor([ms_total], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / K
    value = world * B / (ms_step / 1e3)

    # ---- e2e: the call a user makes (model(c_t, prompt)) with pinned-host input and a device->host read of the result ----
    host_in = synthetic_inputs(B, S, 1024, dt, None, seed_offset=rank)[0].pin_memory()
    host_out = torch.empty(B, 3, S, S, dtype=dt).pin_memory()
    gathered_img = torch.empty(world * B, 3, S, S, device="cuda", dtype=dt) if world > 1 else None
    with torch.no_grad():
        for _ in range(2):
            host_out.copy_(call(host_in.cuda(non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            flush.zero_()
            y = call(host_in.cuda(non_blocking=True))               # H2D + randn(eps) + path (the call a user makes)
            if world > 1:
                dist.all_gather_into_tensor(gathered_img, y)        # sharded users gather the images (dist.sharded_forward)
            host_out.copy_(y, non_blocking=True)                    # D2H of the step's result
            torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / K
    te = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B / te.item()
    io_bytes = B * 3 * S * S * 2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tapgemm = every conv/linear/attention GEMM), per-launch CUDA events ----
    peak_tf, peak_gbs, peak_src = measured_peaks()
    eng.forward(c_t, text_emb, eps, out=out, out_latent=lat)
    torch.cuda.synchronize()
    prof = eng.profile(reps=2)
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump(prof, open(args.profile_out, "w"))
    by_kind = {}
    for p in prof:
        k = by_kind.setdefault(p["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        k["ms"] += p["ms"]; k["flops"] += p["flops"]; k["bytes"] += p["bytes"]; k["n"] += 1
    tg = [p for p in prof if p["kind"].startswith("tapgemm")]
    tg_ms, tg_fl = sum(p["ms"] for p in tg), sum(p["flops"] for p in tg)
    all_ms = sum(p["ms"] for p in prof)
    achieved = tg_fl / (tg_ms * 1e-3) / 1e12 if tg_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "tapgemm_dram_traffic.json")     # from the committed ncu capture of this command
    if os.path.exists(tpath) and args.model == "pix2pix" and B == 8 and S == 512:
        traffic = json.load(open(tpath)).get("dram_bytes_per_step")
    roofline = {"bound": "tensor", "kernel": "tapgemm_kernel + tapgemm2_kernel (tcgen05 implicit GEMM: conv3x3/1x1/linear/attention; "
                                             "all launches of a step, CTA-pair variant for the large layers)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "peak_source": peak_src,
                "launches_per_step": len(tg), "avg_launch_ms": tg_ms / max(1, len(tg)),
                "share_of_step": tg_ms / all_ms if all_ms else None,
                "traffic": traffic, "traffic_note": "sum of dram__bytes_read+write over the step's tapgemm launches (ncu, profiles/)",
                "algorithmic_bytes": sum(p["bytes"] for p in tg),
                "step_tensor_frac": (B * FLOPS_PER_IMAGE / (ms_step * 1e-3)) / 1e12 / peak_tf,
                "by_kind_ms": {k: round(v["ms"], 4) for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1]["ms"])},
                "hbm_kernels_gbs": {k: round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) for k, v in by_kind.items()
                                    if not k.startswith("tapgemm") and v["ms"] > 0 and v["bytes"] > 0}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ips, sec = cpu_oracle_images_per_s(model._sd, S, 1, 0, 1024, W.SD_TURBO) if args.model == "pix2pix" else (None, 0.0)
        cpu = None if ips is None else {"value": ips, "unit": "images/s", "cores": CPU_THREADS, "kind": "port",
               "sample": f"1 x one {S}x{S} image (batch 1), oracle fp32 restatement of the diffusers path, {sec:.1f} s"}

    line = {"metric": "512x512 images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": workload, "per_gpu_batch": B, "global_batch": world * B, "size": S,
                       "parallelism": f"dp{world}", "l2": "256 MiB flush write between timed iterations",
                       "collective": "1 x all_gather_into_tensor(output latents) per step" if world > 1 else "none",
                       "cuda_graph": True},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes},
            "gpu_launches": K * eng.launch_count(B, S, S), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

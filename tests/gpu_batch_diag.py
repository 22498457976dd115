"""Batch-consistency diagnostic: engine forward with batch B (text batch tb) vs the same engine on each image alone,
stage by stage and image by image.  usage: python tests/gpu_batch_diag.py tiny|full B tb [H]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "img2img-turbo_b200"))
import i2it  # noqa: E402
import weights as W  # noqa: E402

STAGES = ["skip0", "skip1", "skip2", "skip3", "enc_mid", "moments", "latent", "unet_mid", "model_pred", "dec_in", "dec_mid",
          "dec_up0", "dec_up1", "dec_up2", "dec_up3"]


def main():
    size, B, tb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    cfg = W.TINY if size == "tiny" else W.SD_TURBO
    H = int(sys.argv[4]) if len(sys.argv) > 4 else (64 if size == "tiny" else 512)
    graph = (len(sys.argv) > 5 and sys.argv[5] == "graph")
    dt = torch.bfloat16
    sd = W.make_state_dict("pix2pix", cfg, seed=0, perturb_norm=True)
    e = i2it.Engine(dt, i2it.PIX2PIX, cfg=cfg, keep_stages=True, use_cuda_graph=graph)
    e.load_state_dict(sd)
    e.set_adapter_scale("default", 1.0)
    e.set_adapter_scale("vae_skip", 2.0)
    e.finalize(1.0, 1.0, 1.0, -1.0)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, 1, H, H, generator=g) < 0.08).float().expand(-1, 3, -1, -1).contiguous().to(dt).cuda()
    text = torch.randn(1, 77, cfg["cross_dim"], generator=g).expand(tb, -1, -1).contiguous().to(dt).cuda()   # same prompt everywhere
    eps = torch.randn(B, 4, H // 8, H // 8, generator=g).to(dt).cuda()
    out = e.forward(x, text, eps).clone()
    torch.cuda.synchronize()
    st = {n: e.read_stage(n) for n in STAGES}
    print(f"== {size} B={B} tb={tb} H={H} graph={graph}: image nan={torch.isnan(out.float()).sum().item()} "
          f"sat={(out.float().abs() >= 1).float().mean().item():.3f}", flush=True)
    for i in range(B):
        o1 = e.forward(x[i:i + 1].contiguous(), text[:1].contiguous(), eps[i:i + 1].contiguous()).clone()
        torch.cuda.synchronize()
        row = []
        for n in STAGES:
            a, b = st[n][i], e.read_stage(n)[0]
            d = (a - b).abs()
            row.append(f"{n}:{d.max().item():.2e}" + ("(nan)" if torch.isnan(a).any() else ""))
        d = (out[i].float() - o1[0].float()).abs()
        print(f" img{i}: out_max_diff={d.max().item():.3e} | " + " ".join(row), flush=True)


if __name__ == "__main__":
    main()

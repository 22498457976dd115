"""Generates tests/golden/*.pt: small seeded input/output vectors of the ORACLE (this repo's restatement).

NOTE (parity unpinned): the reference ships no golden vectors and its arithmetic (diffusers 0.25.1 / peft) cannot be
imported in this image, so these fixtures pin the oracle against ITSELF across refactors and machines; they are
regression vectors, not reference outputs.   Run:  python oracle/gen_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "img2img-turbo_b200"))
import oracle as O  # noqa: E402
import weights as W  # noqa: E402


def inputs(kind, B, H, cfg, seed=1):
    g = torch.Generator().manual_seed(seed)
    if kind == "pix2pix":
        x = (torch.rand(B, 1, H, H, generator=g) < 0.08).float().expand(-1, 3, -1, -1).contiguous()
    else:
        x = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    text = torch.randn(B, 77, cfg["cross_dim"], generator=g)
    eps = torch.randn(B, 4, H // 8, H // 8, generator=g)
    noise = torch.randn(B, 4, H // 8, H // 8, generator=g)
    return x, text, eps, noise


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    cfg = W.TINY
    with torch.no_grad():
        sd = W.make_state_dict("pix2pix", cfg, seed=0, perturb_norm=True)
        x, text, eps, noise = inputs("pix2pix", 2, 64, cfg)
        st = {}
        y = O.pix2pix_forward(sd, x, text, eps, cfg, stages=st)
        torch.save({"image": y.half(), "latent": st["latent"], "model_pred": st["model_pred"], "x_denoised": st["x_denoised"]},
                   os.path.join(out, "pix2pix_tiny_det.pt"))
        sdt = W.make_state_dict("pix2pix", cfg, seed=0, twin=True, perturb_norm=True)
        st = {}
        y = O.pix2pix_forward(sdt, x, text, eps, cfg, deterministic=False, r=0.4, noise_map=noise, stages=st)
        torch.save({"image": y.half(), "model_pred": st["model_pred"], "x_denoised": st["x_denoised"]},
                   os.path.join(out, "pix2pix_tiny_stochastic.pt"))
        sdc = W.make_state_dict("cyclegan", cfg, seed=0, perturb_norm=True)
        x, text, eps, _ = inputs("cyclegan", 2, 64, cfg)
        for d in ("a2b", "b2a"):
            st = {}
            y = O.cyclegan_forward(sdc, x, text, eps, d, cfg, stages=st)
            torch.save({"image": y.half(), "model_pred": st["model_pred"], "x_denoised": st["x_denoised"]},
                       os.path.join(out, f"cyclegan_tiny_{d}.pt"))
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main()

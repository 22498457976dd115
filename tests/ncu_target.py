"""Profiling target: one warm forward, then exactly one forward of BASELINE config #2 (pix2pix, bf16, batch 8, 512x512)
between cudaProfilerStart/Stop.  Run under `ncu --profile-from-start off ...` (never report timings taken this way)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "img2img-turbo_b200"))
import i2it  # noqa: E402
import weights as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dt = torch.bfloat16
sd = W.make_state_dict("pix2pix", W.SD_TURBO, seed=0)
e = i2it.Engine(dt, i2it.PIX2PIX, use_cuda_graph=False)
e.load_state_dict(sd)
e.set_adapter_scale("default", 1.0)
e.set_adapter_scale("vae_skip", 2.0)
e.finalize(1.0, 1.0, 1.0, -1.0)
g = torch.Generator().manual_seed(1)
x = (torch.rand(B, 1, H, H, generator=g) < 0.08).float().expand(-1, 3, -1, -1).contiguous().to(dt).cuda()
text = torch.randn(1, 77, 1024, generator=g).to(dt).cuda()
eps = torch.randn(B, 4, H // 8, H // 8, generator=g).to(dt).cuda()
e.set_text(text)                       # the bench path: prompt K/V projected once, not per forward
out = e.forward(x, None, eps)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = e.forward(x, None, eps)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", torch.isnan(out.float()).sum().item())

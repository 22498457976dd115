"""Independent pins of the oracle's building blocks.

diffusers / peft are not importable here, so the oracle cannot be pinned against the reference stack itself.  What CAN be pinned:
every functional block of oracle/oracle.py against the corresponding torch.nn MODULE (the classes diffusers composes:
nn.GroupNorm, nn.LayerNorm, nn.Conv2d, nn.Linear, nn.MultiheadAttention, F.scaled_dot_product_attention, nn.GELU, nn.Upsample)
loaded with the same tensors, plus the peft LoRA forward written as modules.  Each test names the diffusers / peft file the
oracle function restates."""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import oracle as O
import weights as W


def _close(a, b, tol=2e-5):
    assert a.shape == b.shape
    assert (a - b).abs().max().item() <= tol * (1 + b.abs().max().item()), (a - b).abs().max().item()


def test_group_norm_and_layer_norm_vs_nn_modules():
    # diffusers models/resnet.py ResnetBlock2D.norm1/norm2 = nn.GroupNorm(32, C, eps); attention.py BasicTransformerBlock.norm* = nn.LayerNorm
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 64, 5, 7, generator=g)
    sd = {"n.weight": torch.randn(64, generator=g), "n.bias": torch.randn(64, generator=g)}
    gn = nn.GroupNorm(32, 64, eps=1e-6)
    gn.load_state_dict({"weight": sd["n.weight"], "bias": sd["n.bias"]})
    _close(O.group_norm(sd, "n", x, 32, 1e-6), gn(x))
    t = torch.randn(3, 11, 64, generator=g)
    ln = nn.LayerNorm(64, eps=1e-5)
    ln.load_state_dict({"weight": sd["n.weight"], "bias": sd["n.bias"]})
    _close(O.layer_norm(sd, "n", t), ln(t))


def test_lora_linear_and_conv_vs_peft_style_modules():
    # peft tuners/lora/layer.py: Linear.forward = base(x) + lora_B(lora_A(dropout(x))) * scaling;  Conv2d: lora_A = Conv2d(cin, r, k,
    # stride, padding, bias=False), lora_B = Conv2d(r, cout, 1, bias=False)
    g = torch.Generator().manual_seed(1)
    base, A, B = nn.Linear(24, 40), nn.Linear(24, 8, bias=False), nn.Linear(8, 40, bias=False)
    sd = {"l.weight": base.weight.data, "l.bias": base.bias.data, "l.lora_A.ad.weight": A.weight.data, "l.lora_B.ad.weight": B.weight.data}
    x = torch.randn(5, 24, generator=g)
    _close(O.linear(sd, "l", x, {"ad": 0.7}), base(x) + B(A(x)) * 0.7)
    _close(O.linear(sd, "l", x, {}), base(x))                                   # inactive adapter -> base layer
    for stride, pad in ((1, 1), (2, 1), (2, 0)):
        cb = nn.Conv2d(6, 10, 3, stride=stride, padding=pad)
        ca, cB = nn.Conv2d(6, 4, 3, stride=stride, padding=pad, bias=False), nn.Conv2d(4, 10, 1, bias=False)
        sd = {"c.base_layer.weight": cb.weight.data, "c.base_layer.bias": cb.bias.data, "c.lora_A.v.weight": ca.weight.data,
              "c.lora_B.v.weight": cB.weight.data}
        x = torch.randn(2, 6, 9, 9, generator=g)
        _close(O.conv2d(sd, "c", x, {"v": 2.0}, stride=stride, padding=pad), cb(x) + cB(ca(x)) * 2.0)
        # the fold the product uses at load time is the same function
        wf = O.fold_lora(sd, "c", {"v": 2.0})
        _close(F.conv2d(x, wf, cb.bias.data, stride=stride, padding=pad), cb(x) + cB(ca(x)) * 2.0, 1e-4)


def test_sdpa_vs_torch_sdpa_and_multihead_attention():
    # diffusers models/attention_processor.py AttnProcessor2_0: F.scaled_dot_product_attention(q, k, v) on [B, heads, N, d]
    g = torch.Generator().manual_seed(2)
    B, Nq, Nk, heads, d = 2, 10, 7, 4, 16
    C = heads * d
    q, k, v = torch.randn(B, Nq, C, generator=g), torch.randn(B, Nk, C, generator=g), torch.randn(B, Nk, C, generator=g)
    sp = lambda t, n: t.view(B, n, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q, Nq), sp(k, Nk), sp(v, Nk)).transpose(1, 2).reshape(B, Nq, C)
    _close(O.sdpa(q, k, v, heads), ref)
    # a whole attention layer (to_q/to_k/to_v/to_out) against nn.MultiheadAttention with the same matrices
    x, ctx = torch.randn(B, Nq, C, generator=g), torch.randn(B, Nk, 24, generator=g)
    mha = nn.MultiheadAttention(C, heads, bias=False, kdim=24, vdim=24, batch_first=True)
    mha.out_proj.bias = nn.Parameter(torch.randn(C, generator=g))
    sd = {"a.to_q.weight": mha.q_proj_weight.data, "a.to_k.weight": mha.k_proj_weight.data, "a.to_v.weight": mha.v_proj_weight.data,
          "a.to_out.0.weight": mha.out_proj.weight.data, "a.to_out.0.bias": mha.out_proj.bias.data}
    mine = O.linear(sd, "a.to_out.0", O.sdpa(O.linear(sd, "a.to_q", x, {}), O.linear(sd, "a.to_k", ctx, {}),
                                             O.linear(sd, "a.to_v", ctx, {}), heads), {})
    _close(mine, mha(x, ctx, ctx, need_weights=False)[0], 1e-4)


def test_vae_resnet_and_attention_vs_module_composition():
    """diffusers models/resnet.py ResnetBlock2D (temb=None, output_scale_factor=1) and models/attention_processor.py Attention
    (1 head, residual_connection, group_norm) rebuilt from torch.nn modules."""
    cfg = dict(W.TINY)
    g = torch.Generator().manual_seed(3)
    cin, cout = 64, 128

    class Res(nn.Module):
        def __init__(s):
            super().__init__()
            s.norm1, s.conv1 = nn.GroupNorm(32, cin, eps=1e-6), nn.Conv2d(cin, cout, 3, padding=1)
            s.norm2, s.conv2 = nn.GroupNorm(32, cout, eps=1e-6), nn.Conv2d(cout, cout, 3, padding=1)
            s.conv_shortcut = nn.Conv2d(cin, cout, 1)

        def forward(s, x):
            h = s.conv2(F.silu(s.norm2(s.conv1(F.silu(s.norm1(x))))))
            return s.conv_shortcut(x) + h
    m = Res()
    for p in m.parameters():
        p.data = torch.randn(p.shape, generator=g) * 0.1
    sd = {"r." + k: v for k, v in m.state_dict().items()}
    x = torch.randn(2, cin, 8, 8, generator=g)
    _close(O._vae_resnet(sd, "r", x, cfg, {}), m(x), 1e-4)

    C = 64
    gn, mha = nn.GroupNorm(32, C, eps=1e-6), nn.MultiheadAttention(C, 1, batch_first=True)
    for p in list(gn.parameters()) + list(mha.parameters()):
        p.data = torch.randn(p.shape, generator=g) * 0.1
    wq, wk, wv = mha.in_proj_weight.data.chunk(3)
    bq, bk, bv = mha.in_proj_bias.data.chunk(3)
    sd = {"a.group_norm.weight": gn.weight.data, "a.group_norm.bias": gn.bias.data, "a.to_q.weight": wq, "a.to_q.bias": bq,
          "a.to_k.weight": wk, "a.to_k.bias": bk, "a.to_v.weight": wv, "a.to_v.bias": bv,
          "a.to_out.0.weight": mha.out_proj.weight.data, "a.to_out.0.bias": mha.out_proj.bias.data}
    x = torch.randn(2, C, 4, 4, generator=g)
    t = gn(x).flatten(2).transpose(1, 2)
    ref = mha(t, t, t, need_weights=False)[0].transpose(1, 2).reshape(2, C, 4, 4) + x
    _close(O._vae_attn(sd, "a", x, cfg, {}), ref, 1e-4)


def test_geglu_feed_forward_and_timestep_embedding():
    # diffusers models/activations.py GEGLU: hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate)  (erf GELU = nn.GELU())
    g = torch.Generator().manual_seed(4)
    proj, out, act = nn.Linear(16, 128), nn.Linear(64, 16), nn.GELU()
    x = torch.randn(3, 5, 16, generator=g)
    h, gate = proj(x).chunk(2, dim=-1)
    ref = out(h * act(gate))
    sd = {"f.net.0.proj.weight": proj.weight.data, "f.net.0.proj.bias": proj.bias.data, "f.net.2.weight": out.weight.data,
          "f.net.2.bias": out.bias.data}
    hg = O.linear(sd, "f.net.0.proj", x, {})
    hh, gg = hg.chunk(2, dim=-1)
    _close(O.linear(sd, "f.net.2", hh * F.gelu(gg), {}), ref)
    # diffusers models/embeddings.py get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0), in float64 numpy-style
    dim, t = 320, 999
    half = dim // 2
    expo = -math.log(10000) * torch.arange(half, dtype=torch.float64) / half
    emb = t * torch.exp(expo)
    ref = torch.cat([torch.cos(emb), torch.sin(emb)]).float()[None]
    _close(O.timestep_embedding(t, dim), ref, 1e-4)


def test_sampling_blocks_vs_nn_modules():
    # diffusers models/resnet.py Downsample2D(padding=0): F.pad(x, (0,1,0,1)) then Conv2d(stride 2); Upsample2D: nearest 2x then conv
    g = torch.Generator().manual_seed(5)
    conv = nn.Conv2d(8, 8, 3, stride=2, padding=0)
    x = torch.randn(1, 8, 10, 10, generator=g)
    sd = {"d.weight": conv.weight.data, "d.bias": conv.bias.data}
    _close(O.conv2d(sd, "d", F.pad(x, (0, 1, 0, 1)), {}, stride=2), conv(nn.ZeroPad2d((0, 1, 0, 1))(x)))
    up = nn.Upsample(scale_factor=2.0, mode="nearest")
    assert torch.equal(F.interpolate(x, scale_factor=2.0, mode="nearest"), up(x))
    # DiagonalGaussianDistribution.sample (diffusers models/vae.py): mean + exp(0.5*clamp(logvar,-30,20)) * eps
    mom = torch.randn(2, 8, 4, 4, generator=g) * 20
    mean, logvar = mom.chunk(2, dim=1)
    eps = torch.randn(2, 4, 4, 4, generator=g)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    assert torch.isfinite(mean + std * eps).all()


def test_scheduler_constants_vs_closed_form_float64():
    # diffusers schedulers/scheduling_ddpm.py: betas = linspace(sqrt(b0), sqrt(b1), T)**2 (fp32), alphas_cumprod = cumprod(1-betas)
    b = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    ac64 = torch.cumprod(1 - b, 0)[999].item()
    assert abs(float(O.alphas_cumprod()[999]) - ac64) < 2e-7
    # "trailing" spacing with 1 inference step: timesteps = round(arange(1000, 0, -1000)) - 1 = [999]; prev_t = 999 - 1000 < 0 -> abar_prev = 1
    assert int(round(1000 / 1) * 1 - 1) == 999

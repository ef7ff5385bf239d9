"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the vanilla-NeRF renderer of the reference
(models/vanilla_nerf/model.py:44-216, models/vanilla_nerf/helper.py:415-616), SURVEY.md section 8(a) row a17.

Pinned to the unmodified reference by oracle/make_golden.py (tests/golden/vanilla_reference_vectors.npz)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .neo360_oracle import pos_enc, piecewise_constant_pdf, _jitter

Tensor = torch.Tensor


def sample_along_rays(o, d, n, near, far, u_rand: Optional[Tensor] = None):
    """helper.py:415-442 (lindisp=False): t = near(1-u) + far u along `d` (the caller passes viewdirs, quirk Q15)."""
    u = torch.linspace(0.0, 1.0, n + 1, device=o.device)
    t = near * (1.0 - u) + far * u
    t = _jitter(t, u_rand) if u_rand is not None else torch.broadcast_to(t, (o.shape[0], n + 1))
    return t, o[:, None, :] + t[..., None] * d[:, None, :]


def sample_pdf(o, d, t_old, w, m, u_rand=None):
    """helper.py:610-616 with bins = mids(t_old), weights = w[..., 1:-1] as NeRF.forward forms them (model.py:171-181)."""
    mids = 0.5 * (t_old[..., 1:] + t_old[..., :-1])
    t_new = piecewise_constant_pdf(mids, w[..., 1:-1], m, u_rand)
    t = torch.sort(torch.cat([t_old, t_new], -1), -1).values
    return t, o[:, None, :] + t[..., None] * d[:, None, :]


def mlp_forward(P: Dict[str, Tensor], pre: str, enc: Tensor, dir_enc: Tensor):
    """NeRFMLP.forward (model.py:100-125).  enc (B,N,63), dir_enc (B,27) -> raw rgb (B,N,3), raw sigma (B,N,1)."""
    B, N, _ = enc.shape
    lin = lambda name, x: F.linear(x, P[pre + name + ".weight"], P[pre + name + ".bias"])
    inp = enc.reshape(-1, enc.shape[-1])
    x = inp
    for i in range(8):
        x = torch.relu(lin(f"pts_linears.{i}", x))
        if i == 4:
            x = torch.cat([x, inp], -1)
    raw_sigma = lin("density_layer", x).reshape(B, N, 1)
    beta = lin("bottleneck_layer", x)
    cond = dir_enc[:, None, :].expand(B, N, dir_enc.shape[-1]).reshape(-1, dir_enc.shape[-1])
    y = torch.relu(lin("views_linear.0", torch.cat([beta, cond], -1)))
    return lin("rgb_layer", y).reshape(B, N, 3), raw_sigma


def composite(rgb, sigma, t, dirs, white_bkgd: bool):
    """helper.py:521-559 (quirks Q9, Q10)."""
    dist = torch.cat([t[..., 1:] - t[..., :-1], torch.ones_like(t[..., :1]) * 1e10], -1)
    dist = dist * torch.norm(dirs[..., None, :], dim=-1)
    alpha = 1.0 - torch.exp(-sigma[..., 0] * dist)
    T = torch.cat([torch.ones_like(alpha[..., :1]), torch.cumprod(1.0 - alpha[..., :-1] + 1e-10, -1)], -1)
    w = alpha * T
    out = (w[..., None] * rgb).sum(-2)
    depth = torch.nan_to_num((w * t).sum(-1), float("inf"))
    depth = torch.clamp(depth, torch.min(depth), torch.max(depth))
    acc = w.sum(-1)
    if white_bkgd:
        out = out + (1.0 - acc[..., None])
    return out, acc, w, depth


def render(rays: Dict[str, Tensor], P: Dict[str, Tensor], n_coarse: int, n_fine: int, near: float, far: float,
           white_bkgd: bool = False, rand: Optional[Dict[str, Tensor]] = None, return_aux: bool = False):
    """NeRF.forward (model.py:154-216): marches along viewdirs, composites with |rays_d|."""
    o, d, vd = rays["rays_o"], rays["rays_d"], rays["viewdirs"]
    denc = pos_enc(vd, 0, 4)
    ret, aux = [], []
    t = w = None
    for lvl in range(2):
        if lvl == 0:
            t, pts = sample_along_rays(o, vd, n_coarse, near, far, None if rand is None else rand.get("u0"))
        else:
            t, pts = sample_pdf(o, vd, t, w, n_fine, None if rand is None else rand.get("u1"))
        pre = "coarse_mlp." if lvl == 0 else "fine_mlp."
        raw_rgb, raw_sigma = mlp_forward(P, pre, pos_enc(pts, 0, 10), denc)
        rgb = torch.sigmoid(raw_rgb) * (1 + 2 * 0.001) - 0.001
        sigma = F.softplus(raw_sigma - 1.0)
        comp, acc, w, depth = composite(rgb, sigma, t, d, white_bkgd)
        ret.append((comp, acc, depth))
        aux.append(dict(t=t, rgb=rgb, sigma=sigma, w=w))
    return (ret, aux) if return_aux else ret

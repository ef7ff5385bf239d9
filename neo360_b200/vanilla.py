"""Drop-in for the reference's vanilla NeRF renderer (models/vanilla_nerf/model.py:44-216), SURVEY.md section 8(a) row a17.

`NeRF.forward(rays, randomized, white_bkgd, near, far)` returns the reference's `list[2]` of `(comp_rgb, acc, depth)`
(model.py:214).  Parameter names and shapes equal the reference's (`coarse_mlp.pts_linears.0.weight`, ...), so its checkpoints
load.  Arithmetic (csrc/vanilla.cu): `self.precision = "fp32"` (default) runs the layers on fp32 CUDA cores in the reference
formulation; `"tc"` runs them as fp16 tcgen05 GEMMs with fp32 accumulation (csrc/gemm_tc.cu).  CUDA only, no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
import torch.nn as nn

from . import _lib as L


class NeRFMLP(nn.Module):
    def __init__(self, min_deg_point=0, max_deg_point=10, deg_view=4, netdepth: int = 8, netwidth: int = 256, netdepth_condition: int = 1,
                 netwidth_condition: int = 128, skip_layer: int = 4, input_ch: int = 3, input_ch_view: int = 3, num_rgb_channels: int = 3,
                 num_density_channels: int = 1):
        super().__init__()
        if (min_deg_point, max_deg_point, deg_view, netdepth, netwidth, netdepth_condition, netwidth_condition, skip_layer, input_ch,
                input_ch_view) != (0, 10, 4, 8, 256, 1, 128, 4, 3, 3):
            raise NotImplementedError("the CUDA path implements the reference's default NeRFMLP architecture")
        pos = ((max_deg_point - min_deg_point) * 2 + 1) * input_ch
        view = (deg_view * 2 + 1) * input_ch_view
        layers = [nn.Linear(pos, netwidth)]
        for idx in range(netdepth - 1):
            layers.append(nn.Linear(netwidth + pos if (idx % skip_layer == 0 and idx > 0) else netwidth, netwidth))
        self.pts_linears = nn.ModuleList(layers)
        self.views_linear = nn.ModuleList([nn.Linear(netwidth + view, netwidth_condition)])
        self.bottleneck_layer = nn.Linear(netwidth, netwidth)
        self.density_layer = nn.Linear(netwidth, num_density_channels)
        self.rgb_layer = nn.Linear(netwidth_condition, num_rgb_channels)
        for m in list(self.pts_linears) + [self.bottleneck_layer, self.density_layer, self.rgb_layer]:
            nn.init.xavier_uniform_(m.weight)

    def c_params(self, keep: list) -> L.NeoVanillaMLPParams:
        p = L.NeoVanillaMLPParams()
        f = lambda t: (keep.append(t.detach().contiguous().float()) or keep[-1])
        for i in range(8):
            p.w[i] = L.ptr(f(self.pts_linears[i].weight))
            p.b[i] = L.ptr(f(self.pts_linears[i].bias))
        for wn, bn, lin in (("wb", "bb", self.bottleneck_layer), ("wsig", "bsig", self.density_layer),
                            ("wv0", "bv0", self.views_linear[0]), ("wrgb", "brgb", self.rgb_layer)):
            setattr(p, wn, L.ptr(f(lin.weight)))
            setattr(p, bn, L.ptr(f(lin.bias)))
        return p

    def forward(self, *a, **k):
        raise RuntimeError("NeRFMLP is evaluated inside the CUDA path; call NeRF.forward")


class NeRF(nn.Module):
    def __init__(self, num_levels: int = 2, min_deg_point: int = 0, max_deg_point: int = 10, deg_view: int = 4, num_coarse_samples: int = 64,
                 num_fine_samples: int = 128, use_viewdirs: bool = True, noise_std: float = 0.0, lindisp: bool = False):
        super().__init__()
        if num_levels != 2 or lindisp or noise_std != 0.0 or not use_viewdirs:
            raise NotImplementedError("reference defaults only (models/vanilla_nerf/model.py:129-139)")
        self.num_coarse_samples, self.num_fine_samples = num_coarse_samples, num_fine_samples
        self.coarse_mlp = NeRFMLP(min_deg_point, max_deg_point, deg_view)
        self.fine_mlp = NeRFMLP(min_deg_point, max_deg_point, deg_view)
        self._handle = None
        self._key = None
        self._ws = None

    def _ensure(self, dev):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._handle is None or key != self._key:
            self.release()
            lib = L.load()
            keep = []
            arr = (L.NeoVanillaMLPParams * 2)(self.coarse_mlp.to(dev).c_params(keep), self.fine_mlp.to(dev).c_params(keep))
            h = C.c_void_p()
            with torch.cuda.device(dev):
                L.check(lib.neo_vanilla_create(arr, C.byref(h), torch.cuda.current_stream().cuda_stream))
            self._handle, self._key = h, tuple((p.data_ptr(), p._version) for p in self.parameters())
        return self._handle

    def release(self):
        if self._handle is not None:
            L.load().neo_vanilla_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def forward(self, rays: Dict[str, torch.Tensor], randomized: bool, white_bkgd: bool, near, far, debug: bool = False) -> List[tuple]:
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("backward through the CUDA path is not built yet; call under torch.no_grad() / .eval()")
        o = rays["rays_o"].contiguous().float()
        d = rays["rays_d"].contiguous().float()
        vd = rays["viewdirs"].contiguous().float()
        if not o.is_cuda:
            raise RuntimeError("neo360_b200 needs CUDA tensors (no CPU fallback)")
        lib = L.load()
        h = self._ensure(o.device)
        n, dev = o.shape[0], o.device
        nc, nf = self.num_coarse_samples, self.num_fine_samples
        cfg = L.NeoVanillaCfg()
        cfg.n_coarse, cfg.n_fine, cfg.white_bkgd = nc, nf, int(bool(white_bkgd))
        cfg.near_plane, cfg.far_plane = float(near), float(far)
        cfg.precision = {"fp32": L.NEO_PREC_FP32, "tc": L.NEO_PREC_TC}[getattr(self, "precision", "fp32")]
        keep = []
        if randomized:
            u = rays.get("_uniforms") or [torch.rand((n, nc + 1), device=dev), torch.rand((n, nf), device=dev)]   # helper.py:438, 587
            keep.extend(u)
            cfg.u0, cfg.u1 = L.ptr(u[0].contiguous()), L.ptr(u[1].contiguous())
        r = L.NeoRays()
        r.n_rays, r.chunk = n, 0
        r.rays_o, r.rays_d, r.viewdirs = L.ptr(o), L.ptr(d), L.ptr(vd)
        need = lib.neo_vanilla_workspace_bytes(n, C.byref(cfg))
        if need == 0:
            raise RuntimeError("neo360_b200: " + lib.neo_last_error().decode())
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = L.NeoVanillaOut()
        T = {k: [] for k in L.VANILLA_OUT_FIELDS}
        N = (nc + 1, nc + 1 + nf)
        for lvl in range(2):
            shapes = {"comp_rgb": (n, 3), "acc": (n,), "depth": (n,)}
            if debug:
                shapes.update({"t": (n, N[lvl]), "sigma": (n, N[lvl], 1), "rgb_s": (n, N[lvl], 3), "weights": (n, N[lvl])})
            for k, shp in shapes.items():
                t = torch.empty(*shp, device=dev)
                T[k].append(t)
                getattr(out, k)[lvl] = t.data_ptr()
        with torch.cuda.device(dev):
            L.check(lib.neo_vanilla_render_fwd(h, C.byref(r), C.byref(cfg), C.byref(out), self._ws.data_ptr(), self._ws.numel(),
                                               torch.cuda.current_stream().cuda_stream))
        if debug:
            self.last_debug = T
        return [(T["comp_rgb"][lvl], T["acc"][lvl], T["depth"][lvl]) for lvl in range(2)]

"""TEST INFRASTRUCTURE ONLY -- import shim for the UNMODIFIED reference (zubair-irshad/NeO-360).

Lets the reference's own modules under /root/reference import in this container (CPU torch, no
pytorch_lightning / kornia / lpips / ... installed) so that `oracle/make_golden.py` can mint golden
vectors from the reference itself and pin `oracle/neo360_oracle.py` against it.

/root/reference does not exist on the GPU box: nothing in tests marked `gpu`, `smoke()` or
`bench.py` imports this file.  Nothing under `neo360_b200/` imports anything under `oracle/`.

What is stubbed: pytorch_lightning, kornia.create_meshgrid, piqa, lpips, dotmap,
torch_efficient_distloss, imageio, matplotlib, wandb.  None of the stubs sits on the hot path except
`create_meshgrid` (15-line equivalent of kornia's, used by datasets/ray_utils.py:97).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("NEO360_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "neo360"))


def _create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    # kornia.utils.create_meshgrid semantics: (1, H, W, 2) with [...,0]=x in [0,W-1], [...,1]=y in [0,H-1]
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).unsqueeze(0)


class _Anything(types.ModuleType):
    """Module whose every attribute is a harmless callable/class stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        class _Stub:  # usable as base class, decorator, callable
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                return None

        _Stub.__name__ = name
        setattr(self, name, _Stub)
        return _Stub


def install():
    """Install the stubs and put the reference root on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        ROOTS = ("pytorch_lightning", "kornia", "piqa", "lpips", "dotmap", "torch_efficient_distloss",
                 "imageio", "wandb", "open3d", "matplotlib", "plotly")

        def find_spec(self, fullname, path=None, target=None):
            if fullname.split(".")[0] in self.ROOTS:
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            m = _Anything(spec.name)
            m.__path__ = []
            return m

        def exec_module(self, module):
            name = module.__name__
            if name == "pytorch_lightning":
                class LightningModule(nn.Module):
                    def save_hyperparameters(self, *a, **k):
                        pass

                    def log(self, *a, **k):
                        pass

                module.LightningModule = LightningModule
            if name in ("kornia", "kornia.utils"):
                module.create_meshgrid = _create_meshgrid

    if not any(type(f).__name__ == "_StubFinder" for f in sys.meta_path):
        # only stub what is genuinely missing
        finder = _StubFinder()
        missing = []
        for r in finder.ROOTS:
            try:
                if importlib.util.find_spec(r) is None:
                    missing.append(r)
            except (ImportError, ValueError):
                missing.append(r)
        finder.ROOTS = tuple(missing)
        sys.meta_path.append(finder)
    return REF_ROOT


def load():
    """Return a namespace with the reference symbols that sit on the hot path."""
    install()
    import importlib

    ns = types.SimpleNamespace()
    ns.neo_helper = importlib.import_module("models.neo360.helper")
    ns.neo_util = importlib.import_module("models.neo360.util")
    ns.neo_tp = importlib.import_module("models.neo360.encoder_tp_fusion_conv")
    ns.neo_pn = importlib.import_module("models.neo360.encoder_pn")
    ns.neo_model = importlib.import_module("models.neo360.model")
    ns.ray_utils = importlib.import_module("datasets.ray_utils")
    ns.van_helper = importlib.import_module("models.vanilla_nerf.helper")
    ns.van_model = importlib.import_module("models.vanilla_nerf.model")
    ns.mip_helper = importlib.import_module("models.mipnerf360.helper")
    ns.mip_model = importlib.import_module("models.mipnerf360.model")
    return ns


def make_reference_nerf_tp(ns, num_coarse, num_fine, num_src_views=3, seed=0):
    """Construct the reference NeRF_TP with random (xavier) MLP weights and a no-download ResNet."""
    import torchvision

    orig = torchvision.models.resnet34

    def resnet34_nodl(*a, **k):
        k.pop("pretrained", None)
        return orig(weights=None)

    torchvision.models.resnet34 = resnet34_nodl
    try:
        torch.manual_seed(seed)
        net = ns.neo_model.NeRF_TP(num_coarse_samples=num_coarse, num_fine_samples=num_fine,
                                   num_src_views=num_src_views)
    finally:
        torchvision.models.resnet34 = orig
    return net.eval()


def bypass_encoder(net, planes_xz, planes_xy, planes_yz, latent):
    """Make net.encoder(...) return the given planes and install `latent` as the pixel-aligned feature map
    (the encoder is a producer of 9 planes + 1 latent, outside the hot path)."""
    se = net.encoder.spatial_encoder
    se.latent = latent
    ls = torch.empty(2, dtype=torch.float32)
    ls[0] = latent.shape[-1]
    ls[1] = latent.shape[-2]
    se.latent_scaling = ls / (ls - 1) * 2.0  # encoder_pn.py:204-206

    def fwd(src_imgs, src_poses, src_focal, src_c):
        return planes_xz, planes_xy, planes_yz

    net.encoder.forward = fwd
    return net

"""Tri-plane builder of NeO-360 (`GridEncoder`) behind the reference's own module surface.

    reference                                                     here
    ---------------------------------------------------------     ----------------------------------------------------------------
    models/neo360/encoder_pn.py:13-210   SpatialEncoder            SpatialEncoder: same sub-module names (torchvision ResNet-34 trunk, host
                                                                   framework convolutions), `.latent` / `.latent_scaling` as the renderer reads them
    encoder_tp_fusion_conv.py:262-470    GridEncoder.__init__      GridEncoder.__init__: same sub-modules in the same construction order
                                                                   (state dicts and seeded initialisations are interchangeable)
    encoder_tp_fusion_conv.py:472-597    GridEncoder.forward       forward(): ResNet features and the three floor-plan conv stacks stay in the
                                                                   host framework; everything between them -- 64^3 x NV grid lookup,
                                                                   DepthPillarEncoder, three pillar aggregators, softmax-weighted pillar sums
                                                                   (2.7 TFLOP per scene) -- runs in hand-written CUDA on tcgen05
                                                                   (`neo_grid_encoder_dense`, csrc/encoder.cu + csrc/gemm_tc.cu) when no
                                                                   gradient is required; under autograd the same algebra runs as framework ops.
"""
from __future__ import annotations

import ctypes as C
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L


def _init_linear_kaiming(m):
    """encoder_tp_fusion_conv.py:258-263"""
    if type(m) == nn.Linear:
        nn.init.kaiming_normal_(m.weight)
        nn.init.uniform_(m.bias, -1e-3, 1e-3)


class _ResNet34Trunk(nn.Module):
    """encoder_pn.py:13-30: the first three stages of torchvision's ResNet-34 (the full net is built first so that a seeded
    construction consumes the generator exactly like the reference)."""

    def __init__(self):
        super().__init__()
        import torchvision
        norm = functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
        net = torchvision.models.resnet34(weights=None, norm_layer=norm)
        self.conv1, self.bn1, self.relu, self.maxpool = net.conv1, net.bn1, net.relu, net.maxpool
        self.layer1, self.layer2, self.layer3 = net.layer1, net.layer2, net.layer3


class SpatialEncoder(nn.Module):
    """encoder_pn.py:32-210 with the reference's defaults as GridEncoder passes them (resnet34, 4 layers, bilinear, zeros padding)."""

    def __init__(self):
        super().__init__()
        self.model = _ResNet34Trunk()
        self.latent_size = 512
        self.register_buffer("latent", torch.empty(1, 1, 1, 1), persistent=False)
        self.register_buffer("latent_scaling", torch.empty(2, dtype=torch.float32), persistent=False)

    def forward(self, x):
        x = self.model.relu(self.model.bn1(self.model.conv1(x)))
        feats = [x]
        x = self.model.layer1(self.model.maxpool(x))
        feats.append(x)
        x = self.model.layer2(x)
        feats.append(x)
        x = self.model.layer3(x)
        feats.append(x)
        size = feats[0].shape[-2:]
        self.latent = torch.cat([F.interpolate(f, size, mode="bilinear", align_corners=True) for f in feats], 1)
        ls = torch.tensor([self.latent.shape[-1], self.latent.shape[-2]], dtype=torch.float32, device=self.latent.device)
        self.latent_scaling = ls / (ls - 1) * 2.0
        return self.latent


class DepthPillarEncoder(nn.Module):
    """encoder_tp_fusion_conv.py:234-250"""

    def __init__(self, inp_ch, LS):
        super().__init__()
        self.common_branch = nn.Sequential(nn.Linear(inp_ch, LS), nn.ReLU(inplace=True), nn.Linear(LS, LS), nn.ReLU(inplace=True))
        self.depth_encoder = nn.Linear(LS, LS)
        self.common_branch.apply(_init_linear_kaiming)
        self.depth_encoder.apply(_init_linear_kaiming)

    def forward(self, x):
        return self.depth_encoder(self.common_branch(x))


def _floorplan_convnet():
    """encoder_tp_fusion_conv.py:372-398 (identical for xy / yz / xz): 512 -> 256 (s2) -> 128 (s2) -> 128 -> up x2 -> 128 -> up (120,160) -> 128."""
    return nn.Sequential(
        nn.Conv2d(512, 256, 3, stride=2, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
        nn.Conv2d(256, 128, 3, stride=2, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
        nn.Conv2d(128, 128, 3, stride=1, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
        nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
        nn.Conv2d(128, 128, 3, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
        nn.Upsample(size=(120, 160), mode="bilinear", align_corners=True),
        nn.Conv2d(128, 128, 3, padding=1))


class GridEncoder(nn.Module):
    GRID = 64

    def __init__(self, encoder_type="resnet", **unused):
        super().__init__()
        if encoder_type != "resnet":
            raise NotImplementedError("reference default only (encoder_type='resnet')")
        self.grid_size = [self.GRID] * 3
        self.spatial_encoder = SpatialEncoder()
        LS = self.latent_size = self.spatial_encoder.latent_size
        self.depth_fc = DepthPillarEncoder(inp_ch=LS + 3 + 3, LS=LS)
        mk = lambda: nn.Sequential(nn.Linear(LS + 1, LS), nn.ReLU(inplace=True), nn.Linear(LS, 1))
        self.pillar_aggregator_xz, self.pillar_aggregator_yz, self.pillar_aggregator_xy = mk(), mk(), mk()
        self.floorplan_convnet_xy, self.floorplan_convnet_yz, self.floorplan_convnet_xz = _floorplan_convnet(), _floorplan_convnet(), _floorplan_convnet()
        for m in (self.floorplan_convnet_xy, self.floorplan_convnet_yz, self.floorplan_convnet_xz,
                  self.pillar_aggregator_xz, self.pillar_aggregator_yz, self.pillar_aggregator_xy):
            m.apply(_init_linear_kaiming)
        self._ws = None

    # ---- the dense part, framework ops (autograd; also the fp32 reference of the CUDA path in the tests) ----
    def dense_torch(self, latent, poses, focal, c, W, H):
        """encoder_tp_fusion_conv.py:483-570"""
        G, NV, dev = self.GRID, latent.shape[0], latent.device
        ax = [torch.linspace(-1, 1, G, device=dev), torch.linspace(-1, 1, G, device=dev), torch.linspace(0, 1, G, device=dev)]
        world = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(1, -1, 3).expand(NV, -1, -1)     # (NV, G^3, 3)
        rot = poses[:, :3, :3].transpose(1, 2)
        trans = -torch.bmm(rot, poses[:, :3, 3:])
        cam = torch.matmul(rot[:, None], world.unsqueeze(-1))[..., 0] + trans[:, None, :, 0]
        mask = cam[:, :, 2] < 1e-3
        d = world - poses[:, None, :3, -1]
        d = d / torch.norm(d + 1e-9, dim=-1)[:, :, None] * mask[:, :, None]
        f2 = torch.stack([focal[0], -focal[0]]).reshape(1, 1, 2)
        uv = -cam[..., :2] / (cam[..., 2:] + 1e-9) * f2 + c[0].reshape(1, 1, 2)
        ls = torch.tensor([latent.shape[-1], latent.shape[-2]], dtype=torch.float32, device=dev)
        uv = uv * ((ls / (ls - 1) * 2.0) / torch.tensor([W, H], dtype=torch.float32, device=dev)) - 1.0
        feat = F.grid_sample(latent, uv.unsqueeze(2), align_corners=True, mode="bilinear", padding_mode="zeros")[..., 0]     # (NV, 512, G^3)
        x = torch.cat([feat, cam.permute(0, 2, 1), d.permute(0, 2, 1)], 1).permute(0, 2, 1)
        lat = self.depth_fc(x).reshape(NV, G, G, G, -1)
        wg = world.reshape(NV, G, G, G, 3)
        w_yz = torch.softmax(self.pillar_aggregator_yz(torch.cat([lat, wg[..., 0:1]], -1)), dim=1)
        w_xz = torch.softmax(self.pillar_aggregator_xz(torch.cat([lat, wg[..., 1:2]], -1)), dim=2)
        w_xy = torch.softmax(self.pillar_aggregator_xy(torch.cat([lat, wg[..., 2:3]], -1)), dim=3)
        fp = lambda t: t.permute(0, 3, 1, 2)
        return fp((lat * w_xz).sum(2)), fp((lat * w_xy).sum(3)), fp((lat * w_yz).sum(1))      # xz, xy, yz: (NV, 512, 64, 64)

    # ---- the dense part, hand-written CUDA (tcgen05) ----
    def dense_cuda(self, latent, poses, focal, c, W, H):
        if not latent.is_cuda:
            raise RuntimeError("neo360_b200 needs CUDA tensors (no CPU fallback)")
        lib = L.load()
        dev = latent.device
        lat = latent.detach().contiguous().float()
        NV, _, lh, lw = lat.shape
        keep = []
        f = lambda t: (keep.append(t.detach().contiguous().float()) or keep[-1].data_ptr())
        p = L.NeoGridEncoderParams()
        fc = [self.depth_fc.common_branch[0], self.depth_fc.common_branch[2], self.depth_fc.depth_encoder]
        for i, m in enumerate(fc):
            p.fc_w[i], p.fc_b[i] = f(m.weight), f(m.bias)
        for pl in ("xz", "yz", "xy"):
            agg = getattr(self, f"pillar_aggregator_{pl}")
            setattr(p, f"agg_{pl}_w0", f(agg[0].weight)); setattr(p, f"agg_{pl}_b0", f(agg[0].bias))
            setattr(p, f"agg_{pl}_w1", f(agg[2].weight)); setattr(p, f"agg_{pl}_b1", f(agg[2].bias))
        need = lib.neo_grid_encoder_workspace_bytes(NV, lh, lw)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = [torch.empty(NV, 512, self.GRID, self.GRID, device=dev) for _ in range(3)]
        pose_c = poses.detach().contiguous().float()
        with torch.cuda.device(dev):
            L.check(lib.neo_grid_encoder_dense(C.byref(p), lat.data_ptr(), NV, lh, lw, int(W), int(H), pose_c.data_ptr(), float(focal[0]),
                                               float(c[0, 0]), float(c[0, 1]), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                               self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream().cuda_stream))
        return out[0], out[1], out[2]

    def forward(self, images, poses, focal, c):
        """images (NV,3,H,W), poses (NV,4,4) camera-to-world, focal (NV,), c (NV,2) -> scene_grid_xz, scene_grid_xy, scene_grid_yz (NV,128,120,160)."""
        NV, _, H, W = images.shape
        latent = self.spatial_encoder(images)
        needs_grad = torch.is_grad_enabled() and (latent.requires_grad or any(q.requires_grad for q in self.depth_fc.parameters()))
        if needs_grad:
            fxz, fxy, fyz = self.dense_torch(latent, poses.float(), focal.float(), c.float(), W, H)
        else:
            fxz, fxy, fyz = self.dense_cuda(latent, poses, focal, c, W, H)
        return self.floorplan_convnet_xz(fxz), self.floorplan_convnet_xy(fxy), self.floorplan_convnet_yz(fyz)

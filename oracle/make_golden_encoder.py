"""TEST INFRASTRUCTURE ONLY.  Pins neo360_b200.encoder.GridEncoder (framework-op form, the checker of the CUDA dense path) against the
UNMODIFIED reference `models.neo360.encoder_tp_fusion_conv.GridEncoder`, imported from /root/reference through oracle/ref_shim.py, and
writes tests/golden/encoder_reference_vectors.npz (inputs + strided samples of the reference outputs; the 15.5 M parameters are NOT
stored: both modules are constructed under the same seed and their state dicts are asserted bit-identical here).

    python oracle/make_golden_encoder.py        (CPU, this container only: /root/reference does not travel to the GPU box)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim                     # noqa: E402
from neo360_b200 import synth                   # noqa: E402
from neo360_b200.encoder import GridEncoder     # noqa: E402

SEED, W, H, NV = 5, 64, 48, 3


def reference_encoder(ns):
    import torchvision
    orig = torchvision.models.resnet34

    def nodl(*a, **k):
        k.pop("pretrained", None)
        return orig(weights=None, **{kk: v for kk, v in k.items() if kk == "norm_layer"})

    torchvision.models.resnet34 = nodl
    try:
        torch.manual_seed(SEED)
        ref = ns.neo_tp.GridEncoder(encoder_type="resnet")
    finally:
        torchvision.models.resnet34 = orig
    return ref.eval()


def main():
    ns = ref_shim.load()
    torch.set_grad_enabled(False)
    ref = reference_encoder(ns)
    torch.manual_seed(SEED)
    ours = GridEncoder().eval()
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr.keys()) == list(so.keys()), (set(sr) ^ set(so))
    worst = max(float((sr[k].float() - so[k].float()).abs().max()) for k in sr)
    assert worst == 0.0, f"seeded construction differs from the reference: {worst}"
    print("state dicts identical:", len(sr), "tensors,", sum(v.numel() for v in sr.values()), "values")

    sc = synth.make_scene((W, H), NV, (12, 16), SEED)
    g = torch.Generator().manual_seed(SEED)
    imgs = torch.rand(NV, 3, H, W, generator=g) * 2 - 1
    focal, c = sc["src_focal"], sc["src_c"]
    # the reference hard-codes device="cuda" for the image-size tensor (encoder_tp_fusion_conv.py:465): drop the device on this CPU box
    real_tensor = torch.tensor
    ns.neo_tp.torch.tensor = lambda *a, **k: real_tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    try:
        rxz, rxy, ryz = ref(imgs, sc["src_poses"], focal.clone(), c.clone())
    finally:
        ns.neo_tp.torch.tensor = real_tensor
    oxz, oxy, oyz = ours(imgs, sc["src_poses"], focal, c) if False else (None, None, None)
    lat = ours.spatial_encoder(imgs)
    fl = ours.dense_torch(lat, sc["src_poses"], focal, c, W, H)
    oxz, oxy, oyz = ours.floorplan_convnet_xz(fl[0]), ours.floorplan_convnet_xy(fl[1]), ours.floorplan_convnet_yz(fl[2])
    for name, a, b in (("xz", oxz, rxz), ("xy", oxy, rxy), ("yz", oyz, ryz), ("latent", lat, ref.spatial_encoder.latent)):
        d = float((a - b).abs().max())
        print(f"ours(torch) vs reference {name}: max |diff| {d:.3e}  (scale {float(b.abs().max()):.3f})")
        assert d <= 2e-5 * max(1.0, float(b.abs().max())), name
    out = {"cfg": np.array([SEED, W, H, NV]), "imgs": imgs.numpy(), "latent_s": ref.spatial_encoder.latent[:, ::16, ::3, ::4].numpy()}
    for name, t in (("xz", rxz), ("xy", rxy), ("yz", ryz)):
        out[f"plane_{name}_s"] = t[:, ::8, ::6, ::8].numpy()         # strided sample of the (3,128,120,160) reference planes
        out[f"floor_{name}_s"] = {"xz": fl[0], "xy": fl[1], "yz": fl[2]}[name][:, ::16, ::4, ::4].numpy()     # pillar sums (validated above via the planes)
    path = os.path.join(ROOT, "tests", "golden", "encoder_reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

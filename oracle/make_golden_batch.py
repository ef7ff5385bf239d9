"""TEST INFRASTRUCTURE ONLY.  Pins oracle.neo360_oracle.sample_training_rays (row f3) to the UNMODIFIED reference: the per-view rays come
from the reference's own `get_ray_directions` / `get_rays` (datasets/ray_utils.py, imported through oracle/ref_shim.py) and are stacked,
flattened and indexed with the statements of the training `__getitem__` (datasets/nerds360_ae.py:730-748; the dataset class itself needs the
NERDS360 files on disk, so its tensor statements are driven here with synthetic poses and images).  Writes tests/golden/train_batch_vectors.npz.

    python oracle/make_golden_batch.py        (CPU, this container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim                     # noqa: E402
from oracle import neo360_oracle as orc         # noqa: E402
from neo360_b200 import synth                   # noqa: E402

T, H, W, FOCAL, N, SEED = 6, 30, 40, 32.0, 500, 11


def main():
    ns = ref_shim.load()
    poses = torch.stack([synth.target_pose(5 * k + 2, 100)[:3, :4] for k in range(T)])
    images = torch.rand(T, H, W, 3, generator=torch.Generator().manual_seed(SEED))
    directions = ns.ray_utils.get_ray_directions(H, W, FOCAL)
    rays, rays_d, view_dirs, radii, rgbs = [], [], [], [], []
    for t in range(T):
        o, vd, rd, rad = ns.ray_utils.get_rays(directions.clone(), poses[t], output_view_dirs=True, output_radii=True)
        rays.append(o.view(-1, 3)); view_dirs.append(vd.view(-1, 3)); rays_d.append(rd.view(-1, 3)); radii.append(rad.view(-1))
        rgbs.append(images[t].flatten(0, 1))
    rays, rays_d, view_dirs, radii, rgbs = (torch.stack(x, 0) for x in (rays, rays_d, view_dirs, radii, rgbs))
    torch.manual_seed(SEED)
    pix_inds = torch.randint(0, T * H * W, (N,))
    pix_inds[:4] = torch.tensor([0, W - 1, T * H * W - 1, (H - 1) * W])
    ref = dict(rays_o=rays.reshape(-1, 3)[pix_inds], rays_d=rays_d.reshape(-1, 3)[pix_inds], viewdirs=view_dirs.reshape(-1, 3)[pix_inds],
               radii=radii.reshape(-1, 1)[pix_inds], target=rgbs.reshape(-1, 3)[pix_inds])
    o, vd, rd, rad, tgt = orc.sample_training_rays(pix_inds, H, W, FOCAL, poses, images)
    for name, mine in (("rays_o", o), ("viewdirs", vd), ("rays_d", rd), ("radii", rad), ("target", tgt)):
        d = float((mine - ref[name]).abs().max())
        print(f"oracle vs reference {name:9s} max |diff| = {d:.3g}")
        assert mine.shape == ref[name].shape and d <= 1.2e-7, name
    out = os.path.join(ROOT, "tests", "golden", "train_batch_vectors.npz")
    np.savez_compressed(out, T=T, H=H, W=W, focal=FOCAL, seed=SEED, poses=poses.numpy(), pix_inds=pix_inds.numpy(),
                        **{k: v.numpy() for k, v in ref.items()})
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

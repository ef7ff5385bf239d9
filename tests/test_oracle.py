

def test_sample_training_rays_is_gather_of_full_frames():
    """f3 oracle (nerds360_ae.py:730-748): the batch is the pix_inds rows of the stacked per-view frames; radii keep their (n,1) shape."""
    import torch
    from neo360_b200 import synth
    from oracle import neo360_oracle as orc
    Tn, H, W, f = 3, 6, 8, 6.4
    poses = torch.stack([synth.target_pose(k, 100)[:3, :4] for k in range(Tn)])
    imgs = torch.rand(Tn, H, W, 3, generator=torch.Generator().manual_seed(0))
    pix = torch.tensor([0, H * W - 1, H * W, Tn * H * W - 1, 77])
    o, vd, rd, rad, tgt = orc.sample_training_rays(pix, H, W, f, poses, imgs)
    dirs = orc.ray_directions(H, W, f)
    for q, p in enumerate(pix.tolist()):
        t, r = divmod(p, H * W)
        fo, fvd, frd, frad = orc.rays_from_pose(dirs, poses[t])
        assert torch.equal(o[q], fo[r]) and torch.equal(vd[q], fvd[r]) and torch.equal(rd[q], frd[r]) and rad[q, 0] == frad[r]
        assert torch.equal(tgt[q], imgs[t].reshape(-1, 3)[r])
    assert rad.shape == (5, 1)

"""CPU tests of the dataset-side row: the oracle's restatement of the training-batch pixel sampling and the host logic of
neo360_b200.batches.  (The oracle's render path is pinned to the reference in test_oracle_golden.py.)"""
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


def test_batches_host_side_logic():
    """neo360_b200.batches without a GPU: the pixel draw is the reference's `torch.randint(0, T*H*W, (n,))` on the same generator stream
    (nerds360_ae.py:730-732) and the device-resident views refuse CPU tensors (no CPU fallback)."""
    import pytest
    import torch
    from neo360_b200 import batches
    a = batches.draw_pix_inds(20, 48, 64, 500, torch.Generator().manual_seed(3))
    b = torch.randint(0, 20 * 48 * 64, (500,), generator=torch.Generator().manual_seed(3))
    assert torch.equal(a, b) and a.dtype == torch.int64 and int(a.max()) < 20 * 48 * 64
    assert batches.RAY_BATCH_SIZE == 500 and batches.NUM_TARGET_VIEWS == 20
    with pytest.raises(RuntimeError):
        batches.TargetViews(torch.zeros(2, 3, 4), torch.zeros(2, 4, 4, 3), 1.0)

"""TEST INFRASTRUCTURE ONLY -- mint golden vectors from the UNMODIFIED reference and pin the oracle to it.

Run in the build container (where /root/reference exists):   python oracle/make_golden.py
Writes tests/golden/*.npz (small: ray inputs + reference outputs; the big synthetic feature maps are
regenerated from seeds by `neo360_b200.synth`, and a checksum of them is stored to catch RNG drift).
Every fixture is produced by calling the reference's own functions / nn.Modules (through
oracle/ref_shim.py); the same inputs are then pushed through oracle/neo360_oracle.py and the two
are asserted equal to fp32 re-association noise.  The reference has no tests of its own (SURVEY.md
section 4), so this script IS the parity pin for the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import neo360_oracle as orc  # noqa: E402
from neo360_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def maxdiff(a, b):
    return float((a - b).abs().max())


def checksum(sc):
    return np.array([float(sc[k].double().sum()) for k in ("planes_xz", "planes_xy", "planes_yz", "latent")]
                    + [float(sc[k].double().abs().sum()) for k in ("planes_xz", "latent")])


class RandQueue:
    """Replays given tensors for successive torch.rand calls (helper.py:50,199)."""

    def __init__(self, items):
        self.items = list(items)
        self.orig = torch.rand

    def __enter__(self):
        def fake(*a, **k):
            t = self.items.pop(0)
            shape = tuple(a[0]) if len(a) == 1 and not isinstance(a[0], int) else tuple(a)
            assert tuple(t.shape) == shape, (t.shape, shape)
            return t.clone()
        torch.rand = fake
        return self

    def __exit__(self, *e):
        torch.rand = self.orig
        assert not self.items


def stage_kats(ns, out):
    H = ns.neo_helper
    g = torch.Generator().manual_seed(7)
    # hand-checkable ray from SURVEY.md section 8(c) + random rays inside the sphere
    o = torch.cat([torch.tensor([[0.3, 0.2, 0.4]]), (torch.rand(15, 3, generator=g) - 0.5) * 1.0])
    d = torch.cat([torch.tensor([[0.0, 0.6, -0.8]]), torch.randn(15, 3, generator=g)])
    d = d / d.norm(dim=-1, keepdim=True)
    far = H.intersect_sphere(o, d)
    assert abs(float(far[0]) - 1.0660254) < 1e-6
    near = torch.full_like(far, 1e-4)
    fg_t, fg_p = H.sample_along_rays(o, d, 4, near, far, False, False, True)
    bg_s, bg_p, bg_l = H.sample_along_rays(o, d, 4, near, far, False, False, False, far_uncontracted=3)
    out.update(kat_o=o, kat_d=d, kat_far=far, kat_fg_t=fg_t, kat_fg_p=fg_p, kat_bg_s=bg_s, kat_bg_p=bg_p,
               kat_bg_l=bg_l)
    assert maxdiff(orc.intersect_sphere(o, d), far) == 0
    t2, p2 = orc.sample_fg(o, d, 4, near, far)
    assert maxdiff(t2, fg_t) == 0 and maxdiff(p2, fg_p) == 0
    s2, bp2, bl2 = orc.sample_bg(o, d, 4, far)
    assert maxdiff(s2, bg_s) == 0 and maxdiff(bp2, bg_p) < 1e-6 and maxdiff(bl2, bg_l) == 0
    # randomized sampling with injected uniforms
    u = torch.rand(16, 5, generator=g)
    with RandQueue([u]):
        rt, _ = H.sample_along_rays(o, d, 4, near, far, True, False, True)
    with RandQueue([u]):
        rs, rp, rl = H.sample_along_rays(o, d, 4, near, far, True, False, False, far_uncontracted=3)
    assert maxdiff(orc.sample_fg(o, d, 4, near, far, u)[0], rt) == 0
    assert maxdiff(orc.sample_bg(o, d, 4, far, 3.0, u)[0], rs) == 0
    out.update(kat_u=u, kat_fg_t_rand=rt, kat_bg_s_rand=rs, kat_bg_l_rand=rl)
    # compositing
    rgb = torch.rand(16, 5, 3, generator=g)
    sig = torch.rand(16, 5, 1, generator=g) * 3
    rgb[0] = torch.tensor([[.1, .2, .3], [.4, .5, .6], [.7, .8, .9], [.2, .2, .2], [.9, .1, .5]])
    sig[0, :, 0] = torch.tensor([.5, 1, 2, .1, 3])
    fc = H.volumetric_rendering(rgb, sig, fg_t, d, False, True, t_far=far, out_depth=True)
    bc = H.volumetric_rendering(rgb, sig, bg_s, d, False, False, out_depth=True)
    assert abs(float(fc[0][0, 0]) - .2903506) < 1e-6 and abs(float(bc[4][0]) - .4016956) < 1e-6
    of = orc.composite(rgb, sig, fg_t, d, False, True, far)
    ob = orc.composite(rgb, sig, bg_s, d, False, False)
    for a, b in zip(of, fc):
        assert maxdiff(a, b) == 0
    for a, b in zip((ob[0], ob[1], ob[2], ob[4]), (bc[0], bc[1], bc[2], bc[4])):
        assert maxdiff(a, b) == 0
    out.update(kat_rgb=rgb, kat_sig=sig, kat_fg_comp=fc[0], kat_fg_acc=fc[1], kat_fg_w=fc[2], kat_fg_lam=fc[3],
               kat_fg_depth=fc[4], kat_bg_comp=bc[0], kat_bg_acc=bc[1], kat_bg_w=bc[2], kat_bg_depth=bc[4])
    # inverse CDF: ascending (fg) and descending (bg, quirk Q17) bins, deterministic and randomized
    mids = 0.5 * (fg_t[..., 1:] + fg_t[..., :-1])
    w = fc[2][..., 1:-1]
    pf = H.sorted_piecewise_constant_pdf(mids, w, 6, False)
    bm = 0.5 * (bg_s[..., 1:] + bg_s[..., :-1])
    pb = H.sorted_piecewise_constant_pdf(bm, bc[2][..., 1:-1], 6, False)
    u6 = torch.rand(16, 6, generator=g)
    with RandQueue([u6]):
        pr = H.sorted_piecewise_constant_pdf(mids, w, 6, True)
    assert maxdiff(orc.piecewise_constant_pdf(mids, w, 6), pf) == 0
    assert maxdiff(orc.piecewise_constant_pdf(bm, bc[2][..., 1:-1], 6), pb) == 0
    assert maxdiff(orc.piecewise_constant_pdf(mids, w, 6, u6), pr) == 0
    kp = H.sorted_piecewise_constant_pdf(mids[:1], torch.tensor([[.1, .5, .2]]), 6, False)
    assert maxdiff(kp, torch.tensor([[.1333407, .4317998, .5170738, .6023479, .7195997, .9327847]])) < 1e-6
    out.update(kat_pdf_fg=pf, kat_pdf_bg=pb, kat_u6=u6, kat_pdf_rand=pr)
    # pos-enc + camera transforms
    x = torch.randn(2, 5, 4, generator=g)
    pe = H.pos_enc(x, 0, 10)
    assert maxdiff(orc.pos_enc(x, 0, 10), pe) == 0
    poses = torch.stack([synth.look_at_pose(30.0 + 120 * v, 0.3, 0.8) for v in range(3)])
    pts = torch.randn(1, 11, 3, generator=g)
    wc = ns.neo_util.world2camera(pts, poses, 3)
    assert maxdiff(orc.world2camera(pts[0], poses), wc) == 0
    wd = ns.neo_util.world2camera_viewdirs(pts, poses, 3)
    assert maxdiff(orc.world2camera_dirs(pts[0], poses), wd) == 0
    out.update(kat_pe_in=x, kat_pe=pe, kat_poses=poses, kat_pts=pts[0], kat_w2c=wc, kat_w2c_dirs=wd)
    # ray generation
    dirs = ns.ray_utils.get_ray_directions(6, 8, 6.4)
    ro, vd, rd, rad = ns.ray_utils.get_rays(dirs.clone(), poses[0][:3, :4], output_view_dirs=True, output_radii=True)
    oo, ovd, ord_, orad = orc.rays_from_pose(orc.ray_directions(6, 8, 6.4), poses[0][:3, :4])
    assert maxdiff(oo, ro) == 0 and maxdiff(ovd, vd) < 1e-7 and maxdiff(ord_, rd) < 1e-7 and maxdiff(orad, rad) < 1e-7
    out.update(kat_ray_o=ro, kat_ray_vd=vd, kat_ray_d=rd, kat_ray_radii=rad)
    print("stage KATs: oracle == reference")


def scene_from_synth(sc):
    W, H = sc["img_wh"]
    return orc.Scene(sc["planes_xz"], sc["planes_xy"], sc["planes_yz"], sc["latent"], sc["src_poses"],
                     float(sc["src_focal"][0]), float(sc["src_c"][0, 0]), float(sc["src_c"][0, 1]), W, H)


def e2e(ns, out, tag, img_wh, plane_hw, B, nc, nf, seed):
    W, H = img_wh
    sc = synth.make_scene(img_wh, 3, plane_hw, seed)
    P = synth.make_mlp_params(seed)
    net = ref_shim.make_reference_nerf_tp(ns, nc, nf, 3, seed)
    missing, unexpected = net.load_state_dict(P, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing), (missing, unexpected)
    ref_shim.bypass_encoder(net, sc["planes_xz"], sc["planes_xy"], sc["planes_yz"], sc["latent"])
    # rays: a contiguous run of pixels of one turntable frame (as render_rays_test slices them)
    pose = synth.target_pose(3, 100)
    dirs = ns.ray_utils.get_ray_directions(H, W, 0.8 * W)
    ro, vd, rd = ns.ray_utils.get_rays(dirs, pose[:3, :4], output_view_dirs=True)
    start = (H // 2) * W + W // 3
    sel = slice(start, start + B)
    rays = {"rays_o": ro[sel].contiguous(), "rays_d": rd[sel].contiguous(), "viewdirs": vd[sel].contiguous(),
            "src_imgs": torch.zeros(3, 3, H, W), "src_poses": sc["src_poses"], "src_focal": sc["src_focal"],
            "src_c": sc["src_c"]}
    osc = scene_from_synth(sc)
    with torch.no_grad():
        ev = net(rays, False, False, 0.2, 3.0, out_depth=True)
        tr = net(rays, False, True, 0.2, 3.0, out_depth=False)
        g = torch.Generator().manual_seed(99 + seed)
        rnd = {"fg0": torch.rand(B, nc + 1, generator=g), "bg0": torch.rand(B, nc + 1, generator=g),
               "fg1": torch.rand(B, nf, generator=g), "bg1": torch.rand(B, nf, generator=g)}
        with RandQueue([rnd["fg0"], rnd["bg0"], rnd["fg1"], rnd["bg1"]]):
            rr = net(rays, True, False, 0.2, 3.0, out_depth=True)
        o_ev, aux = orc.render(rays, osc, P, nc, nf, False, True, return_aux=True)
        o_tr = orc.render(rays, osc, P, nc, nf, True, False)
        o_rr = orc.render(rays, osc, P, nc, nf, False, True, rand=rnd)
        o_at = orc.render(rays, osc, P, nc, nf, False, True, lookup_impl="aten")
    worst = 0.0
    for name, got, ref in (("eval", o_ev, ev), ("train", o_tr, tr), ("rand", o_rr, rr), ("aten", o_at, ev)):
        for lvl in range(2):
            for j, (a, b) in enumerate(zip(got[lvl], ref[lvl])):
                dd = maxdiff(a, b)
                worst = max(worst, dd)
                assert dd < 5e-4, (tag, name, lvl, j, dd)  # fp32 re-association noise through the gained MLP (bg resampling is discontinuous at CDF bracket edges, quirk Q17)
    print(f"e2e[{tag}]: oracle vs reference max|diff| = {worst:.3e}")
    out.update({f"{tag}_cfg": np.array([W, H, plane_hw[0], plane_hw[1], B, nc, nf, seed, start]),
                f"{tag}_checksum": checksum(sc), f"{tag}_rays_o": rays["rays_o"], f"{tag}_rays_d": rays["rays_d"],
                f"{tag}_viewdirs": rays["viewdirs"]})
    names_ev = ("comp_rgb", "fg_rgb", "bg_rgb", "fg_acc", "bg_lambda", "depth")
    names_tr = ("comp_rgb", "fg_w", "bg_w", "fg_sdist", "bg_sdist", "bg_acc")
    for lvl in range(2):
        for n, v in zip(names_ev, ev[lvl]):
            out[f"{tag}_eval{lvl}_{n}"] = v
        for n, v in zip(names_tr, tr[lvl]):
            out[f"{tag}_train{lvl}_{n}"] = v
        for n, v in zip(names_ev, rr[lvl]):
            out[f"{tag}_rand{lvl}_{n}"] = v
        for k in ("fg_t", "bg_s", "fg_sigma", "bg_sigma", "fg_rgb", "bg_rgb"):
            out[f"{tag}_aux{lvl}_{k}"] = aux[lvl][k]
    for k, v in rnd.items():
        out[f"{tag}_u_{k}"] = v


def vanilla(ns, path):
    """Golden vectors for the vanilla-NeRF renderer (row a17) from the unmodified reference NeRF module."""
    from oracle import vanilla_oracle as vor
    out = {}
    for tag, (W, H, B, nc, nf, seed) in {"v_tiny": (64, 48, 96, 16, 8, 0), "v_cfg1": (64, 64, 1024, 64, 64, 1)}.items():
        P = synth.make_vanilla_params(seed)
        torch.manual_seed(seed)
        net = ns.van_model.NeRF(num_coarse_samples=nc, num_fine_samples=nf).eval()
        missing, unexpected = net.load_state_dict(P, strict=True)
        pose = synth.target_pose(5, 100)
        dirs = ns.ray_utils.get_ray_directions(H, W, 0.8 * W)
        ro, vd, rd = ns.ray_utils.get_rays(dirs, pose[:3, :4], output_view_dirs=True)
        g = torch.Generator().manual_seed(50 + seed)
        sel = torch.randperm(H * W, generator=g)[:B]
        # the dataset hands un-normalised rays_d to the vanilla model (datasets/nerds360.py); exercise |rays_d| != 1 (quirk Q15)
        scale = 0.5 + torch.rand(B, 1, generator=g)
        rays = {"rays_o": ro[sel].contiguous(), "rays_d": (rd[sel] * scale).contiguous(), "viewdirs": vd[sel].contiguous()}
        near, far = 0.2, 3.0
        with torch.no_grad():
            ev = net(rays, False, True, near, far)
            rnd = {"u0": torch.rand(B, nc + 1, generator=g), "u1": torch.rand(B, nf, generator=g)}
            with RandQueue([rnd["u0"], rnd["u1"]]):
                rr = net(rays, True, False, near, far)
            o_ev, aux = vor.render(rays, P, nc, nf, near, far, True, return_aux=True)
            o_rr = vor.render(rays, P, nc, nf, near, far, False, rand=rnd)
        worst = 0.0
        for got, ref in ((o_ev, ev), (o_rr, rr)):
            for lvl in range(2):
                for a, b in zip(got[lvl], ref[lvl]):
                    worst = max(worst, maxdiff(a, b))
        assert worst < 2e-4, worst
        print(f"vanilla[{tag}]: oracle vs reference max|diff| = {worst:.3e}")
        out.update({f"{tag}_cfg": np.array([W, H, B, nc, nf, seed]), f"{tag}_rays_o": rays["rays_o"], f"{tag}_rays_d": rays["rays_d"],
                    f"{tag}_viewdirs": rays["viewdirs"], f"{tag}_u0": rnd["u0"], f"{tag}_u1": rnd["u1"]})
        for lvl in range(2):
            for n_, a, b in zip(("rgb", "acc", "depth"), ev[lvl], rr[lvl]):
                out[f"{tag}_eval{lvl}_{n_}"] = a
                out[f"{tag}_rand{lvl}_{n_}"] = b
            for k in ("t", "sigma", "rgb"):
                out[f"{tag}_aux{lvl}_{k}"] = aux[lvl][k] if tag == "v_tiny" else aux[lvl][k][:32]
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def mip360(ns, path):
    """Golden vectors for the Mip-NeRF 360 renderer (row a18) from the unmodified reference MipNeRF360 module."""
    from oracle import mip_oracle as mor
    from neo360_b200.mip_basis import POS_BASIS_T
    ref_basis = ns.mip_helper.generate_basis("icosahedron", 2)
    assert maxdiff(POS_BASIS_T, ref_basis) == 0, "embedded basis != reference generate_basis"
    out = {}
    for tag, (W, H, B, npp, nn_, seed) in {"m_tiny": (64, 48, 40, 16, 8, 0), "m_small": (64, 48, 96, 32, 16, 1)}.items():
        P = synth.make_mip_params(seed)
        torch.manual_seed(seed)
        net = ns.mip_model.MipNeRF360(num_prop_samples=npp, num_nerf_samples=nn_).eval()
        net.load_state_dict(P, strict=True)
        pose = synth.target_pose(9, 100)
        dirs = ns.ray_utils.get_ray_directions(H, W, 0.8 * W)
        ro, vd, rd, radii = ns.ray_utils.get_rays(dirs, pose[:3, :4], output_view_dirs=True, output_radii=True)
        g = torch.Generator().manual_seed(70 + seed)
        sel = torch.randperm(H * W, generator=g)[:B]
        batch = {"rays_o": ro[sel].contiguous(), "rays_d": rd[sel].contiguous(), "viewdirs": vd[sel].contiguous(),
                 "radii": radii[sel].reshape(-1, 1).contiguous()}
        near, far = 0.2, 6.0
        with torch.no_grad():
            ren, hist = net(batch, 1.0, False, False, near, far)
            jit = [torch.rand(B, 1, generator=g) for _ in range(3)]
            with RandQueue(jit):
                ren_r, hist_r = net(batch, 0.5, True, False, near, far)
            o_ren, o_hist = mor.render(batch, P, POS_BASIS_T, npp, nn_, near, far, 1.0)
            o_ren_r, o_hist_r = mor.render(batch, P, POS_BASIS_T, npp, nn_, near, far, 0.5, rand=jit)
        worst = 0.0
        for a, b, c, e in ((o_ren, o_hist, ren, hist), (o_ren_r, o_hist_r, ren_r, hist_r)):
            for lvl in range(3):
                worst = max(worst, maxdiff(a[lvl]["rgb"], c[lvl]["rgb"]))
                for k in ("density", "rgb", "sdist", "weights"):
                    worst = max(worst, maxdiff(b[lvl][k], e[lvl][k]))
        assert worst < 5e-4, worst
        print(f"mip360[{tag}]: oracle vs reference max|diff| = {worst:.3e}")
        out.update({f"{tag}_cfg": np.array([W, H, B, npp, nn_, seed]), f"{tag}_near_far": np.array([near, far])})
        for k, v in batch.items():
            out[f"{tag}_{k}"] = v
        for i in range(3):
            out[f"{tag}_jit{i}"] = jit[i]
            out[f"{tag}_eval{i}_rgb"] = ren[i]["rgb"]
            out[f"{tag}_rand{i}_rgb"] = ren_r[i]["rgb"]
            for k in ("density", "rgb", "sdist", "weights"):
                out[f"{tag}_hist{i}_{k}"] = hist[i][k]
            out[f"{tag}_rhist{i}_sdist"] = hist_r[i]["sdist"]
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    ns = ref_shim.load()
    torch.set_grad_enabled(False)
    out = {}
    stage_kats(ns, out)
    # tiny: chunk of 48 rays, 16+8 samples; small: 160 rays, 32+16 (exercises Q1 with B not dividing N)
    e2e(ns, out, "tiny", (64, 48), (24, 32), 48, 16, 8, 0)
    e2e(ns, out, "small", (96, 64), (30, 40), 160, 32, 16, 1)
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "neo360_reference_vectors.npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")
    vanilla(ns, os.path.join(GOLD, "vanilla_reference_vectors.npz"))
    mip360(ns, os.path.join(GOLD, "mip360_reference_vectors.npz"))


if __name__ == "__main__":
    main()

"""CPU: the oracle (oracle/neo360_oracle.py) against the golden vectors minted from the UNMODIFIED
reference by oracle/make_golden.py.  These run everywhere (no /root/reference needed)."""
import numpy as np
import pytest
import torch

from neo360_b200 import synth
from oracle import neo360_oracle as orc

T = lambda a: torch.from_numpy(np.asarray(a))


def md(a, b):
    return float((a - T(b)).abs().max())


def test_stage_known_answers(golden):
    g = golden
    o, d = T(g["kat_o"]), T(g["kat_d"])
    far = orc.intersect_sphere(o, d)
    assert md(far, g["kat_far"]) == 0
    assert abs(float(far[0]) - 1.0660254) < 1e-6          # SURVEY.md 8(c) hand vector
    near = torch.full_like(far, 1e-4)
    t, p = orc.sample_fg(o, d, 4, near, far)
    assert md(t, g["kat_fg_t"]) == 0 and md(p, g["kat_fg_p"]) == 0
    s, bp, bl = orc.sample_bg(o, d, 4, far)
    assert md(s, g["kat_bg_s"]) == 0 and md(bp, g["kat_bg_p"]) < 1e-6 and md(bl, g["kat_bg_l"]) == 0
    u = T(g["kat_u"])
    assert md(orc.sample_fg(o, d, 4, near, far, u)[0], g["kat_fg_t_rand"]) == 0
    s_r, _, l_r = orc.sample_bg(o, d, 4, far, 3.0, u)
    assert md(s_r, g["kat_bg_s_rand"]) == 0
    rgb, sig = T(g["kat_rgb"]), T(g["kat_sig"])
    fc = orc.composite(rgb, sig, t, d, False, True, far)
    for a, k in zip(fc, ("kat_fg_comp", "kat_fg_acc", "kat_fg_w", "kat_fg_lam", "kat_fg_depth")):
        assert md(a, g[k]) == 0
    bc = orc.composite(rgb, sig, s, d, False, False)
    for a, k in zip((bc[0], bc[1], bc[2], bc[4]), ("kat_bg_comp", "kat_bg_acc", "kat_bg_w", "kat_bg_depth")):
        assert md(a, g[k]) == 0
    mids = 0.5 * (t[..., 1:] + t[..., :-1])
    assert md(orc.piecewise_constant_pdf(mids, fc[2][..., 1:-1], 6), g["kat_pdf_fg"]) == 0
    bm = 0.5 * (s[..., 1:] + s[..., :-1])
    assert md(orc.piecewise_constant_pdf(bm, bc[2][..., 1:-1], 6), g["kat_pdf_bg"]) == 0      # quirk Q17
    assert md(orc.piecewise_constant_pdf(mids, fc[2][..., 1:-1], 6, T(g["kat_u6"])), g["kat_pdf_rand"]) == 0
    assert md(orc.pos_enc(T(g["kat_pe_in"]), 0, 10), g["kat_pe"]) == 0
    poses = T(g["kat_poses"])
    assert md(orc.world2camera(T(g["kat_pts"]), poses), g["kat_w2c"]) == 0
    assert md(orc.world2camera_dirs(T(g["kat_pts"]), poses), g["kat_w2c_dirs"]) == 0
    ro, vd, rd, rad = orc.rays_from_pose(orc.ray_directions(6, 8, 6.4), poses[0][:3, :4])
    assert md(ro, g["kat_ray_o"]) == 0 and md(vd, g["kat_ray_vd"]) < 1e-7 and md(rad, g["kat_ray_radii"]) < 1e-7


def _load_case(g, tag):
    W, H, hp, wp, B, nc, nf, seed, start = [int(x) for x in g[f"{tag}_cfg"]]
    sc = synth.make_scene((W, H), 3, (hp, wp), seed)
    chk = np.array([float(sc[k].double().sum()) for k in ("planes_xz", "planes_xy", "planes_yz", "latent")]
                   + [float(sc[k].double().abs().sum()) for k in ("planes_xz", "latent")])
    assert np.allclose(chk, g[f"{tag}_checksum"], rtol=1e-9), "synthetic scene RNG drifted; re-mint the goldens"
    P = synth.make_mlp_params(seed)
    rays = {"rays_o": T(g[f"{tag}_rays_o"]), "rays_d": T(g[f"{tag}_rays_d"]), "viewdirs": T(g[f"{tag}_viewdirs"])}
    osc = orc.Scene(sc["planes_xz"], sc["planes_xy"], sc["planes_yz"], sc["latent"], sc["src_poses"],
                    float(sc["src_focal"][0]), float(sc["src_c"][0, 0]), float(sc["src_c"][0, 1]), W, H)
    return rays, osc, P, nc, nf


EV = ("comp_rgb", "fg_rgb", "bg_rgb", "fg_acc", "bg_lambda", "depth")
TR = ("comp_rgb", "fg_w", "bg_w", "fg_sdist", "bg_sdist", "bg_acc")


@pytest.mark.parametrize("tag", ["tiny", "small"])
def test_end_to_end_vs_reference_vectors(golden, tag):
    g = golden
    rays, osc, P, nc, nf = _load_case(g, tag)
    with torch.no_grad():
        ev = orc.render(rays, osc, P, nc, nf, False, True)
        tr = orc.render(rays, osc, P, nc, nf, True, False)
        rnd = {k: T(g[f"{tag}_u_{k}"]) for k in ("fg0", "bg0", "fg1", "bg1")}
        rr = orc.render(rays, osc, P, nc, nf, False, True, rand=rnd)
    tol = 5e-4  # fp32 re-association noise (oracle pin tolerance, see oracle/make_golden.py)
    for lvl in range(2):
        for n, v in zip(EV, ev[lvl]):
            assert md(v, g[f"{tag}_eval{lvl}_{n}"]) < tol, (lvl, n)
        for n, v in zip(TR, tr[lvl]):
            assert md(v, g[f"{tag}_train{lvl}_{n}"]) < tol, (lvl, n)
        for n, v in zip(EV, rr[lvl]):
            assert md(v, g[f"{tag}_rand{lvl}_{n}"]) < tol, (lvl, n)


def test_chunked_render_matches_single_chunk(golden):
    """render_rays_test's chunk loop (model.py:861-896): chunk == B must equal the unchunked call."""
    rays, osc, P, nc, nf = _load_case(golden, "tiny")
    with torch.no_grad():
        a = orc.render_chunked(rays, osc, P, nc, nf, chunk=rays["rays_o"].shape[0])
    assert md(a["comp_rgb"], golden["tiny_eval1_comp_rgb"]) < 5e-4
    assert md(a["depth"], golden["tiny_eval1_depth"]) < 5e-4


# ---------------- vanilla NeRF (row a17; BASELINE configs[0] is the CPU-runnable plumbing case) ----------------

@pytest.fixture(scope="module")
def vgolden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vanilla_reference_vectors.npz"))


@pytest.mark.parametrize("tag", ["v_tiny", "v_cfg1"])
def test_vanilla_oracle_vs_reference_vectors(vgolden, tag):
    """v_cfg1 = BASELINE configs[0]: 64x64 crop, 1024 rays, 64+64 samples, on CPU."""
    from oracle import vanilla_oracle as vor
    g = vgolden
    W, H, B, nc, nf, seed = [int(x) for x in g[f"{tag}_cfg"]]
    P = synth.make_vanilla_params(seed)
    rays = {k: T(g[f"{tag}_{k}"]) for k in ("rays_o", "rays_d", "viewdirs")}
    with torch.no_grad():
        ev = vor.render(rays, P, nc, nf, 0.2, 3.0, True)
        rr = vor.render(rays, P, nc, nf, 0.2, 3.0, False, rand={"u0": T(g[f"{tag}_u0"]), "u1": T(g[f"{tag}_u1"])})
    for lvl in range(2):
        for n_, a, b in zip(("rgb", "acc", "depth"), ev[lvl], rr[lvl]):
            assert md(a, g[f"{tag}_eval{lvl}_{n_}"]) < 1e-5, (lvl, n_)
            assert md(b, g[f"{tag}_rand{lvl}_{n_}"]) < 1e-5, (lvl, n_)


# ---------------- Mip-NeRF 360 (row a18) ----------------

@pytest.mark.parametrize("tag", ["m_tiny", "m_small"])
def test_mip360_oracle_vs_reference_vectors(tag):
    import os
    from oracle import mip_oracle as mor
    from neo360_b200.mip_basis import POS_BASIS_T
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mip360_reference_vectors.npz"))
    W, H, B, npp, nn_, seed = [int(x) for x in g[f"{tag}_cfg"]]
    near, far = [float(x) for x in g[f"{tag}_near_far"]]
    P = synth.make_mip_params(seed)
    batch = {k: T(g[f"{tag}_{k}"]) for k in ("rays_o", "rays_d", "viewdirs", "radii")}
    with torch.no_grad():
        ren, hist = mor.render(batch, P, POS_BASIS_T, npp, nn_, near, far, 1.0)
        ren_r, hist_r = mor.render(batch, P, POS_BASIS_T, npp, nn_, near, far, 0.5, rand=[T(g[f"{tag}_jit{i}"]) for i in range(3)])
    for i in range(3):
        assert md(ren[i]["rgb"], g[f"{tag}_eval{i}_rgb"]) < 1e-4
        assert md(ren_r[i]["rgb"], g[f"{tag}_rand{i}_rgb"]) < 1e-4
        for k in ("density", "rgb", "sdist", "weights"):
            assert md(hist[i][k], g[f"{tag}_hist{i}_{k}"]) < 1e-4, (i, k)
        assert md(hist_r[i]["sdist"], g[f"{tag}_rhist{i}_sdist"]) < 1e-5


def test_training_batch_golden():
    """Row f3: the oracle's pixel sampling equals the reference's get_rays + stack + index (oracle/make_golden_batch.py)."""
    import os
    import numpy as np
    import torch
    from oracle import neo360_oracle as orc
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_batch_vectors.npz"))
    pix, poses = torch.from_numpy(z["pix_inds"]), torch.from_numpy(z["poses"])
    images = torch.rand(int(z["T"]), int(z["H"]), int(z["W"]), 3, generator=torch.Generator().manual_seed(int(z["seed"])))
    o, vd, rd, rad, tgt = orc.sample_training_rays(pix, int(z["H"]), int(z["W"]), float(z["focal"]), poses, images)
    for name, mine in (("rays_o", o), ("viewdirs", vd), ("rays_d", rd), ("radii", rad), ("target", tgt)):
        assert mine.shape == z[name].shape and float((mine - torch.from_numpy(z[name])).abs().max()) <= 1.2e-7, name

"""Cycle accounting of the TC field kernel (neo_tc_debug): where each warp role spends its time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from neo360_b200 import NeRF_TP, _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ablate = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
sc, P = bench.build_scene_cpu()
net = NeRF_TP(num_coarse_samples=128, num_fine_samples=64, precision="tc").eval()
net.load_state_dict(P); net = net.to(dev)
net.set_scene(*[sc[k].to(dev) for k in ("planes_xz", "planes_xy", "planes_yz", "latent", "src_poses", "src_focal", "src_c")], sc["img_wh"])
o, d = bench.frame_rays_cpu(0)
rays = {"rays_o": o[:n].to(dev), "rays_d": d[:n].to(dev), "viewdirs": d[:n].to(dev)}
lib = L.load()
with torch.no_grad():
    net.render_rays_test(rays, chunk=1024)
    lib.neo_tc_ablate(ablate)
    buf = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
    lib.neo_tc_debug(buf.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net.render_rays_test(rays, chunk=1024); e1.record(); torch.cuda.synchronize()
    lib.neo_tc_debug(None)
ms = e0.elapsed_time(e1)
b = buf.view(148, 64).double().cpu()
names = ["P pts", "P wait ENC_FREE", "P geometry", "P bar", "P wait G_FREE", "P gather", "M wait ENC_READY", "M wait H_READY", "M issue",
         "E wait ACC", "E wait G", "E work", "E head", "M wait G_READY"]
# counters are overwritten by each of the 4 field launches: they hold the LAST launch (bg fine, N=193)
tiles = ((n + 31) // 32) * ((193 + 3) // 4)
halfjobs_per_cta = tiles * 6 / 148
lib.neo_tc_ablate(0)
print(f"ablate={ablate}: {n} rays, frame step {ms:.1f} ms; last launch: {tiles} tiles, {halfjobs_per_cta:.0f} half-jobs per CTA")
for i, nm in enumerate(names):
    print(f"  {nm:18s} {b[:, i].mean() / halfjobs_per_cta:9.0f} cycles / half-job   (total {b[:, i].mean() / 1e6:8.2f} Mcyc)")

jobs_per_bin = halfjobs_per_cta / 6
for nm, off in (("E wait G by job (v*2+h)", 16), ("P wait ENC_FREE by job", 24), ("P wait G_FREE by job", 32), ("P gather by job", 48)):
    print(f"  {nm:26s} " + " ".join(f"{b[:, off + i].mean() / jobs_per_bin:7.0f}" for i in range(6)))
print("  E wait ACC by layer,block   " + " ".join(f"{b[:, 40 + i].mean() / halfjobs_per_cta:6.0f}" for i in range(8)))

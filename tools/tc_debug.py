"""Cycle accounting of the TC field kernel (neo_tc_debug): where each warp role spends its time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from neo360_b200 import NeRF_TP, _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
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
    buf = torch.zeros(148 * 64 + 5 * 1024, dtype=torch.int64, device=dev)
    lib.neo_tc_debug(buf.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net.render_rays_test(rays, chunk=1024); e1.record(); torch.cuda.synchronize()
    mlp = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    if mlp != 3:      # make another branch the LAST launch: re-evaluate its field on the frame's own t-values
        lib.neo_tc_debug(None)
        net.forward(rays, False, False, None, None, out_depth=True, chunk=1024, debug=True)
        dbg = net.last_debug
        tk = ("fg_t", "bg_s")[mlp & 1]
        tv = dbg[tk][mlp >> 1]
        from neo360_b200 import ops
        far = ops.intersect_sphere(rays["rays_o"], rays["rays_d"])
        buf.zero_()
        lib.neo_tc_debug(buf.data_ptr())
        net.field_eval(rays, far, tv, mlp, chunk=1024)
        torch.cuda.synchronize()
    lib.neo_tc_debug(None)
ms = e0.elapsed_time(e1)
tr = buf[148 * 64:].view(5, 128, 8).cpu()
b = buf[:148 * 64].view(148, 64).double().cpu()
names = ["W wait INFO_READY", "G wait slots", "G geometry", "-", "-", "-", "M wait ENC_READY", "M wait H_READY", "M issue",
         "E wait ACC", "E wait G", "E work", "E head", "M wait windows"]
# counters are overwritten by each of the 4 field launches: they hold the LAST launch (bg fine, N=193)
NN = 193 if (int(sys.argv[4]) if len(sys.argv) > 4 else 3) >= 2 else 129
tiles = ((n + 31) // 32) * ((NN + 3) // 4)
halfjobs_per_cta = tiles * 6 / 148
print(f"{n} rays, frame step {ms:.1f} ms; last launch: {tiles} tiles, {halfjobs_per_cta:.0f} half-jobs per CTA")
for i, nm in enumerate(names):
    print(f"  {nm:18s} {b[:, i].mean() / halfjobs_per_cta:9.0f} cycles / half-job   (total {b[:, i].mean() / 1e6:8.2f} Mcyc)")

jobs_per_bin = halfjobs_per_cta / 6
for i, nm in enumerate(["W enumerate", "W bar.sync 2", "W turn", "W acquire + TMA", "W weights + fence", "W latent windows (count)"]):
    print(f"  {nm:36s} {b[:, 24 + i].mean() / halfjobs_per_cta:9.1f} per half-job")
print("  E wait ACC by layer,block   " + " ".join(f"{b[:, 40 + i].mean() / halfjobs_per_cta:6.0f}" for i in range(8)))

# event trace of CTA 0 (last launch), cycles relative to the MMA warp's first stamp
t0 = int(tr[3, 0, 0])
print("job | G0: start done | W0 (latent, xz): info cnt - done nwin | W1 (xy, yz): info cnt - done nwin | MMA: enc cnt win l3 end nwin")
for j in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8, int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    r = lambda role, k: int(tr[role, j, k]) - t0 if int(tr[role, j, k]) else -1
    print(f"{j:3d} | {r(0,0):7d} {r(0,1):7d} | {r(1,0):7d} {r(1,1):7d} {r(1,2):7d} {r(1,3):7d} {int(tr[1,j,7]):2d} | "
          f"{r(2,0):7d} {r(2,1):7d} {r(2,2):7d} {r(2,3):7d} {int(tr[2,j,7]):2d} | {r(3,0):7d} {r(3,1):7d} {r(3,2):7d} {r(3,3):7d} {r(3,4):7d} {int(tr[3,j,7]):2d}")

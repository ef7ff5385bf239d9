"""Per-scene preparation cost (neo_scene_create): H2D of the raw maps, the first (cold: module load, attribute setup) and the following
(warm) builds.  Usage on a GPU box:  python tools/time_scene.py [tc|fp32]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as Bm
from neo360_b200 import NeRF_TP

prec = sys.argv[1] if len(sys.argv) > 1 else "tc"
dev = torch.device("cuda:0")
sc, P = Bm.build_scene_cpu()
net = NeRF_TP(num_coarse_samples=Bm.N_COARSE, num_fine_samples=Bm.N_FINE, num_src_views=Bm.NV, precision=prec).eval()
net.load_state_dict(P)
net = net.to(dev)
keys = ("planes_xz", "planes_xy", "planes_yz", "latent", "src_poses", "src_focal", "src_c")
torch.cuda.synchronize()
t0 = time.perf_counter()
devt = [sc[k].to(dev) for k in keys]
torch.cuda.synchronize()
print(f"h2d (pageable, {sum(t.numel() * 4 for t in devt) / 1e6:.0f} MB): {(time.perf_counter() - t0) * 1e3:.1f} ms")
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net.set_scene(*devt, sc["img_wh"])
    torch.cuda.synchronize()
    print(f"set_scene[{prec}] call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms")

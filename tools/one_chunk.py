"""Where does a 1024-ray drop-in call spend its time?  Host time per call (no sync), device time per call (CUDA events), for the
reference's chunk loop.  Under `ncu --metrics gpu__time_duration.sum` the same script gives the per-kernel durations of one chunk.
Usage on a GPU box:  python tools/one_chunk.py [n_calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as Bm
from neo360_b200 import NeRF_TP

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
sc, P = Bm.build_scene_cpu()
net = NeRF_TP(num_coarse_samples=Bm.N_COARSE, num_fine_samples=Bm.N_FINE, num_src_views=Bm.NV, precision="tc").eval()
net.load_state_dict(P)
net = net.to(dev)
net.set_scene(*[sc[k].to(dev) for k in ("planes_xz", "planes_xy", "planes_yz", "latent", "src_poses", "src_focal", "src_c")], sc["img_wh"])
o, d = Bm.frame_rays_cpu(0)
o, d = o.to(dev), d.to(dev)
C = Bm.CHUNK
with torch.no_grad():
    for i in range(3):
        net({"rays_o": o[:C], "rays_d": d[:C], "viewdirs": d[:C]}, False, False, None, None, out_depth=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n_calls):
        j = (i * C) % (o.shape[0] - C)
        net({"rays_o": o[j:j + C], "rays_d": d[j:j + C], "viewdirs": d[j:j + C]}, False, False, None, None, out_depth=True)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
print(f"{n_calls} chunk calls: host issue {1e3 * t_host / n_calls:.3f} ms/call, device {e0.elapsed_time(e1) / n_calls:.3f} ms/call, wall {1e3 * t_all / n_calls:.3f} ms/call")

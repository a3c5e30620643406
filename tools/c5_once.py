"""C5 (Barnes-Hut N = 2^24 galaxy pair) build + walk through the device entry points: python tools/c5_once.py [reps] [N] [first step number]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from starflood_b200 import engine as e, ics
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
step0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
e.init(0); L = e.lib()
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(); sptr = stream.cuda_stream
pos, vel, mas = ics.galaxy_pair(N)
pm = np.empty((N, 4), np.float32); pm[:, :3] = pos.reshape(-1, 3); pm[:, 3] = mas
pm4 = torch.from_numpy(pm).to(dev).reshape(-1)
acc = torch.zeros(3 * N, dtype=torch.float32, device=dev); pot = torch.zeros(N, dtype=torch.float32, device=dev)
box = np.zeros(6, np.float32)
for r in range(reps):
    l0 = L.sfe_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(stream)
    rc = L.sfe_dev_root_box(pm4.data_ptr(), N, step0 + r, box.ctypes.data, sptr); ev[1].record(stream)
    rc |= L.sfe_dev_tree_build(box.ctypes.data, pm4.data_ptr(), N, e.TREE_DEPTH_MAX, N // 2, sptr); ev[2].record(stream)
    rc |= L.sfe_dev_bh(pm4.data_ptr(), N, 0, N, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, 0.5, e.WALK_EXACT, sptr); ev[3].record(stream)
    torch.cuda.synchronize()
    print('N', N, rc, 'box_ms %.3f build_ms %.3f walk_ms %.3f' % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])), 'launches', L.sfe_launch_count() - l0, e.last_error() if rc else '')

#!/usr/bin/env python3
"""Where does the end-to-end time of a 10-frame batch go? Host time of the call vs. wall time per step, float and mm depth."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvids_b200 import capi, scenes

CFG = scenes.CONFIG2
cam = CFG.cam
H, W = cam.height, cam.width
n = 60
frames = [scenes.stream_frame(CFG, f) for f in range(n)]
hd = torch.empty((n, H, W), dtype=torch.float32).pin_memory()
hm = torch.empty((n, H, W), dtype=torch.int16).pin_memory()
hc = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
for i, (d, c, p) in enumerate(frames):
    hd[i].copy_(torch.from_numpy(d)); hc[i].copy_(torch.from_numpy(c))
    hm[i].copy_(torch.from_numpy(np.clip(np.nan_to_num(d, nan=0.0) * 1000, 0, 65535).astype(np.uint16).view(np.int16)))
integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, CFG.truncation, CFG.weight, CFG.carve, CFG.carve_dist)
for name, depth, asyn in (("float sync", hd, False), ("float async", hd, True), ("mm async", hm, True)):
    m = capi.Chisel(CFG.chunk, CFG.resolution, True, initial_chunks=98304)
    t_call, t0 = 0.0, None
    prev = None
    for s in range(6):
        if s == 1:
            m.synchronize(); t0 = time.perf_counter(); t_call = 0.0
        ids = range(s * 10, s * 10 + 10)
        ds = [depth[i].numpy().view(np.uint16) if depth is hm else depth[i].numpy() for i in ids]
        a = time.perf_counter()
        m.integrate_batch(integ, ds, [frames[i][2] for i in ids], cam.as_array(), [hc[i].numpy() for i in ids], host_async=asyn)
        t_call += time.perf_counter() - a
        tk = m.last_batch_ticket()
        if prev is not None:
            m.wait_batch(prev)
        prev = tk
    m.wait_batch(prev)
    m.synchronize()
    wall = time.perf_counter() - t0
    print("%-12s wall %.0f us/step, host time inside integrate_batch %.0f us/step" % (name, 1e6 * wall / 5, 1e6 * t_call / 5))
    m.close()

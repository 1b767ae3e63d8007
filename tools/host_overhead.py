#!/usr/bin/env python3
"""Host-side cost of one chs_integrate_depth_color call (enqueue only) and GPU time per frame without any L2 flush."""
import ctypes as C
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvids_b200 import capi, scenes

cfg = scenes.CONFIG2
n = 40
frames = [scenes.stream_frame(cfg, f) for f in range(n)]
dev = torch.device("cuda", 0)
dd = torch.stack([torch.from_numpy(f[0]) for f in frames]).to(dev)
dc = torch.stack([torch.from_numpy(f[1]) for f in frames]).to(dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
m = capi.Chisel(cfg.chunk, cfg.resolution, True, device=0, stream=stream.cuda_stream, initial_chunks=16384)
integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, cfg.truncation, cfg.weight, cfg.carve, cfg.carve_dist)
camv = cfg.cam.as_array()
lib = capi.load_library()
cam = capi.make_camera(camv)
istruct = integ.as_struct()
poses = [np.ascontiguousarray(f[2].reshape(12)) for f in frames]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        m.integrate_depth_scan_color(integ, None, poses[i], camv, None, device_ptrs=(dd[i].data_ptr(), dc[i].data_ptr()), channels=3)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("python wrapper: enqueue %.1f us/frame, total %.1f us/frame" % (1e6 * (t1 - t0) / n, 1e6 * (t2 - t0) / n))
ptrs = [(dd[i].data_ptr(), dc[i].data_ptr(), poses[i].ctypes.data_as(C.c_void_p)) for i in range(n)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        d, c, p = ptrs[i]
        lib.chs_integrate_depth_color(m._h, C.byref(istruct), d, 1, p, C.byref(cam), c, 3, p, C.byref(cam))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("raw C ABI:      enqueue %.1f us/frame, total %.1f us/frame" % (1e6 * (t1 - t0) / n, 1e6 * (t2 - t0) / n))
st = m.frame_stats()
print(st)

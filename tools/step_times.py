#!/usr/bin/env python3
"""Per-step device times (CUDA events) with and without the L2 flush, to see patterns (ring stalls, graph uploads)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvids_b200 import capi, scenes

cfg = scenes.CONFIG2
n = 48
frames = [scenes.stream_frame(cfg, f) for f in range(n)]
dev = torch.device("cuda", 0)
dd = torch.stack([torch.from_numpy(f[0]) for f in frames]).to(dev)
dc = torch.stack([torch.from_numpy(f[1]) for f in frames]).to(dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, cfg.truncation, cfg.weight, cfg.carve, cfg.carve_dist)
camv = cfg.cam.as_array()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
for mode in ("noflush", "flush_w", "flush_wr", "flush_wr_sync"):
    m = capi.Chisel(cfg.chunk, cfg.resolution, True, device=0, stream=stream.cuda_stream, initial_chunks=16384)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        if mode != "noflush":
            flush.fill_(i & 0xFF)
            if mode != "flush_w":
                flush_rd.max()
        if mode == "flush_wr_sync":
            torch.cuda.synchronize()
        ev[i][0].record(stream)
        m.integrate_depth_scan_color(integ, None, frames[i][2], camv, None, device_ptrs=(dd[i].data_ptr(), dc[i].data_ptr()), channels=3)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ts = [a.elapsed_time(b) * 1000 for a, b in ev]
    print(mode, "wall/step %.0f us" % (1e6 * wall / n), "mean %.1f" % np.mean(ts[8:]), " ".join("%.0f" % t for t in ts))
    m.close()

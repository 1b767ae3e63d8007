#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native OpenChisel hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] -- the single-agent EuRoC-shape 752x480 synthetic depth+colour
stream of the analytic room, 2 cm voxels, 16^3 chunks, truncation 4 voxels, ConstantWeighter(1), carving on;
one STEP = --batch (default 10) consecutive frames handed over in ONE chs_integrate_batch call (the fused multi-frame
kernels; bit-identical to IntegrateDepthScanColor frame by frame; 10 = the period of the reference's UpdateMeshes gate,
i.e. what the drop-in facade queues without any observable difference). --batch 1: one frame per call. Steps
W .. W+K-1 of the 200-frame orbit are timed after W warm-up steps from an empty map.

Metric: TSDF voxel updates per second (GVox/s; a "voxel update" is one DistVoxel::Integrate), with
frames/s alongside. `value`: inputs resident in HBM, device time from CUDA events per step, L2 flushed between
steps (flush excluded). `e2e`: the same call with pinned HOST frames (H2D inside) plus the D2H read of every step's
per-frame counters (depth-2 pipeline), wall clock per step. `roofline`: algorithmic bytes of the brick kernel / its
event-timed duration, against MEASURED_PEAKS.json; `traffic` from the committed ncu capture. `single_frame_calls`,
`l2_warm`, `e2e_depth_mm`, `mesh`, `side_lines`: context, see DESIGN.md section 7. `cpu_baseline`: the reference CPU OpenChisel (oracle/_ref, else the C port) on a
bounded sample of the same frames on this box's host cores.

N > 1: one process per GPU; the chunk-ID hash space is partitioned (owner = chs_owner(id) % N). Every step's frame block is
ingested in N equal byte ranges, one per rank, and replicated with ONE NCCL all-gather; every rank then
integrates the chunks it owns. Total work is fixed => "scaling": "strong". Time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cvids_b200 import scenes  # noqa: E402

METRIC = "tsdf_voxel_updates_per_s"
UNIT = "GVox/s"
CFG = scenes.CONFIG2
WORKLOAD = "configs[1]: single-agent EuRoC-shape 752x480 depth+colour stream, analytic room, 2 cm voxels, 16^3 chunks, " \
           "trunc 4 voxels, IntegrateDepthScanColor; step = %d consecutive frame(s) in one chs_integrate_batch call"


def frames_for(lo: int, hi: int):
    out = []
    for f in range(lo, hi):
        depth, col, pose = scenes.stream_frame(CFG, f % CFG.n_frames)
        out.append((depth, col, pose))
    return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, indices, enabled=True):
        """indices: the GPUs of the job. ONE sampler (rank 0) queries all of them in one nvidia-smi call: a poller per rank
        contends with the kernel launches of eight processes for the driver."""
        super().__init__(daemon=True)
        self.indices, self.samples, self._halt, self.enabled = list(indices), [], threading.Event(), enabled

    def run(self):
        while self.enabled and not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for row in out.split("\n"):
                    if row.strip():
                        self.samples.append([x.strip() for x in row.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))


def algorithmic_bytes(st: dict, cam: scenes.Camera, channels: int, use_color: bool) -> int:
    """B_int of SURVEY.md 8(d) / BASELINE.md section 4 for one frame."""
    V = CFG.chunk ** 3
    px = cam.width * cam.height
    return (16 * (st["n_upd"] + st["n_carve"]) + 8 * st["n_col"] + V * (8 + 4 * int(use_color)) * st["n_new"]
            + 4 * px + channels * px * int(use_color))


# ---------------------------------------------------------------------------------------------------------------
# reference arm and CPU baseline

def cpu_arm(frames, warm, steps, budget_s):
    """Reference CPU OpenChisel on the host cores: oracle/_ref (as-is, 16 threads: Chisel.h:150) when it was built,
    else the single-threaded C port. Returns dict(value GVox/s, fps, kind, cores, steps_done, seconds)."""
    from oracle import pyoracle
    use_ref = pyoracle.ref_available()
    cls = pyoracle.RefChisel if use_ref else pyoracle.OracleChisel
    ref = cls(CFG.chunk, CFG.resolution, True)
    ref.setup_integrator(pyoracle.TRUNC_CONSTANT, CFG.truncation, CFG.weight, CFG.carve, CFG.carve_dist)
    counter = pyoracle.OracleChisel(CFG.chunk, CFG.resolution, True)       # untimed: counts N_upd for the same frames
    counter.setup_integrator(pyoracle.TRUNC_CONSTANT, CFG.truncation, CFG.weight, CFG.carve, CFG.carve_dist)
    cam = CFG.cam.as_array()
    t_total, upd, done = 0.0, 0, 0
    t_begin = time.perf_counter()
    for i, (depth, col, pose) in enumerate(frames[:warm + steps]):
        t0 = time.perf_counter()
        ref.integrate_color(depth, pose, cam, col, as_is=True)
        dt = time.perf_counter() - t0
        if use_ref:
            counter.integrate_color(depth, pose, cam, col)              # untimed
            n = counter.frame_counters()["n_upd"]
        else:
            n = ref.frame_counters()["n_upd"]
        if i >= warm:
            t_total += dt
            upd += n
            done += 1
            if time.perf_counter() - t_begin > budget_s:
                break
    cores = os.cpu_count() or 1
    return dict(value=upd / t_total / 1e9 if t_total else 0.0, fps=done / t_total if t_total else 0.0,
                kind="reference" if use_ref else "port", cores=min(16, cores) if use_ref else 1, host_cores=cores,
                steps_done=done, seconds=t_total, updates=upd)


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores, same workload and step (B consecutive
    frames, integrated one by one -- the reference has no batch entry). Bounded by --cpu-budget seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = max(1, min(args.batch, 16))
    warm, steps = args.warmup * B, args.steps * B
    n_unique = min(warm + steps, CFG.n_frames)
    frames = frames_for(0, n_unique)
    frames = [frames[i % n_unique] for i in range(warm + steps)]
    r = cpu_arm(frames, warm, steps, budget_s=args.cpu_budget)
    steps_done = r["steps_done"] / B
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": 1000.0 * r["seconds"] / max(steps_done, 1e-9), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "frames_per_step": B, "frames_per_s": r["fps"],
        "config": {"workload": WORKLOAD % B, "threads": "16 std::threads hard-coded by the reference (Chisel.h:150)" if r["kind"] == "reference" else "1 (C port)"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "host_cores": r["host_cores"], "kind": r["kind"],
                         "sample": "%d frames after %d warm-up frames of the same stream, whole frames, %.1f s" % (r["steps_done"], warm, r["seconds"])},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------------------------------------
# parity check inside the bench: the frames the bench times, against the CPU oracle

PARITY_KEYS = ("candidates", "n_upd", "n_carve", "n_col", "n_new", "updated_chunks")


def oracle_counters(frames):
    """Per-frame counters of the CPU oracle (the plain-C restatement, bit-identical to the compiled reference) and the oracle itself."""
    from oracle import pyoracle
    o = pyoracle.OracleChisel(CFG.chunk, CFG.resolution, True)
    o.setup_integrator(pyoracle.TRUNC_CONSTANT, CFG.truncation, CFG.weight, CFG.carve, CFG.carve_dist)
    cam = CFG.cam.as_array()
    out = []
    for depth, col, pose in frames:
        o.integrate_color(depth, pose, cam, col)
        out.append(o.frame_counters())
    return out, o


def state_equal(a, b):
    """Bit-exact comparison of two (ids, sdf, weight, rgbw) states; returns a list of what differs."""
    bad = []
    if a[0].shape != b[0].shape or not np.array_equal(a[0], b[0]):
        return ["chunk-ID set (%d vs %d chunks)" % (len(a[0]), len(b[0]))]
    for name, x, y in (("sdf", a[1], b[1]), ("weight", a[2], b[2]), ("colour", a[3], b[3])):
        xv = np.ascontiguousarray(x).view(np.uint32) if x.dtype == np.float32 else x
        yv = np.ascontiguousarray(y).view(np.uint32) if y.dtype == np.float32 else y
        if not np.array_equal(xv, yv):
            bad.append("%s (%d voxels)" % (name, int((xv != yv).sum())))
    return bad

# ---------------------------------------------------------------------------------------------------------------
# side lines (not the headline): BASELINE configs[4] shape (8 agents) and configs[3] shape (1 cm hall, meshing-heavy)

def side_lines(torch, capi, dev, stream, flush_l2, pool_chunks):
    out = {}
    # --- 8 agents x 640x480 depth only, 2 cm: the eight frames of a time step are one batch (arrival order) ---
    cfg = scenes.CONFIG5
    integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, cfg.truncation, cfg.weight, cfg.carve, cfg.carve_dist)
    camv = cfg.cam.as_array()
    n_t, warm_t = 8, 2
    frames = [[scenes.stream_frame(cfg, t, agent=a) for a in range(cfg.agents)] for t in range(n_t)]
    dd = [[torch.from_numpy(f[0]).to(dev) for f in grp] for grp in frames]
    m = capi.Chisel(cfg.chunk, cfg.resolution, False, stream=stream.cuda_stream, initial_chunks=pool_chunks)
    t_dev, upd = 0.0, 0
    for t in range(n_t):
        flush_l2(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        m.integrate_batch(integ, None, [f[2] for f in frames[t]], camv, device_ptrs=[(d.data_ptr(), None) for d in dd[t]])
        e1.record(stream)
        sts = m.batch_stats()
        if t >= warm_t:
            t_dev += e0.elapsed_time(e1) / 1000.0
            upd += sum(s["n_upd"] for s in sts)
    out["eight_agents_2cm"] = {"workload": "configs[4] shape: 8 agents x 640x480 depth, 2 cm; step = the 8 frames of one time step in one batch",
                               "value": upd / t_dev / 1e9, "unit": UNIT, "frames_per_s": 8 * (n_t - warm_t) / t_dev,
                               "ms_per_step": 1000.0 * t_dev / (n_t - warm_t), "steps": n_t - warm_t, "map_chunks": m.frame_stats()["total_chunks"]}
    m.close()
    # --- 1 cm hall, lawn-mower sweep, then ONE re-mesh of everything dirty (meshing-heavy) ---
    hall = scenes.hall(seed=3)
    cam = scenes.Camera(525.0, 525.0, 319.5, 239.5, 640, 480, near=0.05, far=5.0)
    res = 0.01
    integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, float(np.float32(4.0) * np.float32(res)), 1.0, True, 0.05)
    poses = [scenes.yaw_pose(0.35 * (i % 6), (-20.0 + 1.0 * (i // 6) + 0.15 * (i % 6), -20.0 + 0.9 * (i % 6), 0.0)) for i in range(24)]
    dd = [torch.from_numpy(scenes.render(hall, cam, p)[0]).to(dev) for p in poses]
    m = capi.Chisel(16, res, False, stream=stream.cuda_stream, initial_chunks=pool_chunks)
    m.set_profiling(True)
    t_dev, upd = 0.0, 0
    for i in range(0, 24, 8):
        flush_l2(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        m.integrate_batch(integ, None, poses[i:i + 8], cam.as_array(), device_ptrs=[(d.data_ptr(), None) for d in dd[i:i + 8]])
        e1.record(stream)
        sts = m.batch_stats()
        t_dev += e0.elapsed_time(e1) / 1000.0
        upd += sum(s["n_upd"] for s in sts)
        tm = m.timings()
        kern = {k: kern.get(k, 0.0) + tm[k + "_ms"] / 3.0 for k in ("prepare", "candidates", "integrate")} if i else \
            {k: tm[k + "_ms"] / 3.0 for k in ("prepare", "candidates", "integrate")}
        cand = sum(s["candidates"] for s in sts)
    dirty = m.get_meshes_to_update()
    m._lib.chs_update_meshes(m._h)
    for _ in range(2):
        m.set_dirty(dirty)
        flush_l2(1)
        assert m._lib.chs_update_meshes(m._h) == 0
    mt, mc = m.timings(), m.last_mesh_counts()
    b_mc = mc["n_chunks"] * 17 ** 3 * 8 + mc["n_vertices"] * 24 + mc["n_grids"] * 12
    out["hall_1cm"] = {"workload": "configs[3] shape: 50x50x5 m pillar hall, 1 cm voxels, 640x480 depth, 24 frames in batches of 8, then one re-mesh of the whole dirty set",
                       "integration": {"value": upd / t_dev / 1e9, "unit": UNIT, "frames_per_s": 24 / t_dev, "voxel_updates_per_frame": upd / 24,
                                       "ms_per_8_frame_batch": {"prepare": kern["prepare"], "candidates": kern["candidates"], "bricks": kern["integrate"]},
                                       "candidate_chunks_last_batch": cand},
                       "remesh": {"dirty_ids": int(len(dirty)), "remeshed_chunks": mc["n_chunks"], "triangles": mc["n_vertices"] // 3,
                                  "device_ms": mt["mesh_ms"], "count_ms": mt["mesh_count_ms"], "emit_ms": mt["mesh_emit_ms"],
                                  "algorithmic_bytes": b_mc, "achieved_gbs": b_mc / (mt["mesh_ms"] * 1e-3) / 1e9 if mt["mesh_ms"] > 0 else None}}
    m.close()
    return out


# ---------------------------------------------------------------------------------------------------------------
# CUDA arm

def run_cuda(args):
    import torch
    import torch.distributed as dist
    from cvids_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the one JSON line
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # a real (non-default) stream for everything: handle 0 would make the library create its own stream and the
    # CUDA events of the timed region would then bracket nothing
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm, steps, B = args.warmup, args.steps, max(1, min(args.batch, 16))
    cam = CFG.cam
    camv = cam.as_array()
    channels = 3
    H, W = cam.height, cam.width
    nfr = (warm + steps) * B                       # frames of the stream used by the batched legs
    n_unique = min(nfr, CFG.n_frames)              # the 200-frame orbit wraps around after that

    # rank 0 owns the stream; other ranks receive frames by NCCL broadcast
    frames = frames_for(0, n_unique) if rank == 0 else None
    poses = np.zeros((n_unique, 12), np.float32)
    if rank == 0:
        for i, fr in enumerate(frames):
            poses[i] = fr[2].reshape(12)
    if world > 1:
        pt = torch.from_numpy(poses).to(dev)
        dist.broadcast(pt, 0)
        poses = pt.cpu().numpy()

    # one contiguous byte buffer per frame [depth f32 | colour u8] so that a frame (or a batch of consecutive frames) is ONE
    # NCCL broadcast (sharding.py)
    fbytes = sharding.frame_nbytes(W, H, channels)
    dbytes = 4 * W * H
    d_frames = torch.empty((n_unique, fbytes), dtype=torch.uint8, device=dev)
    h_frames = None
    if rank == 0:
        h_frames = torch.empty((n_unique, fbytes), dtype=torch.uint8).pin_memory()
        for i, fr in enumerate(frames):
            h_frames[i].copy_(torch.from_numpy(sharding.pack_frame(fr[0], fr[1])))
        d_frames.copy_(h_frames)
    # N > 1, sharded ingest: the frames of a step enter the node through ALL ranks -- rank r ingests the r-th of N equal byte
    # ranges of the step's contiguous [depth | colour] x B block (over its own PCIe link in the e2e leg) and ONE all-gather
    # replicates the block. Every byte is still broadcast over NVLink from its ingest rank, all ingest ranks at once, and all
    # ranks finish together (a ring broadcast from a single root reaches the last of 8 ranks only after ~300 us). Set-up,
    # outside every timed region: every rank gets a copy of the synthetic stream so that it can play the ingest rank of its range.
    share = sharding.ingest_share(B * fbytes, world)                     # bytes per rank and step
    if world > 1:
        dist.broadcast(d_frames, 0)
        if rank != 0:
            h_frames = torch.empty((n_unique, fbytes), dtype=torch.uint8).pin_memory()
            h_frames.copy_(d_frames)
        torch.cuda.synchronize(dev)
    recv = torch.empty(max(B * fbytes, share * world) + fbytes, dtype=torch.uint8, device=dev)   # all-gather output = the step's block
    mine = torch.empty(max(share, fbytes), dtype=torch.uint8, device=dev)                        # this rank's range when it needs staging
    d_flat = d_frames.view(-1)
    h_flat = h_frames.view(-1) if h_frames is not None else None
    # L2 flush between timed steps: write a 256 MiB buffer, then read another one, so that L2 ends up full of CLEAN
    # foreign lines (a write-only flush leaves ~126 MB of dirty lines whose write-back would be charged to the step)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush_l2 else None
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev) if args.flush_l2 else None

    def flush_l2(k):
        if flush is not None:
            flush.fill_(k & 0xFF)
            flush_rd.max()
    integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, CFG.truncation, CFG.weight, CFG.carve, CFG.carve_dist)

    def new_map():
        return capi.Chisel(CFG.chunk, CFG.resolution, True, device=local, rank=rank, world=world, stream=stream.cuda_stream,
                           initial_chunks=args.pool_chunks)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def frame_ids(step, b=B):
        return [(step * b + j) % n_unique for j in range(b)]

    prepared_dev = {}

    def my_range(step, b):
        """This rank's byte range of the step's block: (offset into the flat stream, valid bytes), or None when the frames of the
        step are not contiguous in the stream (the orbit wraps inside the step)."""
        ids = frame_ids(step, b)
        lo, n = sharding.ingest_range(b * fbytes, rank, world)
        if ids[-1] - ids[0] != b - 1:
            return None, lo, n
        return ids[0] * fbytes + lo, lo, n

    def exchange(step, b):
        if b == 1:
            i = frame_ids(step, 1)[0]
            sharding.broadcast_frame(d_frames[i:i + 1] if rank == 0 else recv[:fbytes].view(1, fbytes), 0)
            return
        off, lo, n = my_range(step, b)
        if off is not None and n == share:
            block = d_flat[off:off + share]                   # a full range of the resident stream: no staging copy
        else:
            if n > 0:
                if off is not None:
                    mine[:n].copy_(d_flat[off:off + n])
                else:
                    whole = d_frames[frame_ids(step, b)].view(-1)
                    mine[:n].copy_(whole[lo:lo + n])
            block = mine[:share]
        sharding.all_gather_block(recv[:share * world], block)

    def prepared_for(m, step, b):
        key = (b, step)
        if key not in prepared_dev:
            ids = frame_ids(step, b)
            if world > 1:
                base = recv.data_ptr() if (b > 1 or rank != 0) else d_frames[ids[0]:ids[0] + 1].data_ptr()   # b == 1: root integrates in place
                ptrs = [(base + j * fbytes, base + j * fbytes + dbytes) for j in range(b)]
            else:
                base = d_frames.data_ptr()
                ptrs = [(base + i * fbytes, base + i * fbytes + dbytes) for i in ids]
            prepared_dev[key] = ptrs[0] if b == 1 else m.prepare_batch(integ, None, [poses[i] for i in ids], camv, device_ptrs=ptrs, channels=channels)
        return prepared_dev[key]

    mid_events = []

    def step_device(m, step, b=B, mid=None):
        """One step = b consecutive frames, inputs resident in HBM on their ingest ranks -> [NCCL all-gather] -> fused integration.
        The call's arguments (device pointers, poses) are marshalled once per (b, step) -- harness work, not part of the path."""
        ids = frame_ids(step, b)
        if world > 1:
            exchange(step, b)
            if mid is not None:
                mid.record(stream)
        pre = prepared_for(m, step, b)
        if b == 1:
            m.integrate_depth_scan_color(integ, None, poses[ids[0]], camv, None, device_ptrs=pre, channels=channels)
        else:
            m.integrate_prepared(pre)

    align = torch.zeros(1, device=dev)

    def timed_leg(b, n_warm, n_steps, do_flush):
        """CUDA events around every step on the stream the kernels run on; returns (device seconds, wall seconds, clocks)."""
        m = new_map()
        for i in range(n_warm + n_steps):
            prepared_for(m, i, b)
        for i in range(n_warm):
            step_device(m, i, b)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        sampler = ClockSampler(range(world), enabled=(rank == 0))
        sampler.start()
        barrier()
        t0 = time.perf_counter()
        if do_flush:
            for k in range(n_steps):
                flush_l2(k)
                if world > 1:
                    # align the ranks IN STREAM ORDER before the step starts: without it the skew between the ranks' flushes is
                    # waited out by the step's broadcast, inside the timed bracket
                    dist.all_reduce(align)
                ev[k][0].record(stream)
                mid = torch.cuda.Event(enable_timing=True) if world > 1 else None
                step_device(m, n_warm + k, b, mid)
                ev[k][1].record(stream)
                if mid is not None:
                    mid_events.append((ev[k][0], mid, ev[k][1]))
        else:
            ev[0][0].record(stream)
            for k in range(n_steps):
                step_device(m, n_warm + k, b)
            ev[0][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        clk = sampler.stop()
        t = sum(a.elapsed_time(c) for a, c in (ev if do_flush else ev[:1])) / 1000.0
        m.close()
        return t, wall, clk

    # ---------------- leg A: device-resident inputs, CUDA events per step, L2 flushed between steps -------------
    t_dev, wall_a, clocks = timed_leg(B, warm, steps, True)
    bracket = None
    if mid_events:
        bracket = {"broadcast_us": 1000.0 * float(np.mean([a.elapsed_time(b_) for a, b_, _ in mid_events])),
                   "integrate_us": 1000.0 * float(np.mean([b_.elapsed_time(c_) for _, b_, c_ in mid_events]))}
        mid_events.clear()
        allb = [None] * world
        dist.all_gather_object(allb, bracket)
        bracket = {"per_rank_exchange_us": [round(x["broadcast_us"], 1) for x in allb], "per_rank_integrate_us": [round(x["integrate_us"], 1) for x in allb]}
    # ---------------- leg A': same, no flush (the map working set stays in L2 as it does in a live stream) -------
    t_warm, _, _ = timed_leg(B, warm, steps, False) if not args.quick else (t_dev, 0, 0)
    # ---------------- leg S: one frame per call (chs_integrate_depth_color, the reference's call granularity) ----
    s_steps, s_warm = min(60, n_unique - 10), 10
    t_single, _, _ = timed_leg(1, s_warm, s_steps, True) if (B > 1 and not args.quick) else (t_dev, 0, 0)

    # ---------------- leg C: per-frame counters and kernel times (profiling events inside the library) ----------
    m = new_map()
    m.set_profiling(True)
    upd_local = upd_single = 0
    bytes_alg = 0
    t_integrate = t_prepare = t_cand = t_new = t_span = 0.0
    per_step = []
    parity_steps = min(args.parity_steps, warm + steps, max(1, n_unique // B))
    got_counters = []                              # per-frame counters of the first parity_steps steps (this rank's chunks)
    for i in range(warm + steps):
        if i >= warm:
            flush_l2(i)
        step_device(m, i)
        sts = m.batch_stats() if B > 1 else [m.frame_stats()]
        if i < parity_steps:
            got_counters += [[int(st[k]) for k in PARITY_KEYS] for st in sts]
        if i * B < s_warm + s_steps:
            upd_single += sum(st["n_upd"] for j, st in enumerate(sts) if s_warm <= i * B + j < s_warm + s_steps)
        if i >= warm:
            tm = m.timings()
            upd_local += sum(st["n_upd"] for st in sts)
            bytes_alg += sum(algorithmic_bytes(st, cam, channels, True) for st in sts)
            t_integrate += tm["integrate_ms"] / 1000.0
            t_prepare += tm["prepare_ms"] / 1000.0
            t_cand += tm["candidates_ms"] / 1000.0
            t_new += tm["new_chunks_ms"] / 1000.0
            t_span += tm.get("bricks_span_ms", 0.0) / 1000.0
            per_step.append((sum(st["n_upd"] for st in sts), sts[-1]["brick_units"], sum(st["candidates"] for st in sts), tm["integrate_ms"],
                             sum(st["updated_chunks"] for st in sts), sum(st["n_new"] for st in sts), sts[-1]["new_candidates"]))
    total_chunks = m.frame_stats()["total_chunks"]
    # ---- parity check: the first parity_steps steps of exactly this path against the CPU oracle (counters summed over ranks;
    # at N = 1 also the whole voxel state and the dirty set, bit for bit, from a second map fed the same steps)
    parity = None
    if parity_steps > 0:
        gc = torch.tensor(got_counters, dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(gc, op=dist.ReduceOp.SUM)
        if rank == 0:
            t0 = time.perf_counter()
            want, orc = oracle_counters([frames[i] for s_ in range(parity_steps) for i in frame_ids(s_)])
            gcl = gc.cpu().tolist()
            bad = []
            for j, (g, w) in enumerate(zip(gcl, want)):
                for k, name in enumerate(PARITY_KEYS):
                    if g[k] != w[name]:
                        bad.append("frame %d %s: cuda %d oracle %d" % (j, name, g[k], w[name]))
            parity = {"frames": len(want), "counters": list(PARITY_KEYS), "counters_equal": not bad, "oracle": "oracle/chisel_oracle.c (C restatement, pinned to the compiled reference)"}
            if world == 1:
                pm = new_map()
                for s_ in range(parity_steps):
                    step_device(pm, s_)
                diff = state_equal(pm.state(), orc.state())
                if not np.array_equal(pm.dirty_ids(), orc.dirty_ids()):
                    diff.append("dirty set")
                parity["state_bit_exact"] = not diff
                parity["chunks"] = int(len(orc.state()[0]))
                bad += diff
                pm.close()
            parity["seconds"] = round(time.perf_counter() - t0, 1)
            if bad:
                # a fast kernel whose results differ from the reference's is not done: no bench line
                print(json.dumps({"parity_check": parity, "mismatch": bad[:20]}), flush=True)
                raise SystemExit("bench.py: PARITY MISMATCH against the oracle: " + "; ".join(bad[:5]))
    # meshing: re-mesh of everything the run left dirty (Chisel::UpdateMeshes without its every-10th gate): once cold (first
    # launch, cold L2), then the same dirty set again twice (steady state; the dirty set is restored with chs_set_dirty)
    dirty = m.get_meshes_to_update()
    n_dirty = len(dirty)
    flush_l2(0)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    m.recompute_meshes()
    t_mesh_wall = time.perf_counter() - t0
    mt_cold = m.timings()
    mt = mt_cold
    for _ in range(2):
        m.set_dirty(dirty)
        flush_l2(1)
        _check_ok = m._lib.chs_update_meshes(m._h)
        assert _check_ok == 0
        mt = m.timings()
    mc = m.last_mesh_counts()
    V_halo = (CFG.chunk + 1) ** 3
    b_mc = mc["n_chunks"] * V_halo * 12 + mc["n_vertices"] * 36 + mc["n_grids"] * 12        # B_mc of SURVEY 8(d), colour map
    mesh_info = {"dirty_ids": n_dirty, "remeshed_chunks": mc["n_chunks"], "triangles": mc["n_vertices"] // 3, "grids": mc["n_grids"],
                 "device_ms": mt["mesh_ms"], "count_ms": mt["mesh_count_ms"], "emit_ms": mt["mesh_emit_ms"],
                 "first_call_device_ms": mt_cold["mesh_ms"],
                 "wall_ms_first_call_incl_download_and_host_merge": 1000.0 * t_mesh_wall,
                 "algorithmic_bytes": b_mc, "achieved_gbs": b_mc / (mt["mesh_ms"] * 1e-3) / 1e9 if mt["mesh_ms"] > 0 else None}
    m.close()

    # ---------------- leg B: end to end through the C ABI with pinned HOST frames + D2H counters ----------------
    m = new_map()

    # views of the pinned frames per step, built once (numpy view objects are harness overhead, not part of the path)
    host_views = {}
    prepared_host = {}

    def views_of(step):
        if step not in host_views:
            ids = frame_ids(step)
            pairs = [sharding.unpack_frame(h_frames[i].numpy(), W, H, channels) for i in ids]
            host_views[step] = ([p[0] for p in pairs], [p[1] for p in pairs], [poses[i] for i in ids])
        return host_views[step]

    if world == 1:
        for st_ in range(warm + steps):
            views_of(st_)

    def step_host(step, read=True):
        ids = frame_ids(step)
        if world > 1:
            # every rank copies ITS byte range of the step from pinned host memory over its own PCIe link, then one all-gather
            if B == 1:
                if rank == 0:
                    mine[:fbytes].copy_(h_frames[ids[0]], non_blocking=True)
                sharding.broadcast_frame((mine if rank == 0 else recv)[:fbytes].view(1, fbytes), 0)
                p0 = (mine if rank == 0 else recv).data_ptr()
                m.integrate_depth_scan_color(integ, None, poses[ids[0]], camv, None, device_ptrs=(p0, p0 + dbytes), channels=channels)
            else:
                off, lo, n = my_range(step, B)
                if n > 0:
                    if off is not None:
                        mine[:n].copy_(h_flat[off:off + n], non_blocking=True)
                    else:
                        mine[:n].copy_(h_frames[ids].view(-1)[lo:lo + n], non_blocking=True)
                sharding.all_gather_block(recv[:share * world], mine[:share])
                m.integrate_prepared(prepared_for(m, step, B))
        else:
            ds, cs, ps = views_of(step)
            if B == 1:
                m.integrate_depth_scan_color(integ, ds[0], ps[0], camv, cs[0])
            else:
                # the pinned frames are never modified: CHS_MEM_HOST_ASYNC lets the call return right after enqueueing
                if step not in prepared_host:
                    prepared_host[step] = m.prepare_batch(integ, ds, ps, camv, cs, host_async=True)
                m.integrate_prepared(prepared_host[step])
        if not read:
            return 0
        return sum(st["n_upd"] for st in m.batch_stats()) if B > 1 else m.frame_stats()["n_upd"]

    # Pipelined, as a streaming caller uses it: issue batch k, then read the counters of batch k - 1 (chs_wait_batch waits for
    # that one batch only), so that the H2D copies of a batch overlap the kernels of the one before. Every step's result is
    # read back inside the timed region.
    pipelined = B > 1
    if world == 1 and B > 1:
        for st_ in range(warm + steps):
            ds_, cs_, ps_ = views_of(st_)
            prepared_host[st_] = m.prepare_batch(integ, ds_, ps_, camv, cs_, host_async=True)
    for i in range(warm):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    upd_e2e = 0
    if pipelined:
        prev = None
        for k in range(steps):
            step_host(warm + k, read=False)
            tk = m.last_batch_ticket()
            if prev is not None:
                upd_e2e += sum(st["n_upd"] for st in m.wait_batch(prev))
            prev = tk
        upd_e2e += sum(st["n_upd"] for st in m.wait_batch(prev))
    else:
        for k in range(steps):
            upd_e2e += step_host(warm + k)
    barrier()
    t_e2e = time.perf_counter() - t0
    m.close()

    # ---------------- leg B': the same pipeline fed with 16UC1 millimetre depth (the sensor's own encoding; converted on the
    # device like chisel_ros converts it on the host). Different input VALUES (quantised to 1 mm), hence reported beside, not as, e2e.
    e2e_mm = None
    if world == 1 and B > 1 and not args.quick:
        h_mm = torch.empty((n_unique, H, W), dtype=torch.int16).pin_memory()
        for i, fr in enumerate(frames):
            q = np.clip(np.nan_to_num(fr[0], nan=0.0) * 1000.0, 0, 65535).astype(np.uint16)
            h_mm[i].copy_(torch.from_numpy(q.view(np.int16)))
        m = new_map()

        mm_views = {st_: [h_mm[i].numpy().view(np.uint16) for i in frame_ids(st_)] for st_ in range(warm + steps)}

        prepared_mm = {}

        def step_mm(step):
            if step not in prepared_mm:
                _, cs, ps = views_of(step)
                prepared_mm[step] = m.prepare_batch(integ, mm_views[step], ps, camv, cs, host_async=True)
            m.integrate_prepared(prepared_mm[step])

        for st_ in range(warm + steps):
            _, cs_, ps_ = views_of(st_)
            prepared_mm[st_] = m.prepare_batch(integ, mm_views[st_], ps_, camv, cs_, host_async=True)
        for i in range(warm):
            step_mm(i)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        upd_mm, prev = 0, None
        for k in range(steps):
            step_mm(warm + k)
            tk = m.last_batch_ticket()
            if prev is not None:
                upd_mm += sum(st["n_upd"] for st in m.wait_batch(prev))
            prev = tk
        upd_mm += sum(st["n_upd"] for st in m.wait_batch(prev))
        torch.cuda.synchronize(dev)
        t_mm = time.perf_counter() - t0
        m.close()
        e2e_mm = {"value": upd_mm / t_mm / 1e9, "unit": UNIT, "frames_per_s": steps * B / t_mm, "ms_per_step": 1000.0 * t_mm / steps,
                  "h2d_bytes_per_step": (2 * W * H + channels * W * H) * B,
                  "note": "depth handed over as uint16 millimetres (chs_frame.depth_mm), same pipeline as e2e"}

    # ---------------- reduce over ranks -------------------------------------------------------------------------
    vals = torch.tensor([t_dev, t_e2e, wall_a, t_integrate, t_warm, t_single], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(upd_local), float(upd_e2e), float(bytes_alg), float(total_chunks), float(upd_single)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t_dev, t_e2e, wall_a, t_int_max, t_warm, t_single = vals.tolist()
    upd_total, upd_e2e_total, bytes_total, chunks_total, upd_single_total = sums.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        # roofline of the integrate kernels on THIS rank (rank 0): algorithmic bytes they processed / their event time
        t_kernels = t_integrate + t_new
        achieved = bytes_alg / t_kernels / 1e9 if t_kernels > 0 else 0.0
        h2d = (4 * W * H + channels * W * H) * B
        kname = "batch_bricks_kernel<16,1,0>" if B > 1 else "integrate_bricks_kernel<16,color> + integrate_new_chunks_kernel<16,color>"
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(kname)
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"] if (tr and B == 10 and world == 1) else None
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": upd_total / t_dev / 1e9, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": 1000.0 * t_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "frames_per_step": B, "frames_per_s": steps * B / t_dev,
            "l2_warm": {"value": upd_total / t_warm / 1e9, "unit": UNIT, "frames_per_s": steps * B / t_warm, "ms_per_step": 1000.0 * t_warm / steps,
                        "note": "same steps back to back without the L2 flush; includes host launch gaps"},
            "single_frame_calls": {"value": upd_single_total / t_single / 1e9 if t_single > 0 else None, "unit": UNIT,
                                   "frames_per_s": s_steps / t_single if t_single > 0 else None, "ms_per_frame": 1000.0 * t_single / max(s_steps, 1),
                                   "note": "one frame per call (chs_integrate_depth_color, the reference's call granularity), frames %d..%d, L2 flushed" % (s_warm, s_warm + s_steps - 1)},
            "config": {"workload": WORKLOAD % B, "parallelism": "chunk-hash shard x%d; every step's frame block ingested in %d equal byte ranges (one per rank), replicated by one NCCL all-gather" % (world, world) if world > 1 else "1 GPU",
                       "l2": "256 MiB write + 256 MiB read between steps, excluded from the step time" if flush is not None else
                             "no flush: frame stream (%d MB) > L2, map working set stays in L2" % ((h2d * nfr) >> 20),
                       "voxel_updates_per_step": upd_total / steps, "map_chunks": chunks_total, "rank0_bracket": bracket,
                       "rank0_per_step": {"candidate_chunks": float(np.mean([p[2] for p in per_step])),
                                          "brick_units": float(np.mean([p[1] for p in per_step])),
                                          "new_chunk_candidates": float(np.mean([p[6] for p in per_step])),
                                          "updated_chunks": float(np.mean([p[4] for p in per_step])),
                                          "new_chunks": float(np.mean([p[5] for p in per_step]))}},
            "e2e": {"value": upd_e2e_total / t_e2e / 1e9, "unit": UNIT, "frames_per_s": steps * B / t_e2e, "ms_per_step": 1000.0 * t_e2e / steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 88 * B * world,
                    "timing": "wall clock; per step chs_integrate_batch(pinned host frames, CHS_MEM_HOST_ASYNC), then chs_wait_batch of the PREVIOUS step's counters "
                              "(depth-2 pipeline: copies of step k overlap kernels of step k-1)" if B > 1 else
                              "wall clock around chs_integrate_depth_color(host) + chs_get_frame_stats"},
            "e2e_depth_mm": e2e_mm,
            "gpu_launches": (4 if B > 1 else 5) * steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload (profiles/r01_traffic.json)" if traffic else None,
                         "algorithmic_bytes_per_launch": bytes_alg / steps, "kernel_ms_per_launch": 1000.0 * t_kernels / steps,
                         "bricks_ms_per_launch": 1000.0 * t_integrate / steps, "bricks_span_ms_per_launch": 1000.0 * t_span / steps, "new_chunks_ms_per_launch": 1000.0 * t_new / steps,
                         "prepare_ms_per_launch": 1000.0 * t_prepare / steps, "candidates_ms_per_launch": 1000.0 * t_cand / steps,
                         "note": "algorithmic bytes = B_int of SURVEY 8(d) summed over the step's frames; with %d frames fused the voxel state "
                                 "moves through HBM once per step, so DRAM traffic is BELOW the algorithmic bytes" % B},
            "wall_s_timed_region": wall_a,
            "parity_check": parity,
            "mesh": mesh_info,
        }
        if world == 1 and not args.quick and not args.no_side_lines:
            try:
                line["side_lines"] = side_lines(torch, capi, dev, stream, flush_l2, args.pool_chunks)
            except Exception as ex:                              # side lines must never cost the headline
                line["side_lines"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu:
            n_cpu = min(steps * B, args.cpu_frames)
            r = cpu_arm(frames, 0, n_cpu, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "host_cores": r["host_cores"], "kind": r["kind"],
                                    "frames_per_s": r["fps"],
                                    "sample": "first %d frames of the same stream from an empty map, whole frames, %.1f s" % (r["steps_done"], r["seconds"])}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=10, help="frames per step (chs_integrate_batch); 1 = one frame per call")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-flush-l2", dest="flush_l2", action="store_false")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-side-lines", action="store_true", help="skip the 8-agent and 1 cm hall side lines")
    ap.add_argument("--quick", action="store_true", help="A/B runs: only the flushed device leg and the profiling leg are meaningful")
    ap.add_argument("--cpu-frames", type=int, default=12)
    ap.add_argument("--pool-chunks", type=int, default=98304,
                    help="pre-sized chunk pool (chunks) so that no slab / hash growth lands inside the timed region")
    ap.add_argument("--cpu-budget", type=float, default=150.0)
    ap.add_argument("--parity-steps", type=int, default=2,
                    help="steps (from the empty map) whose per-frame counters, voxel state and dirty set are compared with the CPU oracle; 0 = off")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

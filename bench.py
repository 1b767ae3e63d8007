#!/usr/bin/env python3
"""bench.py -- benchmark of the B200-native OpenChisel hot path (TSDF integration + incremental marching cubes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload, at every N (config.workload): BASELINE.json configs[4] -- the scaling-sweep config: 8 agents, 640x480 depth
frames of the analytic room, 2 cm voxels, 16^3 chunks, truncation 4 voxels, carving on, all fused into ONE map. One STEP =
the frames of --time-steps-per-step (default 2) consecutive time steps, one frame per agent each (16 frames, arrival order =
time step, then agent) in one chs_integrate_batch[_distributed] call: the fused kernels take up to 16 frames.
N > 1: one process per GPU; the chunk-ID hash space is partitioned (owner = chs_owner(id) % N); rank r ingests the frames of
agents [r*8/N, (r+1)*8/N) and PUSHES them into the exchange arenas of all ranks over NVLink (peer memory mapped through CUDA
IPC inside the library, flag words instead of a collective; NCCL all-gather as fallback), every rank integrates the chunks it
owns. Same total work at every N => "scaling": "strong".

Metric: TSDF voxel updates per second (GVox/s; a "voxel update" is one DistVoxel::Integrate), frames/s alongside.
  value     inputs resident in HBM on their ingest ranks; W warm-up steps from an empty map, then EXACTLY K steps enqueued back to
            back, bracketed by barrier + synchronize and timed with CUDA events on the map's stream; max over ranks. No L2 flush:
            the frame stream of the timed region is larger than L2 (config.l2 says so), the map working set stays cached as it
            does in a live stream. Median of --passes passes (each from a fresh map, after one untimed rehearsal pass); clocks
            are sampled across all of them. The chunk pool is pre-sized (--pool-chunks) so that no growth lands in a timed pass.
  config.rank0_device_timeline_us  medians over the steps of a pass in which the kernels stamp %globaltimer themselves
            (chs_set_profiling(map, 2) / chs_get_device_timeline): kernel spans, gaps and overlap of the product path.
  e2e       the same call fed from pinned HOST frames (H2D inside the call, arguments marshalled inside the timed region) plus the
            D2H read of every step's per-frame counters (chs_wait_batch of the previous step: depth-2 pipeline), wall clock.
  e2e_depth_mm (N = 1)  the e2e leg with 16UC1 millimetre depth (the sensor's encoding; converted on the device): half the PCIe bytes.
  roofline  the dominant kernel (the fused brick kernel): B_int of SURVEY.md 8(d) summed over the step's frames / the kernel's
            duration from CUDA events recorded by the library around it, on its stream, averaged over the timed steps of a
            separate profiling pass; peak = MEASURED_PEAKS.json. `traffic` from the committed ncu capture of the same command.
  parity_check  the first --parity-steps steps of exactly this path against the CPU oracle: per-frame counters (summed over ranks),
            at N = 1 also the whole voxel state and the dirty set bit for bit. A mismatch aborts the run (no bench line).
  mesh      re-mesh of everything the run left dirty (N > 1: chs_update_meshes_distributed, incl. ghost exchange and gather).
  single_agent_color (N = 1)  the second line: BASELINE configs[1] (752x480 depth+colour, 2 cm, 10 frames per step) measured the
            same way, with its own roofline, e2e and L2-flushed per-step figure.  side_lines.hall_1cm: configs[3] shape.
  cpu_baseline / --impl reference  the reference CPU OpenChisel (oracle/_ref: the unmodified sources; else the C port) on a bounded
            sample of the same frames on this box's host cores (the depth path of the reference runs on ONE thread, Chisel.h:71-72).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cvids_b200 import scenes  # noqa: E402

METRIC = "tsdf_voxel_updates_per_s"
UNIT = "GVox/s"
CFG = scenes.CONFIG5                     # configs[4]: 8 agents x 60 frames, 640x480 depth, 2 cm
CFG2 = scenes.CONFIG2                    # configs[1]: single agent, 752x480 depth + colour, 2 cm
AGENTS = CFG.agents
WORKLOAD = ("configs[4]: scaling sweep, 8 agents x 640x480 depth, analytic room, 2 cm voxels, 16^3 chunks, trunc 4 voxels, carving on, "
            "IntegrateDepthScan into one shared map; step = the 8 frames of each of %d consecutive time step(s), arrival order, in one fused call")
WORKLOAD2 = ("configs[1]: single-agent EuRoC-shape 752x480 depth+colour stream, analytic room, 2 cm voxels, 16^3 chunks, trunc 4 voxels, "
             "IntegrateDepthScanColor; step = %d consecutive frames in one chs_integrate_batch call")
PARITY_KEYS = ("candidates", "n_upd", "n_carve", "n_col", "n_new", "updated_chunks")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clocks and throttle reasons DURING the timed passes (B200_PROFILING.md recipe). The samples come from NVML inside this
    process (the library nvidia-smi reads; no process start-up per sample, so the short timed passes get several samples);
    CHS_BENCH_SAMPLER=smi polls an nvidia-smi process instead, =off disables sampling. Measured: no effect on the timed passes."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, indices, enabled=True):
        super().__init__(daemon=True)
        self.indices, self.samples, self._halt = list(indices), [], threading.Event()
        self.mode = os.environ.get("CHS_BENCH_SAMPLER", "nvml")
        self.enabled = enabled and self.mode != "off"
        self.handles = []
        if self.enabled and self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
                self.nv = pynvml
                self.handles = [pynvml.nvmlDeviceGetHandleByIndex(phys[i] if phys else i) for i in self.indices]
            except Exception:
                self.mode = "smi"

    def _nvml_once(self):
        nv = self.nv
        for h in self.handles:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            act = lambda bit: "Active" if (r & bit) else "Not Active"
            # hw_slowdown 0x8, hw_thermal_slowdown 0x40, sw_thermal_slowdown 0x20, sw_power_cap 0x4
            self.samples.append([str(sm), str(mx), str(pw), act(0x8), act(0x40), act(0x20), act(0x4)])

    def run(self):
        while self.enabled and not self._halt.is_set():
            try:
                if self.mode == "nvml":
                    self._nvml_once()
                    self._halt.wait(0.02)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for row in out.split("\n"):
                    if row.strip():
                        self.samples.append([x.strip() for x in row.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

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
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm), source=self.mode)


def algorithmic_bytes(st: dict, cam, channels: int, use_color: bool, chunk: int) -> int:
    """B_int of SURVEY.md 8(d) / BASELINE.md section 4 for one frame."""
    V = chunk ** 3
    px = cam.width * cam.height
    return (16 * (st["n_upd"] + st["n_carve"]) + 8 * st["n_col"] + V * (8 + 4 * int(use_color)) * st["n_new"]
            + 4 * px + channels * px * int(use_color))


TS = 2                                   # time steps per step (--time-steps-per-step): a step is TS x 8 frames


def step_items(cfg, t):
    """(time step, agent) of the frames of step t in arrival order: time step by time step, agent by agent."""
    return [(t * TS + k, a) for k in range(TS) for a in range(cfg.agents)]


def step_frames(cfg, t, which=None):
    """Frames of step t of the multi-agent stream: [(depth, None, pose), ...]; `which`: indices into the step (default: all)."""
    items = step_items(cfg, t)
    return [scenes.stream_frame(cfg, items[i][0] % cfg.n_frames, agent=items[i][1]) for i in (range(len(items)) if which is None else which)]


def step_poses(cfg, t):
    return [scenes.orbit_pose(ts % cfg.n_frames, cfg.n_frames, a * (2.0 * math.pi / cfg.agents)) for ts, a in step_items(cfg, t)]


# ---------------------------------------------------------------------------------------------------------------
# CPU side: oracle counters for the parity check, reference arm, cpu_baseline

def make_cpu(cfg, use_ref: bool):
    from oracle import pyoracle
    cls = pyoracle.RefChisel if (use_ref and pyoracle.ref_available()) else pyoracle.OracleChisel
    o = cls(cfg.chunk, cfg.resolution, cfg.color)
    o.setup_integrator(pyoracle.TRUNC_CONSTANT, cfg.truncation, cfg.weight, cfg.carve, cfg.carve_dist)
    return o


def cpu_integrate(o, cfg, frame, as_is=False):
    depth, col, pose = frame
    cam = cfg.cam.as_array()
    if cfg.color:
        if as_is:
            o.integrate_color(depth, pose, cam, col, as_is=True)
        else:
            o.integrate_color(depth, pose, cam, col)
    else:
        o.integrate_depth(depth, pose, cam)


def state_equal(a, b):
    """Bit-exact comparison of two (ids, sdf, weight, rgbw) states; returns a list of what differs."""
    if a[0].shape != b[0].shape or not np.array_equal(a[0], b[0]):
        return ["chunk-ID set (%d vs %d chunks)" % (len(a[0]), len(b[0]))]
    bad = []
    for name, x, y in (("sdf", a[1], b[1]), ("weight", a[2], b[2]), ("colour", a[3], b[3])):
        xv = np.ascontiguousarray(x).view(np.uint32) if x.dtype == np.float32 else x
        yv = np.ascontiguousarray(y).view(np.uint32) if y.dtype == np.float32 else y
        if not np.array_equal(xv, yv):
            bad.append("%s (%d voxels)" % (name, int((xv != yv).sum())))
    return bad


class CpuArm:
    """Reference CPU OpenChisel on the host cores, frame by frame from an empty map: oracle/_ref (the unmodified reference, as-is
    threading) when built, else the single-threaded C port. N_upd of a frame comes from an untimed pass of the C port."""

    def __init__(self, cfg):
        from oracle import pyoracle
        self.cfg = cfg
        self.use_ref = pyoracle.ref_available()
        self.ref = make_cpu(cfg, True)
        self.counter = make_cpu(cfg, False) if self.use_ref else None
        cores = os.cpu_count() or 1
        self.host_cores = cores
        self.cores = (min(16, cores) if cfg.color else 1) if self.use_ref else 1      # Chisel.h:150 (colour: 16 threads) / :71-72 (depth: one)
        self.kind = "reference" if self.use_ref else "port"

    def frame(self, fr):
        """-> (seconds of the timed implementation, voxel updates of the frame)"""
        t0 = time.perf_counter()
        cpu_integrate(self.ref, self.cfg, fr, as_is=True)
        dt = time.perf_counter() - t0
        if self.counter is not None:
            cpu_integrate(self.counter, self.cfg, fr)
            return dt, self.counter.frame_counters()["n_upd"]
        return dt, self.ref.frame_counters()["n_upd"]


def cpu_arm(cfg, frames, warm, budget_s):
    """`frames` in order from an empty map, the first `warm` untimed, until the wall-clock budget is used up."""
    arm = CpuArm(cfg)
    t_total, upd, done = 0.0, 0, 0
    t_begin = time.perf_counter()
    for i, fr in enumerate(frames):
        dt, n = arm.frame(fr)
        if i >= warm:
            t_total += dt
            upd += n
            done += 1
            if time.perf_counter() - t_begin > budget_s:
                break
    return dict(value=upd / t_total / 1e9 if t_total else 0.0, fps=done / t_total if t_total else 0.0, kind=arm.kind,
                cores=arm.cores, host_cores=arm.host_cores, frames_done=done, seconds=t_total, updates=upd)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same workload, metric and
    unit. EXACTLY --steps timed steps after --warmup untimed ones; a step of this arm is a BOUNDED SAMPLE of the workload: n
    consecutive frames of the same arrival-ordered stream (from the empty map), n chosen from the cost of the first frame so that
    the whole run fits --cpu-budget seconds of wall clock (the GPU arm's step is the 16 frames of two time steps)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    K, Wm, full = max(1, args.steps), max(0, args.warmup), AGENTS * TS
    arm = CpuArm(CFG)

    def stream():
        t = 0
        while True:
            for fr in step_frames(CFG, t):
                yield fr
            t += 1
    src = stream()
    # calibration: the first frame of the stream (it is also the first frame of the run, timed or not as the step layout says)
    w0 = time.perf_counter()
    first = arm.frame(next(src))
    wall0 = time.perf_counter() - w0
    n = int(min(full, max(1, args.cpu_budget / (1.7 * wall0 * (K + Wm)))))           # later frames see a fuller map: factor 1.7
    t_total, upd, idx = 0.0, 0, 0
    for step in range(Wm + K):
        for _ in range(n):
            dt, nu = first if idx == 0 else arm.frame(next(src))
            idx += 1
            if step >= Wm:
                t_total += dt
                upd += nu
    value = upd / t_total / 1e9 if t_total else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": Wm, "ms_per_step": 1000.0 * t_total / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "frames_per_step": n, "frames_per_s": K * n / t_total if t_total else 0.0,
        "config": {"workload": WORKLOAD % TS, "reference_step": "a bounded sample: %d consecutive frames of the same stream per step (the GPU arm: %d)" % (n, full)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "host_cores": arm.host_cores, "kind": arm.kind,
                         "threads": "1: the reference's depth-only path is serial (Chisel.h:71-72)" if arm.kind == "reference" else "1 (C port)",
                         "sample": "%d timed steps of %d consecutive frames after %d warm-up steps, the same stream from an empty map, whole frames, %.1f s"
                                   % (K, n, Wm, t_total)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# CUDA arm

class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from cvids_b200 import capi
        self.torch, self.dist, self.capi, self.args = torch, dist, capi, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the one JSON line
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        # a real (non-default) stream for everything: the CUDA events of the timed region must be on the stream the kernels run on
        self.stream = torch.cuda.Stream(self.dev)
        torch.cuda.set_stream(self.stream)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            if (AGENTS * TS) % self.world:
                raise SystemExit("bench.py: the 8-agent workload needs --gpus in {1, 2, 4, 8}")
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def flush_l2(self, k=0):
        """write 256 MiB, then read 256 MiB: L2 ends up full of CLEAN foreign lines"""
        self.flush.fill_(k & 0xFF)
        self.flush_rd.max()

    def new_map(self, cfg, sharded=True):
        m = self.capi.Chisel(cfg.chunk, cfg.resolution, cfg.color, device=self.local, rank=self.rank if sharded else 0,
                             world=self.world if sharded else 1, stream=self.stream.cuda_stream, initial_chunks=self.args.pool_chunks)
        if sharded and self.world > 1:
            m.comm_init_torch()
        return m

    def integ(self, cfg):
        return self.capi.ProjectionIntegrator(self.capi.TRUNC_CONSTANT, cfg.truncation, cfg.weight, cfg.carve, cfg.carve_dist)

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()


def summarize_timeline(tl):
    """Medians (us) over the steps of a pass of the device timeline the library records (chs_get_device_timeline)."""
    def med(v):
        v = [x for x in v if x is not None]
        return round(float(np.median(v)) / 1000.0, 2) if v else None

    def d(a, b):
        return (a - b) if (a and b) else None
    rows = {"step": [], "hiz": [], "candidates": [], "bricks": [], "gap_bricks_to_next_hiz": [], "gap_hiz_to_candidates": [], "gap_candidates_to_bricks": [],
            "push": [], "push_end_to_hiz_start": [], "arrival_to_hiz_start": []}
    for i, cur in enumerate(tl):
        rows["hiz"].append(d(cur["hiz_end"], cur["hiz_start"]))
        rows["candidates"].append(d(cur["cand_end"], cur["cand_start"]))
        rows["bricks"].append(d(cur["bricks_end"], cur["bricks_start"]))
        rows["gap_hiz_to_candidates"].append(d(cur["cand_start"], cur["hiz_end"]))
        rows["gap_candidates_to_bricks"].append(d(cur["bricks_start"], cur["cand_end"]))
        rows["push"].append(d(cur["push_end"], cur["push_start"]))
        rows["push_end_to_hiz_start"].append(d(cur["hiz_start"], cur["push_end"]))
        rows["arrival_to_hiz_start"].append(d(cur["hiz_start"], cur["wait_end"]))
        if i > 0:
            rows["step"].append(d(cur["bricks_end"], tl[i - 1]["bricks_end"]))
            rows["gap_bricks_to_next_hiz"].append(d(cur["hiz_start"], tl[i - 1]["bricks_end"]))
    out = {k: med(v) for k, v in rows.items()}
    out["steps"] = [round(x / 1000.0, 1) for x in rows["step"] if x is not None]
    return out


def run_multi_agent(B: Bench, args):
    """The headline workload at N = world GPUs. Returns the JSON line (rank 0) or None."""
    torch = B.torch
    world, rank = B.world, B.rank
    cfg, cam = CFG, CFG.cam
    camv = cam.as_array()
    W, H = cam.width, cam.height
    warm, steps = args.warmup, args.steps
    T = warm + steps
    F = AGENTS * TS                                       # frames per step
    per = F // world
    mine = list(range(rank * per, (rank + 1) * per))      # indices into the step's arrival order that this rank ingests
    integ = B.integ(cfg)
    # frames this rank ingests (its agents), resident in HBM and in pinned host memory; poses of all agents
    frames = [step_frames(cfg, t, mine) for t in range(T)]
    poses = [step_poses(cfg, t) for t in range(T)]
    npx = W * H
    h_depth = torch.empty((T, per, H, W), dtype=torch.float32).pin_memory()
    for t in range(T):
        for j in range(per):
            h_depth[t, j].copy_(torch.from_numpy(frames[t][j][0]))
    d_depth = h_depth.to(B.dev)
    torch.cuda.synchronize(B.dev)

    def dev_ptrs(t):
        base = d_depth[t].data_ptr()
        out = [(base, None)] * F                          # only this rank's entries are read
        for j, a in enumerate(mine):
            out[a] = (base + 4 * npx * j, None)
        return out

    def host_arrays(t):
        out = [None] * F
        for j, a in enumerate(mine):
            out[a] = h_depth[t, j].numpy()
        return out

    prepared = {}

    def step_device(m, t):
        # the call's arguments (device pointers, poses) are marshalled once per step index: harness work of this leg, not part of the
        # path (the e2e leg marshals inside its timed region)
        if world > 1:
            if t not in prepared:
                prepared[t] = m.prepare_batch_distributed(integ, None, poses[t], camv, device_ptrs=dev_ptrs(t), channels=0, device_async=True)
            m.integrate_prepared_distributed(prepared[t])
        else:
            if t not in prepared:
                prepared[t] = m.prepare_batch(integ, None, poses[t], camv, device_ptrs=dev_ptrs(t), channels=0, device_async=True)
            m.integrate_prepared(prepared[t])

    # ---------------- value: K steps back to back, CUDA events on the map's stream, median of several passes --------------
    sampler = ClockSampler(range(world), enabled=(rank == 0))
    sampler.start()
    pass_ms, host_us = [], []
    # pass -1 is an untimed rehearsal of a whole pass (first use of the pool slabs, of the NCCL communicator's channels and of the
    # pinned snapshot ring); the timed passes that follow each start from a fresh map again
    debug_tl = os.environ.get("CHS_BENCH_DEBUG_TIMELINE") is not None
    for ip in range(-1, max(1, args.passes)):
        m = B.new_map(cfg)
        if debug_tl:
            m.set_profiling(2)
        for t in range(warm):
            step_device(m, t)
        m.synchronize()
        B.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(B.stream)
        h0 = time.perf_counter()
        for t in range(warm, T):
            step_device(m, t)
        h1 = time.perf_counter()
        e1.record(B.stream)
        m.synchronize()
        B.barrier()
        ms = B.max_over_ranks([e0.elapsed_time(e1)])[0]
        if debug_tl and rank == 0 and ip == max(1, args.passes) - 1:
            tl = m.device_timeline()
            t0 = tl[warm]["hiz_start"]
            sys.stderr.write("timed pass %.3f ms; per step (us from the first timed Hi-Z start): hiz_start, cand_start, bricks_start, bricks_end\n" % ms)
            for i, r in enumerate(tl):
                sys.stderr.write("  step %2d  %9.1f %9.1f %9.1f %9.1f\n" % (i, (r["hiz_start"] - t0) / 1e3, (r["cand_start"] - t0) / 1e3, (r["bricks_start"] - t0) / 1e3, (r["bricks_end"] - t0) / 1e3))
        if ip >= 0:
            pass_ms.append(ms)
            host_us.append(1e6 * (h1 - h0) / steps)
        if world > 1:
            m.comm_destroy()
        m.close()
    clocks = sampler.stop()
    t_value = float(np.median(pass_ms)) / 1000.0

    # ---------------- device timeline pass: the flow of a timed pass, the kernels stamp %globaltimer themselves ----------------
    m = B.new_map(cfg)
    m.set_profiling(2)
    for t in range(T):
        step_device(m, t)
    m.synchronize()
    B.barrier()
    timeline = summarize_timeline(m.device_timeline()[-steps:])
    if world > 1:
        m.comm_destroy()
    m.close()

    # ---------------- profiling pass: per-frame counters, kernel times (events inside the library), parity ----------------
    m = B.new_map(cfg)
    m.set_profiling(True)
    upd = bytes_alg = 0
    tk = dict(integrate=0.0, prepare=0.0, candidates=0.0, new_chunks=0.0, bricks_span=0.0)
    got_counters, per_step = [], []
    parity_steps = min(args.parity_steps, T)
    for t in range(T):
        step_device(m, t)
        sts = m.batch_stats()
        if t < parity_steps:
            got_counters += [[int(st[k]) for k in PARITY_KEYS] for st in sts]
        if t >= warm:
            tm = m.timings()
            upd += sum(st["n_upd"] for st in sts)
            bytes_alg += sum(algorithmic_bytes(st, cam, 0, False, cfg.chunk) for st in sts)
            for k in tk:
                tk[k] += tm[k + "_ms"] / 1000.0
            per_step.append((sts[-1]["brick_units"], sum(st["candidates"] for st in sts), sum(st["updated_chunks"] for st in sts),
                             sum(st["n_new"] for st in sts), sts[-1]["new_candidates"]))
    total_chunks = m.frame_stats()["total_chunks"]
    # parity: counters summed over ranks vs the oracle; N = 1 also state + dirty set from a second map fed the same steps
    parity = None
    if parity_steps > 0:
        gc = torch.tensor(got_counters, dtype=torch.int64, device=B.dev)
        if world > 1:
            B.dist.all_reduce(gc, op=B.dist.ReduceOp.SUM)
        if rank == 0:
            t0 = time.perf_counter()
            orc = make_cpu(cfg, False)
            want = []
            for t in range(parity_steps):
                for fr in step_frames(cfg, t):
                    cpu_integrate(orc, cfg, fr)
                    want.append(orc.frame_counters())
            bad = []
            for j, (g, w) in enumerate(zip(gc.cpu().tolist(), want)):
                for k, name in enumerate(PARITY_KEYS):
                    if g[k] != w[name]:
                        bad.append("frame %d %s: cuda %d oracle %d" % (j, name, g[k], w[name]))
            parity = {"frames": len(want), "counters": list(PARITY_KEYS), "counters_equal": not bad,
                      "oracle": "oracle/chisel_oracle.c (C restatement, pinned to the compiled reference)"}
            if world == 1:
                pm = B.new_map(cfg)
                for t in range(parity_steps):
                    step_device(pm, t)
                diff = state_equal(pm.state(), orc.state())
                if not np.array_equal(pm.dirty_ids(), orc.dirty_ids()):
                    diff.append("dirty set")
                parity["state_bit_exact"] = not diff
                parity["chunks"] = int(len(orc.state()[0]))
                bad += diff
                pm.close()
            parity["seconds"] = round(time.perf_counter() - t0, 1)
            if bad:
                print(json.dumps({"parity_check": parity, "mismatch": bad[:20]}), flush=True)
                raise SystemExit("bench.py: PARITY MISMATCH against the oracle: " + "; ".join(bad[:5]))

    # ---------------- mesh: re-mesh of everything the run left dirty (cold, then the same dirty set twice more) -----------
    if world > 1:
        m.sync_dirty()
    dirty = m.get_meshes_to_update()
    B.flush_l2(0)
    B.barrier()
    t0 = time.perf_counter()
    if world > 1:
        m.recompute_meshes_distributed(0)
    else:
        m.recompute_meshes()
    t_mesh_first = time.perf_counter() - t0
    mt_cold = m.timings()
    mesh_wall = []
    mt = mt_cold
    for _ in range(2):
        m.set_dirty(dirty)
        B.flush_l2(1)
        B.barrier()
        t0 = time.perf_counter()
        if world > 1:
            assert m._lib.chs_update_meshes_distributed(m._h, 0) == 0
        else:
            assert m._lib.chs_update_meshes(m._h) == 0
        m.synchronize()
        mesh_wall.append(time.perf_counter() - t0)
        if world == 1:
            mt = m.timings()
    mc = m.last_mesh_counts()
    V_halo = (cfg.chunk + 1) ** 3
    b_mc = mc["n_chunks"] * V_halo * 8 + mc["n_vertices"] * 24 + mc["n_grids"] * 12
    mesh_wall_s = B.max_over_ranks([float(np.min(mesh_wall))])[0]
    mesh_info = {"dirty_ids": int(len(dirty)), "remeshed_chunks": mc["n_chunks"], "triangles": mc["n_vertices"] // 3, "grids": mc["n_grids"],
                 "wall_ms": 1000.0 * mesh_wall_s, "first_call_wall_ms": 1000.0 * t_mesh_first,
                 "what": "chs_update_meshes_distributed: dirty-set union, ghost exchange, per-rank meshing, gather on rank 0 (wall clock, max over ranks)"
                 if world > 1 else "chs_update_meshes (wall clock incl. the host read of the vertex count)"}
    if world == 1:
        mesh_info.update({"device_ms": mt["mesh_ms"], "count_ms": mt["mesh_count_ms"], "emit_ms": mt["mesh_emit_ms"],
                          "first_call_device_ms": mt_cold["mesh_ms"], "algorithmic_bytes": b_mc,
                          "achieved_gbs": b_mc / (mt["mesh_ms"] * 1e-3) / 1e9 if mt["mesh_ms"] > 0 else None})
    if world > 1:
        m.comm_destroy()
    m.close()

    # ---------------- e2e: pinned host frames -> the same call -> counters of the previous step (depth-2 pipeline) --------
    m = B.new_map(cfg)

    def step_host(t):
        # arguments are marshalled here, inside the timed region, as a caller would
        if world > 1:
            m.integrate_batch_distributed(integ, host_arrays(t), poses[t], camv, host_async=True)
        else:
            m.integrate_batch(integ, host_arrays(t), poses[t], camv, host_async=True)

    for t in range(warm):
        step_host(t)
    m.synchronize()
    B.barrier()
    t0 = time.perf_counter()
    upd_e2e, prev = 0, None
    host_call = host_wait = 0.0
    for t in range(warm, T):
        c0 = time.perf_counter()
        step_host(t)
        tkt = m.last_batch_ticket()
        c1 = time.perf_counter()
        if prev is not None:
            upd_e2e += sum(st["n_upd"] for st in m.wait_batch(prev))
        host_call += c1 - c0
        host_wait += time.perf_counter() - c1
        prev = tkt
    upd_e2e += sum(st["n_upd"] for st in m.wait_batch(prev))
    B.barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        m.comm_destroy()
    m.close()

    # ---------------- e2e with the sensor's own encoding: 16UC1 millimetres (ROS), converted on the device (N = 1) ----------
    # chisel_ros receives depth as uint16 millimetres and converts on the host (CR Conversions.h:141-152); chs_frame.depth_mm moves
    # half the bytes over PCIe. The stream is the headline's quantised to millimetres, so its update count is its own. A failure
    # of this side leg never costs the headline line.
    e2e_mm = None
    if world == 1:
        try:
            h_mm = torch.empty((T, per, H, W), dtype=torch.uint16).pin_memory()
            mm_np = h_mm.numpy()
            for t in range(T):
                for j in range(per):
                    mm_np[t, j] = np.clip(np.nan_to_num(frames[t][j][0], nan=0.0) * 1000.0, 0, 65535).astype(np.uint16)
            m = B.new_map(cfg)
            for t in range(warm):
                m.integrate_batch(integ, [mm_np[t, j] for j in range(per)], poses[t], camv, host_async=True)
            m.synchronize()
            t0 = time.perf_counter()
            upd_mm, prev = 0, None
            for t in range(warm, T):
                m.integrate_batch(integ, [mm_np[t, j] for j in range(per)], poses[t], camv, host_async=True)
                tkt = m.last_batch_ticket()
                if prev is not None:
                    upd_mm += sum(st["n_upd"] for st in m.wait_batch(prev))
                prev = tkt
            upd_mm += sum(st["n_upd"] for st in m.wait_batch(prev))
            t_mm = time.perf_counter() - t0
            m.close()
            e2e_mm = {"value": upd_mm / t_mm / 1e9, "unit": UNIT, "frames_per_s": steps * F / t_mm, "ms_per_step": 1000.0 * t_mm / steps,
                      "h2d_bytes_per_step": 2 * npx * F, "h2d_gbs": 2 * npx * F * steps / t_mm / 1e9,
                      "what": "the e2e leg with 16UC1 millimetre depth (chs_frame.depth_mm, converted on the device as CR Conversions.h:141-152 does on the "
                              "host): half the PCIe bytes; the headline stream quantised to millimetres"}
        except Exception as exc:                                      # noqa: BLE001
            e2e_mm = {"error": "%s: %s" % (type(exc).__name__, exc)}

    t_e2e = B.max_over_ranks([t_e2e])[0]
    upd_total, upd_e2e_total, chunks_total = B.sum_over_ranks([float(upd), float(upd_e2e), float(total_chunks)])
    ksum = B.max_over_ranks([tk["integrate"], tk["prepare"], tk["candidates"]])
    if rank != 0:
        return None
    peak, peak_src = peaks()
    t_kernel = tk["integrate"]
    achieved = bytes_alg / t_kernel / 1e9 if t_kernel > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("multi_agent_n1")
        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"] if (tr and world == 1) else None
    except Exception:
        pass
    frame_bytes = 4 * npx
    line = {
        "metric": METRIC, "value": upd_total / t_value / 1e9, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": 1000.0 * t_value / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "frames_per_step": F, "frames_per_s": steps * F / t_value,
        "config": {"workload": WORKLOAD % TS,
                   "parallelism": ("chunk-hash shard x%d; rank r ingests frames [r*%d, (r+1)*%d) of the step's arrival order; the images are all-gathered in place over NVLink "
                                   "by the library (NCCL on its copy stream, beside the kernels of the previous step), every rank builds all Hi-Z pyramids and integrates the chunks it owns" % (world, per, per)) if world > 1 else "1 GPU",
                   "l2": "no flush: the timed region streams %d MB of frames (> 126 MB L2 from %d steps on); the map working set stays in L2 as in a live stream"
                         % ((frame_bytes * F * steps) >> 20, (126 << 20) // (frame_bytes * F) + 1),
                   "timing": "CUDA events on the map's stream around the %d timed steps, barrier + synchronize on both sides, max over ranks; median of %d passes %s ms"
                             % (steps, len(pass_ms), [round(x, 3) for x in pass_ms]),
                   "host_enqueue_us_per_step": float(np.median(host_us)),
                   "rank0_device_timeline_us": timeline,
                   "voxel_updates_per_step": upd_total / steps, "map_chunks": chunks_total,
                   "rank0_kernels_us_per_step": {"hiz": 1e6 * tk["prepare"] / steps, "candidates": 1e6 * tk["candidates"] / steps,
                                                 "wait_for_colour_pack": 1e6 * tk["new_chunks"] / steps, "bricks": 1e6 * tk["integrate"] / steps,
                                                 "bricks_first_cta_to_last_cta": 1e6 * tk["bricks_span"] / steps,
                                                 "max_over_ranks": {"bricks": 1e6 * ksum[0] / steps, "hiz": 1e6 * ksum[1] / steps, "candidates": 1e6 * ksum[2] / steps}},
                   "rank0_per_step": {"candidate_chunks": float(np.mean([p[1] for p in per_step])), "brick_units": float(np.mean([p[0] for p in per_step])),
                                      "updated_chunks": float(np.mean([p[2] for p in per_step])), "new_chunks": float(np.mean([p[3] for p in per_step])),
                                      "new_chunk_candidates": float(np.mean([p[4] for p in per_step]))}},
        "e2e": {"value": upd_e2e_total / t_e2e / 1e9, "unit": UNIT, "frames_per_s": steps * F / t_e2e, "ms_per_step": 1000.0 * t_e2e / steps,
                "h2d_bytes_per_step": frame_bytes * F, "d2h_bytes_per_step": 88 * F * world,
                "h2d_gbs_per_rank": frame_bytes * per * steps / t_e2e / 1e9,
                "rank0_host_us_per_step": {"in_the_call_incl_marshalling": 1e6 * host_call / steps, "waiting_for_the_previous_step": 1e6 * host_wait / steps},
                "timing": "wall clock, max over ranks; per step chs_integrate_batch%s(pinned host frames, CHS_MEM_HOST_ASYNC; arguments marshalled inside the "
                          "timed region), then chs_wait_batch of the PREVIOUS step's counters (depth-2 pipeline)" % ("_distributed" if world > 1 else "")},
        "e2e_depth_mm": e2e_mm,
        # Hi-Z, candidates, bricks per step; N > 1: plus the push kernel and the arrival wait of the frame exchange
        "gpu_launches": (3 if world == 1 else 5) * steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "batch_bricks_fast_kernel<16,0,0>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload (profiles/r02_traffic.json)" if traffic else None,
                     "algorithmic_bytes_per_launch": bytes_alg / steps, "kernel_ms_per_launch": 1000.0 * t_kernel / steps,
                     "kernel_span_ms_per_launch": 1000.0 * tk["bricks_span"] / steps,
                     "note": "rank 0's brick kernel: algorithmic bytes = B_int of SURVEY 8(d) of the chunks this rank owns, summed over the step's frames (every rank "
                             "reads all of them); duration from CUDA events recorded by the library around the kernel on its stream (kernel_span: first CTA start to "
                             "last CTA end from %globaltimer, i.e. without launch and event-record overhead)"},
        "parity_check": parity,
        "mesh": mesh_info,
    }
    return line


def run_single_agent(B: Bench, args):
    """Second line (N = 1): BASELINE configs[1], 752x480 depth + colour, 2 cm; step = 10 consecutive frames in one fused call."""
    torch = B.torch
    cfg, cam = CFG2, CFG2.cam
    camv = cam.as_array()
    W, H, ch = cam.width, cam.height, 3
    Bf = 10
    warm, steps = args.warmup, args.steps2
    T = warm + steps
    n_frames = min(T * Bf, cfg.n_frames)
    integ = B.integ(cfg)
    frames = [scenes.stream_frame(cfg, f) for f in range(n_frames)]
    npx = W * H
    h_depth = torch.empty((n_frames, H, W), dtype=torch.float32).pin_memory()
    h_col = torch.empty((n_frames, H, W, ch), dtype=torch.uint8).pin_memory()
    for i, fr in enumerate(frames):
        h_depth[i].copy_(torch.from_numpy(fr[0]))
        h_col[i].copy_(torch.from_numpy(fr[1]))
    d_depth, d_col = h_depth.to(B.dev), h_col.to(B.dev)
    torch.cuda.synchronize(B.dev)

    def ids(t):
        return [(t * Bf + j) % n_frames for j in range(Bf)]

    prepared = {}

    def step_device(m, t):
        if t not in prepared:
            prepared[t] = m.prepare_batch(integ, None, [frames[i][2] for i in ids(t)], camv,
                                          device_ptrs=[(d_depth[i].data_ptr(), d_col[i].data_ptr()) for i in ids(t)], channels=ch, device_async=True)
        m.integrate_prepared(prepared[t])

    pass_ms = []
    for _ in range(max(1, args.passes)):
        m = B.new_map(cfg, sharded=False)
        for t in range(warm):
            step_device(m, t)
        m.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(B.stream)
        for t in range(warm, T):
            step_device(m, t)
        e1.record(B.stream)
        m.synchronize()
        pass_ms.append(e0.elapsed_time(e1))
        m.close()
    t_value = float(np.median(pass_ms)) / 1000.0
    # L2-flushed, one bracket per step (round 1's definition of `value`)
    m = B.new_map(cfg, sharded=False)
    for t in range(warm):
        step_device(m, t)
    m.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        B.flush_l2(k)
        ev[k][0].record(B.stream)
        step_device(m, warm + k)
        ev[k][1].record(B.stream)
    m.synchronize()
    t_flushed = sum(a.elapsed_time(b) for a, b in ev) / 1000.0
    m.close()
    # profiling pass (L2 flushed before every step: cold-cache kernel times, like the ncu captures). Plain CHS_MEM_DEVICE here: every
    # kernel of the step then runs on the map's stream, behind the flush, so the per-kernel event times are clean
    m = B.new_map(cfg, sharded=False)
    m.set_profiling(True)
    upd = bytes_alg = 0
    tk = dict(integrate=0.0, prepare=0.0, candidates=0.0, new_chunks=0.0, bricks_span=0.0)
    prepared_sync = {}
    for t in range(T):
        if t >= warm:
            B.flush_l2(t)
        prepared_sync[t] = m.prepare_batch(integ, None, [frames[i][2] for i in ids(t)], camv,
                                           device_ptrs=[(d_depth[i].data_ptr(), d_col[i].data_ptr()) for i in ids(t)], channels=ch)
        m.integrate_prepared(prepared_sync[t])
        sts = m.batch_stats()
        if t >= warm:
            tm = m.timings()
            upd += sum(st["n_upd"] for st in sts)
            bytes_alg += sum(algorithmic_bytes(st, cam, ch, True, cfg.chunk) for st in sts)
            for k in tk:
                tk[k] += tm[k + "_ms"] / 1000.0
    m.close()
    # e2e
    m = B.new_map(cfg, sharded=False)

    def step_host(t):
        m.integrate_batch(integ, [h_depth[i].numpy() for i in ids(t)], [frames[i][2] for i in ids(t)], camv, [h_col[i].numpy() for i in ids(t)], host_async=True)

    for t in range(warm):
        step_host(t)
    m.synchronize()
    t0 = time.perf_counter()
    upd_e2e, prev = 0, None
    host_call = host_wait = 0.0
    for t in range(warm, T):
        c0 = time.perf_counter()
        step_host(t)
        tkt = m.last_batch_ticket()
        c1 = time.perf_counter()
        if prev is not None:
            upd_e2e += sum(st["n_upd"] for st in m.wait_batch(prev))
        host_call += c1 - c0
        host_wait += time.perf_counter() - c1
        prev = tkt
    upd_e2e += sum(st["n_upd"] for st in m.wait_batch(prev))
    t_e2e = time.perf_counter() - t0
    m.close()
    # parity: the first two steps of this path (state, dirty set, counters) against the oracle
    parity = None
    if args.parity_steps > 0:
        pm = B.new_map(cfg, sharded=False)
        orc = make_cpu(cfg, False)
        bad = []
        for t in range(min(2, T)):
            step_device(pm, t)
            got = pm.batch_stats()
            for j, i in enumerate(ids(t)):
                cpu_integrate(orc, cfg, frames[i])
                w = orc.frame_counters()
                bad += ["frame %d %s: cuda %d oracle %d" % (i, k, got[j][k], w[k]) for k in PARITY_KEYS if got[j][k] != w[k]]
        bad += state_equal(pm.state(), orc.state())
        if not np.array_equal(pm.dirty_ids(), orc.dirty_ids()):
            bad.append("dirty set")
        pm.close()
        parity = {"frames": 2 * Bf, "state_bit_exact": not bad, "counters_equal": not bad}
        if bad:
            print(json.dumps({"parity_check_single_agent": parity, "mismatch": bad[:20]}), flush=True)
            raise SystemExit("bench.py: PARITY MISMATCH (configs[1]) against the oracle: " + "; ".join(bad[:5]))
    peak, peak_src = peaks()
    achieved = bytes_alg / tk["integrate"] / 1e9 if tk["integrate"] > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("single_agent_color")
        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"] if tr else None
    except Exception:
        pass
    h2d = (4 + ch) * npx * Bf
    return {
        "workload": WORKLOAD2 % Bf, "value": upd / t_value / 1e9, "unit": UNIT, "frames_per_s": steps * Bf / t_value, "ms_per_step": 1000.0 * t_value / steps,
        "steps": steps, "frames_per_step": Bf, "passes_ms": [round(x, 3) for x in pass_ms],
        "l2_flushed_per_step": {"value": upd / t_flushed / 1e9, "ms_per_step": 1000.0 * t_flushed / steps,
                                "note": "one CUDA-event bracket per step, 256 MiB write + 256 MiB read between steps (round 1's definition of value)"},
        "kernels_us_per_step_cold_l2": {"hiz": 1e6 * tk["prepare"] / steps, "candidates": 1e6 * tk["candidates"] / steps,
                                        "wait_for_colour_pack": 1e6 * tk["new_chunks"] / steps, "bricks": 1e6 * tk["integrate"] / steps,
                                        "bricks_first_cta_to_last_cta": 1e6 * tk["bricks_span"] / steps},
        "roofline": {"bound": "hbm", "kernel": "batch_bricks_fast_kernel<16,1,1>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_alg / steps,
                     "kernel_ms_per_launch": 1000.0 * tk["integrate"] / steps, "kernel_span_ms_per_launch": 1000.0 * tk["bricks_span"] / steps},
        "e2e": {"value": upd_e2e / t_e2e / 1e9, "unit": UNIT, "frames_per_s": steps * Bf / t_e2e, "ms_per_step": 1000.0 * t_e2e / steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 88 * Bf, "h2d_gbs": h2d * steps / t_e2e / 1e9,
                "host_us_per_step": {"in_the_call_incl_marshalling": 1e6 * host_call / steps, "waiting_for_the_previous_step": 1e6 * host_wait / steps}},
        "parity_check": parity,
    }


def hall_side_line(B: Bench, args):
    """configs[3] shape: 50x50x5 m pillar hall at 1 cm, 640x480 depth, lawn-mower sweep in batches of 8, then ONE re-mesh of everything dirty."""
    torch, capi = B.torch, B.capi
    hall = scenes.hall(seed=3)
    cam = scenes.Camera(525.0, 525.0, 319.5, 239.5, 640, 480, near=0.05, far=5.0)
    res = 0.01
    integ = capi.ProjectionIntegrator(capi.TRUNC_CONSTANT, float(np.float32(4.0) * np.float32(res)), 1.0, True, 0.05)
    n = args.hall_frames
    poses = [scenes.yaw_pose(0.35 * (i % 6) + 0.05, (-20.0 + 1.0 * (i // 6) + 0.15 * (i % 6), -20.0 + 0.9 * (i % 6), 0.0)) for i in range(n)]
    dd = [torch.from_numpy(scenes.render(hall, cam, p)[0]).to(B.dev) for p in poses]
    m = capi.Chisel(16, res, False, device=B.local, stream=B.stream.cuda_stream, initial_chunks=args.pool_chunks)
    m.set_profiling(True)
    t_dev, upd, kern, cand = 0.0, 0, {}, 0
    for i in range(0, n, 8):
        B.flush_l2(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(B.stream)
        m.integrate_batch(integ, None, poses[i:i + 8], cam.as_array(), device_ptrs=[(d.data_ptr(), None) for d in dd[i:i + 8]])
        e1.record(B.stream)
        sts = m.batch_stats()
        t_dev += e0.elapsed_time(e1) / 1000.0
        upd += sum(s["n_upd"] for s in sts)
        tm = m.timings()
        for k in ("prepare", "candidates", "integrate"):
            kern[k] = kern.get(k, 0.0) + tm[k + "_ms"] / (n // 8)
        cand = sum(s["candidates"] for s in sts)
    dirty = m.get_meshes_to_update()
    m._lib.chs_update_meshes(m._h)
    for _ in range(2):
        m.set_dirty(dirty)
        B.flush_l2(1)
        assert m._lib.chs_update_meshes(m._h) == 0
    mt, mc = m.timings(), m.last_mesh_counts()
    b_mc = mc["n_chunks"] * 17 ** 3 * 8 + mc["n_vertices"] * 24 + mc["n_grids"] * 12
    out = {"workload": "configs[3] shape: 50x50x5 m pillar hall, 1 cm voxels, 640x480 depth, %d frames in batches of 8, then one re-mesh of the whole dirty set" % n,
           "integration": {"value": upd / t_dev / 1e9, "unit": UNIT, "frames_per_s": n / t_dev, "voxel_updates_per_frame": upd / n,
                           "ms_per_8_frame_batch": {"hiz": kern["prepare"], "candidates": kern["candidates"], "bricks": kern["integrate"]},
                           "candidate_chunks_last_batch": cand},
           "remesh": {"dirty_ids": int(len(dirty)), "remeshed_chunks": mc["n_chunks"], "triangles": mc["n_vertices"] // 3,
                      "device_ms": mt["mesh_ms"], "count_ms": mt["mesh_count_ms"], "emit_ms": mt["mesh_emit_ms"],
                      "algorithmic_bytes": b_mc, "achieved_gbs": b_mc / (mt["mesh_ms"] * 1e-3) / 1e9 if mt["mesh_ms"] > 0 else None}}
    m.close()
    return out


def run_cuda(args):
    B = Bench(args)
    line = run_multi_agent(B, args)
    if B.rank == 0 and B.world == 1:
        if not args.quick:
            try:
                line["single_agent_color"] = run_single_agent(B, args)
            except SystemExit:
                raise
            except Exception as ex:                              # the second line must never cost the headline
                line["single_agent_color"] = {"error": repr(ex)}
        if not args.quick and not args.no_side_lines:
            try:
                line["side_lines"] = {"hall_1cm": hall_side_line(B, args)}
            except Exception as ex:
                line["side_lines"] = {"error": repr(ex)}
        if not args.no_cpu:
            frames = []
            for t in range(args.cpu_steps):
                frames += step_frames(CFG, t)
            r = cpu_arm(CFG, frames, 0, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "host_cores": r["host_cores"], "kind": r["kind"],
                                    "frames_per_s": r["fps"],
                                    "sample": "first %d frames (%d time steps) of the same stream from an empty map, whole frames, %.1f s; the reference's "
                                              "depth-only path runs on one thread (Chisel.h:71-72)" % (r["frames_done"], r["frames_done"] // AGENTS, r["seconds"])}
    if B.rank == 0:
        print(json.dumps(line), flush=True)
    if B.world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="timed steps of the headline workload (8 frames each)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--steps2", type=int, default=16, help="timed steps of the second line (configs[1], 10 frames each)")
    ap.add_argument("--passes", type=int, default=5, help="timed passes (each from a fresh map); value = the median")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-side-lines", action="store_true", help="skip the 1 cm hall side line")
    ap.add_argument("--quick", action="store_true", help="A/B runs: the headline workload only")
    ap.add_argument("--cpu-steps", type=int, default=1, help="steps of the cpu_baseline sample")
    ap.add_argument("--time-steps-per-step", type=int, default=2, choices=[1, 2],
                    help="time steps (8 frames each, one per agent) handed over per call: the fused kernels take up to 16 frames")
    ap.add_argument("--hall-frames", type=int, default=24)
    ap.add_argument("--pool-chunks", type=int, default=262144,
                    help="pre-sized chunk pool (chunks) so that no slab / hash growth lands inside the timed region")
    ap.add_argument("--cpu-budget", type=float, default=150.0)
    ap.add_argument("--parity-steps", type=int, default=2,
                    help="steps (from the empty map) whose per-frame counters, voxel state and dirty set are compared with the CPU oracle; 0 = off")
    args = ap.parse_args()
    global TS
    TS = args.time_steps_per_step
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

"""Deterministic synthetic analytic-scene depth(+colour) streams.

All scenes are the interior of an axis-aligned box, optionally with axis-aligned box obstacles
(pillars), rendered by exact ray/box intersection in float64 and stored as float32 z-depth in
metres. Camera convention follows the reference's use (PinholeCamera.cpp:38-45): x right, y down,
z forward, pose = camera -> world (Conversions.h:216 inverts the looked-up TF), pixel (col, row)
covers [col, col+1) x [row, row+1) and rays are cast through pixel centres.

Nothing here reads /root/reference; the same NumPy code runs in the build container and on the
GPU box, so oracle and CUDA runs see bit-identical inputs.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass(frozen=True)
class Camera:
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int
    near: float = 0.05
    far: float = 5.0

    def as_array(self) -> np.ndarray:
        """[fx, fy, cx, cy, W, H, near, far] float32 — the layout of chs_camera / the oracle."""
        return np.array([self.fx, self.fy, self.cx, self.cy, self.width, self.height, self.near, self.far],
                        dtype=np.float32)


KINECT_640 = Camera(525.0, 525.0, 319.5, 239.5, 640, 480)
# config/euroc/euroc_config_multi_agent.yaml:11-22 of the reference (752x480 EuRoC cam0)
EUROC_752 = Camera(461.6, 460.3, 363.0, 248.1, 752, 480)


@dataclasses.dataclass(frozen=True)
class Scene:
    lo: tuple          # room min corner
    hi: tuple          # room max corner
    boxes: tuple = ()  # obstacles: ((lo3, hi3), ...)


ROOM = Scene((-3.0, -2.5, -1.5), (3.0, 2.5, 1.5))


def hall(seed: int = 0, size=(50.0, 50.0, 5.0), spacing: float = 5.0) -> Scene:
    """50 x 50 x 5 m hall with a regular grid of square pillars so that surfaces stay within `far`."""
    rng = np.random.RandomState(seed)
    sx, sy, sz = size
    boxes = []
    nx, ny = int(sx / spacing), int(sy / spacing)
    for i in range(nx):
        for j in range(ny):
            cx = -sx / 2 + (i + 0.5) * spacing + rng.uniform(-0.5, 0.5)
            cy = -sy / 2 + (j + 0.5) * spacing + rng.uniform(-0.5, 0.5)
            h = rng.uniform(0.3, 0.6)
            boxes.append(((cx - h, cy - h, -sz / 2), (cx + h, cy + h, sz / 2)))
    return Scene((-sx / 2, -sy / 2, -sz / 2), (sx / 2, sy / 2, sz / 2), tuple(boxes))


def yaw_pose(theta: float, pos) -> np.ndarray:
    """Camera looking horizontally along (cos t, sin t, 0), world z up. Returns 3x4 [R|t] float32."""
    c, s = math.cos(theta), math.sin(theta)
    right = (s, -c, 0.0)
    down = (0.0, 0.0, -1.0)
    fwd = (c, s, 0.0)
    m = np.zeros((3, 4), dtype=np.float64)
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = right, down, fwd, pos
    return m.astype(np.float32)


def orbit_pose(f: int, n: int, phase: float = 0.0) -> np.ndarray:
    """Config-1 trajectory: yaw 2*pi*f/n about world z, position (0.5 cos 2t, 0.5 sin 2t, 0.1 sin 3t)."""
    t = 2.0 * math.pi * f / n + phase
    return yaw_pose(t, (0.5 * math.cos(2 * t), 0.5 * math.sin(2 * t), 0.1 * math.sin(3 * t)))


def _ray_dirs(cam: Camera, pose: np.ndarray):
    u = (np.arange(cam.width, dtype=np.float64) + 0.5 - cam.cx) / cam.fx
    v = (np.arange(cam.height, dtype=np.float64) + 0.5 - cam.cy) / cam.fy
    uu, vv = np.meshgrid(u, v)
    d_cam = np.stack([uu, vv, np.ones_like(uu)], axis=-1)            # z component 1 => t is z-depth
    R = pose[:, :3].astype(np.float64)
    return d_cam @ R.T, pose[:, 3].astype(np.float64)


def render(scene: Scene, cam: Camera, pose: np.ndarray, color: bool = False, channels: int = 3,
           nan_frac: float = 0.0, noise_sigma: float = 0.0, seed: int = 0):
    """Returns (depth float32 [H,W], color uint8 [H,W,channels] | None)."""
    d, o = _ray_dirs(cam, pose)
    lo = np.asarray(scene.lo, dtype=np.float64)
    hi = np.asarray(scene.hi, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # exit distance from inside the room
        t_exit = np.where(d > 0, (hi - o) * inv, np.where(d < 0, (lo - o) * inv, np.inf))
        t = t_exit.min(axis=-1)
        for blo, bhi in scene.boxes:
            blo = np.asarray(blo, dtype=np.float64)
            bhi = np.asarray(bhi, dtype=np.float64)
            t0 = (blo - o) * inv
            t1 = (bhi - o) * inv
            tn = np.minimum(t0, t1)
            tf = np.maximum(t0, t1)
            # axis-parallel rays: inside the slab -> (-inf, inf), else miss
            par = d == 0
            inside = (o >= blo) & (o <= bhi)
            tn = np.where(par, np.where(inside, -np.inf, np.inf), tn)
            tf = np.where(par, np.where(inside, np.inf, -np.inf), tf)
            tn = tn.max(axis=-1)
            tf = tf.min(axis=-1)
            hit = (tn <= tf) & (tn > 0)
            t = np.where(hit & (tn < t), tn, t)
    depth = t
    rng = np.random.RandomState(seed)
    if noise_sigma > 0:
        depth = depth + rng.normal(0.0, noise_sigma, size=depth.shape)
    depth32 = depth.astype(np.float32)
    if nan_frac > 0:
        mask = rng.uniform(size=depth.shape) < nan_frac
        depth32 = np.where(mask, np.float32(np.nan), depth32)
    col = None
    if color:
        p = o + d * t[..., None]
        # procedural colour of the hit point: 25 cm checker modulated by smooth gradients
        chk = (np.floor(p[..., 0] * 4) + np.floor(p[..., 1] * 4) + np.floor(p[..., 2] * 4)) % 2
        r = 96 + 64 * chk + 60 * np.sin(p[..., 0] * 1.7)
        g = 96 + 64 * (1 - chk) + 60 * np.sin(p[..., 1] * 2.3 + 1.0)
        b = 128 + 100 * np.sin(p[..., 2] * 3.1 + 2.0)
        rgb = np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)
        if channels == 1:
            col = rgb[..., 1:2].copy()
        elif channels == 3:
            col = rgb[..., ::-1].copy()                                # BGR (ColorImage.h:80-85)
        elif channels == 4:
            col = np.concatenate([rgb[..., ::-1], np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=-1)
        else:
            raise ValueError("channels must be 1, 3 or 4")
    return np.ascontiguousarray(depth32), col


@dataclasses.dataclass(frozen=True)
class StreamConfig:
    """One BASELINE.json config."""
    name: str
    scene: Scene
    cam: Camera
    resolution: float
    n_frames: int
    color: bool
    trunc_voxels: float = 4.0
    chunk: int = 16
    nan_frac: float = 0.0
    agents: int = 1
    carve: bool = True
    carve_dist: float = 0.05
    weight: float = 1.0

    @property
    def truncation(self) -> float:
        return float(np.float32(self.trunc_voxels) * np.float32(self.resolution))


CONFIG1 = StreamConfig("room-640x480-5cm-depth", ROOM, KINECT_640, 0.05, 100, False)
CONFIG2 = StreamConfig("euroc-752x480-2cm-color", ROOM, EUROC_752, 0.02, 200, True, nan_frac=0.02)
CONFIG3 = StreamConfig("4agent-640x480-5cm-depth", ROOM, KINECT_640, 0.05, 100, False, agents=4)
CONFIG5 = StreamConfig("8agent-640x480-2cm-depth", ROOM, KINECT_640, 0.02, 60, False, agents=8)


def stream_frame(cfg: StreamConfig, f: int, agent: int = 0):
    """Frame `f` of agent `agent`: (depth, color|None, pose 3x4 float32). Agents are phase-shifted
    copies of the orbit trajectory (config 3: offsets pi/2)."""
    phase = agent * (2.0 * math.pi / max(cfg.agents, 1))
    pose = orbit_pose(f, cfg.n_frames, phase)
    depth, col = render(cfg.scene, cfg.cam, pose, color=cfg.color, nan_frac=cfg.nan_frac,
                        seed=1 + f * 131 + agent * 7919)
    return depth, col, pose


def interleaved(cfg: StreamConfig):
    """Canonical multi-agent frame order: round-robin by agent per time step."""
    for f in range(cfg.n_frames):
        for a in range(cfg.agents):
            yield f, a

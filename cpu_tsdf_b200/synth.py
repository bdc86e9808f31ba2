"""Deterministic synthetic depth frames for tests and bench.

Scenes are closed-form so every pixel's depth is known analytically:
  * S1 "sphere-in-room": axis-aligned cubic room of half-side R centred at the origin plus a
    sphere of radius 0.25 R at the origin; the camera orbits on the circle of radius r in the
    plane y = 0 and looks at the origin.
  * S2 "interior" (ICL-NUIM-shaped: every pixel valid): same room, camera looks outward.
Camera convention follows the reference's callers (src/prog/integrate.cpp:66-69, 350-353):
x right, y down (= world +y), z forward; clouds are organized W x H, row-major, in the SENSOR
frame, NaN z = invalid, and the pose is camera->world.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class Camera:
    fx: float = 525.0
    fy: float = 525.0
    cx: float = 319.5
    cy: float = 239.5
    width: int = 640
    height: int = 480


@dataclass
class Scene:
    room_half: float = 1.4      # R
    cam_radius: float = 1.0     # r
    outward: bool = False       # S2 when True
    sphere: bool = True

    @property
    def sphere_radius(self) -> float:
        return 0.25 * self.room_half


S1 = Scene(room_half=1.4, cam_radius=1.0, outward=False)
S2 = Scene(room_half=2.0, cam_radius=0.5, outward=True)


def orbit_pose(scene: Scene, frame: int, nframes: int = 100) -> np.ndarray:
    """camera->world 4x4 (float64) for frame f: theta = 2 pi f / max(N, 100)."""
    theta = 2.0 * math.pi * frame / max(nframes, 100)
    c = np.array([scene.cam_radius * math.cos(theta), 0.0, scene.cam_radius * math.sin(theta)])
    fwd = c / np.linalg.norm(c)
    if not scene.outward:
        fwd = -fwd
    down = np.array([0.0, 1.0, 0.0])
    right = np.cross(down, fwd)
    right /= np.linalg.norm(right)
    pose = np.eye(4)
    pose[:3, 0] = right
    pose[:3, 1] = down
    pose[:3, 2] = fwd
    pose[:3, 3] = c
    return pose


def _ray_depth(scene: Scene, pose: np.ndarray, cam: Camera):
    u = np.arange(cam.width, dtype=np.float64)
    v = np.arange(cam.height, dtype=np.float64)
    xn = (u[None, :] - cam.cx) / cam.fx
    yn = (v[:, None] - cam.cy) / cam.fy
    xn, yn = np.broadcast_arrays(xn, yn)
    d_cam = np.stack([xn, yn, np.ones_like(xn)], axis=-1)          # z component 1 => t is z-depth
    d_w = d_cam @ pose[:3, :3].T
    o = pose[:3, 3]
    R = scene.room_half
    with np.errstate(divide="ignore", invalid="ignore"):
        t_axis = (np.where(d_w > 0, R, -R) - o[None, None, :]) / d_w
    t_axis = np.where(np.isfinite(t_axis) & (t_axis > 0), t_axis, np.inf)
    t = t_axis.min(axis=-1)
    if scene.sphere:
        rs = scene.sphere_radius
        a = (d_w * d_w).sum(-1)
        b = 2.0 * (d_w @ o)
        c = float(o @ o) - rs * rs
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            ts = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * a)
        hit = (disc > 0) & (ts > 0)
        t = np.where(hit & (ts < t), ts, t)
    return t, xn, yn, d_w, o


def make_frame(scene: Scene, pose: np.ndarray, cam: Camera = Camera(), *, color: bool = False,
               noise_seed: int | None = None, frame: int = 0, max_depth: float | None = None,
               dropout: float = 0.0) -> np.ndarray:
    """Organized cloud as a float32 array [H, W, 4] (pcl::PointXYZ, 16 B/pt) or, with
    color=True, [H, W, 8] (pcl::PointXYZRGBA, 32 B/pt: xyz, 1.0, then the bytes b,g,r,a packed
    in float slot 4).  Depth noise sigma(z) = 0.0012 + 0.0019 (z - 0.4)^2 m when noise_seed is
    given.  Pixels deeper than max_depth, and a `dropout` fraction of pixels, become NaN."""
    t, xn, yn, d_w, o = _ray_depth(scene, pose, cam)
    rng = None
    if noise_seed is not None:
        rng = np.random.default_rng([noise_seed, frame])
        sigma = 0.0012 + 0.0019 * (t - 0.4) ** 2
        t = t + rng.standard_normal(t.shape) * sigma
    valid = np.isfinite(t) & (t > 0)
    if max_depth is not None:
        valid &= t <= max_depth
    if dropout > 0.0:
        drng = np.random.default_rng([noise_seed or 0, frame, 7])
        valid &= drng.random(t.shape) >= dropout
    nfl = 8 if color else 4
    out = np.empty((cam.height, cam.width, nfl), dtype=np.float32)
    t32 = np.where(valid, t, np.nan)
    out[..., 0] = (xn * t32).astype(np.float32)
    out[..., 1] = (yn * t32).astype(np.float32)
    out[..., 2] = t32.astype(np.float32)
    out[..., 3] = 1.0
    if color:
        hit = o[None, None, :] + d_w * np.where(valid, t, 0.0)[..., None]
        cell = np.floor(hit / 0.032).astype(np.int64)
        par = (cell.sum(-1) & 1).astype(np.uint8)
        rgba = np.zeros((cam.height, cam.width, 4), dtype=np.uint8)
        # PCL byte order b, g, r, a
        rgba[..., 2] = np.where(par == 1, 220, 40) + (np.mod(cell[..., 0], 8) * 4).astype(np.uint8)
        rgba[..., 1] = np.where(par == 1, 180, 60) + (np.mod(cell[..., 1], 8) * 4).astype(np.uint8)
        rgba[..., 0] = np.where(par == 1, 90, 200) + (np.mod(cell[..., 2], 8) * 4).astype(np.uint8)
        rgba[..., 3] = 255
        out[..., 4] = rgba.view(np.float32)[..., 0]
        out[..., 5:] = 0.0
    return np.ascontiguousarray(out)


def analytic_depth(scene: Scene, pose: np.ndarray, cam: Camera = Camera()) -> np.ndarray:
    """Exact z-depth per pixel (float64), for known-answer tests."""
    return _ray_depth(scene, pose, cam)[0]

#!/usr/bin/env python3
"""Secondary measurements (not the BASELINE metric): renderView and marching-cubes time on the GPU vs
the CPU reference, on BASELINE.json configs[1] (512^3 orbit) and configs[2] (2048^3 interior, colour).
Prints one JSON line per config.  Usage: python tests/perf/bench_extra.py [--frames N] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cpu_tsdf_b200 as pkg  # noqa: E402
from cpu_tsdf_b200 import synth  # noqa: E402

CAM = synth.Camera()


def run(name, res, size, scene, nframes, color, cpu):
    vol = pkg.TSDFVolumeOctree(device=0, pool_log2=19)
    vol.setResolution(res, res, res); vol.setGridSize(size, size, size)
    vol.setCameraIntrinsics(CAM.fx, CAM.fy, CAM.cx, CAM.cy); vol.setIntegrateColor(color); vol.reset()
    orc = None
    if cpu:
        from oracle import oracle_py
        from oracle.oracle_py import OracleVolume
        kind = "reference" if os.path.exists(oracle_py.REF_LIB) else "port"
        orc = OracleVolume(kind=kind, xres=res, yres=res, zres=res, xsize=size, ysize=size, zsize=size, cx=CAM.cx, cy=CAM.cy, integrate_color=int(color))
        orc.reset()
    for f in range(nframes):
        pose = synth.orbit_pose(scene, f * (100 // nframes), 100)
        cloud = synth.make_frame(scene, pose, CAM, color=color, noise_seed=12345, frame=f)
        vol.integrateCloud(cloud, None, pose)
        if orc:
            orc.integrate(cloud, pose)
    vol.sync()
    pose = synth.orbit_pose(scene, 7, 100)
    out = {"config": name, "frames": nframes, "bricks": int(vol.stats().n_bricks)}
    vol.renderView(pose, 1)                                   # warm-up
    t = []
    for _ in range(3):
        vol.profile_begin(); r = vol.renderView(pose, 1); pr = vol.profile_end(); t.append(pr.ms_elapsed)
    out["render_ms_gpu_stream"] = min(t)
    t0 = time.perf_counter(); r = vol.renderView(pose, 1); out["render_ms_gpu_call"] = 1e3 * (time.perf_counter() - t0)
    out["render_hits"] = int(np.isfinite(r[..., 2]).sum())
    mc = pkg.MarchingCubesTSDFOctree(); mc.setInputTSDF(vol); mc.setMinWeight(2.0); mc.setColorByRGB(color)
    mc.reconstruct()
    t0 = time.perf_counter(); v, c, _ = mc.reconstruct(); out["mesh_ms_gpu_call"] = 1e3 * (time.perf_counter() - t0)
    out["mesh_triangles"] = int(len(v) // 3)
    if orc:
        t0 = time.perf_counter(); ro = orc.render(pose, 1); out["render_ms_cpu"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter(); vo, co = orc.mesh(2.0, 1 if color else 0); out["mesh_ms_cpu"] = 1e3 * (time.perf_counter() - t0)
        out["render_depth_max_abs_diff"] = float(np.nanmax(np.abs(ro[..., 2] - r[..., 2]))) if out["render_hits"] else 0.0
        out["render_same_hits"] = bool(np.array_equal(np.isfinite(ro[..., 2]), np.isfinite(r[..., 2])))
        out["mesh_same_count"] = bool(len(vo) == len(v))
        out["cpu_kind"] = kind
        out["cpu_threads"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.only in ("", "c2"):
        run("configs[1]: 512^3 / 3 m, S1 orbit", 512, 3.0, synth.S1, a.frames, False, not a.no_cpu)
    if a.only in ("", "c3"):
        run("configs[2]: 2048^3 / 10 m, S2 interior, colour", 2048, 10.0, synth.S2, a.frames, True, not a.no_cpu)

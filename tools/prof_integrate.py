#!/usr/bin/env python3
"""Short device-resident integrate loop for profiling under ncu (bench.py's workload, no timing of its own)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import cpu_tsdf_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--distinct", type=int, default=12)
a = ap.parse_args()
poses, clouds = bench.make_inputs(a.distinct)
vol = pkg.TSDFVolumeOctree(device=0, pool_log2=18)
vol.setGridSize(bench.SIZE, bench.SIZE, bench.SIZE)
vol.setResolution(bench.RES, bench.RES, bench.RES)
vol.setCameraIntrinsics(bench.CAM.fx, bench.CAM.fy, bench.CAM.cx, bench.CAM.cy)
vol.setIntegrateColor(True)
vol.reset()
d = [torch.from_numpy(c).cuda() for c in clouds]
for k in range(a.frames):
    i = k % a.distinct
    vol.integrateCloudDevice(d[i].data_ptr(), bench.H, bench.W, 32, poses[i], rgba_off=16)
vol.sync()
print("frames", a.frames, "updates last frame", vol.stats().n_updates)

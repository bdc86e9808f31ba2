#!/usr/bin/env python3
"""Diagnostics: per-frame phase timing of k_celltop_up on the bench workload (uses the debug hook b200tsdf_debug_timing)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, cpu_tsdf_b200 as pkg
from cpu_tsdf_b200 import synth
CAM = synth.Camera()
vol = pkg.TSDFVolumeOctree(device=0, pool_log2=18)
vol.setResolution(2048, 2048, 2048); vol.setGridSize(10, 10, 10); vol.setCameraIntrinsics(CAM.fx, CAM.fy, CAM.cx, CAM.cy); vol.setIntegrateColor(True); vol.reset()
poses, clouds = bench.make_inputs(64)
lib = vol._lib; lib.b200tsdf_debug_timing.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros(16, np.uint64)
for rep in range(2):
    for i in range(64):
        vol.integrateCloud(clouds[i], None, poses[i])
        if rep == 0 and i < 63:
            continue
        lib.b200tsdf_debug_timing(vol._h, buf.ctypes.data)
        if rep == 0:
            continue
        def f(v):
            v = int(v); return f"{(v >> 32) / 1965:.1f}us[s2={(v >> 24) & 255} s1={(v >> 20) & 15} cell={(v >> 18) & 3} ci={v & 0x3FFFF}]"
        st = vol.stats()
        print(i, "ms", round(st.ms_last_integrate * 1e3, 1), "total", f(buf[0]), "L3", f(buf[1]), "L2", f(buf[2]), "L1", f(buf[3]), "cell", f(buf[4]), "wait", f(buf[5]),
              "slow2/slow1/cellvisit/cellft/cells", [int(x) for x in buf[6:11]])

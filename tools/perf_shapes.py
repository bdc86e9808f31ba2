#!/usr/bin/env python3
"""Device-resident integrate rate of the grid shapes of BASELINE.json's configs (256^3/3 m, 512^3/3 m, 1024^3/3 m, 2048^3/10 m,
4096^3/10 m): which launch sequence each takes and how fast it is.  Prints one JSON line per shape."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpu_tsdf_b200 as pkg
from cpu_tsdf_b200 import synth

CAM = synth.Camera()
SHAPES = [("256/3m", 256, 3.0, synth.S1, 17), ("512/3m", 512, 3.0, synth.S1, 18), ("1024/3m", 1024, 3.0, synth.S1, 19),
          ("2048/10m", 2048, 10.0, synth.S2, 18), ("4096/10m", 4096, 10.0, synth.S2, 20)]
only = sys.argv[1:] or None
for name, res, size, scene, pool in SHAPES:
    if only and name not in only:
        continue
    v = pkg.TSDFVolumeOctree(device=0, pool_log2=pool)
    v.setGridSize(size, size, size); v.setResolution(res, res, res)
    v.setCameraIntrinsics(CAM.fx, CAM.fy, CAM.cx, CAM.cy); v.setIntegrateColor(True); v.reset()
    n = 16
    poses = [synth.orbit_pose(scene, f, 100) for f in range(n)]
    dev = [torch.from_numpy(synth.make_frame(scene, p, CAM, color=True, noise_seed=12345, frame=f)).cuda() for f, p in enumerate(poses)]
    ptrs = [d.data_ptr() for d in dev]
    for rep in range(2):                                   # warm the volume
        v.integrateBatchDevice(ptrs, CAM.height, CAM.width, 32, poses, rgba_off=16)
    v.sync()
    v.profile_begin()
    reps = 6
    for rep in range(reps):
        v.integrateBatchDevice(ptrs, CAM.height, CAM.width, 32, poses, rgba_off=16)
    pr = v.profile_end()
    st = v.stats()
    print(json.dumps({"shape": name, "frames_per_s": round(n * reps / (pr.ms_elapsed / 1e3), 1), "us_per_frame": round(1e3 * pr.ms_elapsed / (n * reps), 1),
                      "launches_per_frame": pr.total_launches / (n * reps), "graph_launches": pr.graph_launches,
                      "updates_per_frame": pr.n_updates // (n * reps), "bricks": st.n_bricks, "blocks_last_frame": st.n_block_visits,
                      "levels": [st.coarse_level, st.finest_level, st.tiers]}))
    del v, dev
    torch.cuda.empty_cache()

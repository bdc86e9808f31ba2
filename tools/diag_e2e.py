#!/usr/bin/env python3
"""Where does the end-to-end leg spend its time?  H2D bandwidth of the box from the same pinned buffers, then host / device time
per integrateBatchRows call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, cpu_tsdf_b200 as pkg

print("affinity:", bench.bind_to_gpu_numa(0) if "--bind" in sys.argv else "not bound")
poses, clouds = bench.make_inputs(32)
H, W = bench.H, bench.W
h_rows = [torch.from_numpy(np.ascontiguousarray(c)).pin_memory() for c in clouds]
d = [torch.empty_like(t, device="cuda") for t in h_rows]
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for a, b in zip(d, h_rows):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"torch H2D 32 x {h_rows[0].numel()*4/1e6:.1f} MB pinned: {32*h_rows[0].numel()*4/dt/1e9:.1f} GB/s")
vol = pkg.TSDFVolumeOctree(device=0, pool_log2=18)
vol.setGridSize(bench.SIZE, bench.SIZE, bench.SIZE); vol.setResolution(bench.RES, bench.RES, bench.RES)
vol.setCameraIntrinsics(bench.CAM.fx, bench.CAM.fy, bench.CAM.cx, bench.CAM.cy); vol.setIntegrateColor(True); vol.reset()
ptrs = [t.data_ptr() for t in h_rows]
for step in range(6):
    t0 = time.perf_counter()
    vol.integrateBatchRows(ptrs, H, W, 32, poses, rgba_off=16)
    t1 = time.perf_counter()
    vol.sync()
    t2 = time.perf_counter()
    n = vol.stats().n_updates
    t3 = time.perf_counter()
    print(f"step {step}: call {1e3*(t1-t0):.2f} ms, sync {1e3*(t2-t1):.2f} ms, stats {1e3*(t3-t2):.2f} ms -> {32/(t3-t0):.0f} frames/s")
# the old per-frame async path for comparison
for step in range(3):
    t0 = time.perf_counter()
    for i in range(32):
        vol.integrateCloudAsync(ptrs[i], H, W, 32, poses[i], rgba_off=16)
    t1 = time.perf_counter()
    vol.sync()
    t2 = time.perf_counter()
    print(f"per-frame async step {step}: calls {1e3*(t1-t0):.2f} ms, sync {1e3*(t2-t1):.2f} ms -> {32/(t2-t0):.0f} frames/s")

#!/usr/bin/env python3
"""Secondary measurements for SURVEY.md §8(f) rows 2-3 (not the BASELINE metric): the z-buffer re-organisation of
unorganised clouds and the mesh post-processing (flattenVertices, cleanupMesh) on the GPU against the sequential CPU
restatement of integrate.cpp's loops (oracle/prog_oracle.cpp — the reference program itself needs PCL/FLANN/boost and
cannot be built here).  Times are whole C-ABI calls from host buffers (copies included).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cpu_tsdf_b200 as pkg  # noqa: E402
from cpu_tsdf_b200 import synth  # noqa: E402
from oracle import oracle_py  # noqa: E402

CAM = synth.Camera()


def best(fn, n=3):
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); t.append(1e3 * (time.perf_counter() - t0))
    return min(t), r


def main():
    out = {}
    rng = np.random.default_rng(0)
    # ---- organise: 4 registered S2 frames' points dumped into one unorganised cloud (1.2 M points, colour) ----------
    pose = synth.orbit_pose(synth.S2, 0, 100)
    fr = synth.make_frame(synth.S2, pose, CAM, color=True).reshape(-1, 8)
    fr = fr[~np.isnan(fr[:, 2])]
    pts = np.concatenate([fr * np.float32([s, s, s, 1, 1, 1, 1, 1]) for s in (1.0, 0.9, 1.1, 1.0)]).astype(np.float32)
    pts = np.ascontiguousarray(pts[rng.permutation(len(pts))])
    vol = pkg.TSDFVolumeOctree(device=0, pool_log2=18)
    vol.setResolution(2048, 2048, 2048); vol.setGridSize(10, 10, 10); vol.setCameraIntrinsics(CAM.fx, CAM.fy, CAM.cx, CAM.cy)
    vol.setIntegrateColor(True); vol.reset()
    vol.organizeCloud(pts)
    tg, (og, nf) = best(lambda: vol.organizeCloud(pts))
    tc, (oc, nf2) = best(lambda: oracle_py.organize(pts, (CAM.fx, CAM.fy, CAM.cx, CAM.cy), CAM.width, CAM.height, rgba_off=16))
    out["organize"] = {"points": int(len(pts)), "filled": nf, "ms_gpu_call": tg, "ms_cpu": tc, "identical": bool(nf == nf2 and np.array_equal(og.view(np.uint32), oc.view(np.uint32)))}
    vol.integrateUnorganizedCloud(pts, pose); vol.sync()
    vol.profile_begin()
    for _ in range(8):
        vol.integrateUnorganizedCloud(pts, pose)
    pr = vol.profile_end()
    out["integrate_unorganized"] = {"ms_per_cloud_stream": pr.ms_elapsed / 8, "h2d_mb_per_cloud": pr.h2d_bytes / 8 / 1e6}
    # ---- mesh post-processing on the 2048^3 interior mesh --------------------------------------------------------
    for f in range(20):
        p = synth.orbit_pose(synth.S2, f * 5, 100)
        vol.integrateCloud(synth.make_frame(synth.S2, p, CAM, color=True, noise_seed=12345, frame=f), None, p)
    mc = pkg.MarchingCubesTSDFOctree(); mc.setInputTSDF(vol); mc.setMinWeight(0.0)
    v, _, _ = mc.reconstruct()
    v = np.asarray(v, np.float32).reshape(-1, 3).copy()
    t = np.arange(len(v), dtype=np.int32).reshape(-1, 3)
    pkg.flattenVertices(v[:3000], t[:1000])
    tg, fg = best(lambda: pkg.flattenVertices(v, t), 2)
    tc, fc = best(lambda: oracle_py.flatten_vertices(v, t), 1)
    out["flattenVertices"] = {"triangles": int(len(t)), "vertices_out": int(len(fg[0])), "ms_gpu_call": tg, "ms_cpu": tc,
                              "identical": bool(fg[0].shape == fc[0].shape and np.array_equal(fg[0].view(np.uint32), fc[0].view(np.uint32)) and np.array_equal(fg[1], fc[1]))}
    tg, cg = best(lambda: pkg.cleanupMesh(*fg), 2)
    tc, cc = best(lambda: oracle_py.cleanup_mesh(*fc), 1)
    out["cleanupMesh"] = {"faces_in": int(len(fg[1])), "faces_out": int(len(cg[1])), "ms_gpu_call": tg, "ms_cpu": tc,
                          "identical": bool(cg[0].shape == cc[0].shape and np.array_equal(cg[0].view(np.uint32), cc[0].view(np.uint32)) and np.array_equal(cg[1], cc[1]))}
    out["cpu_kind"] = "port (sequential restatement, 1 thread)"
    print(json.dumps(out))


if __name__ == "__main__":
    main()

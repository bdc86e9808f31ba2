#!/usr/bin/env python3
"""The BASELINE.json configs that are not bench.py's headline line (configs[2]): one JSON line per run.

  --config 0   single 640x480 frame into 256^3 / 3 m (S1) + getFxn spot-check at 1000 points
  --config 1   100-frame orbit into 512^3 / 3 m (S1): integrate rate + renderView at frames 0/25/50/75
  --config 3   (torchrun, N ranks) 2048^3 / 10 m sharded by coarse cell, 1000 frames from HOST row slices (NVLink all-gather)
  --config 4   (torchrun, N ranks) 4096^3 / 10 m sharded: integrate -> gather + renderView every 10th frame -> gather + mesh
  --exchange   (torchrun, N ranks) latency of a per-frame all-to-all of active block keys (NCCL) next to the zero-communication
               front end it would replace (SURVEY.md §8e "measure both")
Launch: python tools/bench_configs.py --config 0   |   python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py --config 3
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import cpu_tsdf_b200 as pkg
from cpu_tsdf_b200 import synth

CAM = synth.Camera(); W, H = CAM.width, CAM.height


def make_vol(res, size, device, pool, color=True, max_cell=None, **kw):
    v = pkg.TSDFVolumeOctree(device=device, pool_log2=pool, **kw)
    v.setGridSize(size, size, size); v.setResolution(res, res, res)
    v.setCameraIntrinsics(CAM.fx, CAM.fy, CAM.cx, CAM.cy); v.setIntegrateColor(color)
    if max_cell:
        v.setMaxVoxelSize(max_cell, max_cell, max_cell)
    v.reset()
    return v


def ev_ms(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


def config0():
    v = make_vol(256, 3.0, 0, 17, color=False)
    pose = synth.orbit_pose(synth.S1, 0, 100); cloud = synth.make_frame(synth.S1, pose, CAM)
    t0 = time.perf_counter(); v.integrateCloud(cloud, None, pose); v.sync(); t_first = time.perf_counter() - t0
    st = v.stats()
    nodes = v.download_nodes()
    leaf = (~nodes["split"].astype(bool)) & (nodes["keys"][:, 0] == 8)
    ctr = (nodes["keys"][leaf][:, 1:] + 0.5) * (3.0 / 256) - 1.5
    rng = np.random.default_rng(1)
    pts = (ctr[rng.integers(0, len(ctr), 1000)] + rng.uniform(-0.5, 0.5, (1000, 3)) * (3.0 / 256)).astype(np.float32)
    t0 = time.perf_counter(); ok, val = v.getFxn(pts); t_q = time.perf_counter() - t0
    line = {"config": 0, "workload": "single 640x480 frame (S1) into 256^3 / 3 m + getFxn at 1000 points near finest leaves",
            "integrate_ms_first_call_host_to_volume": 1e3 * t_first, "updates": int(st.n_updates), "nodes": int(len(nodes["keys"])),
            "getFxn_in_bounds": int(ok.sum()), "getFxn_ms_1000_points": 1e3 * t_q}
    try:
        from oracle.oracle_py import OracleVolume
        o = OracleVolume(xres=256, yres=256, zres=256, cx=CAM.cx, cy=CAM.cy); o.reset()
        t0 = time.perf_counter(); o.integrate(cloud, pose); line["cpu_oracle_integrate_ms"] = 1e3 * (time.perf_counter() - t0)
        v2, _, _, ok2 = o.query(pts, 1, 0)
        line["getFxn_bit_identical_to_oracle"] = bool(np.array_equal(ok, ok2) and np.array_equal(val[ok], v2[ok2]))
        line["nodes_match_oracle"] = bool(len(o.dump_nodes()["keys"]) == len(nodes["keys"]))
    except Exception as e:
        line["cpu_oracle"] = f"unavailable ({type(e).__name__})"
    print(json.dumps(line))


def config1():
    v = make_vol(512, 3.0, 0, 20, color=False)
    poses = [synth.orbit_pose(synth.S1, f, 100) for f in range(100)]
    dev = [torch.from_numpy(synth.make_frame(synth.S1, p, CAM, noise_seed=2, frame=f)).cuda() for f, p in enumerate(poses)]
    ptrs = [d.data_ptr() for d in dev]
    v.profile_begin()
    for lo in range(0, 100, 25):
        v.integrateBatchDevice(ptrs[lo:lo + 25], H, W, 16, poses[lo:lo + 25])
    pr = v.profile_end()
    renders = {}
    for f in (0, 25, 50, 75):
        v.renderView(poses[f], 1)
        t0 = time.perf_counter(); r = v.renderView(poses[f], 1); renders[f] = {"ms": 1e3 * (time.perf_counter() - t0), "hits": int(np.isfinite(r[..., 2]).sum())}
    print(json.dumps({"config": 1, "workload": "100-frame S1 orbit into 512^3 / 3 m, device-resident clouds, 4 graph launches of 25 frames; renderView 640x480 at frames 0/25/50/75 (host call incl. D2H)",
                      "integrate_frames_per_s": 100 / (pr.ms_elapsed / 1e3), "updates_per_frame": pr.n_updates // 100, "renderView": renders,
                      "parity": "tests/test_golden.py::test_engine_reproduces_reference_digests_at_baseline_lengths[L1_512_orbit100]"}))


def dist_setup():
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist, rank, world, local


def comm(v, dist, rank, world):
    if world > 1:
        ids = [pkg.TSDFVolumeOctree.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        v.commInit(ids[0], rank, world)


def sharded_stream(res, size, pool, nframes, total_orbit, max_cell, dist, rank, world, local, render_every=0, mesh=False):
    v = make_vol(res, size, local, pool, max_cell=max_cell, shard_rank=rank, shard_count=world)
    comm(v, dist, rank, world)
    r0, r1 = v.rowSlice(H)
    ND = 64
    poses = [synth.orbit_pose(synth.S2, f, total_orbit) for f in range(nframes)]
    # 64 distinct clouds (host memory), re-used along the longer orbit with their own poses' depth maps regenerated per 64
    rows, pcache = [], {}
    for f in range(min(ND, nframes)):
        c = synth.make_frame(synth.S2, poses[f], CAM, color=True, noise_seed=12345, frame=f)
        rows.append(torch.from_numpy(np.ascontiguousarray(c[r0:r1])).pin_memory())
    full = make_vol(res, size, local, pool, max_cell=max_cell) if (rank == 0 and (render_every or mesh)) else None
    t_render, t_gather, n_render, hits = 0.0, 0.0, 0, 0
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    # one untimed batch first: NCCL channel set-up, graph capture and the staging buffers are one-time costs
    v.integrateBatchRows([rows[i].data_ptr() for i in range(min(8, len(rows)))], H, W, 32, poses[:min(8, len(rows))], rgba_off=16)
    v.sync()
    if world > 1 and (render_every or mesh):
        if full is not None:
            full.reset()
        v.gatherVolume(full, 0)
    barrier()
    t0 = time.perf_counter()
    v.profile_begin()
    f = 0
    while f < nframes:
        m = min(10 if render_every else 32, nframes - f)
        idx = [(f + j) % len(rows) for j in range(m)]
        # (a cloud is re-used with the pose it was rendered from: frame f + j uses pose of its cloud index when the stream is longer than 64)
        v.integrateBatchRows([rows[i].data_ptr() for i in idx], H, W, 32, [poses[i if nframes > len(rows) else f + j] for j, i in enumerate(idx)], rgba_off=16)
        f += m
        if render_every and f % render_every == 0:
            v.sync()
            tg = time.perf_counter()
            if full is not None:
                full.reset()
            if world > 1:
                v.gatherVolume(full, 0)
            t_gather += time.perf_counter() - tg
            if rank == 0:
                tr = time.perf_counter()
                r = (full if world > 1 else v).renderView(poses[idx[-1] if nframes > len(rows) else f - 1], 1)     # the view just integrated
                t_render += time.perf_counter() - tr; n_render += 1; hits = int(np.isfinite(r[..., 2]).sum())
    v.sync()
    pr = v.profile_end()
    barrier()
    t_int = time.perf_counter() - t0
    out = {"n_gpus": world, "frames": nframes, "wall_s": t_int, "frames_per_s_whole_pipeline": nframes / t_int,
           "h2d_bytes_per_frame_per_rank": pr.h2d_bytes // nframes, "nvlink_bytes_per_frame_per_rank": pr.nvlink_bytes // nframes,
           "updates_per_frame_this_rank": pr.n_updates // nframes, "rows_per_rank": [r0, r1]}
    if render_every:
        out.update({"renders": n_render, "render_ms_each": 1e3 * t_render / max(1, n_render), "gather_ms_each": 1e3 * t_gather / max(1, n_render), "render_hits_last": hits})
    if mesh:
        tg = time.perf_counter()
        if full is not None:
            full.reset()
        if world > 1:
            v.gatherVolume(full, 0)
        if rank == 0:
            mc = pkg.MarchingCubesTSDFOctree(); mc.setInputTSDF(full if world > 1 else v); mc.setMinWeight(2.0); mc.setColorByRGB(True)
            tm = time.perf_counter(); verts, cols, polys = mc.reconstruct(); out["mesh_s"] = time.perf_counter() - tm; out["mesh_triangles"] = len(polys)
        out["final_gather_plus_mesh_s"] = time.perf_counter() - tg
        barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=-1)
    ap.add_argument("--exchange", action="store_true")
    ap.add_argument("--frames", type=int, default=0)
    a = ap.parse_args()
    if a.config == 0:
        return config0()
    if a.config == 1:
        return config1()
    dist, rank, world, local = dist_setup()
    if a.config == 3:
        n = a.frames or 1000
        out = sharded_stream(2048, 10.0, 18, n, 1000, None, dist, rank, world, local)
        out.update({"config": 3, "workload": f"2048^3 / 10 m sharded by coarse cell over {world} GPU(s), {n} frames (the first 64 views of the S2 stream with theta = 2 pi f / 1000, cycled) from pinned HOST row slices: "
                    "H2D of 1/N of every frame per rank, NCCL all-gather over NVLink inside the library, graph launches of 8 frames"})
    elif a.config == 4:
        n = a.frames or 200
        out = sharded_stream(4096, 10.0, 20, n, 1000, 0.3, dist, rank, world, local, render_every=10, mesh=True)
        out.update({"config": 4, "workload": f"4096^3 / 10 m (2.44 mm voxels; setMaxVoxelSize(0.3) so that the coarse cells are the tier-1 roots and shards gather device to device) over {world} GPU(s): "
                    f"{n} frames integrated from host row slices, every 10th frame all shards are gathered over NVLink into a full volume on rank 0 and renderView (640x480) runs there, marching cubes (w_min 2, rgb) on the final gathered volume"})
    elif a.exchange:
        # what a per-frame exchange of newly active block keys would cost: all_to_all_single of 8-byte keys, ~20 k keys per frame in total
        keys_total = 20000
        per = max(1, keys_total // (world * world))
        send = torch.zeros(world * per, dtype=torch.int64, device="cuda"); recv = torch.empty_like(send)
        times = []
        if world > 1:
            for it in range(120):
                times.append(ev_ms(lambda: dist.all_to_all_single(recv, send)))
        t = float(np.median(times[20:])) if times else 0.0
        out = {"exchange": True, "n_gpus": world, "keys_per_frame_total": keys_total, "bytes_per_rank_pair": per * 8,
               "nccl_all_to_all_us_median": 1e3 * t,
               "zero_comm_front_end": "k_front scans all 307 200 pixels on every rank and keeps the keys of the cells it owns: 12-13 us per frame at any N (profiles/r2_launches.csv); "
                                      "a pixel-sliced front end would save at most (1 - 1/N) of that and pay the all-to-all above plus a second launch on every frame"}
    else:
        raise SystemExit("nothing to do")
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()

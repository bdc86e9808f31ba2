"""Worker for tests/test_multi_gpu.py (one process per GPU, launched by torch.distributed.run on a box with >= 2 GPUs).

Every rank owns the coarse cells hash(cell) % world == rank of a 512^3 volume and holds in HOST memory only its row slice of
each frame.  The frames go in through b200tsdf_integrate_batch_rows (slice upload over the rank's own PCIe link, NCCL
all-gather over NVLink inside the library, one graph launch per batch); the shards are then gathered device to device into
a full volume on rank 0 (b200tsdf_gather_volume), which must equal the single-volume CPU oracle bit for bit — nodes,
renderView and the marching-cubes soup."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cpu_tsdf_b200 as pkg  # noqa: E402
from cpu_tsdf_b200 import synth  # noqa: E402
from tests.common import CAM, CFG_512, assert_same_nodes, frames  # noqa: E402


def make(cfg, device, **kw):
    v = pkg.TSDFVolumeOctree(device=device, pool_log2=18, **kw)
    v.setResolution(cfg["xres"], cfg["yres"], cfg["zres"])
    v.setGridSize(cfg.get("xsize", 3.0), cfg.get("ysize", 3.0), cfg.get("zsize", 3.0))
    v.setCameraIntrinsics(525.0, 525.0, cfg["cx"], cfg["cy"])
    v.setIntegrateColor(True)
    v.reset()
    return v


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [pkg.TSDFVolumeOctree.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    shard = make(CFG_512, local, shard_rank=rank, shard_count=world)
    shard.commInit(ids[0], rank, world)
    fs = list(frames(synth.S1, 11, stride=8, color=True, noise_seed=31))
    H, W = fs[0][1].shape[:2]
    r0, r1 = shard.rowSlice(H)
    assert (r1 - r0) * world >= H
    # this rank's host memory holds ONLY its rows of every frame (pinned)
    mine = [torch.from_numpy(np.ascontiguousarray(c[r0:r1])).pin_memory() for _, c in fs]
    poses = [p for p, _ in fs]
    for lo, hi in ((0, 4), (4, 8), (8, 11)):
        shard.integrateBatchRows([t.data_ptr() for t in mine[lo:hi]], H, W, 32, poses[lo:hi], rgba_off=16)
    shard.sync()
    full = make(CFG_512, local) if rank == 0 else None
    shard.gatherVolume(full, 0)
    if rank == 0:
        from oracle.oracle_py import OracleVolume
        o = OracleVolume(**CFG_512, integrate_color=1); o.reset()
        for pose, cloud in fs:
            o.integrate(cloud, pose)
        assert_same_nodes(o.dump_nodes(), full.download_nodes(), rgb=True)
        pose = synth.orbit_pose(synth.S1, 30, 100)
        ra, rb = o.render(pose, 2), full.renderView(pose, 2)
        assert np.array_equal(ra[..., :3], rb[..., :3], equal_nan=True) and np.array_equal(ra[..., 4:7], rb[..., 4:7], equal_nan=True)
        mc = pkg.MarchingCubesTSDFOctree(); mc.setInputTSDF(full); mc.setMinWeight(2.0); mc.setColorByRGB(True)
        vb, cb, _ = mc.reconstruct()
        va, ca = o.mesh(2.0, 1)
        assert len(va) == len(vb) > 1000
        assert np.array_equal(np.asarray(va).view(np.uint32), np.asarray(vb).reshape(-1, 3).view(np.uint32)) and np.array_equal(ca, cb)
        print(f"MGPU_OK world={world} nodes={len(full.download_nodes()['keys'])} tris={len(vb) // 3}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Worker for tests/test_shard_gloo.py: rank r fuses the same frames into the shard of the volume
it owns (coarse cells with hash(cell) % world == rank); rank 0 checks that the union of the shards
is bit-identical to the unsharded oracle volume.  Runs on CPU (gloo) through the host emulation of
the engine's device code."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cpu_tsdf_b200 import synth  # noqa: E402
from tests.common import CAM, CFG_256, frames  # noqa: E402
from tests.emu.emu_py import EmuVolume  # noqa: E402


def cell_owner(cx, cy, cz, world):
    m = (1 << 64) - 1
    k = ((int(cx) << 40) | (int(cy) << 20) | int(cz)) & m
    k ^= k >> 33; k = (k * 0xff51afd7ed558ccd) & m
    k ^= k >> 33; k = (k * 0xc4ceb9fe1a85ec53) & m
    k ^= k >> 33
    return (k & 0xFFFFFFFF) % world


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    e = EmuVolume(shard_rank=rank, shard_count=world, **CFG_256)
    e.reset()
    fr = list(frames(synth.S1, 3, stride=11, noise_seed=21))
    upd = 0
    for pose, cloud in fr:
        e.integrate(cloud, pose)
        upd += e.stats()["n_updates"]
    mine = e.dump_nodes()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine["keys"], mine["dw"], mine["split"], upd))
    if rank == 0:
        from oracle.oracle_py import OracleVolume
        o = OracleVolume(**CFG_256); o.reset()
        total = 0
        for pose, cloud in fr:
            o.integrate(cloud, pose)
            total += o.stats().n_add_observation
        ref = o.dump_nodes()
        C = o.levels()[0]
        cells = ref["keys"][:, 1:] >> (ref["keys"][:, 0:1] - C)
        owner = np.array([cell_owner(*c, world) for c in cells])
        n_checked = 0
        for r in range(world):
            keys, dw, split, _ = gathered[r]
            rc = keys[:, 1:] >> (keys[:, 0:1] - C)
            ro = np.array([cell_owner(*c, world) for c in rc])
            own = ro == r
            # owned subtrees are identical to the oracle's
            sel = owner == r
            assert np.array_equal(keys[own], ref["keys"][sel]), f"rank {r}: structure differs"
            assert np.array_equal(dw[own].view(np.uint32), ref["dw"][sel].view(np.uint32)), f"rank {r}: sdf/weight differ"
            assert np.array_equal(split[own], ref["split"][sel])
            # cells it does not own stay pristine
            assert (keys[~own][:, 0] == C).all() and (dw[~own] == [-1, 0]).all() and not split[~own].any()
            n_checked += int(own.sum())
        assert n_checked == len(ref["keys"])
        assert sum(g[3] for g in gathered) == total        # every voxel update happened on exactly one rank
        print(f"SHARD_OK world={world} nodes={n_checked} updates={total}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

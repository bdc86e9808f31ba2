"""Unorganised clouds: the z-buffer re-organisation of the reference's `integrate`
program (src/prog/integrate.cpp:548-607).  CPU: the restatement (oracle/prog_oracle.cpp) against the host
emulation of the device code (organize.cuh driven in reverse order); GPU: the CUDA kernels against the
restatement, and unorganised integration against organised integration of the oracle."""
import numpy as np
import pytest

from cpu_tsdf_b200 import synth
from oracle import oracle_py
from tests.common import CAM, CFG_256, assert_same_nodes

INTR = (525.0, 525.0, CAM.cx, CAM.cy)


def unorganized_cloud(seed, *, n_extra=60000, color=True, scale=1.0, world_pose=None):
    """A shuffled S1 frame + a second, nearer/farther layer hitting the same pixels (z-fights, exact z ties),
    + rejects: NaN, z <= 0, outside the image, zeros, infinities."""
    rng = np.random.default_rng(seed)
    pose = synth.orbit_pose(synth.S1, 3, 100)
    fr = synth.make_frame(synth.S1, pose, CAM, color=color).reshape(-1, 8 if color else 4)
    pts = fr[~np.isnan(fr[:, 2])].copy()
    layer = pts[rng.integers(0, len(pts), n_extra)].copy()
    s = rng.choice(np.float32([0.5, 0.75, 1.0, 1.0, 1.25]), n_extra)[:, None]     # 1.0: exact duplicates (tie -> first wins)
    layer[:, :3] *= s
    if color:
        layer[:, 4] = rng.integers(0, 256, (n_extra, 4), dtype=np.uint8).view(np.float32)[:, 0]
    junk = np.zeros((5000, pts.shape[1]), np.float32)
    junk[:, :3] = rng.normal(scale=2.0, size=(5000, 3))                           # behind the camera / outside the image
    junk[:500, 2] = np.nan; junk[500:1000, 0] = np.nan; junk[1000:1500, :3] = 0.0
    junk[1500:1600, 2] = np.inf; junk[1600:1700, 0] = np.inf; junk[1700:1800, 2] = -0.0
    junk[1800:1900, 2] = 1e-38                                                    # projected coordinate overflows int
    allp = np.concatenate([pts, layer, junk]).astype(np.float32)
    allp = allp[rng.permutation(len(allp))]
    if world_pose is not None:                                                    # express in a world frame
        R, t = world_pose[:3, :3], world_pose[:3, 3]
        with np.errstate(invalid="ignore"):
            allp[:, :3] = (allp[:, :3].astype(np.float64) @ R.T + t).astype(np.float32)
    allp[:, :3] /= np.float32(scale)
    return np.ascontiguousarray(allp), pose


def same_cloud(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("color", [False, True])
def test_emulated_device_code_matches_the_restatement(color):
    from tests.emu import emu_py
    pts, _ = unorganized_cloud(1, color=color)
    ro = 16 if color else -1
    want, nf = oracle_py.organize(pts, INTR, CAM.width, CAM.height, rgba_off=ro)
    got, nf2 = emu_py.organize(pts, INTR, CAM.width, CAM.height, rgba_off=ro)
    assert nf == nf2 and nf > 200000
    assert same_cloud(want, got)
    # 16-byte pixels (the layout the fused integrate path keeps in HBM): same xyz and colour bytes
    got16, _ = emu_py.organize(pts, INTR, CAM.width, CAM.height, rgba_off=ro, out_stride=16, out_rgba_off=12)
    assert same_cloud(want[..., :3], got16[..., :3]) and same_cloud(want[..., 4], got16[..., 3])


def numpy_organize(pts, intr, W, H, cloud_units=1.0, zero_nans=False):
    """An independent, vectorised statement of integrate.cpp:548-607 (no world transform): per pixel the accepted point with
    the smallest (z, input index)."""
    fx, fy, cx, cy = np.float32(intr)
    xyz = pts[:, :3].astype(np.float32).copy()
    if cloud_units != 1.0:
        xyz *= np.float32(cloud_units)
    if zero_nans:
        xyz[(xyz == 0).all(1)] = np.nan
    x, y, z = xyz.T
    with np.errstate(all="ignore"):
        fu = (x * fx / z + cx).astype(np.float32); fv = (y * fy / z + cy).astype(np.float32)
        okf = np.isfinite(fu) & np.isfinite(fv) & (np.abs(fu) < 2e9) & (np.abs(fv) < 2e9)
        u = np.where(okf, np.trunc(np.where(okf, fu, 0)), -1).astype(np.int64); v = np.where(okf, np.trunc(np.where(okf, fv, 0)), -1).astype(np.int64)
        ok = okf & ~np.isnan(z) & (z > 0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
    idx = np.nonzero(ok)[0]
    pix = v[idx] * W + u[idx]
    order = np.lexsort((idx, z[idx], pix))                       # by pixel, then z, then input order
    first = np.ones(len(order), bool); first[1:] = pix[order][1:] != pix[order][:-1]
    win = idx[order][first]; wpix = pix[order][first]
    out = np.zeros((H * W, 8), np.float32); out[:, 2] = np.nan; out[:, 3] = 1.0
    out[:, 4] = np.array([0, 0, 0, 255], np.uint8).view(np.float32)[0]
    out[wpix, :3] = xyz[win]
    out[wpix, 4] = pts[win, 4]
    return out.reshape(H, W, 8), len(win)


def test_restatement_agrees_with_an_independent_vectorised_statement():
    for seed, kw, gen in [(1, {}, {}), (4, dict(cloud_units=0.001, zero_nans=True), dict(scale=0.001))]:
        pts, _ = unorganized_cloud(seed, **gen)
        a, na = oracle_py.organize(pts, INTR, CAM.width, CAM.height, rgba_off=16, **kw)
        b, nb = numpy_organize(pts, INTR, CAM.width, CAM.height, **kw)
        assert na == nb and same_cloud(a, b)


def test_units_zero_nans_and_world_frame_options():
    from tests.emu import emu_py
    world = synth.orbit_pose(synth.S1, 17, 100)
    pts, _ = unorganized_cloud(2, scale=0.001, world_pose=world)                  # millimetres, world frame
    w2c = np.linalg.inv(world)
    kw = dict(rgba_off=16, cloud_units=0.001, zero_nans=True, world_to_camera=w2c)
    want, nf = oracle_py.organize(pts, INTR, CAM.width, CAM.height, **kw)
    got, nf2 = emu_py.organize(pts, INTR, CAM.width, CAM.height, **kw)
    assert nf == nf2 and nf > 100000 and same_cloud(want, got)


def test_empty_cloud_gives_an_all_nan_image():
    from tests.emu import emu_py
    for mod in (oracle_py, emu_py):
        out, nf = mod.organize(np.zeros((0, 4), np.float32), INTR, 64, 48)
        assert nf == 0 and np.isnan(out[..., 2]).all() and not out[..., :2].any()
        assert (np.ascontiguousarray(out[..., 4]).view(np.uint8).reshape(48, 64, 4) == [0, 0, 0, 255]).all()


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_organize_matches_the_restatement():
    import cpu_tsdf_b200 as pkg
    v = pkg.TSDFVolumeOctree(device=0, pool_log2=12)
    v.setResolution(256, 256, 256); v.setGridSize(3, 3, 3); v.setCameraIntrinsics(525, 525, CAM.cx, CAM.cy); v.reset()
    world = synth.orbit_pose(synth.S1, 17, 100)
    for seed, kw, gen in [(1, {}, {}), (3, dict(cloud_units=0.001, zero_nans=True, world_to_camera=np.linalg.inv(world)),
                                        dict(scale=0.001, world_pose=world))]:
        pts, _ = unorganized_cloud(seed, **gen)
        want, nf = oracle_py.organize(pts, INTR, CAM.width, CAM.height, rgba_off=16, **kw)
        got, nf2 = v.organizeCloud(pts, **kw)
        assert nf == nf2 and same_cloud(want, got)
    out, nf = v.organizeCloud(np.zeros((0, 4), np.float32))
    assert nf == 0 and np.isnan(out[..., 2]).all()


@pytest.mark.gpu
def test_unorganized_integration_equals_organized_integration_of_the_oracle():
    import cpu_tsdf_b200 as pkg
    o = oracle_py.OracleVolume(**CFG_256, integrate_color=1); o.reset()
    v = pkg.TSDFVolumeOctree(device=0)
    v.setResolution(256, 256, 256); v.setGridSize(3, 3, 3); v.setCameraIntrinsics(525, 525, CAM.cx, CAM.cy)
    v.setIntegrateColor(True); v.reset()
    for f in range(3):
        pts, _ = unorganized_cloud(10 + f)
        pose = synth.orbit_pose(synth.S1, 3 + 5 * f, 100)
        org, _ = oracle_py.organize(pts, INTR, CAM.width, CAM.height, rgba_off=16)
        o.integrate(org, pose)
        v.integrateUnorganizedCloud(pts, pose)
    assert_same_nodes(o.dump_nodes(), v.download_nodes(), rgb=True)

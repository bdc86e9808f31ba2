"""Mesh post-processing of the reference's `integrate` program: flattenVertices
(src/prog/integrate.cpp:103-150) and cleanupMesh (:152-214).  CPU: the sequential restatement
(oracle/prog_oracle.cpp) against the host emulation of the device formulation (meshpost_core.cuh, loops run
backwards); GPU: the CUDA kernels against the restatement."""
import numpy as np
import pytest

from cpu_tsdf_b200 import synth
from oracle import oracle_py
from tests.common import CAM, CFG_256


def mc_soup():
    """marching-cubes soup of a one-frame 256^3 volume (oracle), as an indexed mesh with tris = 0,1,2,..."""
    o = oracle_py.OracleVolume(**CFG_256); o.reset()
    pose = synth.orbit_pose(synth.S1, 0, 100)
    o.integrate(synth.make_frame(synth.S1, pose, CAM), pose)
    verts, _ = o.mesh(0.0, 0)
    verts = np.asarray(verts, np.float32).reshape(-1, 3)
    return verts, np.arange(len(verts), dtype=np.int32).reshape(-1, 3)


def adversarial_mesh(seed):
    """Chains of vertices 0.6 r apart (the weld relation is not transitive: the sweep order decides), exact
    duplicates, and triangles over them, in shuffled index order."""
    rng = np.random.default_rng(seed)
    r = 1e-4
    v = []
    for c in range(300):
        base = rng.uniform(-1, 1, 3)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        n = int(rng.integers(1, 9))
        for k in range(n):
            v.append(base + d * (0.6 * r * k))
        if c % 7 == 0:
            v.append(base.copy())                                                  # exact duplicate
    v = np.float32(v)
    v = v[rng.permutation(len(v))]
    t = rng.integers(0, len(v), (4000, 3)).astype(np.int32)
    return v, t


def scattered_faces(seed, verts0, tris0):
    """The MC soup + little islands of 1..8 faces (some near the surface, some far away)."""
    rng = np.random.default_rng(seed)
    vs, ts = [verts0], [tris0]
    base = len(verts0)
    for c in range(120):
        centre = rng.uniform(-1.4, 1.4, 3) if c % 3 else verts0[rng.integers(0, len(verts0))] + rng.normal(scale=0.03, size=3)
        for k in range(int(rng.integers(1, 9))):
            tri = (centre + rng.normal(scale=0.004, size=(3, 3))).astype(np.float32)
            vs.append(tri); ts.append(np.arange(base, base + 3, dtype=np.int32)[None]); base += 3
    v = np.concatenate(vs).astype(np.float32); t = np.concatenate(ts).astype(np.int32)
    return v, t[rng.permutation(len(t))]


def same(a, b):
    return a[0].shape == b[0].shape and a[1].shape == b[1].shape and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])


@pytest.fixture(scope="module")
def soup():
    return mc_soup()


def test_flatten_emulation_matches_the_sequential_sweep(soup):
    from tests.emu import emu_py
    v, t = soup
    want = oracle_py.flatten_vertices(v, t)
    got = emu_py.flatten_vertices(v, t)
    assert same(want, got)
    # a closed-form sanity check of the restatement itself: every soup vertex sits on a cube edge that up to four
    # cubes share, so welding leaves ~ nverts/6..nverts/3 vertices, and no face of a surface mesh is lost
    assert len(v) / 7 < len(want[0]) < len(v) / 2 and len(want[1]) > 0.98 * len(t)
    # welded mesh: every face index in range, no degenerate face
    tt = want[1]
    assert tt.min() >= 0 and tt.max() < len(want[0]) and not ((tt[:, 0] == tt[:, 1]) | (tt[:, 1] == tt[:, 2]) | (tt[:, 2] == tt[:, 0])).any()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flatten_order_dependent_chains(seed):
    from tests.emu import emu_py
    v, t = adversarial_mesh(seed)
    want = oracle_py.flatten_vertices(v, t)
    got = emu_py.flatten_vertices(v, t, return_rounds=True)
    assert same(want, got[:2]) and got[2] >= 2
    assert len(want[0]) < len(v)


def test_cleanup_emulation_matches_cluster_extraction(soup):
    from tests.emu import emu_py
    v, t = scattered_faces(5, *soup)
    want = oracle_py.cleanup_mesh(v, t)
    got = emu_py.cleanup_mesh(v, t)
    assert same(want, got)
    assert len(t) - len(want[1]) > 50 and len(want[1]) >= len(soup[1]) * 0.99           # islands go, the surface stays
    for k in (1, 3, 16):
        assert same(oracle_py.cleanup_mesh(v, t, 0.02, k), emu_py.cleanup_mesh(v, t, 0.02, k))
    assert same(oracle_py.cleanup_mesh(v, t, 0.005, 5), emu_py.cleanup_mesh(v, t, 0.005, 5))


def test_flatten_then_cleanup_and_empty_meshes(soup):
    from tests.emu import emu_py
    v, t = scattered_faces(6, *soup)
    fo, fe = oracle_py.flatten_vertices(v, t), emu_py.flatten_vertices(v, t)
    assert same(fo, fe)
    assert same(oracle_py.cleanup_mesh(*fo), emu_py.cleanup_mesh(*fe))
    e = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
    for mod in (oracle_py, emu_py):
        assert mod.flatten_vertices(*e)[0].shape == (0, 3) and mod.cleanup_mesh(*e)[1].shape == (0, 3)
        # a zero or negative radius finds no neighbours: nothing is welded (exact duplicates included), and every face is its
        # own cluster of one, so cleanupMesh removes them all
        for r in (0.0, -1.0):
            fv = mod.flatten_vertices(v[:3000], t[:1000] % 3000, r)
            assert len(fv[0]) == 3000
            assert len(mod.cleanup_mesh(v[:300], np.arange(300, dtype=np.int32).reshape(-1, 3), r)[1]) == 0


def test_post_processing_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import cpu_tsdf_b200 as pkg
    v = np.zeros((3, 3), np.float32); t = np.arange(3, dtype=np.int32).reshape(1, 3)
    for fn in (pkg.flattenVertices, pkg.cleanupMesh):
        with pytest.raises(pkg.B200Error, match="no such CUDA device"):
            fn(v, t)


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_flatten_and_cleanup_match_the_restatement(soup):
    import cpu_tsdf_b200 as pkg
    for v, t in (soup, adversarial_mesh(1), adversarial_mesh(4), scattered_faces(5, *soup)):
        want = oracle_py.flatten_vertices(v, t)
        got = pkg.flattenVertices(v, t)
        assert same(want, got)
        assert same(oracle_py.cleanup_mesh(*want), pkg.cleanupMesh(*got))
    v, t = scattered_faces(7, *soup)
    for k, d in ((1, 0.02), (3, 0.02), (16, 0.02), (5, 0.005)):
        assert same(oracle_py.cleanup_mesh(v, t, d, k), pkg.cleanupMesh(v, t, d, k))
    e = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
    assert pkg.flattenVertices(*e)[0].shape == (0, 3) and pkg.cleanupMesh(*e)[1].shape == (0, 3)
    with pytest.raises(pkg.B200Error):
        pkg.cleanupMesh(v, t, 0.02, 17)
    with pytest.raises(pkg.B200Error):
        pkg.flattenVertices(v[:10], t)                                              # index out of range

"""Property tests (hypothesis) for the program-side rows: random small inputs through the sequential restatement
(oracle/prog_oracle.cpp) and the host emulation of the device formulation must agree bit for bit."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle_py
from tests.emu import emu_py

F32 = st.floats(min_value=-4.0, max_value=4.0, allow_nan=False, width=32)


@st.composite
def clouds(draw):
    n = draw(st.integers(0, 200))
    rng = np.random.default_rng(draw(st.integers(0, 2**32 - 1)))
    p = np.zeros((n, 8), np.float32)
    # few distinct depths and coarse x/y -> many pixel collisions and exact z ties
    p[:, 2] = rng.choice(np.float32([0.5, 1.0, 1.0, 2.0, -1.0, 0.0, np.nan, np.inf]), n)
    p[:, 0] = rng.integers(-3, 4, n).astype(np.float32) * np.float32(0.01) * np.abs(np.nan_to_num(p[:, 2], nan=1.0, posinf=1.0))
    p[:, 1] = rng.integers(-3, 4, n).astype(np.float32) * np.float32(0.01) * np.abs(np.nan_to_num(p[:, 2], nan=1.0, posinf=1.0))
    p[:, 4] = rng.integers(0, 2**31, n).astype(np.uint32).view(np.float32)
    return p


@settings(max_examples=60, deadline=None)
@given(clouds(), st.sampled_from([1.0, 0.001, 2.5]), st.booleans(), st.booleans())
def test_organize_random_clouds(cloud, units, zero_nans, world):
    intr = (30.0, 30.0, 7.5, 5.5)
    w2c = None
    if world:
        a = 0.3
        w2c = np.array([[np.cos(a), 0, np.sin(a), 0.01], [0, 1, 0, -0.02], [-np.sin(a), 0, np.cos(a), 0.03], [0, 0, 0, 1]])
    kw = dict(rgba_off=16, cloud_units=units, zero_nans=zero_nans, world_to_camera=w2c)
    a, na = oracle_py.organize(cloud, intr, 16, 12, **kw)
    b, nb = emu_py.organize(cloud, intr, 16, 12, **kw)
    assert na == nb and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@st.composite
def meshes(draw):
    rng = np.random.default_rng(draw(st.integers(0, 2**32 - 1)))
    nc = draw(st.integers(1, 12))
    pts = []
    for _ in range(nc):                                   # clusters of vertices a fraction of the weld radius apart
        c = rng.uniform(-0.05, 0.05, 3)
        k = int(rng.integers(1, 7))
        step = rng.choice([0.0, 3e-5, 6e-5, 9e-5])
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        pts += [c + d * step * i for i in range(k)]
    v = np.float32(pts)
    v = v[rng.permutation(len(v))]
    nt = draw(st.integers(0, 40))
    t = rng.integers(0, len(v), (nt, 3)).astype(np.int32)
    return v, t


@settings(max_examples=80, deadline=None)
@given(meshes(), st.sampled_from([1e-4, 5e-5, 0.0]))
def test_flatten_random_meshes(mesh, r):
    a = oracle_py.flatten_vertices(*mesh, r); b = emu_py.flatten_vertices(*mesh, r)
    assert a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])


@settings(max_examples=80, deadline=None)
@given(meshes(), st.sampled_from([0.02, 0.005, 0.08]), st.integers(1, 16))
def test_cleanup_random_meshes(mesh, d, k):
    v, t = mesh
    a = oracle_py.cleanup_mesh(v, t, d, k); b = emu_py.cleanup_mesh(v, t, d, k)
    assert a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])

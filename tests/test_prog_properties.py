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


# ---- independent brute-force statements (no search structure at all) ------------------------------------------------
def sqdist_f32(v):
    d = v[:, None, :] - v[None, :, :]
    q = (d * d).astype(np.float32)
    return ((q[..., 0] + q[..., 1]).astype(np.float32) + q[..., 2]).astype(np.float32)


def brute_flatten(v, t, r):
    """integrate.cpp:103-150 transcribed literally, radiusSearch = all j != i with float squared distance < float(r*r)."""
    n = len(v)
    r2 = np.float32(float(np.float32(r)) ** 2) if r > 0 else np.float32(0)
    near = sqdist_f32(v) < r2 if n else np.zeros((0, 0), bool)
    remap = [-1] * n; new = []
    for i in range(n):
        if remap[i] >= 0:
            continue
        remap[i] = len(new)
        for j in range(n):
            if j != i and near[i, j]:
                remap[j] = len(new)
        new.append(v[i])
    faces = []
    for a, b, c in t:
        a, b, c = remap[a], remap[b], remap[c]
        if a == b or b == c or c == a:
            continue
        faces.append((a, b, c))
    return np.float32(new).reshape(-1, 3), np.int32(faces).reshape(-1, 3)


def brute_cleanup(v, t, d, k):
    """integrate.cpp:152-214 with scipy's connected components as the cluster extraction."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    m = len(t)
    if m == 0:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    c = ((v[t[:, 0]] + v[t[:, 1]]).astype(np.float32) + v[t[:, 2]]).astype(np.float32) / np.float32(3)
    r2 = np.float32(float(np.float32(d)) ** 2)
    adj = sqdist_f32(c.astype(np.float32)) < r2
    np.fill_diagonal(adj, False)
    _, lab = connected_components(csr_matrix(adj), directed=False)
    size = np.bincount(lab)[lab]
    keep = size > k
    tt = t[keep]
    used = np.zeros(len(v), bool); used[tt.reshape(-1)] = True
    newidx = np.cumsum(used) - 1
    return v[used], newidx[tt].astype(np.int32).reshape(-1, 3)


@settings(max_examples=60, deadline=None)
@given(meshes(), st.sampled_from([1e-4, 5e-5]))
def test_flatten_restatement_agrees_with_a_literal_brute_force_transcription(mesh, r):
    a = oracle_py.flatten_vertices(*mesh, r); b = brute_flatten(*mesh, r)
    assert a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])


@settings(max_examples=60, deadline=None)
@given(meshes(), st.sampled_from([0.02, 0.005, 0.08]), st.integers(1, 8))
def test_cleanup_restatement_agrees_with_connected_components(mesh, d, k):
    v, t = mesh
    a = oracle_py.cleanup_mesh(v, t, d, k); b = brute_cleanup(v, t, d, k)
    assert a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])

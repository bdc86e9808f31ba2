"""Global validation of the marching-cubes case tables (tools/gen_mc_tables.py): applied to
smooth closed implicit surfaces on a dense grid they must produce closed, consistently oriented
2-manifolds with the right Euler characteristic.  Vertices are identified by the grid edge they
sit on, so the check is purely combinatorial (no floating-point welding)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("gen_mc_tables", os.path.join(ROOT, "tools", "gen_mc_tables.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def polygonise(field):
    edge = gen.edge_table()
    nx, ny, nz = field.shape
    tris = []
    for x in range(nx - 1):
        for y in range(ny - 1):
            for z in range(nz - 1):
                corner = [(x + c[0], y + c[1], z + c[2]) for c in gen.CORNERS]
                vals = [field[c] for c in corner]
                ci = sum(1 << k for k in range(8) if vals[k] < 0)
                if edge[ci] == 0:
                    continue
                t = gen.TRI[ci]
                for i in range(0, len(t), 3):
                    tri = []
                    for e in t[i:i + 3]:
                        a, b = gen.EDGES[e]
                        tri.append(tuple(sorted((corner[a], corner[b]))))   # global edge id
                    tris.append(tri)
    return tris


def topology(tris):
    vid = {}
    F = []
    for t in tris:
        F.append([vid.setdefault(v, len(vid)) for v in t])
    F = np.array(F)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    _, dcnt = np.unique(e, axis=0, return_counts=True)
    und, ucnt = np.unique(np.sort(e, axis=1), axis=0, return_counts=True)
    return len(vid), len(und), len(F), dcnt.max(), set(ucnt.tolist())


def test_tables_pass_structural_checks():
    assert gen.validate(gen.TRI, gen.edge_table())


def test_sphere_is_closed_oriented_genus0():
    n = 20
    g = (np.indices((n, n, n)).astype(np.float64) - (n - 1) / 2 + 0.13)
    field = np.sqrt((g ** 2).sum(0)) - 6.3
    V, E, F, dmax, ucnt = topology(polygonise(field))
    assert ucnt == {2} and dmax == 1          # closed and consistently oriented
    assert V - E + F == 2                      # a sphere


def test_random_blobs_and_complement():
    rng = np.random.default_rng(3)
    n = 18
    g = np.indices((n, n, n)).astype(np.float64)
    field = np.full((n, n, n), 10.0)
    for _ in range(9):
        c = rng.uniform(4, n - 5, 3)
        r = rng.uniform(1.7, 3.6)
        field = np.minimum(field, np.sqrt(((g - c[:, None, None, None]) ** 2).sum(0)) - r)
    field += rng.normal(scale=0.15, size=field.shape)     # exercise ambiguous cases
    field[0, :, :] = field[-1, :, :] = field[:, 0, :] = field[:, -1, :] = field[:, :, 0] = field[:, :, -1] = 5.0
    V, E, F, dmax, ucnt = topology(polygonise(field))
    assert ucnt == {2} and dmax == 1
    assert (V - E + F) % 2 == 0
    # the complement uses the mirrored cases: same closedness (grid border now inside, so pad)
    f2 = np.pad(-field, 1, constant_values=5.0)
    V, E, F, dmax, ucnt = topology(polygonise(f2))
    assert ucnt == {2} and dmax == 1

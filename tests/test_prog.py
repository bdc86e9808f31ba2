"""The host programs above the C ABI: PCD / pose / PLY I/O on the CPU, and on the GPU the
whole `b200_integrate` pipeline (PCD directory -> unorganised clouds -> volume -> mesh -> flatten/cleanup -> PLY,
volume.tsdf) against the oracle's pipeline, plus `b200_tsdf2mesh`."""
import os
import struct
import subprocess

import numpy as np
import pytest

from cpu_tsdf_b200 import synth
from cpu_tsdf_b200.build import BIN, build_programs
from tests.common import CAM

PT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("bgra", "u1", 4)])


@pytest.fixture(scope="module")
def progs():
    build_programs()
    return {n: os.path.join(BIN, n) for n in ("b200_integrate", "b200_tsdf2mesh", "b200_pcd_convert")}


def lzf_compress(data: bytes) -> bytes:
    """A small liblzf-style compressor (hash of 3-byte sequences) so that the reader sees real back references."""
    out = bytearray(); lit = bytearray(); table = {}; i = 0; n = len(data)

    def flush():
        for k in range(0, len(lit), 32):
            run = lit[k:k + 32]; out.append(len(run) - 1); out.extend(run)
        lit.clear()
    while i < n:
        key = data[i:i + 3]; ref = table.get(key); table[key] = i
        if ref is not None and len(key) == 3 and i - ref <= 8191 + 1 - 1 and i - ref >= 1:
            ln = 3
            while i + ln < n and ln < 264 and data[ref + ln] == data[i + ln]:
                ln += 1
            flush()
            dist = i - ref - 1; l2 = ln - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8)); out.append(l2 - 7)
            out.append(dist & 0xff)
            i += ln
        else:
            lit.append(data[i]); i += 1
    flush()
    return bytes(out)


def sample_points(n=5000, seed=3):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, PT)
    p["x"], p["y"], p["z"] = rng.normal(size=(3, n)).astype(np.float32)
    p["z"][::50] = np.nan
    p["x"][::7] = np.float32(0.25)                       # repeated values -> back references in the LZF stream
    p["bgra"] = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    return p


def write_pcd(path, pts, mode, *, color="rgba", width=None, height=1):
    n = len(pts)
    width = width or n
    if color == "rgba":
        hdr = "FIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
    elif color == "rgb":
        hdr = "FIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
    else:
        hdr = "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
    head = f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n{hdr}WIDTH {width}\nHEIGHT {height}\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n"
    cbits = np.ascontiguousarray(pts["bgra"]).view("<u4")[:, 0]
    with open(path, "wb") as f:
        f.write(head.encode())
        if mode == "ascii":
            for i in range(n):
                t = " ".join("nan" if np.isnan(v) else f"{v:.9g}" for v in (pts["x"][i], pts["y"][i], pts["z"][i]))
                if color == "rgba":
                    t += f" {cbits[i]}"
                elif color == "rgb":
                    t += f" {cbits[i:i + 1].view('<f4')[0]:.9g}"
                f.write((t + "\n").encode())
        elif mode == "binary":
            f.write(pts.tobytes() if color else np.stack([pts["x"], pts["y"], pts["z"]], 1).astype("<f4").tobytes())
        else:
            cols = [pts["x"].tobytes(), pts["y"].tobytes(), pts["z"].tobytes()] + ([cbits.tobytes()] if color else [])
            soa = b"".join(cols); comp = lzf_compress(soa)
            assert len(comp) < len(soa)
            f.write(struct.pack("<II", len(comp), len(soa))); f.write(comp)


def dump(progs, path, tmp):
    out = os.path.join(tmp, "dump.bin")
    r = subprocess.run([progs["b200_pcd_convert"], path, out, "--dump"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w, h, n, c = map(int, r.stdout.split())
    return np.fromfile(out, PT), (w, h, n, c)


def same_points(a, b, color=True):
    ok = all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in "xyz")
    return ok and (not color or np.array_equal(a["bgra"], b["bgra"]))


@pytest.mark.parametrize("mode", ["ascii", "binary", "binary_compressed"])
def test_pcd_reader_all_encodings(progs, tmp_path, mode):
    pts = sample_points()
    p = str(tmp_path / "a.pcd")
    write_pcd(p, pts, mode)
    got, (w, h, n, c) = dump(progs, p, str(tmp_path))
    assert (w, h, n, c) == (len(pts), 1, len(pts), 1) and same_points(pts, got)
    # re-encode with the tool's own writer into every mode and read back
    for m2 in ("ascii", "binary", "binary_compressed"):
        q = str(tmp_path / f"b_{m2}.pcd")
        assert subprocess.run([progs["b200_pcd_convert"], p, q, m2]).returncode == 0
        assert same_points(pts, dump(progs, q, str(tmp_path))[0])


def test_pcd_reader_field_variants(progs, tmp_path):
    pts = sample_points(600)
    good = ~np.isnan(np.ascontiguousarray(pts["bgra"]).view("<u4")[:, 0].view("<f4"))      # float-typed rgb cannot carry NaN payloads through text
    pts = pts[good]
    p = str(tmp_path / "rgbf.pcd")
    write_pcd(p, pts, "binary", color="rgb")                                # legacy float-typed rgb field
    assert same_points(pts, dump(progs, p, str(tmp_path))[0])
    p = str(tmp_path / "xyz.pcd")
    write_pcd(p, pts[:480], "ascii", color=None, width=24, height=20)       # organized, no colour
    got, meta = dump(progs, p, str(tmp_path))
    assert meta == (24, 20, 480, 0) and same_points(pts[:480], got, color=False)
    assert (got["bgra"] == [0, 0, 0, 255]).all()                           # PointXYZRGBA defaults
    # truncated / corrupt files are reported, not crashed on
    raw = open(p, "rb").read()
    bad = str(tmp_path / "bad.pcd"); open(bad, "wb").write(raw[: len(raw) // 2])
    assert subprocess.run([progs["b200_pcd_convert"], bad, bad + ".o", "--dump"], capture_output=True).returncode == 1
    assert subprocess.run([progs["b200_pcd_convert"], __file__, bad + ".o", "--dump"], capture_output=True).returncode == 1


def test_integrate_program_argument_handling(progs, tmp_path):
    r = subprocess.run([progs["b200_integrate"], "--help"], capture_output=True, text=True)
    assert r.returncode == 1 and "--in" in r.stdout and "--zero-nans" in r.stdout
    assert subprocess.run([progs["b200_integrate"], "--in", str(tmp_path)], capture_output=True).returncode == 1       # --out missing
    assert subprocess.run([progs["b200_integrate"], "--in", str(tmp_path), "--out", str(tmp_path), "--bogus"], capture_output=True).returncode == 1
    assert subprocess.run([progs["b200_integrate"], "--in", str(tmp_path), "--out", str(tmp_path), "--cloud-only"], capture_output=True).returncode == 2
    assert subprocess.run([progs["b200_tsdf2mesh"]], capture_output=True).returncode == 1


def test_cloud_and_pose_files_are_paired_even_below_directories_with_digits(progs, tmp_path):
    d = tmp_path / "run2" / "seq3"; d.mkdir(parents=True)
    for f in range(3):
        write_pcd(str(d / f"frame_{f:04d}.pcd"), sample_points(50), "binary")
        (d / f"frame_{f:04d}.txt").write_text("1 0 0 0\n0 1 0 0\n0 0 1 0\n")
    r = subprocess.run([progs["b200_integrate"], "--in", str(d), "--out", str(tmp_path / "o")], capture_output=True, text=True)
    assert f"Found PCD files with prefix: {d}/frame_, poses with prefix: {d}/frame_" in r.stdout
    assert "Could not find matching transform file" not in r.stderr
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "no CUDA device" in r.stderr                  # the product has no CPU path
    (d / "frame_0001.txt").unlink()
    r = subprocess.run([progs["b200_integrate"], "--in", str(d), "--out", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not find matching transform file" in r.stderr


# ---------------------------------------------------------------------------------------------------------------------
def read_ply(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    color = any("red" in l for l in lines)
    if "binary_little_endian" in lines[1]:
        vd = np.dtype([("xyz", "<f4", 3)] + ([("rgb", "u1", 3)] if color else []))
        v = np.frombuffer(body, vd, nv)
        fd = np.dtype([("n", "u1"), ("idx", "<i4", 3)])
        f = np.frombuffer(body, fd, nf, offset=nv * vd.itemsize)
        assert (f["n"] == 3).all() and nv * vd.itemsize + nf * fd.itemsize == len(body)
        return v["xyz"].copy(), (v["rgb"].copy() if color else None), f["idx"].copy()
    rows = body.decode().split("\n")
    v = np.array([r.split() for r in rows[:nv]], dtype=np.float64)
    f = np.array([r.split() for r in rows[nv:nv + nf]], dtype=np.int64)
    return v[:, :3].astype(np.float32), (v[:, 3:6].astype(np.uint8) if color else None), f[:, 1:].astype(np.int32)


@pytest.mark.gpu
def test_integrate_program_end_to_end_matches_the_oracle_pipeline(progs, tmp_path):
    from oracle import oracle_py
    from tests.test_organize import unorganized_cloud
    d = tmp_path / "seq"; d.mkdir()
    out = tmp_path / "out"
    o = oracle_py.OracleVolume(xres=256, yres=256, zres=256, xsize=3.0, ysize=3.0, zsize=3.0, cx=CAM.cx, cy=CAM.cy, integrate_color=1,
                               min_sensor_dist=0.0)
    o.reset()
    intr = (525.0, 525.0, CAM.cx, CAM.cy)
    for f in range(4):
        cloud, _ = unorganized_cloud(20 + f, n_extra=20000, scale=0.001)            # millimetres
        pose = synth.orbit_pose(synth.S1, 4 * f, 100) if f else np.eye(4)
        pose = pose.astype(np.float32).astype(np.float64)                           # what a 12-float pose file can hold
        pts = np.zeros(len(cloud), PT)
        pts["x"], pts["y"], pts["z"] = cloud[:, 0], cloud[:, 1], cloud[:, 2]
        pts["bgra"] = np.ascontiguousarray(cloud[:, 4]).view(np.uint8).reshape(-1, 4)
        write_pcd(str(d / f"frame_{f:04d}.pcd"), pts, ("binary", "binary_compressed", "ascii", "binary")[f])
        with open(d / f"frame_{f:04d}.txt", "w") as fh:
            for r in range(3):
                fh.write(" ".join(f"{v:.9g}" for v in pose[r]) + "\n")
        # the oracle's pipeline for this frame: integrate.cpp:548-607 then integrateCloud (first pose is the identity,
        # so pose_rel_to_first_frame is the pose itself)
        org, _ = oracle_py.organize(cloud, intr, CAM.width, CAM.height, rgba_off=16, cloud_units=0.001, zero_nans=True)
        o.integrate(org, pose)
    args = [progs["b200_integrate"], "--in", str(d), "--volume-size", "3", "--cell-size", "0.0117", "--color", "--cloud-units", "0.001",
            "--zero-nans", "--save-tsdf", "--flatten", "--cleanup", "--pool-log2", "16"]
    r = subprocess.run(args + ["--out", str(out)], capture_output=True, text=True)                   # the fast update kernels
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Setting resolution: 256" in r.stdout
    # with the variance accumulators kept (general update kernel) the volume on disk is byte-identical to the oracle's
    out2 = tmp_path / "out_exact"
    r = subprocess.run(args + ["--out", str(out2), "--exact-vol"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ref_vol = str(tmp_path / "ref.vol")
    assert o.save(ref_vol) == 0
    assert open(ref_vol, "rb").read() == open(out2 / "volume.tsdf", "rb").read()
    assert open(out / "mesh.ply", "rb").read() == open(out2 / "mesh.ply", "rb").read()
    # the mesh: marching cubes (min weight 0, coloured), flattenVertices, cleanupMesh
    verts, _ = o.mesh(0.0, 1)
    soup = (np.asarray(verts, np.float32).reshape(-1, 3), np.arange(len(verts), dtype=np.int32).reshape(-1, 3))
    want = oracle_py.cleanup_mesh(*oracle_py.flatten_vertices(*soup))
    gv, gc, gt = read_ply(str(out / "mesh.ply"))
    assert gc is None                                                               # colour does not survive flatten (as in the reference)
    assert np.array_equal(want[0].view(np.uint32), gv.view(np.uint32)) and np.array_equal(want[1], gt)
    # tsdf2mesh on the saved volume: the plain soup, binary PLY
    ply2 = str(tmp_path / "m2.ply")
    assert subprocess.run([progs["b200_tsdf2mesh"], str(out2 / "volume.tsdf"), ply2], capture_output=True).returncode == 0
    v2, c2, t2 = read_ply(ply2)
    w2, _ = o.mesh(2.5, 0)                                                          # MarchingCubesTSDFOctree's default min weight
    assert c2 is None and np.array_equal(np.asarray(w2, np.float32).reshape(-1, 3).view(np.uint32), v2.view(np.uint32))
    assert np.array_equal(t2.reshape(-1), np.arange(len(v2), dtype=np.int32))

"""bench.py --impl reference runs on CPU: check the JSON contract of the reference arm (and, by
construction, of the keys the GPU arm shares) without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, B200TSDF_REF_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("integrateCloud frames/s") and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 2 and cb["value"] == line["value"]
    assert "workload" in line["config"] and line["data"] == "synthetic" and line["vs_baseline"] is None


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.parametrize("name", ["r1_bench.json", "r2_bench.json"])
def test_committed_gpu_bench_line_has_every_contract_key(name):
    """profiles/rN_bench.json is the line `python bench.py` printed on the B200 box at the end of round N: check its shape against
    the contract."""
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "l2" in d["config"] and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    # bytes that crossed PCIe per 32-frame step: the caller's 32-byte points (round 1), or the 16-byte pixels the host packed them to
    packed = bool(e.get("host_pack", {}).get("enabled"))
    assert e["unit"] == "frames/s" and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 32 * 640 * 480 * (16 if packed else 32) and e["d2h_bytes_per_step"] > 0
    if packed:
        assert e["host_pack"]["input_bytes_per_step"] == 32 * 640 * 480 * 32 and e["host_pack"]["threads"] >= 1
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "frames/s" and c["sample"]
    assert d["gpu_launches"] == 4 * 32 * d["steps"]                      # four kernels per frame, 32 frames per step
    assert abs(d["value"] - 32 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    if name == "r2_bench.json":
        # round 2: one graph launch per 32-frame step, a batched roofline figure next to the event-timed one, the loaded-host leg
        assert d["graph_launches"] == d["steps"]
        b = r["batched"]
        assert b["frames_per_graph_launch"] == 32 and abs(b["frac"] - b["achieved"] / r["peak"]) < 1e-9 and b["frac"] >= r["frac"]
        hl = d["host_load_leg"]
        assert hl["unit"] == "frames/s" and len(hl["repeats"]) == 3 and abs(hl["value"] - sorted(hl["repeats"])[1]) < 1e-9
        assert hl["value"] > 0.9 * d["value"]                            # the device-resident rate does not depend on an idle host

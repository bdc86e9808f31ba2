"""The C-ABI library loads and exports every symbol include/b200tsdf.h declares; without a CUDA
device it refuses to create a volume (there is no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import cpu_tsdf_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200tsdf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200tsdf_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(engine_lib):
    so = ctypes.CDLL(engine_lib)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(so, s), f"{s} declared in include/b200tsdf.h but not exported"
    assert sorted(pkg.EXPORTS) == syms


def test_config_struct_layout(engine_lib):
    lib = pkg.load_library()
    cfg = pkg.Config()
    lib.b200tsdf_default_config(ctypes.byref(cfg))
    # TSDFVolumeOctree ctor defaults (src/lib/tsdf_volume_octree.cpp:54-85)
    assert (cfg.xres, cfg.yres, cfg.zres) == (512, 512, 512)
    assert (cfg.xsize, cfg.max_weight, cfg.image_width, cfg.image_height) == (3.0, 100.0, 640, 480)
    assert abs(cfg.max_dist_pos - 0.03) < 1e-8 and abs(cfg.min_sensor_dist - 0.3) < 1e-7 and cfg.max_sensor_dist == 3.0
    assert (cfg.fx, cfg.fy, cfg.cx, cfg.cy) == (525.0, 525.0, 320.0, 240.0)
    assert cfg.max_cell_x == 0.5 and cfg.integrate_color == 0 and cfg.shard_count == 1
    assert list(cfg.global_transform) == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    assert ctypes.sizeof(pkg.Config) == 264


def test_no_device_is_a_loud_error(engine_lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    lib = pkg.load_library()
    h = ctypes.c_void_p()
    cfg = pkg.Config()
    lib.b200tsdf_default_config(ctypes.byref(cfg))
    assert lib.b200tsdf_create(ctypes.byref(cfg), ctypes.byref(h)) == -2     # B200TSDF_ENODEVICE
    with pytest.raises(pkg.B200Error):
        pkg.TSDFVolumeOctree()


def test_product_does_not_reference_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "cpu_tsdf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "oracle_py" not in text and "tsdf_oracle" not in text and "libemu" not in text, f

"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/spnb.h,
argument validation mirrors the reference's ValueErrors, module state matches the reference's
buffer names, and the product refuses to run without the GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(spn):
    from smoothparticlenets_b200 import _native, build
    hdr = open(os.path.join(ROOT, "include", "spnb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(spnb_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(build.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    L = _native.lib()
    assert L.spnb_version() >= 100
    assert L.spnb_max_cartesian_dim() == 20
    assert L.spnb_hashgrid_workspace_bytes(8, 65536, 3, 96) > 4 * 4 * 8 * 65536
    assert L.spnb_hashgrid_workspace_bytes(0, 1, 1, 1) == 0


def test_library_is_sm100a_and_torch_free():
    """The C ABI must not depend on torch and must carry sm_100a code."""
    import subprocess
    from smoothparticlenets_b200 import build
    build.build_library()
    ldd = subprocess.run(["ldd", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_host_validation_returns_errors_without_a_gpu(spn):
    """Bad arguments are rejected on the host before any CUDA call (status 0 + message)."""
    from smoothparticlenets_b200 import _native
    L = _native.lib()
    assert L.spnb_convsp_forward(None, None, None, None, None, None, 1, 1, 1, 1, 3, 8, 1, 1, 0.1, None,
                                 None, 0, 0, None, None) == 0
    assert b"null pointer" in L.spnb_last_error()
    assert L.spnb_convsp_forward(None, None, None, None, None, None, 1, 1, 1, 1, 3, 8, 1, 1, 0.1, None,
                                 None, 0, 12, None, None) == 0
    assert b"kernel function" in L.spnb_last_error()
    assert L.spnb_convsdf_forward(None, 1, 1, 4, None, None, None, 1, 4, None, 1, None, None, 1, None, None,
                                  1, 1, None, None, 1.0, None, None) == 0
    assert b"1-, 2- and 3-D" in L.spnb_last_error()
    assert L.spnb_hashgrid_order(None, None, None, None, None, None, 0, 1, (1 << 24) + 1, 3, 0.1, 96, None) == 0
    assert b"2^24" in L.spnb_last_error()


def test_constructor_validation_matches_reference(spn):
    with pytest.raises(ValueError):
        spn.ConvSP(0, 1, 3, 1, 1, 0.1)
    with pytest.raises(ValueError):
        spn.ConvSP(1, 1, 3, 2, 1, 0.1)             # even kernel size
    with pytest.raises(ValueError):
        spn.ConvSP(1, 1, 3, (3, 3), 1, 0.1)        # wrong list length
    with pytest.raises(ValueError):
        spn.ConvSP(1, 1, 20, 1, 1, 0.1)            # ndim < MAX_CARTESIAN_DIM
    with pytest.raises(ValueError):
        spn.ConvSP(1, 1, 3, 1, 1, 0.1, kernel_fn="nope")
    with pytest.raises(ValueError):
        spn.ParticleCollision(3, -1.0)
    with pytest.raises(ValueError):
        spn.ConvSDF([torch.zeros(2, 2)], [1.0], 1, 3, 1, 1, 1.0)   # 2-D SDF for ndim 3
    with pytest.raises(ValueError):
        spn.ConvSDF([torch.zeros(2, 2, 2, 2)], [1.0], 1, 4, 1, 1, 1.0)  # ndim must be 1..3
    c = spn.ConvSP(4, 8, 3, 3, 0.05, 0.1, kernel_fn="spiky")
    assert c.kernel_fn == 11 and c.ncells == 27 and c.weight.shape == (8, 4, 27)
    assert isinstance(c.weight, torch.nn.Parameter)
    c2 = spn.ConvSP(1, 1, 3, 1, 1, 0.1, with_params=False)
    assert not isinstance(c2.weight, torch.nn.Parameter) and "weight" in dict(c2.named_buffers())
    with pytest.raises(ValueError):
        c(torch.zeros(2, 5, 2), torch.zeros(2, 5, 4), torch.zeros(2, 5, 8))  # wrong ndim


def test_state_dict_names_match_reference(spn):
    """Checkpoint compatibility: same registered names as the reference."""
    pc = spn.ParticleCollision(3, 0.1)
    assert set(pc.state_dict()) == {"cellIDs", "cellStarts", "cellEnds", "cuda_buffer"}
    ref_like = {"cellIDs": torch.zeros(4, 10, 1), "cellStarts": torch.zeros(2, 27),
                "cellEnds": torch.zeros(2, 27), "cuda_buffer": torch.zeros(100)}
    pc.load_state_dict(ref_like)  # scratch buffers of any shape load
    cs = spn.ConvSP(2, 3, 2, (3, 1), 0.05, 1.0)
    assert set(cs.state_dict()) == {"weight", "bias", "kernel_size", "dilation"}
    sd = spn.ConvSDF([torch.zeros(3, 4, 5), torch.ones(2, 2, 2)], [0.5, 0.25], 2, 3, 3, 0.01, 1.0)
    assert set(sd.state_dict()) == {"sdfs", "sdf_shapes", "sdf_offsets", "weight", "bias", "kernel_size",
                                    "dilation"}
    assert sd.sdf_offsets.tolist() == [0.0, 60.0]
    assert sd.sdf_shapes.tolist() == [[3, 4, 5, 0.5], [2, 2, 2, 0.25]]
    assert sd.sdfs.numel() == 68


def test_no_cpu_fallback(spn):
    pc = spn.ParticleCollision(2, 0.2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pc(torch.rand(1, 10, 2))
    conv = spn.ConvSP(1, 1, 2, 1, 1, 0.2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.rand(1, 10, 2), torch.rand(1, 10, 1), torch.full((1, 10, 4), -1.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        spn.ReorderData()(torch.zeros(1, 10), torch.rand(1, 10, 2))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under smoothparticlenets_b200/ may reference it."""
    pkg = os.path.join(ROOT, "smoothparticlenets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "spn_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_tile_list_layout_mirror_matches_library(spn):
    """smoothparticlenets_b200/tile_lists.py (the host-side mirror tests decode the sidecar with) and
    csrc/tile_lists.cuh must agree on the buffer size; unsupported shapes report 0 (no GPU needed)."""
    from smoothparticlenets_b200 import _native as nat, tile_lists as tl
    L = nat.lib()
    for B, N, D, K in [(1, 1, 3, 16), (2, 1000, 3, 128), (8, 65536, 3, 128), (3, 777, 2, 64), (2, 300, 1, 32),
                       (1, 4099, 3, 256), (1, 100, 3, 77)]:
        assert L.spnb_tile_lists_bytes(B, N, D, K) == tl.layout(B, N, K)["total"], (B, N, D, K)
    assert L.spnb_tile_lists_bytes(1, 100, 4, 128) == 0    # ndims > 3
    assert L.spnb_tile_lists_bytes(1, 100, 3, 1000) == 0   # longer lists than the format is built for
    assert L.spnb_tile_lists_bytes(1, 100, 3, 100) > 0     # any max_collisions up to 512
    assert L.spnb_tile_lists_bytes(0, 100, 3, 128) == 0
    # null pointers are rejected on the host
    assert L.spnb_compute_collisions_tiled(None, None, None, None, None, None, None, None, 1, 100, 3, 128, 96 ** 3,
                                           0.1, 0.1, 0, None, None, 0, None) == 0
    assert b"null pointer" in L.spnb_last_error()
    assert L.spnb_pbf_stage1_forward(None, None, None, None, None, None, None, 10, 3, 1.0, 1.0, None) == 0
    assert b"bad arguments" in L.spnb_last_error()


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the reference's CPU functions on the host cores, forked workers) on a tiny sample:
    exactly one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-particles", "512"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "particles/s"

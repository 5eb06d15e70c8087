"""The C-ABI library loads on a machine WITHOUT a GPU, exports every symbol include/svcm.h
declares, agrees with the ctypes mirror, and refuses to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import smallvcm_b200 as sv
from smallvcm_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "svcm.h")).read()
    return sorted(set(re.findall(r"SVCM_API[^;(]*?\b(svcm_\w+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    lib = sv.load_library()
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "libsvcm.so does not export %s" % name
    # the ctypes mirror knows exactly the header's symbols
    assert sorted(abi.PROTOTYPES) == names


def test_struct_layouts_match_header():
    # sizes implied by include/svcm.h
    assert C.sizeof(abi.Primitive) == 56
    assert C.sizeof(abi.Material) == 44          # the reference's Material is 44 bytes too
    assert C.sizeof(abi.Light) == 4 + 4 * (3 * 7 + 2)
    assert C.sizeof(abi.Camera) == 4 * (3 + 3 + 2 + 16 + 16 + 1)
    assert C.sizeof(abi.SceneSphere) == 20
    assert C.sizeof(abi.MergeVertex) == 52       # the 52-byte merge record of SURVEY §8 a3
    assert C.sizeof(abi.MergeQuery) == 52
    assert C.sizeof(abi.Params) == 48


def test_version_and_scene_masks():
    lib = sv.load_library()
    assert b"sm_100a" in lib.svcm_version()
    # g_SceneConfigs (reference src/config.hxx:146-151)
    assert lib.svcm_scene_config_mask(0) == 256 | 64 | 128 | 2
    assert lib.svcm_scene_config_mask(1) == 256 | 16 | 1
    assert lib.svcm_scene_config_mask(2) == 256 | 64 | 128 | 4
    assert lib.svcm_scene_config_mask(3) == 256 | 64 | 128 | 8
    assert lib.svcm_scene_config_mask(4) == 0
    assert lib.svcm_scene_config_mask(-1) == 0


def test_invalid_arguments_are_rejected_without_touching_a_gpu():
    lib = sv.load_library()
    scene = sv.Scene.cornell_box(1, 32, 32)
    p = abi.Params()
    p.algorithm, p.radiusFactor, p.radiusAlpha, p.seed = 9, 0.003, 0.75, 1
    p.device, p.rank, p.worldSize = 0, 0, 1
    h = C.c_void_p()
    assert lib.svcm_create(C.byref(scene.desc), C.byref(p), C.byref(h)) == abi.E_INVALID
    assert b"algorithm" in lib.svcm_last_error(None)
    p.algorithm, p.worldSize, p.rank = abi.ALG_VCM, 2, 2
    assert lib.svcm_create(C.byref(scene.desc), C.byref(p), C.byref(h)) == abi.E_INVALID
    assert lib.svcm_create(None, None, C.byref(h)) == abi.E_INVALID
    assert not h.value
    with pytest.raises(sv.SvcmError):
        sv.Scene.cornell_box(7, 32, 32)
    with pytest.raises(sv.SvcmError):
        sv.Scene.cornell_box(1, 0, 32)


def test_no_cpu_fallback():
    """Without a CUDA device every computing entry point fails with SVCM_E_NO_DEVICE."""
    lib = sv.load_library()
    if lib.svcm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    scene = sv.Scene.cornell_box(1, 32, 32)
    with pytest.raises(sv.SvcmError) as e:
        sv.VertexCM(scene, "vcm")
    assert e.value.code == abi.E_NO_DEVICE
    with pytest.raises(sv.SvcmError) as e:
        sv.intersect_batch(scene, [[0, 0, 0, 0, 1, 0]])
    assert e.value.code == abi.E_NO_DEVICE
    with pytest.raises(sv.SvcmError) as e:
        sv.rng_stream_device(1, 0, 0, 0, 4, 4)
    assert e.value.code == abi.E_NO_DEVICE
    with pytest.raises(sv.SvcmError) as e:
        sv.image_convert_batch(np.zeros((2, 2, 3), np.float32), 1, "hdr")
    assert e.value.code == abi.E_NO_DEVICE
    # a scene large enough for the hierarchy is accepted by the host code and then fails the same way
    big = scene.with_triangles(np.random.default_rng(0).uniform(-1, 1, (100, 3, 3)), np.full(100, 5))
    with pytest.raises(sv.SvcmError) as e:
        sv.VertexCM(big, "vcm")
    assert e.value.code == abi.E_NO_DEVICE


def test_host_only_entry_points_reject_bad_arguments():
    lib = sv.load_library()
    scene = sv.Scene.cornell_box(0, 16, 16)
    nodes = (abi.BvhNode * 64)()
    order = np.zeros(22, np.int32)
    count = C.c_int32()
    assert lib.svcm_bvh_build(None, 22, nodes, 64, order.ctypes.data, C.byref(count)) == abi.E_INVALID
    assert lib.svcm_bvh_build(scene.desc.primitives, 0, nodes, 64, order.ctypes.data, C.byref(count)) == abi.E_INVALID
    assert lib.svcm_bvh_build(scene.desc.primitives, 22, nodes, 2, order.ctypes.data, C.byref(count)) == abi.E_NOMEM
    assert 11 <= count.value <= 43                     # the size the caller should have provided
    assert lib.svcm_bvh_build(scene.desc.primitives, 22, nodes, 64, order.ctypes.data, C.byref(count)) == abi.OK
    assert lib.svcm_image_bytes(abi.IMAGE_RGBE, -1, 4) == -1
    assert lib.svcm_download_image(None, 0, 2.2, None, None) == abi.E_INVALID
    assert lib.svcm_save_image(None, b"x.bmp", 2.2) == abi.E_INVALID


def test_shard_map_is_a_partition():
    """Every path is owned by exactly one rank and local/global indices invert each other."""
    lib = sv.load_library()
    for (w, h, ws) in [(7, 5, 1), (7, 5, 2), (16, 9, 4), (5, 3, 8), (33, 17, 3)]:
        total = 0
        seen = set()
        for rank in range(ws):
            n = lib.svcm_shard_local_paths(w, h, rank, ws)
            total += n
            for l in range(n):
                g = lib.svcm_shard_global_index(w, rank, ws, l)
                assert 0 <= g < w * h and g not in seen
                seen.add(g)
                assert lib.svcm_shard_owner(w, ws, g) == rank
                assert lib.svcm_shard_local_index(w, ws, g) == l
        assert total == w * h and len(seen) == w * h
    assert lib.svcm_shard_local_paths(4, 4, 2, 2) == -1


def test_header_is_plain_c_and_the_example_links(tmp_path):
    """include/svcm.h must be usable from C (the drop-in boundary is a C ABI): the C99 example
    compiles warning-free with -pedantic, links against the library, and without a GPU exits with
    the library's "no CUDA device" error rather than computing anything."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "render_box")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                          os.path.join(root, "examples", "render_box.c"), "-L" + os.path.dirname(sv.library_path()),
                          "-lsvcm", "-Wl,-rpath," + os.path.dirname(sv.library_path()), "-o", exe],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    if sv.load_library().svcm_device_count() > 0:
        return
    run = subprocess.run([exe, "1", "vcm", "32", "1", str(tmp_path / "o.bmp")], capture_output=True, text=True)
    assert run.returncode == 2 and "no CUDA device" in run.stderr

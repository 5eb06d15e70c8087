"""svcm_scene_cornell_box reproduces, byte for byte, the reference's own Scene flattened by
the host glue — against committed fixtures (tests/golden/scenes.npz, made from the compiled
reference by tests/golden/make_golden.py) and, where oracle/_ref exists, against it directly."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as oa
import smallvcm_b200 as sv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenes.npz")


def raw(x):
    return np.frombuffer(C.string_at(C.addressof(x), C.sizeof(x)), np.uint8)


def desc_arrays(d):
    return dict(prims=np.concatenate([raw(d.primitives[i]) for i in range(d.numPrimitives)]),
                mats=np.concatenate([raw(d.materials[i]) for i in range(d.numMaterials)]),
                lights=np.concatenate([raw(d.lights[i]) for i in range(d.numLights)]),
                m2l=np.array([d.materialToLight[i] for i in range(d.numMaterials)], np.int32),
                camera=raw(d.camera), sphere=raw(d.sceneSphere),
                background=np.array([d.backgroundLight], np.int32))


@pytest.mark.parametrize("scene_id", [0, 1, 2, 3])
@pytest.mark.parametrize("res", [(64, 48), (512, 512)])
def test_against_golden(scene_id, res):
    gold = np.load(GOLD)
    scene = sv.Scene.cornell_box(scene_id, res[0], res[1])
    mine = desc_arrays(scene.desc)
    key = "s%d_%dx%d_" % (scene_id, res[0], res[1])
    assert sv.load_library().svcm_scene_config_mask(scene_id) == int(gold[key + "mask"][0])
    for name, value in mine.items():
        assert np.array_equal(value, gold[key + name]), (scene_id, res, name)


def test_primitive_counts_and_lights():
    # 22 / 21 / 12 / 22 primitives, 9 materials, 1-2 lights
    expect = {0: (22, 1, sv.abi.LIGHT_DIRECTIONAL), 1: (21, 2, sv.abi.LIGHT_AREA),
              2: (12, 1, sv.abi.LIGHT_POINT), 3: (22, 1, sv.abi.LIGHT_BACKGROUND)}
    for sid, (nprim, nlight, ltype) in expect.items():
        scene = sv.Scene.cornell_box(sid, 32, 32)   # keep the owner alive while reading
        d = scene.desc
        assert (d.numPrimitives, d.numLights, d.numMaterials) == (nprim, nlight, 9)
        assert d.lights[0].type == ltype
        assert d.backgroundLight == (0 if sid == 3 else -1)
    scene = sv.Scene.cornell_box(1, 32, 32)
    d = scene.desc
    assert abs(d.sceneSphere.radius - 2.21706) < 1e-4


@pytest.mark.ref
@pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (no reference checkout)")
@pytest.mark.parametrize("scene_id", [0, 1, 2, 3])
def test_against_reference_directly(scene_id):
    for res in [(96, 64), (1024, 1024), (333, 77)]:
        a, b = sv.Scene.cornell_box(scene_id, *res), oa.ref_flatten_scene(scene_id, *res)
        mine, ref = desc_arrays(a.desc), desc_arrays(b.desc)
        for name in mine:
            assert np.array_equal(mine[name], ref[name]), (scene_id, res, name)

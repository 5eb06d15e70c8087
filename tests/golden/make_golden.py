"""Regenerates the committed fixtures in tests/golden/ from the COMPILED REFERENCE.

Run in a container that has /root/reference (it builds oracle/_ref first):
    python tests/golden/make_golden.py
Everything written here is an OUTPUT of the reference (images, scalars, flattened scene bytes),
never its source. The fixtures let the CPU tests pin oracle/vcm_oracle.cpp and the scene
builder on machines without the reference (the GPU box, CI).

Files
  ref_kat.json        known-answer scalars of the unmodified reference (BASELINE.md §3 recipe)
  ref_images.npz      unmodified reference, 64x64, 2 iterations, 12 (scene, algorithm) pairs
  refctr_images.npz   reference arithmetic + per-path counter RNG, same 8 pairs, + vertex counts
  ref_noise.npz       unmodified reference 128x128 x16 iterations, two seeds (noise floor for
                      the statistical GPU-vs-reference test)
  scenes.npz          the reference's Scene for g_SceneConfigs[0..3], flattened, raw bytes
  ref_extra_scene.npz unmodified reference rendering box 1 + 1 160 extra triangles appended to its
                      GeometryList (a scene that is not one of the hard-coded boxes): the triangles, the
                      reference's flattened primitive bytes, 64x64 vcm and bpt images, 2 iterations
  image_files.npz     the files Framebuffer::SaveBMP / SaveHDR wrote for one small accumulated
                      image with hostile values (NaN, inf, negatives, denormals), 0 and 3 iterations
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_api as oa  # noqa: E402
from smallvcm_b200 import abi  # noqa: E402

PAIRS = [(1, "vcm"), (0, "bpt"), (3, "bpm"), (2, "vcm"), (1, "lt"), (1, "ppm"), (0, "vcm"), (3, "vcm"),
         (0, "pt"), (1, "pt"), (3, "pt"), (2, "el")]


def raw(x):
    return np.frombuffer(C.string_at(C.addressof(x), C.sizeof(x)), np.uint8).copy()


def image_input():
    """A 24x16 accumulated image: smooth values, then a block of hostile ones."""
    rng = np.random.default_rng(77)
    img = (rng.random((16, 24, 3), dtype=np.float32) ** 4 * 3).astype(np.float32)
    flat = img.reshape(-1)
    flat[0:20] = [np.nan, np.inf, -np.inf, -1.5, 0, 1e-33, 1e-32, 1e-31, 255.99998, 1e30, 3.4e38, -0.0, 1, 2, 4,
                  0.5, 0.99999994, 1.9999999, 127.99999, 1e-38]
    flat[20:400] = np.exp(rng.uniform(-80, 80, 380)).astype(np.float32) * rng.choice([1, 1, 1, -1], 380)
    return img


def image_fixture():
    import tempfile
    img = image_input()
    out = {"accum": img}
    with tempfile.TemporaryDirectory() as d:
        for it in (0, 3):
            out["bmp_it%d" % it] = np.frombuffer(oa.ref_save_image(img, it, 0, 2.2, os.path.join(d, "a.bmp")), np.uint8)
            out["hdr_it%d" % it] = np.frombuffer(oa.ref_save_image(img, it, 1, 1.0, os.path.join(d, "a.hdr")), np.uint8)
    np.savez_compressed(os.path.join(HERE, "image_files.npz"), **out)
    print("image_files.npz", {k: v.shape for k, v in out.items()})


def extra_scene_fixture():
    import scene_gen
    b = scene_gen.blob((0.0, 0.25, 0.15), 0.38, 16, 32, 0.25, 5)
    c = scene_gen.confetti(200, (-1.1, -0.9, -1.1), (1.1, 1.1, 1.1), 0.03, 6)
    verts = np.concatenate([b, c])
    mats = np.concatenate([np.where(np.arange(len(b)) % 2 == 0, 2, 8), np.full(len(c), 5)]).astype(np.int32)
    out = {"vertices": verts, "materials": mats}
    for alg in ("vcm", "bpt"):
        img, prims, sphere = oa.ref_render_extra(1, alg, 64, 64, 2, verts, mats)
        out["image_" + alg] = img
    import smallvcm_b200 as sv
    n = sv.Scene.cornell_box(1, 64, 64).desc.numPrimitives + len(verts)
    out["primitives"] = np.frombuffer(C.string_at(prims, n * 56), np.uint8).reshape(n, 56).copy()
    out["sphere"] = np.array(list(sphere.center) + [sphere.radius, sphere.invRadiusSqr], np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_extra_scene.npz"), **out)
    print("ref_extra_scene.npz", {k: v.shape for k, v in out.items()})


def main():
    oa.build_oracle()
    image_fixture()
    extra_scene_fixture()
    assert oa.have_ref() and oa.have_ref_ctr(), "needs the reference checkout"

    # --- known answers of the unmodified reference (1 renderer, seed 1234, minLen 0) ---------
    kat = []
    for sid, alg, w, iters in [(1, "vcm", 512, 1), (1, "vcm", 512, 2), (0, "bpt", 512, 1), (3, "bpm", 512, 1),
                               (2, "vcm", 256, 1), (1, "lt", 256, 1), (1, "ppm", 256, 1), (0, "pt", 512, 1),
                               (0, "el", 256, 1)]:
        _, img = oa.ref_render(sid, alg, w, w, iters, threads=1)
        kat.append(dict(scene=sid, alg=alg, width=w, height=w, iterations=iters, seed=1234, maxPathLength=10,
                        totalLuminance=float(oa.ref_lib().ref_total_luminance(img.ctypes.data, w, w)),
                        meanRGB=[float(v) for v in img.reshape(-1, 3).astype(np.float64).mean(0)]))
        print(kat[-1])
    json.dump(kat, open(os.path.join(HERE, "ref_kat.json"), "w"), indent=1)

    # --- small images, both generators ----------------------------------------------------------
    w = 64
    ref, ctr = {}, {}
    for sid, alg in PAIRS:
        key = "s%d_%s" % (sid, alg)
        _, img = oa.ref_render(sid, alg, w, w, 2, threads=1)   # el: iteration 1 is the unjittered one
        ref[key] = img
        r = oa.RefCtr(sid, abi.ALGORITHMS[alg], w, w, seed=1234)
        r.run_iteration(0); r.run_iteration(1)
        acc, n = r.accum()
        ctr[key] = acc
        ctr[key + "_nverts"] = np.array([r.num_light_vertices()], np.int64)
        r.close()
    np.savez_compressed(os.path.join(HERE, "ref_images.npz"), **ref)
    np.savez_compressed(os.path.join(HERE, "refctr_images.npz"), **ctr)

    # --- noise floor for the statistical test ---------------------------------------------------
    noise = {}
    for sid, alg in [(1, "vcm"), (3, "bpm"), (0, "bpt"), (1, "pt")]:
        for seed in (1234, 1235):
            _, img = oa.ref_render(sid, alg, 128, 128, 16, threads=1, base_seed=seed)
            noise["s%d_%s_seed%d" % (sid, alg, seed)] = img.astype(np.float16)
    np.savez_compressed(os.path.join(HERE, "ref_noise.npz"), **noise)

    # --- flattened scenes -------------------------------------------------------------------------
    scenes = {}
    for sid in range(4):
        for (sw, sh) in [(64, 48), (512, 512)]:
            sc = oa.ref_flatten_scene(sid, sw, sh)
            d = sc.desc
            k = "s%d_%dx%d_" % (sid, sw, sh)
            scenes[k + "prims"] = np.concatenate([raw(d.primitives[i]) for i in range(d.numPrimitives)])
            scenes[k + "mats"] = np.concatenate([raw(d.materials[i]) for i in range(d.numMaterials)])
            scenes[k + "lights"] = np.concatenate([raw(d.lights[i]) for i in range(d.numLights)])
            scenes[k + "m2l"] = np.array([d.materialToLight[i] for i in range(d.numMaterials)], np.int32)
            scenes[k + "camera"] = raw(d.camera)
            scenes[k + "sphere"] = raw(d.sceneSphere)
            scenes[k + "background"] = np.array([d.backgroundLight], np.int32)
            scenes[k + "mask"] = np.array([oa.ref_lib().ref_scene_mask(sid)], np.uint32)
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **scenes)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

"""The image conversions after the path: Framebuffer::SaveBMP / SaveHDR
(reference src/framebuffer.hxx:170-251) restated in the oracle and computed on the device.

CPU tests pin the oracle's pixel conversions and the library's file headers to files the COMPILED
reference wrote (tests/golden/image_files.npz; live against oracle/_ref where it exists).
GPU tests compare the device conversion with the oracle: RGBE must be byte-exact (IEEE multiply,
divide and frexp only); BMP goes through pow(), where the device's double-precision pow and
glibc's powf may round differently on sub-ulp ties, so a byte may differ by one only where
pow(c)*255 lies within 1e-3 of an integer, and in at most 1e-5 of the bytes."""
import os

import numpy as np
import pytest

import oracle_api as oa
import smallvcm_b200 as sv
from smallvcm_b200 import abi

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_files.npz"))


def hostile_image(h, w, seed):
    rng = np.random.default_rng(seed)
    img = (rng.random((h, w, 3), dtype=np.float32) ** 4 * 3).astype(np.float32)
    flat = img.reshape(-1)
    k = min(len(flat) // 3, 20000)
    flat[:k] = np.exp(rng.uniform(-80, 80, k)).astype(np.float32) * rng.choice([1, 1, 1, -1], k)
    special = [np.nan, np.inf, -np.inf, -1.5, 0, 1e-33, 1e-32, 255.99998, 3.4e38, -0.0, 0.99999994, 1e-38]
    flat[:min(12, len(flat))] = special[:min(12, len(flat))]
    return img


# ---------------------------------------------------------------------------------------------
# CPU: oracle and headers against the reference's files
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("it", [0, 3])
def test_oracle_payloads_match_reference_files(it):
    img = GOLD["accum"]
    h, w = img.shape[:2]
    bmp, hdr = GOLD["bmp_it%d" % it].tobytes(), GOLD["hdr_it%d" % it].tobytes()
    assert len(bmp) == 54 + w * h * 3
    assert np.array_equal(np.frombuffer(bmp[54:], np.uint8), oa.oracle_bmp_payload(img, it, 2.2))
    text = b"#?RADIANCE\n# SmallVCM\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w)
    assert hdr.startswith(text) and len(hdr) == len(text) + w * h * 4
    assert np.array_equal(np.frombuffer(hdr[len(text):], np.uint8), oa.oracle_hdr_payload(img, it))


def test_library_headers_match_reference_files():
    img = GOLD["accum"]
    h, w = img.shape[:2]
    assert sv.image_header("bmp", w, h) == GOLD["bmp_it0"].tobytes()[:54]
    hdr = GOLD["hdr_it0"].tobytes()
    assert sv.image_header("hdr", w, h) == hdr[:len(hdr) - 4 * w * h]
    lib = sv.load_library()
    assert lib.svcm_image_bytes(abi.IMAGE_BMP24, 7, 5) == 105 and lib.svcm_image_bytes(abi.IMAGE_RGBE, 7, 5) == 140
    assert lib.svcm_image_bytes(9, 7, 5) == -1 and lib.svcm_image_header(9, 7, 5, None, 0) == -1
    assert lib.svcm_image_header(abi.IMAGE_BMP24, 7, 5, None, 0) == 54            # size query


@pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (no reference checkout)")
@pytest.mark.parametrize("gamma", [2.2, 1.0, 0.7])
def test_oracle_payloads_match_live_reference(tmp_path, gamma):
    img = hostile_image(61, 83, 3)
    for it in (0, 1, 7):
        ref = oa.ref_save_image(img, it, 0, gamma, tmp_path / "a.bmp")
        assert ref[:54] == sv.image_header("bmp", 83, 61)
        assert np.array_equal(np.frombuffer(ref[54:], np.uint8), oa.oracle_bmp_payload(img, it, gamma))
        ref = oa.ref_save_image(img, it, 1, gamma, tmp_path / "a.hdr")
        head = sv.image_header("hdr", 83, 61)
        assert ref[:len(head)] == head
        assert np.array_equal(np.frombuffer(ref[len(head):], np.uint8), oa.oracle_hdr_payload(img, it))


# ---------------------------------------------------------------------------------------------
# GPU: device conversion against the oracle
# ---------------------------------------------------------------------------------------------
def check_bmp(dev, img, it, gamma):
    want = oa.oracle_bmp_payload(img, it, gamma)
    bad = np.nonzero(dev != want)[0]
    assert len(bad) <= max(1, int(1e-5 * len(want))), len(bad)
    if len(bad):
        assert np.all(np.abs(dev[bad].astype(int) - want[bad].astype(int)) == 1)
        h, w = img.shape[:2]
        # undo the payload order: byte k belongs to output pixel k // 3, channel 2 - k % 3
        px = bad // 3
        y, x = px // w, px % w
        c = img[h - 1 - y, x, 2 - bad % 3].astype(np.float64) * (np.float32(1.0 / it) if it > 0 else 1.0)
        g = c ** (1.0 / np.float32(gamma)) * 255.0
        assert np.all(np.abs(g - np.round(g)) < 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(16, 24), (61, 83), (512, 512), (1, 1)])
def test_device_conversions_match_oracle(shape):
    img = hostile_image(shape[0], shape[1], 11) if shape != (16, 24) else GOLD["accum"]
    for it in (0, 3):
        assert np.array_equal(sv.image_convert_batch(img, it, "hdr"), oa.oracle_hdr_payload(img, it))
        for gamma in (2.2, 1.0):
            check_bmp(sv.image_convert_batch(img, it, "bmp", gamma), img, it, gamma)
    if shape == (16, 24):   # and therefore the reference's own file bytes
        assert np.array_equal(sv.image_convert_batch(img, 3, "hdr"), GOLD["hdr_it3"][-4 * 16 * 24:])


@pytest.mark.gpu
def test_empty_image_and_bad_arguments():
    assert sv.image_convert_batch(np.zeros((0, 5, 3), np.float32), 1, "bmp").size == 0
    with pytest.raises(sv.SvcmError):
        sv.image_convert_batch(np.zeros((2, 2, 3), np.float32), 1, "bmp", gamma=0.0)


@pytest.mark.gpu
def test_renderer_image_download_and_saved_files(tmp_path):
    """svcm_download_image / svcm_save_image on a rendered image: payload = oracle conversion of
    the renderer's own accumulated image; the files parse back to header + payload."""
    scene = sv.Scene.cornell_box(1, 160, 96)
    r = sv.VertexCM(scene, "vcm"); r.max_path_length = 10
    r.run_iterations(0, 5)
    accum, n = r.accum()
    assert n == 5
    hdr = r.image8("hdr")
    assert hdr.shape == (96, 160, 4) and np.array_equal(hdr.reshape(-1), oa.oracle_hdr_payload(accum, n))
    bmp = r.image8("bmp", 2.2)
    assert bmp.shape == (96, 160, 3)
    check_bmp(bmp.reshape(-1), accum, n, 2.2)
    assert bmp.max() > 200 and bmp.min() == 0            # a real picture, not a constant
    for name, fmt in (("out.bmp", "bmp"), ("out.hdr", "hdr")):
        r.save_image(tmp_path / name, 2.2)
        data = (tmp_path / name).read_bytes()
        head = sv.image_header(fmt, 160, 96)
        assert data[:len(head)] == head
        assert data[len(head):] == (bmp if fmt == "bmp" else hdr).tobytes()
    with pytest.raises(sv.SvcmError) as e:
        r.save_image(tmp_path / "out.png")
    assert e.value.code == abi.E_INVALID and "unknown extension" in str(e.value)
    # the float image is untouched by the conversions
    assert np.array_equal(r.accum()[0], accum)
    r.close()

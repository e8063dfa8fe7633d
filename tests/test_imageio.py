"""Image writers for the progressive render: PFM / PNG round trips and the accum -> radiance resolve."""
import numpy as np
import pytest

from foundation_b200 import imageio


def test_radiance_from_accum_divides_by_sample_count_and_zeroes_unowned_pixels():
    acc = np.zeros((2, 3, 4), np.float32)
    acc[0, 0] = (2.0, 4.0, 6.0, 4.0)
    acc[1, 2] = (1.0, 0.0, 3.0, 1.0)
    rgb = imageio.radiance_from_accum(acc)
    assert rgb.shape == (2, 3, 3) and rgb.dtype == np.float32
    assert np.array_equal(rgb[0, 0], [0.5, 1.0, 1.5]) and np.array_equal(rgb[1, 2], [1.0, 0.0, 3.0])
    assert not rgb[0, 1].any() and not rgb[1, 0].any()
    with pytest.raises(ValueError):
        imageio.radiance_from_accum(np.zeros((2, 3, 3), np.float32))


def test_pfm_round_trip_is_bit_exact_and_bottom_up(tmp_path):
    rng = np.random.default_rng(1)
    img = rng.random((5, 7, 3), dtype=np.float32) * 100.0
    img[0, 0] = (np.inf, 0.0, 1e-30)
    p = str(tmp_path / "a.pfm")
    imageio.write_pfm(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(b"PF\n7 5\n-1.0\n")
    first_stored_row = np.frombuffer(raw[len(b"PF\n7 5\n-1.0\n"):][:7 * 3 * 4], "<f4").reshape(7, 3)
    assert np.array_equal(first_stored_row, img[-1])                    # the file starts with the BOTTOM row
    assert np.array_equal(imageio.read_pfm(p), img)
    with pytest.raises(ValueError):
        imageio.write_pfm(p, np.zeros((4, 4, 4), np.float32))


@pytest.mark.parametrize("channels", [3, 4])
def test_png_round_trip(tmp_path, channels):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (9, 11, channels), dtype=np.uint8)
    p = str(tmp_path / "a.png")
    imageio.write_png(p, img)
    assert np.array_equal(imageio.read_png(p), img)
    try:
        from PIL import Image
    except ImportError:
        return
    assert np.array_equal(np.asarray(Image.open(p)), img)             # an independent decoder agrees


@pytest.mark.parametrize("channels", [3, 4])
def test_exr_round_trip_and_file_layout(tmp_path, channels):
    """Uncompressed float scan-line OpenEXR: bit-exact round trip, and the bytes a foreign reader relies on — magic, version, the attribute
    list, one offset per scan line pointing at [y][size][channel rows in alphabetical channel order] — checked by hand."""
    import struct
    rng = np.random.default_rng(3)
    img = (rng.random((4, 6, channels), dtype=np.float32) * 50.0).astype(np.float32)
    img[0, 0, 0] = np.inf; img[1, 2, 1] = 1e-30
    p = str(tmp_path / "a.exr")
    imageio.write_exr(p, img)
    assert np.array_equal(imageio.read_exr(p), img)
    raw = open(p, "rb").read()
    assert raw[:8] == struct.pack("<ii", 20000630, 2)                                      # magic 0x762f3101, version 2, no flags (single-part scan lines)
    assert raw[8:8 + 16] == b"channels\0chlist\0"
    names = [b"B", b"G", b"R"] if channels == 3 else [b"A", b"B", b"G", b"R"]
    (n,) = struct.unpack("<i", raw[24:28])
    assert n == 18 * len(names) + 1 and raw[28:30] == names[0] + b"\0" and struct.unpack("<i", raw[30:34])[0] == 2     # FLOAT
    for key in (b"compression\0compression\0", b"dataWindow\0box2i\0", b"displayWindow\0box2i\0", b"lineOrder\0lineOrder\0", b"pixelAspectRatio\0float\0",
                b"screenWindowCenter\0v2f\0", b"screenWindowWidth\0float\0"):
        assert key in raw
    dw = raw.index(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0")
    assert struct.unpack("<iiiii", raw[dw:dw + 20]) == (16, 0, 0, 5, 3)
    end = raw.index(b"screenWindowWidth\0float\0") + len(b"screenWindowWidth\0float\0") + 8 + 1    # size, value, terminating zero of the header
    offs = np.frombuffer(raw[end:end + 8 * 4], "<u8")
    line = 8 + len(names) * 6 * 4
    assert list(offs) == [end + 32 + k * line for k in range(4)] and len(raw) == end + 32 + 4 * line
    y, nbytes = struct.unpack("<ii", raw[int(offs[2]):int(offs[2]) + 8])
    assert (y, nbytes) == (2, line - 8)
    rows = np.frombuffer(raw[int(offs[2]) + 8:int(offs[2]) + line], "<f4").reshape(len(names), 6)
    src = {b"R": 0, b"G": 1, b"B": 2, b"A": 3}
    for k, nm in enumerate(names):
        assert np.array_equal(rows[k], img[2, :, src[nm]])
    with pytest.raises(ValueError):
        imageio.write_exr(p, np.zeros((4, 4), np.float32))
    import os
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    try:
        import cv2
    except ImportError:
        return
    dec = cv2.imread(p, cv2.IMREAD_UNCHANGED)                                              # OpenCV's bundled OpenEXR: an independent decoder (BGR[A] order)
    if dec is not None:
        assert dec.dtype == np.float32 and np.array_equal(dec[..., [2, 1, 0, 3][:channels]], img)

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

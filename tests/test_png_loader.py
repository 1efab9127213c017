"""grt_image_load_png (csrc/png_loader.cpp) - the background loader's PNG decoder - against PIL, on generated images of every
colour type and, where the reference is mounted, on the reference's own background files (graphics_settings.cpp:214-243)."""
import os

import numpy as np
import pytest

PIL_Image = pytest.importorskip("PIL.Image")


def _roundtrip(grt, tmp_path, img, name, **save_kw):
    path = str(tmp_path / name)
    img.save(path, **save_kw)
    got = grt.load_png(path)
    want = np.asarray(PIL_Image.open(path).convert("RGBA"))
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), name


def test_every_colour_type_and_filter(grt, tmp_path):
    rng = np.random.default_rng(2)
    h, w = 37, 53                                   # odd sizes: partial bytes for the packed depths
    smooth = (np.add.outer(np.arange(h) * 3, np.arange(w) * 5) % 256).astype(np.uint8)   # makes the encoder pick Sub/Up/Average/Paeth filters
    noise = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(noise, "RGBA"), "rgba.png")
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(noise[..., :3].copy(), "RGB"), "rgb.png")
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(np.stack([smooth, smooth[::-1], smooth.T[:h, :w] if smooth.T.shape == (h, w) else smooth, smooth], -1), "RGBA"), "smooth_rgba.png", compress_level=9)
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(smooth, "L"), "grey.png")
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(np.stack([smooth, 255 - smooth], -1), "LA"), "grey_alpha.png")
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(noise[..., :3].copy(), "RGB").quantize(16), "palette4bit.png")
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(noise[..., :3].copy(), "RGB").quantize(200), "palette8bit.png")
    _roundtrip(grt, tmp_path, PIL_Image.fromarray((smooth > 127).astype(np.uint8) * 255, "L").convert("1"), "bilevel.png")
    big = rng.integers(0, 256, (300, 1024, 4), dtype=np.uint8)
    big[..., 3] = 255
    _roundtrip(grt, tmp_path, PIL_Image.fromarray(big, "RGBA"), "big.png")


def test_errors_are_reported_not_crashed(grt, tmp_path):
    p = tmp_path / "junk.png"
    p.write_bytes(b"not a png at all")
    with pytest.raises(grt.GrtError, match="not a PNG"):
        grt.load_png(str(p))
    with pytest.raises(grt.GrtError, match="cannot open"):
        grt.load_png(str(tmp_path / "missing.png"))
    good = tmp_path / "ok.png"
    PIL_Image.fromarray(np.zeros((8, 8, 4), np.uint8), "RGBA").save(good)
    data = good.read_bytes()
    (tmp_path / "cut.png").write_bytes(data[:len(data) // 2])
    with pytest.raises(grt.GrtError):
        grt.load_png(str(tmp_path / "cut.png"))


def test_reference_backgrounds_decode_like_pil(grt, have_reference):
    if not have_reference:
        pytest.skip("reference not mounted")
    for rel, size in [("backgrounds/nasa.png", (4096, 2048)), ("backgrounds/esa.png", (4000, 2000)), ("background2.png", (4000, 2000))]:
        path = os.path.join("/root/reference", rel)
        got = grt.load_png(path)
        assert got.shape == (size[1], size[0], 4)
        want = np.asarray(PIL_Image.open(path).convert("RGBA"))
        assert np.array_equal(got, want), rel


def test_png_writer_round_trips(grt, tmp_path):
    """grt_image_save_png (the screenshot file of main.cpp:2783-2812): what it writes, PIL and the library's own decoder read back."""
    from PIL import Image
    rng = np.random.default_rng(3)
    for (w, h) in [(1, 1), (301, 157), (64, 2)]:
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        path = str(tmp_path / f"shot_{w}x{h}.png")
        grt.save_png(path, img)
        assert np.array_equal(np.array(Image.open(path).convert("RGBA")), img)
        assert np.array_equal(grt.load_png(path), img)
    with pytest.raises(grt.GrtError):
        grt.save_png(str(tmp_path / "no_such_dir" / "x.png"), img)

"""CPU: the crop / augment oracle (oracle/augment.py) against (a) Pillow itself, the library the reference's image
arithmetic lives in (anakin/utils/img_augment.py calls ImageEnhance / convert('HSV') / transform(AFFINE), and
rendered_dataset.py:256 ImageFilter.GaussianBlur), and (b) the fixture recorded from the reference's own
RenderedDataset.__getitem__ (tests/golden/make_golden_augment.py)."""
import numpy as np
import pytest

from conftest import golden
from oracle import augment as A


def _img(rng, h, w, kind):
    if kind == 0:
        return rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    if kind == 1:  # blocky: sharp 0 / 255 edges are where the tiny blur rounds up
        a = rng.randint(0, 2, size=(h // 4 + 1, w // 4 + 1, 3)) * 255
        return np.kron(a, np.ones((4, 4, 1)))[:h, :w].astype(np.uint8)
    return (np.linspace(0, 255, w)[None, :, None] * np.ones((h, 1, 3))).astype(np.uint8)


def test_hsv_round_trip_tables_match_pillow_on_every_colour():
    Image = pytest.importorskip("PIL.Image")
    allc = np.arange(1 << 24, dtype=np.uint32)
    cols = np.stack([(allc >> 16) & 255, (allc >> 8) & 255, allc & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(np.array(Image.fromarray(cols).convert("HSV")), A.rgb_to_hsv(cols))
    assert np.array_equal(np.array(Image.fromarray(cols, "HSV").convert("RGB")), A.hsv_to_rgb(cols))
    assert np.array_equal(np.array(Image.fromarray(cols).convert("L")), A.rgb_to_l(cols))


def test_enhancers_and_blur_match_pillow_bit_for_bit():
    Image = pytest.importorskip("PIL.Image")
    from PIL import ImageEnhance, ImageFilter
    rng = np.random.RandomState(0)
    for trial in range(45):
        img = _img(rng, rng.randint(5, 70), rng.randint(5, 90), trial % 3)
        pil = Image.fromarray(img)
        for f in (rng.uniform(0.9, 1.1), rng.uniform(0.0, 2.0), 1.0, 0.0):
            assert np.array_equal(np.array(ImageEnhance.Brightness(pil).enhance(f)), A.adjust_brightness(img, f))
            assert np.array_equal(np.array(ImageEnhance.Contrast(pil).enhance(f)), A.adjust_contrast(img, f))
            assert np.array_equal(np.array(ImageEnhance.Color(pil).enhance(f)), A.adjust_saturation(img, f))
        for r in (0.0, rng.uniform(0, 0.1), 0.1):
            assert np.array_equal(np.array(pil.filter(ImageFilter.GaussianBlur(r))), A.gaussian_blur(img, r)), (trial, r)


def test_affine_nearest_matches_pillow_transform():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(1)
    for trial in range(120):
        hi, wi = rng.randint(20, 100), rng.randint(20, 100)
        img = _img(rng, hi, wi, 0)
        ang = rng.uniform(-0.7, 0.7) if trial % 4 else 0.0     # every 4th: the pure-scaling path (ImagingScaleAffine)
        sc = rng.uniform(0.3, 3.0)
        m = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-40, 40)], [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-40, 40)],
                      [0, 0, 1]], np.float32)
        inv = np.linalg.inv(m)
        co = [float(inv[0, 0]), float(inv[0, 1]), float(inv[0, 2]), float(inv[1, 0]), float(inv[1, 1]), float(inv[1, 2])]
        wo, ho = rng.randint(10, 80), rng.randint(10, 80)
        ref = np.array(Image.fromarray(img).transform((wo, ho), Image.AFFINE, tuple(co)))
        assert np.array_equal(ref, A.affine_nearest(img, co, (wo, ho))), trial
        ours = A.invert_affine(m)  # closed form vs LAPACK: a few fp32 ulps
        np.testing.assert_allclose(ours, np.array(co, np.float32), rtol=2e-6, atol=2e-5)


def _draws(g, i):
    f = g["factors"][i]
    return {"center_jit": g["center_jit"][i], "scale_jit": g["scale_jit"][i], "rot_cs": g["rot_cs"][i],
            "blur_radius": g["blur_radius"][i], "brightness": f[0], "contrast": f[1], "saturation": f[2], "hue": f[3],
            "order": g["order"][i]}


CFG = {"crop_model": "root_obj", "bbox_expand_ratio": 1.2, "aug": True, "center_jit": 0.1, "scale_jit": 0.1, "center_idx": 0}


def test_rendered_sample_matches_the_reference_getitem():
    """Annotations to fp32 accuracy; the jittered (pre-warp) image and -- given the reference's own inverse coefficients
    -- the warped network input bit for bit; with our closed-form inverse < 0.5 % of the pixels move."""
    g = golden("augment.npz")
    cfg = dict(CFG, image_size=tuple(g["out_size"]), raw_size=tuple(g["raw_size"]))
    moved = total = 0
    for i in range(len(g["img"])):
        d = _draws(g, i)
        out = A.rendered_sample(g["img"][i], g["joints"][i], g["pose"][i], g["corners_can"][i], g["K"], d, cfg)
        for k in ("cam_intr", "root_joint", "joints_3d", "joints_2d", "corners_3d", "corners_2d", "obj_transf"):
            np.testing.assert_allclose(out[k], g[k][i], rtol=2e-5, atol=2e-4 if k.endswith("2d") or k == "cam_intr" else 2e-6, err_msg=f"{k}[{i}]")
        assert np.array_equal(out["joints_vis"], g["joints_vis"][i]) and np.array_equal(out["corners_vis"], g["corners_vis"][i])
        pre = A.color_jitter(A.gaussian_blur(g["img"][i], d["blur_radius"]), d)
        assert np.array_equal(pre, g["pre_warp"][i]), f"blur + jitter composition differs for sample {i}"
        ref_u8 = np.rint((g["image"][i] + 0.5) * 255.0).astype(np.uint8).transpose(1, 2, 0)
        same_coef = A.affine_nearest(pre, g["rev"][i], cfg["image_size"])
        assert np.array_equal(same_coef, ref_u8)
        img32 = (same_coef.astype(np.float32) / np.float32(255.0) - np.float32(0.5)).transpose(2, 0, 1)
        assert np.array_equal(img32, g["image"][i]), "to_tensor / normalise arithmetic"
        moved += int((out["image_u8"] != ref_u8).any(-1).sum())
        total += ref_u8.shape[0] * ref_u8.shape[1]
        np.testing.assert_allclose(out["inv_affine"], g["rev"][i].astype(np.float32), rtol=3e-6, atol=3e-5)
    assert g["joints_vis"][3].sum() == 0 and g["joints_vis"][0].sum() > 0
    assert moved / total < 5e-3, (moved, total)


def test_edge_cases():
    rng = np.random.RandomState(3)
    img = _img(rng, 9, 7, 0)
    assert np.array_equal(A.gaussian_blur(img, 0.0), img)
    assert A.affine_nearest(img, [1, 0, 100, 0, 1, 0], (5, 5)).sum() == 0          # entirely outside: fill 0
    assert np.array_equal(A.affine_nearest(img, [1, 0, 0, 0, 1, 0], (7, 9)), img)   # identity
    assert A.hue_shift_byte(-0.05) == (256 - 12) and A.hue_shift_byte(0.05) == 12
    one = np.full((1, 1, 3), 200, np.uint8)
    assert np.array_equal(A.adjust_contrast(one, 1.5), one)                          # mean == pixel
    with pytest.raises(NotImplementedError):
        A.gaussian_blur(img, 3.0)

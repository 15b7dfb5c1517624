"""Parity of the sm_100a ORB extractor (through the C ABI) with the CPU oracle: bit-exact pyramid bytes, keypoint
selection/order, keypoint fields and descriptor bytes."""
import numpy as np
import pytest
import ucoslam_b200
import orb_oracle as oo

pytestmark = pytest.mark.gpu


def _check_frame(ctx, img, prm_kw, stages=True):
    prm = ucoslam_b200.OrbParams(**prm_kw)
    K, D = ctx.orb_extract(img, prm)
    oK, oD, inter = oo.extract(np.ascontiguousarray(img), prm_kw.get("max_features", 2000), prm_kw.get("n_levels", 8),
                               prm_kw.get("scale_factor", 1.2), prm_kw.get("ini_th", 20), prm_kw.get("min_th", 7),
                               prm_kw.get("blur_first", True), want_intermediates=True)
    if stages:
        for l, ref in enumerate(inter["pyramid"]):
            got = ctx.orb_pyramid_level(0, l)
            assert got.shape == ref.shape, "level %d size" % l
            assert np.array_equal(got, ref), "level %d bytes differ at %d px" % (l, (got != ref).sum())
        for l, lv in enumerate(inter["levels"]):
            got = ctx.orb_selected(0, l)
            ref = lv["selected"]
            assert len(got) == len(ref), "level %d count %d vs %d" % (l, len(got), len(ref))
            assert np.array_equal(got & 0xfff, ref["x"].astype(np.uint32)), "level %d x" % l
            assert np.array_equal((got >> 12) & 0xfff, ref["y"].astype(np.uint32)), "level %d y" % l
            assert np.array_equal(got >> 24, ref["response"].astype(np.uint32)), "level %d score" % l
    assert len(K) == len(oK)
    for fld in oo.KP_DTYPE.names:
        assert np.array_equal(K[fld].view(np.uint32), oK[fld].view(np.uint32)), fld
    assert np.array_equal(D, oD)


def test_orb_640x480_2000(ctx):
    for i in (0, 17):
        _check_frame(ctx, oo.synth_frame(i), dict(max_features=2000))


def test_orb_1280x720_4000(ctx):
    _check_frame(ctx, oo.synth_frame(5, 1280, 720), dict(max_features=4000))


@pytest.mark.parametrize("w,h,nf,nl,sf", [(752, 480, 1000, 8, 1.2), (641, 479, 1500, 6, 1.3), (320, 240, 500, 4, 1.5),
                                            (1241, 376, 2000, 8, 1.2), (640, 480, 800, 4, 2.0), (800, 600, 900, 4, 1.7)])
def test_orb_odd_shapes(ctx, w, h, nf, nl, sf):
    _check_frame(ctx, oo.synth_frame(2, w, h), dict(max_features=nf, n_levels=nl, scale_factor=sf))


def test_orb_low_texture_fallback_threshold(ctx):
    """Smooth image: most cells fall back to minThFAST or stay empty; quota redistribution is exercised."""
    import cv2
    img = cv2.GaussianBlur(oo.synth_frame(1), (0, 0), 3)
    img = (img // 3 + 60).astype(np.uint8)
    _check_frame(ctx, img, dict(max_features=2000))


def test_orb_random_noise_and_no_blur(ctx):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    _check_frame(ctx, img, dict(max_features=2000))
    _check_frame(ctx, oo.synth_frame(4), dict(max_features=2000, blur_first=False))


def test_orb_strided_input_and_batch(ctx):
    frames = [oo.synth_frame(i) for i in range(4)]
    big = np.zeros((480, 704), np.uint8)
    big[:, :640] = frames[0]
    K, D = ctx.orb_extract(big[:, :640], ucoslam_b200.OrbParams(2000))
    oK, oD = oo.extract(frames[0])
    assert np.array_equal(D, oD) and np.array_equal(K["x"], oK["x"])
    kps, desc, n = ctx.orb_extract_batch(frames, ucoslam_b200.OrbParams(2000))
    for i, fr in enumerate(frames):
        oK, oD = oo.extract(fr)
        assert n[i] == len(oK)
        assert np.array_equal(desc[i][:n[i]], oD)
        for fld in oo.KP_DTYPE.names:
            assert np.array_equal(kps[i][:n[i]][fld].view(np.uint32), oK[fld].view(np.uint32)), fld


def test_orb_constant_image_gives_no_keypoints(ctx):
    img = np.full((480, 640), 128, np.uint8)
    K, D = ctx.orb_extract(img, ucoslam_b200.OrbParams(2000))
    assert len(K) == 0 and D.shape == (0, 32)
    oK, oD = oo.extract(img)
    assert len(oK) == 0


def test_orb_bad_arguments(ctx):
    img = np.zeros((32, 32), np.uint8)
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.orb_extract(img, ucoslam_b200.OrbParams(2000))
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.orb_extract(oo.synth_frame(0), ucoslam_b200.OrbParams(2000, n_levels=40))

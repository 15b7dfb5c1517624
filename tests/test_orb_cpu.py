"""CPU-side pins for the extractor: the exact-arithmetic helpers compiled into the product library (host copies of the device
code) against the real libraries the reference calls, and the ORB oracle's own invariants."""
import ctypes
import numpy as np
import cv2
import pytest
import ucoslam_b200
import orb_oracle as oo


def test_fast_atan2_matches_opencv():
    rng = np.random.default_rng(0)
    y = rng.integers(-300000, 300000, 60000).astype(np.float32)
    x = rng.integers(-300000, 300000, 60000).astype(np.float32)
    y[:500] = 0; x[500:1000] = 0; x[:5] = 0
    got = ucoslam_b200.probe_fast_atan2(y, x)
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_sincos_matches_libm():
    rng = np.random.default_rng(1)
    deg = np.concatenate([rng.uniform(0, 360, 400000), np.arange(0, 360, 0.25)]).astype(np.float32)
    ang = (deg * oo.FACTOR_PI).astype(np.float32)
    s, c = ucoslam_b200.probe_sincos(ang)
    rs = np.empty_like(ang); rc = np.empty_like(ang)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    oo.stl().stl_sincosf_array(p(ang), len(ang), p(rc), p(rs))
    assert np.array_equal(s.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), rc.view(np.uint32))


@pytest.mark.parametrize("seed", range(6))
def test_retain_best_matches_libstdcxx(seed):
    rng = np.random.default_rng(seed)
    for it in range(400):
        n = int(rng.integers(1, 3000 if it % 40 == 0 else 160))
        rngmax = int(rng.integers(1, 5 if it % 3 == 0 else 120))
        mode = it % 4
        sc = {0: rng.integers(0, rngmax, n), 1: np.arange(n) % rngmax, 2: (n - np.arange(n)) % rngmax,
              3: (np.arange(n) * 7919) % rngmax}[mode].astype(np.uint32)
        want = int(rng.integers(0, n + 2))
        packed = (sc << 24) | np.arange(n, dtype=np.uint32)
        got = ucoslam_b200.probe_retain_best(packed, want)
        k = np.zeros(n, oo.KP_DTYPE)
        k["response"] = sc.astype(np.float32)
        k["class_id"] = np.arange(n)
        ref = oo.retain_best(k, want)[:want]
        assert len(got) == len(ref)
        assert np.array_equal(got & 0xffffff, ref["class_id"].astype(np.uint32))


def test_oracle_params_match_survey_table():
    P = oo.Params(2000, 8, 1.2)
    assert P.n_per_level == [434, 362, 302, 251, 209, 175, 145, 122]
    assert [P.level_size(640, 480, l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                              (257, 193), (214, 161), (179, 134)]
    assert oo.Params(4000, 8, 1.2).n_per_level == [869, 724, 603, 503, 419, 349, 291, 242]
    assert oo.UMAX == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_oracle_extract_invariants():
    img = oo.synth_frame(3)
    K, D = oo.extract(img)
    assert len(K) == 2000 and D.shape == (2000, 32)
    assert (np.diff(K["octave"]) >= 0).all()
    assert (K["angle"] >= 0).all() and (K["angle"] < 360).all()
    assert (K["x"] >= 19).all() and (K["x"] <= 640).all()


def test_vectorised_orientation_equals_the_per_keypoint_restatement():
    """oracle/orb_oracle.py: ic_angles (one gather per level, used by extract and by bench.py's CPU baseline) == ic_angle (the
    line-by-line restatement of IC_Angle, ORBextractor.cpp:79-106) on random patches"""
    import orb_oracle
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    xs, ys = rng.integers(20, 140, 200), rng.integers(20, 100, 200)
    a = orb_oracle.ic_angles(img, xs, ys)
    b = np.array([orb_oracle.ic_angle(img, None, int(x), int(y)) for x, y in zip(xs, ys)], np.float32)
    assert np.array_equal(a, b)


def test_strip_kernel_index_arithmetic_matches_opencv():
    """scripts/emulate_pyramid_strip.py: the index arithmetic of blur7_strip_kernel / resize_cubic_strip_kernel (bordered output domain, reflected
    coordinates, packed u16x2 horizontal blur, 7-row register window, staged source tile, byte-permute selectors) emulated statement by statement
    against cv2.GaussianBlur / cv2.resize / cv2.copyMakeBorder on small images, aligned and unaligned sources, scale factors 1.2 ... 2.0"""
    import importlib.util, os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "emulate_pyramid_strip.py")
    spec = importlib.util.spec_from_file_location("emulate_pyramid_strip", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.main() == 0

"""Parity of stereo.cu (through the C ABI) with the oracle: depths bit-exact (f32), associations identical; synthetic pairs and, at
BASELINE config 3's size (1280x720, 4000 keypoints per image), keypoints and descriptors from the device ORB extractor itself."""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200.synth import synth_stereo
from test_stereo_oracle import CASES, edge_scene

pytestmark = pytest.mark.gpu


def check(ctx, sc, md=50.0):
    want = oracle_py.stereo_depth(sc, md)
    got = ctx.stereo_depth(sc, md)
    assert got[2] == want[2]
    assert np.array_equal(got[1], want[1])
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32))
    return got


@pytest.mark.parametrize("name", list(CASES))
def test_matches_oracle(ctx, name):
    assert check(ctx, synth_stereo(**CASES[name]))[2] > 50


def test_edges_thresholds_and_strides(ctx):
    sc = edge_scene()
    for md in (50.0, 20.0, 1.0, 300.0):
        check(ctx, sc, md)
    empty = dict(sc)
    empty["kps_r"], empty["desc_r"] = sc["kps_r"][:0], sc["desc_r"][:0]
    assert check(ctx, empty)[2] == 0
    none = dict(sc)
    none["kps_l"], none["desc_l"] = sc["kps_l"][:0], sc["desc_l"][:0]
    assert ctx.stereo_depth(none)[2] == 0
    # strided images and descriptor rows (cv::Mat ROIs)
    big_l = np.zeros((120, 200), np.uint8); big_l[:, :160] = sc["img_l"]
    wide = np.zeros((len(sc["desc_l"]), 48), np.uint8); wide[:, :32] = sc["desc_l"]
    st = dict(sc); st["img_l"] = big_l[:, :160]; st["desc_l"] = wide[:, :32]
    check(ctx, st)


def test_right_keypoint_near_border_fails_like_the_reference(ctx):
    sc = synth_stereo(6, w=160, h=120, n=50, max_disp=12.0)
    sc["kps_l"]["x"][0], sc["kps_l"]["y"][0] = 40.0, 60.0
    sc["kps_r"]["x"][0], sc["kps_r"]["y"][0] = 5.0, 60.0          # 3 px inside, but the -7 offset leaves the image
    sc["kps_r"]["octave"][0] = sc["kps_l"]["octave"][0]
    sc["desc_r"][0] = sc["desc_l"][0]
    assert oracle_py.stereo_depth(sc)[2] == -1
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.stereo_depth(sc)


def test_config3_pair_with_device_orb_keypoints(ctx):
    sc = synth_stereo(9, w=1280, h=720, n=10, max_disp=64.0)
    prm = ucoslam_b200.OrbParams(4000)
    kl, dl = ctx.orb_extract(sc["img_l"], prm)
    kr, dr = ctx.orb_extract(sc["img_r"], prm)
    assert len(kl) > 3000 and len(kr) > 3000
    pair = dict(sc, kps_l=kl, desc_l=dl, kps_r=kr, desc_r=dr, fx=np.float32(1050.0))
    d, m, n = check(ctx, pair)
    assert n > 500
    # the recovered disparities follow the rendered ones
    rows = np.clip(np.round(kl["y"]).astype(int), 0, 719)
    disp = 8.0 + (64.0 - 8.0) * (0.5 + 0.5 * np.sin(rows / 720 * 2 * np.pi))
    ok = d != 0
    assert np.median(np.abs(pair["bl"] * pair["fx"] / d[ok] - disp[ok])) < 1.0

"""Oracle restatement of the callers either side of the path (nerfmlp/data.py ray tables and
preprocessing, scripts/render_example.py post-processing) pinned to vectors produced by the
reference itself (tests/golden/make_golden.py data_case).  CPU only."""
import numpy as np

from oracle import nerf_oracle as O
from tests.conftest import load_golden


def test_dataset_rays_and_colours_match_reference():
    g = load_golden("data_3x12x12")
    N, S = g["rgba"].shape[0], g["rgba"].shape[1]
    focal = 0.5 * S / np.tan(0.5 * float(g["camera_angle_x"]))                  # data.py:73
    assert focal == float(g["focal"])
    ro, rd = O.dataset_rays(g["poses"], S, S, focal)
    assert np.array_equal(ro, g["all_rays_o"]) and np.array_equal(rd, g["all_rays_d"])   # bit-exact (same numpy ops)
    rgb = O.preprocess_rgba(g["rgba"]).reshape(-1, 3)
    assert np.array_equal(rgb, g["all_rgbs"])
    idx = g["idx"]
    assert np.array_equal(ro[idx], g["batch_ray_o"]) and np.array_equal(rd[idx], g["batch_ray_d"])
    assert np.array_equal(rgb[idx], g["batch_rgb"])


def test_postprocess_matches_reference():
    g = load_golden("data_3x12x12")
    for boost in (1.0, 1.5):
        for gamma in (False, True):
            got = O.to_uint8(g["pp_in"], boost, gamma)
            assert np.array_equal(got, g[f"pp_out_b{boost}_g{int(gamma)}"]), (boost, gamma)

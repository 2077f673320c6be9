"""Device-side ray generation / batching and output post-processing
against the oracle and the reference's golden vectors.  Bars:

  * rays_o: bit-exact (a copy of pose[:3,3]).
  * rays_d: the kernel evaluates the reference's float64 expression and rounds once to float32;
    numpy's matmul may order/fuse the three float64 products differently, which can move the
    float32 rounding of an exact tie by 1 ulp: >= 99.9 % bit-exact, all within 1 float32 ulp.
  * target colours from raw RGBA: float64 compositing then float32 sRGB decode; CUDA powf vs numpy
    powf differ by <= 2 ulp: <= 3e-7 abs; from preprocessed float images: bit-exact.
  * uint8 post-processing: bit-exact without gamma; with gamma (powf) <= 1 LSB on <= 0.1 % of values.
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def nb():
    import nerf_mlp_b200
    return nerf_mlp_b200


def ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def check_dirs(got, ref):
    u = ulps(got, ref)
    assert u.max() <= 1 and np.mean(u == 0) >= 0.999, (u.max(), np.mean(u == 0))


def test_dataset_golden(nb):
    g = load_golden("data_3x12x12")
    ds = nb.data.DeviceRayDataset(g["rgba"], g["poses"], camera_angle_x=float(g["camera_angle_x"]))
    assert len(ds) == g["all_rays_o"].shape[0] and ds.focal == float(g["focal"])
    idx = torch.from_numpy(g["idx"]).to(DEV)
    o, d, c = ds.batch(idx)
    assert np.array_equal(o.cpu().numpy(), g["batch_ray_o"])
    check_dirs(d.cpu().numpy(), g["batch_ray_d"])
    np.testing.assert_allclose(c.cpu().numpy(), g["batch_rgb"], atol=3e-7, rtol=0)
    # the whole table, view by view
    hw = ds.H * ds.W
    for img in range(ds.N):
        o, d, c = ds.view_rays(img, with_rgb=True)
        assert np.array_equal(o.cpu().numpy(), g["all_rays_o"][img * hw:(img + 1) * hw])
        check_dirs(d.cpu().numpy(), g["all_rays_d"][img * hw:(img + 1) * hw])
        np.testing.assert_allclose(c.cpu().numpy(), g["all_rgbs"][img * hw:(img + 1) * hw], atol=3e-7, rtol=0)
    it = ds[int(g["idx"][5])]
    assert np.array_equal(it["ray_o"].cpu().numpy(), g["batch_ray_o"][5])
    # preprocessed float images: colours pass through bit-exactly
    ds2 = nb.data.DeviceRayDataset(g["all_rgbs"].reshape(ds.N, ds.H, ds.W, 3), g["poses"], focal=float(g["focal"]))
    _, _, c2 = ds2.batch(idx)
    assert np.array_equal(c2.cpu().numpy(), g["batch_rgb"])


def test_full_size_rays_vs_oracle(nb):
    """800x800 view (BASELINE configs[2] size) + a 100-image 400x400 training table (data.py defaults):
    oracle on a slice, size-independent properties on the rest."""
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = q
    pose[:3, 3] = [0.5, -1.0, 4.0]
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911)
    o, d = nb.data.pose_rays(pose, H, W, focal)
    ro, rd = O.dataset_rays(pose[None], H, W, focal)
    assert np.array_equal(o.cpu().numpy(), ro)
    check_dirs(d.cpu().numpy(), rd)
    # shards concatenate to the whole view (the multi-GPU render split)
    lo, hi = nb.dist.shard_range(H * W, 1, 3)
    o1, d1 = nb.data.pose_rays(pose, H, W, focal, lo=lo, hi=hi)
    assert torch.equal(d1, d[lo:hi]) and torch.equal(o1, o[lo:hi])
    # |R^T d| is preserved by the rotation: z component in camera frame is -1
    cam = d.double().cpu().numpy() @ pose[:3, :3].astype(np.float64)
    np.testing.assert_allclose(cam[:, 2], -1.0, atol=1e-6)
    # big virtual table: 100 x 400 x 400 rays, random batch equals the per-view generation
    N, S = 100, 400
    poses = np.tile(np.eye(4, dtype=np.float32), (N, 1, 1))
    poses[:, :3, 3] = rng.standard_normal((N, 3)).astype(np.float32)
    imgs = torch.zeros((N, S, S, 4), dtype=torch.uint8)
    imgs[..., 3] = 255
    imgs[:, :, :, 0] = torch.arange(S, dtype=torch.uint8)[None, None, :]
    ds = nb.data.DeviceRayDataset(imgs, poses, camera_angle_x=0.6911)
    assert len(ds) == 16_000_000
    gen = torch.Generator(DEV).manual_seed(1)
    seen = 0
    for k, (bo, bd, bc) in enumerate(ds.epoch(4096, generator=gen)):
        seen += bo.shape[0]
        if k == 3:
            break
    assert seen == 4 * 4096
    idx = torch.tensor([0, S * S - 1, 57 * S * S + 123 * S + 45, len(ds) - 1], device=DEV)
    bo, bd, bc = ds.batch(idx)
    vo, vd, vc = ds.view_rays(57, with_rgb=True)
    assert torch.equal(bd[2], vd[123 * S + 45]) and torch.equal(bc[2], vc[123 * S + 45]) and torch.equal(bo[2], vo[0])
    ref = O.preprocess_rgba(np.array([[45, 0, 0, 255]], np.uint8))
    np.testing.assert_allclose(bc[2].cpu().numpy(), ref[0], atol=3e-7)


def test_bad_indices_and_errors(nb):
    g = load_golden("data_3x12x12")
    ds = nb.data.DeviceRayDataset(g["rgba"], g["poses"], camera_angle_x=float(g["camera_angle_x"]))
    o, d, c = ds.batch(torch.tensor([len(ds), -1, 0], device=DEV))
    assert torch.isnan(o[:2]).all() and torch.isnan(c[:2]).all() and not torch.isnan(o[2]).any()
    with pytest.raises(RuntimeError):
        ds.batch(torch.tensor([0, 1]))                      # CPU index tensor: no fallback
    with pytest.raises(IndexError):
        ds.view_rays(3)
    o, d, c = ds.batch(torch.empty(0, dtype=torch.int64, device=DEV))
    assert o.shape == (0, 3)


def test_postprocess_uint8(nb):
    g = load_golden("data_3x12x12")
    x = torch.from_numpy(g["pp_in"]).to(DEV)
    for boost in (1.0, 1.5):
        got = nb.data.to_uint8(x, boost, False).cpu().numpy()
        assert np.array_equal(got, g[f"pp_out_b{boost}_g0"])
        got = nb.data.to_uint8(x, boost, True).cpu().numpy().astype(np.int16)
        ref = g[f"pp_out_b{boost}_g1"].astype(np.int16)
        assert np.abs(got - ref).max() <= 1 and np.mean(got != ref) <= 1e-3
    # full-size image (800x800x3) against the oracle, exact without gamma
    big = torch.rand(800, 800, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(2)) * 1.2
    assert np.array_equal(nb.data.to_uint8(big).cpu().numpy(), O.to_uint8(big.cpu().numpy()))
    a, b = nb.data.to_uint8(big, 1.0, True).cpu().numpy().astype(np.int16), O.to_uint8(big.cpu().numpy(), 1.0, True).astype(np.int16)
    assert np.abs(a - b).max() <= 1 and np.mean(a != b) <= 1e-3


def test_render_maps_and_train_from_dataset(nb):
    """End to end across the widened boundary: DeviceRayDataset -> TrainStep -> render_maps -> uint8."""
    g = load_golden("data_3x12x12")
    ds = nb.data.DeviceRayDataset(g["rgba"], g["poses"], camera_angle_x=float(g["camera_angle_x"]))
    torch.manual_seed(0)
    m = nb.NeRFMLP().to(DEV)
    r = nb.NeRFRenderer(m, DEV, perturb=1.0)
    opt = nb.FlatAdam(m, lr=5e-4)
    step = nb.TrainStep(r, opt, 144)
    gen = torch.Generator(DEV).manual_seed(0)
    n = 0
    for _ in range(2):
        for bo, bd, bc in ds.epoch(144, generator=gen, drop_last=True):
            step(bo, bd, bc)
            n += 1
    assert n == 6 and opt._step == 6 and np.isfinite(step.read_metrics()["psnr"])
    r.perturb = 0.0
    o, d, _ = ds.view_rays(1)
    maps = r.render_maps(o, d, ds.H, ds.W, ds.focal, chunk=100)
    img = r.render(o, d, ds.H, ds.W, ds.focal, chunk=100)
    assert torch.equal(maps["rgb_map"], img) and maps["depth_map"].shape == (12, 12) and maps["acc_map_coarse"].shape == (12, 12)
    u8 = nb.data.to_uint8(img, gamma_correction=True)
    assert u8.shape == (12, 12, 3) and u8.dtype == torch.uint8

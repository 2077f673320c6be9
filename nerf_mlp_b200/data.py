"""Device-side replacements for the callers either side of the hot path (SURVEY.md section 8f rows
2 and 4): ray generation + batching (reference nerfmlp/data.py:76-104 + the DataLoader at
scripts/train.py:219,368-371) and output post-processing (scripts/render_example.py:256-271).

The reference's NeRFDataset builds three host tables with one row per ray of every training image
(float64 -> 36 B/ray after .float()) and serves them through a per-item ``__getitem__`` + collate.
``DeviceRayDataset`` keeps only the poses (64 B/image) and the images on the device and rebuilds a
ray from (pose, pixel) inside the gather kernel; a training batch is one launch, an epoch is one
``torch.randperm`` on the device (what ``DataLoader(shuffle=True)`` does on the host).

File I/O (transforms_*.json, PNG decoding, LANCZOS resize: data.py:35-46) stays with the caller:
it is host-side work outside the path; the arrays it yields are what this class takes.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import check, dll, ptr, stream_ptr


class DeviceRayDataset:
    """``images``: uint8 ``[N,H,W,4]`` raw RGBA (preprocessing of data.py:47-62 fused into the gather)
    or float32 ``[N,H,W,3]`` already-linear RGB; ``poses``: ``[N,4,4]`` camera-to-world
    ('transform_matrix' of transforms_*.json); ``focal`` in pixels, or ``camera_angle_x`` to derive
    it as data.py:73 does.  Row order of the virtual ray table = NeRFDataset.all_rays_* (image-major,
    then row, then column), so ``dataset[idx]`` matches the reference's ``__getitem__`` (data.py:99-104).
    """

    def __init__(self, images, poses, focal=None, camera_angle_x=None, white_bkgd=True, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("nerf_mlp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        images = torch.as_tensor(images)
        poses = torch.as_tensor(poses, dtype=torch.float32)
        if images.dim() != 4 or poses.shape != (images.shape[0], 4, 4):
            raise RuntimeError(f"DeviceRayDataset: images [N,H,W,C] / poses [N,4,4] expected, got "
                               f"{tuple(images.shape)} / {tuple(poses.shape)}")
        self.N, self.H, self.W = int(images.shape[0]), int(images.shape[1]), int(images.shape[2])
        if images.dtype == torch.uint8 and images.shape[3] == 4:
            self.rgba, self.rgb_lin = images.contiguous().to(self.device), None
        elif images.shape[3] == 3:
            self.rgba, self.rgb_lin = None, images.to(torch.float32).contiguous().to(self.device)
        else:
            raise RuntimeError("DeviceRayDataset: images must be uint8 RGBA [N,H,W,4] or float RGB [N,H,W,3]")
        if focal is None:
            if camera_angle_x is None:
                raise ValueError("DeviceRayDataset: give focal or camera_angle_x")
            focal = 0.5 * self.W / math.tan(0.5 * float(camera_angle_x))          # data.py:73 (img_wh[0] = W)
        self.focal = float(focal)
        self.white_bkgd = bool(white_bkgd)
        self.poses = poses.contiguous().to(self.device)

    def __len__(self):
        return self.N * self.H * self.W                                           # data.py:96-97

    def _run(self, idx, first, n, want_rgb, out=None):
        f32 = dict(device=self.device, dtype=torch.float32)
        if out is None:
            out = (torch.empty((n, 3), **f32), torch.empty((n, 3), **f32),
                   torch.empty((n, 3), **f32) if want_rgb else None)
        rays_o, rays_d, rgb = out
        _lib.require_cuda(rays_o, rays_d, rgb)
        for t in (rays_o, rays_d) + ((rgb,) if want_rgb else ()):
            if tuple(t.shape) != (n, 3):
                raise RuntimeError(f"DeviceRayDataset: output buffers must be [{n},3], got {tuple(t.shape)}")
        check(dll().nerf_generate_rays(ptr(self.poses), self.N, self.H, self.W, self.focal, ptr(idx), int(first), int(n),
                                       ptr(rays_o), ptr(rays_d), ptr(self.rgba), ptr(self.rgb_lin), int(self.white_bkgd),
                                       ptr(rgb) if want_rgb else None, stream_ptr(self.device)), "nerf_generate_rays")
        return rays_o, rays_d, rgb

    def batch(self, idx, out=None):
        """Rays + target colours of the flat ray ids ``idx`` (int64 CUDA tensor): one launch.
        Returns ``(ray_o, ray_d, rgb)``, the three fields of the reference's batch dict.  ``out`` =
        optional preallocated ``(rays_o, rays_d, rgb)`` (e.g. TrainStep's static buffers)."""
        if not (idx.is_cuda and idx.dtype == torch.int64 and idx.is_contiguous() and idx.dim() == 1):
            raise RuntimeError("DeviceRayDataset.batch: idx must be a contiguous 1-D int64 CUDA tensor (no CPU fallback)")
        return self._run(idx, 0, idx.numel(), True, out)

    def __getitem__(self, i):
        """data.py:99-104 for one ray (a convenience for tests; training uses batch())."""
        o, d, c = self.batch(torch.tensor([int(i)], device=self.device, dtype=torch.int64))
        return {"ray_o": o[0], "ray_d": d[0], "rgb": c[0]}

    def view_rays(self, img, lo=0, hi=None, with_rgb=False):
        """All rays (or the pixel range [lo, hi), e.g. one rank's shard) of training view ``img`` in
        render order (scripts/render_example.py:245-250)."""
        hw = self.H * self.W
        hi = hw if hi is None else hi
        if not (0 <= img < self.N and 0 <= lo <= hi <= hw):
            raise IndexError("DeviceRayDataset.view_rays: view or pixel range out of bounds")
        return self._run(None, img * hw + lo, hi - lo, with_rgb)

    def epoch(self, batch_size, generator=None, drop_last=False):
        """Iterate one shuffled epoch: the device-side equivalent of
        ``DataLoader(dataset, batch_size, shuffle=True)`` (scripts/train.py:219)."""
        perm = torch.randperm(len(self), device=self.device, generator=generator)
        for i in range(0, len(self), batch_size):
            idx = perm[i:i + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            yield self.batch(idx.contiguous())


def pose_rays(pose, H, W, focal, device="cuda", lo=0, hi=None):
    """rays_o, rays_d of one camera pose (4x4 camera-to-world) for pixels [lo, hi) in row-major
    order -- scripts/render_example.py:245-250 without the host meshgrid."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("nerf_mlp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    hi = H * W if hi is None else hi
    pose = torch.as_tensor(pose, dtype=torch.float32).reshape(1, 4, 4).contiguous().to(device)
    n = hi - lo
    o = torch.empty((n, 3), device=device, dtype=torch.float32)
    d = torch.empty((n, 3), device=device, dtype=torch.float32)
    check(dll().nerf_generate_rays(ptr(pose), 1, int(H), int(W), float(focal), None, int(lo), int(n), ptr(o), ptr(d),
                                   None, None, 0, None, stream_ptr(device)), "nerf_generate_rays")
    return o, d


def to_uint8(rgb, brightness=1.0, gamma_correction=False):
    """scripts/render_example.py:256-271 on the device: ``rgb * brightness`` -> optional linear->sRGB
    (:12-26) -> clip [0,1] -> *255 -> uint8 (truncation).  Any shape; returns a uint8 CUDA tensor."""
    rgb = _lib.f32c(rgb)
    out = torch.empty(rgb.shape, device=rgb.device, dtype=torch.uint8)
    check(dll().nerf_postprocess_rgb8(ptr(rgb), rgb.numel(), float(brightness), int(bool(gamma_correction)), ptr(out),
                                      stream_ptr(rgb.device)), "nerf_postprocess_rgb8")
    return out

"""Multi-GPU plumbing (new relative to the reference, which is single-process): one process per
GPU, rays sharded in contiguous blocks.

  * rendering is embarrassingly parallel: no collective on the data path, only an optional
    all_gather of the 12 B/ray rgb_map to assemble the image;
  * training is data-parallel: one all-reduce of the flat 595 844-float gradient buffer per step
    (done inside FlatAdam.step).

Everything here is device-agnostic index arithmetic + torch.distributed calls, so it is covered by
world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as td


def shard_range(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n` items owned by `rank`; the first n % world ranks get one
    extra item so block sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _rank_world(group=None):
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


def render_sharded(render_fn, rays_o, rays_d, H, W, focal, chunk=1024 * 16, group=None, gather=True):
    """Shard `render` over the ranks of `group`: each rank renders its contiguous block of rays with
    `render_fn(rays_o_blk, rays_d_blk) -> [n_blk, 3]` (e.g. a closure over NeRFRenderer._render_rays
    chunks) and, if `gather`, every rank receives the full (H, W, 3) image.

    render_fn is passed explicitly so that the sharding logic is testable without a GPU."""
    rank, world = _rank_world(group)
    n = rays_o.shape[0]
    if n != H * W:
        raise RuntimeError(f"render_sharded: N_rays ({n}) != H*W ({H * W})")   # same contract as renderer.py:45
    lo, hi = shard_range(n, rank, world)
    parts = []
    for i in range(lo, hi, chunk):
        j = min(i + chunk, hi)
        parts.append(render_fn(rays_o[i:j], rays_d[i:j]))
    local = torch.cat(parts, 0) if parts else rays_o.new_zeros((0, 3))
    if world == 1 or not gather:
        return local.view(H, W, 3) if world == 1 else local
    # all_gather needs equal sizes: pad to the largest block
    max_blk = -(-n // world)
    pad = local.new_zeros((max_blk, 3))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    td.all_gather(bufs, pad, group=group)
    out = []
    for r in range(world):
        l, h = shard_range(n, r, world)
        out.append(bufs[r][: h - l])
    return torch.cat(out, 0).view(H, W, 3)


def make_render_fn(renderer):
    """render_fn for render_sharded from a NeRFRenderer (fine rgb_map only, no grad; renderer.py:40-44)."""
    def fn(o, d):
        with torch.no_grad():
            return renderer._render_rays(o, d)["rgb_map"]
    return fn


def broadcast_params(model, src=0, group=None):
    """Make every rank start from rank `src`'s weights (one flat broadcast)."""
    rank, world = _rank_world(group)
    if world > 1:
        model._ensure_flat()
        td.broadcast(model.flat_params, src=src, group=group)
        model.mark_dirty()


def allreduce_flat(flat, group=None):
    """SUM all-reduce of one flat buffer (the DP gradient exchange when not using FlatAdam)."""
    rank, world = _rank_world(group)
    if world > 1:
        td.all_reduce(flat, op=td.ReduceOp.SUM, group=group)
    return world


def peer_exchange_layout(n_params: int, world: int, two_shot=None, peer_max: int = 8):
    """Layout of one rank's symmetric block for TrainStep's peer-memory gradient exchange
    (nerf_adam_step_fused_peer): ``{"two_shot", "n_pad", "grad_off", "red_off", "flag_off", "floats"}`` in floats.
    One-shot (every rank reads all gradients) below 4 ranks, two-shot (reduce-scatter + all-gather through a second
    n-float region) from 4 ranks up; ``two_shot`` (or the environment variable NERF_PEER_TWO_SHOT=0|1) overrides."""
    import os
    if not 1 <= world <= peer_max:
        raise ValueError(f"peer-memory gradient exchange supports 1..{peer_max} ranks, got {world}")
    if two_shot is None:
        env = os.environ.get("NERF_PEER_TWO_SHOT", "")
        two_shot = (world >= 4) if env == "" else (env != "0")
    n_pad = (int(n_params) + 127) // 128 * 128                    # 512-byte granules: every region stays 16-byte aligned
    n_data = (2 if two_shot else 1) * n_pad
    n_flags = 2 * peer_max + 32                                   # ready flags | slice / read-done flags | epoch, counter
    return {"two_shot": bool(two_shot), "n_pad": n_pad, "grad_off": 0, "red_off": n_pad if two_shot else None,
            "flag_off": n_data, "floats": n_data + n_flags}


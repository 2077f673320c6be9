"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
the drop-in classes mirror the reference's surface, and the multi-GPU host logic (ray sharding,
gather, flat all-reduce) works under world_size-2 gloo."""
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import nerf_mlp_b200 as nb
from oracle import nerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 17
    dll = nb._lib.dll()
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/nerf_b200.h but not exported"
    assert sorted(nb._lib.SIGNATURES) == syms, "ctypes SIGNATURES out of sync with the header"
    assert b"sm_100a" in dll.nerf_version()
    assert dll.nerf_packed_weight_bytes() > 1_000_000
    # workspace sizing is pure host arithmetic
    assert dll.nerf_mlp_workspace_bytes(1024, nb._lib.PREC_FP32, 0) == 1024 * 2522 * 4
    assert dll.nerf_mlp_workspace_bytes(0, nb._lib.PREC_BF16, 1) == 0


def test_argument_checks_do_not_need_a_gpu():
    dll = nb._lib.dll()
    rc = dll.nerf_composite_fwd(None, None, None, None, 4, 0, 1, None, None, None, None, None)
    assert rc != 0 and b"bad shape" in dll.nerf_last_error()
    rc = dll.nerf_mlp_fwd_rays(None, None, None, 4, 64, 1.0, None, None, None, None, 0, 7, 0, None)
    assert rc != 0 and b"precision" in dll.nerf_last_error()


def test_train_step_has_no_cpu_fallback():
    m = nb.NeRFMLP()
    r = nb.NeRFRenderer(m, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.TrainStep(r, nb.FlatAdam(m), 16)
    dll = nb._lib.dll()
    assert dll.nerf_train_prepare(None, None, None, 0, None) != 0 and b"nerf_train_prepare" in dll.nerf_last_error()


def test_model_surface_matches_reference():
    torch.manual_seed(0)
    m = nb.NeRFMLP()
    assert (m.D, m.W, m.input_ch, m.input_ch_views, m.skips, m.use_viewdirs) == (8, 256, 63, 27, [5], True)
    sd = m.state_dict()
    assert list(sd) == list(O.PARAM_NAMES)
    ref = O.init_params(0)
    assert all(tuple(sd[k].shape) == ref[k].shape for k in sd)
    assert sum(p.numel() for p in m.parameters()) == 595844
    # default init follows nn.Linear (U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
    assert float(sd["pts_linears.5.weight"].abs().max()) <= 1 / np.sqrt(319) + 1e-6
    # flat views survive load_state_dict; a fresh flat buffer is built by .to()/_apply
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ref.items()})
    assert np.array_equal(m.flat_params.numpy(), O.flatten_params(ref))
    m2 = m.to(torch.device("cpu"))
    assert m2 is m and np.array_equal(m.flat_params.numpy(), O.flatten_params(ref))
    assert m.sigma_linear.weight.data_ptr() == m.flat_params.data_ptr() + 4 * (595844 - 1 - 256 - 65792 - 36352 - 387)
    with pytest.raises(RuntimeError):
        m.double()


def test_load_from_numpy_order():
    p = O.init_params(5)
    arrs = []
    for i in range(8):
        arrs += [p[f"pts_linears.{i}.weight"].T.copy(), p[f"pts_linears.{i}.bias"]]
    for n in ("bottleneck_linear", "view_linear", "rgb_linear", "sigma_linear"):
        arrs += [p[f"{n}.weight"].T.copy(), p[f"{n}.bias"]]
    m = nb.NeRFMLP()
    m.load_from_numpy(arrs)
    assert np.array_equal(m.flat_params.numpy(), O.flatten_params(p))


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        nb.NeRFMLP(D=4)
    with pytest.raises(NotImplementedError):
        nb.NeRFMLP(use_viewdirs=False)
    with pytest.raises(ValueError):
        nb.NeRFMLP(precision="fp8")
    with pytest.raises(TypeError):
        nb.NeRFRenderer(torch.nn.Linear(3, 3), "cpu")
    m = nb.NeRFMLP()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(2, 63))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 63), torch.zeros(2, 27))
    r = nb.NeRFRenderer(m, "cpu", N_samples=32, N_importance=16, perturb=0.0)
    assert (r.N_samples, r.N_importance, r.near, r.far, r.white_bkgd, r.perturb, r.raw_noise_std, r.coord_scale) == \
        (32, 16, 2.0, 6.0, True, 0.0, 0.0, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r.render(torch.zeros(4, 3), torch.ones(4, 3), 2, 2, 1.0)


def test_positional_encoding_attributes():
    pe = nb.PositionalEncoding(10)
    assert pe.num_freqs == 10 and pe.include_input and pe.log_sampling
    assert np.array_equal(pe.freq_bands.numpy(), 2.0 ** np.arange(10, dtype=np.float32))
    assert len(pe.state_dict()) == 0
    lin = nb.PositionalEncoding(4, log_sampling=False)
    np.testing.assert_allclose(lin.freq_bands.numpy(), np.linspace(1, 8, 4))


def test_shard_range_partitions():
    for n in (0, 1, 7, 640000, 10007):
        for world in (1, 2, 3, 8):
            blocks = [nb.dist.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        nb.dist.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W = 7, 9                                    # 63 rays: uneven split 32 / 31
        o = torch.arange(H * W * 3, dtype=torch.float32).view(-1, 3)
        d = -o
        calls = []

        def render_fn(ob, db):                        # stand-in for the GPU renderer: a per-ray function
            calls.append(ob.shape[0])
            return ob * 2 + db * 0.5 + 1

        img = nb.dist.render_sharded(render_fn, o, d, H, W, 1.0, chunk=10)
        ok_img = torch.equal(img, (o * 2 + d * 0.5 + 1).view(H, W, 3))
        lo, hi = nb.dist.shard_range(H * W, rank, world)
        ok_calls = sum(calls) == hi - lo and max(calls) <= 10
        local = nb.dist.render_sharded(render_fn, o, d, H, W, 1.0, chunk=100, gather=False)
        ok_local = local.shape[0] == hi - lo
        # flat gradient exchange: SUM all-reduce, then 1/world scaling is applied by the optimiser
        flat = torch.full((1000,), float(rank + 1))
        w = nb.dist.allreduce_flat(flat)
        ok_ar = w == world and torch.equal(flat, torch.full((1000,), float(sum(range(1, world + 1)))))
        # broadcast of flat parameters from rank 0
        torch.manual_seed(rank)
        m = nb.NeRFMLP()
        nb.dist.broadcast_params(m, src=0)
        torch.manual_seed(0)
        ok_bc = torch.equal(m.flat_params, nb.NeRFMLP().flat_params)
        q.put((rank, ok_img, ok_calls, ok_local, ok_ar, ok_bc))
    finally:
        td.destroy_process_group()


def test_multi_rank_host_logic_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert all(r[1:]), r


# ---- checkpoint formats ---------------------------------------------------
def _adam_steps(opt, model, seed, n):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n):
        for p in model.parameters():
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
        opt.step()


def test_optimizer_state_interchanges_with_torch_adam(tmp_path):
    """A checkpoint written by the reference's loop (torch.optim.Adam.state_dict(), scripts/train.py:471-475)
    resumes in FlatAdam, and FlatAdam's state_dict loads into torch.optim.Adam."""
    torch.manual_seed(1)
    m_ref = nb.NeRFMLP()
    opt_ref = torch.optim.Adam(m_ref.parameters(), lr=5e-4)
    _adam_steps(opt_ref, m_ref, 3, 2)
    path = tmp_path / "metrics_latest.pth"
    torch.save({"model_state_dict": m_ref.state_dict(), "optimizer_state_dict": opt_ref.state_dict(),
                "metrics": {"step": 2, "best_val_psnr": 12.5}}, path)
    m = nb.NeRFMLP()
    opt = nb.FlatAdam(m, lr=1e-3)
    metrics = nb.checkpoint.load_checkpoint(path, m, opt, map_location="cpu")
    assert metrics["step"] == 2 and opt._step == 2 and opt.param_groups[0]["lr"] == 5e-4
    assert torch.equal(m.flat_params, m_ref.flat_params)
    for i, (mv, vv) in enumerate(zip(m._views_of(opt._m), m._views_of(opt._v))):
        st = opt_ref.state[opt_ref.param_groups[0]["params"][i]]
        assert torch.equal(mv, st["exp_avg"]) and torch.equal(vv, st["exp_avg_sq"])
    # and back: FlatAdam -> file -> torch.optim.Adam continues identically to the uninterrupted run
    path2 = tmp_path / "from_flat.pth"
    nb.checkpoint.save_checkpoint(path2, m, opt, {"step": 2})
    ck = torch.load(path2)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "metrics"}
    assert list(ck["model_state_dict"]) == list(O.PARAM_NAMES)
    m2 = nb.NeRFMLP()
    m2.load_state_dict(ck["model_state_dict"])
    opt2 = torch.optim.Adam(m2.parameters(), lr=1.0)
    opt2.load_state_dict(ck["optimizer_state_dict"])
    _adam_steps(opt_ref, m_ref, 9, 1)
    _adam_steps(opt2, m2, 9, 1)
    assert torch.equal(m2.flat_params, m_ref.flat_params)
    # unsupported Adam variants are refused, not silently ignored
    bad = opt_ref.state_dict()
    bad["param_groups"][0]["weight_decay"] = 0.1
    with pytest.raises(NotImplementedError):
        opt.load_state_dict(bad)


def test_fresh_flat_adam_state_loads_into_torch_adam():
    """A FlatAdam that never loaded a torch checkpoint writes param_groups with EVERY torch.optim.Adam key, so the
    reference's optimizer (scripts/train.py:303-306: optimizer.load_state_dict) can resume from it and step --
    both before the first step (empty state) and after some steps (moments carried over)."""
    ref_keys = set(torch.optim.Adam(torch.nn.Linear(2, 2).parameters()).state_dict()["param_groups"][0])
    torch.manual_seed(3)
    m = nb.NeRFMLP()
    opt = nb.FlatAdam(m, lr=5e-4)
    sd0 = opt.state_dict()
    assert set(sd0["param_groups"][0]) == ref_keys and sd0["state"] == {}
    m_t = nb.NeRFMLP()
    m_t.load_state_dict(m.state_dict())
    opt_t = torch.optim.Adam(m_t.parameters(), lr=1.0)
    opt_t.load_state_dict(sd0)                          # raised KeyError('weight_decay') on step() before the fix
    _adam_steps(opt_t, m_t, 5, 1)
    assert opt_t.param_groups[0]["lr"] == 5e-4
    # an empty-state checkpoint loaded back into FlatAdam leaves it steppable (moments re-created lazily)
    opt.load_state_dict(sd0)
    assert opt._step == 0 and opt._m is None
    opt._ensure_moments()
    assert opt._m is not None and float(opt._m.abs().sum()) == 0.0


def test_weight_files(tmp_path):
    """.pth state_dict and the official .npy weight list (render_example.py:166-208, model.py:83-127)."""
    p = O.init_params(4)
    m = nb.NeRFMLP()
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    nb.checkpoint.save_weights(m, tmp_path / "model_best.pth")
    m2 = nb.checkpoint.load_weights(nb.NeRFMLP(), tmp_path / "model_best.pth", map_location="cpu")
    assert torch.equal(m2.flat_params, m.flat_params) and m2._dirty
    # official list order: 8 trunk pairs, bottleneck, view, rgb, sigma; arrays are [in, out]
    names = [f"pts_linears.{i}" for i in range(8)] + ["bottleneck_linear", "view_linear", "rgb_linear", "sigma_linear"]
    arrs = []
    for n in names:
        arrs += [p[n + ".weight"].T.copy(), p[n + ".bias"].copy()]
    obj = np.empty(len(arrs), dtype=object)
    for i, a in enumerate(arrs):
        obj[i] = a
    np.save(tmp_path / "model_fine.npy", obj, allow_pickle=True)
    m3 = nb.checkpoint.load_weights(nb.NeRFMLP(), str(tmp_path / "model_fine.npy"))
    assert torch.equal(m3.flat_params, m.flat_params)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): exactly one JSON line on stdout
    with the contract's keys; needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--rays", "128"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "train_rays_per_s" and d["unit"] == "rays/s" and d["value"] > 0
    # oracle/_ref (the unmodified reference, compiled by oracle/build_ref.py) when it is built, else the bit-identical port
    from oracle import build_ref
    kind = "reference" if build_ref.available()[0] else "port"
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1
    assert ("unmodified reference" if kind == "reference" else "torch CPU restatement") in d["cpu_baseline"]["sample"]
    assert d["config"]["rays_per_gpu"] == 128 and d["steps"] == 1 and d["warmup"] == 0      # --steps / --warmup are honoured
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # the forced port arm still works (what a box without oracle/_ref runs)
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--rays", "64",
                          "--cpu-port", "torch"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    d2 = json.loads(res.stdout.strip())
    assert d2["cpu_baseline"]["kind"] == "port" and "torch CPU restatement" in d2["cpu_baseline"]["sample"]
    assert {k: v for k, v in d2["config"].items() if k not in ("workload", "rays_per_gpu")} == \
           {k: v for k, v in d["config"].items() if k not in ("workload", "rays_per_gpu")}


def test_oracle_is_only_reachable_from_checker_and_baseline_code():
    """The oracle is test infrastructure: nothing under nerf_mlp_b200/ may import it, and bench.py may execute it only
    in its CPU legs (cpu_step_fn -> cpu_baseline / --impl reference) and in the separately reported torch-eager
    baseline -- never inside main()'s product path."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "nerf_mlp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f.endswith((".cu", ".cuh")) and "import" not in txt, os.path.join(dirpath, f)
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    allowed = {"cpu_step_fn", "cpu_kind", "time_torch_eager_gpu"}
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mod = getattr(node, "module", None) or ""
            assert not mod.startswith("oracle") and all(not a.name.startswith("oracle") for a in node.names)
        if isinstance(node, ast.FunctionDef):
            uses = [n for n in ast.walk(node) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
            assert not uses or node.name in allowed, node.name


def test_peer_exchange_layout(monkeypatch):
    """Host side of TrainStep's peer-memory gradient exchange: one-shot below 4 ranks, two-shot from 4 up (second
    n-float region), every region 16-byte aligned, the flag block behind the data, the environment override."""
    from nerf_mlp_b200 import _lib
    from nerf_mlp_b200.dist import peer_exchange_layout
    monkeypatch.delenv("NERF_PEER_TWO_SHOT", raising=False)
    n = 595_844                                                    # the reference network's parameter count
    for world in range(1, 9):
        lay = peer_exchange_layout(n, world, peer_max=_lib.PEER_MAX)
        assert lay["two_shot"] == (world >= 4)
        assert lay["n_pad"] >= n and lay["n_pad"] % 128 == 0 and lay["grad_off"] == 0
        assert lay["red_off"] == (lay["n_pad"] if world >= 4 else None)
        assert lay["flag_off"] == (2 if world >= 4 else 1) * lay["n_pad"]
        assert lay["floats"] - lay["flag_off"] >= 2 * _lib.PEER_MAX + 2      # ready | slice flags, epoch, block counter
    monkeypatch.setenv("NERF_PEER_TWO_SHOT", "1")
    assert peer_exchange_layout(n, 2)["two_shot"] is True
    monkeypatch.setenv("NERF_PEER_TWO_SHOT", "0")
    assert peer_exchange_layout(n, 8)["two_shot"] is False
    assert peer_exchange_layout(n, 8, two_shot=True)["two_shot"] is True     # the explicit argument wins
    with pytest.raises(ValueError):
        peer_exchange_layout(n, 9)

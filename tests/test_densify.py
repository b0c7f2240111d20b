"""Densification: split_and_prune with its kernels D2 classify_gaussians and D3
build_densify_output_map (Trainer/GaussianTrainer.swift:344-427, 766-908).

CPU part: the C port against the reference's own inline Metal kernels compiled for the host (oracle/_ref) and the
host logic (guards, counts).  GPU part (-m gpu): libgsb.so through the C ABI against the oracle — bit-exact actions,
counts, offsets, gather indices and noise modes; bit-exact gathered tensors; positions within f32 rounding of the
device expf (<= 2 ulp) when the same base noise is supplied.
"""
import numpy as np
import pytest

from gaussiansplattingmlx_b200.scene import make_gaussians, make_cameras, make_targets
from oracle import pipeline as pl


def densify_case(n=6000, seed=5):
    params = make_gaussians(n, seed, 3)
    rng = np.random.default_rng(seed + 100)
    accum = (rng.random(n) * 0.02).astype(np.float32)          # avg grad = accum / denom straddles 2e-4 for denom 60
    params["_opacity"][: n // 20] = -8.0                        # sigmoid < 0.005: pruned
    params["_opacity"][n // 20: n // 10] = np.float32(np.log(0.005 / 0.995))   # right at the threshold
    noise = rng.standard_normal((2 * n, 3)).astype(np.float32)
    return params, accum, 60, noise


def test_port_matches_reference_inline_kernels(port, ref):
    params, accum, denom, noise = densify_case()
    for allow in (True, False):
        a_p, c_p = port.classify_gaussians(accum, denom, params["_scales"], params["_opacity"], 0.0002, 0.05, 0.005, allow)
        a_r, c_r = ref.classify_gaussians(accum, denom, params["_scales"], params["_opacity"], 0.0002, 0.05, 0.005, allow)
        assert np.array_equal(a_p, a_r) and np.array_equal(c_p, c_r)
        assert set(np.unique(a_p)) <= {0, 1, 2, 3}
        if allow:
            assert all((a_p == k).any() for k in (0, 1, 2, 3)), "the case must exercise keep, split, clone and prune"
        else:
            assert not ((a_p == 1) | (a_p == 2)).any()
    offsets = np.cumsum(c_p, dtype=np.int64).astype(np.int32) - c_p
    a_p, c_p = port.classify_gaussians(accum, denom, params["_scales"], params["_opacity"], 0.0002, 0.05, 0.005, True)
    offsets = np.cumsum(c_p, dtype=np.int64).astype(np.int32) - c_p
    total = int(c_p.sum())
    g_p, m_p = port.build_densify_output_map(a_p, offsets, total)
    g_r, m_r = ref.build_densify_output_map(a_p, offsets, total)
    assert np.array_equal(g_p, g_r) and np.array_equal(m_p, m_r)
    # D1 through the reference's inline kernel == the port
    g = np.random.default_rng(1).standard_normal((accum.shape[0], 3)).astype(np.float32)
    acc_p, acc_r = accum.copy(), accum.copy()
    port.accum_grad_norm(g, acc_p); ref.accum_grad_norm(g, acc_r)
    assert np.array_equal(acc_p.view(np.uint32), acc_r.view(np.uint32))


def test_split_and_prune_host_logic(port):
    params, accum, denom, noise = densify_case(2000, 9)
    # iteration guard (GaussianTrainer.swift:767)
    new, info = pl.split_and_prune(port, params, accum, denom, 499, noise)
    assert new is None and not info["ran"]
    new, info = pl.split_and_prune(port, params, accum, denom, 15001, noise)
    assert new is None and not info["ran"]
    new, info = pl.split_and_prune(port, params, accum, denom, 600, noise, maxScale=0.05)
    n = params["_xyz"].shape[0]
    assert info["keep"] + info["split"] + info["clone"] + info["prune"] == n
    assert info["total"] == info["keep"] + 2 * info["split"] + 2 * info["clone"] == new["_xyz"].shape[0]
    gather, mode = info["gather"], info["noise_mode"]
    # structure: keep -> one slot mode 0; split -> modes (1,2); clone -> modes (0,3); pruned sources never appear
    assert not np.isin(gather, np.nonzero(info["actions"] == 3)[0]).any()
    src_split = info["actions"][gather] == 1
    assert set(np.unique(mode[src_split])) == {1, 2} and set(np.unique(mode[info["actions"][gather] == 2])) == {0, 3}
    # split children: scale / 1.6 in log space, symmetric offsets; clone originals untouched, copies moved by 0.01 * noise
    red = np.float32(-np.log(1.6))
    assert np.array_equal(new["_scales"][mode == 1], params["_scales"][gather[mode == 1]] + red)
    assert np.array_equal(new["_xyz"][mode == 0], params["_xyz"][gather[mode == 0]])
    d3 = new["_xyz"][mode == 3] - params["_xyz"][gather[mode == 3]]
    assert np.allclose(d3, np.float32(0.01) * noise[: info["total"]][mode == 3], atol=1e-6)
    i1 = np.nonzero(mode == 1)[0]
    off1 = new["_xyz"][i1] - params["_xyz"][gather[i1]]
    mean = np.exp(params["_scales"][gather[i1]]).mean(axis=1, keepdims=True)
    assert np.allclose(off1, mean * 0.1 * noise[i1], rtol=1e-4, atol=1e-7)
    # over budget: prune only (:785); nothing to do at all: unchanged model
    new2, info2 = pl.split_and_prune(port, params, accum, denom, 600, noise, maxGaussians=n)
    assert info2["split"] == 0 and info2["clone"] == 0 and info2["prune"] > 0 and new2["_xyz"].shape[0] == n - info2["prune"]
    p3 = {k: v.copy() for k, v in params.items()}
    p3["_opacity"][:] = 2.0
    new3, info3 = pl.split_and_prune(port, p3, np.zeros(n, np.float32), denom, 600, noise)
    assert new3 is None and info3["ran"] and info3["total"] == n


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gsb():
    from gaussiansplattingmlx_b200.context import Context
    from gaussiansplattingmlx_b200 import _lib
    return Context, _lib


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(6000, 5), (257, 6), (1, 7)])
def test_densify_kernels_vs_oracle(gsb, best_oracle, n, seed):
    Context, L = gsb
    o = best_oracle
    params, accum, denom, noise = densify_case(n, seed)
    ctx = Context(64, 64)
    for allow in (True, False):
        a_o, c_o = o.classify_gaussians(accum, denom, params["_scales"], params["_opacity"], 0.0002, 0.05, 0.005, allow)
        a_g, c_g = ctx.densify_classify(_dev(accum), denom, _dev(params["_scales"]), _dev(params["_opacity"]), 0.0002, 0.05, 0.005, allow)
        assert np.array_equal(a_g.cpu().numpy(), a_o) and np.array_equal(c_g.cpu().numpy(), c_o)
        off_o = np.cumsum(c_o, dtype=np.int64).astype(np.int32) - c_o
        total = int(c_o.sum())
        off_g, g_g, m_g = ctx.densify_map(a_g, c_g)
        assert np.array_equal(off_g.cpu().numpy(), off_o) and g_g.shape[0] == total
        if total == 0:
            continue
        g_o, m_o = o.build_densify_output_map(a_o, off_o, total)
        assert np.array_equal(g_g.cpu().numpy(), g_o) and np.array_equal(m_g.cpu().numpy(), m_o)
        new_o = o.densify_apply(params, g_o, m_o, noise[:total])
        new_g = ctx.densify_apply({k: _dev(v) for k, v in params.items()}, g_g, m_g, base_noise=_dev(noise[:total]))
        for k in ("_features_dc", "_features_rest", "_scales", "_rotation", "_opacity"):
            assert np.array_equal(new_g[k].cpu().numpy().view(np.uint32), new_o[k].view(np.uint32)), k
        # positions: the split offset multiplies mean(exp(scale)) (device expf vs libm: <= 2 ulp apart)
        assert np.abs(new_g["_xyz"].cpu().numpy() - new_o["_xyz"]).max() <= 1e-6
        assert np.array_equal(new_g["_xyz"].cpu().numpy()[m_o == 0], new_o["_xyz"][m_o == 0])
    ctx.close()


@pytest.mark.gpu
def test_densify_generated_noise_is_standard_normal_and_reproducible(gsb):
    Context, L = gsb
    import torch
    n = 200_000
    params = make_gaussians(n, 3, 3)
    ctx = Context(64, 64)
    dp = {k: _dev(v) for k, v in params.items()}
    gather = torch.arange(n, dtype=torch.int32, device="cuda")
    mode = torch.full((n,), 3, dtype=torch.int32, device="cuda")          # clone copies: xyz + 0.01 * noise
    a = ctx.densify_apply(dp, gather, mode, seed=1234)["_xyz"]
    b = ctx.densify_apply(dp, gather, mode, seed=1234)["_xyz"]
    c = ctx.densify_apply(dp, gather, mode, seed=1235)["_xyz"]
    assert torch.equal(a, b) and not torch.equal(a, c)
    z = ((a - dp["_xyz"]) / 0.01).double()
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z ** 3).mean())) < 0.03 and abs(float((z ** 4).mean()) - 3.0) < 0.1   # skewness 0, kurtosis 3
    zc = z - z.mean(0)
    corr = (zc.T @ zc / n).cpu().numpy()
    assert np.abs(corr - np.eye(3)).max() < 0.02                                             # components uncorrelated
    ctx.close()


@pytest.mark.gpu
def test_trainer_densify_vs_oracle_and_keeps_training(gsb, best_oracle):
    """Train, densify on the trainer's own tensors with supplied noise, compare with the oracle's split_and_prune applied
    to the same tensors, then keep training on the new Gaussian count."""
    Context, L = gsb
    import torch
    o = best_oracle
    n, W, H = 3000, 96, 64
    params = make_gaussians(n, 51, 3)
    cams = make_cameras(W, H, 3)
    targets = make_targets(W, H, 3, 51)
    ctx = Context(W, H)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    gc = [L.make_camera(c) for c in cams]
    tg = [torch.from_numpy(t).cuda() for t in targets]
    for it in range(4):
        ctx.train_step(gc, tg, it, 100)
    n0, steps = ctx.trainer_count()
    assert (n0, steps) == (n, 4)
    tt = ctx.trainer_tensors()
    p_host = {k: v.cpu().numpy().copy() for k, v in tt["params"].items()}
    acc_host = tt["accum"].cpu().numpy().copy()
    thr = float(np.median(acc_host) / steps)                     # make roughly half of the Gaussians densify
    noise = np.random.default_rng(2).standard_normal((2 * n, 3)).astype(np.float32)
    new_o, info_o = pl.split_and_prune(o, p_host, acc_host, steps, 600, noise, gradientThreshold=thr, maxScale=0.05)
    info_g = ctx.trainer_densify(thr, 0.05, 0.005, 1_000_000, base_noise=_dev(noise))
    for k in ("keep", "split", "clone", "prune", "total"):
        assert info_g[k] == info_o[k], k
    assert info_g["n"] == info_o["total"] != n and ctx.trainer_count() == (info_o["total"], 0)
    tt = ctx.trainer_tensors()
    for k in ("_features_dc", "_features_rest", "_scales", "_rotation", "_opacity"):
        assert np.array_equal(tt["params"][k].cpu().numpy().view(np.uint32), new_o[k].view(np.uint32)), k
    assert np.abs(tt["params"]["_xyz"].cpu().numpy() - new_o["_xyz"]).max() <= 1e-6
    for k in tt["m"]:
        assert float(tt["m"][k].abs().max()) == 0.0 and float(tt["v"][k].abs().max()) == 0.0 and float(tt["grads"][k].abs().max()) == 0.0
    assert float(tt["accum"].abs().max()) == 0.0
    # training continues on the new count; a second densification without evidence changes nothing but prunes
    l0 = ctx.train_step(gc, tg, 4, 100)
    l1 = ctx.train_step(gc, tg, 5, 100)
    assert np.isfinite(l0) and np.isfinite(l1)
    info2 = ctx.trainer_densify(1e9, 0.05, 0.0, 1_000_000)      # threshold unreachable, nothing pruned
    assert info2["total"] == info2["keep"] == info_g["n"] and ctx.trainer_count() == (info_g["n"], 0)
    ctx.close()


@pytest.mark.gpu
def test_python_trainer_runs_split_and_prune_on_cadence(gsb):
    from gaussiansplattingmlx_b200.model import GaussModel
    from gaussiansplattingmlx_b200.renderer import GaussianRenderer
    from gaussiansplattingmlx_b200.trainer import GaussianTrainer, TrainData
    n, W, H = 1500, 64, 48
    params = make_gaussians(n, 61, 3)
    cams = make_cameras(W, H, 2)
    targets = make_targets(W, H, 2, 61)
    model = GaussModel.from_arrays(params, 3)
    r = GaussianRenderer(active_sh_degree=3, W=W, H=H, TILE_SIZE=(16, 16), whiteBackground=False)
    tr = GaussianTrainer(model, TrainData(cams, targets), r, iterationCount=12, views_per_step=2, seed=3)
    tr.split_and_prune_per_iteration = 5      # GaussianTrainer.swift:1098: this cadence alone drives densify + state reset
    tr.densifyFromIter, tr.densifyUntilIter = 5, 10
    tr.gradientThreshold, tr.maxScale = 1e-7, 0.05
    tr.startTrain(earlyStoppingThreshold=-1.0)
    its = [d["iteration"] for d in tr.densify_log]
    assert its == [5, 10]
    assert tr.densify_log[0]["n"] > n and model._xyz.shape[0] == tr.densify_log[-1]["n"]
    assert model._features_rest.shape == (model._xyz.shape[0], 15, 3)

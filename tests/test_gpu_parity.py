"""GPU parity: every stage of libgsb.so (through the C ABI) against the CPU oracle on seeded scenes.

Tolerances are the ones BASELINE.json's north_star states: tile-intersection counts and sorted key
lists bit-exact, forward pixels <= 1e-4 max-abs, gradients <= 1e-3 relative
(||delta||_inf / max(||g||_inf, 1e-8) per tensor).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gaussiansplattingmlx_b200.scene import make_workload, make_gaussians, make_cameras, make_targets
from oracle import pipeline as pl
from oracle.api import tile_bit_count, ssim_window

PIX_TOL = 1e-4
GRAD_TOL = 1e-3


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-8))


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def u32(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.fixture(scope="module")
def gsb():
    from gaussiansplattingmlx_b200.context import Context
    from gaussiansplattingmlx_b200 import _lib
    return Context, _lib


def make_ctx(gsb, wl_or_wh, **kw):
    Context, _ = gsb
    if isinstance(wl_or_wh, tuple):
        W, H = wl_or_wh
        return Context(W, H, **kw)
    wl = wl_or_wh
    return Context(wl.width, wl.height, tile_w=wl.tile, tile_h=wl.tile, sh_degree=wl.sh_degree, **kw)


SCENES = {
    "C1": dict(n=1000, W=64, H=64, seed=1, degree=3),
    "deg4": dict(n=700, W=80, H=48, seed=11, degree=4),
    "ragged": dict(n=1500, W=100, H=70, seed=12, degree=2),      # image not a multiple of the tile
    "deg0": dict(n=300, W=32, H=32, seed=13, degree=0),
}


def scene(name):
    s = SCENES[name]
    params = make_gaussians(s["n"], s["seed"], s["degree"])
    cam = make_cameras(s["W"], s["H"], 3)[1]
    target = make_targets(s["W"], s["H"], 1, s["seed"])[0]
    return s, params, cam, target


@pytest.mark.parametrize("name", list(SCENES))
def test_stagewise_parity(gsb, best_oracle, name):
    Context, L = gsb
    o = best_oracle
    s, params, cam, target = scene(name)
    W, H, degree = s["W"], s["H"], s["degree"]
    ctx = Context(W, H, sh_degree=degree)
    gcam = L.make_camera(cam)

    # --- activations
    act_o = o.activate_fwd(params)
    dparams = {k: dev(v) for k, v in params.items()}
    act_g = ctx.activate_fwd(dparams)
    for k in ("shs", "scales", "rotations", "opacity"):
        assert rel_err(act_g[k].cpu().numpy(), act_o[k]) < 1e-6, k

    # --- K1 on the oracle's activated tensors: geometry must be BIT-EXACT
    act_d = {k: dev(v) for k, v in act_o.items()}
    proj_o = o.project_fwd(act_o, cam, degree)
    proj_g = ctx.project_fwd(act_d, gcam)
    for k in ("means2d", "depths", "radii", "rectMin", "rectMax", "cov2d", "conic", "color"):
        a, b = proj_g[k].cpu().numpy(), proj_o[k]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"project_fwd {k} not bit-exact (max diff {np.abs(a-b).max()})"

    # --- K3..K8 on identical inputs: bit-exact counts, keys, sorted lists, ranges
    bins_o = o.bin(proj_o, W, H, 16, 16)
    bins_g = ctx.bin({k: dev(v) for k, v in proj_o.items()})
    assert bins_g["M"] == bins_o["M"]
    assert np.array_equal(u32(bins_g["tilesTouched"]), bins_o["tilesTouched"])
    for k in ("keysHigh", "keysLow", "gaussIdx", "sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"):
        assert np.array_equal(u32(bins_g[k]), bins_o[k]), k
    assert np.array_equal(u32(bins_g["tileCounts"]), bins_o["tileCounts"])
    nz = bins_o["tileCounts"] > 0
    assert np.array_equal(u32(bins_g["tileRanges"])[nz], bins_o["tileRanges"][nz])

    # --- K9
    packed = o.pack(proj_o, act_o["opacity"])
    for white in (False, True):
        ctxw = Context(W, H, sh_degree=degree, white_background=white)
        ctxw.bin({k: dev(v) for k, v in proj_o.items()}, read_lists=False)
        fwd_o = o.raster_fwd(packed, bins_o, W, H, 16, 16, white)
        fwd_g = ctxw.raster_fwd(dev(packed))
        assert np.abs(fwd_g["color"].cpu().numpy() - fwd_o["color"]).max() <= PIX_TOL
        assert np.abs(fwd_g["depth"].cpu().numpy() - fwd_o["depth"]).max() <= PIX_TOL * 10  # depth ~ 4 units
        assert np.abs(fwd_g["alpha"].cpu().numpy() - fwd_o["alpha"]).max() <= PIX_TOL
        lc_g, lc_o = u32(fwd_g["lastContrib"]).astype(np.int64), fwd_o["lastContrib"].astype(np.int64)
        assert (np.abs(lc_g - lc_o) <= 1).all() and (lc_g != lc_o).mean() < 0.01

        # --- K10 on the ORACLE's saved forward (isolates the backward)
        rng = np.random.default_rng(5)
        cot = {"color": rng.standard_normal((W * H, 3)).astype(np.float32),
               "depth": rng.standard_normal((W * H, 1)).astype(np.float32) * 0.1,
               "alpha": rng.standard_normal((W * H, 1)).astype(np.float32)}
        g_o = o.raster_bwd(packed, bins_o, W, H, 16, 16, white, cot, fwd_o)
        g_g = ctxw.raster_bwd(dev(packed), {k: dev(v) for k, v in cot.items()}, {k: dev(v) for k, v in fwd_o.items()})
        g_g = g_g.cpu().numpy()
        for lo, hi, nm in ((0, 2, "means2d"), (2, 6, "conic"), (6, 9, "color"), (9, 10, "opacity"), (10, 11, "depth")):
            assert rel_err(g_g[:, lo:hi], g_o[:, lo:hi]) < GRAD_TOL, f"raster_bwd {nm} white={white}"
        ctxw.close()

    # --- K2
    rng = np.random.default_rng(6)
    n = s["n"]
    cotp = {"depths": rng.standard_normal(n).astype(np.float32), "means2d": rng.standard_normal((n, 2)).astype(np.float32),
            "cov2d": rng.standard_normal((n, 4)).astype(np.float32) * 0.0, "color": rng.standard_normal((n, 3)).astype(np.float32),
            "conic": rng.standard_normal((n, 4)).astype(np.float32)}
    gp_o = o.project_bwd(act_o, cam, degree, cotp)
    gp_g = ctx.project_bwd(act_d, gcam, {k: dev(v) for k, v in cotp.items()})
    for k in ("scales", "rotations", "means3d", "shs", "cameraCenterPoint"):
        assert rel_err(gp_g[k].cpu().numpy(), gp_o[k]) < GRAD_TOL, f"project_bwd {k}"
    ctx.close()


def test_project_bwd_with_cov2d_cotangent(gsb, best_oracle):
    Context, L = gsb
    o = best_oracle
    s, params, cam, _ = scene("C1")
    act_o = o.activate_fwd(params)
    ctx = Context(64, 64)
    rng = np.random.default_rng(7)
    n = s["n"]
    cotp = {"depths": rng.standard_normal(n).astype(np.float32), "means2d": rng.standard_normal((n, 2)).astype(np.float32),
            "cov2d": rng.standard_normal((n, 4)).astype(np.float32), "color": rng.standard_normal((n, 3)).astype(np.float32),
            "conic": rng.standard_normal((n, 4)).astype(np.float32)}
    gp_o = o.project_bwd(act_o, cam, 3, cotp)
    gp_g = ctx.project_bwd({k: dev(v) for k, v in act_o.items()}, L.make_camera(cam), {k: dev(v) for k, v in cotp.items()})
    for k in ("scales", "rotations", "means3d", "shs"):
        assert rel_err(gp_g[k].cpu().numpy(), gp_o[k]) < GRAD_TOL, k


@pytest.mark.parametrize("M,tile_bits", [(0, 4), (1, 4), (2, 1), (777, 4), (4096, 8), (4097, 8), (100_000, 13), (1_000_003, 15)])
def test_sort_matches_stable_sort_and_cub(gsb, M, tile_bits):
    Context, _ = gsb
    ctx = Context(64, 64)
    rng = np.random.default_rng(M + tile_bits)
    kh = rng.integers(0, 1 << tile_bits, M, dtype=np.uint32)
    # depths: positive floats with many duplicates so stability is exercised
    kl = (rng.integers(0, 50, M).astype(np.float32) * 0.25 + 0.2).view(np.uint32)
    vals = np.arange(M, dtype=np.uint32)
    order = np.lexsort((kl, kh))   # stable, like the reference's LSD radix sort
    for use_cub in (False, True):
        sh, sl, sv = ctx.sort_tile_keys(dev(kh.view(np.int32)), dev(kl.view(np.int32)), dev(vals.view(np.int32)), tile_bits, use_cub)
        assert np.array_equal(u32(sh), kh[order]), f"keys high (cub={use_cub})"
        assert np.array_equal(u32(sl), kl[order])
        assert np.array_equal(u32(sv), vals[order]), "payload order (stability)"


def test_sort_full_range_keys(gsb):
    """Full 32-bit low words + 15 tile bits: sortedness + permutation checksum at 4M pairs."""
    Context, _ = gsb
    ctx = Context(64, 64)
    M = 4_000_000
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    kh = torch.randint(0, 1 << 15, (M,), device="cuda", dtype=torch.int32, generator=g)
    kl = torch.randint(0, 2 ** 31 - 1, (M,), device="cuda", dtype=torch.int32, generator=g)
    vals = torch.arange(M, device="cuda", dtype=torch.int32)
    sh, sl, sv = ctx.sort_tile_keys(kh, kl, vals, 15)
    key = (sh.to(torch.int64) << 32) | sl.to(torch.int64)
    assert bool((key[1:] >= key[:-1]).all())
    assert int(sv.to(torch.int64).sum()) == M * (M - 1) // 2
    assert bool((kh[sv.long()] == sh).all()) and bool((kl[sv.long()] == sl).all())
    eq = key[1:] == key[:-1]
    assert bool((sv[1:][eq] > sv[:-1][eq]).all()), "ties must keep emission order"


@pytest.mark.parametrize("shape", [(64, 64, 3), (37, 53, 3), (8, 5, 1), (11, 200, 3)])
def test_ssim_parity(gsb, best_oracle, shape):
    Context, _ = gsb
    o = best_oracle
    H, W, Cn = shape
    ctx = Context(max(W, 1), max(H, 1))
    rng = np.random.default_rng(H * W)
    a = rng.random(shape, dtype=np.float32); b = rng.random(shape, dtype=np.float32)
    _, win = ssim_window()
    so = o.ssim_fwd(a, b, win)
    sg = ctx.ssim_fwd(dev(a), dev(b))
    for k in ("ssim", "mu1", "mu2", "sigma1", "sigma2", "sigma12"):
        assert np.abs(sg[k].cpu().numpy().reshape(-1) - so[k]).max() < 2e-5, k
    up = rng.standard_normal(H * W * Cn).astype(np.float32)
    g1o, _ = o.ssim_bwd(up, a, b, win, so)
    g1g = ctx.ssim_bwd(dev(up.reshape(shape)), dev(a), dev(b)).cpu().numpy()
    assert rel_err(g1g, g1o) < GRAD_TOL


def test_loss_fwd_bwd_parity(gsb, best_oracle):
    Context, _ = gsb
    o = best_oracle
    H, W = 48, 72
    ctx = Context(W, H)
    rng = np.random.default_rng(9)
    render = rng.random((H, W, 3), dtype=np.float32); target = rng.random((H, W, 3), dtype=np.float32)
    lo = pl.loss_forward_backward(o, render, target, 0.2)
    loss, cot = ctx.loss_fwd_bwd(dev(render), dev(target), 1.0)
    assert abs(float(loss.item()) - lo["loss"]) < 1e-5
    assert rel_err(cot.cpu().numpy(), lo["cot_render"]) < GRAD_TOL


def test_autograd_through_reference_shaped_renderer(gsb, best_oracle):
    """``GaussianRenderer.forward`` (GaussianRenderer.swift:882-934) is differentiable through two custom functions
    standing where the reference has its CustomFunction{Forward; VJP} pairs (:150-184 tile composite, :576-602 fused
    projection).  torch.autograd through that path (activations as torch ops, like the reference's MLX ops) must give
    the oracle's gradients w.r.t. the activated AND the raw tensors, and agree with the fused ``forward_raw`` path."""
    Context, L = gsb
    from gaussiansplattingmlx_b200.renderer import GaussianRenderer
    o = best_oracle
    s, params, cam, target = scene("deg4")
    W, H, degree = s["W"], s["H"], s["degree"]
    fr, lo, bw = pl.loss_and_grads(o, params, cam, target, degree)
    r = GaussianRenderer(degree, W, H)
    raw = {k: dev(v).requires_grad_(True) for k, v in params.items()}
    act = {"means3d": r.get_xyz_from(raw["_xyz"]), "shs": r.get_features_from(raw["_features_dc"], raw["_features_rest"]),
           "opacity": r.get_opacity_from(raw["_opacity"]), "scales": r.get_scales_from(raw["_scales"]),
           "rotations": r.get_rotation_from(raw["_rotation"])}
    for v in act.values():
        v.retain_grad()
    render, depth, alpha, vis, radii = r.forward(cam, act["means3d"], act["shs"], act["opacity"], act["scales"], act["rotations"])
    assert np.abs(render.detach().cpu().numpy() - fr["render"]).max() <= PIX_TOL
    assert (vis.cpu().numpy() == fr["visibility_filter"]).all()
    (render * dev(lo["cot_render"])).sum().backward()          # VJP with the loss cotangent
    for k in ("means3d", "shs", "scales", "rotations", "opacity"):
        assert rel_err(act[k].grad.cpu().numpy().reshape(bw["g_act"][k].shape), bw["g_act"][k]) < GRAD_TOL, k
    for k in params:
        assert rel_err(raw[k].grad.cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, k
    # the fused path on the same cotangent
    ctx = r.ctx
    ctx.render_forward({k: v.detach() for k, v in raw.items()}, L.make_camera(cam))
    gf = ctx.render_backward(dev(lo["cot_render"]))
    for k in params:
        assert rel_err(raw[k].grad.cpu().numpy(), gf[k].cpu().numpy()) < GRAD_TOL, k
    # depth / alpha outputs are differentiable too (cotDepth / cotAlpha of the composite VJP, :187-226)
    rng = np.random.default_rng(3)
    cd, ca = rng.standard_normal((H, W, 1)).astype(np.float32), rng.standard_normal((H, W, 1)).astype(np.float32)
    bw2 = pl.backward(o, params, cam, degree, fr, lo["cot_render"], cd, ca)
    act2 = {k: v.detach().clone().requires_grad_(True) for k, v in act.items()}
    render2, depth2, alpha2, _, _ = r.forward(cam, act2["means3d"], act2["shs"], act2["opacity"], act2["scales"], act2["rotations"])
    ((render2 * dev(lo["cot_render"])).sum() + (depth2 * dev(cd)).sum() + (alpha2 * dev(ca)).sum()).backward()
    for k in act2:
        assert rel_err(act2[k].grad.cpu().numpy().reshape(bw2["g_act"][k].shape), bw2["g_act"][k]) < GRAD_TOL, ("depth/alpha", k)
    # one saved forward per renderer, like the reference's closure-captured state: a second forward invalidates the first
    act3 = {k: v.detach().clone().requires_grad_(True) for k, v in act.items()}
    ra, *_ = r.forward(cam, act3["means3d"], act3["shs"], act3["opacity"], act3["scales"], act3["rotations"])
    other = make_cameras(W, H, 3)[2]
    r.forward(other, act3["means3d"].detach(), act3["shs"].detach(), act3["opacity"].detach(), act3["scales"].detach(),
              act3["rotations"].detach())
    with pytest.raises(RuntimeError, match="tile lists were rebuilt"):
        ra.sum().backward()


def test_depth_supervised_loss_and_training_path(gsb, best_oracle):
    """lossFn's depth term (GaussianTrainer.swift:693-714, on whenever the dataset has depth, :949): masked L1 with the
    safe weight, its cotangent through the raster / projection backward, and the trainer entry point with depth targets
    on the device and in pinned host memory - against the oracle's restatement."""
    Context, L = gsb
    o = best_oracle
    n, W, H, degree, lam = 900, 80, 56, 3, 0.35
    params = make_gaussians(n, 91, degree)
    cam = make_cameras(W, H, 3)[1]
    target = make_targets(W, H, 1, 91)[0]
    rng = np.random.default_rng(92)
    tdepth = (rng.random((H, W), dtype=np.float32) * 6.0).astype(np.float32)
    mask = rng.random((H, W)) > 0.4
    fr, lo, bw = pl.loss_and_grads(o, params, cam, target, degree, target_depth=tdepth, depth_mask=mask, lambda_depth=lam)
    ctx = Context(W, H, sh_degree=degree)
    gcam = L.make_camera(cam)
    dparams = {k: dev(v) for k, v in params.items()}
    render, depth, alpha, vis, radii = ctx.render_forward(dparams, gcam)
    dmask = dev(mask.astype(np.uint8))
    loss, cot, cot_depth = ctx.loss_fwd_bwd(render, dev(target), 1.0, depth=depth, depth_mask=dmask, target_depth=dev(tdepth),
                                            lambda_depth=lam)
    assert abs(float(loss.item()) - lo["loss"]) < 3e-5 and lo["depth_loss"] > 0.1
    # sign(depth - target) can flip where the two differ by less than the raster tolerance
    near = np.abs(fr["depth"][..., 0] - tdepth) < 1e-3
    d = np.abs(cot_depth.cpu().numpy() - lo["cot_depth"])[..., 0]
    assert d[~near].max() <= 1e-6 * np.abs(lo["cot_depth"]).max() and near.mean() < 0.01
    grads = ctx.render_backward(cot, cot_depth=cot_depth)
    for k, g in grads.items():
        assert rel_err(g.cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, k
    # an all-false mask: weight = max(0, 1e-6), zero loss term, zero cotangent
    l0, _, cd0 = ctx.loss_fwd_bwd(render, dev(target), 1.0, depth=depth, depth_mask=torch.zeros_like(dmask), target_depth=dev(tdepth),
                                  lambda_depth=lam)
    l_rgb, _ = ctx.loss_fwd_bwd(render, dev(target), 1.0)
    assert float(cd0.abs().max()) == 0.0 and abs(float(l0.item()) - float(l_rgb.item())) < 1e-7
    # trainer path: device-resident and pinned-host depth targets give the oracle's loss and gradient block
    for on_host in (False, True):
        tctx = Context(W, H, sh_degree=degree)
        tctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
        put = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()) if on_host else dev
        l = tctx.trainer_accumulate([gcam], [put(target)], grad_scale=1.0, target_depths=[put(tdepth)],
                                    depth_masks=[put(mask.astype(np.uint8))], lambda_depth=lam)
        assert abs(l - lo["loss"]) < 3e-5
        tg = tctx.trainer_tensors()["grads"]
        for k in params:
            assert rel_err(tg[k].cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, (on_host, k)
        tctx.close()
    ctx.close()


def test_adam_bit_exact(gsb, port):
    Context, _ = gsb
    ctx = Context(64, 64)
    n = 1237
    rng = np.random.default_rng(10)
    shapes = [(n, 3), (n, 1, 3), (n, 15, 3), (n, 3), (n, 4), (n, 1)]
    p = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    g = [rng.standard_normal(s).astype(np.float32) * 1e-3 for s in shapes]
    m = [rng.standard_normal(s).astype(np.float32) * 1e-3 for s in shapes]
    v = [(rng.standard_normal(s).astype(np.float32) * 1e-3) ** 2 for s in shapes]
    acc = rng.random(n, dtype=np.float32)
    lrs = pl.learning_rates(7, 100)
    dp, dg, dm, dv = ([dev(x) for x in xs] for xs in (p, g, m, v))
    dacc = dev(acc)
    ctx.adam_step(dp, dg, dm, dv, lrs, dacc)
    port.accum_grad_norm(g[0], acc)
    for i in range(6):
        port.adam(p[i], g[i], m[i], v[i], lrs[i])
        assert np.array_equal(dp[i].cpu().numpy().view(np.uint32), p[i].view(np.uint32)), f"param {i}"
        assert np.array_equal(dm[i].cpu().numpy().view(np.uint32), m[i].view(np.uint32))
        assert np.array_equal(dv[i].cpu().numpy().view(np.uint32), v[i].view(np.uint32))
    assert np.array_equal(dacc.cpu().numpy().view(np.uint32), acc.view(np.uint32))


@pytest.mark.parametrize("name", ["C1", "deg4", "ragged"])
def test_fused_render_and_backward_vs_oracle(gsb, best_oracle, name):
    """gsb_render_forward / gsb_loss_fwd_bwd / gsb_render_backward (raw tensors, fused activations)
    against the oracle's restatement of lossFn + VJPs (GaussianTrainer.swift:634-716)."""
    Context, L = gsb
    o = best_oracle
    s, params, cam, target = scene(name)
    W, H, degree = s["W"], s["H"], s["degree"]
    fr, lo, bw = pl.loss_and_grads(o, params, cam, target, degree)
    ctx = Context(W, H, sh_degree=degree)
    dparams = {k: dev(v) for k, v in params.items()}
    render, depth, alpha, vis, radii = ctx.render_forward(dparams, L.make_camera(cam))
    assert np.abs(render.cpu().numpy() - fr["render"]).max() <= PIX_TOL
    assert np.abs(alpha.cpu().numpy() - fr["alpha"]).max() <= PIX_TOL
    assert np.abs(depth.cpu().numpy() - fr["depth"]).max() <= PIX_TOL * 10
    assert (vis.cpu().numpy() == fr["visibility_filter"]).all()
    # the activation exponential is pinned by convention (common.cuh gsb_expf == the oracle's gso_expf, bit for bit), so
    # the fused path's geometry - radii, pair count, tile lists - equals the oracle's exactly
    assert np.array_equal(radii.cpu().numpy(), fr["radii"])
    st = ctx.stats()
    assert st["pairs_last_view"] == fr["bins"]["M"]
    lists = ctx.bin_read()
    for k in ("sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"):
        assert np.array_equal(u32(lists[k]), fr["bins"][k]), k
    loss, cot = ctx.loss_fwd_bwd(render, dev(target), 1.0)
    assert abs(float(loss.item()) - lo["loss"]) < 2e-5
    grads = ctx.render_backward(cot)
    for k, g in grads.items():
        assert rel_err(g.cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, k
    # accumulate=True adds onto existing buffers
    grads2 = ctx.render_backward(cot, grads={k: v.clone() for k, v in grads.items()}, accumulate=True)
    for k in grads:
        assert rel_err(grads2[k].cpu().numpy(), 2.0 * grads[k].cpu().numpy()) < 1e-5, k


@pytest.mark.parametrize("white,depth_cot,tile", [(False, False, (16, 16)), (True, True, (16, 16)), (False, True, (24, 20))])
def test_segmented_backward_long_lists_vs_oracle(gsb, best_oracle, white, depth_cot, tile):
    """Lists of ~1500 translucent Gaussians per tile: most blocks cross several 256-Gaussian checkpoints, pixels
    terminate inside, at and past segment boundaries.  Fused forward + segmented backward against the oracle, and
    against the whole-block backward (GSB_FLAG_NO_SEGMENTS)."""
    Context, L = gsb
    o = best_oracle
    n, W, H, degree = 6000, 48, 40, 1
    params = make_gaussians(n, 21, degree)
    params["_scales"] = params["_scales"] + np.float32(1.1)        # big footprints: every tile sees most Gaussians
    params["_opacity"] = params["_opacity"] - np.float32(2.0)      # translucent: long lists before T < 1e-4
    cam = make_cameras(W, H, 3)[2]
    rng = np.random.default_rng(5)
    cot = rng.standard_normal((H, W, 3)).astype(np.float32)
    cot_depth = rng.standard_normal((H, W, 1)).astype(np.float32) if depth_cot else None
    cot_alpha = rng.standard_normal((H, W, 1)).astype(np.float32) if depth_cot else None
    tw, th = tile
    fr = pl.render_forward(o, params, cam, degree, tileW=tw, tileH=th, white_bg=white)
    last = fr["fwd"]["lastContrib"].reshape(-1)
    assert last.max() > 8 * 256
    if tile == (16, 16):
        assert last.min() < 2 * 256 and (last % 256 == 0).any()   # incl. termination AT a checkpoint
    bw = pl.backward(o, params, cam, degree, fr, cot, cot_depth, cot_alpha, tileW=tw, tileH=th, white_bg=white)
    ctx = Context(W, H, tile_w=tw, tile_h=th, sh_degree=degree, white_background=white)
    dparams = {k: dev(v) for k, v in params.items()}
    render, depth, alpha, vis, radii = ctx.render_forward(dparams, L.make_camera(cam))
    assert np.abs(render.cpu().numpy() - fr["render"]).max() <= PIX_TOL
    kw = dict(cot_depth=dev(cot_depth), cot_alpha=dev(cot_alpha)) if depth_cot else {}
    g = {k: v.clone() for k, v in ctx.render_backward(dev(cot), **kw).items()}
    for k in g:
        assert rel_err(g[k].cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, k
    ctx.set_flags(L.GSB_FLAG_NO_SEGMENTS)
    ctx.render_forward(dparams, L.make_camera(cam))
    g0 = ctx.render_backward(dev(cot), **kw)
    for k in g:
        assert rel_err(g[k].cpu().numpy(), g0[k].cpu().numpy()) < 1e-4, k


def test_segmented_backward_checkpoint_pool_exhausted(gsb, monkeypatch):
    """With a pool of 7 slots most blocks cannot checkpoint (or stop half way, or get a slot for one of the two a
    first checkpoint needs): the backward must fall back to longer items and give the same gradients."""
    Context, L = gsb
    n, W, H, degree = 6000, 48, 40, 1
    params = make_gaussians(n, 21, degree)
    params["_scales"] = params["_scales"] + np.float32(1.1)
    params["_opacity"] = params["_opacity"] - np.float32(2.0)
    cam = L.make_camera(make_cameras(W, H, 3)[2])
    cot = dev(np.random.default_rng(5).standard_normal((H, W, 3)).astype(np.float32))
    dparams = {k: dev(v) for k, v in params.items()}
    grads = {}
    for slots in (None, "7", "2", "1"):
        if slots is None:
            monkeypatch.delenv("GSB_CKPT_SLOTS", raising=False)
        else:
            monkeypatch.setenv("GSB_CKPT_SLOTS", slots)
        ctx = Context(W, H, sh_degree=degree)
        render, *_ = ctx.render_forward(dparams, cam)
        grads[slots] = ({k: v.clone() for k, v in ctx.render_backward(cot).items()}, render.clone())
        ctx.close()
    for slots in ("7", "2", "1"):
        assert float((grads[slots][1] - grads[None][1]).abs().max()) < 1e-5
        for k in grads[None][0]:
            assert rel_err(grads[slots][0][k].cpu().numpy(), grads[None][0][k].cpu().numpy()) < 1e-4, (slots, k)


def test_train_steps_vs_oracle(gsb, best_oracle, port):
    """Three batched train steps (B = 2 views).  Every iteration is checked in two tight halves instead of through
    parameter deltas (Adam without bias correction turns a first-step gradient into ~3.2 lr sign(g), which amplifies
    last-bit noise of near-zero gradients to O(lr)):
      * the PRE-ADAM gradient block of gsb_trainer_accumulate against the oracle's batch-mean gradient at the SAME
        parameters (<= 1e-3 relative per tensor), and the loss;
      * gsb_trainer_apply against the C port's Adam + accum_grad_norm fed with that very gradient block: bit-exact
        parameters, m, v and D1 accumulator."""
    Context, L = gsb
    o = best_oracle
    n, W, H = 600, 64, 48
    params = make_gaussians(n, 21, 3)
    cams = make_cameras(W, H, 2)
    targets = make_targets(W, H, 2, 21)
    ctx = Context(W, H)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    gcams = [L.make_camera(c) for c in cams]
    tg = [torch.from_numpy(t).pin_memory() for t in targets]
    worst = 0.0
    for it in range(3):
        tt = ctx.trainer_tensors()
        host = {k: tt["params"][k].cpu().numpy().reshape(params[k].shape).copy() for k in params}
        m_h = {k: tt["m"][k].cpu().numpy().copy() for k in params}
        v_h = {k: tt["v"][k].cpu().numpy().copy() for k in params}
        acc_h = tt["accum"].cpu().numpy().copy()
        loss = ctx.trainer_accumulate(gcams, tg, zero_grads=True, grad_scale=0.5)
        g_dev = {k: tt["grads"][k].cpu().numpy().copy() for k in params}
        gsum = {k: np.zeros(params[k].shape, np.float64) for k in params}
        loss_o = 0.0
        for b in range(2):
            _, lo, bw = pl.loss_and_grads(o, host, cams[b], targets[b], 3)
            loss_o += lo["loss"] / 2
            for k in gsum:
                gsum[k] += bw["grads"][k].reshape(gsum[k].shape).astype(np.float64) / 2
        assert abs(loss - loss_o) < 5e-5
        for k in params:
            e = rel_err(g_dev[k].reshape(gsum[k].shape), gsum[k])
            worst = max(worst, e)
            assert e < GRAD_TOL, (it, k, e)
        ctx.trainer_apply(it, 100)
        lrs = pl.learning_rates(it, 100)
        port.accum_grad_norm(np.ascontiguousarray(g_dev["_xyz"].reshape(n, 3)), acc_h)
        for i, k in enumerate(pl.PARAM_ORDER):
            p_h = np.ascontiguousarray(host[k].reshape(g_dev[k].shape))
            port.adam(p_h, np.ascontiguousarray(g_dev[k]), m_h[k], v_h[k], lrs[i])
            assert np.array_equal(tt["params"][k].cpu().numpy().view(np.uint32), p_h.view(np.uint32)), (it, k)
            assert np.array_equal(tt["m"][k].cpu().numpy().view(np.uint32), m_h[k].view(np.uint32)), (it, k)
            assert np.array_equal(tt["v"][k].cpu().numpy().view(np.uint32), v_h[k].view(np.uint32)), (it, k)
        assert np.array_equal(tt["accum"].cpu().numpy().view(np.uint32), acc_h.view(np.uint32)), it
    print(f"train steps: worst pre-Adam gradient error {worst:.2e} relative")


@pytest.mark.parametrize("chunks,phased", [(1, 1), (2, 1), (4, 1), (2, 0)])
def test_step_protocol_on_one_replica_matches_plain_steps(gsb, monkeypatch, chunks, phased):
    """gsb_trainer_step_peers with a world of ONE replica (its own flags, its own slab): the whole device-side protocol -
    flag waits, the projection split into a geometry and a colour kernel around the binning, the last view's projection
    backward in Gaussian chunks, the exchange kernels per chunk and tensor group (geometry first, SH second), the
    announcements - must reproduce gsb_train_step: same losses, same D1 accumulators, Adam deltas within the run-to-run
    noise of float atomics."""
    Context, L = gsb
    monkeypatch.setenv("GSB_PEER_CHUNKS", str(chunks))
    monkeypatch.setenv("GSB_PEER_PHASED", str(phased))
    n, W, H = 5000, 96, 64
    params = make_gaussians(n, 33, 3)
    cams = [L.make_camera(c) for c in make_cameras(W, H, 3)]
    tg = [torch.from_numpy(t).cuda() for t in make_targets(W, H, 3, 33)]
    ref = Context(W, H)
    ref.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    want = [ref.train_step(cams, tg, it, 100) for it in range(3)]
    rt = ref.trainer_tensors()
    ctx = Context(W, H)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    ctx.trainer_peers_import(1, 0, [ctx.trainer_peers_export()])
    got = [ctx.trainer_step_peers(cams, tg, 1.0 / 3, it, 100, want_loss=True) for it in range(3)]
    ctx.trainer_peers_check()
    tt = ctx.trainer_tensors()
    for a, b in zip(got, want):
        assert abs(a - b) < 1e-6
    assert rel_err(tt["accum"].cpu().numpy(), rt["accum"].cpu().numpy()) < 1e-4
    for k in params:
        d_g = tt["params"][k].cpu().numpy() - params[k].reshape(tt["params"][k].shape)
        d_r = rt["params"][k].cpu().numpy() - params[k].reshape(tt["params"][k].shape)
        assert rel_err(d_g, d_r) < 5e-2, k
    # a step without views (B = 0) in the same protocol: zero gradients, Adam still decays m / v identically everywhere
    ctx.trainer_step_peers([], [], 1.0, 3, 100)
    ctx.trainer_peers_check()
    ctx.trainer_peers_close()
    ctx.close(); ref.close()


def test_view_pipeline_matches_serial_and_regrows(gsb):
    """The two-stream view pipeline of gsb_trainer_accumulate (front of view b+1 overlapping the back of view b)
    must give the same steps as the serial schedule, including when the pair buffers overflow mid-batch."""
    Context, L = gsb
    from gaussiansplattingmlx_b200.camera import Camera
    n, W, H = 40000, 160, 96
    params = make_gaussians(n, 41, 3)
    cams = make_cameras(W, H, 5)
    c2w = np.linalg.inv(cams[0].worldViewTransform.astype(np.float64).T)
    c2w[:3, 0] *= -1.0; c2w[:3, 2] *= -1.0                                            # look AWAY from the scene:
    far = Camera(W, H, float(cams[0].focalX), float(cams[0].focalY), c2w)                # everything culled, M = 0
    targets = make_targets(W, H, 5, 41)
    out = {}
    for flags in (0, L.GSB_FLAG_NO_OVERLAP):
        ctx = Context(W, H, flags=flags)
        ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
        tg = [torch.from_numpy(t).cuda() for t in targets]
        losses = [ctx.train_step([L.make_camera(far)], tg[:1], 0, 100)]
        cap0 = ctx.stats()["pair_capacity"]
        losses += [ctx.train_step([L.make_camera(c) for c in cams], tg, it, 100) for it in (1, 2)]
        cap1 = ctx.stats()["pair_capacity"]
        assert cap1 > cap0, "the batch must have outgrown the first sizing (regrow path)"
        tt = ctx.trainer_tensors()
        out[flags] = (losses, {k: v.cpu().numpy().copy() for k, v in tt["params"].items()}, tt["accum"].cpu().numpy().copy())
        ctx.close()
    (la, pa, aa), (lb, pb, ab) = out[0], out[L.GSB_FLAG_NO_OVERLAP]
    for x, y in zip(la, lb):
        assert abs(x - y) < 1e-6
    # Float atomics make two runs of the SAME schedule differ in the last bits of the gradients, and Adam without bias
    # correction turns a first-step gradient into a step of ~3.2 lr * sign(g): measured run-to-run noise of the
    # parameter deltas is up to 2e-2 of the largest delta (tools/dbg/pipeline_noise.py); a schedule bug (a view
    # dropped, a stale buffer) shows up at O(1)
    for k in pa:
        assert rel_err(pa[k] - params[k].reshape(pa[k].shape), pb[k] - params[k].reshape(pb[k].shape)) < 5e-2, k
    assert rel_err(aa, ab) < 1e-4


@pytest.mark.parametrize("W,H,tw,th", [(128, 96, 32, 24), (100, 70, 25, 18), (64, 64, 8, 8)])
def test_app_default_tile_sizes(gsb, best_oracle, W, H, tw, th):
    """The reference app renders with TILE_SIZE = W/4 x H/4 (UI/TrainView.swift:158-163), not 16 x 16: tile lists,
    forward and backward must match the oracle for tiles larger (covered by several 16 x 16 blocks) and smaller
    than the raster block."""
    Context, L = gsb
    o = best_oracle
    n, degree = 900, 3
    params = make_gaussians(n, 71, degree)
    cam = make_cameras(W, H, 3)[2]
    target = make_targets(W, H, 1, 71)[0]
    fr, lo, bw = pl.loss_and_grads(o, params, cam, target, degree, 0.2, tw, th)
    ctx = Context(W, H, tile_w=tw, tile_h=th, sh_degree=degree)
    gcam = L.make_camera(cam)
    # tile lists on the oracle's projection: bit-exact
    bins_o = fr["bins"]
    bins_g = ctx.bin({k: dev(v) for k, v in fr["proj"].items()})
    assert bins_g["M"] == bins_o["M"]
    for k in ("tilesTouched", "sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx", "tileCounts"):
        assert np.array_equal(u32(bins_g[k]), bins_o[k]), k
    # fused path
    render, depth, alpha, vis, radii = ctx.render_forward({k: dev(v) for k, v in params.items()}, gcam)
    assert np.abs(render.cpu().numpy() - fr["render"]).max() <= PIX_TOL
    assert np.abs(alpha.cpu().numpy() - fr["alpha"]).max() <= PIX_TOL
    loss, cot = ctx.loss_fwd_bwd(render, dev(target), 1.0)
    assert abs(float(loss.item()) - lo["loss"]) < 2e-5
    grads = ctx.render_backward(cot)
    for k, g in grads.items():
        assert rel_err(g.cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, k
    ctx.close()


def test_empty_and_culled_scenes(gsb):
    Context, L = gsb
    cam = make_cameras(32, 32, 1)[0]
    ctx = Context(32, 32)
    params = make_gaussians(64, 3, 3)
    params["_xyz"][:] = np.array([0, 0, -100.0], np.float32)   # everything behind the camera
    render, depth, alpha, vis, radii = ctx.render_forward({k: dev(v) for k, v in params.items()}, L.make_camera(cam))
    assert float(render.abs().max()) == 0.0 and float(alpha.abs().max()) == 0.0 and not bool(vis.any())
    assert ctx.stats()["pairs_last_view"] == 0
    grads = ctx.render_backward(torch.ones_like(render))
    assert all(float(g.abs().max()) == 0.0 for g in grads.values())


def test_error_paths(gsb):
    Context, L = gsb
    from gaussiansplattingmlx_b200._lib import GsbError
    with pytest.raises(GsbError):
        Context(0, 64)
    with pytest.raises(GsbError):
        Context(64, 64, sh_degree=5)
    ctx = Context(32, 32)
    with pytest.raises(GsbError):   # backward without forward
        ctx._saved_params = {k: torch.zeros(1, device="cuda") for k in ("_xyz", "_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")}
        ctx.render_backward(torch.zeros(32, 32, 3, device="cuda"))


# ------------------------------------------------------------------------------------------------
# committed golden fixtures (outputs of the reference's own kernels, tests/golden/make_golden.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "deg4_ragged"])
def test_cuda_path_vs_reference_golden(gsb, port, name):
    from test_oracle_pinning import load_case
    Context, L = gsb
    g, params, cam, target, degree = load_case(name)
    n, W, H = (int(x) for x in g["meta"][:3])
    ctx = Context(W, H, sh_degree=degree)
    gcam = L.make_camera(cam)
    act = port.activate_fwd(params)                     # MLX-op activations: host-side inputs of K1
    proj = ctx.project_fwd({k: dev(v) for k, v in act.items()}, gcam)
    for k in ("means2d", "depths", "radii", "conic", "color"):
        assert np.array_equal(proj[k].cpu().numpy().view(np.uint32), g[k].view(np.uint32)), k
    bins = ctx.bin(proj)
    assert bins["M"] == int(g["M"][0])
    for k in ("tilesTouched", "tileCounts", "sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"):
        assert np.array_equal(u32(bins[k]), g[k]), k
    # fused path against the golden image, loss and gradients
    render, depth, alpha, vis, radii = ctx.render_forward({k: dev(v) for k, v in params.items()}, gcam)
    assert np.abs(render.cpu().numpy() - g["render"]).max() <= PIX_TOL
    assert np.abs(alpha.cpu().numpy() - g["alpha"]).max() <= PIX_TOL
    loss, cot = ctx.loss_fwd_bwd(render, dev(target), 1.0)
    assert abs(float(loss.item()) - float(g["loss"][0])) < 2e-5 * max(1.0, float(g["loss"][0]))
    grads = ctx.render_backward(cot)
    for k, gr in grads.items():
        assert rel_err(gr.cpu().numpy().reshape(g["grad" + k].shape), g["grad" + k]) < GRAD_TOL, k


# ------------------------------------------------------------------------------------------------
# BASELINE.json's own sizes against the reference-kernel oracle (one view each: ~3 s / ~9 s of CPU on the GPU box)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wl_name", ["C2", "C3"])
def test_baseline_sizes_vs_reference_oracle(gsb, best_oracle, wl_name):
    """C2 (300 k Gaussians, 800x800) and C3 (1 M Gaussians, 1080p), seed / view 0 of SURVEY.md 8d, through the oracle
    (the reference's shipped kernels compiled for the host when oracle/_ref is present):
      * stage API on the oracle's projection: M, tiles-touched, tile counts, sorted (tile, depth) keys and Gaussian
        lists BIT-EXACT (slang/gaussian_tile_global_kernels.slang:17-404);
      * fused product path (raw tensors in): the same M, radii and sorted lists bit-exact (activation exp pinned by
        convention), pixels <= 1e-4 max-abs (:523-614) wherever both sides terminate a pixel at the same Gaussian - the
        few pixels (< 1e-3 of the image, counted and printed) whose transmittance crosses the 1e-4 cut within rounding of
        it are held to the bound of that discontinuity, 1e-4 x (1 + max colour) -, loss, and all six gradient tensors
        <= 1e-3 relative (:648-881, gaussian_projection_kernels.slang:205-398).
    The measured errors are printed (pytest -s / the GPU test log)."""
    Context, L = gsb
    o = best_oracle
    wl, params, cams, targets = make_workload(wl_name, views_override=1)
    N = params["_xyz"].shape[0]
    W, H, degree = wl.width, wl.height, wl.sh_degree
    fr, lo, bw = pl.loss_and_grads(o, params, cams[0], targets[0], degree, 0.2, wl.tile, wl.tile)
    bo = fr["bins"]
    ctx = Context(W, H, tile_w=wl.tile, tile_h=wl.tile, sh_degree=degree, max_gaussians=N)
    gcam = L.make_camera(cams[0])
    # --- stage API on the oracle's projection
    bg = ctx.bin({k: dev(v) for k, v in fr["proj"].items()})
    assert bg["M"] == bo["M"]
    for k in ("tilesTouched", "tileCounts", "sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"):
        assert np.array_equal(u32(bg[k]), bo[k]), f"stage API {k}"
    del bg
    # K9 on the oracle's packed records: lastContrib tells where the two implementations disagree about the Gaussian at
    # which a pixel's transmittance crosses 1e-4 (:599-603).  That decision is a discontinuity: with T within one ulp of
    # the threshold one side stops and the other blends on, adding up to T * max colour ~ 1e-4 * c_max to the pixel.
    fwd_g = ctx.raster_fwd(dev(fr["packed"]))
    flip = (u32(fwd_g["lastContrib"]).reshape(-1).astype(np.int64) != fr["fwd"]["lastContrib"].reshape(-1).astype(np.int64))
    c_max = float(max(fr["packed"][:, 6:9].max(), 1.0))
    del fwd_g
    # --- fused product path
    dparams = {k: dev(v) for k, v in params.items()}
    render, depth, alpha, vis, radii = ctx.render_forward(dparams, gcam)
    M = ctx.stats()["pairs_last_view"]
    flips = int((radii.cpu().numpy() != fr["radii"]).sum())
    assert M == bo["M"] and flips == 0, f"fused path: M {M} vs {bo['M']}, {flips} radius flips"
    lists = ctx.bin_read()
    for k in ("sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"):
        assert np.array_equal(u32(lists[k]), bo[k]), f"fused path {k}"
    del lists
    d_pix = np.abs(render.cpu().numpy() - fr["render"]).reshape(-1, 3).max(axis=1)
    e_pix = float(d_pix[~flip].max())
    e_flip = float(d_pix[flip].max()) if flip.any() else 0.0
    e_alpha = float(np.abs(alpha.cpu().numpy() - fr["alpha"]).max())
    e_depth = float(np.abs(depth.cpu().numpy() - fr["depth"]).max())
    print(f"\n[{wl_name}] pixels with the same terminating Gaussian: max |d| {e_pix:.2e}; {int(flip.sum())} of {flip.size} pixels "
          f"({flip.mean():.1e}) terminate one Gaussian apart: max |d| {e_flip:.2e} (bound 1e-4 x c_max = {1e-4 * c_max:.2e})")
    assert e_pix <= PIX_TOL and e_alpha <= PIX_TOL and e_depth <= PIX_TOL * 10
    assert flip.mean() < 1e-3 and e_flip <= PIX_TOL * (1.0 + c_max)
    assert (vis.cpu().numpy() == fr["visibility_filter"]).all()
    loss, cot = ctx.loss_fwd_bwd(render, dev(targets[0]), 1.0)
    e_loss = abs(float(loss.item()) - lo["loss"])
    assert e_loss < 2e-5
    e_cot = rel_err(cot.cpu().numpy(), lo["cot_render"])
    assert e_cot < GRAD_TOL
    grads = ctx.render_backward(cot)
    errs = {k: rel_err(g.cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) for k, g in grads.items()}
    print(f"[{wl_name} vs {o.kind} oracle] N={N} M={M} (bit-exact lists) pixels {e_pix:.2e} alpha {e_alpha:.2e} depth {e_depth:.2e} "
          f"loss {e_loss:.2e} cot {e_cot:.2e} grads " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < GRAD_TOL, (k, v)
    # the whole-block backward (no forward checkpoints) meets the same bound
    ctx.set_flags(L.GSB_FLAG_NO_SEGMENTS)
    ctx.render_forward(dparams, gcam)
    g0 = ctx.render_backward(cot)
    for k in g0:
        assert rel_err(g0[k].cpu().numpy().reshape(bw["grads"][k].shape), bw["grads"][k]) < GRAD_TOL, ("no segments", k)
    ctx.close()


# ------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties (the oracle would take minutes here)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wl_name,n", [("C2", None), ("C3", None)])
def test_full_size_properties(gsb, wl_name, n):
    Context, L = gsb
    wl, params, cams, targets = make_workload(wl_name, views_override=1)
    N = params["_xyz"].shape[0]
    ctx = Context(wl.width, wl.height, sh_degree=wl.sh_degree, max_gaussians=N)
    dparams = {k: dev(v) for k, v in params.items()}
    gcam = L.make_camera(cams[0])
    render, depth, alpha, vis, radii = ctx.render_forward(dparams, gcam)
    M = ctx.stats()["pairs_last_view"]
    lists = ctx.bin_read()
    key = (lists["sortedKeysHigh"].to(torch.int64) << 32) | (lists["sortedKeysLow"].to(torch.int64) & 0xffffffff)
    assert key.numel() == M and bool((key[1:] >= key[:-1]).all()), "sorted (tile, depth) keys must be non-decreasing"
    eq = key[1:] == key[:-1]
    sv = lists["sortedGaussIdx"]
    assert bool((sv[1:][eq] > sv[:-1][eq]).all()), "equal keys keep emission (Gaussian index) order"
    # every listed Gaussian is visible, and the multiset of values matches tiles-touched
    cnt = torch.bincount(sv.long(), minlength=N)
    assert bool((vis | (cnt == 0)).all()), "only visible Gaussians are listed"
    assert int((cnt > 0).sum()) > 0.9 * int(vis.sum())
    # the stage API on the stage API's projection produces the same pairs as the fused path (same pinned activations)
    act = ctx.activate_fwd(dparams)
    proj = ctx.project_fwd(act, gcam)
    bins = ctx.bin(proj, read_lists=False)
    assert bins["M"] == M
    tc = bins["tileCounts"].long()
    assert int(tc.sum()) == bins["M"] == int(bins["tilesTouched"].long().sum())
    rg = bins["tileRanges"].long()
    nz = tc > 0
    assert bool((rg[nz, 1] - rg[nz, 0] == tc[nz]).all())
    starts = rg[nz, 0]
    assert bool((starts[1:] == rg[nz, 1][:-1]).all()) and int(starts[0]) == 0 and int(rg[nz, 1][-1]) == bins["M"]
    # image sanity
    a = alpha
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0 and bool(torch.isfinite(render).all())
    # re-render after the API calls (state was replaced) and check backward linearity in the cotangent
    render2, *_ = ctx.render_forward(dparams, gcam)
    assert torch.equal(render, render2), "the forward is deterministic"
    cot = torch.from_numpy(targets[0]).cuda() - 0.5
    g1 = ctx.render_backward(cot)
    g1 = {k: v.clone() for k, v in g1.items()}
    g2 = ctx.render_backward(cot * 2.0)
    for k in g1:
        assert rel_err(g2[k].cpu().numpy(), 2.0 * g1[k].cpu().numpy()) < 1e-4, k
    assert all(bool(torch.isfinite(v).all()) for v in g1.values())
    # the segmented backward (forward checkpoints every 256 Gaussians) against whole-block work items
    ctx.set_flags(L.GSB_FLAG_NO_SEGMENTS)
    render3, *_ = ctx.render_forward(dparams, gcam)
    assert float((render - render3).abs().max()) < 1e-5     # per-segment colour sums vs one running sum
    g3 = ctx.render_backward(cot)
    ctx.set_flags(0)
    for k in g1:
        assert rel_err(g1[k].cpu().numpy(), g3[k].cpu().numpy()) < 1e-4, k
    if wl_name == "C3":
        # SURVEY.md 8d-workload quotes 12 031 308 for libm expf activations; with the pinned activation exponential the
        # reference kernels give the count asserted exactly in test_baseline_sizes_vs_reference_oracle
        assert abs(M - 12_031_308) <= 200


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs 4 and 5 at full size: size-independent properties
# ------------------------------------------------------------------------------------------------
def test_c5_forward_only_render_3m(gsb):
    """Forward-only render (GaussianRenderer.forward, GaussianRenderer.swift:882-934) of 3 M Gaussians at 1080p."""
    Context, L = gsb
    n, W, H = 3_000_000, 1920, 1080
    params = make_gaussians(n, 5, 3)
    ctx = Context(W, H, max_gaussians=n)
    dp = {k: dev(v) for k, v in params.items()}
    cam = L.make_camera(make_cameras(W, H, 8)[0])
    render, depth, alpha, vis, radii = ctx.render_forward(dp, cam)
    M = ctx.stats()["pairs_last_view"]
    assert M > 10 * n // 4 and bool(torch.isfinite(render).all()) and float(alpha.min()) >= 0.0 and float(alpha.max()) <= 1.0
    lists = ctx.bin_read()
    key = (lists["sortedKeysHigh"].to(torch.int64) << 32) | (lists["sortedKeysLow"].to(torch.int64) & 0xffffffff)
    assert key.numel() == M and bool((key[1:] >= key[:-1]).all())
    eq = key[1:] == key[:-1]
    sv = lists["sortedGaussIdx"]
    assert bool((sv[1:][eq] > sv[:-1][eq]).all())
    del key, eq, lists
    render2, *_ = ctx.render_forward(dp, cam)
    assert torch.equal(render, render2)
    assert ctx.last_contrib_sum() <= M * 256
    ctx.close()


def test_c4_densification_stress_6m_4k(gsb):
    """6 M Gaussians at 3840x2160: train, clone/split/prune on the accumulated gradient norms, train on the new count."""
    Context, L = gsb
    wl, params, cams, targets = make_workload("C4")
    n = params["_xyz"].shape[0]
    ctx = Context(wl.width, wl.height, sh_degree=wl.sh_degree)           # max_gaussians = 0: buffers grow with the model
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    gc = [L.make_camera(c) for c in cams]
    tg = [torch.from_numpy(t).cuda() for t in targets]
    losses = [ctx.train_step(gc, tg, it, 30000) for it in range(2)]
    assert all(np.isfinite(l) for l in losses) and ctx.trainer_count() == (n, 2)
    acc = ctx.trainer_tensors()["accum"]
    thr = float(torch.quantile(acc[::8], 0.95)) / 2.0                        # ~5 % of the Gaussians densify
    op = ctx.trainer_tensors()["params"]["_opacity"]
    pruned_expected = int((torch.sigmoid(op[:, 0]) < 0.005).sum())
    info = ctx.trainer_densify(thr, 0.01, 0.005, 8_000_000, seed=7)
    assert info["prune"] == pruned_expected
    assert info["keep"] + info["split"] + info["clone"] + info["prune"] == n
    assert info["n"] == info["keep"] + 2 * (info["split"] + info["clone"]) == info["total"]
    assert 0.03 * n < info["split"] + info["clone"] < 0.07 * n
    tt = ctx.trainer_tensors()
    assert tt["params"]["_xyz"].shape[0] == info["n"] and float(tt["m"]["_xyz"].abs().max()) == 0.0
    assert bool(torch.isfinite(tt["params"]["_xyz"]).all())
    l2 = ctx.train_step(gc, tg, 2, 30000)
    assert np.isfinite(l2) and ctx.trainer_count() == (info["n"], 1)
    ctx.close()


def test_async_loss_flag_matches_synchronous_loss(gsb):
    """GSB_FLAG_ASYNC_LOSS: the loss lands in the caller's pinned slot without a synchronisation inside the call."""
    Context, L = gsb
    n, W, H = 800, 64, 48
    params = make_gaussians(n, 81, 3)
    cams = [L.make_camera(c) for c in make_cameras(W, H, 2)]
    tg = [torch.from_numpy(t).cuda() for t in make_targets(W, H, 2, 81)]
    ref = Context(W, H)
    ref.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    want = [ref.train_step(cams, tg, it, 100) for it in range(2)]
    ctx = Context(W, H, flags=L.GSB_FLAG_ASYNC_LOSS)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    slots = [torch.full((1,), -1.0).pin_memory() for _ in range(2)]
    for it in range(2):
        ctx.trainer_accumulate(cams, tg, loss_out=slots[it])
        ctx.trainer_apply(it, 100)
    torch.cuda.synchronize()
    for it in range(2):
        assert abs(float(slots[it][0]) - want[it]) < 1e-6

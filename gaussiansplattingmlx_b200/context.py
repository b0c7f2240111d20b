"""``Context`` — thin object wrapper over a ``gsb_ctx`` for callers that hold torch CUDA tensors.

torch is used for device memory and streams only (plumbing); every computation is a call into
``libgsb.so``.  All methods enqueue on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import GsbCamera, GsbConfig, GsbError, GsbStats

PARAM_NAMES = ("_xyz", "_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "gsb takes contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32, "gsb arithmetic is f32"
    return t.contiguous()


class Context:
    def __init__(self, width: int, height: int, tile_w: int = 16, tile_h: int = 16, sh_degree: int = 3,
                 sh_coeffs: Optional[int] = None, white_background: bool = False, max_gaussians: int = 0,
                 device: int = 0, flags: int = 0, lambda_dssim: float = 0.2):
        self.lib = _lib.load()
        cfg = GsbConfig()
        self.lib.gsb_default_config(C.byref(cfg))
        cfg.width, cfg.height, cfg.tile_w, cfg.tile_h = width, height, tile_w, tile_h
        cfg.sh_degree = sh_degree
        cfg.sh_coeffs = sh_coeffs if sh_coeffs is not None else (sh_degree + 1) ** 2
        cfg.white_background = int(bool(white_background))
        cfg.max_gaussians = max_gaussians
        cfg.device = device
        cfg.flags = flags
        cfg.lambda_dssim = lambda_dssim
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.W, self.H, self.K = width, height, cfg.sh_coeffs
        self.P = width * height
        self.grid_w = (width + tile_w - 1) // tile_w
        self.grid_h = (height + tile_h - 1) // tile_h
        self.num_tiles = self.grid_w * self.grid_h
        h = C.c_void_p()
        rc = self.lib.gsb_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise GsbError(rc, (self.lib.gsb_last_error(None) or b"").decode())
        self.h = h
        self._stream = None

    # ---- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.gsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise GsbError(rc, (self.lib.gsb_last_error(self.h) or b"").decode())

    def _sync_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            self._check(self.lib.gsb_set_stream(self.h, C.c_void_p(s)))
            self._stream = s

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def synchronize(self):
        self._check(self.lib.gsb_synchronize(self.h))

    def set_flags(self, flags: int):
        """Replace gsb_config.flags (e.g. GSB_FLAG_NO_OVERLAP for per-kernel timing)."""
        self._check(self.lib.gsb_set_flags(self.h, int(flags)))
        self.cfg.flags = int(flags)

    # ---- activations ------------------------------------------------------------------------
    def activate_fwd(self, params: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        n = params["_xyz"].shape[0]
        shs, scales, rot, op = self._new(n, self.K, 3), self._new(n, 3), self._new(n, 4), self._new(n, 1)
        self._check(self.lib.gsb_activate_fwd(self.h, n, _ptr(_f32(params["_features_dc"])), _ptr(_f32(params["_features_rest"])),
                                              _ptr(_f32(params["_scales"])), _ptr(_f32(params["_rotation"])),
                                              _ptr(_f32(params["_opacity"])), _ptr(shs), _ptr(scales), _ptr(rot), _ptr(op)))
        return {"means3d": params["_xyz"], "shs": shs, "scales": scales, "rotations": rot, "opacity": op}

    def activate_bwd(self, params, g_act) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        n = params["_xyz"].shape[0]
        out = {k: torch.empty_like(params[k]) for k in PARAM_NAMES[1:]}
        self._check(self.lib.gsb_activate_bwd(self.h, n, _ptr(_f32(params["_scales"])), _ptr(_f32(params["_rotation"])),
                                              _ptr(_f32(params["_opacity"])), _ptr(_f32(g_act["shs"])), _ptr(_f32(g_act["scales"])),
                                              _ptr(_f32(g_act["rotations"])), _ptr(_f32(g_act["opacity"])),
                                              _ptr(out["_features_dc"]), _ptr(out["_features_rest"]), _ptr(out["_scales"]),
                                              _ptr(out["_rotation"]), _ptr(out["_opacity"])))
        out["_xyz"] = g_act["means3d"]
        return out

    # ---- K1 / K2 ----------------------------------------------------------------------------
    def project_fwd(self, act, cam: GsbCamera) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        n = act["means3d"].shape[0]
        o = {"means2d": self._new(n, 2), "depths": self._new(n), "color": self._new(n, 3), "cov2d": self._new(n, 2, 2),
             "conic": self._new(n, 2, 2), "radii": self._new(n), "rectMin": self._new(n, 2), "rectMax": self._new(n, 2)}
        self._check(self.lib.gsb_project_fwd(self.h, n, _ptr(_f32(act["scales"])), _ptr(_f32(act["rotations"])),
                                             _ptr(_f32(act["means3d"])), _ptr(_f32(act["shs"])), C.byref(cam), _ptr(o["means2d"]),
                                             _ptr(o["depths"]), _ptr(o["color"]), _ptr(o["cov2d"]), _ptr(o["conic"]),
                                             _ptr(o["radii"]), _ptr(o["rectMin"]), _ptr(o["rectMax"])))
        return o

    def project_bwd(self, act, cam: GsbCamera, cot) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        n = act["means3d"].shape[0]
        g = {"scales": self._new(n, 3), "rotations": self._new(n, 4), "means3d": self._new(n, 3),
             "shs": self._new(n, self.K, 3), "cameraCenterPoint": self._new(n, 3)}
        self._check(self.lib.gsb_project_bwd(self.h, n, _ptr(_f32(act["scales"])), _ptr(_f32(act["rotations"])),
                                             _ptr(_f32(act["means3d"])), _ptr(_f32(act["shs"])), C.byref(cam),
                                             _ptr(_f32(cot["depths"])), _ptr(_f32(cot["means2d"])), _ptr(_f32(cot["cov2d"])),
                                             _ptr(_f32(cot["color"])), _ptr(_f32(cot["conic"])), _ptr(g["scales"]),
                                             _ptr(g["rotations"]), _ptr(g["means3d"]), _ptr(g["shs"]), _ptr(g["cameraCenterPoint"])))
        return g

    # ---- K3..K8 -----------------------------------------------------------------------------
    def bin(self, proj, read_lists: bool = True) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        n = proj["radii"].shape[0]
        touched = self._new(n, dtype=torch.int32)
        ranges = self._new(self.num_tiles, 2, dtype=torch.int32)
        counts = self._new(self.num_tiles, dtype=torch.int32)
        m = C.c_uint32(0)
        self._check(self.lib.gsb_bin(self.h, n, _ptr(_f32(proj["rectMin"])), _ptr(_f32(proj["rectMax"])), _ptr(_f32(proj["radii"])),
                                     _ptr(_f32(proj["depths"])), _ptr(touched), _ptr(ranges), _ptr(counts), C.byref(m)))
        out = {"tilesTouched": touched, "tileRanges": ranges, "tileCounts": counts, "M": int(m.value)}
        if read_lists:
            out.update(self.bin_read(unsorted=True))
        return out

    def bin_read(self, unsorted: bool = False) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        st = GsbStats()
        self._check(self.lib.gsb_stats_get(self.h, C.byref(st)))
        m = int(st.pairs_last_view)
        names = ["sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"]
        if unsorted:
            names = ["keysHigh", "keysLow", "gaussIdx"] + names
        bufs = {k: self._new(m, dtype=torch.int32) for k in names}
        args = [_ptr(bufs[k]) if k in bufs else None for k in
                ("keysHigh", "keysLow", "gaussIdx", "sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx")]
        self._check(self.lib.gsb_bin_read(self.h, *args))
        return bufs

    def sort_tile_keys(self, keys_high, keys_low, values, tile_bits: int, use_cub: bool = False):
        self._sync_stream()
        m = keys_high.numel()
        outs = [torch.empty_like(keys_high), torch.empty_like(keys_low), torch.empty_like(values)]
        self._check(self.lib.gsb_sort_tile_keys(self.h, m, tile_bits, _ptr(keys_high.contiguous()), _ptr(keys_low.contiguous()),
                                                _ptr(values.contiguous()), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]),
                                                int(use_cub)))
        return outs

    # ---- K9 / K10 ---------------------------------------------------------------------------
    def raster_fwd(self, packed) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        n = packed.shape[0]
        o = {"color": self._new(self.P, 3), "depth": self._new(self.P, 1), "alpha": self._new(self.P, 1),
             "lastContrib": self._new(self.P, 1, dtype=torch.int32)}
        self._check(self.lib.gsb_raster_fwd(self.h, n, _ptr(_f32(packed)), _ptr(o["color"]), _ptr(o["depth"]), _ptr(o["alpha"]),
                                            _ptr(o["lastContrib"])))
        return o

    def raster_bwd(self, packed, cot, fwd) -> torch.Tensor:
        self._sync_stream()
        n = packed.shape[0]
        g = self._new(n, 11)
        self._check(self.lib.gsb_raster_bwd(self.h, n, _ptr(_f32(packed)), _ptr(_f32(cot["color"])),
                                            _ptr(_f32(cot["depth"])) if cot.get("depth") is not None else None,
                                            _ptr(_f32(cot["alpha"])) if cot.get("alpha") is not None else None,
                                            _ptr(fwd["color"]), _ptr(fwd["depth"]), _ptr(fwd["alpha"]), _ptr(fwd["lastContrib"]),
                                            _ptr(g)))
        return g

    # ---- K11 / K12 --------------------------------------------------------------------------
    def ssim_fwd(self, img1, img2, saved: bool = True) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        H, W, Cn = img1.shape
        names = ("ssim", "mu1", "mu2", "sigma1", "sigma2", "sigma12")
        o = {k: self._new(H, W, Cn) for k in (names if saved else names[:1])}
        self._check(self.lib.gsb_ssim_fwd(self.h, H, W, Cn, _ptr(_f32(img1)), _ptr(_f32(img2)),
                                          *[_ptr(o[k]) if k in o else None for k in names]))
        return o

    def ssim_bwd(self, grad_out, img1, img2) -> torch.Tensor:
        self._sync_stream()
        H, W, Cn = img1.shape
        g = self._new(H, W, Cn)
        self._check(self.lib.gsb_ssim_bwd(self.h, H, W, Cn, _ptr(_f32(grad_out)), _ptr(_f32(img1)), _ptr(_f32(img2)), _ptr(g)))
        return g

    # ---- fused renderer / loss ---------------------------------------------------------------
    def render_forward(self, params: Dict[str, torch.Tensor], cam: GsbCamera, want_outputs: bool = True):
        self._sync_stream()
        n = params["_xyz"].shape[0]
        render = self._new(self.H, self.W, 3) if want_outputs else None
        depth = self._new(self.H, self.W, 1) if want_outputs else None
        alpha = self._new(self.H, self.W, 1) if want_outputs else None
        vis = self._new(n, dtype=torch.uint8) if want_outputs else None
        radii = self._new(n) if want_outputs else None
        self._saved_params = params  # keep the tensors alive until the backward
        self._check(self.lib.gsb_render_forward(self.h, n, *[_ptr(_f32(params[k])) for k in PARAM_NAMES], C.byref(cam),
                                                _ptr(render), _ptr(depth), _ptr(alpha), _ptr(vis), _ptr(radii)))
        return render, depth, alpha, (vis.bool() if vis is not None else None), radii

    def render_backward(self, cot_render, cot_depth=None, cot_alpha=None, grads: Optional[Dict[str, torch.Tensor]] = None,
                        accumulate: bool = False) -> Dict[str, torch.Tensor]:
        self._sync_stream()
        params = self._saved_params
        if grads is None:
            grads = {k: torch.empty_like(params[k]) for k in PARAM_NAMES}
            accumulate = False
        self._check(self.lib.gsb_render_backward(self.h, _ptr(_f32(cot_render)), _ptr(cot_depth), _ptr(cot_alpha),
                                                 *[_ptr(grads[k]) for k in PARAM_NAMES], int(accumulate)))
        return grads

    def loss_fwd_bwd(self, render, target, grad_scale: float = 1.0, depth=None, depth_mask=None, target_depth=None,
                     lambda_depth: float = 0.0):
        """Returns (loss tensor [1] on device, cot_render[H,W,3]); with depth supervision (``depth`` = the renderer's depth
        output, ``depth_mask`` bool/uint8 [H,W], ``target_depth`` f32 [H,W]) also cot_depth[H,W,1] as a third value."""
        self._sync_stream()
        cot = torch.empty_like(render)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        if depth is None:
            self._check(self.lib.gsb_loss_fwd_bwd(self.h, _ptr(_f32(render)), _ptr(_f32(target)), C.c_float(grad_scale), _ptr(cot),
                                                  _ptr(loss)))
            return loss, cot
        mask = depth_mask.to(torch.uint8).contiguous()
        cot_depth = torch.empty_like(depth)
        self._check(self.lib.gsb_loss_fwd_bwd_depth(self.h, _ptr(_f32(render)), _ptr(_f32(depth)), _ptr(_f32(target)), _ptr(mask),
                                                    _ptr(_f32(target_depth)), C.c_float(lambda_depth), C.c_float(grad_scale),
                                                    _ptr(cot), _ptr(cot_depth), _ptr(loss)))
        return loss, cot, cot_depth

    # ---- Adam -------------------------------------------------------------------------------
    def adam_step(self, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], m: Sequence[torch.Tensor],
                  v: Sequence[torch.Tensor], lrs: Sequence[float], grad_norm_accum: Optional[torch.Tensor] = None):
        self._sync_stream()
        arr = lambda ts: (C.c_void_p * 6)(*[t.data_ptr() for t in ts])
        counts = (C.c_int64 * 6)(*[t.numel() for t in params])
        lr = (C.c_float * 6)(*lrs)
        n = params[0].shape[0]
        self._check(self.lib.gsb_adam_step(self.h, n, arr(params), arr(grads), arr(m), arr(v), counts, lr, _ptr(grad_norm_accum)))

    # ---- trainer ----------------------------------------------------------------------------
    def trainer_init(self, params: Dict[str, torch.Tensor]):
        """``params`` may be host (numpy-backed CPU tensors) or CUDA tensors."""
        self._sync_stream()
        n = params["_xyz"].shape[0]
        ptrs = [C.c_void_p(params[k].contiguous().data_ptr()) for k in PARAM_NAMES]
        self._trainer_src = params
        self._check(self.lib.gsb_trainer_init(self.h, n, *ptrs))
        self.tN = n

    def trainer_tensors(self):
        """Zero-copy torch views of the context-owned params / grads / m / v / accum buffers."""
        P = (C.c_void_p * 6)(); G = (C.c_void_p * 6)(); M = (C.c_void_p * 6)(); V = (C.c_void_p * 6)()
        A = C.c_void_p()
        self._check(self.lib.gsb_trainer_param_ptrs(self.h, P, G, M, V, C.byref(A)))
        K = self.K
        shapes = [(self.tN, 3), (self.tN, 1, 3), (self.tN, K - 1, 3), (self.tN, 3), (self.tN, 4), (self.tN, 1)]
        def view(ptr, shape):
            return _wrap_device_memory(ptr, shape, self.device)
        out = {}
        for name, arr in (("params", P), ("grads", G), ("m", M), ("v", V)):
            out[name] = {k: view(arr[i], shapes[i]) for i, k in enumerate(PARAM_NAMES)}
        out["accum"] = view(A.value, (self.tN,))
        return out

    def trainer_grad_block(self) -> torch.Tensor:
        p = C.c_void_p(); n = C.c_int64()
        self._check(self.lib.gsb_trainer_grad_block(self.h, C.byref(p), C.byref(n)))
        return _wrap_device_memory(p.value, (int(n.value),), self.device)

    def _cams_targets(self, cams: Sequence[GsbCamera], targets: Sequence[torch.Tensor]):
        B = len(cams)
        cam_arr = (GsbCamera * B)(*cams)
        on_host = not targets[0].is_cuda
        for t in targets:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda != on_host
            if on_host:
                assert t.is_pinned(), "host targets must be pinned"
        tp = (C.c_void_p * B)(*[t.data_ptr() for t in targets])
        return B, cam_arr, tp, int(on_host)

    def trainer_accumulate(self, cams, targets, zero_grads: bool = True, grad_scale: Optional[float] = None,
                           want_loss: bool = True, loss_out: Optional[torch.Tensor] = None,
                           target_depths: Optional[Sequence[torch.Tensor]] = None,
                           depth_masks: Optional[Sequence[torch.Tensor]] = None, lambda_depth: float = 0.0) -> Optional[float]:
        """``loss_out``: a pinned 1-element f32 CPU tensor; with GSB_FLAG_ASYNC_LOSS set the mean loss is copied into it
        asynchronously (no synchronisation; read it after an event/stream sync) and the return value is None.
        ``target_depths`` (f32 [H,W]) + ``depth_masks`` (uint8 [H,W]), living where the targets live, switch on the
        depth-supervision term with weight ``lambda_depth``."""
        self._sync_stream()
        B, cam_arr, tp, on_host = self._cams_targets(cams, targets)
        scale = (1.0 / B) if grad_scale is None else grad_scale
        if target_depths is not None:
            assert depth_masks is not None and len(target_depths) == B and len(depth_masks) == B
            for d, m in zip(target_depths, depth_masks):
                assert d.dtype == torch.float32 and m.dtype == torch.uint8 and d.is_contiguous() and m.is_contiguous()
                assert d.is_cuda != bool(on_host) and m.is_cuda != bool(on_host) and d.numel() == self.P and m.numel() == self.P
            dp = (C.c_void_p * B)(*[t.data_ptr() for t in target_depths])
            mp = (C.c_void_p * B)(*[t.data_ptr() for t in depth_masks])
            call = lambda loss_ptr: self.lib.gsb_trainer_accumulate_depth(self.h, B, cam_arr, tp, dp, mp, C.c_float(lambda_depth), on_host,
                                                                          int(zero_grads), C.c_float(scale), loss_ptr)
        else:
            call = lambda loss_ptr: self.lib.gsb_trainer_accumulate(self.h, B, cam_arr, tp, on_host, int(zero_grads), C.c_float(scale),
                                                                    loss_ptr)
        if loss_out is not None:
            assert self.cfg.flags & _lib.GSB_FLAG_ASYNC_LOSS and loss_out.is_pinned() and loss_out.dtype == torch.float32
            self._check(call(C.c_void_p(loss_out.data_ptr())))
            return None
        assert not (want_loss and self.cfg.flags & _lib.GSB_FLAG_ASYNC_LOSS), "async-loss contexts need loss_out"
        loss = C.c_float(0.0)
        self._check(call(C.cast(C.pointer(loss), C.c_void_p) if want_loss else None))
        return float(loss.value) if want_loss else None

    def trainer_apply(self, iteration: int, total_iterations: int, reset_state: bool = False):
        self._sync_stream()
        self._check(self.lib.gsb_trainer_apply(self.h, iteration, total_iterations, int(reset_state)))

    # ---- peer-memory data parallelism (gsb.h: gsb_trainer_peers_*) -------------------------------
    def trainer_peers_export(self) -> bytes:
        """This replica's IPC blob (trainer slab handles + layout); exchange the blobs of all ranks, in rank order."""
        buf = C.create_string_buffer(_lib.GSB_PEER_BLOB_BYTES)
        self._check(self.lib.gsb_trainer_peers_export(self.h, C.cast(buf, C.c_void_p), _lib.GSB_PEER_BLOB_BYTES))
        return buf.raw

    def trainer_peers_import(self, world: int, rank: int, blobs: Sequence[bytes]):
        assert len(blobs) == world and all(len(b) == _lib.GSB_PEER_BLOB_BYTES for b in blobs)
        buf = C.create_string_buffer(b"".join(blobs), world * _lib.GSB_PEER_BLOB_BYTES)
        self._check(self.lib.gsb_trainer_peers_import(self.h, world, rank, C.cast(buf, C.c_void_p), _lib.GSB_PEER_BLOB_BYTES))

    def trainer_peers_close(self):
        self._check(self.lib.gsb_trainer_peers_close(self.h))

    def trainer_apply_peers(self, iteration: int, total_iterations: int, reset_state: bool = False):
        """Fused gradient reduction + Adam + parameter broadcast over NVLink peer memory; the caller brackets it with
        two stream-ordered barriers (see dp.ViewParallel.peer_step)."""
        self._sync_stream()
        self._check(self.lib.gsb_trainer_apply_peers(self.h, iteration, total_iterations, int(reset_state)))

    def trainer_step_peers(self, cams, targets, grad_scale: float, iteration: int, total_iterations: int, reset_state: bool = False,
                           want_loss: bool = False, loss_out: Optional[torch.Tensor] = None) -> Optional[float]:
        """The whole data-parallel step with device-side synchronisation (gsb.h: gsb_trainer_step_peers); ``cams`` may be empty
        on a replica without views.  Loss semantics as in ``trainer_accumulate``."""
        self._sync_stream()
        if len(cams):
            B, cam_arr, tp, on_host = self._cams_targets(cams, targets)
        else:
            B, cam_arr, tp, on_host = 0, None, None, 0
        if loss_out is not None:
            assert self.cfg.flags & _lib.GSB_FLAG_ASYNC_LOSS and loss_out.is_pinned() and loss_out.dtype == torch.float32
            self._check(self.lib.gsb_trainer_step_peers(self.h, B, cam_arr, tp, on_host, C.c_float(grad_scale), iteration, total_iterations,
                                                        int(reset_state), C.c_void_p(loss_out.data_ptr())))
            return None
        loss = C.c_float(0.0)
        self._check(self.lib.gsb_trainer_step_peers(self.h, B, cam_arr, tp, on_host, C.c_float(grad_scale), iteration, total_iterations,
                                                    int(reset_state), C.cast(C.pointer(loss), C.c_void_p) if want_loss else None))
        return float(loss.value) if want_loss else None

    def trainer_peers_tune(self, chunks: int = 0, peer_blocks: int = 0, multicast_blocks: int = 0):
        self._check(self.lib.gsb_trainer_peers_tune(self.h, int(chunks), int(peer_blocks), int(multicast_blocks)))

    def trainer_peers_check(self):
        self._check(self.lib.gsb_trainer_peers_check(self.h))

    def trainer_attach_symmetric(self, world: int, rank: int, params_local: int, grads_local: int, params_mc: int, grads_mc: int,
                                 floats: int):
        """Device addresses (ints) of this replica's symmetric parameter / gradient buffers and of their multicast mappings."""
        self._sync_stream()
        self._check(self.lib.gsb_trainer_attach_symmetric(self.h, world, rank, C.c_void_p(params_local), C.c_void_p(grads_local),
                                                          C.c_void_p(params_mc), C.c_void_p(grads_mc), int(floats)))

    def trainer_apply_multicast(self, iteration: int, total_iterations: int, reset_state: bool = False):
        self._sync_stream()
        self._check(self.lib.gsb_trainer_apply_multicast(self.h, iteration, total_iterations, int(reset_state)))

    def train_step(self, cams, targets, iteration: int, total_iterations: int, want_loss: bool = True) -> Optional[float]:
        self._sync_stream()
        B, cam_arr, tp, on_host = self._cams_targets(cams, targets)
        loss = C.c_float(0.0)
        self._check(self.lib.gsb_train_step(self.h, B, cam_arr, tp, on_host, iteration, total_iterations,
                                            C.byref(loss) if want_loss else None))
        return float(loss.value) if want_loss else None

    # ---- densification (split_and_prune, GaussianTrainer.swift:766-908) ----------------------
    def densify_classify(self, grad_accum, denom: float, scales_log, opacity_logit, grad_threshold: float, max_scale: float,
                         min_opacity: float, allow_densify: bool):
        self._sync_stream()
        n = grad_accum.shape[0]
        actions, counts = self._new(n, dtype=torch.int32), self._new(n, dtype=torch.int32)
        self._check(self.lib.gsb_densify_classify(self.h, n, _ptr(_f32(grad_accum)), C.c_float(denom), _ptr(_f32(scales_log)),
                                                  _ptr(_f32(opacity_logit)), C.c_float(grad_threshold), C.c_float(max_scale),
                                                  C.c_float(min_opacity), int(bool(allow_densify)), _ptr(actions), _ptr(counts)))
        return actions, counts

    def densify_map(self, actions, counts):
        """cumsum + build_densify_output_map; returns (offsets[N], gather[N'], noise_mode[N'])."""
        self._sync_stream()
        n = actions.shape[0]
        cap = 2 * n
        offsets = self._new(n, dtype=torch.int32)
        gather, mode = self._new(max(cap, 1), dtype=torch.int32), self._new(max(cap, 1), dtype=torch.int32)
        total = C.c_int32(0)
        self._check(self.lib.gsb_densify_map(self.h, n, _ptr(actions), _ptr(counts), _ptr(offsets), cap, _ptr(gather), _ptr(mode),
                                             C.byref(total)))
        t = int(total.value)
        return offsets, gather[:t], mode[:t]

    def densify_apply(self, params: Dict[str, torch.Tensor], gather, noise_mode, base_noise=None, seed: int = 0):
        self._sync_stream()
        nout = gather.shape[0]
        K = self.K
        out = {"_xyz": self._new(nout, 3), "_features_dc": self._new(nout, 1, 3), "_features_rest": self._new(nout, K - 1, 3),
               "_scales": self._new(nout, 3), "_rotation": self._new(nout, 4), "_opacity": self._new(nout, 1)}
        self._check(self.lib.gsb_densify_apply(self.h, nout, _ptr(gather.contiguous()), _ptr(noise_mode.contiguous()),
                                               _ptr(base_noise), C.c_uint64(seed), *[_ptr(_f32(params[k])) for k in PARAM_NAMES],
                                               *[_ptr(out[k]) for k in PARAM_NAMES]))
        return out

    def trainer_densify(self, grad_threshold: float = 0.0002, max_scale: float = 0.01, min_opacity: float = 0.005,
                        max_gaussians: int = 1_000_000, seed: int = 0, base_noise=None) -> Dict[str, int]:
        """split_and_prune on the context-owned tensors; returns keep/split/clone/prune/total and the new count."""
        self._sync_stream()
        counts = (C.c_int32 * 5)()
        self._check(self.lib.gsb_trainer_densify(self.h, C.c_float(grad_threshold), C.c_float(max_scale), C.c_float(min_opacity),
                                                 int(max_gaussians), C.c_uint64(seed), _ptr(base_noise), counts))
        n, steps = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.gsb_trainer_count(self.h, C.byref(n), C.byref(steps)))
        self.tN = int(n.value)
        return {"keep": counts[0], "split": counts[1], "clone": counts[2], "prune": counts[3], "total": counts[4], "n": self.tN}

    def trainer_count(self):
        n, steps = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.gsb_trainer_count(self.h, C.byref(n), C.byref(steps)))
        return int(n.value), int(steps.value)

    # ---- stats ------------------------------------------------------------------------------
    def stats(self) -> Dict:
        st = GsbStats()
        self._check(self.lib.gsb_stats_get(self.h, C.byref(st)))
        names = [self.lib.gsb_stage_name(i).decode() for i in range(_lib.STAGE_COUNT)]
        return {"kernel_launches": int(st.kernel_launches), "pairs_last_view": int(st.pairs_last_view),
                "pairs_total": int(st.pairs_total), "views": int(st.views), "pair_capacity": int(st.pair_capacity),
                "sb_pairs_last_view": int(st.sb_pairs_last_view),
                "stage_ms": {n: float(st.stage_ms[i]) for i, n in enumerate(names)},
                "stage_calls": {n: int(st.stage_calls[i]) for i, n in enumerate(names)}}

    def tile_list_info(self) -> Dict[str, int]:
        a, b, n, p = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.lib.gsb_tile_list_info(self.h, C.byref(a), C.byref(b), C.byref(n), C.byref(p)))
        return {"sb_w": a.value, "sb_h": b.value, "num_superblocks": n.value, "sort_passes": p.value}

    def last_contrib_sum(self) -> int:
        v = C.c_uint64(0)
        self._check(self.lib.gsb_last_contrib_sum(self.h, C.byref(v)))
        return int(v.value)

    def bin_generation(self) -> int:
        v = C.c_uint64(0)
        self._check(self.lib.gsb_bin_generation(self.h, C.byref(v)))
        return int(v.value)

    def stage_sections(self) -> Dict[str, str]:
        """stage name -> section of the reference's profiler report (GaussianTrainer.swift:122-241)."""
        return {self.lib.gsb_stage_name(i).decode(): self.lib.gsb_stage_section(i).decode() for i in range(_lib.STAGE_COUNT)}

    def stats_reset(self):
        self._check(self.lib.gsb_stats_reset(self.h))

    def enable_stage_timing(self, on: bool):
        self._check(self.lib.gsb_enable_stage_timing(self.h, int(on)))


class _DevMem:
    """``__cuda_array_interface__`` carrier so torch can view library-owned device memory."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _wrap_device_memory(ptr: int, shape, device) -> torch.Tensor:
    return torch.as_tensor(_DevMem(ptr, shape), device=device)

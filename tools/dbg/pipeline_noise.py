"""Run-to-run noise of three pipelined train steps (float atomics + Adam's sign-like first steps): the same schedule
twice, then pipelined vs serial; prints the relative difference of the parameter deltas per tensor."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gaussiansplattingmlx_b200.context import Context
from gaussiansplattingmlx_b200 import _lib as L
from gaussiansplattingmlx_b200.scene import make_gaussians, make_cameras, make_targets

def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-8))

n, W, H = 40000, 160, 96
params = make_gaussians(n, 41, 3)
cams = make_cameras(W, H, 5)
targets = make_targets(W, H, 5, 41)
def run(flags):
    ctx = Context(W, H, flags=flags)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    tg = [torch.from_numpy(t).cuda() for t in targets]
    for it in (0, 1, 2):
        ctx.train_step([L.make_camera(c) for c in cams], tg, it, 100)
    tt = ctx.trainer_tensors()
    out = {k: v.cpu().numpy().copy() for k, v in tt["params"].items()}
    ctx.close()
    return out
for fa, fb in ((0, 0), (L.GSB_FLAG_NO_OVERLAP, L.GSB_FLAG_NO_OVERLAP), (0, L.GSB_FLAG_NO_OVERLAP), (L.GSB_FLAG_NO_SEGMENTS, L.GSB_FLAG_NO_SEGMENTS),
               (L.GSB_FLAG_NO_SEGMENTS, L.GSB_FLAG_NO_SEGMENTS | L.GSB_FLAG_NO_OVERLAP), (0, L.GSB_FLAG_NO_SEGMENTS)):
    a, b = run(fa), run(fb)
    print(fa, fb, {k: round(rel_err(a[k] - params[k].reshape(a[k].shape), b[k] - params[k].reshape(b[k].shape)), 6) for k in a})

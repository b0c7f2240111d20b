"""C2-size check of the segmented raster backward: gradients of the fused path with and without forward checkpoints,
each against the CPU oracle (reference kernels when built, else the C port)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gaussiansplattingmlx_b200.context import Context
from gaussiansplattingmlx_b200 import _lib as L
from gaussiansplattingmlx_b200.scene import make_workload
from oracle import api, pipeline as pl

def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-8))

wl, params, cams, targets = make_workload(sys.argv[1] if len(sys.argv) > 1 else "C2", views_override=1)
o = api.Port()
cot = (targets[0] - 0.5).astype(np.float32)
t = time.time()
fr = pl.render_forward(o, params, cams[0], wl.sh_degree)
bw = pl.backward(o, params, cams[0], wl.sh_degree, fr, cot)
print("oracle s", time.time() - t)
ctx = Context(wl.width, wl.height, sh_degree=wl.sh_degree, max_gaussians=params["_xyz"].shape[0])
dparams = {k: torch.from_numpy(v).cuda() for k, v in params.items()}
res = {}
for flags in (0, L.GSB_FLAG_NO_SEGMENTS):
    ctx.set_flags(flags)
    ctx.render_forward(dparams, L.make_camera(cams[0]))
    g = ctx.render_backward(torch.from_numpy(cot).cuda())
    res[flags] = {k: v.cpu().numpy().copy() for k, v in g.items()}
    print("flags", flags, "vs oracle", {k: round(rel_err(res[flags][k].reshape(bw["grads"][k].shape), bw["grads"][k]), 7) for k in res[flags]})
print("segmented vs whole-block", {k: round(rel_err(res[0][k], res[8][k]), 7) for k in res[0]})

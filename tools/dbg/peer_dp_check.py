"""2+ GPU check of the peer-memory data-parallel step against the NCCL all-reduce path (run under torchrun):
same initial model, same views, two steps each; compares the D1 accumulators (no Adam amplification: tight), the
parameter deltas (fraction of elements off by more than 1e-3 of the largest delta) and replica consistency."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from gaussiansplattingmlx_b200.context import Context
from gaussiansplattingmlx_b200 import _lib as L
from gaussiansplattingmlx_b200.dp import ViewParallel
from gaussiansplattingmlx_b200.scene import make_gaussians, make_cameras, make_targets

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, W, H = int(os.environ.get("N", 50001)), 320, 192      # N not a multiple of 4 x world: ragged last slice
params = make_gaussians(n, 7, 3)
views = 2 * world
cams = make_cameras(W, H, views)
targets = make_targets(W, H, views, 7)
mine = [v for v in range(views) if v % world == rank]
gc = [L.make_camera(cams[v]) for v in mine]
tg = [torch.from_numpy(targets[v]).cuda() for v in mine]
vp = ViewParallel(rank, world)

MODE = os.environ.get("MODE", "peers")   # peers | multicast

def run(peers: bool):
    ctx = Context(W, H, device=local)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    if peers:
        assert vp.enable_multicast(ctx) if MODE == "multicast" else vp.enable_peers(ctx)
    gb = ctx.trainer_grad_block()
    for it in range(2):
        ctx.trainer_accumulate(gc, tg, zero_grads=True, grad_scale=1.0 / views, want_loss=False)
        if peers:
            (vp.multicast_step if MODE == "multicast" else vp.peer_step)(ctx, it, 100)
        else:
            dist.all_reduce(gb)
            ctx.trainer_apply(it, 100)
    torch.cuda.synchronize()
    tt = ctx.trainer_tensors()
    out = ({k: v.cpu().numpy().copy() for k, v in tt["params"].items()}, tt["accum"].cpu().numpy().copy())
    dist.barrier()
    ctx.close()
    return out

pa, aa = run(False)
pb, ab = run(True)
ok = True
rel = float(np.abs(aa - ab).max() / max(np.abs(aa).max(), 1e-12))
print(f"[rank {rank}] D1 accumulator: max rel diff {rel:.2e}", flush=True)
ok &= rel < 1e-4
for k in pa:
    da, db = pa[k] - params[k].reshape(pa[k].shape), pb[k] - params[k].reshape(pb[k].shape)
    scale = max(np.abs(da).max(), 1e-12)
    bad = float((np.abs(da - db) > 1e-3 * scale).mean())
    print(f"[rank {rank}] {k}: elements off by > 1e-3 of the largest delta: {bad:.2e}", flush=True)
    ok &= bad < 2e-2
# replicas must hold bit-identical parameters (and D1 accumulators) after the fused step
pb = dict(pb, _accum=ab)
for k in pb:
    t = torch.from_numpy(pb[k].view(np.int32).astype(np.int64)).cuda().sum()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok &= bool(lo == hi)
    if rank == 0:
        print(f"{k}: replica checksums {'identical' if bool(lo == hi) else 'DIFFER'}", flush=True)
print(f"[rank {rank}] PEER_DP_{'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)

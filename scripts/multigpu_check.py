"""torchrun --nproc-per-node N scripts/multigpu_check.py : frame-sharded solve == single-GPU solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import multicam_calibration_b200 as mcc
from multicam_calibration_b200 import distributed
from multicam_calibration_b200.synthetic import make_scene

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 997
sc = make_scene(6, F, sigma=0.3, p_missing_view=0.2, seed=5)
args = sc.init_args()
np.random.seed(0)
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    e, i, p, use, res = mcc.bundle_adjust(*args, n_frames=None, ftol=1e-12, xtol=1e-12, verbose=0)
ok = True
if rank == 0:
    prob = mcc.BAProblem(sc.uvs[:, use], sc.objpoints, device=local)
    x0 = mcc.serialize_params(args[1], args[2], args[4][use])
    xs, rs = prob.solve(x0, ftol=1e-12, xtol=1e-12, verbose=0)
    d_cost = abs(rs.cost - res.cost) / rs.cost
    d_rms = abs(rs.rms - res.rms)
    d_x = np.abs(xs - res.x).max()
    print(f"world={world} frames={len(use)} sharded: cost {res.cost:.9f} it {res.iterations} rms {res.rms:.9f} | "
          f"single: cost {rs.cost:.9f} it {rs.iterations} | rel dcost {d_cost:.2e} drms {d_rms:.2e} max|dx| {d_x:.2e} "
          f"shard0 {res.shard} solve_ms {res.solve_ms:.2f} vs {rs.solve_ms:.2f} peer_memory={res.peer_memory}")
    ok = d_cost < 1e-9 and d_rms < 1e-8 and res.x.shape == xs.shape and res.success
    print("MULTIGPU_OK" if ok else "MULTIGPU_FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

"""torchrun --nproc-per-node N scripts/multigpu_check.py [frames]: the frame-sharded public call
against the single-GPU solve of the same problem (run by tests/test_multigpu.py; its output is kept
under profiles/ as the record of multi-GPU correctness).

Checks, on every rank where it applies: cost / RMS / parameters of the sharded solve equal the
single-GPU ones; ``use_frames`` and the calibration are identical on every rank although the ranks'
numpy RNGs are seeded differently (rank 0 draws the sub-sample); ``result.fun`` is the reference-order
residual vector; uneven shards; per-rank host->device traffic is the rank's shard."""
import contextlib
import io
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 997
sc = make_scene(6, F, sigma=0.3, p_missing_view=0.2, seed=5)
uvs = sc.uvs.copy()
uvs[:, 3] += 60.0                      # one gross outlier frame: excluded by the global 5 x median rule
args = (uvs,) + sc.init_args()[1:]
ok = True
lines = []


def same_on_all_ranks(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


for tag, nf in (("all", None), ("sub", 301), ("all again: cached problem and communicator", None)):
    np.random.seed(1000 + 17 * rank)                     # ranks disagree on purpose
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        e, i, p, use, res = mcc.bundle_adjust(*args, n_frames=nf, ftol=1e-12, xtol=1e-12, verbose=0)
    ident = same_on_all_ranks(use) and same_on_all_ranks(res.x) and same_on_all_ranks([res.cost])
    fun = res.fun
    if rank == 0:
        np.random.seed(1000)                              # what rank 0 drew
        with contextlib.redirect_stdout(io.StringIO()) as buf1:
            use1 = mcc.select_frames(*args, n_frames=nf)
        prob = mcc.BAProblem(uvs[:, use1], sc.objpoints, device=local)
        x0 = mcc.serialize_params(args[1], args[2], args[4][use1])
        xs, rs = prob.solve(x0, ftol=1e-12, xtol=1e-12, verbose=0)
        d_cost = abs(rs.cost - res.cost) / rs.cost
        d_rms = abs(rs.rms - res.rms)
        d_x = np.abs(xs - res.x).max()
        d_fun = np.abs(fun - prob.residuals(res.x)).max() if fun.shape == (prob.n_residuals,) else np.inf
        msg_same = buf.getvalue().strip().splitlines()[:1] == buf1.getvalue().strip().splitlines()[:1]
        good = (d_cost < 1e-9 and d_rms < 1e-8 and d_x < 1e-6 and d_fun < 1e-9 and np.array_equal(use, use1) and msg_same
                and 3 not in use and res.success and ident)
        lines.append(f"[{tag}] world={world} frames={len(use)} sharded: cost {res.cost:.9f} it {res.iterations} rms {res.rms:.9f} | "
                     f"single: cost {rs.cost:.9f} it {rs.iterations} | rel dcost {d_cost:.2e} drms {d_rms:.2e} max|dx| {d_x:.2e} "
                     f"max|dfun| {d_fun:.2e} use_frames equal {np.array_equal(use, use1)} message equal {msg_same} "
                     f"identical on all ranks {ident} shard0 {res.shard} solve_ms {res.solve_ms:.2f} vs {rs.solve_ms:.2f} "
                     f"collective={res.collective} peer_memory={res.peer_memory}")
        ok = ok and good
        prob.close()
if rank == 0:
    print("\n".join(lines))
    print("MULTIGPU_OK" if ok else "MULTIGPU_FAIL")
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)

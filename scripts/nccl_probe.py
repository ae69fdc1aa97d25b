"""torchrun --nproc-per-node 2 scripts/nccl_probe.py : host-side cost of the per-step NCCL all-reduce."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from multicam_calibration_b200 import _native, distributed
from multicam_calibration_b200.engine import BAProblem
from multicam_calibration_b200.synthetic import make_scene
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
comm = (distributed.broadcast_unique_id(), rank, world)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
sc = make_scene(6, F, sigma=0.5, p_missing_view=0.2, seed=0, shard=rank)
for use_comm in (True, False):
    prob = BAProblem(sc.uvs, sc.objpoints, device=local, comm=comm if use_comm else None)
    lib, h = prob.lib, prob._h
    null = ctypes.c_void_p()
    with torch.cuda.device(local), torch.cuda.stream(prob.stream):
        d_x = torch.as_tensor(sc.x0()).cuda()
        def step():
            _native.check(lib.mcba_build_reduced(h, ctypes.c_void_p(d_x.data_ptr()), 1e-3, 1, 1.0, null, null, null, null))
        for _ in range(5): step()
        torch.cuda.synchronize(); dist.barrier()
        host = []
        t0 = time.perf_counter()
        for _ in range(20):
            t = time.perf_counter(); step(); host.append(time.perf_counter() - t)
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        if rank == 0:
            print(f"comm={use_comm}: enqueue {t_enq/20*1e3:.3f} ms/step, total {t_all/20*1e3:.3f} ms/step, "
                  f"host per call min {min(host)*1e3:.3f} max {max(host)*1e3:.3f} ms", flush=True)
    # torch's own all_reduce on the same size for comparison
    t = torch.zeros(5408, dtype=torch.float64, device="cuda")
    for _ in range(5): dist.all_reduce(t)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): dist.all_reduce(t)
    torch.cuda.synchronize()
    if rank == 0: print(f"torch all_reduce 43 KB: {(time.perf_counter()-t0)/20*1e3:.3f} ms each", flush=True)
    prob.close()
dist.destroy_process_group()

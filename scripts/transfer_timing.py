import time, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from multicam_calibration_b200 import _native
a = np.random.default_rng(0).standard_normal(21_000_000)   # 168 MB pageable
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d = _native.to_device(a); torch.cuda.synchronize(); t1 = time.perf_counter()
    h = _native.to_host(d); t2 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.2f} ms ({a.nbytes/(t1-t0)/1e9:.1f} GB/s)  download {1e3*(t2-t1):.2f} ms ({a.nbytes/(t2-t1)/1e9:.1f} GB/s)")

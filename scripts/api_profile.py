"""cProfile of the public bundle_adjust(...) call at BASELINE configs[2]: where the host time goes."""
import cProfile, io, contextlib, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene

F = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
sc = make_scene(6, F, sigma=0.5, p_missing_view=0.2, seed=0)
args = sc.init_args()
for rep in range(3):
    np.random.seed(0)
    buf = io.StringIO()
    pr = cProfile.Profile()
    with contextlib.redirect_stdout(buf):
        pr.enable()
        out = mcc.bundle_adjust(*args, n_frames=None, verbose=0)
        torch.cuda.synchronize()
        pr.disable()
    del out
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())

import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene
out = {}
for C, F in [(6, 3000), (3, 500), (16, 1000), (1, 100), (7, 333)]:
    sc = make_scene(C, F, sigma=0.4, p_missing_view=0.25, seed=C * 1000 + F)
    x, res = mcc.BAProblem(sc.uvs, sc.objpoints).solve(sc.x0(), verbose=0)
    out[f"{C}x{F}"] = np.concatenate([[res.cost, res.optimality, res.nfev, res.rms], res.grad, x])
np.savez(sys.argv[1], **out)

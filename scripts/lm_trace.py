"""Print the LM iteration table for a synthetic scene (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene

F = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-10
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = float(v)
sc = make_scene(6, F, sigma=0.5, p_missing_view=0.2, seed=0)
prob = mcc.BAProblem(sc.uvs, sc.objpoints)
x, res = prob.solve(sc.x0(), ftol=tol, xtol=tol if tol < 1e-4 else 1e-8, verbose=2, **kw)
print({k: res[k] for k in ("cost", "rms", "nfev", "njev", "iterations", "status", "optimality", "solve_ms", "damping")})

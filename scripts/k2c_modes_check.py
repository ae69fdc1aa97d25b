"""K2c variants against each other (MCBA_K2C_MODE is read once per process): run with
   python scripts/k2c_modes_check.py dump <mode-tag>    in three processes, then    ... compare
Compares S, b, g, cost, a damped step and a converged solve at several shapes to rounding level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
SHAPES = [(6, 40), (6, 3000), (3, 500), (7, 333), (16, 2000), (1, 100)]
def dump(tag):
    import multicam_calibration_b200 as mcc
    from multicam_calibration_b200.synthetic import make_scene
    out = {}
    for C, F in SHAPES:
        sc = make_scene(C, F, sigma=0.4, p_missing_view=0.25, seed=C * 1000 + F)
        x0 = sc.x0()
        prob = mcc.BAProblem(sc.uvs, sc.objpoints)
        S, b, g, cost = prob.build_reduced(x0, lam=1e-3)
        step = prob.solve_step(1e-3)
        grad = prob.gradient()
        x, res = prob.solve(x0, ftol=1e-12, xtol=1e-12, verbose=0)
        k = f"{C}x{F}"
        out.update({k + "_S": S, k + "_b": b, k + "_g": g, k + "_cost": cost, k + "_step": step, k + "_grad": grad,
                    k + "_x": x, k + "_fc": res.cost, k + "_nfev": res.nfev, k + "_opt": res.optimality})
    np.savez(f"gpurun_out/k2c_mode_{tag}.npz", **out)
def compare():
    tags = ["general", "ring", "stream"]
    d = {t: np.load(f"gpurun_out/k2c_mode_{t}.npz") for t in tags}
    for C, F in SHAPES:
        k = f"{C}x{F}"
        line = [k]
        for t in tags[1:]:
            for q in ("S", "b", "g", "step", "grad", "x"):
                a, r = d[t][k + "_" + q], d["general"][k + "_" + q]
                line.append(f"{t}.{q} {np.abs(a - r).max() / max(np.abs(r).max(), 1e-300):.1e}")
            line.append(f"{t}.cost {abs(float(d[t][k + '_cost']) - float(d['general'][k + '_cost'])) / float(d['general'][k + '_cost']):.1e}")
            line.append(f"{t}.nfev {int(d[t][k + '_nfev'])}/{int(d['general'][k + '_nfev'])} opt {float(d[t][k + '_opt']):.1e}/{float(d['general'][k + '_opt']):.1e}")
        print("  ".join(line))
if sys.argv[1] == "dump": dump(sys.argv[2])
else: compare()

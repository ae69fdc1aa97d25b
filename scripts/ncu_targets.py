"""Every hot kernel of the path once or twice, for an Nsight Compute capture:

    ncu --set full --clock-control none --import-source on \
        -k regex:'k2p_kernel|k2c_|k2_syrk|finalize_kernel|solve_reduced|backsub_kernel|sum_scalars|residual_chunks|triangulate_kernel|project_points_multi|homography_transfer' \
        -c 48 -o gpurun_out/r02_full python scripts/ncu_targets.py
    python scripts/ncu_summary.py gpurun_out/r02_full.ncu-rep "<command>" > profiles/r02_ncu_full_summary.json

BASELINE configs[2] (6 cameras x 50,000 frames): three LM evaluations (K2p, K2c, SYRK, finalize, solve,
back-substitution, step scalars), K1 residuals; configs[4]: triangulate and project_points of 1M
keypoints x 6 cameras; the reprojection-error QC on 6 x 5,000 frames."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene, make_keypoints

C = int(sys.argv[1]) if len(sys.argv) > 1 else 6
F = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
sc = make_scene(C, F, sigma=0.5, p_missing_view=0.2, seed=0)
prob = mcc.BAProblem(sc.uvs, sc.objpoints)
x0 = sc.x0()
x, res = prob.solve(x0, verbose=0, max_nfev=3)
r = prob.residuals(x)
print(f"LM: nfev {res.nfev} cost {res.cost:.6e}; residuals {r.size}")
prob.close()
uvs, ext, intr, pts = make_keypoints(1_000_000, 6, sigma=0.3, p_missing=0.2, seed=0)
X = mcc.triangulate(uvs, list(ext), intr)
uv = mcc.project_points_multi(pts, list(ext), intr)
print("triangulated", int(np.isfinite(X).all(1).sum()), "projected", uv.shape)
q = make_scene(6, 5000, sigma=0.3, p_missing_view=0.2, seed=1)
e, i = q._split(q.true_cams)
med, _, _ = mcc.reprojection_residuals(q.uvs, e, i, q.objpoints, q.true_poses)
print("QC median error per camera", np.round(med, 4))

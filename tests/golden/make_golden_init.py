"""Golden fixtures for the initialisation algebra from the UNMODIFIED
reference ``multicam_calibration.calibration`` (calibration.py:116-277), imported through the
stub package of make_golden.py.  Build container only (needs /root/reference and networkx):

    python tests/golden/make_golden_init.py [--check]

Also asserts that ``oracle/np_oracle.py`` reproduces the reference on every fixture and that the
host-side spanning tree of the package (no networkx) orders edges like the reference, ties
included.
"""
import importlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import check_fixture, close, load_reference   # noqa: E402


def main():
    from oracle import np_oracle as orc
    from multicam_calibration_b200.synthetic import make_camera_poses
    from multicam_calibration_b200.calibration import get_camera_spanning_tree
    geo, _ = load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cal = importlib.import_module("multicam_calibration.calibration")
    import networkx
    out = {"versions": np.array([np.__version__, networkx.__version__])}
    for tag, C, F, p in (("a", 5, 400, 0.25), ("b", 9, 150, 0.6)):   # odd rings: no camera pair exactly 180 deg apart
        poses, ext_true, board_true = make_camera_poses(C, F, p_missing=p, seed=3)
        poses[:, 5] = np.nan                      # a frame nobody saw
        poses[1:, 6] = np.nan                     # a frame only camera 0 saw
        pair = cal.estimate_pairwise_camera_transform(poses[0], poses[1])
        ext, tree = cal.estimate_all_extrinsics(poses, root=0)
        ext_r2, tree_r2 = cal.estimate_all_extrinsics(poses, root=2)
        cons = cal.consensus_calib_poses(poses, ext)
        close(pair, orc.estimate_pairwise_camera_transform(poses[0], poses[1]), 1e-12, f"{tag}: pairwise transform")
        e_o, t_o = orc.estimate_all_extrinsics(poses, root=0)
        assert [tuple(map(int, e)) for e in t_o] == [tuple(map(int, e)) for e in tree]
        close(ext, e_o, 1e-12, f"{tag}: all extrinsics")
        close(cons, orc.consensus_calib_poses(poses, ext), 1e-12, f"{tag}: consensus poses")
        assert get_camera_spanning_tree(poses, root=0) == [tuple(map(int, e)) for e in tree]
        assert get_camera_spanning_tree(poses, root=2) == [tuple(map(int, e)) for e in tree_r2]
        # the estimate is close to the truth the poses were generated from
        dT = orc.transformation_matrix(ext) - orc.transformation_matrix(ext_true)   # vectors are ambiguous near pi
        assert np.abs(dT[:, :3, :3]).max() < 0.05 and np.abs(dT[:, :3, 3]).max() < 40, np.abs(dT).max((0, 1))
        out.update({f"{tag}_poses": poses, f"{tag}_pair01": pair, f"{tag}_ext": ext, f"{tag}_tree": np.array(tree),
                    f"{tag}_ext_root2": ext_r2, f"{tag}_tree_root2": np.array(tree_r2), f"{tag}_consensus": cons,
                    f"{tag}_ext_true": ext_true, f"{tag}_board_true": board_true})
    # spanning-tree ordering under heavy ties (co-detection counts from few frames), many roots
    rng = np.random.default_rng(0)
    n_checked = 0
    for trial in range(200):
        C, F = int(rng.integers(2, 9)), int(rng.integers(1, 6))
        poses = rng.normal(size=(C, F, 6))
        poses[rng.random((C, F)) < 0.5] = np.nan
        for root in range(C):
            ref_tree = [tuple(map(int, e)) for e in cal.get_camera_spanning_tree(poses, root=root)]
            assert get_camera_spanning_tree(poses, root=root) == ref_tree, (trial, root)
            n_checked += 1
    print(f"  spanning tree: {n_checked} tied graphs ordered like the reference")
    # rodrigues_inv / get_transformation_vector
    R = geo.rodrigues(rng.normal(0, 1.0, (64, 3)))
    T = geo.get_transformation_matrix(rng.normal(0, 1.0, (64, 6)))
    out.update({"R": R, "R_vec": geo.rodrigues_inv(R), "T": T, "T_vec": geo.get_transformation_vector(T)})
    close(out["R_vec"], orc.rodrigues_inv(R), 1e-15, "rodrigues_inv")
    close(out["T_vec"], orc.transformation_vector(T), 1e-15, "get_transformation_vector")
    if "--check" in sys.argv[1:]:
        check_fixture("init", out)
    else:
        np.savez_compressed(os.path.join(HERE, "init.npz"), **out)
        print("wrote init.npz")


if __name__ == "__main__":
    main()

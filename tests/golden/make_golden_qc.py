"""Golden fixtures for the two rows next to the hot path that the reference implements in modules it
cannot import here without help:

* ``qc.npz``: ``multicam_calibration.viz.plot_residuals`` (viz.py:70-210; numeric core :155-177:
  project + undistort + per-frame ``cv2.findHomography`` transfer) run UNMODIFIED with matplotlib,
  vidio, imageio and h5py replaced by mocks (the numeric core touches none of them);
* ``calibration_ref.json`` / ``calibration_ref_jarvis/``: what ``multicam_calibration.io.save_calibration``
  (io.py:8-99) writes for a known calibration, and ``io_loaded.npz``: what ``load_calibration``
  (:102-245) reads back from the jarvis directory (the reference's own JSON loader cannot read its own
  JSON files: it looks for 'rotation'/'translation' where the writer stored 'R'/'T', :59-60 vs :161-164).

Run in the build container only (needs ``/root/reference`` and OpenCV):

    python tests/golden/make_golden_qc.py [--check]
"""
import importlib
import json
import os
import shutil
import sys
import types
import warnings
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/multicam_calibration"


def load_reference():
    for name in ("vidio", "vidio.read", "matplotlib", "matplotlib.pyplot", "imageio", "h5py"):
        sys.modules.setdefault(name, mock.MagicMock())
    plt = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].pyplot = plt
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    pkg = types.ModuleType("multicam_calibration")
    pkg.__path__ = [REF]
    sys.modules["multicam_calibration"] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        viz = importlib.import_module("multicam_calibration.viz")
        rio = importlib.import_module("multicam_calibration.io")
    return viz, rio


def qc_scene():
    from multicam_calibration_b200.synthetic import make_scene
    sc = make_scene(6, 48, sigma=0.3, p_missing_view=0.2, p_missing_corner=0.01, seed=11)
    ext, intr = sc._split(sc.true_cams)
    intr = [(K, np.array([d[0], d[1], 1e-4 * (c + 1), -2e-4, 1e-3])) for c, (K, d) in enumerate(intr)]   # all five coefficients in play
    poses = sc.true_poses + np.random.default_rng(3).normal(0, [1e-3] * 3 + [0.2] * 3, sc.true_poses.shape)
    return sc.uvs, ext, intr, poses, sc.objpoints


def build():
    import cv2
    viz, rio = load_reference()
    out = {}
    uvs, ext, intr, poses, obj = qc_scene()
    _, median_error, reproj, transformed = viz.plot_residuals(uvs, ext, intr, obj, poses)
    out["qc"] = dict(uvs=uvs, ext=ext, K=np.stack([K for K, _ in intr]), dist=np.stack([d for _, d in intr]), poses=poses,
                     objpoints=obj, median_error=median_error, reprojections=reproj, transformed_reprojections=transformed,
                     versions=np.array([np.__version__, cv2.__version__]))
    # ---- calibration files
    names = [f"cam{c}" for c in (3, 0, 5, 1, 4, 2)]      # deliberately not sorted
    tmp = os.path.join(HERE, "_tmp_io")
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    rio.save_calibration(ext, intr, names, os.path.join(tmp, "calibration_ref"), save_format="json")
    rio.save_calibration(ext, intr, names, os.path.join(tmp, "calibration_ref_jarvis"), save_format="jarvis")
    e2, i2, n2 = rio.load_calibration(os.path.join(tmp, "calibration_ref_jarvis"), load_format="jarvis")
    out["io_loaded"] = dict(ext=np.stack(e2), K=np.stack([K for K, _ in i2]), dist=np.stack([d for _, d in i2]), names=np.array(n2),
                            saved_ext=ext, saved_K=np.stack([K for K, _ in intr]), saved_dist=np.stack([d for _, d in intr]),
                            saved_names=np.array(names))
    out["_files"] = tmp
    return out


def main():
    check = "--check" in sys.argv
    out = build()
    tmp = out.pop("_files")
    for name, arrays in out.items():
        path = os.path.join(HERE, name + ".npz")
        if check:
            old = np.load(path, allow_pickle=False)
            for k, v in arrays.items():
                a, b = np.asarray(v), old[k]
                assert a.shape == b.shape and (a.dtype.kind in "US" and (a == b).all() or np.array_equal(a, b, equal_nan=True)), (name, k)
            print(f"{name}: regenerated == committed")
        else:
            np.savez_compressed(path, **arrays)
            print("wrote", path, {k: np.asarray(v).shape for k, v in arrays.items()})
    # text fixtures
    for rel in ["calibration_ref.json"] + [os.path.join("calibration_ref_jarvis", f) for f in sorted(os.listdir(os.path.join(tmp, "calibration_ref_jarvis")))]:
        src, dst = os.path.join(tmp, rel), os.path.join(HERE, rel)
        if check:
            assert open(src).read() == open(dst).read(), rel
        else:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    if check:
        print("calibration files: regenerated == committed")
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

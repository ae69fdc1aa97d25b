"""Calibration files: drop-in for ``multicam_calibration.io`` (reference io.py:8-245).

Same two functions, formats and on-disk conventions (world -> camera rotation matrices and 3x1
translations, 3x3 camera matrices, ``k1, k2, p1, p2, k3``), so files written by either package are
read by the other -- with one deliberate difference: the reference's JSON loader looks for the keys
``rotation`` / ``translation`` (io.py:161-164) while its writer stores ``R`` / ``T`` (io.py:59-60),
so it cannot read its own files.  :func:`load_calibration` accepts both spellings; the writer keeps
the reference's ``R`` / ``T``.

Host code only (the step after the solve; no device content).  ``jarvis`` needs OpenCV's
``FileStorage`` and ``gimbal`` needs ``h5py``; each is imported when that format is used.
"""
import json
import os

import numpy as np

from .geometry import get_transformation_matrix, rodrigues_inv

FORMATS = ("json", "jarvis", "gimbal")


def _need(module, fmt):
    try:
        return __import__(module)
    except ImportError as e:   # loud: there is no silent alternative writer
        raise ImportError(f"calibration format {fmt!r} needs the {module!r} module") from e


def _vector(R, T):
    return np.concatenate([rodrigues_inv(np.asarray(R, dtype=float)), np.asarray(T, dtype=float).reshape(3)])


def save_calibration(all_extrinsics, all_intrinsics, camera_names, save_path, save_format="json"):
    """Write extrinsics ``(C,6)``, intrinsics ``[(K, dist5)]`` and names in ``save_format``
    (io.py:8-99): ``json`` one file, ``jarvis`` one OpenCV YAML per camera in a directory (matrices
    transposed relative to JSON), ``gimbal`` one HDF5 file with a ``camera_parameters`` group."""
    if not (len(all_extrinsics) == len(all_intrinsics) == len(camera_names)):
        raise AssertionError("Number of camera names must match number of extrinsics and intrinsics")
    if save_format not in FORMATS:
        raise ValueError(f"Unknown format {save_format}")
    T = get_transformation_matrix(np.array(all_extrinsics))
    Ks = [np.asarray(K) for K, _ in all_intrinsics]
    ds = [np.asarray(d) for _, d in all_intrinsics]

    if save_format == "json":
        data = {name: {"R": T[i, :3, :3].tolist(), "T": T[i, :3, 3:].tolist(),
                       "camera_matrix": Ks[i].tolist(), "distortion_coefs": ds[i].tolist()}
                for i, name in enumerate(camera_names)}
        if not save_path.endswith(".json"):
            save_path += ".json"
        with open(save_path, "w") as f:
            json.dump(data, f, indent=4)
    elif save_format == "jarvis":
        cv2 = _need("cv2", save_format)
        os.makedirs(save_path, exist_ok=True)
        for i, name in enumerate(camera_names):
            fs = cv2.FileStorage(os.path.join(save_path, f"{name}.yaml"), cv2.FILE_STORAGE_WRITE)
            fs.write("intrinsicMatrix", np.ascontiguousarray(Ks[i].T))
            fs.write("distortionCoefficients", ds[i].reshape(1, -1))
            fs.write("R", np.ascontiguousarray(T[i, :3, :3].T))
            fs.write("T", np.ascontiguousarray(T[i, :3, 3:]))
            fs.release()
    else:
        h5py = _need("h5py", save_format)
        if not save_path.endswith(".h5"):
            save_path += ".h5"
        with h5py.File(save_path, "w") as h5:
            grp = h5.create_group("camera_parameters")
            grp.create_dataset("dist_coefs", data=np.stack(ds))
            grp.create_dataset("intrinsic", data=np.stack(Ks))
            grp.create_dataset("rotation", data=T[:, :3, :3])
            grp.create_dataset("translation", data=T[:, :3, 3])
            grp.create_dataset("camera_names", data=camera_names)


def load_calibration(load_path, load_format="json", camera_names=None):
    """Read a calibration back (io.py:102-245): returns ``(all_extrinsics list of (6,), all_intrinsics
    list of (K, dist), camera_names)``.  ``camera_names`` fixes the order (and, for jarvis / gimbal,
    selects a subset); ``None`` means alphabetical (json, jarvis) or file order (gimbal)."""
    if load_format not in FORMATS:
        raise ValueError(f"Unknown format {load_format}")
    if load_format == "json":
        with open(load_path, "r") as f:
            data = json.load(f)
        if camera_names is None:
            camera_names = sorted(data.keys())
        elif set(camera_names) != set(data.keys()):
            raise AssertionError("Camera names must match keys in calibration file")
        ext, intr = [], []
        for name in camera_names:
            cam = data[name]
            R = cam["R"] if "R" in cam else cam["rotation"]          # the reference writes R / T and reads rotation / translation
            t = cam["T"] if "T" in cam else cam["translation"]
            ext.append(_vector(R, t))
            intr.append((np.array(cam["camera_matrix"]), np.array(cam["distortion_coefs"])))
        return ext, intr, camera_names
    if load_format == "jarvis":
        cv2 = _need("cv2", load_format)
        files = {os.path.splitext(f)[0]: f for f in sorted(os.listdir(load_path))
                 if os.path.splitext(f)[1] in (".yaml", ".YAML")}
        if camera_names is None:
            camera_names = sorted(files)
        elif not set(camera_names) <= set(files):
            raise AssertionError("Camera names must be a subset of yaml files in calibration directory")
        ext, intr = [], []
        for name in camera_names:
            fs = cv2.FileStorage(os.path.join(load_path, files[name]), cv2.FILE_STORAGE_READ)
            ext.append(_vector(fs.getNode("R").mat().T, fs.getNode("T").mat()))
            intr.append((fs.getNode("intrinsicMatrix").mat().T, fs.getNode("distortionCoefficients").mat().squeeze()))
            fs.release()
        return ext, intr, camera_names
    h5py = _need("h5py", load_format)
    if not load_path.endswith(".h5"):
        load_path += ".h5"
    with h5py.File(load_path, "r") as h5:
        grp = h5["camera_parameters"]
        names = [n.decode("utf-8") if isinstance(n, bytes) else str(n) for n in grp["camera_names"][()].tolist()]
        intr = list(zip(grp["intrinsic"][()], grp["dist_coefs"][()]))
        ext = np.concatenate([rodrigues_inv(grp["rotation"][()]), grp["translation"][()]], axis=1)
    if camera_names is None:
        camera_names = names
    else:
        if not set(camera_names) <= set(names):
            raise AssertionError("Camera names must be a subset of names in calibration file")
        ix = [names.index(n) for n in camera_names]
        ext = ext[ix]
        intr = [intr[i] for i in ix]
    return list(ext), intr, camera_names

"""CPU: calibration loaders (mvdetr_b200/calibration.py) on files written in the datasets' own formats -- OpenCV
FileStorage XML written by cv2 itself and Wildtrack's plain-text extrinsic XML -- against the generating values,
cv2.Rodrigues, and (in the build container only) the reference's own dataset classes."""
import os
import sys

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from mvdetr_b200 import calibration as cal  # noqa: E402


def _write_scene(root, names, intr_dir, wildtrack, seed):
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "calibrations", intr_dir))
    os.makedirs(os.path.join(root, "calibrations", "extrinsic"))
    truth = []
    for c in names:
        K = np.array([[1700 + rng.rand() * 100, 0, 960 + rng.randn()], [0, 1700 + rng.rand() * 100, 540 + rng.randn()],
                      [0, 0, 1]])
        rvec = rng.randn(3) * 0.8
        tvec = rng.randn(3) * (500 if wildtrack else 5)
        fs = cv2.FileStorage(os.path.join(root, "calibrations", intr_dir, f"intr_{c}.xml"), cv2.FILE_STORAGE_WRITE)
        fs.write("camera_matrix", K)
        fs.write("distortion_coefficients", np.zeros((5, 1)))
        fs.release()
        path = os.path.join(root, "calibrations", "extrinsic", f"extr_{c}.xml")
        if wildtrack:  # the dataset's hand-written format: numbers as text
            rvec, tvec = rvec.astype(np.float32), tvec.astype(np.float32)
            with open(path, "w") as f:
                f.write('<?xml version="1.0"?>\n<opencv_storage>\n<rvec>%s</rvec>\n<tvec>%s</tvec>\n</opencv_storage>\n'
                        % (" ".join(repr(float(v)) for v in rvec), " ".join(repr(float(v)) for v in tvec)))
        else:
            fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
            fs.write("rvec", rvec.reshape(3, 1))
            fs.write("tvec", tvec.reshape(3, 1))
            fs.release()
        truth.append((K, np.asarray(rvec, dtype=np.float64), np.asarray(tvec, dtype=np.float64)))
    return truth


@pytest.mark.parametrize("which", ["wildtrack", "multiviewx"])
def test_loaders_read_the_dataset_formats(tmp_path, which):
    wild = which == "wildtrack"
    names = cal.WILDTRACK_CAMERAS if wild else cal.MULTIVIEWX_CAMERAS
    truth = _write_scene(str(tmp_path), names, "intrinsic_zero" if wild else "intrinsic", wild, seed=3)
    ds = (cal.load_wildtrack if wild else cal.load_multiviewx)(str(tmp_path))
    assert ds.num_cam == len(names)
    assert ds.Rimg_shape == [90, 160] and ds.Rworld_shape == ([120, 360] if wild else [160, 250])
    for (K, rvec, tvec), Ki, Rt in zip(truth, ds.base.intrinsic_matrices, ds.base.extrinsic_matrices):
        assert np.allclose(Ki, K, rtol=0, atol=1e-9)
        R_cv, _ = cv2.Rodrigues(rvec)
        assert np.allclose(Rt[:, :3], R_cv, atol=1e-12) and np.allclose(Rt[:, 3], tvec, atol=1e-12)
        assert np.allclose(Rt[:, :3] @ Rt[:, :3].T, np.eye(3), atol=1e-12)
    assert np.allclose(cal.rodrigues(np.zeros(3)), np.eye(3))


def test_loaded_scene_drives_the_projection_chain(tmp_path):
    """The loaded namespace plugs into projection.* / MultiviewFusion set-up like the synthetic scenes do."""
    from mvdetr_b200.projection import create_reference_map, world_grid_projection_mats
    _write_scene(str(tmp_path), cal.MULTIVIEWX_CAMERAS, "intrinsic", False, seed=5)
    ds = cal.load_multiviewx(str(tmp_path))
    mats = world_grid_projection_mats(ds)
    assert tuple(mats.shape) == (6, 3, 3) and bool(np.isfinite(np.asarray(mats)).all())
    ref = create_reference_map(ds, 4)
    assert tuple(ref.shape) == (80 * 125, 6, 4, 2)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("which", ["wildtrack", "multiviewx"])
def test_same_matrices_as_the_reference_dataset_classes(tmp_path, which, monkeypatch):
    """multiview_detector/datasets/{Wildtrack,MultiviewX}.py on the same files (np.float was removed from numpy; the
    reference still spells it, so it is aliased for the call)."""
    wild = which == "wildtrack"
    names = cal.WILDTRACK_CAMERAS if wild else cal.MULTIVIEWX_CAMERAS
    _write_scene(str(tmp_path), names, "intrinsic_zero" if wild else "intrinsic", wild, seed=7)
    monkeypatch.setattr(np, "float", float, raising=False)
    import importlib.util
    name = "Wildtrack" if wild else "MultiviewX"  # loaded by path: the package __init__ pulls in kornia / matplotlib
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", f"/root/reference/multiview_detector/datasets/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = getattr(mod, name)(str(tmp_path))
    ds = (cal.load_wildtrack if wild else cal.load_multiviewx)(str(tmp_path))
    for a, b in zip(ds.base.intrinsic_matrices, ref.intrinsic_matrices):
        assert np.allclose(a, b, atol=1e-9)
    for a, b in zip(ds.base.extrinsic_matrices, ref.extrinsic_matrices):
        assert np.allclose(a, b, atol=1e-6)   # cv2.Rodrigues runs in float32 for Wildtrack's float32 rvec
    assert np.allclose(ds.base.worldcoord_from_worldgrid_mat, ref.worldcoord_from_worldgrid_mat)
    assert np.allclose(ds.base.world_indexing_from_xy_mat, ref.world_indexing_from_xy_mat)
    assert ds.base.worldcoord_unit == ref.worldcoord_unit and ds.base.indexing == ref.indexing

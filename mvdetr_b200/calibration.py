"""Camera calibration loaders for the two datasets MVDeTr ships with, so that REAL Wildtrack / MultiviewX calibration
can drive the fusion path (input side). Host-side only; no images are read here.

The directory layouts and file formats are the datasets' own (as the reference reads them):
  Wildtrack    <root>/calibrations/intrinsic_zero/intr_{CVLab1..4,IDIAP1..3}.xml   OpenCV FileStorage, node camera_matrix
               <root>/calibrations/extrinsic/extr_*.xml                            plain XML, <rvec>/<tvec> as text
               ref: multiview_detector/datasets/Wildtrack.py:8-11,21-34,79-100
  MultiviewX   <root>/calibrations/intrinsic/intr_Camera{1..6}.xml                 OpenCV FileStorage, node camera_matrix
               <root>/calibrations/extrinsic/extr_Camera{1..6}.xml                 OpenCV FileStorage, nodes rvec, tvec
               ref: multiview_detector/datasets/MultiviewX.py:8-11,21-34,79-98
The objects returned expose exactly what the model set-up reads from the reference's `frameDataset`
(ref: multiview_detector/models/mvdetr.py:34,46-56,78-95; multiview_detector/datasets/frameDataset.py:57-71):
    num_cam, img_shape, worldgrid_shape, Rimg_shape, Rworld_shape, img_reduce, world_reduce,
    base.{intrinsic_matrices, extrinsic_matrices, worldcoord_from_worldgrid_mat, world_indexing_from_xy_mat,
          worldcoord_unit, indexing}
and plug into mvdetr_b200.fusion.MultiviewFusion / projection.* unchanged (same attributes as synthetic.py).

Parsing is dependency-free (ElementTree + numpy; OpenCV's XML FileStorage is plain XML with
<name type_id="opencv-matrix"><rows/><cols/><dt/><data/></name>), and the Rodrigues rotation is computed here.
"""
import os
import types
import xml.etree.ElementTree as ET

import numpy as np

WILDTRACK_CAMERAS = ["CVLab1", "CVLab2", "CVLab3", "CVLab4", "IDIAP1", "IDIAP2", "IDIAP3"]
MULTIVIEWX_CAMERAS = ["Camera%d" % i for i in range(1, 7)]


def rodrigues(rvec):
    """Axis-angle vector -> 3x3 rotation matrix (the forward direction of cv2.Rodrigues), float64."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = float(np.linalg.norm(r))
    if theta < 1e-12:
        return np.eye(3)
    k = r / theta
    K = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return np.cos(theta) * np.eye(3) + (1.0 - np.cos(theta)) * np.outer(k, k) + np.sin(theta) * K


def _numbers(text):
    return np.array([float(t) for t in text.replace(",", " ").split()], dtype=np.float64)


def read_opencv_matrix(path, node):
    """Reads node `node` of an OpenCV XML FileStorage file (type_id="opencv-matrix") as a float64 array [rows, cols]."""
    root = ET.parse(path).getroot()
    elem = root.find(node)
    if elem is None:
        raise KeyError(f"{path}: no <{node}> node")
    rows, cols = int(elem.findtext("rows")), int(elem.findtext("cols"))
    data = _numbers(elem.findtext("data"))
    if data.size != rows * cols:
        raise ValueError(f"{path}: <{node}> has {data.size} values for a {rows}x{cols} matrix")
    return data.reshape(rows, cols)


def read_text_vector(path, node):
    """Reads <node>a b c</node> (Wildtrack's extrinsic files) as float64. Values pass through float32 like the
    reference's parser does (Wildtrack.py:91-95)."""
    root = ET.parse(path).getroot()
    elem = root.find(node)
    if elem is None or elem.text is None:
        raise KeyError(f"{path}: no <{node}> node")
    return _numbers(elem.text).astype(np.float32).astype(np.float64)


def _extrinsic(rvec, tvec):
    return np.hstack((rodrigues(rvec), np.asarray(tvec, dtype=np.float64).reshape(3, 1)))


def _scene(name, num_cam, img_shape, worldgrid_shape, indexing, unit, grid_mat, intrinsics, extrinsics, world_reduce,
           img_reduce):
    base = types.SimpleNamespace()
    base.__name__ = name
    base.num_cam = num_cam
    base.img_shape, base.worldgrid_shape = list(img_shape), list(worldgrid_shape)
    base.indexing = indexing
    base.worldcoord_unit = unit
    base.world_indexing_from_xy_mat = (np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]], dtype=float) if indexing == "ij"
                                       else np.eye(3))
    base.worldcoord_from_worldgrid_mat = np.asarray(grid_mat, dtype=float)
    base.intrinsic_matrices, base.extrinsic_matrices = tuple(intrinsics), tuple(extrinsics)
    ds = types.SimpleNamespace()
    ds.base = base
    ds.num_cam = num_cam
    ds.img_shape, ds.worldgrid_shape = base.img_shape, base.worldgrid_shape
    ds.world_reduce, ds.img_reduce = world_reduce, img_reduce
    ds.Rworld_shape = [s // world_reduce for s in worldgrid_shape]                      # frameDataset.py:70
    ds.Rimg_shape = np.ceil(np.array(img_shape) / img_reduce).astype(int).tolist()      # frameDataset.py:71
    return ds


def load_wildtrack(root, world_reduce=4, img_reduce=12):
    """Wildtrack: 7 cameras, 1080p, 480 x 1440 grid of 2.5 cm cells, ij indexing, centimetres."""
    cal = os.path.join(root, "calibrations")
    K = [read_opencv_matrix(os.path.join(cal, "intrinsic_zero", f"intr_{c}.xml"), "camera_matrix")
         for c in WILDTRACK_CAMERAS]
    Rt = [_extrinsic(read_text_vector(os.path.join(cal, "extrinsic", f"extr_{c}.xml"), "rvec"),
                     read_text_vector(os.path.join(cal, "extrinsic", f"extr_{c}.xml"), "tvec"))
          for c in WILDTRACK_CAMERAS]
    return _scene("Wildtrack", 7, [1080, 1920], [480, 1440], "ij", 0.01, [[2.5, 0, -300], [0, 2.5, -900], [0, 0, 1]],
                  K, Rt, world_reduce, img_reduce)


def load_multiviewx(root, world_reduce=4, img_reduce=12):
    """MultiviewX: 6 cameras, 1080p, 640 x 1000 grid of 2.5 cm cells, xy indexing, metres."""
    cal = os.path.join(root, "calibrations")
    K = [read_opencv_matrix(os.path.join(cal, "intrinsic", f"intr_{c}.xml"), "camera_matrix")
         for c in MULTIVIEWX_CAMERAS]
    Rt = [_extrinsic(read_opencv_matrix(os.path.join(cal, "extrinsic", f"extr_{c}.xml"), "rvec").reshape(3),
                     read_opencv_matrix(os.path.join(cal, "extrinsic", f"extr_{c}.xml"), "tvec").reshape(3))
          for c in MULTIVIEWX_CAMERAS]
    return _scene("MultiviewX", 6, [1080, 1920], [640, 1000], "xy", 1.0, [[0.025, 0, 0], [0, 0.025, 0], [0, 0, 1]],
                  K, Rt, world_reduce, img_reduce)

"""`inv_proj_mats.txt` without the adapted COLMAP.

The reference reads one 4x4 inverse projection matrix per image from `colmap/mvs/inv_proj_mats.txt`
(aggregate_2p5d_util.py:54-61); that file is written by the authors' COLMAP fork.  Its content is fully determined by
two things the Python side owns: the camera dict (`w, h, fx, fy, cx, cy, s, qw, qx, qy, qz, tx, ty, tz`,
colmap/extract_sfm.py:79-94) and `last_rows.txt` (reparam_depth.py:171-174): P4 = [K [R | t]; last_row]
(reparam_depth.py:120-141), M = inv(P4).  These helpers rebuild the file from those inputs.
"""
import os

import numpy as np


def quaternion_to_rotation(qw, qx, qy, qz):
    """Rotation matrix of the unit quaternion (w, x, y, z) (pyquaternion's `rotation_matrix`, reparam_depth.py:117)."""
    q = np.array([qw, qx, qy, qz], dtype=np.float64)
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def proj_mat_4by4(camera_params, last_row):
    """camera_params = (w, h, fx, fy, cx, cy, s, qw, qx, qy, qz, tx, ty, tz); last_row = the image's 4-vector."""
    _, _, fx, fy, cx, cy, s, qw, qx, qy, qz, tx, ty, tz = [float(v) for v in camera_params]
    K = np.array([[fx, s, cx],
                  [0., fy, cy],
                  [0., 0., 1.]])
    R = quaternion_to_rotation(qw, qx, qy, qz)
    P_3by4 = np.dot(K, np.hstack((R, np.array([[tx], [ty], [tz]]))))
    return np.vstack((P_3by4, np.asarray(last_row, dtype=np.float64).reshape((1, 4))))


def read_last_rows(path):
    """`img_name v0 v1 v2 v3` per line (reparam_depth.py:171-174)."""
    rows = {}
    with open(path) as fp:
        for line in fp:
            tmp = line.split()
            if len(tmp) == 5:
                rows[tmp[0]] = np.array([float(v) for v in tmp[1:]])
    return rows


def inv_proj_mats_from_cameras(camera_dict, last_rows):
    """{img_name: 4x4 float64 inverse projection matrix} for every image that has a last row."""
    return {name: np.linalg.inv(proj_mat_4by4(camera_dict[name], last_rows[name]))
            for name in sorted(last_rows.keys()) if name in camera_dict}


def write_inv_proj_mats(mats, path):
    """One line per image: name + 16 row-major values, parsed by aggregate_2p5d_util.py:56-61."""
    with open(path, 'w') as fp:
        for name in sorted(mats.keys()):
            fp.write('{} {}\n'.format(name, ' '.join(repr(float(v)) for v in np.asarray(mats[name]).reshape(-1))))


def ensure_inv_proj_mats(mvs_dir, camera_dict):
    """Write <mvs_dir>/inv_proj_mats.txt from <mvs_dir>/last_rows.txt when the file is missing."""
    path = os.path.join(mvs_dir, 'inv_proj_mats.txt')
    if not os.path.exists(path):
        write_inv_proj_mats(inv_proj_mats_from_cameras(camera_dict, read_last_rows(os.path.join(mvs_dir,
                                                                                                'last_rows.txt'))), path)
    return path

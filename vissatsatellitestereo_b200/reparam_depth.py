"""Drop-in for the reference's reparam_depth.py: `reparam_depth(sparse_dir, save_dir, camera_model)` (:69-195), the
step that fixes the depth re-parametrisation the MVS depth maps (and hence inv_proj_mats.txt) are expressed in, plus
`inv_proj_mats.txt` without the adapted COLMAP.

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


# ---------------------------------------------------------------------------------------------------------------
# reparam_depth.py:42-195
# ---------------------------------------------------------------------------------------------------------------
def robust_depth_range(depth_range):
    """:42-66: per image, the [2 %, 98 %] order statistics of its depths stretched by (-10, +30) m; images without a
    point get (-1e20, -1e20)."""
    import logging
    for img_name in depth_range:
        if depth_range[img_name]:
            tmp = depth_range[img_name]
            max_val, min_val = max(tmp), min(tmp)
            logging.info('img_name: {}, depth min: {}, max: {}, ratio: {}'.format(img_name, min_val, max_val,
                                                                                max_val / min_val))
            tmp = sorted(tmp)
            cnt = len(tmp)
            min_depth, max_depth = tmp[int(0.02 * cnt)], tmp[int(0.98 * cnt)]
            min_depth_new, max_depth_new = min_depth - 10, max_depth + 30
            if max_depth_new <= min_depth_new:
                min_depth_new, max_depth_new = min_depth, max_depth
            depth_range[img_name] = (min_depth_new, max_depth_new)
        else:
            depth_range[img_name] = (-1e20, -1e20)
    return depth_range


def reparam_depth(sparse_dir, save_dir, camera_model='perspective'):
    """:69-195.  Reads the sparse model (text form), and writes raw_depth.txt, reparam_depth.txt, last_rows.txt,
    reference_plane.txt and depth_ranges.txt into save_dir, with the reference's line formats.
    The fourth row of every image's projection matrix is depth_min(image) * [0, 0, 1, -min_z] with min_z = 1st percentile
    of the sparse points' z minus 20 m (:97-105): the reference plane z = min_z, scaled per image."""
    import logging
    from .colmap.read_model import read_model
    assert (camera_model == 'perspective' or camera_model == 'pinhole')
    cameras, images, points3D = read_model(sparse_dir, ext='.txt')

    # per-image rotation / translation / intrinsics, once
    rot, tvec, Kmat = {}, {}, {}
    for img_id, im in images.items():
        rot[img_id] = quaternion_to_rotation(im.qvec[0], im.qvec[1], im.qvec[2], im.qvec[3])
        tvec[img_id] = im.tvec.reshape((3, 1))
        params = cameras[im.camera_id].params
        if camera_model == 'pinhole':
            fx, fy, cx, cy = params
            s = 0.
        else:
            fx, fy, cx, cy, s = params
        Kmat[img_id] = np.array([[fx, s, cx], [0., fy, cy], [0., 0., 1.]])

    depth_range = {im.name: [] for im in images.values()}
    z_values = []
    for p in points3D.values():                                              # :78-91
        x = p.xyz.reshape((3, 1))
        z_values.append(x[2, 0])
        for img_id in p.image_ids:
            depth = (np.dot(rot[img_id], x) + tvec[img_id])[2, 0]
            if depth > 0:
                depth_range[images[img_id].name].append(depth)
    depth_range = robust_depth_range(depth_range)

    margin = 20.0                                                            # :96-99
    min_z_value = np.percentile(z_values, 1) - margin
    logging.info('min_z_value: {}'.format(min_z_value))
    last_row = np.array([0., 0., 1., -min_z_value]).reshape((1, 4))          # :103
    last_rows = {}
    reparam_depth_range = {im.name: [] for im in images.values()}
    common_reparam_depth_range = []
    for p in points3D.values():                                              # :113-151
        x = p.xyz.reshape((3, 1))
        x1 = np.vstack((x, np.array([[1., ]])))
        depth = 0
        for img_id in p.image_ids:
            img_name = images[img_id].name
            P_3by4 = np.dot(Kmat[img_id], np.hstack((rot[img_id], tvec[img_id])))
            depth_min = depth_range[img_name][0]
            P_4by4 = np.vstack((P_3by4, depth_min * last_row))
            if img_name not in last_rows:
                last_rows[img_name] = depth_min * last_row
            tmp = np.dot(P_4by4, x1)
            depth = tmp[3, 0] / tmp[2, 0]                                    # the fourth component, not its inverse
            if depth > 0:
                reparam_depth_range[img_name].append(depth)
        if depth > 0:                                                        # (sic) the LAST image's value, :150-151
            common_reparam_depth_range.append(depth)
    reparam_depth_range = robust_depth_range(reparam_depth_range)

    with open(os.path.join(save_dir, 'raw_depth.txt'), 'w') as fp:
        fp.write('# format: img_name, depth_min, depth_max\n')
        for img_name in sorted(depth_range.keys()):
            fp.write('{} {} {}\n'.format(img_name, depth_range[img_name][0], depth_range[img_name][1]))
    with open(os.path.join(save_dir, 'reparam_depth.txt'), 'w') as fp:
        fp.write('# format: img_name, depth_min, depth_max\n')
        for img_name in sorted(reparam_depth_range.keys()):
            fp.write('{} {} {}\n'.format(img_name, reparam_depth_range[img_name][0], reparam_depth_range[img_name][1]))
    with open(os.path.join(save_dir, 'last_rows.txt'), 'w') as fp:
        for img_name in sorted(last_rows.keys()):
            vec = last_rows[img_name]
            fp.write('{} {} {} {} {}\n'.format(img_name, vec[0, 0], vec[0, 1], vec[0, 2], vec[0, 3]))
    with open(os.path.join(save_dir, 'reference_plane.txt'), 'w') as fp:
        fp.write('{} {} {} {}\n'.format(last_row[0, 0], last_row[0, 1], last_row[0, 2], last_row[0, 3]))
    common = sorted(common_reparam_depth_range)
    cnt = len(common)
    min_depth = common[int(0.02 * cnt)] - 10
    max_depth = common[int(0.98 * cnt)] + 100.
    logging.info('{} points, depth_min: {}, depth_max: {}'.format(cnt, min_depth, max_depth))
    with open(os.path.join(save_dir, 'depth_ranges.txt'), 'w') as fp:
        for img_name in sorted(last_rows.keys()):
            fp.write('{} {} {}\n'.format(img_name, min_depth, max_depth))

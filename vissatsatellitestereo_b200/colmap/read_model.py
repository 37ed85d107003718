"""COLMAP sparse model in its TEXT form (cameras.txt / images.txt / points3D.txt), as much of the reference's
colmap/read_model.py:83-107,138-167,204-229,261-270 as reparam_depth needs: `read_model(path, ext='.txt')` ->
(cameras, images, points3D) dicts keyed by id, with the same attribute names.

File formats (COLMAP documentation): lines starting with '#' are comments;
  cameras.txt    CAMERA_ID MODEL WIDTH HEIGHT PARAMS[]
  images.txt     IMAGE_ID QW QX QY QZ TX TY TZ CAMERA_ID NAME            (then one line: X Y POINT3D_ID triples)
  points3D.txt   POINT3D_ID X Y Z R G B ERROR (IMAGE_ID POINT2D_IDX)*
"""
import collections
import os

import numpy as np

Camera = collections.namedtuple('Camera', ['id', 'model', 'width', 'height', 'params'])
Image = collections.namedtuple('Image', ['id', 'qvec', 'tvec', 'camera_id', 'name', 'xys', 'point3D_ids'])
Point3D = collections.namedtuple('Point3D', ['id', 'xyz', 'rgb', 'error', 'image_ids', 'point2D_idxs'])


def _data_lines(path):
    with open(path, 'r') as fid:
        for line in fid:
            line = line.strip()
            if line and not line.startswith('#'):
                yield line


def read_cameras_text(path):
    cameras = {}
    for line in _data_lines(path):
        t = line.split()
        cid = int(t[0])
        cameras[cid] = Camera(id=cid, model=t[1], width=int(t[2]), height=int(t[3]),
                              params=np.array([float(v) for v in t[4:]]))
    return cameras


def read_images_text(path):
    images = {}
    with open(path, 'r') as fid:
        lines = [ln.rstrip('\n') for ln in fid]
    i = 0
    while i < len(lines):
        line = lines[i].strip()
        i += 1
        if not line or line.startswith('#'):
            continue
        t = line.split()
        iid = int(t[0])
        pts = lines[i].split() if i < len(lines) else []      # the observation line may be empty
        i += 1
        xys = np.column_stack([np.array(pts[0::3], dtype=np.float64), np.array(pts[1::3], dtype=np.float64)]) \
            if pts else np.zeros((0, 2))
        images[iid] = Image(id=iid, qvec=np.array([float(v) for v in t[1:5]]), tvec=np.array([float(v) for v in t[5:8]]),
                            camera_id=int(t[8]), name=t[9], xys=xys,
                            point3D_ids=np.array(pts[2::3], dtype=np.int64))
    return images


def read_points3D_text(path):
    points3D = {}
    for line in _data_lines(path):
        t = line.split()
        pid = int(t[0])
        points3D[pid] = Point3D(id=pid, xyz=np.array([float(v) for v in t[1:4]]), rgb=np.array([int(v) for v in t[4:7]]),
                                error=float(t[7]), image_ids=np.array(t[8::2], dtype=np.int64),
                                point2D_idxs=np.array(t[9::2], dtype=np.int64))
    return points3D


def read_model(path, ext='.txt'):
    if ext != '.txt':
        raise NotImplementedError('only the text form of the sparse model is read here (the reference calls '
                                  "read_model(sparse_dir, ext='.txt'), reparam_depth.py:72)")
    return (read_cameras_text(os.path.join(path, 'cameras' + ext)), read_images_text(os.path.join(path, 'images' + ext)),
            read_points3D_text(os.path.join(path, 'points3D' + ext)))

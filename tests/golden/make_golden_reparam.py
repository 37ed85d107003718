#!/usr/bin/env python
"""Golden files for reparam_depth: a small synthetic COLMAP sparse model (text form) and what the REFERENCE's
reparam_depth.reparam_depth (/root/reference/reparam_depth.py:69-195, with the reference's own colmap/read_model.py)
writes for it -> tests/golden/reparam/.  pyquaternion is absent from the image; its Quaternion(w, x, y, z).rotation_matrix
is stood in for by the closed-form rotation matrix of the normalised quaternion (the only third-party piece).

Usage: python tests/golden/make_golden_reparam.py        (needs /root/reference; run in the build container)"""
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, 'reparam')


def rotation_from_quaternion(w, x, y, z):
    q = np.array([w, x, y, z], dtype=np.float64)
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def write_model(sparse_dir, camera_model, seed=3):
    """Five satellite-like cameras looking down on a 600 m scene, 400 points each seen by 2..5 of them."""
    rng = np.random.default_rng(seed)
    os.makedirs(sparse_dir, exist_ok=True)
    n_img, n_pts = 5, 400
    with open(os.path.join(sparse_dir, 'cameras.txt'), 'w') as fp:
        fp.write('# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n')
        for i in range(n_img):
            f = 2.0e6 + 1e4 * i
            params = [f, f * 1.01, 1024.0 + i, 1000.0 - i] + ([0.3 * i] if camera_model == 'perspective' else [])
            fp.write('{} {} 2048 2048 {}\n'.format(i + 1, 'PERSPECTIVE' if camera_model == 'perspective' else 'PINHOLE',
                                                  ' '.join(repr(float(p)) for p in params)))
    quats, tvecs = [], []
    for i in range(n_img):
        ang = np.radians(rng.uniform(150, 210))           # looking roughly down (-z)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        axis = 0.15 * axis + np.array([1.0, 0.0, 0.0])
        axis /= np.linalg.norm(axis)
        q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])
        R = rotation_from_quaternion(*q)
        C = np.array([rng.uniform(-5e4, 5e4), rng.uniform(-5e4, 5e4), 6.0e5 + rng.uniform(-1e4, 1e4)])
        quats.append(q)
        tvecs.append(-R @ C)
    tracks = []
    for j in range(n_pts):
        k = int(rng.integers(2, n_img + 1))
        tracks.append(sorted(rng.choice(n_img, size=k, replace=False) + 1))
    with open(os.path.join(sparse_dir, 'images.txt'), 'w') as fp:
        fp.write('# Image list with two lines of data per image:\n#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n'
                 '#   POINTS2D[] as (X, Y, POINT3D_ID)\n')
        for i in range(n_img):
            fp.write('{} {} {} {} 000{}.png\n'.format(i + 1, ' '.join(repr(float(v)) for v in quats[i]),
                                                     ' '.join(repr(float(v)) for v in tvecs[i]), i + 1, i))
            obs = ['{} {} {}'.format(repr(float(rng.uniform(0, 2048))), repr(float(rng.uniform(0, 2048))), j + 1)
                   for j in range(n_pts) if (i + 1) in tracks[j]]
            fp.write(' '.join(obs) + '\n')
    with open(os.path.join(sparse_dir, 'points3D.txt'), 'w') as fp:
        fp.write('# 3D point list with one line of data per point:\n#   POINT3D_ID, X, Y, Z, R, G, B, ERROR, TRACK[] as (IMAGE_ID, POINT2D_IDX)\n')
        for j in range(n_pts):
            xyz = [rng.uniform(-300, 300), rng.uniform(-300, 300), rng.uniform(-20, 90)]
            fp.write('{} {} 128 128 128 0.5 {}\n'.format(j + 1, ' '.join(repr(float(v)) for v in xyz),
                                                        ' '.join('{} {}'.format(i, 0) for i in tracks[j])))


def main():
    sys.path.insert(0, '/root/reference')
    stub = types.ModuleType('pyquaternion')

    class Quaternion(object):
        def __init__(self, w, x, y, z):
            self.q = (w, x, y, z)

        @property
        def rotation_matrix(self):
            return rotation_from_quaternion(*self.q)
    stub.Quaternion = Quaternion
    sys.modules['pyquaternion'] = stub
    import reparam_depth as ref                       # the reference module
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    for model in ('perspective', 'pinhole'):
        sparse = os.path.join(OUT, model, 'sparse')
        write_model(sparse, model)
        save = os.path.join(OUT, model, 'want')
        os.makedirs(save)
        ref.reparam_depth(sparse, save, camera_model=model)
        print(model, sorted(os.listdir(save)))


if __name__ == '__main__':
    main()

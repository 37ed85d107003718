"""Drop-in for the reference's colmap/read_dense.py:read_array (:36-51): COLMAP dense array reader.
File format: ASCII 'width&height&channels&' followed by little-endian float32 in Fortran (W,H,C) order."""
import numpy as np


def read_array(path):
    with open(path, 'rb') as fid:
        head = fid.read(64)
        parts = head.split(b'&', 3)
        if len(parts) < 4:
            raise ValueError('not a COLMAP array file: {}'.format(path))
        width, height, channels = int(parts[0]), int(parts[1]), int(parts[2])
        offset = len(parts[0]) + len(parts[1]) + len(parts[2]) + 3
        fid.seek(offset)
        array = np.fromfile(fid, np.float32)
    array = array.reshape((width, height, channels), order='F')
    return np.transpose(array, (1, 0, 2)).squeeze()


def read_array_hw(path):
    """Same file, returned as a C-contiguous (H, W) float32 array ready for a pinned-memory upload
    (single-channel files only; the payload already is row-major H x W)."""
    with open(path, 'rb') as fid:
        head = fid.read(64)
        parts = head.split(b'&', 3)
        width, height, channels = int(parts[0]), int(parts[1]), int(parts[2])
        if channels != 1:
            raise ValueError('depth maps are single-channel')
        fid.seek(len(parts[0]) + len(parts[1]) + len(parts[2]) + 3)
        array = np.fromfile(fid, np.float32, count=width * height)
    return array.reshape((height, width))

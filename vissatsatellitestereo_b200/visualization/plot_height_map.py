"""Preview images for DSMs / height maps (reference: visualization/plot_height_map.py:39-57 and
visualization/save_image_only.py:41-108), without matplotlib / imageio.

Same call signature, same file set (<name>.jpg, <name>.mask.jpg, optionally <name>.cbar.jpg) and the same mapping from
height to colour: clip to the [1, 99] NaN-percentiles (or `force_range`), pin pixels [0, 0] / [0, 1] to the range ends
(plot_height_map.py:48-49), normalise to [0, 1] over that range, look the colour up in the reference's 197-entry table
as matplotlib's ListedColormap does (index = int(x * N), x == 1 -> N - 1), NaN / masked pixels black
(save_image_only.py:62-101).  What differs is the rasterisation path (the reference draws a matplotlib figure and
resizes the canvas with nearest-neighbour; here the array is colour-mapped directly) and the JPEG encoder (OpenCV
instead of imageio), so files are compared by PSNR, not bytes (tests/test_host_io.py).  Skipped if OpenCV is missing."""
import numpy as np

from ._colormap_height import COLORMAP_HEIGHT

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def height_to_rgb(height_map, force_range=None, maskout=None, pin_range_pixels=True):
    """(H, W) heights -> (uint8 (H, W, 3) RGB, bool (H, W) invalid mask, (min_val, max_val)); the colour mapping of
    plot_height_map.py:39-57 + save_image_only.py:62-70,101 as one array operation."""
    height_map = np.array(height_map, dtype=np.float64, copy=True)
    if force_range is None:
        if np.isnan(height_map).all():
            force_range = (0.0, 1.0)
        else:
            min_val, max_val = np.nanpercentile(height_map, [1, 99])
            force_range = (min_val, max_val)
    min_val, max_val = float(force_range[0]), float(force_range[1])
    height_map = np.clip(height_map, min_val, max_val)
    if pin_range_pixels and height_map.shape[0] > 0 and height_map.shape[1] > 1:
        height_map[0, 0] = min_val
        height_map[0, 1] = max_val
    nan_mask = np.isnan(height_map)
    # save_image_only.py:62-63: NaN -> nanmin for the drawing, black afterwards
    filled = np.where(nan_mask, min_val, height_map)
    span = max_val - min_val
    x = (filled - min_val) / span if span > 0 else np.zeros_like(filled)
    n = COLORMAP_HEIGHT.shape[0]
    idx = np.minimum((x * n).astype(np.int64), n - 1)          # matplotlib Colormap.__call__: int(x * N), x == 1 -> N - 1
    rgb = COLORMAP_HEIGHT[np.clip(idx, 0, n - 1)]
    if maskout is not None:
        nan_mask = np.logical_or(nan_mask, maskout)
    rgb = rgb.copy()
    rgb[nan_mask] = 0
    return rgb, nan_mask, (min_val, max_val)


def plot_height_map(height_map, out_file, maskout=None, save_cbar=False, force_range=None):
    if cv2 is None:
        return
    rgb, nan_mask, (lo, hi) = height_to_rgb(height_map, force_range=force_range, maskout=maskout)
    cv2.imwrite(out_file, np.ascontiguousarray(rgb[:, :, ::-1]))
    idx = out_file.rfind('.')
    cv2.imwrite(out_file[:idx] + '.mask.jpg', np.uint8((1.0 - nan_mask.astype(np.float32)) * 255.0))
    if save_cbar:
        n = COLORMAP_HEIGHT.shape[0]
        ramp = COLORMAP_HEIGHT[np.minimum((np.arange(512) * n) // 512, n - 1)]
        bar = np.ascontiguousarray(np.tile(ramp[None, :, ::-1], (28, 1, 1)))
        cv2.putText(bar, '{:.1f}'.format(lo), (2, 18), cv2.FONT_HERSHEY_SIMPLEX, 0.45, (255, 255, 255), 1)
        cv2.putText(bar, '{:.1f}'.format(hi), (440, 18), cv2.FONT_HERSHEY_SIMPLEX, 0.45, (255, 255, 255), 1)
        cv2.imwrite(out_file[:idx] + '.cbar.jpg', bar)

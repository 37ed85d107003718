"""Preview images for DSMs / height maps (reference: visualization/plot_height_map.py:39-57 and
visualization/save_image_only.py:41-108).  The reference renders through matplotlib with its own colour table;
previews are lossy JPEGs and not part of the parity contract, so this keeps the call
signature and the file set (<name>.jpg, <name>.mask.jpg) and renders with OpenCV: clip to the [1, 99] NaN
percentiles (or force_range), linear colour ramp, NaN pixels black.  Skipped silently if OpenCV is missing."""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def plot_height_map(height_map, out_file, maskout=None, save_cbar=False, force_range=None):
    if cv2 is None:
        return
    height_map = np.array(height_map, dtype=np.float32, copy=True)
    nan_mask = np.isnan(height_map)
    if nan_mask.all():
        lo, hi = 0.0, 1.0
    elif force_range is None:
        lo, hi = np.nanpercentile(height_map, [1, 99])
    else:
        lo, hi = force_range
    scale = 255.0 / (hi - lo) if hi > lo else 0.0
    gray = np.clip((np.nan_to_num(height_map, nan=lo) - lo) * scale, 0, 255).astype(np.uint8)
    im = cv2.applyColorMap(gray, cv2.COLORMAP_TURBO)
    if maskout is not None:
        nan_mask = np.logical_or(nan_mask, maskout)
    im[nan_mask] = 0
    cv2.imwrite(out_file, im)
    idx = out_file.rfind('.')
    cv2.imwrite(out_file[:idx] + '.mask.jpg', np.uint8((1.0 - nan_mask.astype(np.float32)) * 255.0))
    if save_cbar:
        bar = cv2.applyColorMap(np.tile(np.arange(256, dtype=np.uint8)[None, :], (24, 1)), cv2.COLORMAP_TURBO)
        cv2.putText(bar, '{:.1f}'.format(lo), (2, 16), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (255, 255, 255), 1)
        cv2.putText(bar, '{:.1f}'.format(hi), (200, 16), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (0, 0, 0), 1)
        cv2.imwrite(out_file[:idx] + '.cbar.jpg', bar)

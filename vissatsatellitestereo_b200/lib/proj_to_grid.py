"""Drop-in for the reference's lib/proj_to_grid.py (:41-81): same signature, returns float64 (ysize, xsize).
Scatter (per-cell nanmax, 64-bit order-preserving keys) and the 3x3 NaN-hole fill both run on the GPU."""
from .. import engine


# points: each row is (xx, yy, zz); xoff: ul_e; yoff: ul_n; xsize: width; ysize: height
def proj_to_grid(points, xoff, yoff, xresolution, yresolution, xsize, ysize, propagate=False):
    # `propagate` is accepted and ignored, as in the reference (the fill always runs)
    dsm = engine.proj_to_grid_device(points, xoff, yoff, xresolution, yresolution, int(xsize), int(ysize))
    return dsm.cpu().numpy()

"""ORACLE (test infrastructure, CPU numpy; parity unpinned at the cv2 boundary -- see below).

Restates the `use_sphere_mask` branch of the dataset, fmc/data/dataset.py:5350-5403: per object, the minimum enclosing
circle (centre, radius) of its segmentation mask becomes a Gaussian disc

    dist = sqrt((x - cx)^2 + (y - cy)^2);  sigma = radius / 2                       (:5366-5371)
    gaussian = exp(-0.5 * (dist / sigma)^2);  gaussian /= gaussian.max()            (:5374-5378)
    gaussian *= disc                                                                 (:5381)

Third-party pieces that are absent offline: `cv2.minEnclosingCircle` (:5359; here the circle is an INPUT) and
`cv2.circle(mask, (int(cx), int(cy)), int(radius), 1, -1)` (:5363; restated as the Euclidean disc
(x - int(cx))^2 + (y - int(cy))^2 <= int(radius)^2 -- OpenCV's midpoint rasterisation can differ from it by single
boundary pixels, which is why this piece says "unpinned").  dtypes follow numpy's promotion in the reference: centre and
radius are float32 (cv2), the pixel grids int64, so the arithmetic runs in float64."""
import numpy as np


def gaussian_sphere_mask(center, radius, H, W):
    """center = (cx, cy), radius: floats as cv2.minEnclosingCircle returns them -> float64 [H, W]."""
    cx, cy, radius = np.float32(center[0]), np.float32(center[1]), np.float32(radius)
    if not radius > 0:
        return np.zeros((H, W), dtype=np.float64)  # empty segmentation mask: the all-zero mask is kept (:5357-5358)
    y, x = np.ogrid[:H, :W]
    disc = (x - int(cx)) ** 2 + (y - int(cy)) ** 2 <= int(radius) ** 2
    dist_from_center = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
    sigma = radius / 2
    gaussian = np.exp(-0.5 * (dist_from_center / sigma) ** 2)
    gaussian = gaussian / gaussian.max()
    return disc * gaussian


def sphere_masks(circles, H, W):
    """circles [..., 3] = (cx, cy, r) -> float64 [..., H, W]."""
    circles = np.asarray(circles)
    flat = circles.reshape(-1, 3)
    out = np.stack([gaussian_sphere_mask((c[0], c[1]), c[2], H, W) for c in flat])
    return out.reshape(circles.shape[:-1] + (H, W))

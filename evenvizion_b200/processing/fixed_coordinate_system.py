"""Original <-> fixed coordinate system (reference evenvizion/processing/fixed_coordinate_system.py:19-122);
the per-point transform runs in libevz.so's remap kernel."""
from copy import deepcopy

import numpy as np
import torch


def _remap(coordinates, homography_dict, sx, sy, inverse):
    from .. import default_engine
    eng = default_engine()
    frames = list(coordinates.keys())
    S = np.array([np.asarray(homography_dict[f], np.float64).reshape(3, 3) for f in frames]).reshape(-1, 9)
    pts, fidx = [], []
    for i, f in enumerate(frames):
        for rect in coordinates[f]:
            pts.append((rect["x1"], rect["y1"]))
            fidx.append(i)
    out = {f: [] for f in frames}
    if not pts:
        return out
    res = eng.remap(torch.tensor(pts, dtype=torch.float64, device=eng.device),
                    torch.tensor(fidx, dtype=torch.int32, device=eng.device),
                    torch.from_numpy(S).to(eng.device), sx, sy, inverse).cpu().numpy()
    j = 0
    for f in frames:
        for rect in coordinates[f]:
            new_rect = deepcopy(rect)
            new_rect["x1"], new_rect["y1"] = np.float64(res[j, 0]), np.float64(res[j, 1])
            out[f].append(new_rect)
            j += 1
    return out


def from_original_to_fix(original_coordinates, homography_dict, original_image_shape, resize_image_shape):
    """reference fixed_coordinate_system.py:19-69: scale to the resized frame, apply the frame's
    superposition, round to 2 decimals.  homography_dict: {frame_no: 3x3} (superposition_dict output)."""
    original_h, original_w = original_image_shape
    resize_h, resize_w = resize_image_shape
    return _remap(original_coordinates, homography_dict, int(resize_w) / original_w, int(resize_h) / original_h, False)


def from_fix_to_original(fix_coordinates, homography_dict, original_image_shape, resize_image_shape):
    """reference fixed_coordinate_system.py:72-122 (scales first, then applies the inverse, as the reference does)."""
    original_h, original_w = original_image_shape
    resize_h, resize_w = resize_image_shape
    return _remap(fix_coordinates, homography_dict, original_w / resize_w, original_h / resize_h, True)

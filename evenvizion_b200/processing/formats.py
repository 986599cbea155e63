"""On-disk formats of the geometry path at scale: streaming writers and array readers
for `dict_with_homography_matrix.json` (written by the reference with `json.dump(result, f)`,
evenvizion/examples/evenvizion_component.py:139-140; read back by utils.py:148-181) and for the coordinate JSONs
(`original_coordinates.json` / the fixed-coordinate JSON, utils.py:233-255, fixed_coordinate_system.py:19-122).

The writers take the arrays the engine returns (`GeometryEngine.video_geometry`: H_fixed / status) and emit,
chunk by chunk, exactly the bytes `json.dump` produces for the reference's dict -- string keys in ascending frame
order, `repr` floats, `null` for a missing H, `resize_info` last -- without building 100 000 nested Python lists
first.  Host-side only: nothing here touches the GPU.
"""
import io
import json

import numpy as np

_CHUNK = 4096          # frames per write() call


def _row(r):
    return "[" + ", ".join(float.__repr__(float(v)) for v in r) + "]"


def _matrix(H):
    return "[" + ", ".join(_row(r) for r in H) + "]"


def dump_homography_dict(dst, H, resize_info, valid=None, first_frame=2):
    """Write {"2": {"H": [[...],[...],[...]]}, ..., "resize_info": {...}}.

    dst: path or text file object.  H: (P, 3, 3) float64 (frame first_frame + p holds H[p]); valid: optional
    (P,) bool, False -> {"H": null} (the reference's `none_H_processing=False` entries).  Byte-identical to
    json.dump of the dict `get_homography_dict` returns."""
    H = np.asarray(H, np.float64).reshape(-1, 3, 3)
    own = isinstance(dst, (str, bytes)) or hasattr(dst, "__fspath__")
    f = open(dst, "w") if own else dst
    try:
        f.write("{")
        for lo in range(0, len(H), _CHUNK):
            parts = []
            for p in range(lo, min(lo + _CHUNK, len(H))):
                body = "null" if valid is not None and not valid[p] else _matrix(H[p])
                parts.append('"%d": {"H": %s}' % (first_frame + p, body))
            f.write(", ".join(parts))
            f.write(", ")
        f.write('"resize_info": ' + json.dumps(resize_info) + "}")
    finally:
        if own:
            f.close()


def load_homography_arrays(src):
    """Inverse of dump_homography_dict: returns (frames int64 (P,), H float64 (P,3,3) with NaN where H is null,
    valid bool (P,), resize_info).  Raises the reference's ValueError when resize_info is missing (utils.py:173-177)."""
    if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__"):
        with open(src, "r") as f:
            d = json.load(f)
    else:
        d = json.load(src)
    if "resize_info" not in d:
        raise ValueError("Specify the height and width of the frame "
                         "for which the homography matrix was obtained")
    resize_info = d.pop("resize_info")
    frames = np.fromiter((int(k) for k in d), np.int64, len(d))
    H = np.full((len(d), 3, 3), np.nan)
    valid = np.zeros(len(d), bool)
    for i, v in enumerate(d.values()):
        if v["H"] is not None:
            H[i] = v["H"]
            valid[i] = True
    return frames, H, valid, resize_info


def dump_coordinates(dst, frames, counts, xy, extra=None):
    """Write {"1": [{"x1": x, "y1": y}, ...], ...} from flat arrays: frames (F,) frame numbers, counts (F,) points
    per frame, xy (sum counts, 2) float64 (already rounded by the remap kernel).  extra: optional list (one dict per
    point) of the other keys of each rectangle, which the reference preserves by deepcopy
    (fixed_coordinate_system.py:64).  Byte-identical to json.dump of the dict `from_original_to_fix` returns."""
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    own = isinstance(dst, (str, bytes)) or hasattr(dst, "__fspath__")
    f = open(dst, "w") if own else dst
    try:
        f.write("{")
        j = 0
        first = True
        buf = io.StringIO()
        for fr, n in zip(frames, counts):
            if not first:
                buf.write(", ")
            first = False
            pts = []
            for _ in range(int(n)):
                rect = dict(extra[j]) if extra is not None else {}
                rect["x1"], rect["y1"] = float(xy[j, 0]), float(xy[j, 1])
                pts.append(json.dumps(rect))
                j += 1
            buf.write('"%d": [%s]' % (int(fr), ", ".join(pts)))
            if buf.tell() > 1 << 20:
                f.write(buf.getvalue())
                buf = io.StringIO()
        f.write(buf.getvalue())
        f.write("}")
    finally:
        if own:
            f.close()


def load_coordinates_arrays(src):
    """Returns (frames int64 (F,), counts int64 (F,), xy float64 (sum counts, 2)) from a coordinate JSON."""
    if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__"):
        with open(src, "r") as f:
            d = json.load(f)
    else:
        d = json.load(src)
    frames = np.fromiter((int(k) for k in d), np.int64, len(d))
    counts = np.fromiter((len(v) for v in d.values()), np.int64, len(d))
    xy = np.array([(r["x1"], r["y1"]) for v in d.values() for r in v], np.float64).reshape(-1, 2)
    return frames, counts, xy

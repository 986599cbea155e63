"""Drop-in for the reference's example CLI (evenvizion/examples/evenvizion_component.py:39-146) on the GPU path.

Same flags, same output locations under ``<cwd>/<experiment_name>/<video stem>/``:

    dict_with_homography_matrix.json      json.dump of get_homography_dict's result        (component.py:139-140)
    metrics_file.txt                      "Maximum movement during the entire video: ..."  (component.py:62-66)
    recalculated_coordinates.json         the fixed-coordinate JSON = from_original_to_fix's return value
                                          (README.md:49; the reference computes it at component.py:92-97 and only
                                          hands it to the visualiser)

The drawing layer (heat-map PNGs, match lines, the original-vs-fixed overlay) is out of scope: the flags that ask for
it are accepted, the numbers behind it are written, no picture is.  Quirks kept on purpose: ``--resize_width`` is parsed
and not passed on (component.py:135-136), and every flag value given on the command line is a non-empty string, i.e.
true (component.py:108-115 declare no type).
"""
import argparse
import json
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from evenvizion_b200.processing.fixed_coordinate_system import from_original_to_fix        # noqa: E402
from evenvizion_b200.processing.utils import (read_homography_dict, read_json_with_coordinates,   # noqa: E402
                                              superposition_dict)
from evenvizion_b200.processing.video_processing import get_homography_dict                 # noqa: E402

log = logging.getLogger("evenvizion_component")


def max_movement(path_to_homography_dict):
    """The number heatmap_video_processing returns (processing_visualization.py:404-418): the largest displacement of a
    pixel of the resized frame over frames 1 .. F-1, on the device (evz_max_movement)."""
    import torch
    from evenvizion_b200 import default_engine
    homography_matrices, resize_info = read_homography_dict(path_to_homography_dict)
    sup = superposition_dict(homography_matrices)
    keys = sorted(sup)
    S = np.array([np.asarray(sup[k], np.float64).reshape(9) for k in keys])
    eng = default_engine()
    S_dev = torch.from_numpy(S).to(eng.device)
    return float(eng.max_movement(S_dev, len(keys) - 1, resize_info["h"], resize_info["w"]).item())


def run(cap, original_shape, save_folder, path_to_original_coordinate=None, none_H_processing=True,
        heatmap_visualization=True, show_matching_visualization=True, **kw):
    """Body of the reference's __main__ for an opened capture; returns the paths it wrote."""
    os.makedirs(save_folder, exist_ok=True)
    if show_matching_visualization:
        log.info("matching visualisation is a drawing step (out of scope): no PNGs are written")
    result = get_homography_dict(cap, matching_path=None, none_H_processing=none_H_processing, **kw)
    out = {"homography": os.path.join(save_folder, "dict_with_homography_matrix.json")}
    with open(out["homography"], "w") as json_:
        json.dump(result, json_)
    if heatmap_visualization:
        mm = max_movement(out["homography"])
        out["metrics"] = os.path.join(save_folder, "metrics_file.txt")
        with open(out["metrics"], "w") as txt_:
            txt_.write("Maximum movement during the entire video: {}".format(mm))
            if not np.isfinite(mm):
                txt_.write("There are some frames with undefined coordinates")
    if path_to_original_coordinate:
        homography_matrices, resize_info = read_homography_dict(out["homography"])
        sup = superposition_dict(homography_matrices)
        original_coordinates = read_json_with_coordinates(path_to_original_coordinate)
        recalculated = from_original_to_fix(original_coordinates, sup, original_shape,
                                            [resize_info["h"], resize_info["w"]])
        out["fixed"] = os.path.join(save_folder, "recalculated_coordinates.json")
        with open(out["fixed"], "w") as json_:
            json.dump(recalculated, json_)
    return out


if __name__ == "__main__":
    import cv2
    parser = argparse.ArgumentParser(description="custom arguments")
    parser.add_argument("--path_to_video", type=str, default="test_video/test_video.mp4")
    parser.add_argument("--experiment_name", type=str, default="test_video_processing")
    parser.add_argument("--resize_width", type=int, help="wights to resize image", default=400)
    parser.add_argument("--path_to_original_coordinate", help="path to json with original coordinate",
                        default="test_video/original_coordinates.json")
    parser.add_argument("--none_H_processing", help="If True we use H_prev as H, False- do nothing", default=True)
    parser.add_argument("--heatmap_visualization", help="Getting heatmap visualization", default=True)
    parser.add_argument("--show_matching_visualization", help="Getting matching visualization", default=True)
    args = parser.parse_args()
    logging.basicConfig(level=logging.INFO)
    save_folder = os.path.join(os.getcwd(), args.experiment_name, os.path.split(args.path_to_video)[-1].split(".")[0])
    cap = cv2.VideoCapture(args.path_to_video)
    original_shape = [int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))]
    written = run(cap, original_shape, save_folder, args.path_to_original_coordinate, args.none_H_processing,
                  args.heatmap_visualization, args.show_matching_visualization)
    for k, v in written.items():
        print(k, v)

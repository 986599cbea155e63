"""Frame feature front-end (reference evenvizion/processing/frame_processing.py:20-108).

Keypoint detection / description stays OpenCV on the CPU (out of scope of the CUDA path, timed
separately); `concatenate_all_features_types` sends the descriptors through the GPU matcher.
SURF (non-free, non-integer float descriptors) cannot use the int8 tensor-core matcher and is
rejected; the default feature list is therefore ["SIFT", "ORB"] instead of the reference's
["SURF", "SIFT", "ORB"].
"""
import numpy as np

from .matching import KeyPoints, NoMatchesException
from .utils import remove_double_matching

DEFAULT_FEATURES = ["SIFT", "ORB"]


def resize(image, width=None, height=None):
    """imutils.resize as the reference uses it (video_processing.py:62): INTER_AREA, aspect kept."""
    import cv2
    h, w = image.shape[:2]
    if width is None and height is None:
        return image
    if width is None:
        r = height / float(h)
        dim = (int(w * r), height)
    else:
        r = width / float(w)
        dim = (width, int(h * r))
    return cv2.resize(image, dim, interpolation=cv2.INTER_AREA)


class FrameProcessing:
    """reference frame_processing.py:20-108."""

    def __init__(self, frame, features_type_list=None):
        self.frame = frame
        self.features_types = features_type_list or list(DEFAULT_FEATURES)

    def detect_and_describe_features(self, features_name):
        """OpenCV keypoints + descriptors (reference frame_processing.py:42-71).  CPU, out of scope."""
        import cv2
        if features_name == "ORB":
            det = cv2.ORB_create()
        elif features_name == "SIFT":
            det = cv2.SIFT_create() if hasattr(cv2, "SIFT_create") else cv2.xfeatures2d.SIFT_create()
        elif features_name == "SURF":
            raise ValueError("SURF descriptors are non-integer floats: the exact int8 matcher cannot take them "
                             "(and SURF is non-free); use SIFT and/or ORB")
        else:
            raise ValueError("You need to choose descriptors type")
        kps, descriptors = det.detectAndCompute(self.frame, None)
        coordinates = np.float32([kp.pt for kp in kps]).reshape(-1, 2)
        return coordinates, descriptors

    def concatenate_all_features_types(self, acceding_image):
        """Static matches of every feature type, concatenated and de-duplicated
        (reference frame_processing.py:73-108).  self = new frame, acceding_image = previous frame."""
        all_a, all_b = [], []
        for feature_type in self.features_types:
            coords_a, desc_a = self.detect_and_describe_features(feature_type)
            coords_b, desc_b = acceding_image.detect_and_describe_features(feature_type)
            static_a, static_b = KeyPoints(coords_a, desc_a).match_static_kps(KeyPoints(coords_b, desc_b))
            all_a.extend(static_a)
            all_b.extend(static_b)
        return remove_double_matching(all_a, all_b)

"""Constants of the geometry path; values are part of the contract
(reference evenvizion/processing/constants.py:15-32)."""

#: coordinates at or above this value are considered undefined
INFINITY_COORDINATE = 10000
#: minimum inlier fraction of the second RANSAC for H to be accepted
LENGTH_ACCOUNTED_POINTS = 0.7
#: reprojection threshold (pixels) of both RANSAC levels
THRESHOLD_FOR_FIND_HOMOGRAPHY = 3.0
#: Lowe ratio
LOWES_RATIO = 0.5
#: minimum number of matches after the many-to-one filter
MINIMUM_MATCHING_POINTS = 4
#: heat-map normalisation constant (visualisation only; kept for import compatibility)
HEATMAP_CONSTANT = 1000

# ---- knobs the CUDA path adds (the reference's cv2.findHomography draws its own samples)
#: RANSAC hypotheses evaluated per findHomography call
RANSAC_HYPOTHESES = 1024
#: seed of the counter-based hypothesis sampler
RANSAC_SEED = 0

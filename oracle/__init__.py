"""CPU oracle for the EvenVizion frame-to-frame geometry hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in NumPy (plus live OpenCV calls where the reference itself calls
OpenCV), the algorithm of the reference's hot path
(`evenvizion/processing/{matching,utils,video_processing,fixed_coordinate_system}.py`)
so that the CUDA implementation in `evenvizion_b200/` can be checked against it.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it.  The product package never does:
`evenvizion_b200` fails loudly when its CUDA library is missing.

Parity pinning (see DESIGN.md "Oracle"):
  * the reference has no tests, so there are no upstream golden vectors for
    match / RANSAC.  The oracle is pinned instead against the reference run in
    the build container (`tests/golden/make_golden.py` imports
    `/root/reference` and OpenCV 4.13 and commits the vectors), and against
    the one known-answer artefact the reference ships
    (`metrics_file.txt` = 863.0428982580879, scan + remap).
  * RANSAC hypotheses cannot be injected into `cv2.findHomography`, so the
    seeded hypothesis generator / 4-point solver here is the *definition* the
    CUDA kernels must reproduce bit-for-bit; only the scoring formula and the
    refit are pinned against OpenCV.
"""

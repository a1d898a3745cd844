"""CPU oracle for the video_unscreen per-pixel matte hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under
``video_unscreen_b200/`` imports ``oracle``; the product path raises when the
CUDA library is missing instead of falling back to anything in here.

What it is
----------
A numpy restatement of the reference's (AnyiRao/video_unscreen) arithmetic on
the hot path named in SURVEY.md section 8.  The reference itself is pure Python
on top of un-vendored, un-pinned third-party libraries (OpenCV, NumPy, torch
CPU, scikit-learn), so the restatement has two layers:

* ``oracle.cvmodel``   closed-form integer / IEEE models of the OpenCV
  primitives the reference calls (cvtColor BGR2HSV / BGR2GRAY / HSV2BGR,
  resize LINEAR / NEAREST, dilate / erode with MORPH_ELLIPSE, inRange).
  No cv2 import; verified bit-exact against cv2 4.13.0 in ``tests/``.
* ``oracle.refport``   the reference's functions (file:line cited on each)
  restated on top of ``cvmodel``.

Parity pinning
--------------
The reference ships no tests, golden vectors or fixtures.
The oracle is therefore pinned against outputs of the reference itself,
executed in the build container under a non-invasive import shim
(``tests/golden/make_golden.py``; library versions recorded in
``tests/golden/MANIFEST.json``) and committed as fixtures in ``tests/golden``.
``tests/test_oracle_vs_reference.py`` additionally re-runs the live reference
whenever ``/root/reference`` is present.  The temporal median (SURVEY.md
section 8 row a23) has NO reference implementation; its oracle is the survey's
specification ``np.median(stack, 0).astype(np.uint8)`` and that row is
"parity unpinned" by construction.
"""

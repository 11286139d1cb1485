"""CPU oracle for the Pix2Pose per-detection inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pix2pose_b200/`` (the product) may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs do, and only as the checker / the timed CPU arm.

Parity status: the reference has no tests, golden vectors or fixtures
and cannot be imported here (needs keras 2.2.1 / tensorflow-gpu 1.9 / scikit-image);
*parity unpinned* for the network and resize halves -- the restatement below is pinned
only by algebraic self-checks (tests/test_oracle_*.py).  The PnP half calls the real
``cv2.solvePnPRansac`` (container OpenCV 4.13.0; reference pins 3.4.2.17).
"""

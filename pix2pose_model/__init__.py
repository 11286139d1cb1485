"""Import-compatible alias of the reference package name: ``from pix2pose_model import recognition as recog``
(tools/5_evaluation_bop_basic.py:30) resolves to the B200 implementation in ``pix2pose_b200``."""

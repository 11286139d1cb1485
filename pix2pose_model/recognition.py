"""``pix2pose_model.recognition`` -> B200 drop-in (see pix2pose_b200/recognition.py)."""
from pix2pose_b200.recognition import pix2pose, PoseBatchResult  # noqa: F401

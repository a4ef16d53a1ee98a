"""Import shim for the reference's ``from label_smoothing import *`` (train.py:21) -- see compat/mtn.py."""
from mtn_b200.label_smoothing import LabelSmoothing  # noqa: F401

"""Alias of the reference's `ext` package (ext/__init__.py:18-23): `from nksr_b200 import ext` then
`ext.sdfgen.sdf_from_points(...)` as at dataset/av_gt_geometry.py:63-78 and models/loss.py:85."""
from . import sdfgen  # noqa: F401

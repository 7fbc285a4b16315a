"""Mirrors of the reference's ``utility`` package for the anchor hot path."""
from . import anchor_manipulator, bbox_util, custom_op  # noqa: F401

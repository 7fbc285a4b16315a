"""dan_b200 -- Blackwell-native (sm_100a) anchor hot path of HiKapok/DAN.

    from dan_b200.utility import anchor_manipulator, bbox_util

mirrors ``utility/anchor_manipulator.py`` and ``utility/bbox_util.py`` of the
reference on CUDA tensors; everything runs in hand-written CUDA kernels reached
through the C ABI of ``include/dan_b200.h`` (``dan_b200/libdan_b200.so``)."""
from . import _lib
from . import functional
from .utility import anchor_manipulator, bbox_util, custom_op
from .utility.anchor_manipulator import AnchorCreator, AnchorEncoder

__all__ = ["functional", "anchor_manipulator", "bbox_util", "custom_op", "AnchorCreator", "AnchorEncoder"]
__version__ = "0.1.0"

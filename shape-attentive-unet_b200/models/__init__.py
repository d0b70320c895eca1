from .models import ModelBuilder, SegmentationModule, SAUNet, DecoderBlock  # noqa: F401
from .attention_blocks import DualAttBlock  # noqa: F401
from .GSConv import GatedSpatialConv2d  # noqa: F401
from .resnet import BasicBlock  # noqa: F401
from .norm import Norm2d  # noqa: F401

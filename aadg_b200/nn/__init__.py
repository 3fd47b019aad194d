"""Segmentation network host side: smp-style constructors over the CUDA layers of libaadg_b200.so."""
from .network import DeepLabV3Plus, Unet, SegNet  # noqa: F401

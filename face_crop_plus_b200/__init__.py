"""face_crop_plus_b200 — B200-native hot path of face-crop-plus behind the reference's own Python API.

``from face_crop_plus_b200 import Cropper`` mirrors ``from face_crop_plus import Cropper`` (__init__.py:1).
"""
from .cropper import Cropper  # noqa: F401

__all__ = ["Cropper"]

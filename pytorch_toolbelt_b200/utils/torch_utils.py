"""Host-side marshalling helpers the tiled-inference loop uses on either side of the hot path.

Mirrors the subset of reference ``pytorch_toolbelt/utils/torch_utils.py`` that the README loop and the
reference tests touch: ``to_numpy`` (:144), ``image_to_tensor`` (:204), ``tensor_from_rgb_image`` (:234),
``rgb_image_from_tensor`` (:244). Pure host glue: no kernels.
"""
from collections.abc import Iterable
from typing import Any, Union

import numpy as np
import torch

from .support import pytorch_toolbelt_deprecated

__all__ = ["to_numpy", "image_to_tensor", "tensor_from_rgb_image", "tensor_from_mask_image", "rgb_image_from_tensor"]


def to_numpy(x: Union[torch.Tensor, np.ndarray, Any, None]) -> Union[np.ndarray, None]:
    if x is None:
        return None
    if torch.is_tensor(x):
        x = x.detach()
        if x.dtype == torch.bfloat16:
            x = x.float()
        return x.cpu().numpy()
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, (Iterable, int, float)):
        return np.array(x)
    raise ValueError("Unsupported type")


def image_to_tensor(image: np.ndarray, dummy_channels_dim=True) -> torch.Tensor:
    """[H,W,C] -> contiguous [C,H,W] tensor; [H,W] -> [1,H,W] (or [H,W] with dummy_channels_dim=False)."""
    if image.ndim not in (2, 3):
        raise ValueError(f"Image must have shape [H,W] or [H,W,C]. Got image with shape {image.shape}")
    if image.ndim == 2:
        chw = image[None] if dummy_channels_dim else image
    else:
        chw = image.transpose(2, 0, 1)
    return torch.from_numpy(np.ascontiguousarray(chw))


@pytorch_toolbelt_deprecated("This function is deprecated, please use image_to_tensor instead")
def tensor_from_rgb_image(image: np.ndarray) -> torch.Tensor:
    return image_to_tensor(image)


@pytorch_toolbelt_deprecated("This function is deprecated, please use image_to_tensor instead")
def tensor_from_mask_image(mask: np.ndarray) -> torch.Tensor:
    return image_to_tensor(mask)


def rgb_image_from_tensor(
    image: torch.Tensor,
    mean=(0.485, 0.456, 0.406),
    std=(0.229, 0.224, 0.225),
    min_pixel_value=0.0,
    max_pixel_value=255.0,
    dtype=np.uint8,
) -> np.ndarray:
    """[C,H,W] normalised tensor -> [H,W,C] image: clip(max_pixel_value * (x * std + mean))."""
    hwc = np.moveaxis(to_numpy(image), 0, -1)
    restored = max_pixel_value * (hwc * to_numpy(std) + to_numpy(mean))
    return np.clip(restored, a_min=min_pixel_value, a_max=max_pixel_value).astype(dtype)

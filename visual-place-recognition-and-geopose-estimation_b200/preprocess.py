"""GPU image preprocessing: the reference's ``Resize((224, 224)) -> ToTensor() -> Normalize(mean, std)``
(dinov2salad/dinov2salad_finetuning.py:45-50) applied to decoded uint8 RGB images on the device instead of on the
data loader's main thread (``num_workers=0``, :89).  Bit-identical to torchvision + Pillow (vpr_preprocess_image).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class GpuPreprocess:
    """``pre(images) -> [B, 3, h, w]``; ``images`` is a sequence of ``[H, W, 3]`` uint8 tensors (any sizes; CUDA, or
    host tensors that are copied over first -- pin them to overlap the copy)."""

    def __init__(self, size: Sequence[int] = (224, 224), mean: Sequence[float] = IMAGENET_MEAN,
                 std: Sequence[float] = IMAGENET_STD, out_dtype: torch.dtype = torch.float32, device="cuda"):
        self.out_h, self.out_w = int(size[0]), int(size[1])
        self._mean = (C.c_float * 3)(*[float(v) for v in mean])
        self._std = (C.c_float * 3)(*[float(v) for v in std])
        self.out_dtype = out_dtype
        self.device = torch.device(device)
        self._ws = None

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def __call__(self, images, return_uint8: bool = False):
        if isinstance(images, torch.Tensor) and images.dim() == 3:
            images = [images]
        if self.device.type != "cuda":
            raise RuntimeError("GpuPreprocess needs a CUDA device: the B200 path has no CPU fallback")
        lib = L.lib()
        B = len(images)
        out = torch.empty(B, 3, self.out_h, self.out_w, dtype=self.out_dtype, device=self.device)
        u8 = torch.empty(B, self.out_h, self.out_w, 3, dtype=torch.uint8, device=self.device) if return_uint8 else None
        st = L.stream_ptr(self.device)
        stacked = images if isinstance(images, torch.Tensor) and images.dim() == 4 else None
        if stacked is None and B > 1 and all(isinstance(im, torch.Tensor) and im.shape == images[0].shape for im in images):
            for im in images:
                if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                    raise ValueError("each image must be a [H, W, 3] uint8 tensor (RGB)")
            stacked = torch.stack([im.to(self.device, non_blocking=True) for im in images])
        if stacked is not None:                  # one size: the whole batch in the same four launches
            if stacked.dtype != torch.uint8 or stacked.shape[3] != 3:
                raise ValueError("a stacked batch must be [B, H, W, 3] uint8 (RGB)")
            stacked = stacked.to(self.device, non_blocking=True).contiguous()
            H, W = int(stacked.shape[1]), int(stacked.shape[2])
            nbytes = lib.vpr_preprocess_batch_workspace_bytes(B, H, W, self.out_h, self.out_w)
            ws = self._workspace(nbytes)
            with L.on_device(stacked):
                L.check(lib.vpr_preprocess_batch(L.ptr(stacked), B, H, W, self.out_h, self.out_w, self._mean, self._std, L.ptr(out),
                                                 L.dtype_code(self.out_dtype), L.ptr(u8), L.ptr(ws), ws.numel(), st),
                        "vpr_preprocess_batch")
            return (out, u8) if return_uint8 else out
        for i, img in enumerate(images):
            if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
                raise ValueError("each image must be a [H, W, 3] uint8 tensor (RGB)")
            img = img.to(self.device, non_blocking=True).contiguous()
            H, W = int(img.shape[0]), int(img.shape[1])
            nbytes = lib.vpr_preprocess_workspace_bytes(H, W, self.out_h, self.out_w)
            ws = self._workspace(nbytes)
            with L.on_device(img):
                L.check(lib.vpr_preprocess_image(L.ptr(img), H, W, self.out_h, self.out_w, self._mean, self._std, L.ptr(out[i]),
                                                 L.dtype_code(self.out_dtype), L.ptr(u8[i]) if u8 is not None else None,
                                                 L.ptr(ws), ws.numel(), st), "vpr_preprocess_image")
        return (out, u8) if return_uint8 else out

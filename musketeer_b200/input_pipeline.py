"""Input hand-off between the data loader and the model (trainer.py:1246-1284 `_prepare_sample`: move_to_cuda
followed by the bf16 cast of every floating tensor; the images were normalised on the host by the dataset transforms,
data/mm_data/*_dataset.py: ToTensor + Normalize).

B200 side of the same contract:
  * `pin(sample)`: page-locked copies of every tensor of a collated sample, so the copy engine can run asynchronously;
  * `DevicePrefetcher`: the host -> device copy of batch i+1 runs on a copy stream while batch i computes; the consumer only
    waits on an event (the role `GraphedTrainStep.prefetch` plays for the graph's static inputs);
  * images may arrive as decoded uint8 HWC pixels under `net_input["patch_images_u8"]` ([B, H, W, 3]): 1 byte per value over
    PCIe instead of 4, normalised on the device by csrc/pool.cu normalize_u8_kernel with the host transforms' exact fp32
    arithmetic, written as the [B, 3, H, W] `patch_images` tensor the model takes.
Nothing here touches the dataset classes: the uint8 key is what a maintainer's collater adds when it skips the two transforms."""
import ctypes as C

import torch

from . import ops
from ._lib import call
from .synthetic import map_tensors

IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def normalize_images(u8, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), dtype=torch.bfloat16):
    """uint8 [B, H, W, 3] (device) -> (x / 255 - mean) / std as [B, 3, H, W] `dtype` (fp32 or bf16)."""
    ops._need_cuda(u8)
    assert u8.dtype == torch.uint8 and u8.dim() == 4 and u8.shape[-1] == 3 and u8.is_contiguous()
    B, H, W, _ = u8.shape
    y = torch.empty(B, 3, H, W, dtype=dtype, device=u8.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    call("ofa_normalize_u8", ops._p(u8), ops._p(y), B, H, W, C.cast(m, C.c_void_p), C.cast(s, C.c_void_p), ops._dt(y), ops._st(),
         work=("byte", u8.numel() * (1.0 + y.element_size())))
    return y


def pin(sample):
    """Page-locked copy of every tensor in a (nested) sample."""
    return map_tensors(sample, lambda t: t.pin_memory() if not t.is_cuda and not t.is_pinned() else t)


def prepare_sample(sample, device, float_dtype=torch.bfloat16, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """trainer.py:1246-1284 on the current stream: copy to the device (asynchronous for pinned tensors), cast floating tensors
    to `float_dtype`, and turn `patch_images_u8` into the normalised `patch_images`."""
    def move(t):
        t = t.to(device, non_blocking=True)
        return t.to(float_dtype) if t.is_floating_point() else t
    out = map_tensors(sample, move)
    for s in (out if isinstance(out, (list, tuple)) else [out]):
        ni = s.get("net_input", s) if isinstance(s, dict) else None
        if ni is not None and "patch_images_u8" in ni:
            ni["patch_images"] = normalize_images(ni.pop("patch_images_u8"), mean, std, float_dtype)
    return out


class DevicePrefetcher:
    """Iterates over host samples; the copy + normalisation of the next sample is issued on a side stream while the caller
    works on the current one.  Yields device samples whose producing work the current stream already waits for."""

    def __init__(self, iterable, device, float_dtype=torch.bfloat16, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
        self.it, self.device, self.dtype, self.mean, self.std = iter(iterable), device, float_dtype, mean, std
        self.stream = torch.cuda.Stream(device=device)
        self._next = None
        self._fill()

    def _fill(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            dev = prepare_sample(host, self.device, self.dtype, self.mean, self.std)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (dev, ev, host)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev, _host = self._next
        torch.cuda.current_stream().wait_event(ev)
        map_tensors(dev, lambda t: (t.record_stream(torch.cuda.current_stream()), t)[1])
        self._fill()
        return dev

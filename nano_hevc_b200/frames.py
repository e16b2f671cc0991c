"""Device-side frame containers: the reference's ``PackedFrame`` and
``FrameBufferPool`` (``nano_hevc/frame.py:121-308``) with the pixels resident in HBM.

* ``DevicePackedFrame``: one contiguous allocation ``[Y | U | V]`` with the three planes as views
  (``frame.py:132-148``), int16 samples by default because that is what every kernel of the path
  consumes.  ``from_yuv420p`` uploads the raw uint8 bytes once (pinned staging) and widens them on
  the device (``nh_convert_u8_to_i16``); ``to_yuv420p`` narrows on the device with numpy's
  ``astype(np.uint8)`` semantics -- the low 8 bits (``frame.py:172-178``) -- and downloads bytes.
* ``DeviceFramePool``: ``pool_size`` frames carved out of ONE arena tensor, with the reference's
  acquire / release bookkeeping, LIFO order and error behaviour (``frame.py:224-293``), so that a
  steady-state encoder never allocates.

Only torch tensors hold memory here; the conversions call the C ABI and therefore need the GPU (the
container bookkeeping itself also works on CPU tensors, which is what the CPU-only tests use).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib


def _sizes(height: int, width: int) -> Tuple[int, int, int, int]:
    uv_h, uv_w = height // 2, width // 2
    return height * width, uv_h * uv_w, uv_h, uv_w


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _convert(src: torch.Tensor, dst: torch.Tensor) -> None:
    """uint8 -> int16 or int16 -> uint8 between two contiguous CUDA tensors of equal length."""
    if not (src.is_cuda and dst.is_cuda):
        raise RuntimeError("nano_hevc_b200: sample conversion runs on the GPU only (no CPU fallback)")
    n = src.numel()
    L = _lib.lib()
    with torch.cuda.device(src.device):
        if src.dtype == torch.uint8 and dst.dtype == torch.int16:
            _lib.check(L.nh_convert_u8_to_i16(src.data_ptr(), dst.data_ptr(), n, _stream()))
        elif src.dtype == torch.int16 and dst.dtype == torch.uint8:
            _lib.check(L.nh_convert_i16_to_u8(src.data_ptr(), dst.data_ptr(), n, _stream()))
        else:
            raise TypeError(f"unsupported conversion {src.dtype} -> {dst.dtype}")


class DevicePackedFrame:
    """``PackedFrame`` (frame.py:121-186) on a torch device."""

    __slots__ = ("_buffer", "y", "u", "v", "height", "width", "_y_size", "_uv_size")

    def __init__(self, height: int, width: int, dtype: torch.dtype = torch.int16, device="cuda",
                 _storage: torch.Tensor | None = None):
        self.height, self.width = int(height), int(width)
        self._y_size, self._uv_size, uv_h, uv_w = _sizes(self.height, self.width)
        total = self._y_size + 2 * self._uv_size
        if _storage is None:
            self._buffer = torch.zeros(total, dtype=dtype, device=device)  # frame.py:143
        else:  # a slice of a pool arena
            if _storage.numel() != total or not _storage.is_contiguous():
                raise ValueError("storage does not match the frame size")
            self._buffer = _storage
        self.y = self._buffer[:self._y_size].view(self.height, self.width)
        self.u = self._buffer[self._y_size:self._y_size + self._uv_size].view(uv_h, uv_w)
        self.v = self._buffer[self._y_size + self._uv_size:].view(uv_h, uv_w)

    # ------------------------------------------------------------------ properties
    @property
    def dtype(self) -> torch.dtype:
        return self._buffer.dtype

    @property
    def device(self) -> torch.device:
        return self._buffer.device

    @property
    def buffer(self) -> torch.Tensor:
        return self._buffer

    # ------------------------------------------------------------------ I/O
    def load_yuv420p(self, buffer) -> "DevicePackedFrame":
        """Fill this frame from raw planar YUV420p bytes (frame.py:150-156): one H2D copy of the
        uint8 samples, widened on the device when the frame holds int16."""
        total = self._buffer.numel()
        if isinstance(buffer, torch.Tensor):
            host = buffer.reshape(-1)
            if host.dtype != torch.uint8:
                raise TypeError("YUV420p data must be uint8")
        else:
            import numpy as np
            host = torch.from_numpy(np.frombuffer(buffer, dtype=np.uint8).copy())  # bytes are read-only
        if host.numel() < total:
            raise ValueError(f"buffer holds {host.numel()} bytes, frame needs {total}")
        host = host[:total]
        if self._buffer.dtype == torch.uint8:
            self._buffer.copy_(host, non_blocking=True)
        elif self._buffer.dtype == torch.int16:
            staged = host.to(self._buffer.device, non_blocking=True)
            _convert(staged, self._buffer)
        else:
            raise TypeError(f"unsupported frame dtype {self._buffer.dtype}")
        return self

    @classmethod
    def from_yuv420p(cls, buffer, height: int, width: int, dtype: torch.dtype = torch.int16,
                     device="cuda") -> "DevicePackedFrame":
        return cls(height, width, dtype=dtype, device=device).load_yuv420p(buffer)

    @classmethod
    def from_planes(cls, y: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> "DevicePackedFrame":
        """``PackedFrame.from_frame`` (frame.py:158-165): copies three planes into one allocation."""
        f = cls(y.shape[0], y.shape[1], dtype=y.dtype, device=y.device)
        f.y.copy_(y); f.u.copy_(u); f.v.copy_(v)
        return f

    def to_yuv420p(self) -> bytes:
        """Raw planar YUV420p bytes (frame.py:167-178): every sample's low 8 bits."""
        if self._buffer.dtype == torch.uint8:
            out = self._buffer
        else:
            out = torch.empty(self._buffer.numel(), dtype=torch.uint8, device=self._buffer.device)
            _convert(self._buffer, out)
        return out.cpu().numpy().tobytes()

    def planes(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return self.y, self.u, self.v

    def clear(self) -> None:
        self._buffer.zero_()  # frame.py:184-186

    def __repr__(self) -> str:
        return (f"DevicePackedFrame(height={self.height}, width={self.width}, dtype={self._buffer.dtype}, "
                f"device={self._buffer.device})")


class DeviceFramePool:
    """``FrameBufferPool`` (frame.py:189-308): ``pool_size`` packed frames in one arena tensor."""

    __slots__ = ("_arena", "_pool", "_available", "_in_use", "height", "width", "dtype")

    def __init__(self, height: int, width: int, pool_size: int = 4, dtype: torch.dtype = torch.int16,
                 device="cuda"):
        self.height, self.width, self.dtype = int(height), int(width), dtype
        y, uv, _, _ = _sizes(self.height, self.width)
        per = y + 2 * uv
        self._arena = torch.zeros(pool_size * per, dtype=dtype, device=device)
        self._pool: List[DevicePackedFrame] = [
            DevicePackedFrame(height, width, dtype=dtype, device=device, _storage=self._arena[i * per:(i + 1) * per])
            for i in range(pool_size)]
        self._available: List[int] = list(range(pool_size))
        self._in_use: set = set()

    def acquire(self, clear: bool = True) -> Tuple[int, DevicePackedFrame]:
        if not self._available:
            raise RuntimeError(f"No buffers available in pool. In use: {len(self._in_use)}, Total: {len(self._pool)}")
        idx = self._available.pop()
        self._in_use.add(idx)
        frame = self._pool[idx]
        if clear:
            frame.clear()
        return idx, frame

    def release(self, idx: int) -> None:
        if idx not in self._in_use:
            raise ValueError(f"Buffer {idx} is not currently in use")
        self._in_use.remove(idx)
        self._available.append(idx)

    @property
    def available_count(self) -> int:
        return len(self._available)

    @property
    def in_use_count(self) -> int:
        return len(self._in_use)

    @property
    def pool_size(self) -> int:
        return len(self._pool)

    def __repr__(self) -> str:
        return (f"DeviceFramePool(height={self.height}, width={self.width}, "
                f"available={self.available_count}/{self.pool_size})")

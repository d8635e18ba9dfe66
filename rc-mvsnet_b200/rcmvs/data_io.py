"""f4 — PFM wire format (reference datasets/data_io.py:8-73): `read_pfm(filename) -> (array, scale)` and
`save_pfm(filename, image, scale=1)` with the reference's signatures, on the native reader / writer of librcmvs
(csrc/pfm_io.cu: rows streamed bottom-up, no flipped temporary).  `save_pfm` also accepts a CUDA tensor: it is copied into
a pinned staging buffer first."""
import ctypes

import numpy as np
import torch

from . import _ffi


def read_pfm(filename):
    L = _ffi.lib()
    w, h, c, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    rc = L.rcmvs_pfm_read_header(filename.encode(), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c), ctypes.byref(s))
    if rc != 0:
        raise Exception(L.rcmvs_last_error().decode())
    shape = (h.value, w.value, 3) if c.value == 3 else (h.value, w.value)
    data = np.empty(shape, dtype=np.float32)
    rc = L.rcmvs_pfm_read(filename.encode(), data.ctypes.data_as(ctypes.c_void_p), h.value, w.value, c.value)
    if rc != 0:
        raise Exception(L.rcmvs_last_error().decode())
    return data, abs(s.value)


def save_pfm(filename, image, scale=1):
    if isinstance(image, torch.Tensor):
        if image.dtype != torch.float32:
            raise Exception("Image dtype must be float32.")
        host = torch.empty(image.shape, dtype=torch.float32, pin_memory=image.is_cuda)
        host.copy_(image)
        image = host.numpy()
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        ch = 3
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        ch = 1
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    img = np.ascontiguousarray(image)
    L = _ffi.lib()
    rc = L.rcmvs_pfm_write(filename.encode(), img.ctypes.data_as(ctypes.c_void_p), img.shape[0], img.shape[1], ch, float(scale))
    if rc != 0:
        raise Exception(L.rcmvs_last_error().decode())

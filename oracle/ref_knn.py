"""ctypes driver of oracle/_ref/libknn_ref.so — the reference simple-knn (distCUDA2) compiled for sm_100a
from /root/reference by oracle/ref_build/Makefile.

TEST INFRASTRUCTURE.  Only tests/ and tests/golden/make_golden_knn.py may import this; nothing under
eogs2_b200/ does.  Mirrors submodules/simple-knn/spatial.cu:15-26 (allocate `means`, call SimpleKNN::knn).
The reference launches on the legacy default stream and synchronises internally.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "_ref" / "libknn_ref.so"
_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle/ref_build` where /root/reference exists")
        _lib = C.CDLL(os.fspath(LIB_PATH))
        _lib.eogs_ref_knn.restype = C.c_int
    pts = points.contiguous().float()
    P = pts.shape[0]
    means = torch.full((P,), 0.0, dtype=torch.float32, device=pts.device)
    torch.cuda.synchronize()
    rc = _lib.eogs_ref_knn(C.c_int(P), C.c_void_p(pts.data_ptr()), C.c_void_p(means.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"reference SimpleKNN::knn failed with cudaError {rc}")
    return means

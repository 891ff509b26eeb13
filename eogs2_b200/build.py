"""Build recipe for libeogs_raster.so (sm_100a only, in-tree).

`python -m eogs2_b200.build` compiles every csrc/*.cu with nvcc for
`-gencode arch=compute_100a,code=sm_100a -lineinfo` and links one shared library,
`eogs2_b200/libeogs_raster.so`, whose only exports are the extern "C" entry points
declared in include/eogs_raster.h.  nvcc cross-compiles without a GPU.  The library
does not link against torch: the Python mirror talks to it through ctypes.

No fast-math: key/range parity with the reference depends on IEEE div/sqrt and the
accurate expf (the reference build uses nvcc defaults, DGR/setup.py:32-38).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
# Developer A/B builds: EOGS_NVCC_DEFS="-DX=1 -DY=2" EOGS_LIB_SUFFIX=_x builds libeogs_raster_x.so next to the
# product library (objects under csrc/build_x/); select it at run time with EOGS_RASTER_LIB=<path>.
_SUFFIX = os.environ.get("EOGS_LIB_SUFFIX", "")
_EXTRA_DEFS = os.environ.get("EOGS_NVCC_DEFS", "").split()
LIB = PKG_DIR / f"libeogs_raster{_SUFFIX}.so"
BUILD = PKG_DIR / "csrc" / f"build{_SUFFIX}"
# The instrumented twin of the product library: the blend kernels count evaluated / blended (pixel, Gaussian) pairs
# (SURVEY.md section 8d).  bench.py runs it ONCE, outside every timed region, to report those counts; nothing else loads it.
COUNT_LIB = PKG_DIR / "libeogs_raster_count.so"
SOURCES = ["cabi.cu", "preprocess.cu", "binning.cu", "blend_fwd.cu", "blend_bwd.cu", "preprocess_bwd.cu", "resample.cu", "ssim_loss.cu", "optim.cu", "knn.cu", "dsm.cu", "nvls.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                     "-Xptxas", "-v", "--expt-relaxed-constexpr", "-Wno-deprecated-declarations"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libeogs_raster.so for sm_100a)")


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "eogs_raster.h", Path(__file__)]
    return max(p.stat().st_mtime for p in hdrs)


def _compile(src: str, nvcc: str, force: bool, log: list, build_dir: Path, defs: list) -> Path:
    s = CSRC / src
    o = build_dir / (src.replace(".cu", ".o"))
    if not force and o.exists() and o.stat().st_mtime > max(s.stat().st_mtime, _deps_mtime()):
        return o
    cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", str(s), "-o", str(o)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append((src, r.stderr))
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return o


def build(force: bool = False, verbose: bool = False, lib: Path = None, build_dir: Path = None, defs: list = None) -> Path:
    nvcc = nvcc_path()
    lib = lib or LIB
    build_dir = build_dir or BUILD
    defs = list(_EXTRA_DEFS if defs is None else defs)
    build_dir.mkdir(parents=True, exist_ok=True)
    log: list = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, nvcc, force, log, build_dir, defs), SOURCES))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not lib.exists() or lib.stat().st_mtime < newest:
        cmd = [nvcc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", str(lib), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for src, err in log:
            print(f"--- {src}\n{err}")
    if log:
        (build_dir / "ptxas.log").write_text("\n".join(f"--- {s}\n{e}" for s, e in log))
    return lib


def build_count(force: bool = False) -> Path:
    """The instrumented twin (only the two blend translation units differ; everything is recompiled for simplicity)."""
    return build(force, False, COUNT_LIB, PKG_DIR / "csrc" / "build_count", ["-DEOGS_COUNT_PAIRS=1"])


# the rasterizer path: what a profile of the render / backward kernels depends on (the loss, resample, optimiser, knn and
# DSM kernels live in their own files and do not change it)
RASTER_SOURCES = ["cabi.cu", "preprocess.cu", "binning.cu", "blend_fwd.cu", "blend_bwd.cu", "preprocess_bwd.cu",
                  "common.cuh", "blend_common.cuh", "f32x2.cuh", "geom_math.cuh"]


def source_hash() -> str:
    """sha256 over the CUDA sources of the rasterizer path and the public header: ties a committed ncu capture
    (profiles/ncu_current.json) to the kernels it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for p in sorted([CSRC / n for n in RASTER_SOURCES] + [PKG_DIR.parent / "include" / "eogs_raster.h"]):
        h.update(p.name.encode()); h.update(p.read_bytes())
    return h.hexdigest()


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
    if not _SUFFIX and "--no-count" not in sys.argv:
        print(build_count(force="--force" in sys.argv))

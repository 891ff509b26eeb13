"""CPU: the C-ABI library loads and exports every symbol include/eogs_raster.h declares; host-only
size queries behave; the Python surface has the reference's names and field order."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "eogs_raster.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"EOGS_API\s+[\w\s\*]+?\b(eogs_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("eogs_forward_geometry", "eogs_forward_render", "eogs_rasterize_forward", "eogs_backward",
                 "eogs_mark_visible", "eogs_export_state", "eogs_geom_bytes", "eogs_image_bytes",
                 "eogs_binning_bytes", "eogs_grad_scratch_floats", "eogs_point_list_words", "eogs_last_error", "eogs_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from eogs2_b200 import _cabi, build
    build.build()
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the ctypes table covers the header one to one
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_only_the_c_abi_is_exported():
    from eogs2_b200 import _cabi
    out = subprocess.run(["nm", "-D", "--defined-only", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert exported and all(s.startswith("eogs_") for s in exported), exported


def test_size_queries_are_host_only_and_monotonic():
    from eogs2_b200 import _cabi
    lib = _cabi.load()
    assert lib.eogs_abi_version() == _cabi.ABI_VERSION == 7
    assert lib.eogs_grad_scratch_floats(1000) == 16 * 1000 + 16
    assert lib.eogs_point_list_words(1001) == 1001 + 251
    g1, g2 = lib.eogs_geom_bytes(1000), lib.eogs_geom_bytes(1_000_000)
    assert 0 < g1 < g2 and g2 >= 1_000_000 * (48 + 4 + 8 + 4 * 6)
    assert lib.eogs_image_bytes(2048, 2048) >= 2048 * 2048 * 8 + 16384 * 8
    assert lib.eogs_binning_bytes(2048, 2048, 10_000_000) >= 10_000_000 * 8      # one 8-byte run per instance at most
    assert lib.eogs_geom_bytes(0) > 0


def test_python_surface_matches_reference_names():
    import diff_gaussian_rasterization as d
    assert d.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug", "antialiasing")     # DGR __init__.py:219-232
    import inspect
    sig = inspect.signature(d.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                    "rotations", "cov3D_precomp"]          # DGR __init__.py:250-260
    assert hasattr(d.GaussianRasterizer, "markVisible") and callable(d.rasterize_gaussians)


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ types in the signatures) and a C program
    must link against the shared library and call the host-only entry points."""
    from eogs2_b200 import _cabi
    src = tmp_path / "use.c"
    src.write_text('#include <stdio.h>\n#include "eogs_raster.h"\n'
                   'int main(void) { printf("%d %zu %zu\\n", eogs_abi_version(), eogs_geom_bytes(1000), eogs_knn_bytes(1000)); return 0; }\n')
    exe = tmp_path / "use"
    lib_dir = str(_cabi.LIB_PATH.parent)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
           f"-L{lib_dir}", "-leogs_raster", f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == _cabi.ABI_VERSION and int(out[1]) > 0 and int(out[2]) > 0


def declared_parameters():
    """name -> list of parameter declarations, parsed from the header."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"EOGS_API\s+[\w\s\*]+?\b(eogs_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params == ["void"] else params
    return out


def test_ctypes_signatures_follow_the_header():
    """Every binding has as many arguments as the C declaration, pointers are bound as pointers and scalars as the
    matching ctypes scalar — a drifted ctypes table corrupts the call silently."""
    import ctypes as C
    from eogs2_b200 import _cabi
    decl = declared_parameters()
    assert sorted(decl) == sorted(_cabi.SIGNATURES)
    scalar = {"int": C.c_int, "float": C.c_float, "double": C.c_double, "uint32_t": C.c_uint32, "size_t": C.c_size_t,
              "long long": C.c_longlong, "unsigned long long": C.c_ulonglong}
    for name, params in decl.items():
        res, args = _cabi.SIGNATURES[name]
        assert len(args) == len(params), (name, len(args), params)
        for a, p in zip(args, params):
            is_ptr = "*" in p or p.startswith("eogs_stream_t") or p.startswith("eogs_alloc_fn")
            if is_ptr:
                assert a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents") or issubclass(a, C._CFuncPtr) \
                    or a.__name__.startswith("LP_"), (name, p, a)
            else:
                ctype = " ".join(p.replace("const ", "").split()[:-1])
                assert scalar.get(ctype) is a, (name, p, a)

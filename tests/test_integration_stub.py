"""CPU: the reference-side binding shown in INTEGRATION.md section 2 compiles.  integration/rasterize_points_eogs.cpp (the pybind
module `_C` of DGR/ext.cpp:15-18 on top of include/eogs_raster.h) and integration/spatial_eogs.cpp (simple_knn._C) are
syntax-checked by g++ against the torch headers, and the code blocks quoted in INTEGRATION.md are the files' contents."""
import re
import subprocess
import sysconfig
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("name", ["rasterize_points_eogs.cpp", "spatial_eogs.cpp"])
def test_stub_compiles_against_torch_headers(name):
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include", str(ROOT / "include")]
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-DTORCH_EXTENSION_NAME=_C", *[f"-I{p}" for p in inc], str(ROOT / "integration" / name)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]


def test_integration_md_quotes_the_files():
    md = (ROOT / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```cpp\n(.*?)```", md, re.S)
    for name in ("rasterize_points_eogs.cpp", "spatial_eogs.cpp"):
        src = (ROOT / "integration" / name).read_text()
        assert any(b.strip() == src.strip() for b in blocks), f"INTEGRATION.md does not quote integration/{name} verbatim"

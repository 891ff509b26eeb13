"""Developer probe (GPU): one fuzz case in detail — every gradient of this library and of the compiled reference (two
runs) against the float64-accumulating C oracle.   python tools/check_case.py P W H kind seed aa mod sun"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_parity_gpu as T                       # noqa: E402
from oracle import c_oracle as O                  # noqa: E402
from oracle import ref_rasterizer as R            # noqa: E402

NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", None, "dL_dscales", "dL_drotations"]
P, W, H, kind, seed, aa, mod, sun = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), \
    sys.argv[6] == "1", float(sys.argv[7]), sys.argv[8] == "1"
dev = torch.device("cuda:0")
c = T.make_case(P, W, H, kind, seed, aa, mod, sun)
st, ex, g = T.run_mine(dev, c)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in c.items()}
empty, campos = torch.empty(0, device=dev), torch.zeros(3, device=dev)
rs = R.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], mod, empty, d["view"], d["view"],
               1.0, 1.0, H, W, campos, False, aa)
gr = [R.backward(rs, d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], mod, empty, d["view"],
                 d["view"], 1.0, 1.0, d["dL_dcolor"], d["dL_dinvdepth"], campos, aa) for _ in range(2)]
o = O.forward(c["means3D"].numpy(), c["scales"].numpy(), c["rotations"].numpy(), c["opacities"].numpy(), c["colors"].numpy(),
              c["view"].numpy(), c["bg"].numpy(), W, H, mod, aa)
go = O.backward(o, c["dL_dcolor"].numpy(), c["dL_dinvdepth"].numpy())
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel()) /
                         (np.linalg.norm(np.asarray(b, np.float64).ravel()) + 1e-300))      # noqa: E731
out = {"instances": st.num_rendered}
for nm, t in zip(NAMES, g):
    if nm is None:
        continue
    out[nm] = {"ours_vs_f64": rel(t.cpu().numpy(), go[nm]), "ref_vs_f64": rel(gr[0][nm].cpu().numpy(), go[nm]),
               "ref_rerun_vs_f64": rel(gr[1][nm].cpu().numpy(), go[nm]), "ours_vs_ref": rel(t.cpu().numpy(), gr[0][nm].cpu().numpy()),
               "ref_vs_ref_rerun": rel(gr[1][nm].cpu().numpy(), gr[0][nm].cpu().numpy())}
print(json.dumps(out, indent=1))

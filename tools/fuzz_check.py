"""For the fuzz cases whose gradients differ from the compiled reference by more than 1e-3: which side is off?
Both are compared with the float64-accumulating C oracle (oracle/eogs_oracle.c).  Development tool."""
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

rows = json.load(open(sys.argv[1]))["rows"]
dev = torch.device("cuda:0")
out = []
for r in rows:
    if not r["bad"]:
        continue
    base_W, base_H = (r["W"] // 2, r["H"] // 2) if r["sun"] else (r["W"], r["H"])
    c = T.make_case(r["P"], base_W, base_H, r["kind"], r["seed"], r["aa"], r["mod"], r["sun"])
    W, H = c["W"], c["H"]
    st, ex, g = T.run_mine(dev, c)
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in c.items()}
    empty, campos = torch.empty(0, device=dev), torch.zeros(3, device=dev)
    rs = R.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], r["mod"], empty,
                   d["view"], d["view"], 1.0, 1.0, H, W, campos, False, r["aa"])
    gr = R.backward(rs, d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], r["mod"], empty,
                    d["view"], d["view"], 1.0, 1.0, d["dL_dcolor"], d["dL_dinvdepth"], campos, r["aa"])
    torch.cuda.synchronize()
    o = O.forward(c["means3D"].numpy(), c["scales"].numpy(), c["rotations"].numpy(), c["opacities"].numpy(),
                  c["colors"].numpy(), c["view"].numpy(), c["bg"].numpy(), W, H, r["mod"], r["aa"])
    go = O.backward(o, c["dL_dcolor"].numpy(), c["dL_dinvdepth"].numpy())
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", None, "dL_dscales", "dL_drotations"]
    row = {"case": r["case"], "P": r["P"], "W": W, "H": H, "kind": r["kind"], "n_contrib_max": int(o["n_contrib"].max())}
    for nm, t in zip(names, g):
        if nm is None:
            continue
        row[nm] = {"ours_vs_oracle": float(T.rel(t.cpu().numpy(), go[nm])),
                   "ref_vs_oracle": float(T.rel(gr[nm].cpu().numpy(), go[nm])),
                   "ours_vs_ref": float(T.rel(t.cpu().numpy(), gr[nm].cpu().numpy()))}
    out.append(row)
    print(row, file=sys.stderr)
print(json.dumps(out, indent=1))

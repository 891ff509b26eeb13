"""Debug: one fuzz case, per-Gaussian gradient comparison ours vs the float64 oracle."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_parity_gpu as T
from oracle import c_oracle as O
rows = [r for r in json.load(open(sys.argv[1]))["rows"] if r["case"] == int(sys.argv[2])]
r = rows[0]; dev = torch.device("cuda:0")
bw, bh = (r["W"] // 2, r["H"] // 2) if r["sun"] else (r["W"], r["H"])
c = T.make_case(r["P"], bw, bh, r["kind"], r["seed"], r["aa"], r["mod"], r["sun"])
st, ex, g = T.run_mine(dev, c)
o = O.forward(c["means3D"].numpy(), c["scales"].numpy(), c["rotations"].numpy(), c["opacities"].numpy(),
              c["colors"].numpy(), c["view"].numpy(), c["bg"].numpy(), c["W"], c["H"], r["mod"], r["aa"])
go = O.backward(o, c["dL_dcolor"].numpy(), c["dL_dinvdepth"].numpy())
np.set_printoptions(precision=8, linewidth=200)
print("ours dL_dcolors\n", g[1].cpu().numpy()[:4]); print("oracle\n", go["dL_dcolors"][:4])
print("ours dL_dopacity", g[2].cpu().numpy()[:4].ravel(), "oracle", go["dL_dopacity"][:4].ravel())
print("ours dL_dmeans2D\n", g[0].cpu().numpy()[:4]); print("oracle\n", go["dL_dmeans2D"][:4])
print("conic_opacity", o["conic_opacity"][:4], "means2D", o["means2D"][:4], "radii", o["radii"][:4])

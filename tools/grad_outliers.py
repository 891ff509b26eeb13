"""Developer probe (GPU): per-Gaussian, per-element view of the gradient differences between this library
and the compiled reference (and, at small sizes, the float64-accumulating C oracle), to choose per-element
tolerances for tests/test_parity_gpu.py instead of a global L2 ratio.  Prints JSON."""
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


def stats(a, b, vis):
    a, b = np.asarray(a, np.float64).reshape(len(vis), -1)[vis], np.asarray(b, np.float64).reshape(len(vis), -1)[vis]
    d = np.abs(a - b)
    row_ref = np.abs(b).max(1)
    row_err = d.max(1)
    gmax = float(np.abs(b).max())
    med = float(np.median(row_ref[row_ref > 0])) if (row_ref > 0).any() else 0.0
    out = {"global_rel_l2": float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)), "ref_absmax": gmax,
           "ref_row_median": med, "err_absmax": float(row_err.max())}
    with np.errstate(divide="ignore", invalid="ignore"):
        rr = row_err / row_ref
    rr = rr[np.isfinite(rr)]
    for q in (0.5, 0.99, 0.9999, 1.0):
        out[f"row_rel_q{q}"] = float(np.quantile(rr, q)) if rr.size else None
    # element-wise: smallest atol (as a multiple of the tensor's row-median scale) under which |a-b| <= 1e-3|b| + atol holds
    need = np.maximum(d - 1e-3 * np.abs(b), 0.0).max()
    out["atol_needed"] = float(need)
    out["atol_needed_over_row_median"] = float(need / med) if med else None
    out["atol_needed_over_absmax"] = float(need / gmax) if gmax else None
    # rows violating 1e-3 * row max + 1e-5 * global median
    out["rows_over_1e-3_rowmax"] = int((row_err > 1e-3 * row_ref + 1e-4 * med).sum())
    return out


def main():
    dev = torch.device("cuda:0")
    res = []
    for (P, W, H, kind, seed, aa, sun) in [(50_000, 512, 512, "trained", 1337, False, False), (50_000, 512, 512, "init", 1337, False, False),
                                           (200_000, 1000, 700, "init", 3, True, False), (1_000_000, 2048, 2048, "trained", 1337, False, False)]:
        c = T.make_case(P, W, H, kind, seed, aa, 1.0, sun)
        st, ex, g = T.run_mine(dev, c)
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in c.items()}
        empty, campos = torch.empty(0, device=dev), torch.zeros(3, device=dev)
        rs = R.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                       d["view"], d["view"], 1.0, 1.0, c["H"], c["W"], campos, False, aa)
        gr = R.backward(rs, d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                        d["view"], d["view"], 1.0, 1.0, d["dL_dcolor"], d["dL_dinvdepth"], campos, aa)
        gr2 = R.backward(rs, d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                         d["view"], d["view"], 1.0, 1.0, d["dL_dcolor"], d["dL_dinvdepth"], campos, aa)
        torch.cuda.synchronize()
        vis = (st.radii > 0).cpu().numpy()
        row = {"case": [P, W, H, kind, aa], "instances": st.num_rendered}
        go = None
        if P <= 50_000:
            o = O.forward(c["means3D"].numpy(), c["scales"].numpy(), c["rotations"].numpy(), c["opacities"].numpy(),
                          c["colors"].numpy(), c["view"].numpy(), c["bg"].numpy(), W, H, 1.0, aa)
            go = O.backward(o, c["dL_dcolor"].numpy(), c["dL_dinvdepth"].numpy())
        for nm, t in zip(NAMES, g):
            if nm is None:
                continue
            row[nm] = {"ours_vs_ref": stats(t.cpu().numpy(), gr[nm].cpu().numpy(), vis),
                       "ref_vs_ref_rerun": stats(gr2[nm].cpu().numpy(), gr[nm].cpu().numpy(), vis)}
            if go is not None:
                row[nm]["ours_vs_f64oracle"] = stats(t.cpu().numpy(), go[nm], vis)
                row[nm]["ref_vs_f64oracle"] = stats(gr[nm].cpu().numpy(), go[nm], vis)
        res.append(row)
        print(json.dumps(row), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/grad_outliers.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_*.npz by running the REFERENCE rasterizer (oracle/_ref/libeogs_ref.so,
built from /root/reference by oracle/ref_build/Makefile) on a B200:

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

The reference has no golden vectors of its own; these fixtures are what pins
oracle/eogs_oracle.c (CPU test, bit-exact on integers and on every float that feeds a key) and
the CUDA path (GPU test).  Cases are small so the files stay in the tens of kilobytes.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from eogs2_b200 import scene as S                    # noqa: E402
from oracle import ref_rasterizer as R               # noqa: E402

CASES = {
    # name: (P, W, H, kind, seed, antialiasing, scale_modifier, use_cov3D_precomp, sun)
    "trained_ragged": (2000, 100, 84, "trained", 1337, False, 1.0, False, False),
    "init_aa": (1500, 64, 64, "init", 11, True, 1.0, False, False),
    "cov_precomp_mod": (600, 48, 80, "trained", 5, False, 1.3, True, False),
    "sun_view": (1200, 40, 40, "trained", 23, False, 1.0, False, True),
}


def case_inputs(name):
    P, W, H, kind, seed, aa, mod, precomp, sun = CASES[name]
    sc = S.make_scene(P, kind, seed)
    view = S.make_camera(seed)
    if sun:
        view = S.sun_camera(view); W, H = 2 * W, 2 * H
    colors = S.colors_precomp(sc, view)
    bg = S.background(seed)
    dcol, dinv = S.upstream_grads(5, H, W, seed, True)
    return dict(P=P, W=W, H=H, aa=aa, mod=mod, precomp=precomp, means3D=sc.means3D, scales=sc.scales,
                rotations=sc.rotations, opacities=sc.opacities, colors=colors, view=view, bg=bg,
                dL_dcolor=dcol, dL_dinvdepth=dinv)


def main():
    out_dir = ROOT / "gpurun_out" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    dev = torch.device("cuda:0")
    campos = torch.zeros(3, device=dev)
    empty = torch.empty(0, device=dev)
    for name in CASES:
        c = case_inputs(name)
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in c.items()}
        scales, rots, cov = d["scales"], d["rotations"], empty
        if c["precomp"]:
            # take the reference's own Sigma3D (scale_modifier applied) as the precomputed covariance
            st0 = R.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], scales, rots, c["mod"], empty,
                            d["view"], d["view"], 1.0, 1.0, c["H"], c["W"], campos, False, c["aa"])
            cov = R.export_state(st0)["cov3D"].clone()
            scales, rots = empty, empty
        st = R.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], scales, rots, c["mod"], cov,
                       d["view"], d["view"], 1.0, 1.0, c["H"], c["W"], campos, False, c["aa"])
        ex = R.export_state(st)
        g = R.backward(st, d["bg"], d["means3D"], d["colors"], d["opacities"], scales, rots, c["mod"], cov,
                       d["view"], d["view"], 1.0, 1.0, d["dL_dcolor"], d["dL_dinvdepth"], campos, c["aa"])
        torch.cuda.synchronize()
        vis = (ex["radii"] > 0)
        save = dict(num_rendered=np.int64(st.num_rendered), color=st.color.cpu().numpy(),
                    invdepth=st.invdepth.cpu().numpy(), cov3D_precomp=cov.cpu().numpy())
        for k in ("radii", "tiles_touched", "point_list", "keys_sorted", "ranges", "n_contrib", "final_T"):
            save[k] = ex[k].cpu().numpy()
        for k in ("depths", "means2D", "conic_opacity", "cov3D"):      # undefined for culled Gaussians: zero them
            t = ex[k].clone()
            t[~vis] = 0
            save[k] = t.cpu().numpy()
        for k in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dscales",
                  "dL_drotations", "dL_dconic"):
            save["g_" + k] = g[k].cpu().numpy()
        terms = R.grad_viewmatrix_terms(g, d["means3D"], d["view"], c["H"], c["W"])
        save["g_view_mean_bias"] = (terms["mean_term"] + terms["bias_term"]).cpu().numpy()
        np.savez_compressed(out_dir / f"ref_{name}.npz", **save)
        print(name, "I =", st.num_rendered, "visible =", int(vis.sum()), "bytes =",
              (out_dir / f"ref_{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()

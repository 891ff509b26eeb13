"""Developer probe (GPU): eogs2_b200 vs the compiled reference (oracle/_ref) on one synthetic scene.
Prints bit-level mismatch counts for the binning state, image / gradient errors and timings.
Not a test and not the bench; see tests/ and bench.py for those."""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import eogs2_b200 as E                                   # noqa: E402
from eogs2_b200 import scene as S                        # noqa: E402
from oracle import ref_rasterizer as R                   # noqa: E402


def bits_equal(a, b, mask=None):
    a = a.contiguous().view(torch.int32) if a.dtype == torch.float32 else a
    b = b.contiguous().view(torch.int32) if b.dtype == torch.float32 else b
    ne = a != b
    if mask is not None:
        ne = ne & mask
    return int(ne.sum().item())


def rel_err(a, b):
    return float((a - b).abs().max().item()), float(((a - b).norm() / (b.norm() + 1e-30)).item())


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(reps)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=50000)
    ap.add_argument("--W", type=int, default=512)
    ap.add_argument("--H", type=int, default=512)
    ap.add_argument("--kind", default="trained")
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--aa", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--sun", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sc = S.make_scene(args.P, args.kind, args.seed).to(dev)
    view = S.make_camera(args.seed)
    W, H = args.W, args.H
    if args.sun:
        view = S.sun_camera(view); W, H = 2 * W, 2 * H
    view = view.to(dev)
    colors = S.colors_precomp(sc, view)
    bg = S.background(args.seed).to(dev)
    dcol, dinv = S.upstream_grads(5, H, W, args.seed, True)
    dcol, dinv = dcol.to(dev), dinv.to(dev)
    campos = torch.zeros(3, device=dev)
    empty = torch.empty(0, device=dev)

    st = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view,
                                 H, W, args.aa)
    mine = E.export_state(st)
    gm = E.rasterize_backward_raw(st, bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty,
                                  view, view, dcol, dinv, args.aa)
    torch.cuda.synchronize()
    print(f"P={args.P} {W}x{H} kind={args.kind} aa={args.aa}: I={st.num_rendered} I/P={st.num_rendered/args.P:.2f} "
          f"visible={(st.radii > 0).sum().item()} mean list={st.num_rendered / (((W+15)//16)*((H+15)//16)):.1f}")

    if R.available():
        rs = R.forward(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, view,
                       1.0, 1.0, H, W, campos, False, args.aa)
        ref = R.export_state(rs)
        gr = R.backward(rs, bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, view,
                        1.0, 1.0, dcol, dinv, campos, args.aa)
        torch.cuda.synchronize()
        vis = ref["radii"] > 0
        print("num_rendered ref/mine:", rs.num_rendered, st.num_rendered)
        print("radii mismatches      :", bits_equal(mine["radii"], ref["radii"]))
        print("tiles_touched mismatch:", bits_equal(mine["tiles_touched"], ref["tiles_touched"]))
        print("depth bit mismatches  :", bits_equal(mine["depths"], ref["depths"], vis))
        print("means2D bit mismatches:", bits_equal(mine["means2D"], ref["means2D"], vis[:, None]))
        print("conic_op bit mismatch :", bits_equal(mine["conic_opacity"], ref["conic_opacity"], vis[:, None]))
        if rs.num_rendered == st.num_rendered:
            print("point_list mismatches :", bits_equal(mine["point_list"], ref["point_list"]))
            print("sorted key mismatches :", int((mine["keys_sorted"] != ref["keys_sorted"]).sum().item()))
        print("ranges mismatches     :", bits_equal(mine["ranges"], ref["ranges"]))
        print("n_contrib mismatches  :", bits_equal(mine["n_contrib"], ref["n_contrib"]))
        print("final_T bit mismatches:", bits_equal(mine["final_T"], ref["final_T"]))
        print("color bit mismatches  :", bits_equal(st.color, rs.color), "max abs", rel_err(st.color, rs.color)[0])
        print("invdepth bit mismatch :", bits_equal(st.invdepth, rs.invdepth), "max abs", rel_err(st.invdepth, rs.invdepth)[0])
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", None, "dL_dscales", "dL_drotations"]
        for nm, t in zip(names, gm):
            if nm is None or t is None:
                continue
            ma, rl = rel_err(t, gr[nm])
            print(f"{nm:14s} max abs {ma:.3e} rel l2 {rl:.3e}  |ref| {gr[nm].abs().max().item():.3e}")
        terms = R.grad_viewmatrix_terms(gr, sc.means3D, view, H, W)
        gv = E.assemble_grad_viewmatrix(gm[7], view, W, H)
        print("grad_view mine:\n", gv)
        print("grad_view ref (mean+bias terms):\n", terms["mean_term"] + terms["bias_term"])
        print("cam_sums mine:", gm[7])
        print("ref cov term (racy):\n", terms["cov_term"])

    if args.time:
        def mine_fwd():
            return E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty,
                                           view, H, W, args.aa)

        def mine_fb():
            s = mine_fwd()
            E.rasterize_backward_raw(s, bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty,
                                     view, view, dcol, dinv, args.aa)
        print(f"mine fwd {timeit(mine_fwd):.3f} ms   fwd+bwd {timeit(mine_fb):.3f} ms")
        if R.available():
            def ref_fwd():
                return R.forward(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, view,
                                 1.0, 1.0, H, W, campos, False, args.aa)

            def ref_fb():
                s = ref_fwd()
                R.backward(s, bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, view,
                           1.0, 1.0, dcol, dinv, campos, args.aa)
            print(f"ref  fwd {timeit(ref_fwd):.3f} ms   fwd+bwd {timeit(ref_fb):.3f} ms")


if __name__ == "__main__":
    main()

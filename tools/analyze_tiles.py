"""Developer probe (GPU): where do the blend kernels spend their pair evaluations?
For a sample of tiles of the bench scene, counts list entries / evaluated / accepted pairs at tile,
warp (8x4 patch) and pixel granularity.  Guides culling work; not a test."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import eogs2_b200 as E                                   # noqa: E402
from eogs2_b200 import scene as S                        # noqa: E402


def main():
    P, W, H = 1_000_000, 2048, 2048
    dev = torch.device("cuda:0")
    sc = S.make_scene(P, "trained", 1337).to(dev)
    view = S.make_camera(1337).to(dev)
    colors = S.colors_precomp(sc, view)
    bg = S.background(1337).to(dev)
    st = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0,
                                 torch.empty(0, device=dev), view, H, W)
    ex = E.export_state(st)
    ranges = ex["ranges"].long()
    pl = ex["point_list"].long()
    m2, co = ex["means2D"], ex["conic_opacity"]
    ncon = ex["n_contrib"].view(H, W).long()
    gx = W // 16
    g = torch.Generator().manual_seed(0)
    tiles = torch.randint(0, ranges.shape[0], (96,), generator=g).tolist()
    tot = dict(entries=0, upto_max=0, tile_surv=0, warp_pairs=0, warp_surv=0, pix_eval=0, pix_acc=0, pix_evalT=0)
    for t in tiles:
        r0, r1 = ranges[t].tolist()
        n = r1 - r0
        if n <= 0:
            continue
        ty, tx = divmod(t, gx)
        ys = torch.arange(ty * 16, ty * 16 + 16, device=dev).float()
        xs = torch.arange(tx * 16, tx * 16 + 16, device=dev).float()
        py, px = torch.meshgrid(ys, xs, indexing="ij")
        ids = pl[r0:r1]
        dx = m2[ids, 0][None, None, :] - px[:, :, None]
        dy = m2[ids, 1][None, None, :] - py[:, :, None]
        c = co[ids]
        power = -0.5 * (c[:, 0] * dx * dx + c[:, 2] * dy * dy) - c[:, 1] * dx * dy
        alpha = torch.clamp(c[:, 3] * torch.exp(power), max=0.99)
        acc = (power <= 0) & (alpha >= 1 / 255.)                       # [16,16,n] would-accept (ignoring termination)
        nc = ncon[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16]           # last contributor per pixel
        nmax = int(nc.max())
        idx = torch.arange(n, device=dev)[None, None, :]
        alive = idx < nc[:, :, None]                                    # evaluated & possibly accepted by this pixel
        # forward evaluates entry j for a pixel while the pixel is not done; done happens right after nc (approx.)
        tot["entries"] += n
        tot["upto_max"] += nmax
        tile_s = acc[:, :, :nmax].any(0).any(0)
        tot["tile_surv"] += int(tile_s.sum())
        # warps = 8x4 patches: [4 rows of patches, 2 cols]
        a4 = acc[:, :, :nmax].view(4, 4, 2, 8, nmax).permute(0, 2, 1, 3, 4).reshape(8, 32, nmax)
        al4 = alive[:, :, :nmax].view(4, 4, 2, 8, nmax).permute(0, 2, 1, 3, 4).reshape(8, 32, nmax)
        warp_alive = al4.any(1)                                         # warp still has a live pixel at entry j
        tot["warp_pairs"] += int(warp_alive.sum())
        tot["warp_surv"] += int(((a4 & al4).any(1)).sum())
        tot["pix_eval"] += int(al4.sum())
        tot["pix_acc"] += int((a4 & al4).sum())
    k = len(tiles)
    print({a: round(b / k, 1) for a, b in tot.items()}, "per tile")
    print(f"tile-level survivors / entries up to max n_contrib: {tot['tile_surv'] / max(tot['upto_max'], 1):.3f}")
    print(f"warp-pairs with any accepting lane / live warp-pairs: {tot['warp_surv'] / max(tot['warp_pairs'], 1):.3f}")
    print(f"accepted pixel-pairs / evaluated pixel-pairs: {tot['pix_acc'] / max(tot['pix_eval'], 1):.3f}")
    print(f"mean n_contrib {ncon.float().mean().item():.1f}  mean list {st.num_rendered / ranges.shape[0]:.1f}")


if __name__ == "__main__":
    main()

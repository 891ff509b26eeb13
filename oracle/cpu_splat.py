"""PyTorch-on-CPU evaluation of the EOGS++ splat equations, with autograd for the backward.

TEST INFRASTRUCTURE and the reported CPU baseline (bench.py `cpu_baseline`, kind "port"):
the reference has no CPU backend, so this is "the reference's CPU path" that BASELINE.json's
north_star names.  Only tests/, __graft_entry__.smoke() and bench.py may import it.

It follows the same equations and thresholds as the reference kernels
(DGR/cuda_rasterizer/forward.cu:74-151 covariances, :186-283 preprocess, auxiliary.h:40-78,
rasterizer_impl.cu:70-138 keys/ranges, forward.cu:337-410 blend) but in plain fp32 torch ops,
so it is NOT bit-exact (no FMA contraction control) — the bit-exact checker is eogs_oracle.c.
Its value is independence: gradients come from autograd through the forward equations, not
from a restatement of the reference's hand-written backward, so it cross-checks backward.cu's
math as well as ours.  Non-differentiable decisions (tile rects, sort, alpha thresholds,
early termination) are taken on detached values, and min(0.99, .) is straight-through like
the reference backward (backward.cu:575-577,623 ignore the clamp).
"""
from __future__ import annotations

import math

import torch

TILE = 16


def _rot_matrix(q):
    r, x, y, z = q.unbind(-1)
    # glm column-major R[c][k] (forward.cu:135-139) -> tensor [P, c, k]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], -2)
    return R


def geometry(means3D, scales, rotations, opacities, view, W, H, scale_modifier=1.0, antialiasing=False,
             cov3D_precomp=None):
    """Per-Gaussian stage.  Returns a dict; 'p' (NDC u, v, altitude) and 'T' (2x3) are graph
    nodes whose .grad give dL_dmeans2D and sum_p dL_dT after backward."""
    v = view.reshape(-1)
    A_t = torch.stack([v[0:3], v[4:7], v[8:11]], 0)          # rows: x, y, z coefficients -> p = m @ A_t + b
    p = means3D @ A_t + v[12:15]
    p.retain_grad() if p.requires_grad else None
    if cov3D_precomp is not None and cov3D_precomp.numel():
        c = cov3D_precomp
        Sigma = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], -1),
                             torch.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                             torch.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], -2)
    else:
        R = _rot_matrix(rotations)                            # [P, c, k]
        M = R * (scale_modifier * scales)[:, None, :]         # M[c][k] = s_k R[c][k]
        Sigma = M @ M.transpose(1, 2)                         # Sigma[i][j] = sum_k M[i][k] M[j][k]
    hW, hH = W / 2.0, H / 2.0
    T = torch.stack([hW * torch.stack([v[0], v[4], v[8]]), hH * torch.stack([v[1], v[5], v[9]])], 0)  # [2,3]
    if T.requires_grad:
        T.retain_grad()
    cov2 = T @ Sigma @ T.t()                                   # [P,2,2]
    cxx, cxy, cyy = cov2[:, 0, 0], cov2[:, 0, 1], cov2[:, 1, 1]
    det_cov = cxx * cyy - cxy * cxy
    a, c_ = cxx + 0.3, cyy + 0.3
    det = a * c_ - cxy * cxy
    op = opacities.reshape(-1)
    if antialiasing:
        ratio = det_cov / det
        op = op * torch.sqrt(torch.clamp(ratio, min=0.000025))
    ok = det.detach() != 0
    det_safe = torch.where(ok, det, torch.ones_like(det))
    conic = torch.stack([c_ / det_safe, -cxy / det_safe, a / det_safe], -1)
    with torch.no_grad():
        mid = 0.5 * (a + c_)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        radius = torch.ceil(3.0 * torch.sqrt(lam)).to(torch.int32)
    pix = torch.stack([((p[:, 0].double() + 1.0) * W - 1.0) * 0.5, ((p[:, 1].double() + 1.0) * H - 1.0) * 0.5], -1).float()
    with torch.no_grad():
        gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
        rf = radius.float()
        x0 = ((pix[:, 0] - rf) / TILE).trunc().clamp(0, gx).int()
        y0 = ((pix[:, 1] - rf) / TILE).trunc().clamp(0, gy).int()
        x1 = ((pix[:, 0] + rf + TILE - 1) / TILE).trunc().clamp(0, gx).int()
        y1 = ((pix[:, 1] + rf + TILE - 1) / TILE).trunc().clamp(0, gy).int()
        area = (x1 - x0) * (y1 - y0)
        depth = 200.0 - p[:, 2]
        if bool(((depth < 0) & ok & (area > 0)).any()):
            raise RuntimeError("Point is too high: altitude above 200")
        vis = ok & (area > 0)
        radii = torch.where(vis, radius, torch.zeros_like(radius))
    return dict(p=p, T=T, pix=pix, conic=conic, op=op, radii=radii, rect=(x0, y0, x1, y1), depth=depth, vis=vis)


def binning(geo, W, H):
    x0, y0, x1, y1 = geo["rect"]
    vis = geo["vis"]
    gx = (W + TILE - 1) // TILE
    ids = torch.nonzero(vis).flatten()
    w = (x1 - x0)[ids].long()
    h = (y1 - y0)[ids].long()
    cnt = w * h
    rep = torch.repeat_interleave(torch.arange(ids.numel()), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(int(cnt.sum())) - start[rep]
    ty = y0[ids].long()[rep] + local // w[rep]
    tx = x0[ids].long()[rep] + local % w[rep]
    tile = ty * gx + tx
    gid = ids[rep]
    dbits = geo["depth"].detach().float().contiguous().view(torch.int32).long()[gid]
    key = (tile << 32) | dbits
    order = torch.sort(key, stable=True).indices
    return tile[order], gid[order], key[order]


def render(means3D, scales, rotations, opacities, colors, view, bg, W, H, scale_modifier=1.0,
           antialiasing=False, cov3D_precomp=None, tile_window=None):
    """Returns (color[C,H,W], invdepth[1,H,W], radii[P], aux) — differentiable w.r.t. every float input.
    tile_window=(tx0, ty0, tx1, ty1) blends only those tiles (bounded sample for the CPU baseline);
    pixels outside keep the background."""
    geo = geometry(means3D, scales, rotations, opacities, view, W, H, scale_modifier, antialiasing, cov3D_precomp)
    tile_sorted, gid_sorted, keys = binning(geo, W, H)
    C = colors.shape[1]
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    ntiles = gx * gy
    counts = torch.bincount(tile_sorted, minlength=ntiles)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    out = bg[:, None, None].expand(C, H, W).clone() * 1.0
    outd = torch.zeros(1, H, W)
    final_T = torch.ones(H, W)
    n_contrib = torch.zeros(H, W, dtype=torch.int32)
    inv_depth = (1.0 / geo["depth"]).detach()
    pieces = []
    for t in range(ntiles):
        n = int(counts[t])
        ty, tx = divmod(t, gx)
        if tile_window is not None and not (tile_window[0] <= tx < tile_window[2] and tile_window[1] <= ty < tile_window[3]):
            continue
        ys = torch.arange(ty * TILE, min((ty + 1) * TILE, H))
        xs = torch.arange(tx * TILE, min((tx + 1) * TILE, W))
        if n == 0:
            continue
        g = gid_sorted[int(starts[t]):int(ends[t])]
        py, px = torch.meshgrid(ys.float(), xs.float(), indexing="ij")
        px, py = px.reshape(-1, 1), py.reshape(-1, 1)                      # [npix,1]
        dx = geo["pix"][g, 0][None, :] - px
        dy = geo["pix"][g, 1][None, :] - py
        con = geo["conic"][g]
        power = -0.5 * (con[:, 0] * dx * dx + con[:, 2] * dy * dy) - con[:, 1] * dx * dy
        G = torch.exp(power)
        a_raw = geo["op"][g][None, :] * G
        alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()     # straight-through min(0.99, .)
        with torch.no_grad():
            valid = (power <= 0) & (alpha >= 1.0 / 255.0)
            one_m = torch.where(valid, 1 - alpha, torch.ones_like(alpha))
            T_incl = torch.cumprod(one_m, 1)
            stop = valid & (T_incl < 0.0001)
            stopped = torch.cummax(stop.int(), 1).values.bool()
            contrib = valid & ~stopped
            idx = torch.arange(1, n + 1)[None, :].expand_as(contrib)
            last = torch.where(contrib, idx, torch.zeros_like(idx)).max(1).values
        am = torch.where(contrib, alpha, torch.zeros_like(alpha))
        T_in = torch.cumprod(1 - am, 1)
        T_ex = torch.cat([torch.ones_like(T_in[:, :1]), T_in[:, :-1]], 1)
        wgt = am * T_ex                                                    # [npix, n]
        col = wgt @ colors[g]                                              # [npix, C]
        Tf = T_in[:, -1]
        col = col + Tf[:, None] * bg[None, :]
        dep = wgt @ inv_depth[g]
        pieces.append((ys, xs, col, dep, Tf.detach(), last))
    # assemble (index_put keeps autograd)
    for ys, xs, col, dep, Tf, last in pieces:
        hh, ww = ys.numel(), xs.numel()
        out[:, ys[0]:ys[0] + hh, xs[0]:xs[0] + ww] = col.t().reshape(C, hh, ww)
        outd[0, ys[0]:ys[0] + hh, xs[0]:xs[0] + ww] = dep.reshape(hh, ww)
        final_T[ys[0]:ys[0] + hh, xs[0]:xs[0] + ww] = Tf.reshape(hh, ww)
        n_contrib[ys[0]:ys[0] + hh, xs[0]:xs[0] + ww] = last.reshape(hh, ww).int()
    aux = dict(geo=geo, keys=keys, point_list=gid_sorted, final_T=final_T, n_contrib=n_contrib,
               num_rendered=int(gid_sorted.numel()))
    return out, outd, geo["radii"], aux


def render_fwd_bwd(scene_tensors, view, bg, W, H, dL_dcolor, dL_dinvdepth=None, antialiasing=False,
                   tile_window=None):
    """One fwd+bwd 'render' on the CPU (the unit of BASELINE.json's metric).  Returns gradients."""
    means3D, scales, rotations, opacities, colors = [t.detach().clone().requires_grad_(True) for t in scene_tensors]
    v = view.detach().clone().requires_grad_(True)
    out, outd, radii, aux = render(means3D, scales, rotations, opacities, colors, v, bg, W, H,
                                   antialiasing=antialiasing, tile_window=tile_window)
    loss = (out * dL_dcolor).sum()
    if dL_dinvdepth is not None:
        loss = loss + (outd * dL_dinvdepth).sum()
    if loss.requires_grad:        # an empty tile window blends nothing
        loss.backward()
    geo = aux["geo"]
    return dict(color=out.detach(), invdepth=outd.detach(), radii=radii, aux=aux,
                dL_dmeans3D=means3D.grad, dL_dscales=scales.grad, dL_drotations=rotations.grad,
                dL_dopacity=opacities.grad, dL_dcolors=colors.grad, dL_dview=v.grad,
                dL_dmeans2D=geo["p"].grad, dL_dT_sum=geo["T"].grad)

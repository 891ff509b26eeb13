"""GPU: the reference-shaped Python API (GaussianRasterizer + autograd) end to end, and the edge
cases of the boundary: empty scene, nothing visible, precomputed covariances, non-contiguous inputs,
altitude > 200, markVisible, a non-default stream, the single-call C entry point."""
import ctypes as C

import numpy as np
import pytest
import torch

import eogs2_b200 as E
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from eogs2_b200 import _cabi, scene as S
from oracle import c_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def settings(dev, view, bg, H, W, aa=False, debug=False, proj=None):
    return GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg.to(dev), scale_modifier=1.0, viewmatrix=view,
        projmatrix=view if proj is None else proj, sh_degree=0, campos=torch.zeros(3, device=dev),
        prefiltered=False, debug=debug, antialiasing=aa)


def test_autograd_end_to_end_like_renderer_py(cuda_dev):
    """The call pattern of gaussian_renderer/renderer.py:66-122, with viewmatrix requiring grad
    (camera refinement, renderer.py:59-65)."""
    dev = cuda_dev
    P, W, H = 30_000, 320, 256
    sc = S.make_scene(P, "trained", 21)
    view0 = S.make_camera(21)
    colors0 = S.colors_precomp(sc, view0)
    bg = S.background(21)
    dcol, dinv = S.upstream_grads(5, H, W, 21, True)
    leaves = {k: getattr(sc, k).to(dev).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities")}
    colors = colors0.to(dev).requires_grad_(True)
    view = view0.to(dev).requires_grad_(True)
    screenspace_points = torch.zeros_like(leaves["means3D"], requires_grad=True)
    rast = GaussianRasterizer(raster_settings=settings(dev, view, bg, H, W, debug=True))
    rendered, radii, invd = rast(means3D=leaves["means3D"], means2D=screenspace_points, shs=None, colors_precomp=colors,
                                 opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"],
                                 cov3D_precomp=None)
    assert rendered.shape == (5, H, W) and radii.shape == (P,) and radii.dtype == torch.int32 and invd.shape == (1, H, W)
    ((rendered * dcol.to(dev)).sum() + (invd * dinv.to(dev)).sum()).backward()
    torch.cuda.synchronize()

    o = O.forward(sc.means3D.numpy(), sc.scales.numpy(), sc.rotations.numpy(), sc.opacities.numpy(), colors0.numpy(),
                  view0.numpy(), bg.numpy(), W, H)
    g = O.backward(o, dcol.numpy(), dinv.numpy())
    assert np.array_equal(radii.cpu().numpy(), o["radii"])
    assert (np.abs(rendered.detach().cpu().numpy() - o["color"]) > 1e-4).mean() <= 1e-4
    for name, t in (("dL_dmeans3D", leaves["means3D"]), ("dL_dscales", leaves["scales"]),
                    ("dL_drotations", leaves["rotations"]), ("dL_dopacity", leaves["opacities"]),
                    ("dL_dcolors", colors), ("dL_dmeans2D", screenspace_points)):
        assert t.grad is not None and t.grad.shape == t.shape, name
        assert rel(t.grad.cpu().numpy(), g[name]) < 1e-3, name
    assert view.grad.shape == (4, 4)
    assert rel(view.grad.cpu().numpy(), g["grad_viewmatrix"]) < 1e-3
    assert float(screenspace_points.grad[:, 2].abs().max()) == 0.0


def test_empty_scene_returns_zero_image_not_bg(cuda_dev):
    """P == 0: nothing is launched and out_color stays 0, not bg (rasterize_points.cu:88)."""
    dev = cuda_dev
    view = S.make_camera(1).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3, -30.0, 0.0])
    rast = GaussianRasterizer(settings(dev, view, bg, 40, 56))
    z = lambda *s: torch.zeros(*s, device=dev)
    color, radii, invd = rast(z(0, 3), z(0, 3), z(0, 1), colors_precomp=z(0, 5), scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (5, 40, 56) and float(color.abs().max()) == 0.0 and radii.numel() == 0


def test_nothing_visible_renders_background(cuda_dev):
    dev = cuda_dev
    view = S.make_camera(1).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3, -30.0, 0.0])
    rast = GaussianRasterizer(settings(dev, view, bg, 33, 47))
    m = torch.tensor([[40.0, 40.0, 0.0], [-50.0, 3.0, 0.1]], device=dev, requires_grad=True)
    color, radii, invd = rast(m, torch.zeros_like(m), torch.ones(2, 1, device=dev), colors_precomp=torch.ones(2, 5, device=dev),
                              scales=torch.full((2, 3), 0.01, device=dev), rotations=torch.tensor([[1.0, 0, 0, 0]] * 2, device=dev))
    assert int(radii.sum()) == 0
    assert torch.allclose(color, bg.to(dev)[:, None, None].expand_as(color))
    color.sum().backward()
    assert float(m.grad.abs().max()) == 0.0


def test_cov3D_precomp_path_matches_scale_rotation_path(cuda_dev):
    dev = cuda_dev
    P, W, H = 5000, 128, 96
    sc = S.make_scene(P, "trained", 31).to(dev)
    view = S.make_camera(31).to(dev)
    colors = S.colors_precomp(sc, view)
    bg = S.background(31)
    o = O.forward(sc.means3D.cpu().numpy(), sc.scales.cpu().numpy(), sc.rotations.cpu().numpy(), sc.opacities.cpu().numpy(),
                  colors.cpu().numpy(), view.cpu().numpy(), bg.numpy(), W, H)
    cov = torch.from_numpy(o["cov3D"]).to(dev).requires_grad_(True)
    rast = GaussianRasterizer(settings(dev, view, bg, H, W))
    a, ra, _ = rast(sc.means3D, torch.zeros_like(sc.means3D), sc.opacities, colors_precomp=colors, scales=sc.scales,
                    rotations=sc.rotations)
    b, rb, _ = rast(sc.means3D, torch.zeros_like(sc.means3D), sc.opacities, colors_precomp=colors, cov3D_precomp=cov)
    assert torch.equal(ra, rb) and torch.equal(a, b)
    b.sum().backward()
    assert cov.grad.shape == (P, 6) and float(cov.grad.abs().max()) > 0


def test_non_contiguous_and_offset_inputs(cuda_dev):
    """Inputs are made contiguous by the binding (rasterize_points.cu:101-120); a storage offset breaks
    16-byte alignment and must take the scalar staging path with identical results."""
    dev = cuda_dev
    P, W, H = 3000, 96, 96
    sc = S.make_scene(P + 1, "trained", 41).to(dev)
    view = S.make_camera(41).to(dev)
    colors = S.colors_precomp(sc, view)
    bg = S.background(41)
    rast = GaussianRasterizer(settings(dev, view, bg, H, W))
    base = rast(sc.means3D[1:].clone(), torch.zeros(P, 3, device=dev), sc.opacities[1:].clone(), colors_precomp=colors[1:].clone(),
                scales=sc.scales[1:].clone(), rotations=sc.rotations[1:].clone())
    off = rast(sc.means3D[1:], torch.zeros(P, 3, device=dev), sc.opacities[1:], colors_precomp=colors[1:],
               scales=sc.scales[1:], rotations=sc.rotations[1:].clone())
    wide = torch.cat([sc.means3D, sc.means3D], 1)[1:, :3]                 # non-contiguous view
    nc = rast(wide, torch.zeros(P, 3, device=dev), sc.opacities[1:], colors_precomp=colors[1:].clone(),
              scales=sc.scales[1:].clone(), rotations=sc.rotations[1:].clone())
    assert torch.equal(base[0], off[0]) and torch.equal(base[1], off[1]) and torch.equal(base[0], nc[0])


def test_altitude_above_200_raises_instead_of_trapping(cuda_dev):
    dev = cuda_dev
    view = S.make_camera(1).to(dev)
    rast = GaussianRasterizer(settings(dev, view, S.background(1), 32, 32))
    m = torch.tensor([[0.0, 0.0, 0.9]], device=dev)                       # 270 m
    with pytest.raises(RuntimeError, match="too high"):
        rast(m, torch.zeros_like(m), torch.ones(1, 1, device=dev), colors_precomp=torch.ones(1, 5, device=dev),
             scales=torch.full((1, 3), 0.01, device=dev), rotations=torch.tensor([[1.0, 0, 0, 0]], device=dev))
    # the context survives (the reference's __trap would have killed it)
    assert float(torch.ones(4, device=dev).sum()) == 4.0


def test_mark_visible_all_true(cuda_dev):
    dev = cuda_dev
    view = S.make_camera(1).to(dev)
    rast = GaussianRasterizer(settings(dev, view, S.background(1), 32, 32))
    vis = rast.markVisible(torch.randn(1000, 3, device=dev))
    assert vis.dtype == torch.bool and vis.shape == (1000,) and bool(vis.all())


def test_runs_on_the_current_stream(cuda_dev):
    dev = cuda_dev
    P, W, H = 4000, 128, 128
    sc = S.make_scene(P, "trained", 51).to(dev)
    view = S.make_camera(51).to(dev)
    colors = S.colors_precomp(sc, view)
    rast = GaussianRasterizer(settings(dev, view, S.background(51), H, W))
    args = dict(means3D=sc.means3D, means2D=torch.zeros_like(sc.means3D), opacities=sc.opacities, colors_precomp=colors,
                scales=sc.scales, rotations=sc.rotations)
    a = rast(**args)[0]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        b = rast(**args)[0]
    side.synchronize()
    assert torch.equal(a, b)


def test_single_call_c_entry_point_with_alloc_callbacks(cuda_dev):
    """eogs_rasterize_forward: the Rasterizer::forward-shaped entry point with allocation callbacks."""
    dev = cuda_dev
    lib = _cabi.load()
    P, W, H = 6000, 144, 112
    sc = S.make_scene(P, "trained", 61).to(dev)
    view = S.make_camera(61).to(dev)
    colors = S.colors_precomp(sc, view)
    bg = S.background(61).to(dev)
    st = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0,
                                 torch.empty(0, device=dev), view, H, W)
    keep = {}

    def alloc(user, which, n):
        t = torch.empty(max(int(n), 1), dtype=torch.uint8, device=dev)
        keep[which] = t
        return t.data_ptr()
    cb = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)(alloc)
    color = torch.empty(5, H, W, device=dev); invd = torch.empty(1, H, W, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    geom, plist, image, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
    rc = lib.eogs_rasterize_forward(
        torch.cuda.current_stream(dev).cuda_stream, P, W, H, 5, sc.means3D.data_ptr(), sc.scales.data_ptr(),
        sc.rotations.data_ptr(), None, sc.opacities.data_ptr(), colors.data_ptr(), view.data_ptr(), 1.0, 0,
        bg.data_ptr(), C.cast(cb, C.c_void_p), None, radii.data_ptr(), color.data_ptr(), invd.data_ptr(),
        C.byref(geom), C.byref(plist), C.byref(image), C.byref(n))
    _cabi.check(rc, "eogs_rasterize_forward")
    torch.cuda.synchronize()
    assert n.value == st.num_rendered and torch.equal(color, st.color) and torch.equal(radii, st.radii)


def test_backward_writes_into_caller_buffers(cuda_dev):
    """out=: the backward kernels write the gradients straight into views of one flat bucket
    (the data-parallel all-reduce buffer of bench.py / dp.py) — same values as fresh tensors."""
    dev = cuda_dev
    P, W, H = 5000, 128, 96
    sc = S.make_scene(P, "trained", 71).to(dev)
    view = S.make_camera(71).to(dev)
    colors = S.colors_precomp(sc, view)
    bg = S.background(71).to(dev)
    dcol, dinv = (t.to(dev) for t in S.upstream_grads(5, H, W, 71, True))
    empty = torch.empty(0, device=dev)
    st = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, H, W)
    args = (st, bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, view, dcol, dinv)
    ref = E.rasterize_backward_raw(*args)
    bucket = torch.full((16 * P + 16,), float("nan"), device=dev)
    views = {"means3D": bucket[0:3 * P], "colors": bucket[3 * P:8 * P], "opacity": bucket[8 * P:9 * P],
             "scales": bucket[9 * P:12 * P], "rotations": bucket[12 * P:16 * P], "cam_sums": bucket[16 * P:]}
    got = E.rasterize_backward_raw(*args, out=views)
    assert got[3].data_ptr() == bucket.data_ptr() and got[7].data_ptr() == bucket[16 * P:].data_ptr()
    assert not torch.isnan(bucket).any()
    for a, b in zip(got, ref):
        if a is not None:
            assert rel(a.cpu().numpy(), b.cpu().numpy()) < 1e-4          # float atomics: order differs run to run
    with pytest.raises(_cabi.EogsRasterError):
        E.rasterize_backward_raw(*args, out={"means3D": torch.empty(7, device=dev)})


def test_three_channel_specialisation_matches_five_channel_render(cuda_dev):
    """C = 3 (grey/PAN, altitude, opacity — SURVEY.md §8 a19) behind the same API: channels are
    blended independently, so the C = 3 render must equal the first three channels of the C = 5
    render bit for bit, and with no upstream gradient on channels 3, 4 the geometry gradients agree."""
    dev = cuda_dev
    P, W, H = 8000, 200, 136
    sc = S.make_scene(P, "trained", 81).to(dev)
    view = S.make_camera(81).to(dev)
    c5 = S.colors_precomp(sc, view)
    c3 = torch.stack([c5[:, :3].mean(1), c5[:, 3], c5[:, 4]], 1).contiguous()         # grey, altitude, 1
    c5b = torch.cat([c3, c5[:, :2]], 1).contiguous()
    bg5 = S.background(81).to(dev)
    bg3 = bg5[[0, 3, 4]].contiguous()
    bg5b = torch.cat([bg3, bg5[:2]]).contiguous()
    dcol, dinv = (t.to(dev) for t in S.upstream_grads(5, H, W, 81, True))
    dcol5 = dcol.clone(); dcol5[3:] = 0
    dcol3 = dcol[:3].contiguous()
    empty = torch.empty(0, device=dev)

    def run(colors, bg, dc):
        st = E.rasterize_forward_raw(bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty, view, H, W)
        g = E.rasterize_backward_raw(st, bg, sc.means3D, colors, sc.opacities, sc.scales, sc.rotations, 1.0, empty,
                                     view, view, dc, dinv)
        return st, g
    st3, g3 = run(c3, bg3, dcol3)
    st5, g5 = run(c5b, bg5b, dcol5)
    assert st3.color.shape == (3, H, W)
    assert torch.equal(st3.color, st5.color[:3]) and torch.equal(st3.invdepth, st5.invdepth)
    assert torch.equal(st3.radii, st5.radii) and st3.num_rendered == st5.num_rendered
    assert torch.equal(E.export_state(st3)["n_contrib"], E.export_state(st5)["n_contrib"])
    assert rel(g3[1].cpu().numpy(), g5[1][:, :3].cpu().numpy()) < 1e-4                 # dL_dcolors
    for k in (0, 2, 3, 5, 6, 7):                                                       # means2D, opacity, means3D, scales, rotations, cam sums
        assert rel(g3[k].cpu().numpy(), g5[k].cpu().numpy()) < 1e-4, k

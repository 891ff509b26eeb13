"""Golden vectors for SURVEY.md section 8 rows a19 / N1: the REFERENCE's own callers of the rasterizer,
executed here (CPU, build container) from /root/reference —

    gaussian_renderer/renderer.py:27-144            render()
    gaussian_renderer/renderer_cc_shadow.py:6-54    render_resample_virtual_camera()
    scene/cameras/affine_cameras.py:350-370,303-348 AffineCamera.get_sun_camera(), render_pipeline()
    scene/gaussian_model.py:41-53,109-137           GaussianModel activations
    train_pan.py:272-329                            how one camera's renders are composed

— on top of the CPU stand-in rasterizer of tests/golden/ref_import.py (oracle/eogs_oracle.c behind the
reference's GaussianRasterizer surface).  The GPU tests (tests/test_reference_callers_gpu.py) feed the same
raw GaussianModel parameters and cameras to (a) the torch formulation of those callers on the drop-in
`diff_gaussian_rasterization` package and (b) the fused kernels (eogs2_b200/fused.py, shadow.py), and compare
with what the reference's code produced here.

    python tests/golden/make_golden_renderer.py          # writes tests/golden/renderer_ref.npz
"""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import ref_import                                   # noqa: E402
from eogs2_b200 import scene as S                   # noqa: E402  (synthetic scene generator; plain torch on CPU)

C0 = 0.28209479177387814                            # utils/sh_utils.py
P, W, H, SEED = 800, 64, 48, 11
SUN_Q = (-0.0030, -0.0025)                          # SURVEY.md section 8d sun shear (NDC per metre of altitude)
OUT = HERE / "renderer_ref.npz"


def raw_parameters(seed=SEED, n=P):
    sc = S.make_scene(n, "trained", seed)
    g = torch.Generator().manual_seed(seed + 1)
    return dict(xyz=sc.means3D.clone(), f_dc=((sc.rgb - 0.5) / C0).unsqueeze(1).clone(),
                opacity=torch.logit(sc.opacities).reshape(-1, 1).clone(), scaling=torch.log(sc.scales).clone(),
                rotation=(sc.rotations * (0.5 + torch.rand(n, 1, generator=g))).clone())


def camera_arrays(seed=SEED):
    view = S.make_camera(seed)                      # transposed [[A, b], [0, 1]]
    A, b = view[:3, :3].t().contiguous(), view[3, :3].clone()
    c2s = torch.eye(3)
    c2s[0, 2], c2s[1, 2] = SUN_Q
    return dict(affine_coef=A.numpy(), affine_inter=b.numpy(), camera_to_sun=c2s.numpy(),
                sun_affine_coef=(c2s @ A).numpy(), sun_affine_inter=(c2s @ b).numpy())


def make_model(R, raw):
    pc = R.GaussianModel(0)
    pc._xyz = torch.nn.Parameter(raw["xyz"].clone())
    pc._features_dc = torch.nn.Parameter(raw["f_dc"].clone())
    pc._features_rest = torch.nn.Parameter(torch.zeros(raw["xyz"].shape[0], 0, 3))
    pc._opacity = torch.nn.Parameter(raw["opacity"].clone())
    pc._scaling = torch.nn.Parameter(raw["scaling"].clone())
    pc._rotation = torch.nn.Parameter(raw["rotation"].clone())
    return pc


def make_camera(R, cam, learn_last=False, last_row=None):
    caminfo = SimpleNamespace(
        image_name="synthetic", reference_altitude=np.zeros(1, np.float32), min_world=np.full(3, -1, np.float32),
        max_world=np.ones(3, np.float32), width=W, height=H, centerofscene_ECEF=np.zeros(3, np.float32),
        affine_coef=cam["affine_coef"], affine_inter=cam["affine_inter"], load_sun=True,
        sun_affine_coef=cam["sun_affine_coef"], sun_affine_inter=cam["sun_affine_inter"],
        camera_to_sun=cam["camera_to_sun"], altitude_bounds=np.array([-30.0, 75.0], np.float32),
        learn_wv_transform=learn_last, use_cc=True, use_shadow=True)
    args = SimpleNamespace(camera_params=SimpleNamespace(learn_wv_only_lastparam=True, use_exposure=False),
                           transient_params=SimpleNamespace(use_transient=False))
    c = R.AffineCamera(caminfo, torch.zeros(3, H, W), None, data_device="cpu", args=args)
    if learn_last:
        with torch.no_grad():
            c.last_row.copy_(torch.as_tensor(last_row))
        c.last_row.requires_grad_(True)
    return c


def grads_of(pc):
    return {n: getattr(pc, "_" + a).grad.detach().numpy().copy()
            for n, a in (("xyz", "xyz"), ("f_dc", "features_dc"), ("opacity", "opacity"), ("scaling", "scaling"),
                         ("rotation", "rotation"))}


def generate() -> dict:
    R = ref_import.load()
    raw, cam = raw_parameters(), camera_arrays()
    bg = S.background(SEED).clone()
    bg[3], bg[4] = -30.0, 0.0                                        # train_pan.py:272-277
    dcol = S.upstream_grads(5, H, W, SEED, False)[0]
    out = {"P": P, "W": W, "H": H, "bg": bg.numpy(), "dcol": dcol.numpy(), **{"raw_" + k: v.numpy() for k, v in raw.items()},
           **{"cam_" + k: v for k, v in cam.items()}}

    def render_case(tag, aa=False, mod=1.0, cov3d=False, learn_last=False, last_row=(0.01, -0.02, 0.5, 0.0)):
        pc = make_model(R, raw)
        c = make_camera(R, cam, learn_last, last_row)
        pipe = SimpleNamespace(debug=False, antialiasing=aa, compute_cov3D_python=cov3d, require_radii=True)
        pkg = R.render(c, pc, pipe, bg, scaling_modifier=mod)
        (pkg["render"] * dcol).sum().backward()
        out[f"{tag}_render"] = pkg["render"].detach().numpy()
        out[f"{tag}_radii"] = pkg["radii"].numpy()
        out[f"{tag}_visibility_filter"] = pkg["visibility_filter"].numpy()
        out[f"{tag}_grad_viewspace"] = pkg["viewspace_points"].grad.numpy()
        for k, v in grads_of(pc).items():
            out[f"{tag}_grad_{k}"] = v
        if learn_last:
            out[f"{tag}_last_row"] = np.asarray(last_row, np.float32)
            out[f"{tag}_grad_last_row"] = c.last_row.grad.numpy()

    render_case("main")
    render_case("aa_lastrow", aa=True, learn_last=True)
    render_case("mod", mod=0.8)
    render_case("cov3d", cov3d=True)

    # ---- one camera of train_pan.py:272-329: main render -> sun render, resampled -> shadow -> shaded image
    pc = make_model(R, raw)
    c = make_camera(R, cam)
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, require_radii=True)
    pkg = R.render(c, pc, pipe, bg)
    raw_render, altitude_render = pkg["render"][:3], pkg["render"][3]
    rendered_uva = torch.stack(c.UV_grid + (altitude_render,), dim=-1)
    sun_camera, camera_to_sun = c.get_sun_camera()
    sun_rgb, sun_alt, sun_uv, sun_render = R.render_resample_virtual_camera(
        virtual_camera=sun_camera, cam2virt=camera_to_sun, rendered_uva=rendered_uva, gaussians=pc, pipe=pipe,
        background=bg, return_extra=True)
    sun_altitude_diff = altitude_render - sun_alt
    output = c.render_pipeline(raw_render=raw_render, sun_altitude_diff=sun_altitude_diff)
    g = torch.Generator().manual_seed(SEED + 7)
    d_shaded = torch.randn(3, H, W, generator=g) / (W * H)
    d_sun_rgb = torch.randn(3, H, W, generator=g) / (W * H)
    loss = (output["shaded"] * d_shaded).sum() + (sun_rgb * d_sun_rgb).sum()
    loss.backward()
    out.update(shadow_sun_view=sun_camera.world_view_transform.detach().numpy(), shadow_cam2virt=camera_to_sun.numpy(),
               shadow_sun_W=sun_camera.image_width, shadow_sun_H=sun_camera.image_height,
               shadow_sun_render=sun_render.detach().numpy(), shadow_sun_rgb=sun_rgb.detach().numpy(),
               shadow_sun_alt=sun_alt.detach().numpy(), shadow_sun_uv=sun_uv.detach().numpy(),
               shadow_shadowmap=output["shadowmap"].detach().numpy(), shadow_shaded=output["shaded"].detach().numpy(),
               shadow_d_shaded=d_shaded.numpy(), shadow_d_sun_rgb=d_sun_rgb.numpy(), shadow_loss=float(loss.detach()),
               shadow_inshadow_cc=c.inshadow_color_correction.detach().numpy())
    for k, v in grads_of(pc).items():
        out[f"shadow_grad_{k}"] = v
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    data = generate()
    np.savez_compressed(OUT, **data)
    print(OUT, OUT.stat().st_size, "bytes;", len(data), "arrays")

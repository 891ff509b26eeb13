"""Deterministic synthetic scenes and affine cameras of the EOGS++ shapes.

Used by tests, bench.py and __graft_entry__.smoke(); there is no dataset on the GPU box.
Everything is generated on the CPU from a seeded torch.Generator and then moved, so the same
seed gives the same bits everywhere.

Conventions (reference: scene/cameras/affine_cameras.py:151-157,186-188):
  world   : normalised UTM cube, x,y in [-0.7, 0.7], z in [-0.10, 0.25]; 300 m per unit
  camera  : [u, v, altitude] = A @ xyz + b with u, v in NDC [-1, 1] and altitude in metres
  viewmatrix = projmatrix = [[A, b], [0, 1]]^T   (4x4, TRANSPOSED, flat index 4*col + row)
  colours : colors_precomp = [r, g, b, altitude, 1]  (gaussian_renderer/renderer.py:99-105)
  bg      : [rand, rand, rand, altitude_min, 0]      (train_pan.py:272-277)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

METRES_PER_UNIT = 300.0
ALTITUDE_MIN = -30.0
BOX_XY = 0.70
BOX_Z = (-0.10, 0.25)


@dataclass
class Scene:
    means3D: torch.Tensor      # [P,3]
    scales: torch.Tensor       # [P,3] (post-exp)
    rotations: torch.Tensor    # [P,4] (post-normalise)
    opacities: torch.Tensor    # [P,1] (post-sigmoid)
    rgb: torch.Tensor          # [P,3]

    def to(self, device):
        return Scene(*(t.to(device) for t in (self.means3D, self.scales, self.rotations, self.opacities, self.rgb)))

    @property
    def P(self):
        return self.means3D.shape[0]


def make_scene(P: int, kind: str = "trained", seed: int = 1337) -> Scene:
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    xy = (u(P, 2) * 2 - 1) * BOX_XY
    z = BOX_Z[0] + u(P, 1) * (BOX_Z[1] - BOX_Z[0])
    means = torch.cat([xy, z], 1)
    if kind == "init":
        # dataset_affine.py:247-295 + gaussian_model.py:179-186: isotropic, sqrt(distCUDA2)-like
        vol = (2 * BOX_XY) ** 2 * (BOX_Z[1] - BOX_Z[0])
        s = 0.55 * (vol / max(P, 1)) ** (1.0 / 3.0) * torch.exp(n(P, 1) * 0.1)
        scales = s.repeat(1, 3)
        rot = torch.zeros(P, 4); rot[:, 0] = 1.0
        opac = torch.full((P, 1), 0.01)
        rgb = torch.full((P, 3), 1.1)
    elif kind == "trained":
        logs = math.log(0.004) + n(P, 3) * 0.6
        scales = torch.exp(logs)
        scales[:, 2] *= 0.3                       # flat roofs / ground
        rot = n(P, 4)
        rot = rot / rot.norm(dim=1, keepdim=True)
        opac = torch.sigmoid(n(P, 1) * 2.0).clamp(0.005, 0.99)
        rgb = u(P, 3)
    else:
        raise ValueError(kind)
    return Scene(means.float().contiguous(), scales.float().contiguous(), rot.float().contiguous(),
                 opac.float().contiguous(), rgb.float().contiguous())


def affine_to_viewmatrix(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    M = torch.eye(4)
    M[:3, :3] = A
    M[:3, 3] = b
    return M.t().contiguous().float()      # BEWARE OF THE TRANSPOSE (affine_cameras.py:155-157)


def make_camera(seed: int = 1337) -> torch.Tensor:
    """RPC-fit-like affine camera: rotation theta about nadir, off-nadir parallax k_u, k_v."""
    g = torch.Generator().manual_seed(seed)
    a = 1.0 / 0.72
    theta = (torch.rand(1, generator=g).item() * 2 - 1) * math.pi
    ku, kv = ((torch.rand(2, generator=g) - 0.5)).tolist()
    A = torch.tensor([[a * math.cos(theta), -a * math.sin(theta), ku],
                      [a * math.sin(theta), a * math.cos(theta), kv],
                      [0.0, 0.0, METRES_PER_UNIT]])
    return affine_to_viewmatrix(A, torch.zeros(3))


def sun_camera(viewmatrix: torch.Tensor, q=(-0.0030, -0.0025), f: int = 2) -> torch.Tensor:
    """Sun-view camera at f x resolution: sun_affine @ diag(1/f, 1/f, 1, 1) (affine_cameras.py:350-370).
    sun_affine = camera_to_sun @ [A|b] with a shear of q NDC per metre of altitude."""
    M = viewmatrix.t().clone()                       # [[A,b],[0,1]]
    c2s = torch.eye(4)
    c2s[0, 2], c2s[1, 2] = q
    sun = (c2s @ M).t()                              # stored transposed
    scaling = torch.eye(4)
    scaling[0, 0] = scaling[1, 1] = 1.0 / f
    return (sun @ scaling).contiguous().float()


def random_camera(viewmatrix: torch.Tensor, extent: float = 0.01, seed: int = 0) -> torch.Tensor:
    """affine_cameras.py:403-430 with centerofscene = 0: shear u, v by N(0,1).clip(-1,1)*extent per metre."""
    g = torch.Generator().manual_seed(seed)
    M = viewmatrix.t().clone()
    myM = torch.eye(4)
    myM[:2, 2] += torch.randn(2, generator=g).clip(-1, 1) * extent
    return (myM @ M).t().contiguous().float()


def colors_precomp(scene: Scene, viewmatrix: torch.Tensor) -> torch.Tensor:
    """[rgb, altitude, 1] (renderer.py:99-105); altitude = third row of the affine map."""
    A = viewmatrix.t()[:3, :3].to(scene.means3D.device)
    b = viewmatrix.t()[:3, 3].to(scene.means3D.device)
    alt = scene.means3D @ A[2] + b[2]
    return torch.cat([scene.rgb, alt[:, None], torch.ones_like(alt[:, None])], 1).contiguous()


def background(seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + 7919)
    bg = torch.rand(5, generator=g)
    bg[3] = ALTITUDE_MIN
    bg[4] = 0.0
    return bg.float()


def upstream_grads(C: int, H: int, W: int, seed: int = 0, with_invdepth: bool = False):
    g = torch.Generator().manual_seed(seed + 104729)
    dcol = torch.randn(C, H, W, generator=g) / (W * H)
    dinv = torch.randn(1, H, W, generator=g) / (W * H) if with_invdepth else torch.zeros(1, H, W)
    return dcol.float(), dinv.float()

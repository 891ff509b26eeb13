"""GPU: alpha_cut (csrc/geom_math.cuh) — the per-Gaussian exponent threshold from which the blend backward takes
the forward's accept decision !(opacity * expf(power) < 1/255) (forward.cu:367-372).  Checked by definition
(accepted at the cut, rejected one float below), on random and adversarial opacities, and end to end: gradients
of scenes whose opacities sit right at 1/255 against the compiled reference."""
import ctypes as C

import numpy as np
import pytest
import torch

from eogs2_b200 import _cabi

pytestmark = pytest.mark.gpu


def alpha_cut(op: torch.Tensor):
    lib = _cabi.load()
    cut = torch.empty_like(op)
    flags = torch.empty(op.numel(), dtype=torch.int32, device=op.device)
    _cabi.check(lib.eogs_debug_alpha_cut(C.c_void_p(torch.cuda.current_stream().cuda_stream), op.numel(),
                                         C.c_void_p(op.data_ptr()), C.c_void_p(cut.data_ptr()),
                                         C.c_void_p(flags.data_ptr())), "eogs_debug_alpha_cut")
    return cut, flags


def test_cut_is_the_exact_boundary(cuda_dev):
    thr = np.float32(1.0) / np.float32(255.0)
    near = [np.nextafter(thr, np.float32(1), dtype=np.float32)]
    for _ in range(40):
        near.append(np.nextafter(near[-1], np.float32(1), dtype=np.float32))
    g = torch.Generator().manual_seed(0)
    ops = torch.cat([torch.tensor(np.array(near + [thr], np.float32)),
                     torch.rand(200_000, generator=g) * 0.99 + 1e-4,                       # the whole opacity range
                     float(thr) * (1 + torch.rand(50_000, generator=g) * 0.1),             # barely above 1/255: expf is flat
                     torch.tensor([0.99, 1.0, 0.5, 0.0039215689, 1e-30, 0.0])]).to(cuda_dev)
    cut, flags = alpha_cut(ops)
    accepted_somewhere = ops >= float(thr)                                                 # power = 0 gives alpha = opacity
    assert torch.isinf(cut[~accepted_somewhere]).all() and (cut[~accepted_somewhere] > 0).all()
    fin = cut[accepted_somewhere]
    assert torch.isfinite(fin).all() and (fin <= 0).all()
    assert (flags[accepted_somewhere] == 1).all(), "cut must be accepted and the next float below it rejected"
    # and it tracks -log(255 opacity) to a few ulps of the exponential
    est = -torch.log(255.0 * ops[accepted_somewhere].double())
    assert float((fin.double() - est).abs().max()) < 1e-6
    nan_cut, _ = alpha_cut(torch.tensor([float("nan")], device=cuda_dev))
    assert float(nan_cut[0]) == float("-inf")


def test_gradients_with_opacities_at_the_threshold(cuda_dev):
    """Every Gaussian's opacity within 6 % above 1/255 (where a linear walk from -log(255 op) would not reach the
    boundary): gradients still match the compiled reference."""
    import test_parity_gpu as T
    from oracle import ref_rasterizer as R
    if not R.available():
        pytest.skip("oracle/_ref/libeogs_ref.so did not travel")
    c = T.make_case(3000, 300, 200, "init", 77)
    g = torch.Generator().manual_seed(1)
    c["opacities"] = (1.0 / 255.0) * (1 + 0.06 * torch.rand(c["opacities"].shape, generator=g))
    st, ex, mine = T.run_mine(cuda_dev, c)
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in c.items()}
    empty, campos = torch.empty(0, device=cuda_dev), torch.zeros(3, device=cuda_dev)
    rs = R.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                   d["view"], d["view"], 1.0, 1.0, c["H"], c["W"], campos, False, False)
    gr = R.backward(rs, d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], 1.0, empty,
                    d["view"], d["view"], 1.0, 1.0, d["dL_dcolor"], d["dL_dinvdepth"], campos, False)
    assert torch.equal(st.color.view(torch.int32), rs.color.view(torch.int32))
    assert int(ex["n_contrib"].max()) > 0
    for nm, t in zip(["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", None, "dL_dscales"], mine):
        if nm is not None:
            assert T.rel(t.cpu().numpy(), gr[nm].cpu().numpy()) < 1e-4, nm

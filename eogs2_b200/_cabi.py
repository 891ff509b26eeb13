"""ctypes binding of libeogs_raster.so (the C ABI declared in include/eogs_raster.h).

There is no fallback: if the library is missing or a symbol is absent, importing the
product path raises.  The library is built in-tree by `python -m eogs2_b200.build`
(or `__graft_entry__.build()`); it is never built implicitly at import time on a GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# EOGS_RASTER_LIB selects another build of the SAME library (developer A/B runs, eogs2_b200/build.py)
LIB_PATH = Path(os.environ.get("EOGS_RASTER_LIB") or Path(__file__).resolve().parent / "libeogs_raster.so")

c_f32p = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())
c_ptr = C.c_void_p

# name -> (restype, argtypes); mirrors include/eogs_raster.h one to one
SIGNATURES = {
    "eogs_abi_version": (C.c_int, []),
    "eogs_last_error": (C.c_char_p, []),
    "eogs_geom_bytes": (C.c_size_t, [C.c_int]),
    "eogs_image_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "eogs_binning_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_uint32]),
    "eogs_grad_scratch_floats": (C.c_size_t, [C.c_int]),
    "eogs_point_list_words": (C.c_size_t, [C.c_uint32]),
    "eogs_forward_geometry": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int,                 # stream, P, W, H, channels
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,            # means3D, scales, rotations, cov3D, opacities, colors
        c_f32p, C.c_float, C.c_int,                                # viewmatrix, scale_modifier, antialiasing
        c_ptr, c_ptr, c_ptr, c_ptr]),                              # radii, geom, info_dev, info_host
    "eogs_forward_render": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,     # stream, P, W, H, channels, I
        c_ptr, c_ptr, c_ptr, c_ptr, c_f32p, c_f32p, c_f32p]),      # geom, point_list, binning, image, bg, out_color, out_invdepth
    "eogs_rasterize_forward": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int,
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
        c_f32p, C.c_float, C.c_int, c_f32p, c_ptr, c_ptr,
        c_ptr, c_f32p, c_f32p, c_ptr, c_ptr, c_ptr, c_ptr]),
    "eogs_backward": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,     # stream, P, W, H, channels, I
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,            # means3D, scales, rotations, cov3D, opacities, colors
        c_f32p, c_f32p, C.c_float, C.c_int, c_f32p,                # view, proj, scale_modifier, antialiasing, bg
        c_ptr, c_ptr, c_ptr, c_ptr, c_f32p, c_f32p,                # radii, geom, point_list, image, dL_dpix, dL_dinvdepth
        c_f32p,                                                    # grad_scratch
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_image_bytes_band": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "eogs_forward_geometry_band": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,   # stream, P, W, H, channels, row_begin, row_end
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
        c_f32p, C.c_float, C.c_int,
        c_ptr, c_ptr, c_ptr, c_ptr]),
    "eogs_forward_render_band": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
        c_ptr, c_ptr, c_ptr, c_ptr, c_f32p, c_f32p, c_f32p]),
    "eogs_backward_band": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
        c_f32p, c_f32p, C.c_float, C.c_int, c_f32p,
        c_ptr, c_ptr, c_ptr, c_ptr, c_f32p, c_f32p,
        c_f32p,
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_export_state_band": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, c_ptr, c_ptr, c_ptr,
        c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "eogs_forward_geometry_params_band": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,            # stream, P, W, H, row_begin, row_end
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,                # xyz, log_scales, raw_rot, logits, f_dc, alt_affine
        c_f32p, C.c_float, C.c_int,                                    # viewmatrix, scale_modifier, antialiasing
        c_ptr, c_ptr, c_ptr, c_ptr]),                                  # radii, geom, info_dev, info_host
    "eogs_backward_params_band": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,                        # xyz, log_scales, raw_rot, logits, alt_affine
        c_f32p, c_f32p, C.c_float, C.c_int, c_f32p,                    # view, proj, scale_modifier, antialiasing, bg
        c_ptr, c_ptr, c_ptr, c_ptr, c_f32p, c_f32p,                    # radii, geom, point_list, image, dL_dpix, dL_dinvdepth
        c_f32p,                                                        # grad_scratch
        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_resample_forward": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_resample_backward": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_photometric_forward": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                           c_f32p, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p]),
    "eogs_photometric_backward": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                            c_f32p, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p]),
    "eogs_adam_step": (C.c_int, [c_ptr, C.c_ulonglong, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_float),
                                 C.c_double, C.c_double, C.c_double, C.c_int, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_prune_temp_bytes": (C.c_size_t, [C.c_int]),
    "eogs_prune_offsets": (C.c_int, [c_ptr, C.c_int, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "eogs_prune_gather": (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr, c_f32p, c_f32p]),
    "eogs_densify_select": (C.c_int, [c_ptr, C.c_int, c_f32p, c_f32p, c_f32p, C.c_float, C.c_float, c_ptr, c_ptr]),
    "eogs_densify_split_children": (C.c_int, [c_ptr, C.c_int, C.c_int, c_f32p, c_f32p, c_f32p, c_f32p]),
    "eogs_densify_keep": (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_f32p, c_f32p, C.c_float, C.c_float, c_ptr]),
    "eogs_knn_bytes": (C.c_size_t, [C.c_int]),
    "eogs_knn_dist2": (C.c_int, [c_ptr, C.c_int, c_f32p, c_ptr, C.c_size_t, c_f32p]),
    "eogs_dsm_splat": (C.c_int, [c_ptr, C.c_longlong, c_ptr, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_int, C.c_float, c_ptr, c_f32p]),
    "eogs_nvls_allreduce": (C.c_int, [c_ptr, c_ptr, C.c_ulonglong, C.c_int, C.c_int]),
    "eogs_p2p_allreduce": (C.c_int, [c_ptr, c_ptr, C.c_ulonglong, C.c_int, C.c_int]),
    "eogs_mark_visible": (C.c_int, [c_ptr, C.c_int, c_f32p, c_f32p, c_f32p, c_ptr]),
    "eogs_profile_enable": (C.c_int, [C.c_int]),
    "eogs_profile_read": (C.c_int, [c_ptr, C.c_int]),
    "eogs_debug_counters": (C.c_int, [c_ptr, C.c_int]),
    "eogs_debug_alpha_cut": (C.c_int, [c_ptr, C.c_int, c_f32p, c_f32p, c_ptr]),
    "eogs_debug_depth_order_bytes": (C.c_size_t, [C.c_int]),
    "eogs_debug_depth_order": (C.c_int, [c_ptr, C.c_int, c_ptr, c_ptr, c_ptr]),
    "eogs_export_state": (C.c_int, [
        c_ptr, C.c_int, C.c_int, C.c_int, C.c_uint32, c_ptr, c_ptr, c_ptr,
        c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
}

ABI_VERSION = 7
_lib = None


class EogsRasterError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libeogs_raster.so and bind every declared symbol; raise loudly otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EogsRasterError(
            f"{LIB_PATH} is missing: build it with `python -m eogs2_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the rasterizer.")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise EogsRasterError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.eogs_abi_version() != ABI_VERSION:
        raise EogsRasterError(f"ABI version mismatch: library {lib.eogs_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().eogs_last_error().decode(errors="replace")
        raise EogsRasterError(f"{what} failed (rc={rc}): {msg}")

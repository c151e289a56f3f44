"""ctypes binding of libtensorf_b200.so (include/tensorf_b200.h).

The library is the product: there is NO fallback.  If the shared object is missing (not
built) every op raises `TensorfLibraryError` — loudly, at first use.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import threading
from typing import Optional

_HERE = pathlib.Path(__file__).resolve().parent
LIB_NAME = "libtensorf_b200.so"
LIB_PATH = pathlib.Path(os.environ.get("TENSORF_B200_LIB", _HERE / LIB_NAME))


class TensorfLibraryError(RuntimeError):
    """The CUDA library is missing or failed to load."""


class TensorfError(RuntimeError):
    """An entry point returned a non-zero status (message from tensorf_last_error())."""

    def __init__(self, status: int, message: str):
        super().__init__(f"tensorf_b200 status {status}: {message}")
        self.status = status
        self.message = message


class RenderDesc(C.Structure):
    """struct tensorf_render_desc"""

    _fields_ = [
        ("R", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("G", C.c_int32),
        ("cd", C.c_int32), ("ca", C.c_int32), ("mode", C.c_int32), ("contracted", C.c_int32),
        ("squash", C.c_int32), ("units", C.c_int32), ("feat_freqs", C.c_int32), ("view_freqs", C.c_int32),
        ("num_cameras", C.c_int32), ("mlp_impl", C.c_int32), ("loss_scale", C.c_float), ("flags", C.c_int32),
    ]


PARAM_FIELDS = (
    "density_vector", "density_matrix", "appearance_vector", "appearance_matrix",
    "w0", "w1", "b1", "w2", "b2", "w3", "b3", "embed",
)


class Params(C.Structure):
    """struct tensorf_params"""

    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


INPUT_FIELDS = ("origins", "directions", "camera_indices", "aabb", "jitter", "gumbel", "base_ts", "deltas", "colors")


class RenderInputs(C.Structure):
    """struct tensorf_render_inputs"""

    _fields_ = [(n, C.c_void_p) for n in INPUT_FIELDS]


class AdamDesc(C.Structure):
    """struct tensorf_adam_desc"""

    _fields_ = [
        ("n_leaves", C.c_int32), ("reserved", C.c_int32),
        ("b1", C.c_float), ("b2", C.c_float), ("eps", C.c_float), ("eps_root", C.c_float),
        ("bias_correction1", C.c_float), ("bias_correction2", C.c_float), ("lr_decay", C.c_float), ("reserved2", C.c_float),
    ]


class PeerAdamDesc(C.Structure):
    """struct tensorf_peer_adam_desc"""

    _fields_ = [("adam", AdamDesc), ("rank", C.c_int32), ("world", C.c_int32), ("total", C.c_int64),
                ("shard_begin", C.c_int64), ("shard_end", C.c_int64)]


ADAM_MAX_LEAVES = 16
PEER_MAX_WORLD = 16
PEER_MAX_LEAVES = 32

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
_pd, _pp, _pi = C.POINTER(RenderDesc), C.POINTER(Params), C.POINTER(RenderInputs)

# name -> (restype, argtypes). Must list every symbol include/tensorf_b200.h declares.
SIGNATURES = {
    "tensorf_last_error": (C.c_char_p, []),
    "tensorf_version": (_i, []),
    "tensorf_launch_count": (_i64, []),
    "tensorf_profile_enable": (_i, [_i]),
    "tensorf_profile_read": (_i, [_i, C.c_char_p, C.POINTER(C.c_float), C.POINTER(_i), C.POINTER(_i)]),
    "tensorf_vm_packed_floats": (_i64, [_i, _i]),
    "tensorf_vm_pack": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "tensorf_vm_unpack": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "tensorf_vm_interp_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i64, _i]),
    "tensorf_vm_interp_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _i]),
    "tensorf_topk_select": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "tensorf_segment_probabilities": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "tensorf_tc_rowgemm_test": (_i, [_vp, _vp, _i64, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i]),
    "tensorf_tc_trace_read": (_i, [_vp, _i]),
    "tensorf_tc_umma_bench": (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    "tensorf_tc_redgemm_test": (_i, [_vp, _vp, _i, _vp, _i, _i64, _vp]),
    "tensorf_tc_umma_probe": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "tensorf_mlp_workspace_bytes": (_i64, [_pd, _i64]),
    "tensorf_mlp_workspace_layout": (_i, [_pd, _i64, C.POINTER(_i64)]),
    "tensorf_mlp_fwd": (_i, [_vp, _pd, _pp, _vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "tensorf_mlp_bwd": (_i, [_vp, _pd, _pp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _pp]),
    "tensorf_render_workspace_bytes": (_i, [_pd, C.POINTER(_i64)]),
    "tensorf_render_rgb_fwd": (_i, [_vp, _pd, _pp, _pi, _vp, _vp, _vp]),
    "tensorf_render_rgb_bwd": (_i, [_vp, _pd, _pp, _pi, _vp, _vp, _pp]),
    "tensorf_render_rgb_bwd_phase": (_i, [_vp, _pd, _pp, _pi, _vp, _vp, _pp, _i]),
    "tensorf_render_depth": (_i, [_vp, _pd, _pp, _pi, _vp, _vp]),
    "tensorf_render_workspace_view": (_i, [_pd, _vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_i64)]),
    "tensorf_adam_scratch_bytes": (_i64, [C.POINTER(_i64), _i]),
    "tensorf_adam_step": (_i, [_vp, C.POINTER(AdamDesc), C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                               C.POINTER(_vp), C.POINTER(C.c_float), _vp, _vp, _i64]),
    "tensorf_peer_shard": (None, [_i64, _i, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "tensorf_peer_adam_scratch_bytes": (_i64, [_i64]),
    "tensorf_adam_step_peer": (_i, [_vp, C.POINTER(PeerAdamDesc), C.POINTER(_i64), C.POINTER(C.c_float), C.POINTER(_vp),
                                    C.POINTER(_vp), _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp, _i64]),
    "tensorf_peer_allreduce": (_i, [_vp, _i, _i, _i64, C.POINTER(_vp), _vp]),
    "tensorf_peer_set_max_ctas": (_i, [_i]),
    "tensorf_peer_allreduce_sync": (_i, [_vp, _i, _i, _i64, C.POINTER(_vp), _vp, C.POINTER(_vp), _vp, C.c_uint32]),
    "tensorf_peer_grad_norm": (_i, [_vp, _vp, _i, _vp]),
    "tensorf_vm_resize_scratch_bytes": (_i64, [_i, _i, _i]),
    "tensorf_vm_resize": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i64]),
    "tensorf_threefry2x32": (None, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "tensorf_prng_uniform": (_i, [_vp, C.c_uint32, C.c_uint32, _i64, C.c_float, C.c_float, _vp]),
    "tensorf_prng_uniform_slice": (_i, [_vp, C.c_uint32, C.c_uint32, _i64, _i64, C.c_float, C.c_float, _vp]),
    "tensorf_prng_gumbel": (_i, [_vp, C.c_uint32, C.c_uint32, _i64, _vp]),
    "tensorf_pixel_rays": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _i, _i, _i, C.c_uint32, _vp, _vp, _vp]),
    "tensorf_pixel_rays_striped": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _i, _i, _i, _i, _i, C.c_uint32, _vp, _vp, _vp, C.POINTER(_i64)]),
    "tensorf_gather_rays": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "tensorf_rgba_over_white": (_i, [_vp, _vp, _i64, _vp]),
}

_lock = threading.Lock()
_lib: Optional[C.CDLL] = None


def load(path: Optional[os.PathLike] = None) -> C.CDLL:
    """Load (once) and return the library with all signatures bound."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = pathlib.Path(path) if path is not None else LIB_PATH
        if not p.exists():
            raise TensorfLibraryError(
                f"{p} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C tensorf-jax_b200/csrc`. There is no CPU fallback."
            )
        try:
            lib = C.CDLL(str(p))
        except OSError as e:  # pragma: no cover
            raise TensorfLibraryError(f"failed to load {p}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise TensorfLibraryError(f"{p} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().tensorf_last_error()
        raise TensorfError(status, msg.decode() if msg else "")

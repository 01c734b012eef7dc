"""ctypes binding of libfolax_b200.so (the C ABI of include/folax_b200.h).

There is deliberately no CPU fallback: if the shared library is missing or a call fails the
binding raises.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfolax_b200.so")

F32, F64 = 0, 1
ELEMENTS = {"hexahedron": 0, "quad": 1, "tetra": 2, "triangle": 3}
PHYSICS = {"mechanical": 0, "thermal": 1, "neohooke": 2, "j2plasticity": 3, "stvenant": 4,
           "transient_thermal": 5, "allen_cahn": 6, "neohooke_ad": 7, "stvenant_ad": 8}
NUM_PARAMS = 12
MESH_AFFINE = 1     # FOL_MESH_AFFINE

_vp, _i32p, _u8p, _i64, _int, _dbl = C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double

# name -> (restype, argtypes); every symbol declared in include/folax_b200.h
SIGNATURES = {
    "fol_last_error": (C.c_char_p, []),
    "fol_version": (_int, []),
    "fol_launch_count": (_i64, []),
    "fol_set_tuned_kernels": (_int, [_int]),
    "fol_set_grid_margin": (_int, [_int]),
    "fol_element_info": (_int, [_int, _int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "fol_dofs_per_node": (_int, [_int, _int]),
    "fol_bcoo_indices": (_int, [_vp, _i32p, _i64, _int, _int, _i32p]),
    "fol_dirichlet_flags": (_int, [_vp, _i32p, _i64, _i64, _u8p]),
    "fol_node_adjacency": (_int, [_vp, _i32p, _i64, _int, _i64, _i32p, _i32p, _i32p]),
    "fol_assemble_elements": (_int, [_vp, _int, _int, _int, _int, _int, _i64, _i64, _vp, _i32p, _vp, _vp,
                                     _u8p, C.POINTER(_dbl), _vp, _vp, _vp, _vp]),
    "fol_residual_gather": (_int, [_vp, _int, _i64, _int, _int, _i32p, _i32p, _vp, _vp]),
    "fol_csr_plan_count_host": (_int, [_i32p, _i64, _int, _i64, _i32p, _i32p, _i32p]),
    "fol_csr_plan_fill_host": (_int, [_i32p, _i64, _int, _i64, _int, _i32p, _i32p, _vp, _i32p, _i32p, _i32p, _i32p,
                                      _i32p, _i32p]),
    "fol_sell_plan_fill_host": (_int, [_vp, _i32p, _i64, _int, _int, _vp, _i32p, _i32p, _i32p, _i32p,
                                       C.POINTER(_int)]),
    "fol_csr_values": (_int, [_vp, _int, _i64, _int, _int, _i32p, _i32p, _i32p, _i32p, _vp, _vp]),
    "fol_apply_jacobian_elements": (_int, [_vp, _int, _int, _int, _int, _int, _i64, _i64, _vp, _i32p, _vp, _vp, _u8p,
                                           C.POINTER(_dbl), _vp, _vp, _vp]),
    "fol_apply_jacobian_elements_batched": (_int, [_vp, _int, _int, _int, _int, _int, _i64, _i64, _i64, _vp, _i32p, _vp,
                                                   _vp, _u8p, C.POINTER(_dbl), _vp, _vp]),
    "fol_residual_gather_batched": (_int, [_vp, _int, _i64, _int, _int, _i64, _i64, _i32p, _i32p, _vp, _vp]),
    "fol_residual_adjoint_elements_batched": (_int, [_vp, _int, _int, _int, _int, _i64, _i64, _i64, _vp, _i32p, _vp, _vp,
                                                     _vp, _vp, C.POINTER(_dbl), _vp]),
    "fol_geometry_cache": (_int, [_vp, _int, _int, _int, _i64, _vp, _i32p, _vp]),
    "fol_geometry_width": (_int, [_int, _int]),
    "fol_geometry_cache_physics": (_int, [_vp, _int, _int, _int, _int, _i64, _vp, _i32p, _vp, _vp]),
    "fol_energy_work_size": (_i64, [_i64, _i64]),
    "fol_energy_and_grads": (_int, [_vp, _int, _int, _int, _int, _i64, _i64, _i64, _vp, _i32p, _i32p, _i32p,
                                    _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i64, _i64, _i64, _i64,
                                    _vp, _vp, _vp, _u8p, _dbl, C.POINTER(_dbl), _vp, _vp, _vp, _vp]),
    "fol_energy_and_grads_flags": (_int, [_vp, _int, _int, _int, _int, _i64, _i64, _i64, _vp, _i32p, _i32p, _i32p,
                                    _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i64, _i64, _i64, _i64,
                                    _vp, _vp, _vp, _u8p, _dbl, C.POINTER(_dbl), _vp, _vp, _vp, _vp, _i64]),
    "fol_energy_grid_work_size": (_i64, [_i64, _i64, _i64]),
    "fol_energy_and_grads_grid": (_int, [_vp, _int, _i64, _i64, _i64, C.POINTER(_dbl), _dbl, _vp, _vp, _vp, _u8p, _u8p,
                                         _dbl, C.POINTER(_dbl), _vp, _vp, _vp, _vp]),
    "fol_energy_grid_mech_work_size": (_i64, [_i64, _i64, _i64]),
    "fol_energy_and_grads_grid_mech": (_int, [_vp, _int, _i64, _i64, _i64, C.POINTER(_dbl), _dbl, _vp, _vp, _vp, _u8p,
                                              _u8p, _dbl, C.POINTER(_dbl), _vp, _vp, _vp]),
    "fol_loss_reduce": (_int, [_vp, _int, _i64, _dbl, _vp, _vp, _vp]),
    "fol_scale_grads": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _dbl, _vp, _int, _u8p, _vp, _vp]),
    "fol_apply_dirichlet": (_int, [_vp, _int, _i64, _i64, _i32p, _i64, _vp, _int, _dbl, _vp]),
    "fol_gauss_interpolate": (_int, [_vp, _int, _int, _int, _int, _i64, _i32p, _vp, _vp, _vp, _vp]),
    "fol_response_elements": (_int, [_vp, _int, _int, _int, _int, _i64, _vp, _i32p, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp]),
    "fol_residual_adjoint_elements": (_int, [_vp, _int, _int, _int, _int, _int, _i64, _vp, _i32p, _vp, _vp, _vp, _vp,
                                             C.POINTER(_dbl), _vp, _vp]),
    "fol_element_energies": (_int, [_vp, _int, _int, _int, _int, _i64, _vp, _i32p, _vp, _vp, _vp, C.POINTER(_dbl),
                                    _vp]),
    "fol_sum": (_int, [_vp, _int, _i64, _vp, _vp]),
    "fol_gather_values": (_int, [_vp, _int, _i64, _i32p, _vp, _vp]),
    "fol_sell_spmv": (_int, [_vp, _int, _i64, _vp, _i32p, _vp, _vp, _vp]),
    "fol_sell_spmv_block": (_int, [_vp, _int, _int, _i64, _vp, _i32p, _vp, _vp, _vp]),
    "fol_vec_op": (_int, [_vp, _int, _int, _i64, _dbl, _vp, _dbl, _vp, _vp]),
    "fol_bicg_scalar_count": (_int, []),
    "fol_bicg_scalars": (_int, [_vp, _int, _int, _vp]),
    "fol_vec_op_dev": (_int, [_vp, _int, _i64, _vp, _int, _int, _dbl, _vp, _int, _dbl, _vp, _vp]),
    "fol_dot_work_size": (_i64, []),
    "fol_bicgstab_fused_work_size": (_i64, [_i64]),
    "fol_bicgstab_fused": (_int, [_vp, _int, _int, _i64, _vp, _i32p, _vp, _vp, _vp, _vp, _dbl, _dbl, _i64, _vp]),
    "fol_dot": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _vp]),
    "fol_plan_create": (_int, [C.POINTER(_vp), _int, _int, _int, _int, _i64, _i64, _vp, _i32p, _i32p, _i64,
                               C.POINTER(_dbl)]),
    "fol_plan_destroy": (None, [_vp]),
    "fol_plan_set_csr": (_int, [_vp, _i64, _i64, _i32p, _i32p, _i32p, _i32p]),
    "fol_plan_assemble_host_csr": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "fol_halo_create": (_int, [C.POINTER(_vp), _int, _i64]),
    "fol_halo_destroy": (None, [_vp]),
    "fol_halo_export": (_int, [_vp, _vp]),
    "fol_halo_connect": (_int, [_vp, _int, _vp]),
    "fol_halo_gather_push": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _int, _i32p, _i32p, _vp, _vp]),
    "fol_halo_add": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _int, _vp]),
    "fol_halo_timeouts": (_i64, [_vp]),
    "fol_assemble_elements_halo": (_int, [_vp, _int, _int, _int, _int, _i64, _i64, _vp, _i32p, _vp, _vp, _u8p, C.POINTER(_dbl), _vp,
                                          _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32p, _i32p, _vp]),
    "fol_residual_gather_halo": (_int, [_vp, _vp, _i64, _i64, _i64, _i32p, _i32p, _vp, _vp]),
    "fol_plan_assemble_host": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "fol_plan_assemble_device": (_int, [_vp, _int, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "fol_plan_stream": (_vp, [_vp]),
    "fol_measure_fma_peak": (_int, [_int, C.POINTER(_dbl)]),
    "fol_measure_write_bandwidth": (_int, [_i64, C.POINTER(_dbl)]),
}

_lib = None


class FolaxError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:   # build in-tree on first use (nvcc cross-compiles sm_100a without a GPU)
                from . import build as _build
                _build.build()
            except Exception as ex:
                raise FolaxError(f"{LIB_PATH} is missing and could not be built ({ex}); run "
                                 "`python -m folax_b200.build` (nvcc, sm_100a). folax_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise FolaxError(f"libfolax_b200 error {rc}: {load().fol_last_error().decode()}")


def dtype_code(dt):
    if dt == torch.float64:
        return F64
    if dt == torch.float32:
        return F32
    raise FolaxError(f"unsupported dtype {dt}")


try:                                   # the raw handle without building a torch.cuda.Stream object: current_stream()
    _raw_stream, _cur_device = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice   # costs ~20 us per call
except AttributeError:                 # (a torch without these private entry points: the public, slower route)
    _raw_stream = _cur_device = None


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (what every C-ABI call launches on)."""
    if _raw_stream is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise FolaxError("expected a contiguous CUDA tensor")
    return t.data_ptr()


def params_array(values):
    arr = (_dbl * NUM_PARAMS)()
    for i, v in enumerate(values):
        arr[i] = float(v)
    return arr


def require_cuda():
    if not torch.cuda.is_available():
        raise FolaxError("folax_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")


def to_device(x, dtype, device=None):
    """numpy / torch (any device) -> contiguous CUDA tensor of `dtype`."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x)), device=device).to(dtype).contiguous()

"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol that include/folax_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from folax_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "folax_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fol_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/folax_b200.h but not exported"


def test_binding_covers_header(lib):
    assert set(_declared_symbols()) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"


def test_element_table(lib):
    a, d, g = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    expect = {("hexahedron", 2): (8, 3, 8), ("hexahedron", 3): (8, 3, 27), ("quad", 2): (4, 2, 4),
              ("tetra", 1): (4, 3, 1), ("tetra", 2): (4, 3, 4), ("triangle", 3): (3, 2, 4)}
    for (name, order), want in expect.items():
        assert lib.fol_element_info(_lib.ELEMENTS[name], order, ctypes.byref(a), ctypes.byref(d),
                                    ctypes.byref(g)) == 0
        assert (a.value, d.value, g.value) == want
    assert lib.fol_element_info(7, 2, None, None, None) != 0
    assert b"bad element" in lib.fol_last_error()
    assert lib.fol_dofs_per_node(_lib.PHYSICS["thermal"], 0) == 1
    assert lib.fol_dofs_per_node(_lib.PHYSICS["mechanical"], 1) == 2
    assert lib.fol_version() >= 100


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import folax_b200
    from folax_b200.loss_functions import ThermalLoss2DQuad
    mesh = folax_b200.create_2D_square_mesh(1.0, 3)
    loss = ThermalLoss2DQuad("t", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}}, mesh)
    with pytest.raises(_lib.FolaxError):
        loss.Initialize()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "folax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_xla_ffi_shim_compiles_against_a_stub_of_the_ffi_header(tmp_path):
    """JAX / jaxlib are absent, so folax_b200/ffi/xla_ffi_shim.cc is normally preprocessed away.  Compiled here
    against tests/xla_stub (a stand-in for the part of xla/ffi/api/ffi.h it uses): valid C++, calls that match the
    current include/folax_b200.h, and handler signatures that match their Ffi::Bind() chains.  Not a behavioural
    test -- XLA itself is not involved."""
    import shutil
    import subprocess
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.isdir(cuda_inc):
        pytest.skip("needs g++ and the CUDA headers")
    obj = str(tmp_path / "shim.o")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror=return-type", "-c",
                        os.path.join(ROOT, "folax_b200", "ffi", "xla_ffi_shim.cc"),
                        "-I", os.path.join(ROOT, "tests", "xla_stub"), "-I", cuda_inc, "-o", obj],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", obj], capture_output=True, text=True).stdout
    for name in ("FolAssembleElements", "FolResidualGather", "FolEnergyAndGrads", "FolApplyJacobianElements",
                 "FolGaussInterpolate", "FolResponseElements", "FolResidualAdjointElements", "FolGeometryCache"):
        assert f"{name}_stub_marker" in syms, f"{name} was not compiled (is the shim still header-guarded away?)"


def test_jax_binding_module_is_importable_and_says_what_it_needs():
    """folax_b200/ffi/jax_binding.py (the reference-side jax.ffi + custom_vjp classes) cannot run here: it must
    import cleanly and fail with a clear message instead of an ImportError deep inside."""
    import importlib.util
    from folax_b200.ffi import jax_binding
    assert set(jax_binding.HANDLERS) >= {"FolAssembleElements", "FolResidualGather", "FolEnergyAndGrads"}
    if importlib.util.find_spec("jax") is None:
        with pytest.raises(_lib.FolaxError, match="needs JAX"):
            jax_binding.register()
        with pytest.raises(_lib.FolaxError, match="needs JAX"):
            jax_binding.accelerate(object, "mechanical")
    # every handler the module registers exists in the shim
    shim_src = open(os.path.join(ROOT, "folax_b200", "ffi", "xla_ffi_shim.cc")).read()
    for name in jax_binding.HANDLERS:
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({name}," in shim_src, name


@pytest.mark.parametrize("name", ["store_path_bench", "cabi_selfcheck"])
def test_micro_programs_compile_for_sm100a(tmp_path, name):
    """scripts/micro/*.cu (the store-path microbenchmark behind profiles/r2/store_path_micro_*.jsonl and the torch-free
    tuned-vs-generic check through the C ABI) stay buildable: nvcc cross-compiles them for sm_100a without a GPU."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "micro",
                       name + ".cu")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-c", src, "-o",
                        str(tmp_path / (name + ".o"))], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

"""The XLA-FFI layer (tensorf-jax_b200/jax_ffi) cannot run in this image (no JAX).  What CAN be checked here:
the shim type-checks against the C ABI header and a minimal stand-in for xla/ffi/api/ffi.h whose Bind().To() asserts
that every handler is invocable with exactly the types its binding declares; the stand-in itself rejects a mismatched
handler; the Python binding compiles and only names handlers the shim defines; INTEGRATION.md only names functions that
exist."""
import ast
import pathlib
import re
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent
FFI = ROOT / "tensorf-jax_b200" / "jax_ffi"
FLAGS = ["g++", "-fsyntax-only", "-std=c++17", f"-I{ROOT / 'tests' / 'mock_xla'}", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include"]


def test_shim_type_checks_against_the_c_abi():
    r = subprocess.run(FLAGS + [str(FFI / "xla_ffi_shim.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_mock_rejects_a_mismatched_handler(tmp_path):
    src = tmp_path / "bad.cc"
    src.write_text('#include "xla/ffi/api/ffi.h"\nnamespace ffi = xla::ffi;\n'
                   "static ffi::Error H(ffi::Buffer<ffi::F32> a, ffi::Result<ffi::Buffer<ffi::U8>> r) { return ffi::Error::Success(); }\n"
                   "XLA_FFI_DEFINE_HANDLER_SYMBOL(Bad, H, ffi::Ffi::Bind().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());\n")
    r = subprocess.run(FLAGS + [str(src)], capture_output=True, text=True)
    assert r.returncode != 0 and "does not match its binding" in r.stderr


def test_python_binding_names_only_existing_handlers_and_functions():
    shim = (FFI / "xla_ffi_shim.cc").read_text()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\(\s*(\w+)", shim))
    assert handlers == {"TensorfRenderRgbFwd", "TensorfRenderRgbBwd", "TensorfRenderDepth", "TensorfVmInterpFwd", "TensorfVmInterpBwd",
                        "TensorfAdamStep"}
    src = (FFI / "tensorf_jax.py").read_text()
    tree = ast.parse(src)
    used = set(re.findall(r'"(Tensorf\w+)"', src))
    assert used == handlers
    funcs = {n.name for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
    assert {"render_rays", "interpolate", "adam_step"} <= funcs
    assert "custom_vjp" in src and "defvjp" in src
    assert "import torch" not in src and "tensorf_b200 import" not in src  # torch-free
    # every C-ABI function the shim calls is declared in the header
    header = (ROOT / "include" / "tensorf_b200.h").read_text()
    for fn in set(re.findall(r"\b(tensorf_\w+)\(", shim)):
        assert re.search(rf"\b{fn}\(", header), fn
    # `from tensorf_jax import a, b` must name functions of the module
    doc = (ROOT / "INTEGRATION.md").read_text()
    for m in re.finditer(r"from tensorf_jax import ([\w, ]+)", doc):
        for name in m.group(1).split(","):
            name = name.split(" as ")[0]
            assert name.strip() in funcs, f"INTEGRATION.md names tensorf_jax.{name.strip()}, which does not exist"

"""Compile the reference's own CPU op into oracle/_ref/custom_ops.so (TEST INFRASTRUCTURE ONLY).

The source is compiled from where it lies -- /root/reference/ggml-cpu/custom_ops.cpp (which includes
ggml-quants.hpp and ggml-common.hpp next to it) -- with one g++ command against this image's torch
headers.  Nothing is copied into the repo; the only output is the shared object under oracle/_ref/,
which is git-ignored but travels to the GPU box with the gpurun snapshot.

The module is the pybind11 extension the reference's tests import as `custom_ops`
(ggml-cpu/custom_ops.cpp:38-40, tests/test_dequantize.py:8).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/ggml-cpu/custom_ops.cpp"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "custom_ops.so")


def build(force: bool = False) -> str | None:
    """Returns the path of the built module, or None when the reference tree is absent."""
    if not os.path.exists(REF_SRC):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(REF_SRC):
        return OUT
    import torch
    from torch.utils import cpp_extension

    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in cpp_extension.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [
        "/usr/bin/g++", "-O2", "-fPIC", "-shared", "-std=c++17",
        "-DTORCH_EXTENSION_NAME=custom_ops", "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
        *inc, REF_SRC, "-o", OUT,
        f"-L{libdir}", f"-Wl,-rpath,{libdir}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
    ]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv)
    print(path if path else "reference tree not present; nothing built")

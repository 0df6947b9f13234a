"""Compile the reference's own CUDA ops for sm_100a into oracle/_ref_cuda/_ggml.so (TEST INFRASTRUCTURE ONLY).

The reference's three CUDA ops -- ggml_dequantize, ggml_mul_mat_vec_a8 (quantize_q8_1 + mul_mat_vec_q) and
ggml_mul_mat_a8 (quantize_mmq_q8_1 + mul_mat_q) -- are compiled UNMODIFIED from where they lie:

    /root/reference/hf-kernels/ggml-kernels/ggml/ggml_kernel.cu, ggml/mmq.cu,
    ggml/kernel_instances/mmq_*.cu (30 instance files), torch-ext/torch_binding.cpp

one nvcc command per source (in parallel) against this image's torch headers, then one link.  This is not the
reference's own build system (kernel-builder / nix / build2cmake, absent here): it is the source list of its
build.toml handed to nvcc.  Nothing is copied into the repo; the only outputs are under oracle/_ref_cuda/
(git-ignored; the .so travels to the GPU box with the gpurun snapshot, the objects do not).

What it is used for (tests/ and bench.py's `gpu_reference` legs only -- never the product path):
  * parity anchor: q8_1 + MMVQ and q8_1_mmq + MMQ outputs of the reference's kernels on the same W / X
    (tests/test_gpu_vs_ref_cuda.py), which pins the oracle's restatement of those steps to the reference itself;
  * same-box GPU comparator: the reference ops timed next to ours (bench.py).

The library registers torch.ops._ggml.{ggml_dequantize, ggml_mul_mat_vec_a8, ggml_mul_mat_a8}
(torch-ext/torch_binding.cpp:6-21 with TORCH_EXTENSION_NAME=_ggml, the name torch-ext/ggml/__init__.py:8-10
falls back to); load it with `oracle.ref_cuda_ops()`.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/hf-kernels/ggml-kernels"
OUT_DIR = os.path.join(HERE, "_ref_cuda")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
OUT = os.path.join(OUT_DIR, "_ggml.so")
SHIM_SRC = os.path.join(HERE, "ref_cuda_q8_shim.cu")       # ours: exposes the reference's file-static q8_1 quantizers
SHIM_OUT = os.path.join(OUT_DIR, "_ggml_q8shim.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources() -> list[str]:
    return ([os.path.join(REF_ROOT, "ggml", "ggml_kernel.cu"), os.path.join(REF_ROOT, "ggml", "mmq.cu")]
            + sorted(glob.glob(os.path.join(REF_ROOT, "ggml", "kernel_instances", "mmq_*.cu")))
            + [os.path.join(REF_ROOT, "torch-ext", "torch_binding.cpp")])


def build(force: bool = False, jobs: int | None = None) -> str | None:
    """Returns the path of the built library, or None when the reference tree is absent and nothing was built before."""
    if not os.path.isdir(REF_ROOT):
        return OUT if os.path.exists(OUT) else None
    srcs = sources()
    newest = max(os.path.getmtime(p) for p in srcs + glob.glob(os.path.join(REF_ROOT, "ggml", "*.h*")) + [SHIM_SRC])
    if (os.path.exists(OUT) and os.path.exists(SHIM_OUT) and not force
            and min(os.path.getmtime(OUT), os.path.getmtime(SHIM_OUT)) >= newest):
        return OUT
    import torch
    from torch.utils import cpp_extension

    os.makedirs(OBJ_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in cpp_extension.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    common = [
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
        "--expt-relaxed-constexpr", "--expt-extended-lambda", "-DTORCH_EXTENSION_NAME=_ggml",
        f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc,
    ]

    def one(job) -> str:
        src, extra, name = job if isinstance(job, tuple) else (job, [], os.path.basename(job).rsplit(".", 1)[0])
        obj = os.path.join(OBJ_DIR, name + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
            return obj
        r = subprocess.run([NVCC, *common, *extra, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
        return obj

    ref_inc = ["-I" + os.path.join(REF_ROOT, "ggml")]
    shim_jobs = [(SHIM_SRC, ["-DSHIM_PART=1", *ref_inc], "q8shim_1"), (SHIM_SRC, ["-DSHIM_PART=2", *ref_inc], "q8shim_2")]
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
        all_objs = list(ex.map(one, srcs + shim_jobs))
    objs, shim_objs = all_objs[:len(srcs)], all_objs[len(srcs):]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           f"-L{libdir}", f"-Xlinker=-rpath,{libdir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-ltorch_python"]
    for c in (cmd, [NVCC, "-shared", "-o", SHIM_OUT, *shim_objs, *cmd[4 + len(objs):]]):
        r = subprocess.run(c, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv)
    print(path if path else "reference tree not present; nothing built")

// ref_cuda_q8_shim.cu -- TEST INFRASTRUCTURE ONLY (built by oracle/build_ref_cuda.py into
// oracle/_ref_cuda/_ggml_q8shim.so; never linked into the product).
//
// The reference's two activation quantizers are file-static and reachable only through its matmul ops:
//   quantize_q8_1      + quantize_row_q8_1_cuda   hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66
//   quantize_mmq_q8_1  + quantize_mmq_q8_1_cuda   hf-kernels/ggml-kernels/ggml/mmq.cu:84-177
// so no reference test can look at the q8_1 bytes ("parity unpinned").  This shim compiles the
// reference's translation unit UNMODIFIED, from where it lies, inside a namespace (one #include; nothing is copied),
// and exposes the static launchers as two torch ops:
//   torch.ops._ggml_q8shim.quantize_q8_1(Tensor X) -> Tensor              int32 [rows, Kp/32*9],  Kp = ceil(K/512)*512
//   torch.ops._ggml_q8shim.quantize_q8_1_mmq(Tensor X, int type) -> Tensor int32 [rows, Kp/32*9], Kp = K - K%512 + 512
// allocated exactly as the reference's ops allocate them (ggml_kernel.cu:89-90, mmq.cu:207-208).  The GPU parity tests
// compare ggb_quantize_q8_1's bytes with these, which pins rows a7 / a11 to the reference's own kernels.
//
// SHIM_PART selects the translation unit (the reference's headers have no include guards, so one per object file).
#include <stdio.h>
#include <assert.h>
#include <limits.h>

#include <algorithm>
#include <cstdint>

#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <torch/all.h>
#include <torch/library.h>
#include <c10/cuda/CUDAGuard.h>
#include <ATen/cuda/CUDAContext.h>

#if SHIM_PART == 1

namespace ref_q8 {
#include "ggml_kernel.cu"
}  // namespace ref_q8

static torch::Tensor shim_quantize_q8_1(torch::Tensor X) {
    TORCH_CHECK(X.is_cuda() && X.dim() == 2 && X.is_contiguous(), "X must be a contiguous 2-D CUDA tensor");
    const at::cuda::OptionalCUDAGuard guard(device_of(X));
    const int col = (int)X.size(1), rows = (int)X.size(0);
    const int padded = (col + 512 - 1) / 512 * 512;                                   // ggml_kernel.cu:84
    auto out = torch::zeros({rows, padded / 32 * 9}, torch::TensorOptions().dtype(torch::kInt32).device(X.device()));
    cudaStream_t stream = at::cuda::getCurrentCUDAStream().stream();
    VLLM_DISPATCH_FLOATING_TYPES(X.scalar_type(), "shim_quantize_q8_1", [&] {
        ref_q8::quantize_row_q8_1_cuda<scalar_t>((scalar_t *)X.data_ptr(), (void *)out.data_ptr(), col, rows, stream);
    });
    return out;
}

TORCH_LIBRARY_FRAGMENT(_ggml_q8shim, m) {
    m.def("quantize_q8_1(Tensor X) -> Tensor");
}
TORCH_LIBRARY_IMPL(_ggml_q8shim, CUDA, m) {
    m.impl("quantize_q8_1", &shim_quantize_q8_1);
}

#else

namespace ref_q8m {
#include "mmq.cu"
// mmq.cu's ggml_mul_mat_a8 references the 30 explicit instantiations that live in kernel_instances/*.cu; the shim never
// calls it, so the symbols are satisfied with empty bodies (this library is separate from _ggml.so, which has the real ones)
template <typename scalar_t, ggml_type type>
void mul_mat_q_case(const mmq_args<scalar_t> &, cudaStream_t) {}
#define SHIM_ALL_TYPES(S)                                                                                            \
    DECL_MMQ_CASE(S, GGML_TYPE_Q4_0); DECL_MMQ_CASE(S, GGML_TYPE_Q4_1); DECL_MMQ_CASE(S, GGML_TYPE_Q5_0);            \
    DECL_MMQ_CASE(S, GGML_TYPE_Q5_1); DECL_MMQ_CASE(S, GGML_TYPE_Q8_0); DECL_MMQ_CASE(S, GGML_TYPE_Q2_K);            \
    DECL_MMQ_CASE(S, GGML_TYPE_Q3_K); DECL_MMQ_CASE(S, GGML_TYPE_Q4_K); DECL_MMQ_CASE(S, GGML_TYPE_Q5_K);            \
    DECL_MMQ_CASE(S, GGML_TYPE_Q6_K);
SHIM_ALL_TYPES(float)
SHIM_ALL_TYPES(c10::Half)
SHIM_ALL_TYPES(c10::BFloat16)
}  // namespace ref_q8m

static torch::Tensor shim_quantize_q8_1_mmq(torch::Tensor X, int64_t type) {
    TORCH_CHECK(X.is_cuda() && X.dim() == 2 && X.is_contiguous(), "X must be a contiguous 2-D CUDA tensor");
    const at::cuda::OptionalCUDAGuard guard(device_of(X));
    const int col = (int)X.size(1), batch = (int)X.size(0);
    const int padded = col - col % 512 + 512;                                          // mmq.cu:190-191
    auto out = torch::zeros({batch, padded / 32 * 9}, torch::TensorOptions().dtype(torch::kInt32).device(X.device()));
    cudaStream_t stream = at::cuda::getCurrentCUDAStream().stream();
    VLLM_DISPATCH_FLOATING_TYPES(X.scalar_type(), "shim_quantize_q8_1_mmq", [&] {
        ref_q8m::quantize_mmq_q8_1_cuda((scalar_t *)X.data_ptr(), (void *)out.data_ptr(), (int64_t)col, (int64_t)batch, type, stream);
    });
    return out;
}

TORCH_LIBRARY_FRAGMENT(_ggml_q8shim, m) {
    m.def("quantize_q8_1_mmq(Tensor X, int type) -> Tensor");
}
TORCH_LIBRARY_IMPL(_ggml_q8shim, CUDA, m) {
    m.impl("quantize_q8_1_mmq", &shim_quantize_q8_1_mmq);
}

#endif

// tp_barrier.cu -- the cross-rank barrier of a tensor-parallel group as ONE tiny programmatic-dependent launch.
//
// The fused exchange (ggb_mmvq_gather / ggb_mmvq_fused: every rank's GEMV epilogue stores its output slice into every rank's
// copy of a symmetric buffer) needs one barrier on the stream before the gathered result is read.  A library barrier is an
// ordinary kernel launch: it breaks the chain of programmatic dependent launches on both sides (the next GEMV's CTAs cannot
// become resident and pre-fill their weight rings while it runs) and costs 6-8 us per use at decode, two uses per MLP block
// (round-1 measurement: 48.6 us per block at tp = 8 against ~9 us of weight streaming).  This one is a PDL launch like
// every other kernel of the library: it releases its dependents at once, waits for the producing GEMV, then
//     thread t (< n ranks): st.release.sys  flags_of_rank_t[my rank] = epoch
//                           spin on ld.relaxed.sys  my_flags[t] until it reaches epoch, then one ld.acquire.sys (bounded: a lost peer sets an error
//                           word instead of hanging the GPU)
// `epoch` lives on the device (word 8 of the rank's own flag block, incremented by each launch), so the launch is
// CUDA-graph capturable: nothing on the host changes between replays.  Flags are monotonic: no reset, no second phase.
// There is no reference counterpart (the reference is single-GPU).
#include "ggb_common.cuh"

namespace {

struct Peers {
    uint32_t *p[8];
};

__device__ __forceinline__ void st_release_sys(uint32_t *addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *addr) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

constexpr int kEpochWord = 8, kErrorWord = 9;

__global__ void __launch_bounds__(32)
tp_barrier_kernel(const Peers peers, const int n, const int rank) {
    ggb_pdl_launch_dependents();        // the next kernel's CTAs may become resident (weight rings, barriers) meanwhile
    ggb_pdl_wait();                     // the producing kernel has completed: its peer stores have been issued
    uint32_t *mine = peers.p[rank];
    const int t = threadIdx.x;
    const uint32_t e = *reinterpret_cast<volatile uint32_t *>(mine + kEpochWord) + 1u;
    __syncwarp();
    if (t < n) {
        st_release_sys(peers.p[t] + rank, e);                 // (release at system scope: the producing kernel's peer stores first)
        unsigned spins = 0;
        while ((int32_t)(ld_relaxed_sys(mine + t) - e) < 0) {
            if (++spins > (1u << 25)) {                       // ~10 s: a peer is gone; do not hang the GPU
                mine[kErrorWord] = 1u + (uint32_t)t;
                break;
            }
        }
        (void)ld_acquire_sys(mine + t);                       // what rank t stored before its signal is visible from here on
    }
    __syncwarp();
    if (t == 0) *reinterpret_cast<volatile uint32_t *>(mine + kEpochWord) = e;
}

}  // namespace

extern "C" int ggb_tp_barrier(void *const *flag_peers, int n_peers, int rank, void *stream) {
    if (!flag_peers || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers) {
        ggb_set_error("ggb_tp_barrier: need 1..8 peer flag blocks and 0 <= rank < n_peers");
        return GGB_E_ARG;
    }
    Peers peers = {};
    for (int i = 0; i < n_peers; i++) {
        if (!flag_peers[i] || (reinterpret_cast<uintptr_t>(flag_peers[i]) & 3)) {
            ggb_set_error("ggb_tp_barrier: null or misaligned flag block %d", i);
            return GGB_E_ARG;
        }
        peers.p[i] = static_cast<uint32_t *>(flag_peers[i]);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<tp_barrier_kernel>(ggb_device_info().dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tp_barrier_kernel, peers, n_peers, rank));
    return GGB_OK;
}

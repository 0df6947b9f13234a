// Write-bandwidth probe: which store path fills HBM fastest on this GPU?  (denominator study for dequant)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbw_probe wbw_probe.cu && ./wbw_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: st.v4 default, 1: st.cs (streaming), 2: st.wt? (L2 evict_first hint), 3: st.cg
__global__ void fill_st(uint4 *out, size_t n16) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) out[i] = v;
        else if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        else if (MODE == 2) { if ((i & 1) == 0 && i + 1 < n16) asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
        else if (MODE == 4) { if ((i & 1) == 0 && i + 1 < n16) asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
        else asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

// each CTA owns contiguous tiles of TILE bytes (like the dequant kernel: one CTA = one 16 KB output tile)
template <int TILE>
__global__ void fill_tile(uint4 *out, size_t ntiles) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint4 *dst = out + t * (TILE / 16);
        for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) dst[i] = v;
    }
}

// TMA bulk store: smem tile -> global, TILE bytes per copy
template <int TILE>
__global__ void fill_tma(uint8_t *out, size_t ntiles) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * TILE), "r"(smem_u32(sm)), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// dequant-like mix: read 1 byte per 3.56 bytes written (Q4_K -> fp16): read R bytes per tile, write TILE
template <int TILE, int RD>
__global__ void mix_tile(const uint4 *in, uint4 *out, size_t ntiles) {
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint4 *src = in + t * (RD / 16);
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < RD / 16; i += blockDim.x) { const uint4 a = src[i]; acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w; }
        uint4 *dst = out + t * (TILE / 16);
        for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) dst[i] = acc;
    }
}

template <typename F>
float timeit(F f, int n = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) f(i);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < n; i++) f(i);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / n;
}

int main() {
    const size_t MB = 1 << 20;
    for (size_t bytes : {117 * MB + 448 * 1024, (size_t)1024 * MB}) {
        const int NBUF = bytes < 512 * MB ? 6 : 2;          // rotate buffers: > L2 between reuses
        uint8_t *buf[6]; for (int i = 0; i < NBUF; i++) cudaMalloc(&buf[i], bytes);
        uint8_t *in; cudaMalloc(&in, bytes / 3); cudaMemset(in, 1, bytes / 3);
        const size_t n16 = bytes / 16;
        printf("--- %zu MB per launch\n", bytes / MB);
        float ms;
        ms = timeit([&](int i) { fill_st<0><<<148 * 8, 256>>>((uint4 *)buf[i % NBUF], n16); });  printf("st.v4 grid-stride (148x8 CTAs)   %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_st<0><<<148 * 32, 256>>>((uint4 *)buf[i % NBUF], n16); }); printf("st.v4 grid-stride (148x32 CTAs)  %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_st<1><<<148 * 8, 256>>>((uint4 *)buf[i % NBUF], n16); });  printf("st.cs.v4                         %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_st<2><<<148 * 8, 256>>>((uint4 *)buf[i % NBUF], n16); });  printf("st.L2::evict_first.v8 (32 B/thr, half the threads idle) %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_st<4><<<148 * 8, 256>>>((uint4 *)buf[i % NBUF], n16); });  printf("st.v8.b32 (32 B/thr)             %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_st<3><<<148 * 8, 256>>>((uint4 *)buf[i % NBUF], n16); });  printf("st.cg.v4                         %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_tile<16384><<<(unsigned)(bytes / 16384), 256>>>((uint4 *)buf[i % NBUF], bytes / 16384); }); printf("16 KB tile per CTA (1 tile/CTA)   %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_tile<16384><<<148 * 8, 256>>>((uint4 *)buf[i % NBUF], bytes / 16384); }); printf("16 KB tiles, persistent 148x8     %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        cudaFuncSetAttribute(fill_tma<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        cudaFuncSetAttribute(fill_tma<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        ms = timeit([&](int i) { fill_tma<16384><<<148 * 4, 128, 16384>>>(buf[i % NBUF], bytes / 16384); }); printf("TMA bulk store 16 KB (148x4)      %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { fill_tma<65536><<<148 * 2, 128, 65536>>>(buf[i % NBUF], bytes / 65536); }); printf("TMA bulk store 64 KB (148x2)      %7.1f us %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&](int i) { mix_tile<16384, 4608><<<148 * 8, 256>>>((const uint4 *)in, (uint4 *)buf[i % NBUF], bytes / 16384); });
        printf("mix: read 4608 + write 16384 /tile %6.1f us %6.0f GB/s total (%.0f write)\n", ms * 1e3, (bytes + bytes / 16384 * 4608) / ms / 1e6, bytes / ms / 1e6);
        for (int i = 0; i < NBUF; i++) cudaFree(buf[i]);
        cudaFree(in);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

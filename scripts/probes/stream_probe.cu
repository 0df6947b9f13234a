// stream_probe.cu -- measurement probe (not part of the library): how fast can 148 SMs pull a weight matrix out of HBM
// through per-warp self-refilling cp.async.bulk rings (the MMVQ kernels' structure), as a function of slot size, ring
// depth and warps per CTA -- against a plain coalesced LDG.128 stream of the same bytes?  No math, no dependency chain,
// ONE launch over 8 x 33 MB (> 2 x L2): the steady-state ceiling of the load structure itself.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/probes/stream_probe scripts/probes/stream_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// every CTA owns a contiguous byte range; warp w streams slots w, w + NW, ... of it through its private ring
__global__ void tma_ring_kernel(const uint8_t *src, size_t total, int slot_bytes, int depth, unsigned *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *slots = smem + 1024;
    if (threadIdx.x < nw * depth) mbar_init(&bars[threadIdx.x], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t per_cta = total / gridDim.x / slot_bytes * slot_bytes;
    const uint8_t *base = src + (size_t)blockIdx.x * per_cta;
    const int nslots_cta = (int)(per_cta / slot_bytes);
    const int mine = (nslots_cta - warp + nw - 1) / nw;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int issued = 0;
    auto issue = [&]() {
        const int r = warp * depth + issued % depth;
        mbar_expect_tx(&bars[r], (uint32_t)slot_bytes);
        bulk_g2s(slots + (size_t)r * slot_bytes, base + (size_t)(warp + (size_t)issued * nw) * slot_bytes, (uint32_t)slot_bytes, &bars[r], pol);
        issued++;
    };
    if (lane == 0) while (issued < min(depth, mine)) issue();
    unsigned acc = 0;
    for (int k = 0; k < mine; k++) {
        const int r = warp * depth + k % depth;
        mbar_wait(&bars[r], (uint32_t)(k / depth) & 1u);
        acc ^= *reinterpret_cast<const unsigned *>(slots + (size_t)r * slot_bytes + (lane * 16) % slot_bytes);
        __syncwarp();
        if (lane == 0 && issued < mine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue();
        }
    }
    if (acc == 0x12345u) sink[0] = acc;
}

// plain coalesced stream: every thread 16 bytes per load, UNR loads in flight
template <int UNR>
__global__ void ldg_kernel(const uint4 *src, size_t n16, unsigned *sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0;
    for (; i + (UNR - 1) * stride < n16; i += UNR * stride) {
        uint4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNR; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345u) sink[0] = acc;
}


// per-warp LDGSTS (cp.async.cg 16 B per lane, coalesced) rings: warp w streams chunks w, w + NW, ... of its CTA's contiguous
// range; DEPTH chunks in flight per warp; optionally the chunk `l2_ahead` steps further is pulled into L2 (prefetch.global.L2)
template <int DEPTH>
__global__ void ldgsts_ring_kernel(const uint8_t *src, size_t total, int chunk_bytes, int l2_ahead, unsigned *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_cta = total / gridDim.x / chunk_bytes * chunk_bytes;
    const uint8_t *base = src + (size_t)blockIdx.x * per_cta;
    const int nchunks = (int)(per_cta / chunk_bytes);
    const int mine = (nchunks - warp + nw - 1) / nw;
    const int n16 = chunk_bytes / 16;
    const uint32_t ring = smem_u32(smem) + warp * DEPTH * chunk_bytes + lane * 16;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int issued = 0;
    auto issue = [&]() {
        if (issued < mine) {
            const uint8_t *g = base + (size_t)(warp + (size_t)issued * nw) * chunk_bytes + lane * 16;
            const uint32_t d = ring + (issued % DEPTH) * chunk_bytes;
            for (int q = 0; lane + 32 * q < n16; q++)
                asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d + 512 * q), "l"(g + 512 * q), "l"(pol) : "memory");
            if (l2_ahead > 0 && issued + l2_ahead < mine && lane * 128 < chunk_bytes)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(warp + (size_t)(issued + l2_ahead) * nw) * chunk_bytes + lane * 128));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        issued++;
    };
    for (int s = 0; s < DEPTH; s++) issue();
    unsigned acc = 0;
    for (int k = 0; k < mine; k++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        __syncwarp();
        acc ^= *reinterpret_cast<const unsigned *>(smem + (size_t)(warp * DEPTH + k % DEPTH) * chunk_bytes + (lane * 16) % chunk_bytes);
        __syncwarp();
        issue();
    }
    if (acc == 0x12345u) sink[0] = acc;
}

// plain coalesced LDG.128 stream with a bounded number of loads in flight per thread (bandwidth vs bytes in flight)
template <int UNR>
__global__ void ldg_bounded_kernel(const uint4 *src, size_t n16, unsigned *sink) {
    const size_t per_cta = n16 / gridDim.x;
    const uint4 *p = src + (size_t)blockIdx.x * per_cta;
    unsigned acc = 0;
    for (size_t i = threadIdx.x; i + (size_t)(UNR - 1) * blockDim.x < per_cta; i += (size_t)UNR * blockDim.x) {
        uint4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) v[u] = __ldcs(p + i + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < UNR; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345u) sink[0] = acc;
}

int main() {
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t total = (size_t)8 * 33030144;
    uint8_t *buf;
    unsigned *sink;
    CK(cudaMalloc(&buf, total));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(buf, 1, total));
    CK(cudaFuncSetAttribute(tma_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](auto launch) {
        launch();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        return (double)total / (best * 1e-3) / 1e9;
    };
    printf("{\"bytes\": %zu, \"ldg\": [", total);
    {
        const double a = run([&] { ldg_kernel<4><<<sms * 4, 512>>>((const uint4 *)buf, total / 16, sink); });
        const double b = run([&] { ldg_kernel<8><<<sms * 4, 512>>>((const uint4 *)buf, total / 16, sink); });
        const double c = run([&] { ldg_kernel<8><<<sms * 2, 1024>>>((const uint4 *)buf, total / 16, sink); });
        printf("{\"unroll\": 4, \"ctas_per_sm\": 4, \"GBps\": %.0f}, {\"unroll\": 8, \"ctas_per_sm\": 4, \"GBps\": %.0f}, {\"unroll\": 8, \"threads\": 1024, \"ctas_per_sm\": 2, \"GBps\": %.0f}", a, b, c);
    }
    printf("],\n \"tma_ring\": [\n");
    const int slot_sizes[] = {1152, 2304, 4608, 9216, 18432};
    const int warps[] = {4, 8, 12, 16};
    const int ctas_per_sm[] = {1, 2};
    bool first = true;
    for (int cps : ctas_per_sm)
        for (int nw : warps)
            for (int sb : slot_sizes) {
                const int budget = (cps == 1 ? 200 : 100) * 1024;
                int depth = budget / (nw * sb);
                if (depth < 2) continue;
                if (depth > 8) depth = 8;
                if (nw * depth > 128) continue;
                const int smem = 1024 + nw * depth * sb;
                const double g = run([&] { tma_ring_kernel<<<sms * cps, nw * 32, smem>>>(buf, total, sb, depth, sink); });
                printf("%s  {\"ctas_per_sm\": %d, \"warps\": %d, \"slot_bytes\": %d, \"depth\": %d, \"in_flight_KB_per_sm\": %d, \"GBps\": %.0f}",
                       first ? "" : ",\n", cps, nw, sb, depth, cps * nw * depth * sb / 1024, g);
                first = false;
            }

    printf("\n ],\n \"ldgsts_ring\": [\n");
    {
        bool f2 = true;
        auto one = [&](auto kern, int depth, int cps, int nw, int cb, int l2a) {
            const int smem = nw * depth * cb;
            if (smem > (cps == 1 ? 220 : 110) * 1024) return;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            const double g = run([&] { kern<<<sms * cps, nw * 32, smem>>>(buf, total, cb, l2a, sink); });
            printf("%s  {\"ctas_per_sm\": %d, \"warps\": %d, \"chunk_bytes\": %d, \"depth\": %d, \"l2_ahead\": %d, \"in_flight_KB_per_sm\": %d, \"GBps\": %.0f}",
                   f2 ? "" : ",\n", cps, nw, cb, depth, l2a, cps * smem / 1024, g);
            f2 = false;
        };
        for (int cps : {1, 2})
            for (int nw : {8, 12, 16})
                for (int cb : {2304, 4608})
                    for (int l2a : {0, 4, 12}) {
                        one(ldgsts_ring_kernel<2>, 2, cps, nw, cb, l2a);
                        one(ldgsts_ring_kernel<3>, 3, cps, nw, cb, l2a);
                        one(ldgsts_ring_kernel<4>, 4, cps, nw, cb, l2a);
                        one(ldgsts_ring_kernel<6>, 6, cps, nw, cb, l2a);
                        one(ldgsts_ring_kernel<8>, 8, cps, nw, cb, l2a);
                    }
    }
    printf("\n ],\n \"ldg_bounded\": [\n");
    {
        bool f3 = true;
        auto one = [&](auto kern, int unr, int cps, int thr) {
            const double g = run([&] { kern<<<sms * cps, thr>>>((const uint4 *)buf, total / 16, sink); });
            printf("%s  {\"ctas_per_sm\": %d, \"threads\": %d, \"unroll\": %d, \"in_flight_KB_per_sm\": %d, \"GBps\": %.0f}", f3 ? "" : ",\n", cps, thr, unr,
                   cps * thr * unr * 16 / 1024, g);
            f3 = false;
        };
        for (int cps : {1, 2})
            for (int thr : {256, 384, 512}) {
                one(ldg_bounded_kernel<2>, 2, cps, thr);
                one(ldg_bounded_kernel<4>, 4, cps, thr);
                one(ldg_bounded_kernel<8>, 8, cps, thr);
                one(ldg_bounded_kernel<16>, 16, cps, thr);
            }
    }
    printf("\n ]}\n");
    return 0;
}

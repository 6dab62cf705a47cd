// Microbenchmark: how many bytes per second can all SMs pull from an L2-resident buffer into shared memory
// (a) with the bulk-copy engine (cp.async.bulk), (b) with LDG.128 + STS.  Sets the ceiling for the GEMM's operand feed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fabric_bench fabric_bench.cu && ./fabric_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../ophelia_b200/csrc/oph_ptx.cuh"
using namespace oph;

constexpr int SLOT = 32768, NSLOT = 6;

__global__ void __launch_bounds__(128, 1) bulk_kernel(const uint8_t* src, size_t src_bytes, int iters, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSLOT * SLOT);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init(); fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t nchunks = src_bytes / SLOT;
        size_t c = (blockIdx.x * 7) % nchunks;
        for (int i = 0; i < NSLOT; ++i) {
            mbar_arrive_expect_tx(smem_u32(bars + i), SLOT);
            bulk_g2s(smem_u32(smem + i * SLOT), src + c * SLOT, SLOT, smem_u32(bars + i));
            c = (c + 1) % nchunks;
        }
        for (int it = 0; it < iters; ++it) {
            const int s = it % NSLOT;
            mbar_wait(smem_u32(bars + s), (it / NSLOT) & 1);
            if (it + NSLOT < iters) {
                mbar_arrive_expect_tx(smem_u32(bars + s), SLOT);
                bulk_g2s(smem_u32(smem + s * SLOT), src + c * SLOT, SLOT, smem_u32(bars + s));
                c = (c + 1) % nchunks;
            }
        }
        sink[blockIdx.x] = smem[5];
    }
}

__global__ void __launch_bounds__(256, 1) ldg_kernel(const float4* src, size_t n4, int iters, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float4* s4 = reinterpret_cast<float4*>(smem);
    float acc = 0.f;
    size_t base = ((size_t)blockIdx.x * 4099) % (n4 - 8192);
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + base + j * 256 + threadIdx.x);
#pragma unroll
        for (int j = 0; j < 8; ++j) s4[j * 256 + threadIdx.x] = v[j];
        base = (base + 2048) % (n4 - 8192);
        acc += v[0].x;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const size_t bytes = 24u << 20;   // L2 resident
    uint8_t* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    unsigned long long* sink; cudaMalloc(&sink, 1024 * 8);
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NSLOT * SLOT + 256);
    cudaFuncSetAttribute(ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148, 74, 296}) {
        const int iters = 400;
        bulk_kernel<<<grid, 128, NSLOT * SLOT + 256>>>(d, bytes, iters, sink);
        cudaEventRecord(e0);
        bulk_kernel<<<grid, 128, NSLOT * SLOT + 256>>>(d, bytes, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("bulk copy: grid %d: %.2f TB/s (%s)\n", grid, (double)grid * iters * SLOT / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    for (int grid : {148, 296}) {
        const int iters = 400;
        ldg_kernel<<<grid, 256, 65536>>>((const float4*)d, bytes / 16, iters, (float*)sink);
        cudaEventRecord(e0);
        ldg_kernel<<<grid, 256, 65536>>>((const float4*)d, bytes / 16, iters, (float*)sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("ldg+sts: grid %d: %.2f TB/s (%s)\n", grid, (double)grid * iters * 32768 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

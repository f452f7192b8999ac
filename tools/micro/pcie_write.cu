// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 tools/micro/pcie_write.cu -o tools/micro/pcie_write   (results: profiles/experiments/r01_pool_kernel.txt)
// micro-benchmark: SM-driven writes to pinned host memory, 32-byte records at scattered positions
//   mode 0: one 32 B record per thread, random record positions            (what a retiring ray does)
//   mode 1: one 32 B record per thread, consecutive threads = consecutive records (fully coalesced)
//   mode 2: one thread writes 4 consecutive records (128 B line), random line positions
//   mode 3: 4 consecutive lanes write the 4 records of a random line
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(void* p, uint32_t v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t h) { h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16; return h; }
__global__ void k(char* dst, uint32_t n_rec, int mode, uint32_t salt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 0) { if (i < n_rec) st256(dst + (size_t)(hash32(i ^ salt) % n_rec) * 32, i); }
    else if (mode == 1) { if (i < n_rec) st256(dst + (size_t)i * 32, i); }
    else if (mode == 2) { uint32_t nl = n_rec / 4; if (i < nl) { char* p = dst + (size_t)(hash32(i ^ salt) % nl) * 128; st256(p, i); st256(p + 32, i); st256(p + 64, i); st256(p + 96, i); } }
    else { uint32_t nl = n_rec / 4; if (i < n_rec) { char* p = dst + (size_t)(hash32((i >> 2) ^ salt) % nl) * 128 + (i & 3) * 32; st256(p, i); } }
}
int main() {
    const uint32_t n = 2073600;   // records of a 1080p frame
    char* h; cudaHostAlloc(&h, (size_t)n * 32, cudaHostAllocMapped);
    char* d; cudaHostGetDevicePointer(&d, h, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 4; ++mode) {
        uint32_t threads = mode == 2 ? n / 4 : n;
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(a);
            k<<<(threads + 127) / 128, 128>>>(d, n, mode, rep * 977u);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
        }
        printf("mode %d: %.3f ms  %.1f GB/s  (%.0f Mrecords/s)\n", mode, best, n * 32.0 / best / 1e6, n / best / 1e3);
    }
    return 0;
}

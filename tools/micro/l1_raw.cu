// Micro-benchmark (GPU box): latency of dependent global loads by one thread, with and without a store to the same
// line in between - does a global store keep / update / invalidate the line in L1 on sm_100?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/l1_raw tools/micro/l1_raw.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chase(uint32_t *a, uint32_t n, uint32_t steps, int mode, long long *out, uint32_t *sink)
{
    extern __shared__ uint32_t sm[];
    if (mode == 3)
        for (uint32_t i = 0; i < n; ++i)
            sm[i] = a[i];
    uint32_t *p = mode == 3 ? sm : a;
    uint32_t i = 0;
    // warm the lines
    for (uint32_t s = 0; s < n; ++s)
        i = p[i];
    const long long t0 = clock64();
    for (uint32_t s = 0; s < steps; ++s)
    {
        const uint32_t nx = p[i];
        if (mode == 1)       // store to the SAME word that the next-but-one hop reads: read-after-write through L1
            p[nx] = p[nx];   // (a load + a store of the same value; keeps the chain intact)
        else if (mode == 2)  // store to another word of the same line as the next hop
            p[nx ^ 1u] = s;
        i = nx;
    }
    const long long t1 = clock64();
    out[0] = t1 - t0;
    sink[0] = i;
}

int main()
{
    for (uint32_t n : {1024u, 16384u, 262144u, 4194304u})
    {
        uint32_t *h = new uint32_t[n];
        // a permutation cycle with stride 2*k+... : even slots only so that slot^1 is free
        const uint32_t m = n / 2;
        uint32_t stride = 0x9E37u | 1u;
        for (uint32_t j = 0; j < m; ++j)
            h[2 * j] = 2 * ((j * 1ull * stride + 12345u) % m), h[2 * j + 1] = 0;
        // make it a single chain: next[j] = (j + stride) mod m over even slots
        for (uint32_t j = 0; j < m; ++j)
            h[2 * j] = 2 * ((j + stride) % m);
        uint32_t *d, *sink;
        long long *out;
        cudaMalloc(&d, n * 4);
        cudaMalloc(&sink, 4);
        cudaMalloc(&out, 8);
        for (int mode = 0; mode < 4; ++mode)
        {
            if (mode == 3 && n * 4 > 200 * 1024)
                continue;
            cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
            if (mode == 3)
                cudaFuncSetAttribute(chase, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            const uint32_t steps = 20000;
            chase<<<1, 1, mode == 3 ? n * 4 : 0>>>(d, n, steps, mode, out, sink);
            long long c;
            cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            printf("array %8u B  mode %d (%s): %.1f cycles per hop  (%s)\n", n * 4, mode,
                   mode == 0 ? "load only" : mode == 1 ? "load + store same word" : mode == 2 ? "load + store same line" : "shared memory",
                   double(c) / steps, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(d);
        delete[] h;
    }
    return 0;
}

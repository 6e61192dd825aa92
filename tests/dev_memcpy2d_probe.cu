// Developer probe: rate of strided (AoS member-run) copies between a pinned host AoS buffer and the device:
// cudaMemcpy2DAsync with tiny rows vs. whole-record DMA.  nvcc -O3 tests/dev_memcpy2d_probe.cu -o tests/dev_memcpy2d_probe.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)
int main()
{
    const size_t n = 16u << 20, rec = 208;
    char * h = nullptr, * d = nullptr;
    CK(cudaMallocHost(&h, n * rec));
    CK(cudaMalloc(&d, n * rec));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int dir = 0; dir < 2; ++dir) {
        for (size_t w : {8, 16, 24, 32, 56, 104, 208}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (dir == 0) CK(cudaMemcpy2DAsync(h + 24, rec, d, w, w, n, cudaMemcpyDeviceToHost, 0));
                else CK(cudaMemcpy2DAsync(d, w, h + 24, rec, w, n, cudaMemcpyHostToDevice, 0));
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                cudaEventElapsedTime(&ms, a, b);
            }
            printf("%s width %3zu B, pitch 208, %zu rows: %8.2f ms  payload %6.2f GB/s\n", dir == 0 ? "D2H" : "H2D", w, n, ms, n * w / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}

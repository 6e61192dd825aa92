// Developer probe (not part of the product): can two PROCESSES on one box map each other's cudaMalloc
// memory with CUDA IPC and read it from a kernel over NVLink?  Measures the bandwidth of a record pull
// (32-byte records, contiguous runs of 16) and of a flat copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tests/dev_ipc_probe.cu -o gpurun_out/ipc_probe && gpurun_out/ipc_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("rank %d: %s -> %s\n", rank, #x, cudaGetErrorString(e)); exit(2); } } while (0)

__global__ void k_fill(double4 * a, size_t n, double v) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) a[i] = make_double4(v, (double)i, 0, 0); }
// one warp per run of 16 records: lanes 0..15 copy one 32-byte record each (two runs per warp)
__global__ void k_pull(const double4 * __restrict__ src, double4 * __restrict__ dst, const int * __restrict__ runs, int nruns)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = t >> 4, k = t & 15;
    if (r >= nruns) return;
    const size_t i = (size_t)runs[r] * 16 + k;
    dst[i] = src[i];
}
__global__ void k_copy(const double4 * __restrict__ src, double4 * __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_check(const double4 * a, size_t n, double v, int * bad) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n && (a[i].x != v || a[i].y != (double)i)) atomicAdd(bad, 1); }

int main()
{
    int p01[2], p10[2];
    if (pipe(p01) || pipe(p10)) return 1;
    pid_t pid = fork();
    const int rank = pid == 0 ? 1 : 0;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { if (rank == 0) printf("needs 2 GPUs, have %d\n", ndev); return 0; }
    CK(cudaSetDevice(rank));
    const size_t n = 16u << 20;                     // 16 M records = 512 MB
    double4 * mine = nullptr, * local = nullptr;
    CK(cudaMalloc(&mine, n * sizeof(double4)));
    CK(cudaMalloc(&local, n * sizeof(double4)));
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(mine, n, 100.0 + rank);
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h, hp;
    CK(cudaIpcGetMemHandle(&h, mine));
    const int wfd = rank == 0 ? p01[1] : p10[1], rfd = rank == 0 ? p10[0] : p01[0];
    if (write(wfd, &h, sizeof(h)) != sizeof(h)) return 1;
    if (read(rfd, &hp, sizeof(hp)) != sizeof(hp)) return 1;
    double4 * peer = nullptr;
    CK(cudaIpcOpenMemHandle((void **)&peer, hp, cudaIpcMemLazyEnablePeerAccess));
    int can = 0;
    cudaDeviceCanAccessPeer(&can, rank, 1 - rank);
    // runs: every 4th run of 16 records (a 25 % halo, scattered)
    const int nruns = (int)(n / 16 / 4);
    int * hr = (int *)malloc(nruns * sizeof(int)), * dr = nullptr;
    for (int k = 0; k < nruns; ++k) hr[k] = 4 * k + (k % 3);
    CK(cudaMalloc(&dr, nruns * sizeof(int)));
    CK(cudaMemcpy(dr, hr, nruns * sizeof(int), cudaMemcpyHostToDevice));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k_pull<<<(nruns * 16 + 255) / 256, 256>>>(peer, local, dr, nruns);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
    }
    const double pull_gb = (double)nruns * 16 * 32 / 1e9;
    printf("rank %d: canAccessPeer=%d pull of %.1f MB in runs of 512 B: %.3f ms = %.1f GB/s\n", rank, can, pull_gb * 1e3, ms, pull_gb / (ms * 1e-3));
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k_copy<<<148 * 8, 256>>>(peer, local, n);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
    }
    printf("rank %d: flat peer copy of 512 MB: %.3f ms = %.1f GB/s\n", rank, ms, 0.512 * 1.048576 / (ms * 1e-3));
    int * bad = nullptr, hb = 0;
    CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    k_check<<<(unsigned)((n + 255) / 256), 256>>>(local, n, 100.0 + (1 - rank), bad);
    CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    printf("rank %d: peer data check bad=%d\n", rank, hb);
    // keep the exporter alive until the peer is done
    char c = 1;
    if (write(wfd, &c, 1) != 1) return 1;
    if (read(rfd, &c, 1) != 1) return 1;
    CK(cudaIpcCloseMemHandle(peer));
    if (rank == 0) { int st; waitpid(pid, &st, 0); }
    return hb != 0;
}

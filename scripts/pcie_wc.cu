// PCIe probe: pinned vs write-combined pinned host buffers, H2D alone / D2H alone / both directions (diagnostic)
//   nvcc -O2 -o scripts/pcie_wc scripts/pcie_wc.cu && scripts/pcie_wc
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <chrono>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t n = (size_t)256 * 3508 * 2480;      // one bench step of pages
    void *d_in, *d_out;
    cudaMalloc(&d_in, n); cudaMalloc(&d_out, n);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    for (int wc = 0; wc < 2; ++wc) {
        void *h_in, *h_out;
        cudaHostAlloc(&h_in, n, wc ? (cudaHostAllocPortable | cudaHostAllocWriteCombined) : cudaHostAllocPortable);
        cudaHostAlloc(&h_out, n, cudaHostAllocPortable);
        memset(h_in, 7, n); memset(h_out, 0, n);
        for (int mode = 0; mode < 3; ++mode) {
            double best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaDeviceSynchronize();
                const double t0 = now();
                if (mode != 1) cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, s1);
                if (mode != 0) cudaMemcpyAsync(h_out, d_out, n, cudaMemcpyDeviceToHost, s2);
                cudaDeviceSynchronize();
                const double dt = now() - t0;
                if (dt < best) best = dt;
            }
            printf("%s input buffer, %s: %.1f GB/s per direction\n", wc ? "write-combined" : "pinned        ",
                   mode == 0 ? "H2D alone" : mode == 1 ? "D2H alone" : "both     ", n / best / 1e9);
        }
        cudaFreeHost(h_in); cudaFreeHost(h_out);
    }
    return 0;
}

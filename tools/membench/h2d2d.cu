// h2d2d.cu -- host-to-device copy rate of row-pitched (2-D) copies with small rows vs one flat copy.  Diagnostic tool.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
int main(int argc, char **argv) {
    const size_t n = argc > 1 ? atol(argv[1]) : 60000000;
    const size_t sp = argc > 2 ? atol(argv[2]) : 28, dp = argc > 3 ? atol(argv[3]) : 32;
    unsigned char *h, *d;
    cudaMallocHost(&h, n * dp); cudaMalloc(&d, n * dp);
    memset(h, 65, n * dp);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) cudaMemcpyAsync(d, h, n * dp, cudaMemcpyHostToDevice);
            else cudaMemcpy2DAsync(d, dp, h, sp, sp, n, cudaMemcpyHostToDevice);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double bytes = mode == 0 ? (double)n * dp : (double)n * sp;
        printf("{\"mode\": \"%s\", \"rows\": %zu, \"ms\": %.2f, \"GB_per_s_payload\": %.1f, \"rows_per_s\": %.3g}\n",
               mode == 0 ? "flat dp-byte rows" : "2D sp->dp", n, best, bytes / best / 1e6, n / (best * 1e-3));
    }
    return cudaGetLastError() != cudaSuccess;
}

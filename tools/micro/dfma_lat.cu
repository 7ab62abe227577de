// Microbenchmark (development aid): FP64 DFMA dependent-issue latency and throughput per SM sub-partition on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_lat dfma_lat.cu && ./dfma_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a + c + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
void run(int warps_per_sm, double* out, long long* cyc) {
    int iters = 2000;
    k<CHAINS><<<148, 32 * warps_per_sm>>>(out, iters, 1.0, 0.999, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 16.0);
    printf("chains=%d warps/SM=%2d (%.1f/SMSP): %.2f cycles per step of %d DFMA/warp -> %.2f cycles per warp-DFMA per SMSP\n", CHAINS, warps_per_sm,
           warps_per_sm / 4.0, per, CHAINS, per / (CHAINS * warps_per_sm / 4.0));
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 12, 16, 24, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc); run<4>(w, out, cyc); run<6>(w, out, cyc); run<8>(w, out, cyc); }
    return 0;
}

// fp64_latency.cu - dependent-issue latency of the FP64 pipe on this GPU (DFMA / DADD / DMUL chains),
// for 1..16 warps per SM with 1..8 independent chains each.  Kernel 9 (heom_dataflow_tma.cuh) is
// bound by exactly this: a few warps per SM walking short dependent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) acc[i] = fma(acc[i], a, b);
            if (OP == 1) acc[i] = acc[i] + b;
            if (OP == 2) acc[i] = acc[i] * a;
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP>
void run(const char* name, int warps, double* out, long long* cyc) {
    const int iters = 2048;
    chain<ILP, OP><<<1, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%s warps/SM %2d chains/thread %d : %.1f cycles per dependent step, %.2f warp-instr/clk/SM\n", name, warps, ILP,
           (double)c / iters, (double)warps * ILP * iters / c);
}
int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024);
    cudaMalloc(&cyc, sizeof(long long));
    for (int w : {1, 4, 16, 32}) {
        run<1, 0>("DFMA", w, out, cyc);
        run<2, 0>("DFMA", w, out, cyc);
        run<4, 0>("DFMA", w, out, cyc);
        run<8, 0>("DFMA", w, out, cyc);
    }
    run<1, 1>("DADD", 1, out, cyc);
    run<1, 2>("DMUL", 1, out, cyc);
    run<4, 1>("DADD", 16, out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

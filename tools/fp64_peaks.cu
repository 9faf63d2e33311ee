// fp64_peaks.cu - FP64 throughput of one B200: plain DFMA against the FP64 tensor-core path
// (mma.sync ... f64, "DMMA"; tcgen05 has no FP64 type).  Decides whether the N = 32 dense
// commutator (BASELINE configs[3]) should go through DMMA.  Build + run: tools/fp64_peaks.sh
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 8x4 (1 double / lane), B 4x8 (1 / lane), C 8x8 (2 / lane); 2*8*8*4 = 512 flop per warp instruction
template <int ILP>
__global__ void dmma884_kernel(double* out, double a, double b, int iters) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: A 16x8 (4 / lane), B 8x8 (2 / lane), C 16x8 (4 / lane); 2*16*8*8 = 2048 flop per warp instruction
template <int ILP>
__global__ void dmma1688_kernel(double* out, double a, double b, int iters) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, threads = 256, ctas = sms * 8, iters = 4096;
    double* out;
    CHECK(cudaMalloc(&out, sizeof(double) * ctas * threads));
    constexpr int ILP = 8;
    const double warps = (double)ctas * threads / 32;
    const double t_fma = time_ms([&] { dfma_kernel<ILP><<<ctas, threads>>>(out, 1.0000001, 1e-9, iters); });
    const double t_884 = time_ms([&] { dmma884_kernel<ILP><<<ctas, threads>>>(out, 1.0000001, 1e-9, iters); });
    const double t_1688 = time_ms([&] { dmma1688_kernel<ILP><<<ctas, threads>>>(out, 1.0000001, 1e-9, iters); });
    CHECK(cudaGetLastError());
    const double f_fma = warps * 32 * ILP * 2.0 * iters / (t_fma * 1e-3) / 1e12;
    const double f_884 = warps * ILP * 512.0 * iters / (t_884 * 1e-3) / 1e12;
    const double f_1688 = warps * ILP * 2048.0 * iters / (t_1688 * 1e-3) / 1e12;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_nominal\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
           "\"dmma_m16n8k8_tflops\": %.2f, \"how\": \"independent chains (ILP %d), %d CTAs x %d threads, %d iterations, best of 5, CUDA events\"}\n",
           prop.name, sms, clk, f_fma, f_884, f_1688, ILP, ctas, threads, iters);
    return 0;
}

// fp64_microbench.cu — how many FP64 FMAs per clock per SM does this B200 sustain?
// (DESIGN.md budget check: the estimator needs ~20k FP64 ops per 16 KB window.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_fp64_microbench tools/fp64_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int ILP>
__global__ void dfma(double* out, int iters)
{
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static void run(int warps_per_sm)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = warps_per_sm * 32, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma<ILP><<<sms, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma<ILP><<<sms, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)sms * threads * iters * ILP;
    printf("ILP=%d warps/SM=%2d : %.2f TFLOP/s FP64 (%.1f FMA/ns/SM)\n", ILP, warps_per_sm, 2 * fmas / ms / 1e9,
           fmas / ms / 1e6 / sms);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<1>(4); run<2>(4); run<4>(4); run<8>(4);
    run<1>(8); run<4>(8); run<8>(8);
    run<4>(16); run<8>(32);
    return 0;
}

// float_p = double instantiations of the fused estimator kernel.
#include "pmu_launch.cuh"
int pmu_launch_f64(int M, int KC, const KernelArgs& a, int grid, int block, int smem, cudaStream_t s)
{
    return launch_m<double>(M, KC, a, grid, block, smem, s);
}

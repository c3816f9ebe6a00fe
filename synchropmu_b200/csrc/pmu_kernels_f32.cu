// float_p = float instantiations of the fused estimator kernel (float I/O, double arithmetic).
#include "pmu_launch.cuh"
int pmu_launch_f32(int M, int KC, const KernelArgs& a, int grid, int block, int smem, cudaStream_t s)
{
    return launch_m<float>(M, KC, a, grid, block, smem, s);
}

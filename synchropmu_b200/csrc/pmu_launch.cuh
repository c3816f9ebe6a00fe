// pmu_launch.cuh — instantiates pmu_fused_kernel for one float_p width (DFT arithmetic in that width) and
// every (codelet size M, bins per lane KC) specialisation, and dispatches at run time.
// Included by pmu_kernels_f64.cu / pmu_kernels_f32.cu so the two widths compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "pmu_host.hpp"
#include "pmu_kernel.cuh"
#include "pmu_wide.cuh"

#ifndef PMU_F32X2
#define PMU_F32X2 1
#endif

namespace {

template <typename T, int M, int KC, int PITCH, int WP = 1, typename CT = T>
int launch_inst(const KernelArgs& a, int grid, int block, int smem, cudaStream_t s)
{
#if PMU_F32X2
    if constexpr (sizeof(T) == 4 && !IsPacked<CT>::value) {
        // float build: two adjacent residues per lane on the packed FP32 pipe (the scalar instantiation is not compiled)
        return launch_inst<T, M, KC, PITCH, WP, F2>(a, grid, block, smem, s);
    } else
#endif
    {
    // per device: has this instantiation been opted in to the device's maximum dynamic shared memory?  The attribute
    // is only ever set to that one value, so concurrent host threads (batches that share an instantiation but differ
    // in smem_bytes) cannot lower it under each other's launches.
    static std::atomic<int> configured[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire))) {
        int max_optin = smem;
        e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(pmu_fused_kernel<T, CT, M, KC, PITCH, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
        if (e == cudaSuccess && dev >= 0 && dev < 64) configured[dev].store(1, std::memory_order_release);
    }
    if (e == cudaSuccess) {
        // Programmatic dependent launch: when the previous operation in the stream is a kernel, this grid's CTAs may be
        // scheduled (and run their prologue: tables to shared memory, barrier set-up) before that kernel has drained;
        // the kernel itself waits (griddepcontrol.wait) for its complete predecessor before it touches any input,
        // output or state, so ordering is exactly stream order.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)block);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = a.pdl ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, pmu_fused_kernel<T, CT, M, KC, PITCH, WP>, a);
    }
    if (e != cudaSuccess) {
        pmu::set_error(std::string("[pmu_estimate] CUDA ERROR launching the estimator kernel: ") + cudaGetErrorString(e));
        return -1;
    }
    return 0;
    }
}

// The 16-point codelet always runs on stages with a fixed row pitch of PMU_PITCH16 elements (the plan never makes
// its slabs wider), so its sample loads carry immediate offsets.  (A 64-element pitch for 1024- and 3072-sample
// windows was measured and dropped: no gain for 1024, 16 % slower for 3072 - profiles/r1_sweep_smem.jsonl.)
template <typename T, int M, int PITCH, int WP = 1>
int launch_kc(int KC, const KernelArgs& a, int grid, int block, int smem, cudaStream_t s)
{
    switch (KC) {
        case 8: return launch_inst<T, M, 8, PITCH, WP>(a, grid, block, smem, s);
        case 10: return launch_inst<T, M, 10, PITCH, WP>(a, grid, block, smem, s);
        case 12: return launch_inst<T, M, 12, PITCH, WP>(a, grid, block, smem, s);
        case 16: return launch_inst<T, M, 16, PITCH, WP>(a, grid, block, smem, s);
        case 24:
            // odd window lengths (direct DFT, M = 1) with 16..23 bins in double run on the general kernel instead: the
            // 48 accumulators plus 24 twiddle loads of the tiny loop body do not fit the register file without spills
            if constexpr (!(M == 1 && sizeof(T) == 8)) return launch_inst<T, M, 24, PITCH, WP>(a, grid, block, smem, s);
            break;
    }
    pmu::set_error("[pmu_estimate] ERROR: unsupported bin-count specialisation");
    return -1;
}

// The general kernel (pmu_wide.cuh): plans with M == 0.
template <typename T>
int launch_wide(const KernelArgs& a, int grid, int smem, cudaStream_t s)
{
    static std::atomic<int> configured[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire))) {
        int max_optin = smem;
        e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pmu_wide_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
        if (e == cudaSuccess && dev >= 0 && dev < 64) configured[dev].store(1, std::memory_order_release);
    }
    if (e == cudaSuccess) {
        pmu_wide_kernel<T><<<grid, PMU_WIDE_THREADS, smem, s>>>(a);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        pmu::set_error(std::string("[pmu_estimate] CUDA ERROR launching the general estimator kernel: ") + cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

template <typename T>
int launch_m(int M, int KC, const KernelArgs& a, int grid, int block, int smem, cudaStream_t s)
{
    switch (M) {
        case 0: return launch_wide<T>(a, grid, smem, s);
        case 16:
            // 1024-sample windows (64 residues): two per warp pass at their natural pitch
            if (a.wp == 2 && a.pitch == PMU_PITCH16_SHORT) return launch_kc<T, 16, PMU_PITCH16_SHORT, 2>(KC, a, grid, block, smem, s);
            if (a.pitch == PMU_PITCH16_SHORT) return launch_kc<T, 16, PMU_PITCH16_SHORT, 1>(KC, a, grid, block, smem, s);   // its small launches
            if (a.pitch == PMU_PITCH16_96) return launch_kc<T, 16, PMU_PITCH16_96, 1>(KC, a, grid, block, smem, s);
            if constexpr (sizeof(T) == 4) {   // 2048 floats = 8 KB: two windows per 16 KB stage
                if (a.wp == 2 && a.pitch == PMU_PITCH16) return launch_kc<T, 16, PMU_PITCH16, 2>(KC, a, grid, block, smem, s);
            }
            return launch_kc<T, 16, PMU_PITCH16>(KC, a, grid, block, smem, s);
        case 8:
            // short windows: four per warp pass
            if (a.wp == 4) return launch_kc<T, 8, 0, 4>(KC, a, grid, block, smem, s);
            return launch_kc<T, 8, 0>(KC, a, grid, block, smem, s);
        case 4: return launch_kc<T, 4, 0>(KC, a, grid, block, smem, s);
        case 1: return launch_kc<T, 1, 0>(KC, a, grid, block, smem, s);
    }
    pmu::set_error("[pmu_estimate] ERROR: unsupported DFT codelet specialisation");
    return -1;
}

}  // namespace

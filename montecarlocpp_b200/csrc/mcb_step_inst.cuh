// mcb_step_inst.cuh — the k_step instantiations of ONE payload width (NCOMP), compiled as their own translation unit
// (mcb_step_n1.cu / _n3.cu / _n4.cu) so that the 60-odd kernel variants build in parallel.
#pragma once
#include "mcb_kernels.cuh"

namespace mcb {

template <int NCOMP, int TM, int NDM, bool BOX, int PAD>
static cudaError_t launch_step_inst(const StepParams& P, int grid, int block, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(k_step<NCOMP, TM, NDM, BOX, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_step<NCOMP, TM, NDM, BOX, PAD><<<grid, block, smem, s>>>(P);
    return cudaGetLastError();
}
template <int NCOMP, int TM, int NDM>
static cudaError_t launch_step_box(const StepParams& P, int box, int grid, int block, size_t smem, cudaStream_t s) {
    return box ? launch_step_inst<NCOMP, TM, NDM, true, 0>(P, grid, block, smem, s) : launch_step_inst<NCOMP, TM, NDM, false, 0>(P, grid, block, smem, s);
}
template <int NCOMP, int TM>
static cudaError_t launch_step_nd(const StepParams& P, int ndm, int box, int grid, int block, size_t smem, cudaStream_t s) {
    if (ndm == 3) return launch_step_box<NCOMP, TM, 3>(P, box, grid, block, smem, s);
    if (ndm == 2) return launch_step_box<NCOMP, TM, 2>(P, box, grid, block, smem, s);
    if (ndm == 1) return launch_step_box<NCOMP, TM, 1>(P, box, grid, block, smem, s);
    return launch_step_box<NCOMP, TM, 0>(P, box, grid, block, smem, s);
}
// tm / ndm / box / pad as chosen by plan_run (mcb_api.cu).  pad > 0 only with tm == MCB_TM_WARP, ndm == 0, box.
template <int NCOMP>
static cudaError_t launch_step_n(const StepParams& P, int tm, int ndm, int box, int pad, int grid, int block, size_t smem, cudaStream_t s) {
    if (tm == MCB_TM_WARP && ndm == 0 && box) {
        switch (pad) {
        case 32: return launch_step_inst<NCOMP, MCB_TM_WARP, 0, true, 32>(P, grid, block, smem, s);
        case 128: return launch_step_inst<NCOMP, MCB_TM_WARP, 0, true, 128>(P, grid, block, smem, s);
        case 512: return launch_step_inst<NCOMP, MCB_TM_WARP, 0, true, 512>(P, grid, block, smem, s);
        default: return launch_step_inst<NCOMP, MCB_TM_WARP, 0, true, 0>(P, grid, block, smem, s);
        }
    }
    switch (tm) {
    case MCB_TM_WARP: return launch_step_box<NCOMP, MCB_TM_WARP, 0>(P, box, grid, block, smem, s);     // 1-D histograms: no N-D grid (plan_run)
    case MCB_TM_BLOCK: return launch_step_nd<NCOMP, MCB_TM_BLOCK>(P, ndm, box, grid, block, smem, s);
    default: return launch_step_nd<NCOMP, MCB_TM_GLOBAL>(P, ndm, box, grid, block, smem, s);
    }
}

} // namespace mcb

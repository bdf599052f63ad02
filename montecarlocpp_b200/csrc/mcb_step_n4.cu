// k_step for payload width NCOMP = 4 (see mcb_step_inst.cuh)
#include "mcb_step_inst.cuh"
namespace mcb {
cudaError_t launch_step_n4(const StepParams& P, int tm, int ndm, int box, int pad, int grid, int block, size_t smem, cudaStream_t s) {
    return launch_step_n<4>(P, tm, ndm, box, pad, grid, block, smem, s);
}
}

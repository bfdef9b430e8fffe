// mrs_step_mode.cu -- one object file per action mode: nvcc -DMRS_INSTANTIATE_MODE=<MrsActionType> -c.
// The 8 modes x group widths x CTA shapes x {generic, baked} instantiations of step_group_kernel dominate the
// compile time; split per mode they build in parallel (Makefile / __graft_entry__.build()).
#include "mrs_step.cuh"

#ifndef MRS_INSTANTIATE_MODE
#error "compile with -DMRS_INSTANTIATE_MODE=<0..7>"
#endif

namespace mrs {
template int dispatch_step<MRS_INSTANTIATE_MODE>(const MrsConfig&, const MrsBuffers&, StepArgs, cudaStream_t);
}

// nl_stack_linfit.cu -- instantiations of stack_column_kernel for this mode family (see nl_stack_kernel.cuh).
#include "nl_stack_kernel.cuh"

namespace nl {

int launch_linfit(nl_stack_job *job, const StackArgs &args) { return launch_column_s<ST_LINFIT, false>(job, args); }

}  // namespace nl

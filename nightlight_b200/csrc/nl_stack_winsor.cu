// nl_stack_winsor.cu -- instantiations of stack_column_kernel for this mode family (see nl_stack_kernel.cuh).
#include "nl_stack_kernel.cuh"

namespace nl {

int launch_winsor(nl_stack_job *job, const StackArgs &args, bool weighted) {
    return weighted ? launch_column_s<ST_WINSOR, true>(job, args) : launch_column_s<ST_WINSOR, false>(job, args);
}
int launch_mad(nl_stack_job *job, const StackArgs &args) { return launch_column_s<ST_MAD, false>(job, args); }

}  // namespace nl

// nl_stack_sigma.cu -- instantiations of stack_column_kernel for this mode family (see nl_stack_kernel.cuh).
#include "nl_stack_kernel.cuh"

namespace nl {

int launch_median(nl_stack_job *job, const StackArgs &args) { return launch_column_s<ST_MEDIAN, false>(job, args); }
int launch_sigma(nl_stack_job *job, const StackArgs &args, bool weighted) {
    return weighted ? launch_column_s<ST_SIGMA, true>(job, args) : launch_column_s<ST_SIGMA, false>(job, args);
}

}  // namespace nl

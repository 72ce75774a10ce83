// Small-column-tile instantiations of the FP32 forward-solve and adjoint kernels (see train_tile in cpz_k_solve.cu).
// One translation unit per tile width (cpz_k_small.cu: 4, cpz_k_small8.cu: 8, cpz_k_small16.cu: 16) so that they compile
// in parallel; this file also holds the dispatch.
#include "cpz_small_impl.h"

namespace cpz {

CPZ_SMALL_DEFINE(4)

int launch_solve_small(cpz_model* m, const SolveArgs& a, int CT) {
  switch (CT) {
    case 4: return solve_small_4(m, a);
    case 8: return solve_small_8(m, a);
    case 16: return solve_small_16(m, a);
    default: return fail(CPZ_ERR_INVALID, "no forward kernel for %d-column tiles", CT);
  }
}
int launch_adjoint_small(cpz_model* m, const AdjArgs& a, int grid, int CT) {
  switch (CT) {
    case 4: return adjoint_small_4(m, a, grid);
    case 8: return adjoint_small_8(m, a, grid);
    case 16: return adjoint_small_16(m, a, grid);
    default: return fail(CPZ_ERR_INVALID, "no adjoint kernel for %d-column tiles", CT);
  }
}

}  // namespace cpz

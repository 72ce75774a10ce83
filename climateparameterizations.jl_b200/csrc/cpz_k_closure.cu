// Closure kernel instantiations and launcher.
#include <algorithm>

#include "cpz_launch.h"

namespace cpz {

template <int CT, int NT, bool WS>
static int launch_closure_t(cpz_model* m, const ClosureD& cd, const ClosureArgs& a) {
  const ClosureSmem L = closure_smem_layout(m->fwd.M, CT);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "closure kernel needs %zu B shared memory", smem);
  auto kern = closure_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(a.n_tiles, m->ctx->sm_count);
  kern<<<grid, NT, smem, m->ctx->stream>>>(m->fwd.M, cd, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

int launch_closure(cpz_model* m, const ClosureD& cd, const ClosureArgs& a) {
  if (m->fwd.M.w_in_smem) return launch_closure_t<32, 256, true>(m, cd, a);
  return launch_closure_t<32, 256, false>(m, cd, a);
}

}  // namespace cpz

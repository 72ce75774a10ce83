// CT_SMALL-column-tile instantiations of the FP32 forward-solve and adjoint kernels (see train_tile in cpz_k_solve.cu).
#include <cstdlib>

#include "cpz_launch.h"

namespace cpz {

template <bool WS>
static int solve_small_t(cpz_model* m, const SolveArgs& a) {
  constexpr int CT = cpz_model::CT_SMALL, NT = 256;
  const SolveSmem L = solve_smem_layout(m->fwd_s.M, CT, m->tab.n_stages);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "small-tile forward kernel needs %zu B shared memory, device allows %zu", smem, m->ctx->smem_optin);
  auto kern = solve_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(a.ncol + CT - 1) / CT, NT, smem, m->ctx->stream>>>(m->fwd_s.M, m->tab, m->tm, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

int launch_solve_small(cpz_model* m, const SolveArgs& a) {
  return m->fwd_s.M.w_in_smem ? solve_small_t<true>(m, a) : solve_small_t<false>(m, a);
}

template <bool WS>
static int adjoint_small_t(cpz_model* m, const AdjArgs& a, int grid) {
  constexpr int CT = cpz_model::CT_SMALL, NT = 256;
  const AdjSmem L = adjoint_smem_layout(m->bwd_s.M, CT);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "small-tile adjoint kernel needs %zu B shared memory, device allows %zu", smem, m->ctx->smem_optin);
  auto kern = adjoint_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NT, smem, m->ctx->stream>>>(m->bwd_s.M, m->tab, m->tm, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

int launch_adjoint_small(cpz_model* m, const AdjArgs& a, int grid) {
  return m->bwd_s.M.w_in_smem ? adjoint_small_t<true>(m, a, grid) : adjoint_small_t<false>(m, a, grid);
}

}  // namespace cpz

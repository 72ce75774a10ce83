// Launch helpers of the small-column-tile kernels, instantiated per tile width by CPZ_SMALL_DEFINE(CT).
#pragma once
#include "cpz_adjoint_launch.h"

namespace cpz {

int solve_small_4(cpz_model* m, const SolveArgs& a);
int solve_small_8(cpz_model* m, const SolveArgs& a);
int solve_small_16(cpz_model* m, const SolveArgs& a);
int adjoint_small_4(cpz_model* m, const AdjArgs& a, int grid);
int adjoint_small_8(cpz_model* m, const AdjArgs& a, int grid);
int adjoint_small_16(cpz_model* m, const AdjArgs& a, int grid);

constexpr int small_index(int CT) { return CT == 4 ? 0 : (CT == 8 ? 1 : 2); }

template <int CT, bool WS>
static int solve_small_t(cpz_model* m, const SolveArgs& a) {
  constexpr int NT = 256;
  const Plan& plan = m->fwd_s[small_index(CT)];
  const SolveSmem L = solve_smem_layout(plan.M, CT, m->tab.n_stages);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "small-tile forward kernel needs %zu B shared memory, device allows %zu", smem, m->ctx->smem_optin);
  auto kern = solve_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(a.ncol + CT - 1) / CT, NT, smem, m->ctx->stream>>>(plan.M, m->tab, m->tm, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

#define CPZ_SMALL_DEFINE(CT)                                                                                              \
  int solve_small_##CT(cpz_model* m, const SolveArgs& a) {                                                                \
    if (!m->has_small[small_index(CT)]) return fail(CPZ_ERR_INVALID, "no plan for %d-column tiles", CT);                  \
    return m->fwd_s[small_index(CT)].M.w_in_smem ? solve_small_t<CT, true>(m, a) : solve_small_t<CT, false>(m, a);         \
  }                                                                                                                       \
  int adjoint_small_##CT(cpz_model* m, const AdjArgs& a, int grid) {                                                      \
    if (!m->has_small[small_index(CT)]) return fail(CPZ_ERR_INVALID, "no plan for %d-column tiles", CT);                  \
    const Plan& plan = m->bwd_s[small_index(CT)];                                                                          \
    return plan.M.w_in_smem ? launch_adjoint_t<CT, 256, true>(m, plan, a, grid) : launch_adjoint_t<CT, 256, false>(m, plan, a, grid); \
  }

}  // namespace cpz

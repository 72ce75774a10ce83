// Forward-solve kernel instantiations and launcher.
#include <cstdlib>

#include "cpz_launch.h"

namespace cpz {

// ---- forward launch ---------------------------------------------------------------------------------------------
template <int CT, int NT, bool WS>
static int launch_solve_t(cpz_model* m, const SolveArgs& a) {
  const SolveSmem L = solve_smem_layout(m->fwd.M, CT, m->tab.n_stages);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "forward kernel needs %zu B shared memory, device allows %zu", smem, m->ctx->smem_optin);
  auto kern = solve_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_tiles = (a.ncol + CT - 1) / CT;
  static const bool prof = getenv("CPZ_PROF") != nullptr;
  if (prof && !a.rhs_only) {  // debug: per-phase cycle counters of CTA 0
    SolveArgs ap = a;
    unsigned long long* d = nullptr;
    CPZ_CUDA(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
    CPZ_CUDA(cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), m->ctx->stream));
    ap.prof = d;
    kern<<<n_tiles, NT, smem, m->ctx->stream>>>(m->fwd.M, m->tab, m->tm, ap);
    unsigned long long hcnt[8];
    CPZ_CUDA(cudaMemcpyAsync(hcnt, d, sizeof(hcnt), cudaMemcpyDeviceToHost, m->ctx->stream));
    CPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
    cudaFree(d);
    const double n_rhs = (double)m->tm.n_steps * m->tm.n_substeps * m->tab.n_stages;
    fprintf(stderr, "[cpz prof] cycles per RHS: phase0 %.0f phase1 %.0f phase2 %.0f phase3+ %.0f stencil+update %.0f | per step: save %.0f | total/RHS %.0f\n",
            hcnt[0] / n_rhs, hcnt[1] / n_rhs, hcnt[2] / n_rhs, hcnt[3] / n_rhs, hcnt[4] / n_rhs, hcnt[5] / (double)m->tm.n_steps, hcnt[7] / n_rhs);
    m->ctx->launches++;
    return CPZ_OK;
  }
  kern<<<n_tiles, NT, smem, m->ctx->stream>>>(m->fwd.M, m->tab, m->tm, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

int launch_solve(cpz_model* m, const SolveArgs& a) {
  static const bool prof = getenv("CPZ_PROF") != nullptr;
  if (!prof) {
    int rc = launch_solve_tc(m, a);  // tcgen05 path for the production u/v/T nets
    if (rc <= 0) return rc;
    rc = launch_solve_nnfree(m, a);  // NN-free u/v/T model
    if (rc <= 0) return rc;
    rc = launch_solve_fc_tc(m, a);   // T-only free-convection nets
    if (rc <= 0) return rc;
  }
  if (m->fwd.M.w_in_smem) return launch_solve_t<32, 256, true>(m, a);
  return launch_solve_t<32, 256, false>(m, a);
}


}  // namespace cpz

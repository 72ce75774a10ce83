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

// Column tile of the training pass. The adjoint kernel is latency-bound per Runge–Kutta stage (measured: a 4-column tile
// takes 0.59 of the time of a 32-column one), so the smallest tile that still fits ONE wave of the device is the fastest:
// 4-, 8- or 16-column tiles for batches up to 4, 8 or 16 x SMs columns (one column in BASELINE config 1, the reference's
// 9-18 simulations, 1 152 / 2 304 columns per GPU of config 3 on eight / four GPUs), 32-column tiles beyond.
// CPZ_SMALL_NCOL overrides.
int train_tile(const cpz_model* m, size_t ncol) {
  const char* ov = getenv("CPZ_SMALL_NCOL");  // "0": always 32-column tiles; n > 0: 4-column tiles up to n columns
  const size_t sms = (size_t)(m->ctx->sm_count > 0 ? m->ctx->sm_count : 148);
  if (ov) return (m->has_small[0] && ncol <= (size_t)atol(ov)) ? 4 : m->CT;
  for (int i = 0; i < cpz_model::N_SMALL; ++i) {
    const int ct = cpz_model::small_ct(i);
    if (m->has_small[i] && (ncol + ct - 1) / ct <= sms) return ct;
  }
  return m->CT;
}

template <bool WS>
static int launch_flux_t(cpz_model* m, const SolveArgs& a) {
  constexpr int CT = 32, NT = 256;
  const ModelD& M = m->bwd.M;  // keeps the 3 (Nz+1) face-flux scratch rows
  TableauD tab1 = m->tab;
  tab1.n_stages = 1;  // a single evaluation needs no Runge–Kutta stage slots (the adjoint plan's arena is the larger one)
  const SolveSmem L = solve_smem_layout(M, CT, tab1.n_stages);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "flux kernel needs %zu B shared memory, device allows %zu", smem, m->ctx->smem_optin);
  auto kern = solve_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(a.ncol + CT - 1) / CT, NT, smem, m->ctx->stream>>>(M, tab1, m->tm, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}
int launch_flux(cpz_model* m, const SolveArgs& a) {
  if (!m->has_bwd) return fail(CPZ_ERR_INVALID, "predict_flux needs the adjoint plan: %s", m->bwd_err.c_str());
  return m->bwd.M.w_in_smem ? launch_flux_t<true>(m, a) : launch_flux_t<false>(m, a);
}

int launch_solve(cpz_model* m, const SolveArgs& a) {
  static const bool prof = getenv("CPZ_PROF") != nullptr;
  if (!prof) {
    int rc = launch_solve_tc(m, a);  // tcgen05 path for the production u/v/T nets
    if (rc <= 0) return rc;
    rc = launch_solve_nnfree(m, a);  // NN-free u/v/T model
    if (rc <= 0) return rc;
    // a small-tile training pass of a T-only model: a handful of columns would occupy one 128-column tensor-core tile per
    // CTA; the 4-column FP32 tiles are the lower-latency choice there (BASELINE config 1)
    if (!(a.small_tiles && a.ckpt != nullptr && a.ncol <= 64)) {
      rc = launch_solve_fc_tc(m, a);   // T-only free-convection nets
      if (rc <= 0) return rc;
    }
  }
  if (a.small_tiles && a.ckpt != nullptr) return launch_solve_small(m, a, a.small_tiles);
  if (m->fwd.M.w_in_smem) return launch_solve_t<32, 256, true>(m, a);
  return launch_solve_t<32, 256, false>(m, a);
}


}  // namespace cpz

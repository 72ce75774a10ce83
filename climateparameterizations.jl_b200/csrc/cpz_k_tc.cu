// Tensor-core (tcgen05) forward solve: eligibility, weight image, launcher.
#include <cstdlib>

#include "cpz_launch.h"
#include "cpz_tc.cuh"

namespace cpz {

// The tcgen05 kernel covers the production wind_mixing nets (wind_mixing/train_NDE.jl:103: three Chains
// Dense(96,h1,act) -> Dense(h1,h2,act) -> Dense(h2,31)) with 3*h1 <= 160 and h2 <= 32; everything else runs on the
// FP32 SIMT kernel (cpz_solve.cuh).
bool tc_plan(const cpz_model* m, TcD& T, std::string& why) {
  const cpz_model_desc& d = m->desc;
  if (d.n_fields != 3 || d.Nz != 32 || d.n_nets != 3) { why = "needs the u/v/T model with three nets at Nz = 32"; return false; }
  if (d.variant != CPZ_RHS_TRAIN && d.variant != CPZ_RHS_INFER) { why = "variant"; return false; }
  if (d.flags & (CPZ_FLAG_SMOOTH_NN | CPZ_FLAG_SMOOTH_RI)) { why = "smoothing filters"; return false; }
  const cpz_net_desc& n0 = d.nets[0];
  if (n0.n_layers != 3 || n0.sizes[0] != 96 || n0.sizes[3] != 31) { why = "net depth/shape"; return false; }
  for (int q = 1; q < 3; ++q) {
    const cpz_net_desc& n = d.nets[q];
    if (n.n_layers != 3) { why = "net depth"; return false; }
    for (int l = 0; l <= 3; ++l) if (n.sizes[l] != n0.sizes[l]) { why = "nets differ in shape"; return false; }
    for (int l = 0; l < 3; ++l) if (n.act[l] != n0.act[l]) { why = "nets differ in activation"; return false; }
  }
  if (n0.act[2] != CPZ_ACT_IDENTITY) { why = "output activation"; return false; }
  T = TcD{};
  T.h1 = n0.sizes[1]; T.h2 = n0.sizes[2]; T.nout = 31;
  T.act1 = n0.act[0]; T.act2 = n0.act[1]; T.act3 = n0.act[2];
  if (3 * T.h1 > 160 || T.h2 > 32 || T.h1 < 1 || T.h2 < 1) { why = "hidden widths outside 3*h1 <= 160, h2 <= 32"; return false; }
  T.n1b = 3 * T.h1 > 128 ? 3 * T.h1 - 128 : 0;
  auto windows = [](int width, int* start, int& steps, int& rows) {
    steps = 0; rows = 0;
    for (int q = 0; q < 3; ++q) {
      start[q] = (width * q) & ~3;
      const int s = (width * q + width - start[q] + 7) / 8;
      if (s > steps) steps = s;
    }
    for (int q = 0; q < 3; ++q) if (start[q] + 8 * steps > rows) rows = start[q] + 8 * steps;
    rows = (rows + 3) & ~3;
  };
  windows(T.h1, T.k2_start, T.k2_steps, T.h1_rows);
  windows(T.h2, T.k3_start, T.k3_steps, T.h2_rows);
  if (T.h1_rows < 128 + 32) T.h1_rows = 160;  // the layer-1 epilogue writes rows 0..127 (+ block 1) unconditionally
  const int K2 = 8 * T.k2_steps, K3 = 8 * T.k3_steps;
  if (2 * K2 + 2 * K3 > 192) { why = "layer-2/3 stacks exceed the 192 free TMEM columns"; return false; }
  T.c_a2hi = 192; T.c_a2lo = 192 + K2; T.c_a3hi = 192 + 2 * K2; T.c_a3lo = 192 + 2 * K2 + K3;
  const ModelD& M = m->fwd.M;
  int found = 0;
  for (int i = 0; i < M.n_gemm; ++i) {
    const GemmD& g = M.gemm[i];
    if (g.net < 0 || g.net >= 3 || g.layer < 0 || g.layer >= 3) { why = "plan"; return false; }
    T.w_off[g.net][g.layer] = g.w_off; T.b_off[g.net][g.layer] = g.b_off;
    ++found;
  }
  if (found != 9) { why = "plan"; return false; }
  const TcSmem L = tc_smem_layout(T, m->tab.n_stages);
  if ((size_t)L.total > m->ctx->smem_optin) { why = "shared memory"; return false; }
  return true;
}

std::string tc_describe(const cpz_model* m) {
  TcD T;
  std::string why;
  char line[512];
  if (getenv("CPZ_NO_TC") != nullptr) return "forward kernel: fp32-simt (CPZ_NO_TC set)\n";
  if (!tc_plan(m, T, why)) return "forward kernel: fp32-simt (tcgen05 path not eligible: " + why + ")\n";
  const TcSmem L = tc_smem_layout(T, m->tab.n_stages);
  snprintf(line, sizeof(line),
           "forward kernel: tcgen05 3xTF32 (weights in TMEM; %d column groups x %d columns, %d threads; layer-1 rows %d+%d, "
           "layer-2 K windows %d/%d/%d x%d steps, layer-3 K windows %d/%d/%d x%d steps; MMAs per RHS and group %d; smem %d B)\n",
           TC_NG, TC_GN, TC_NT, 3 * T.h1 - T.n1b, T.n1b, T.k2_start[0], T.k2_start[1], T.k2_start[2], T.k2_steps, T.k3_start[0],
           T.k3_start[1], T.k3_start[2], T.k3_steps, 36 * (T.n1b > 0 ? 2 : 1) + 9 * T.k2_steps + 9 * T.k3_steps, L.total);
  return line;
}

template <int ACT>
static int launch_tc_t(cpz_model* m, const TcD& T, const SolveArgs& a, const TcArgs& ta) {
  const TcSmem L = tc_smem_layout(T, m->tab.n_stages);
  auto kern = solve_tc_kernel<ACT>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const int n_tiles = (a.ncol + TC_CT - 1) / TC_CT;
  kern<<<n_tiles, TC_NT, L.total, m->ctx->stream>>>(m->fwd.M, T, m->tab, m->tm, a, ta);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

// returns 1 when the model is not eligible (caller falls back to the SIMT kernel), 0 on success, <0 on error
int launch_solve_tc(cpz_model* m, const SolveArgs& a) {
  if (getenv("CPZ_NO_TC") != nullptr) return 1;
  TcD T;
  std::string why;
  if (!tc_plan(m, T, why)) return 1;
  const size_t need = (size_t)TC_WCOLS * 128;
  if (m->b_wimg.cap < need) {
    if (m->b_wimg.p) cudaFree(m->b_wimg.p);
    m->b_wimg.p = nullptr; m->b_wimg.cap = 0;
    CPZ_CUDA(cudaMalloc(&m->b_wimg.p, need * sizeof(float)));
    m->b_wimg.cap = need;
  }
  tc_image_kernel<<<(int)((need + 255) / 256), 256, 0, m->ctx->stream>>>(T, a.theta, m->b_wimg.p);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  TcArgs ta{m->b_wimg.p};
  if (T.act1 == T.act2 && T.act1 == ACT_MISH) return launch_tc_t<ACT_MISH>(m, T, a, ta);
  if (T.act1 == T.act2 && T.act1 == ACT_RELU) return launch_tc_t<ACT_RELU>(m, T, a, ta);
  return launch_tc_t<-1>(m, T, a, ta);
}

}  // namespace cpz

// Closure kernel instantiations and launcher.
#include <algorithm>

#include <cstdlib>

#include "cpz_launch.h"
#include "cpz_closure_tc.cuh"
#include "cpz_closure_uvt.cuh"
#include "cpz_fc_tc.cuh"

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

// tcgen05 closure: T-only model, one net Nz -> h1 -> h2 -> Nz-1 with h1, h2 <= 128 (free_convection/train_free_convection_nde.jl:119-121).
static bool closure_tc_plan(const cpz_model* m, ClosureTcD& C) {
  const cpz_model_desc& d = m->desc;
  if (getenv("CPZ_NO_TC") != nullptr) return false;
  if (d.n_fields != 1 || d.n_nets != 1 || d.Nz != 32) return false;
  const cpz_net_desc& n = d.nets[0];
  if (n.n_layers != 3 || n.sizes[0] != 32 || n.sizes[3] != 31 || n.act[2] != CPZ_ACT_IDENTITY) return false;
  if (n.sizes[1] < 1 || n.sizes[1] > 128 || n.sizes[2] < 1 || n.sizes[2] > 128) return false;
  C = ClosureTcD{};
  C.h1 = n.sizes[1]; C.h2 = n.sizes[2]; C.nout = 31;
  C.n1 = (C.h1 + 15) & ~15; C.n2 = (C.h2 + 15) & ~15; C.n3 = 32;
  C.k2 = (C.h1 + 7) & ~7; C.k3 = (C.h2 + 7) & ~7;
  C.act1 = n.act[0]; C.act2 = n.act[1];
  const ModelD& M = m->fwd.M;
  int found = 0;
  for (int i = 0; i < M.n_gemm; ++i) {
    const GemmD& g = M.gemm[i];
    if (g.net != 0 || g.layer < 0 || g.layer > 2) return false;
    C.w_off[g.layer] = g.w_off; C.b_off[g.layer] = g.b_off; ++found;
  }
  if (found != 3) return false;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) & ~127; return r; };
  C.o_w1h = take(C.n1 * 32 * 4); C.o_w1l = take(C.n1 * 32 * 4);
  C.o_w2h = take(C.n2 * C.k2 * 4); C.o_w2l = take(C.n2 * C.k2 * 4);
  C.o_w3h = take(C.n3 * C.k3 * 4); C.o_w3l = take(C.n3 * C.k3 * 4);
  C.o_b = take((C.n1 + C.n2 + C.n3) * 4);
  C.img_bytes = o;
  return (size_t)o + 1024 <= m->ctx->smem_optin;
}

template <int ACT>
static int launch_closure_tc_t(cpz_model* m, const ClosureTcD& C, const ClosureD& cd, ClosureArgs a) {
  auto kern = closure_tc_kernel<ACT>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C.img_bytes));
  a.n_tiles = (a.ncol + CTC_TILE - 1) / CTC_TILE;
  const int grid = std::min(a.n_tiles, m->ctx->sm_count);
  kern<<<grid, CTC_NT, C.img_bytes, m->ctx->stream>>>(C, cd, a, m->b_cimg.p);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

bool closure_uses_tc(const cpz_model* m) {
  ClosureTcD C;
  return closure_tc_plan(m, C);
}

static int ensure_image(cpz_model* m, const ClosureTcD& C, const float* theta) {
  const size_t need = (size_t)C.img_bytes / 4;
  if (m->b_cimg.cap < need) {
    if (m->b_cimg.p) cudaFree(m->b_cimg.p);
    m->b_cimg.p = nullptr; m->b_cimg.cap = 0; m->cimg_ver = 0;
    CPZ_CUDA(cudaMalloc(&m->b_cimg.p, need * sizeof(float)));
    m->b_cimg.cap = need;
  }
  if (m->cimg_ver != m->theta_ver) {  // weights changed since the image was built
    const int n_el = C.n1 * 32 + C.n2 * C.k2 + C.n3 * C.k3 + C.n1 + C.n2 + C.n3;
    closure_tc_image_kernel<<<(n_el + 255) / 256, 256, 0, m->ctx->stream>>>(C, theta, m->b_cimg.p);
    CPZ_CUDA(cudaGetLastError());
    m->ctx->launches++;
    m->cimg_ver = m->theta_ver;
  }
  return CPZ_OK;
}

template <int ACT, bool MPP>
static int launch_fc_tc_t(cpz_model* m, const ClosureTcD& C, const SolveArgs& a) {
  auto kern = solve_fc_tc_kernel<ACT, MPP>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C.img_bytes));
  const int n_tiles = (a.ncol + CTC_TILE - 1) / CTC_TILE;
  // full waves of two-tile CTAs; a remainder that fits one wave of single-tile CTAs runs as such (a single tile per CTA
  // finishes an evaluation sooner than a pair)
  const int sms = m->ctx->sm_count > 0 ? m->ctx->sm_count : 148;
  int n_pair = (n_tiles + 1) / 2, n_single = 0;
  const int full = n_tiles / (2 * sms), rem = n_tiles - full * 2 * sms;
  if (rem > 0 && rem <= sms) { n_pair = full * sms; n_single = rem; }
  kern<<<n_pair + n_single, CTC_NT, C.img_bytes, m->ctx->stream>>>(C, m->fwd.M, m->tab, m->tm, a, m->b_cimg.p, m->b_fcscr.p, n_pair);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

// T-only free-convection forward solve / RHS on tcgen05 (same net limits as the closure). 1 = not eligible.
int launch_solve_fc_tc(cpz_model* m, const SolveArgs& a) {
  ClosureTcD C;
  if (m->desc.variant != CPZ_RHS_FREE_CONVECTION || !closure_tc_plan(m, C)) return 1;
  if (m->desc.flags & CPZ_FLAG_IMPLICIT_DIFFUSION) return 1;  // the implicit-diffusion step lives in the FP32 tile kernels
  int rc = ensure_image(m, C, a.theta);
  if (rc) return rc;
  const size_t n_tiles = 2 * (((size_t)(a.ncol + CTC_TILE - 1) / CTC_TILE + 1) / 2);  // tile pairs
  const size_t need = n_tiles * (size_t)m->tab.n_stages * 32 * CTC_TILE;
  if (m->b_fcscr.cap < need) {
    if (m->b_fcscr.p) cudaFree(m->b_fcscr.p);
    m->b_fcscr.p = nullptr; m->b_fcscr.cap = 0;
    CPZ_CUDA(cudaMalloc(&m->b_fcscr.p, need * sizeof(float)));
    m->b_fcscr.cap = need;
  }
  const bool same = C.act1 == C.act2;
  if (m->desc.flags & CPZ_FLAG_MPP) {  // T-only model with the mPP base (BASELINE config 1)
    if (same && C.act1 == ACT_RELU) return launch_fc_tc_t<ACT_RELU, true>(m, C, a);
    return launch_fc_tc_t<-1, true>(m, C, a);
  }
  if (same && C.act1 == ACT_RELU) return launch_fc_tc_t<ACT_RELU, false>(m, C, a);
  if (same && C.act1 == ACT_MISH) return launch_fc_tc_t<ACT_MISH, false>(m, C, a);
  return launch_fc_tc_t<-1, false>(m, C, a);
}

template <int CT, int NT, bool WS>
static int launch_closure_uvt_t(cpz_model* m, const ClosureUvtD& cd, const ClosureUvtArgs& a) {
  const ClosureUvtSmem L = closure_uvt_smem_layout(m->fwd.M, CT);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > m->ctx->smem_optin) return fail(CPZ_ERR_INVALID, "u/v/T closure kernel needs %zu B shared memory", smem);
  auto kern = closure_uvt_kernel<CT, NT, WS>;
  CPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(a.n_tiles, m->ctx->sm_count);
  kern<<<grid, NT, smem, m->ctx->stream>>>(m->fwd.M, cd, a);
  CPZ_CUDA(cudaGetLastError());
  m->ctx->launches++;
  return CPZ_OK;
}

// the shared-memory weight arena of the FP32 MLP phases as a device image (same gather as load_weights_smem, done once per theta)
static __global__ void __launch_bounds__(256) closure_uvt_image_kernel(const __grid_constant__ ModelD M, const float* __restrict__ theta,
                                                                       float* __restrict__ img) {
  load_weights_smem<256>(M, img, theta);
}

int launch_closure_uvt(cpz_model* m, const ClosureUvtD& cd, const ClosureUvtArgs& a0) {
  {  // production nets at Nz = 32: the tcgen05 kernel's CLOSURE instantiation (CPZ_NO_TC=1 keeps the FP32 SIMT kernel below)
    const int rc = launch_closure_uvt_tc(m, cd, a0);
    if (rc != 1) return rc;
  }
  ClosureUvtArgs a = a0;
  a.wimg = nullptr;
  if (m->fwd.M.w_in_smem) {
    const size_t need = (size_t)m->fwd.M.smem_w_floats;
    if (m->b_cimg.cap < need) {
      if (m->b_cimg.p) cudaFree(m->b_cimg.p);
      m->b_cimg.p = nullptr; m->b_cimg.cap = 0; m->cimg_ver = 0;
      CPZ_CUDA(cudaMalloc(&m->b_cimg.p, need * sizeof(float)));
      m->b_cimg.cap = need;
    }
    if (m->cimg_ver != m->theta_ver) {  // weights changed since the image was built
      closure_uvt_image_kernel<<<1, 256, 0, m->ctx->stream>>>(m->fwd.M, a.theta, m->b_cimg.p);
      CPZ_CUDA(cudaGetLastError());
      m->ctx->launches++;
      m->cimg_ver = m->theta_ver;
    }
    a.wimg = m->b_cimg.p;
    return launch_closure_uvt_t<32, 256, true>(m, cd, a);
  }
  return launch_closure_uvt_t<32, 256, false>(m, cd, a);
}

int launch_closure(cpz_model* m, const ClosureD& cd, const ClosureArgs& a) {
  ClosureTcD C;
  if (closure_tc_plan(m, C)) {
    const int rci = ensure_image(m, C, a.theta);
    if (rci) return rci;
    const bool same = C.act1 == C.act2;
    if (same && C.act1 == ACT_RELU) return launch_closure_tc_t<ACT_RELU>(m, C, cd, a);
    if (same && C.act1 == ACT_MISH) return launch_closure_tc_t<ACT_MISH>(m, C, cd, a);
    return launch_closure_tc_t<-1>(m, C, cd, a);
  }
  if (m->fwd.M.w_in_smem) return launch_closure_t<32, 256, true>(m, cd, a);
  return launch_closure_t<32, 256, false>(m, cd, a);
}

}  // namespace cpz

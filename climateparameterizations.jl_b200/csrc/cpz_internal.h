// Internal host-side structures behind the opaque handles of include/cpz.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/cpz.h"
#include "cpz_plan.h"

struct cpz_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cpz_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  int rank = 0, world = 1;
  uint64_t launches = 0;
  int sm_count = 0;
  cudaStream_t copy_stream = nullptr;  // D2H of finished trajectory chunks, overlapped with the next chunk's kernel
  cudaEvent_t chunk_ev[2] = {nullptr, nullptr};
  size_t smem_optin = 0;
  // device counters of non-finite values seen in results (CPZ_ERR_NONFINITE): [0] cumulative since the context was
  // created (what the *_dev flavours report through cpz_ctx_nonfinite_count), [1] of the current host-flavour call
  unsigned int* d_nonfinite = nullptr;
};

struct DevBuf {
  float* p = nullptr;
  size_t cap = 0;  // floats
};

struct cpz_model {
  cpz_ctx* ctx = nullptr;
  cpz_model_desc desc;
  size_t P = 0;
  cpz::Plan fwd;  // forward plan (aliasing arena, free TO)
  cpz::Plan bwd;  // adjoint plan (all activations kept)
  bool has_bwd = false;
  bool implicit_adjoint_ok = false;  // set once the adjoint kernels carry the VJP of the implicit-diffusion step
  // the same two plans for CT_SMALL-column tiles: small batches (BASELINE config 1: one column; the reference's 9-18
  // simulations) and shards too small to give every SM a 32-column tile run on these
  static constexpr int N_SMALL = 3;
  static constexpr int small_ct(int i) { return 4 << i; }  // 4, 8, 16
  cpz::Plan fwd_s[N_SMALL], bwd_s[N_SMALL];
  bool has_small[N_SMALL] = {false, false, false};
  std::string bwd_err;
  TableauD tab;
  TimeD tm;
  int CT = 32, NT = 256;
  // parameters and optimiser state (device)
  float* d_theta = nullptr;
  float* d_m = nullptr;
  float* d_v = nullptr;
  int* d_gmap = nullptr;           // destructure index -> offset in the adjoint's gradient slab
  float beta_pow[2] = {0.f, 0.f};  // 0 = uninitialised (set to (beta1,beta2) on the first step)
  // grow-only device scratch
  DevBuf b_x0, b_bcs, b_q, b_traj, b_tgt, b_ckpt, b_scr, b_part, b_red, b_out, b_w, b_wimg, b_cimg, b_fcscr, b_kstore;
  DevBuf b_aux, b_bwimg, b_tcadj;  // tensor-core adjoint: per-stage records, transposed weight image, xbar + loss sums + accumulator images
  uint64_t theta_ver = 1;   // bumped whenever d_theta changes (set_theta, ADAM step)
  uint64_t cimg_ver = 0;    // theta_ver the closure weight image was built from
  uint64_t wimg_ver = 0;    // theta_ver the tensor-memory weight image (b_wimg) was built from
};

namespace cpz {
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
}  // namespace cpz

#define CPZ_CUDA(call)                                        \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return cpz::cuda_fail(e__, #call); \
  } while (0)

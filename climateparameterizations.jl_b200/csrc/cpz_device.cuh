// Device building blocks of the NDE column engine (sm_100a).
//
// Data layout inside a CTA ("column tile" of CT columns):
//   every per-column vector lives in shared memory FEATURE-MAJOR: v[row][c], c = column in tile fastest.
//   - elementwise / stencil work maps lanes to consecutive columns  -> conflict-free LDS/STS
//   - the MLP runs as a small dense contraction out[N][CT] = act(W^T in[K][CT] + b): a thread owns a
//     4-column x TO-output register tile, reads 4 columns of one input row as one LDS.128 (broadcast within
//     the column group) and TO weights of one W row (W is [K][N], Flux's column-major out x in).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cpz_model.h"

namespace cpz {

// ---- flags / enums mirrored from include/cpz.h (kept numeric here to avoid including the C header in device code)
enum { RHS_TRAIN = 0, RHS_INFER = 1, RHS_FC = 2 };
enum {
  F_MPP = 1, F_CA = 2, F_ZERO_WEIGHTS = 4, F_SMOOTH_NN = 8, F_SMOOTH_RI = 16, F_DIURNAL = 32, F_CA_LITERAL_U = 64,
  F_DIURNAL_UNSHIFTED = 128, F_IMPLICIT = 256
};
enum { ACT_ID = 0, ACT_RELU = 1, ACT_MISH = 2, ACT_SWISH = 3, ACT_LEAKY = 4, ACT_TANH = 5 };

// ---- activations (NNlib 0.7.20 definitions) -----------------------------------------------------------------------
// exp via MUFU.EX2 (__expf, <= 2+|1.16x| ulp) and division via MUFU.RCP (__fdividef, 2 ulp): errors ~1e-6 relative,
// an order of magnitude inside the 1e-5 RHS tolerance, at a quarter of the instruction count of expf + IEEE divide.
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case ACT_RELU: return x < 0.f ? 0.f : x;  // NaN propagates, as Julia's max(0, x) does (fmaxf would return 0)
    case ACT_MISH: {  // x*tanh(softplus(x)) = x*n/(n+2), n = e^x(e^x+2)
      const float e = __expf(fminf(x, 20.f));
      const float n = e * (e + 2.f);
      return x * __fdividef(n, n + 2.f);
    }
    case ACT_SWISH: return __fdividef(x, 1.f + __expf(-x));
    case ACT_LEAKY: return x < 0.f ? 0.01f * x : x;
    case ACT_TANH: {  // 1 - 2/(e^{2x}+1)
      const float e = __expf(2.f * (x > 40.f ? 40.f : x));  // not fminf: a NaN argument must stay NaN
      return 1.f - __fdividef(2.f, e + 1.f);
    }
    default: return x;
  }
}

// derivative d act / d z at the pre-activation z
__device__ __forceinline__ float act_grad(int act, float z) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_MISH: {
      const float e = __expf(fminf(z, 20.f));
      const float n = e * (e + 2.f);
      const float t = __fdividef(n, n + 2.f);
      const float sg = __fdividef(e, 1.f + e);
      return t + z * (1.f - t * t) * sg;
    }
    case ACT_SWISH: {
      const float sg = __fdividef(1.f, 1.f + __expf(-z));
      return sg + z * sg * (1.f - sg);
    }
    case ACT_LEAKY: return z > 0.f ? 1.f : 0.01f;
    case ACT_TANH: {
      const float e = __expf(2.f * fminf(z, 40.f));
      const float t = 1.f - __fdividef(2.f, e + 1.f);
      return 1.f - t * t;
    }
    default: return 1.f;
  }
}

// ---- async bulk copies (TMA engine, non-tensor form: UBLKCP) --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion on an mbarrier. dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, bulk-group completion.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace cpz
#include "cpz_gemm.cuh"
namespace cpz {

// Copy theta into the shared weight arena in the plan's padded layout (rows of Npad floats, zero padded).
template <int NT>
__device__ __noinline__ void load_weights_smem(const ModelD& M, float* __restrict__ wsm, const float* __restrict__ theta) {
  for (int i = threadIdx.x; i < M.smem_w_floats; i += NT) wsm[i] = 0.f;
  __syncthreads();
  for (int gi = 0; gi < M.n_gemm; ++gi) {
    const GemmD& g = M.gemm[gi];
    const int K = g.K, N = g.N, Np = g.Npad;
    for (int i = threadIdx.x; i < K * N; i += NT) {
      const int k = i / N, j = i - k * N;
      wsm[g.sw_off + k * Np + j] = __ldg(theta + g.w_off + i);
    }
    for (int j = threadIdx.x; j < N; j += NT) wsm[g.sb_off + j] = __ldg(theta + g.b_off + j);
  }
}

// ---- per-column boundary data ------------------------------------------------------------------------------------
// bcf[q][0/1][CT]: effective boundary-face fluxes (bottom, top) for field q after the variant's shift rule.
// For the diurnal case the top wT entry is recomputed every stage from Q and t.
__device__ __forceinline__ void bc_effective(const ModelD& M, const float* __restrict__ bc /*nbc raw values*/, float* out6) {
  if (M.variant == RHS_FC) {
    out6[0] = bc[0]; out6[1] = bc[1];
    return;
  }
  const bool mpp = (M.flags & F_MPP) || M.variant == RHS_INFER;
  const bool shift = M.variant == RHS_INFER || (M.flags & F_ZERO_WEIGHTS);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    float b = bc[2 * q], t = bc[2 * q + 1];
    if (mpp && shift) { b -= M.rc.z0[q]; t -= M.rc.z0[q]; }
    else if (M.flags & F_ZERO_WEIGHTS) { b = 0.f; t = 0.f; }
    out6[2 * q] = b; out6[2 * q + 1] = t;
  }
}

__device__ __forceinline__ float diurnal_top_eff(const ModelD& M, float Q, float t) {
  // s_wT(Q sin(2 pi t tau / period)/(alpha g))   NDE_training.jl:72-73, data_containers.jl:135
  float top = (Q * sinf(M.rc.di_w * t) * M.rc.di_amp - M.rc.mu_wT) * M.rc.inv_sig_wT;
  const bool mpp = (M.flags & F_MPP) || M.variant == RHS_INFER;
  const bool shift = M.variant == RHS_INFER || (M.flags & F_ZERO_WEIGHTS);
  if (mpp && shift) {
    if (!(M.variant == RHS_INFER && (M.flags & F_DIURNAL_UNSHIFTED))) top -= M.rc.z0[2];
  } else if (M.flags & F_ZERO_WEIGHTS) {
    top = 0.f;
  }
  return top;
}

// width-3 running mean of filtering_operators.jl:1-15 applied to v[0..n) at index i (v row-strided by CT)
__device__ __forceinline__ float filt3(const float* __restrict__ v, int i, int n, int stride) {
  if (i == 0) return (v[0] + v[stride]) * 0.5f;
  if (i == n - 1) return (v[(n - 2) * stride] + v[(n - 1) * stride]) * 0.5f;
  return (v[(i - 1) * stride] + v[i * stride] + v[(i + 1) * stride]) * (1.f / 3.f);
}

// Richardson number at face k (0..N) of column c from the state tile (u,v,T rows), boundary faces have zero gradient.
__device__ __forceinline__ float ri_face(const ModelD& M, const float* __restrict__ X, int k, int c, int CT_, float eps) {
  const int N = M.Nz;
  float Gu = 0.f, Gv = 0.f, GT = 0.f;
  if (k > 0 && k < N) {
    Gu = M.rc.Nf * (X[k * CT_ + c] - X[(k - 1) * CT_ + c]);
    Gv = M.rc.Nf * (X[(N + k) * CT_ + c] - X[(N + k - 1) * CT_ + c]);
    GT = M.rc.Nf * (X[(2 * N + k) * CT_ + c] - X[(2 * N + k - 1) * CT_ + c]);
  }
  const float su = M.rc.sig_u * (Gu + eps), sv = M.rc.sig_v * (Gv + eps);
  return __fdividef(M.rc.BzC * (GT + eps), su * su + sv * sv);
}

// nu = nu0 + nu_m*(1 - tanh(y))/2 with y = (Ri - Ric)/dRi, evaluated as nu0 + nu_m/(1 + e^{2y}) (identical function,
// no cancellation for large y). NDE_training.jl:54,125
__device__ __forceinline__ float nu_of_ri(const ModelD& M, float Ri) {
  const float y2 = 2.f * (Ri - M.rc.Ric) * M.rc.inv_dRi;
  return M.rc.nu0 + __fdividef(M.rc.nu_m, 1.f + __expf(y2));
}

// T-only model with the mPP base diffusivity (BASELINE config 1): the rule of NDE_training.jl:114-139 at u = v = 0, where
// both shear gradients are D_face*0 + eps. Returns c_T nu/Pr at a face with scaled gradient G; *dG (optional) receives
// d(c_T nu/Pr)/dG. The exponent is clamped: at these shear-free Richardson numbers (|Ri| ~ 1e15 |G|) the tanh step is
// saturated everywhere except within ~1e-16 of G = -eps.
__device__ __forceinline__ float fc_mpp_cnu(const ModelD& M, float G, float* dG = nullptr) {
  const float k = M.rc.BzC * M.rc.fc_iS2;
  const float y2 = 2.f * (k * (G + M.rc.eps) - M.rc.Ric) * M.rc.inv_dRi;
  const float s = __fdividef(1.f, 1.f + __expf(y2 > 80.f ? 80.f : y2));
  const float cp = M.rc.c[2] * M.rc.inv_Pr;
  if (dG) *dG = -cp * M.rc.nu_m * s * (1.f - s) * (2.f * M.rc.inv_dRi * k);
  return cp * (M.rc.nu0 + M.rc.nu_m * s);
}

// ---- face phase: E_q[k][c] for all faces k = 0..N ------------------------------------------------------------------
// E is the total (NN + diffusive [+ boundary]) flux whose cell-difference gives the tendency.
// nn rows come from the activation arena (M.nn_off), or are zero when the model has no nets.
template <int CT, int NT>
__device__ __noinline__ void faces_phase(const ModelD& M, const float* __restrict__ X, const float* __restrict__ arena,
                                            float* __restrict__ E, const float* __restrict__ bcf /*[nbc][CT]*/) {
  const int N = M.Nz;
  const int nfaces = N + 1;
  const bool has_nn = M.n_nets > 0;
  if (M.variant == RHS_FC) {
    const float* nn = has_nn ? arena + M.nn_off[0] * CT : nullptr;
    const bool expl = !(M.flags & F_IMPLICIT);  // implicit diffusion: the RHS carries the NN and boundary fluxes only
    const bool ca = (M.flags & F_CA) && expl, mpp1 = (M.flags & F_MPP) && expl;
    for (int i = threadIdx.x; i < nfaces * CT; i += NT) {
      const int k = i / CT, c = i - k * CT;
      float e;
      if (k == 0) e = bcf[c];
      else if (k == N) e = bcf[CT + c];
      else {
        e = has_nn ? nn[(k - 1) * CT + c] : 0.f;
        const float G = M.rc.Nf * (X[k * CT + c] - X[(k - 1) * CT + c]);
        if (mpp1) e -= fc_mpp_cnu(M, G) * G;
        if (ca) e -= fminf(0.f, M.rc.K_ca * G);
      }
      E[i] = e;
    }
    return;
  }
  const bool expl = !(M.flags & F_IMPLICIT);
  const bool mpp = ((M.flags & F_MPP) || M.variant == RHS_INFER) && expl;
  const bool smooth_nn = M.variant == RHS_TRAIN && (M.flags & F_SMOOTH_NN);
  const bool smooth_ri = M.variant == RHS_TRAIN && (M.flags & F_SMOOTH_RI);
  const float eps = M.variant == RHS_TRAIN ? M.rc.eps : 0.f;
  for (int i = threadIdx.x; i < nfaces * CT; i += NT) {
    const int k = i / CT, c = i - k * CT;
    float e[3];
    if (k == 0 || k == N) {
      const int tb = k == 0 ? 0 : 1;
#pragma unroll
      for (int q = 0; q < 3; ++q) e[q] = bcf[(2 * q + tb) * CT + c];
    } else {
      float nn[3] = {0.f, 0.f, 0.f};
      if (has_nn) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float* r = arena + M.nn_off[q] * CT + c;
          nn[q] = smooth_nn ? filt3(r, k - 1, N - 1, CT) : r[(k - 1) * CT];
        }
      }
      const float Gu = M.rc.Nf * (X[k * CT + c] - X[(k - 1) * CT + c]);
      const float Gv = M.rc.Nf * (X[(N + k) * CT + c] - X[(N + k - 1) * CT + c]);
      const float GT = M.rc.Nf * (X[(2 * N + k) * CT + c] - X[(2 * N + k - 1) * CT + c]);
      if (mpp) {
        float Ri;
        if (smooth_ri) {
          const float r0 = ri_face(M, X, k - 1, c, CT, eps), r1 = ri_face(M, X, k, c, CT, eps),
                      r2 = ri_face(M, X, k + 1, c, CT, eps);
          // interior faces 1..N-1 of an (N+1)-vector never hit the filter's shrinking end rows
          Ri = (r0 + r1 + r2) * (1.f / 3.f);
        } else {
          const float su = M.rc.sig_u * (Gu + eps), sv = M.rc.sig_v * (Gv + eps);
          Ri = __fdividef(M.rc.BzC * (GT + eps), su * su + sv * sv);
        }
        const float nu = nu_of_ri(M, Ri);
        float nuT = nu * M.rc.inv_Pr;
        if (M.variant == RHS_INFER && (M.flags & F_CA)) {
          const float test = (M.flags & F_CA_LITERAL_U) ? Gu : GT;
          nuT = test > 0.f ? nu * M.rc.inv_Pr : M.rc.kappa;
        }
        e[0] = nn[0] - M.rc.c[0] * nu * Gu;
        e[1] = nn[1] - M.rc.c[1] * nu * Gv;
        e[2] = nn[2] - M.rc.c[2] * nuT * GT;
      } else if ((M.flags & F_CA) && expl) {
        e[0] = nn[0]; e[1] = nn[1];
        e[2] = nn[2] - M.rc.c[2] * M.rc.kappa * fminf(0.f, GT);
      } else {
        e[0] = nn[0]; e[1] = nn[1]; e[2] = nn[2];
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) E[(q * nfaces + k) * CT + c] = e[q];
  }
}

// tendency of field q at level k, column c from the face fluxes and the stage input X
__device__ __forceinline__ float tendency(const ModelD& M, const float* __restrict__ E, const float* __restrict__ X, int q,
                                          int k, int c, int CT_) {
  const int N = M.Nz, nfaces = N + 1;
  const float dE = E[(q * nfaces + k + 1) * CT_ + c] - E[(q * nfaces + k) * CT_ + c];
  if (M.variant == RHS_FC) return -M.rc.A[2] * M.rc.Nf * dE;
  float r = -M.rc.A[q] * M.rc.Nf * dE;
  if (q == 0) r += M.rc.cor_u_s * X[(N + k) * CT_ + c] + M.rc.cor_u_m;
  else if (q == 1) r -= M.rc.cor_v_s * X[k * CT_ + c] + M.rc.cor_v_m;
  return r;
}

// ---- fused, register-blocked stencil ------------------------------------------------------------------------------------
// One work item = 4 consecutive levels of one column, all fields: the thread loads the 6 levels it needs per field,
// evaluates the 5 bounding face fluxes in registers (mPP diffusivity, convective adjustment, NN fluxes, boundary
// fluxes) and hands the 4 x NF tendencies to `sink(k0, c, dx)`. No flux scratch in shared memory, no barrier between
// "faces" and "centres"; a warp is 32 consecutive columns of one level group (conflict-free LDS).
// Not used for the smooth_NN / smooth_Ri variants (they need neighbouring faces' values; see faces_phase).
template <int CT, int NT, int NF, class Sink>
__device__ __forceinline__ void stencil_fused(const ModelD& M, const float* __restrict__ X, const float* __restrict__ arena,
                                           const float* __restrict__ bcf, Sink sink) {
  const int N = M.Nz;
  const bool has_nn = M.n_nets > 0;
  const bool expl = !(M.flags & F_IMPLICIT);  // implicit diffusion: the RHS carries the NN and boundary fluxes only
  const bool mpp = NF == 3 && ((M.flags & F_MPP) || M.variant == RHS_INFER) && expl;
  const bool ca = (M.flags & F_CA) != 0 && expl;
  const float eps = M.variant == RHS_TRAIN ? M.rc.eps : 0.f;
  const float Nf = M.rc.Nf;
  const int n_items = (N / 4) * CT;
  for (int it = threadIdx.x; it < n_items; it += NT) {
    const int kg = it / CT, c = it - kg * CT;
    const int k0 = 4 * kg;
    // levels k0-1 .. k0+4 (clamped; the clamped values only feed boundary faces, which use bcf instead)
    float xl[NF][6];
#pragma unroll
    for (int q = 0; q < NF; ++q)
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        const int k = min(max(k0 - 1 + l, 0), N - 1);
        xl[q][l] = X[(q * N + k) * CT + c];
      }
    float E[NF][5];
#pragma unroll
    for (int fi = 0; fi < 5; ++fi) {
      const int f = k0 + fi;  // face between levels f-1 and f
      if (f == 0 || f == N) {
        const int tb = f == 0 ? 0 : 1;
#pragma unroll
        for (int q = 0; q < NF; ++q) E[q][fi] = bcf[(2 * q + tb) * CT + c];
        continue;
      }
      float nn[NF];
#pragma unroll
      for (int q = 0; q < NF; ++q) nn[q] = has_nn ? arena[(M.nn_off[q] + f - 1) * CT + c] : 0.f;
      float G[NF];
#pragma unroll
      for (int q = 0; q < NF; ++q) G[q] = Nf * (xl[q][fi + 1] - xl[q][fi]);
      if constexpr (NF == 1) {
        float e = nn[0];
        if ((M.flags & F_MPP) && expl) e -= fc_mpp_cnu(M, G[0]) * G[0];
        if (ca) e -= fminf(0.f, M.rc.K_ca * G[0]);
        E[0][fi] = e;
      } else {
        if (mpp) {
          const float su = M.rc.sig_u * (G[0] + eps), sv = M.rc.sig_v * (G[1] + eps);
          const float Ri = __fdividef(M.rc.BzC * (G[2] + eps), su * su + sv * sv);
          const float nu = nu_of_ri(M, Ri);
          float nuT = nu * M.rc.inv_Pr;
          if (M.variant == RHS_INFER && ca) {
            const float test = (M.flags & F_CA_LITERAL_U) ? G[0] : G[2];
            nuT = test > 0.f ? nuT : M.rc.kappa;
          }
          E[0][fi] = nn[0] - M.rc.c[0] * nu * G[0];
          E[1][fi] = nn[1] - M.rc.c[1] * nu * G[1];
          E[2][fi] = nn[2] - M.rc.c[2] * nuT * G[2];
        } else {
          E[0][fi] = nn[0];
          E[1][fi] = nn[1];
          E[2][fi] = ca ? nn[2] - M.rc.c[2] * M.rc.kappa * fminf(0.f, G[2]) : nn[2];
        }
      }
    }
    float dx[NF][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if constexpr (NF == 1) {
        dx[0][kk] = -M.rc.A[2] * Nf * (E[0][kk + 1] - E[0][kk]);
      } else {
        dx[0][kk] = -M.rc.A[0] * Nf * (E[0][kk + 1] - E[0][kk]) + (M.rc.cor_u_s * xl[1][kk + 1] + M.rc.cor_u_m);
        dx[1][kk] = -M.rc.A[1] * Nf * (E[1][kk + 1] - E[1][kk]) - (M.rc.cor_v_s * xl[0][kk + 1] + M.rc.cor_v_m);
        dx[2][kk] = -M.rc.A[2] * Nf * (E[2][kk + 1] - E[2][kk]);
      }
    }
    sink(k0, c, dx);
  }
}

// ---- implicit vertical diffusion (CPZ_FLAG_IMPLICIT_DIFFUSION) ---------------------------------------------------------------
// Diffusivity D_q at interior face k of column c such that the diffusive flux is -D_q G_q (the terms stencil_fused leaves out
// when the flag is set): u, v: c_q nu; T: c_T nu_T with the variant's convective-adjustment rule; T-only: [mPP] c_T nu/Pr +
// [CA] K [G < 0].
__device__ __forceinline__ float face_diffusivity(const ModelD& M, const float* __restrict__ X, int q, int k, int c, int CT_) {
  const int N = M.Nz;
  const float Nf = M.rc.Nf;
  if (M.variant == RHS_FC) {
    const float G = Nf * (X[k * CT_ + c] - X[(k - 1) * CT_ + c]);
    float D = 0.f;
    if (M.flags & F_MPP) D += fc_mpp_cnu(M, G);
    if ((M.flags & F_CA) && M.rc.K_ca * G < 0.f) D += M.rc.K_ca;
    return D;
  }
  const bool mpp = (M.flags & F_MPP) || M.variant == RHS_INFER;
  const float Gu = Nf * (X[k * CT_ + c] - X[(k - 1) * CT_ + c]);
  const float GT = Nf * (X[(2 * N + k) * CT_ + c] - X[(2 * N + k - 1) * CT_ + c]);
  if (!mpp) return (q == 2 && (M.flags & F_CA) && GT < 0.f) ? M.rc.c[2] * M.rc.kappa : 0.f;
  const float eps = M.variant == RHS_TRAIN ? M.rc.eps : 0.f;
  const float nu = nu_of_ri(M, ri_face(M, X, k, c, CT_, eps));
  if (q < 2) return M.rc.c[q] * nu;
  float nuT = nu * M.rc.inv_Pr;
  if (M.variant == RHS_INFER && (M.flags & F_CA)) {
    const float test = (M.flags & F_CA_LITERAL_U) ? Gu : GT;
    nuT = test > 0.f ? nuT : M.rc.kappa;
  }
  return M.rc.c[2] * nuT;
}

// x_q <- L_q \ x_q for every field of the tile, L_q = tridiag(-r_k, 1 + r_k + r_{k+1}, -r_{k+1}), r_k = h A_q Nz^2 D_q,k on the
// interior faces (0 on the boundary faces), diffusivities of the incoming state (NDE_oceananigans.jl:61-101,
// oceananigans_nn.jl:13-40). One thread per (field, column) runs the Thomas recurrence over the levels; `scr` holds 2 S CT
// floats (r, then the forward-sweep coefficients) — the dead Runge–Kutta stage slots at the start of a step.
// Smoothed-Ri models are not supported (refused by the host). Block barriers inside; the caller syncs afterwards.
template <int CT, int NT>
__device__ __noinline__ void implicit_diffusion_tile(const ModelD& M, float* __restrict__ x, float* __restrict__ scr, float h) {
  const int N = M.Nz, nf = M.nf;
  float* r = scr;                 // [nf][N][CT]: r at face k (below level k), k = 1..N-1; index 0 unused (= 0)
  float* cp = scr + nf * N * CT;  // [nf][N][CT]
  for (int i = threadIdx.x; i < nf * N * CT; i += NT) {
    const int c = i % CT, k = (i / CT) % N, q = i / (CT * N);
    const float A = M.rc.A[nf == 1 ? 2 : q];
    r[i] = k == 0 ? 0.f : h * A * M.rc.Nf * M.rc.Nf * face_diffusivity(M, x, q, k, c, CT);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nf * CT; t += NT) {
    const int c = t % CT, q = t / CT;
    float* xq = x + q * N * CT + c;
    float* rq = r + q * N * CT + c;
    float* cq = cp + q * N * CT + c;
    // incremental form: L (x' - x) = x - L x = r_k (x_{k-1} - x_k) + r_{k+1} (x_{k+1} - x_k); the forward-sweep values of the
    // increment go to the r slot of their row (already consumed), x itself is only touched by the final update
    float cprev = 0.f, dprev = 0.f, xdn = 0.f, xk = xq[0];
    for (int k = 0; k < N; ++k) {
      const float rlo = rq[k * CT], rhi = k + 1 < N ? rq[(k + 1) * CT] : 0.f;
      const float xup = k + 1 < N ? xq[(k + 1) * CT] : xk;
      const float den = 1.f + rlo + rhi + rlo * cprev;  // diag - lower * cp_{k-1}, lower = -rlo
      const float inv = 1.f / den;
      cprev = -rhi * inv;
      dprev = (rlo * (xdn - xk) + rhi * (xup - xk) + rlo * dprev) * inv;
      cq[k * CT] = cprev;
      rq[k * CT] = dprev;
      xdn = xk;
      xk = xup;
    }
    float y = rq[(N - 1) * CT];
    xq[(N - 1) * CT] += y;
    for (int k = N - 2; k >= 0; --k) {
      y = rq[k * CT] - cq[k * CT] * y;
      xq[k * CT] += y;
    }
  }
}

// ---- tile <-> global transposes --------------------------------------------------------------------------------------
// Load columns [col0, col0+CT) of a [ncol][S] global array into dst[S][CT] through `stage` ([CT][S+4] floats),
// using one bulk async copy per column. Columns past ncol replicate the last valid one (keeps the math finite).
template <int CT, int NT>
__device__ __noinline__ void load_tile(float* __restrict__ dst, float* __restrict__ stage, uint64_t* bar, uint32_t& parity,
                                          const float* __restrict__ src, size_t row_stride, int S, int col0, int ncol) {
  const int SP = S + 4;
  if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(CT * S * sizeof(float)));
  __syncthreads();
  if (threadIdx.x < CT) {
    const int col = min(col0 + (int)threadIdx.x, ncol - 1);
    bulk_g2s(stage + threadIdx.x * SP, src + (size_t)col * row_stride, (uint32_t)(S * sizeof(float)), bar);
  }
  mbar_wait(bar, parity);
  parity ^= 1u;
  const int S4 = S / 4;
  for (int i = threadIdx.x; i < S4 * CT; i += NT) {
    const int c = i % CT, s4 = i / CT;
    const float4 v = *reinterpret_cast<const float4*>(stage + c * SP + 4 * s4);
    dst[(4 * s4 + 0) * CT + c] = v.x;
    dst[(4 * s4 + 1) * CT + c] = v.y;
    dst[(4 * s4 + 2) * CT + c] = v.z;
    dst[(4 * s4 + 3) * CT + c] = v.w;
  }
}

// Store src[S][CT] as rows of a [ncol][row_stride] global array (one bulk async copy per valid column).
// The caller must __syncthreads() before (src complete) — this function syncs internally before issuing and the
// issuing threads must call bulk_wait_read0() before `stage` is overwritten again.
template <int CT, int NT>
__device__ __noinline__ void store_tile(const float* __restrict__ src, float* __restrict__ stage, float* __restrict__ dstg,
                                           size_t row_stride, int S, int col0, int ncol) {
  const int SP = S + 4;
  const int S4 = S / 4;
  for (int i = threadIdx.x; i < S4 * CT; i += NT) {
    const int c = i % CT, s4 = i / CT;
    float4 v;
    v.x = src[(4 * s4 + 0) * CT + c];
    v.y = src[(4 * s4 + 1) * CT + c];
    v.z = src[(4 * s4 + 2) * CT + c];
    v.w = src[(4 * s4 + 3) * CT + c];
    *reinterpret_cast<float4*>(stage + c * SP + 4 * s4) = v;
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < CT && col0 + (int)threadIdx.x < ncol) {
    bulk_s2g(dstg + (size_t)(col0 + threadIdx.x) * row_stride, stage + threadIdx.x * SP, (uint32_t)(S * sizeof(float)));
  }
  if (threadIdx.x < CT) bulk_commit();
}

}  // namespace cpz

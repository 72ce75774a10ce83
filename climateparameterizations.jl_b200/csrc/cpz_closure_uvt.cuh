// Per-step closure of the u/v/T NDE embedded in a host ocean model: for every column of three (Nx,Ny,Nz) fields
// the three NN forcing chains on the incoming state followed by the backward-Euler modified Pacanowski–Philander step.
// Replaces the host-model callback progress_neural_network (wind_mixing/src/NDE_oceananigans.jl:380-405):
//   NN_uw_forcing / NN_vw_forcing / NN_wT_forcing  (:288-344)  -> dz_uw_NN, dz_vw_NN, dz_wT_NN
//   modified_pacanowski_philander!                 (:61-101)   -> u', v', T'   (diffusivities :17-58)
// Everything is dimensional (the callback scales the NN input and unscales its output itself).
//
// Same tile scheme as closure_kernel: the fields are x-fastest, so CT consecutive columns at one level are one coalesced
// row of the [row][column] shared-memory layout the MLP phases read. One thread per (field, column) runs the Thomas sweep.
#pragma once
#include "cpz_device.cuh"
#include "cpz_solve.cuh"

namespace cpz {

struct ClosureUvtD {
  int Nx, Ny, Nz;
  float inv_dz, r;          // 1/dz, dt/dz^2
  float top[3];             // uw, vw, wT at the surface face
  int ca;                   // nu_T = Ri > 0 ? nu/Pr : kappa_ca
  float kappa_ca;
  float g_alpha;            // g alpha: Ri = g alpha dT/dz / (du/dz^2 + dv/dz^2)
  float nu0, nu_m, Ric, inv_dRi, inv_Pr;
  float mu[6], sig[6], inv_sig[3];
};

struct ClosureUvtArgs {
  const float* theta;
  const float* wimg;  // the shared-memory weight arena as a device image (rebuilt when theta changes), or null
  const float* f[3];  // u, v, T: [Nz][Ny*Nx]
  float* dzf;         // [3][Nz][Ny*Nx]  dz_uw_NN, dz_vw_NN, dz_wT_NN
  float* out;         // [3][Nz][Ny*Nx]  u', v', T'
  int ncol, n_tiles;
};

struct ClosureUvtSmem {
  int w, st, xin, arena, cp, nu, model, total_floats;
};
__host__ __device__ inline ClosureUvtSmem closure_uvt_smem_layout(const ModelD& M, int CT) {
  ClosureUvtSmem L;
  int o = 0;
  L.w = o; o += M.w_in_smem ? M.smem_w_floats : 0;
  L.st = o; o += 3 * M.Nz * CT;     // incoming u, v, T (dimensional)
  L.xin = o; o += 3 * M.Nz * CT;    // scaled NN input; after the MLP: Thomas d'
  L.arena = o; o += M.arena_floats * CT;
  L.cp = o; o += 3 * M.Nz * CT;     // Thomas c'
  L.nu = o; o += 2 * M.Nz * CT;     // face diffusivities nu and nu_T (face f of column c at [f][c]; face 0 unused)
  L.model = o; o += (int)((sizeof(ModelD) + 15) / 16) * 4;
  L.total_floats = o + 4;
  return L;
}

__device__ __forceinline__ float uvt_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// a / b for the Thomas pivots (|b| >= 1): MUFU.RCP, one Newton step on the reciprocal, one residual correction of the quotient.
// Besides being ~4x shorter than the IEEE division sequence, this keeps ptxas' division slow-path subroutines out of this
// 255-register kernel: with them, two lanes of the temperature warp returned NaN on B200 (same source with a debug printf
// compiled in was clean; not understood further).
__device__ __forceinline__ float uvt_div(float a, float b) {
  float r = uvt_rcp(b);
  r = fmaf(r, fmaf(-b, r, 1.f), r);
  const float q = a * r;
  return fmaf(r, fmaf(-b, q, a), q);
}

template <int CT, int NT, bool WS>
__global__ void __launch_bounds__(NT, 1) closure_uvt_kernel(const __grid_constant__ ModelD Mp, const __grid_constant__ ClosureUvtD cd,
                                                            const __grid_constant__ ClosureUvtArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ClosureUvtSmem L = closure_uvt_smem_layout(Mp, CT);
  const ModelD& M = model_to_smem<NT>(Mp, smem + L.model);
  float* wsm = smem + L.w;
  float* st = smem + L.st;
  float* xin = smem + L.xin;
  float* arena = smem + L.arena;
  float* cp = smem + L.cp;
  float* nus = smem + L.nu;
  const int N = M.Nz;
  if (WS) {
    if (a.wimg != nullptr) {  // one coalesced copy of the prebuilt arena instead of the per-element gather (17 % of a call)
      for (int i = threadIdx.x; i < M.smem_w_floats; i += NT) wsm[i] = __ldg(a.wimg + i);
    } else {
      load_weights_smem<NT>(M, wsm, a.theta);
    }
  }
  __syncthreads();
  PhaseCache pc;
  build_phase_cache<WS, CT, NT>(M, pc);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int col0 = tile * CT;
    // coalesced loads: row (q, k) of the tile = CT consecutive floats of field q at level k
    for (int i = threadIdx.x; i < 3 * N * CT; i += NT) {
      const int row = i / CT, c = i - row * CT;
      const int q = row / N, k = row - q * N;
      const int col = min(col0 + c, a.ncol - 1);
      const float v = __ldg(a.f[q] + (size_t)k * a.ncol + col);
      st[i] = v;
      xin[i] = (v - cd.mu[q]) * cd.inv_sig[q];
    }
    __syncthreads();
    for (int p = 0; p < M.n_phase; ++p) {
      run_phase_cached<CT, NT, WS, false>(M, p, pc, xin, arena, nullptr, wsm, a.theta);
      __syncthreads();
    }
    // F = [0; unscaled NN - shift; top flux]; output dF/dz at the centres
    for (int i = threadIdx.x; i < 3 * N * CT; i += NT) {
      const int row = i / CT, c = i - row * CT;
      if (col0 + c >= a.ncol) continue;
      const int q = row / N, k = row - q * N;
      const float* nn = arena + M.nn_off[q] * CT;
      const float sg = cd.sig[3 + q], mu = cd.mu[3 + q];
      const float u0 = fmaf(sg, nn[c], mu);
      const float shift = q < 2 ? fmaf(sg, u0, mu) : u0;  // :292,301 (momentum, literal) / :324 (temperature)
      const float lo = k == 0 ? 0.f : fmaf(sg, nn[(k - 1) * CT + c], mu) - shift;
      const float hi = k == N - 1 ? cd.top[q] : fmaf(sg, nn[k * CT + c], mu) - shift;
      a.dzf[((size_t)q * N + k) * a.ncol + col0 + c] = (hi - lo) * cd.inv_dz;
    }
    __syncthreads();  // the d' sweep below reuses xin
    // face diffusivities nu and nu_T of every interior face, all threads (the Thomas sweeps below only run the recurrence)
    for (int i = threadIdx.x; i < (N - 1) * CT; i += NT) {
      const int f = 1 + i / CT, c = i - (f - 1) * CT;
      const float du = (st[f * CT + c] - st[(f - 1) * CT + c]) * cd.inv_dz;
      const float dv = (st[(N + f) * CT + c] - st[(N + f - 1) * CT + c]) * cd.inv_dz;
      const float dT = (st[(2 * N + f) * CT + c] - st[(2 * N + f - 1) * CT + c]) * cd.inv_dz;
      // MUFU.RCP keeps the IEEE corner cases of the reference's division: +-Inf at zero shear, NaN at 0/0
      const float Ri = cd.g_alpha * dT * uvt_rcp(du * du + dv * dv);
      const float nu = cd.nu0 + cd.nu_m * (0.5f * (1.f - tanhf((Ri - cd.Ric) * cd.inv_dRi)));
      nus[f * CT + c] = nu;
      nus[(N + f) * CT + c] = (cd.ca && !(Ri > 0.f)) ? cd.kappa_ca : nu * cd.inv_Pr;
    }
    __syncthreads();
    // backward-Euler mPP step: one thread per (field, column)
    if (threadIdx.x < 3 * CT) {
      const int q = threadIdx.x / CT, c = threadIdx.x - q * CT;
      const float* sq = st + q * N * CT + c;
      const float* nv = nus + (q == 2 ? N * CT : 0) + c;
      float* dpv = xin + q * N * CT + c;
      float* cpv = cp + q * N * CT + c;
      // diffusivity on face f (between levels f-1 and f); 0 on the boundary faces
      auto nuq = [&](int f) -> float { return (f <= 0 || f >= N) ? 0.f : nv[f * CT]; };
      // incremental form: L (x' - x) = x - L x = r nu_k (x_{k-1} - x_k) + r nu_{k+1} (x_{k+1} - x_k); the small right-hand side keeps
      // the rounding of the sweep relative to the increment, not to the state
      float nk = nuq(0), nn1 = nuq(1);
      float diag = 1.f + cd.r * (nk + nn1);
      float cprev = uvt_div(-cd.r * nn1, diag);
      float xk = sq[0], xup = sq[CT];
      float dprev = uvt_div(cd.r * nn1 * (xup - xk), diag);
      cpv[0] = cprev;
      dpv[0] = dprev;
      for (int k = 1; k < N; ++k) {
        nk = nn1;
        nn1 = nuq(k + 1);
        const float xdn = xk;
        xk = xup;
        xup = k + 1 < N ? sq[(k + 1) * CT] : xk;
        const float lo = -cd.r * nk;
        diag = 1.f + cd.r * (nk + nn1);
        const float den = diag - lo * cprev;
        cprev = uvt_div(-cd.r * nn1, den);
        dprev = uvt_div(cd.r * (nk * (xdn - xk) + nn1 * (xup - xk)) - lo * dprev, den);
        cpv[k * CT] = cprev;
        dpv[k * CT] = dprev;
      }
      const bool live = col0 + c < a.ncol;
      float* o = a.out + (size_t)q * N * a.ncol + col0 + c;
      float dn = dprev;  // increment of the top level
      if (live) o[(size_t)(N - 1) * a.ncol] = sq[(N - 1) * CT] + dn;
      for (int k = N - 2; k >= 0; --k) {
        dn = dpv[k * CT] - cpv[k * CT] * dn;
        if (live) o[(size_t)k * a.ncol] = (q == 2 && k == 0) ? sq[0] : sq[k * CT] + dn;  // T'[1] = T_bottom (:93)
      }
    }
    __syncthreads();
  }
}

}  // namespace cpz

// Per-step NN closure inside a 3-D host model (BASELINE config 5): implicit convective adjustment followed by the
// neural-network temperature-flux forcing, for every column of a (Nx,Ny,Nz) field. Replaces convective_adjustment!
// and compute_neural_network_forcing! (free_convection/double_gyre_nn.jl:27-62,149-168) called from the host model's
// per-iteration callback (:211-217).
//
// The field is x-fastest, so CT consecutive columns at one level are CT consecutive floats: global rows map 1:1 onto
// the [level][column] shared-memory layout of the MLP with fully coalesced 128-byte accesses and no transpose.
#pragma once
#include "cpz_device.cuh"
#include "cpz_solve.cuh"

namespace cpz {

struct ClosureD {
  int Nx, Ny, Nz;
  float inv_dz, r;         // 1/dz, dt/dz^2
  float K;
  float T_shift, inv_T_div;
  float mu_relax, T_mid, dT_over_Ly;
  float mu_T, inv_sig_T, sig_wT, mu_wT;
};

struct ClosureArgs {
  const float* theta;
  const float* T;     // [Nz][Ny*Nx]
  const float* y;     // [Ny]
  float* forcing;     // [Nz][Ny*Nx]
  float* T_out;       // [Nz][Ny*Nx]
  int ncol, n_tiles;
};

struct ClosureSmem {
  int w, tt, xin, arena, cp, model, total_floats;
};
__host__ __device__ inline ClosureSmem closure_smem_layout(const ModelD& M, int CT) {
  ClosureSmem L;
  int o = 0;
  L.w = o; o += M.w_in_smem ? M.smem_w_floats : 0;
  L.tt = o; o += M.Nz * CT;       // temperature (deg C), adjusted in place
  L.xin = o; o += M.Nz * CT;      // scaled NN input
  L.arena = o; o += M.arena_floats * CT;
  L.cp = o; o += M.Nz * CT;       // Thomas c' coefficients
  L.model = o; o += (int)((sizeof(ModelD) + 15) / 16) * 4;
  L.total_floats = o + 4;
  return L;
}

template <int CT, int NT, bool WS>
__global__ void __launch_bounds__(NT, 1) closure_kernel(const __grid_constant__ ModelD Mp, const ClosureD cd, const ClosureArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ClosureSmem L = closure_smem_layout(Mp, CT);
  const ModelD& M = model_to_smem<NT>(Mp, smem + L.model);
  float* wsm = smem + L.w;
  float* tt = smem + L.tt;
  float* xin = smem + L.xin;
  float* arena = smem + L.arena;
  float* cp = smem + L.cp;
  const int N = M.Nz;
  if (WS) load_weights_smem<NT>(M, wsm, a.theta);
  __syncthreads();
  PhaseCache pc;
  build_phase_cache<WS, CT, NT>(M, pc);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int col0 = tile * CT;
    // coalesced load: row k of the tile = CT consecutive floats of level k
    for (int i = threadIdx.x; i < N * CT; i += NT) {
      const int k = i / CT, c = i - k * CT;
      const int col = min(col0 + c, a.ncol - 1);
      tt[i] = __ldg(a.T + (size_t)k * a.ncol + col);
    }
    __syncthreads();
    // implicit convective adjustment (oceananigans_nn.jl:13-40): one thread per column, Thomas algorithm.
    // kappa_k = K where the centre-located dT/dz (mean of the two adjacent face gradients, zero at the boundary
    // faces: free_convection/convective_adjustment.jl:106-112) is negative.
    if (threadIdx.x < CT) {
      const int c = threadIdx.x;
      auto G = [&](int f) -> float { return (f <= 0 || f >= N) ? 0.f : (tt[f * CT + c] - tt[(f - 1) * CT + c]) * cd.inv_dz; };
      auto kap = [&](int k) -> float { return (0.5f * (G(k) + G(k + 1)) < 0.f) ? cd.K : 0.f; };
      // row k: lower = -r*kap[k] (k>=1), upper = -r*kap[k+1] (k<=N-2), diag = 1 + r*(kap[k]+kap[k+1]) (k<N-1), 1 + r*kap[N-1]
      float kk = kap(0), kn = kap(1);
      float diag = 1.f + cd.r * (kk + kn);
      float up = -cd.r * kn;
      float cprev = up / diag;
      float dprev = tt[c] / diag;
      cp[c] = cprev;
      float dsave_prev = dprev;
      // forward sweep stores d' in xin (scratch) and c' in cp
      xin[c] = dprev;
      for (int k = 1; k < N; ++k) {
        kk = kn;
        kn = (k + 1 < N) ? kap(k + 1) : 0.f;
        const float lo = -cd.r * kk;
        diag = (k < N - 1) ? 1.f + cd.r * (kk + kn) : 1.f + cd.r * kk;
        up = (k < N - 1) ? -cd.r * kn : 0.f;
        const float den = diag - lo * cprev;
        cprev = up / den;
        dprev = (tt[k * CT + c] - lo * dsave_prev) / den;
        dsave_prev = dprev;
        cp[k * CT + c] = cprev;
        xin[k * CT + c] = dprev;
      }
      float tn = xin[(N - 1) * CT + c];
      tt[(N - 1) * CT + c] = tn;
      for (int k = N - 2; k >= 0; --k) {
        tn = xin[k * CT + c] - cp[k * CT + c] * tn;
        tt[k * CT + c] = tn;
      }
    }
    __syncthreads();
    // NN input: T_scaling(T_shift + T/T_div)   (double_gyre_nn.jl:155-158); also write the adjusted T
    for (int i = threadIdx.x; i < N * CT; i += NT) {
      const int k = i / CT, c = i - k * CT;
      const float tv = tt[i];
      xin[i] = ((cd.T_shift + tv * cd.inv_T_div) - cd.mu_T) * cd.inv_sig_T;
      if (col0 + c < a.ncol) a.T_out[(size_t)k * a.ncol + col0 + c] = tv;
    }
    __syncthreads();
    for (int p = 0; p < M.n_phase; ++p) {
      run_phase_cached<CT, NT, WS, false>(M, p, pc, xin, arena, nullptr, wsm, a.theta);
      __syncthreads();
    }
    // wT = [0; inv(wT_scaling)(NN); surface_flux]; forcing = d(wT)/dz at centres (double_gyre_nn.jl:159-166,140-147)
    const float* nn = arena + M.nn_off[0] * CT;
    for (int i = threadIdx.x; i < N * CT; i += NT) {
      const int k = i / CT, c = i - k * CT;
      if (col0 + c >= a.ncol) continue;
      const int col = col0 + c;
      const float lo = (k == 0) ? 0.f : cd.sig_wT * nn[(k - 1) * CT + c] + cd.mu_wT;
      float hi;
      if (k == N - 1) {
        const int jy = col / cd.Nx;
        const float T_ref = cd.T_mid + cd.dT_over_Ly * __ldg(a.y + jy);
        hi = -cd.mu_relax * (tt[(N - 1) * CT + c] - T_ref);
      } else {
        hi = cd.sig_wT * nn[k * CT + c] + cd.mu_wT;
      }
      a.forcing[(size_t)k * a.ncol + col] = (hi - lo) * cd.inv_dz;
    }
    __syncthreads();
  }
}

}  // namespace cpz

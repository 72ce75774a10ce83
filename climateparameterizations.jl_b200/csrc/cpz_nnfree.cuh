// NN-free forward solve: the modified Pacanowski–Philander column model with the networks identically zero
// (`DE`, wind_mixing/src/diffusivity_parameter_optimisation.jl:1-33; the RHS of NDE_training.jl:83-165 /
// training_postprocessing.jl:105-153 without the NN terms). This is the HBM-fair variant of the benchmark (SURVEY 8d,
// "2-base"): ~7 kflop and 384 B per column-step.
//
// One WARP integrates NC columns on its own: lane <-> level (Nz = 32), the three fields of NC columns in registers, the
// D^f / D^c stencils are lane shuffles, the Runge–Kutta stage slots are thread-private shared memory. There is no block
// barrier and no shared state between warps, so an SM hides latency with up to 32 independent warps, and a frame is
// written as three 128-byte row segments per column straight from registers.
#pragma once
#include "cpz_tc.cuh"

namespace cpz {

constexpr int NF_WARPS = 4;  // warps per CTA

template <int NC>
__global__ void __launch_bounds__(NF_WARPS * 32) solve_nnfree_kernel(const __grid_constant__ ModelD M, const __grid_constant__ SideC SC,
                                                                    const int side_mode, const __grid_constant__ TableauD tab,
                                                                    const TimeD tm, const SolveArgs a) {
  extern __shared__ __align__(16) float ks_smem[];  // [stage][thread][3][NC]
  const int lane = threadIdx.x & 31;
  const int wglobal = blockIdx.x * NF_WARPS + (threadIdx.x >> 5);
  const int c0 = wglobal * NC;  // first column of this warp
  // a checkpointing solve fills every column slot of the last 32-column checkpoint tile (duplicates of the last column):
  // the adjoint's tiles read them, and 0 * (uninitialised NaN) would poison its sums
  if (c0 >= (a.ckpt != nullptr ? ((a.ncol + 31) & ~31) : a.ncol)) return;
  float* ks = ks_smem + (size_t)threadIdx.x * 3 * NC;
  const int ks_stride = NF_WARPS * 32 * 3 * NC;
  const size_t xs = a.x0_stride ? a.x0_stride : (size_t)96;
  const bool diurnal = (M.flags & F_DIURNAL) != 0;

  float x[3][NC], X[3][NC], bnd[3][NC], Qd[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = min(c0 + c, a.ncol - 1);
    float raw[6], eff[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) raw[j] = __ldg(a.bcs + (size_t)col * 6 + j);
    bc_effective(M, raw, eff);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      x[q][c] = __ldg(a.x0 + (size_t)col * xs + 32 * q + lane);
      X[q][c] = x[q][c];
      bnd[q][c] = lane == 0 ? eff[2 * q] : eff[2 * q + 1];  // bottom flux for lane 0, top flux for lane 31
    }
    Qd[c] = (diurnal && a.Q != nullptr) ? __ldg(a.Q + col) : 0.f;
  }
  const float Au = M.rc.A[0] * M.rc.Nf, Av = M.rc.A[1] * M.rc.Nf, AT = M.rc.A[2] * M.rc.Nf;
  // CPZ_FLAG_IMPLICIT_DIFFUSION: the stages see no diffusive flux; the diffusivities act in a backward-Euler solve per sub-step
  const bool implicit = (M.flags & F_IMPLICIT) != 0;
  const int rhs_mode = implicit ? (int)SIDE_NONE : side_mode;
  auto diffusivity = [&](int mode, float du, float dv, float dT, float& Du, float& Dv, float& DT) {
    switch (mode) {
      case SIDE_MPP: side_column<SIDE_MPP>(SC, du, dv, dT, Du, Dv, DT); break;
      case SIDE_MPP_CA_T: side_column<SIDE_MPP_CA_T>(SC, du, dv, dT, Du, Dv, DT); break;
      case SIDE_MPP_CA_U: side_column<SIDE_MPP_CA_U>(SC, du, dv, dT, Du, Dv, DT); break;
      case SIDE_CA_ONLY: side_column<SIDE_CA_ONLY>(SC, du, dv, dT, Du, Dv, DT); break;
      default: side_column<SIDE_NONE>(SC, du, dv, dT, Du, Dv, DT); break;
    }
  };
  // x <- L(nu(x)) \ x per field: lane <-> level, parallel cyclic reduction across the warp (see pcr32)
  auto implicit_step = [&](float hsub) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float u = x[0][c], v = x[1][c], T = x[2][c];
      const float du = __shfl_down_sync(0xffffffffu, u, 1) - u;
      const float dv = __shfl_down_sync(0xffffffffu, v, 1) - v;
      const float dT = __shfl_down_sync(0xffffffffu, T, 1) - T;
      float D[3];
      diffusivity(side_mode, du, dv, dT, D[0], D[1], D[2]);
      const float Aq[3] = {Au, Av, AT};
      const float dq[3] = {du, dv, dT};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float rup = lane == 31 ? 0.f : hsub * Aq[q] * D[q];
        float rdn = __shfl_up_sync(0xffffffffu, rup, 1);
        const float dqdn = __shfl_up_sync(0xffffffffu, dq[q], 1);
        if (lane == 0) rdn = 0.f;
        // incremental form (see implicit_step in cpz_tc.cuh): L (x' - x) = r_up (x_up - x) - r_dn (x - x_dn)
        x[q][c] += pcr32(-rdn, 1.f + rdn + rup, -rup, rup * dq[q] - rdn * dqdn);
        X[q][c] = x[q][c];
      }
    }
  };

  auto rhs = [&](float t_stage, float (&dx)[3][NC]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float u = X[0][c], v = X[1][c], T = X[2][c];
      const float du = __shfl_down_sync(0xffffffffu, u, 1) - u;
      const float dv = __shfl_down_sync(0xffffffffu, v, 1) - v;
      const float dT = __shfl_down_sync(0xffffffffu, T, 1) - T;
      float Du, Dv, DT;
      diffusivity(rhs_mode, du, dv, dT, Du, Dv, DT);
      float bT = bnd[2][c];
      if (diurnal && lane == 31) bT = diurnal_top_eff(M, Qd[c], t_stage);
      // flux at face lane+1 (top boundary flux for lane 31), flux at face lane from the lane below (bottom flux for lane 0)
      float Fu = -Du * du, Fv = -Dv * dv, FT = -DT * dT;
      if (lane == 31) { Fu = bnd[0][c]; Fv = bnd[1][c]; FT = bT; }
      float Gu = __shfl_up_sync(0xffffffffu, Fu, 1), Gv = __shfl_up_sync(0xffffffffu, Fv, 1), GT = __shfl_up_sync(0xffffffffu, FT, 1);
      if (lane == 0) { Gu = bnd[0][c]; Gv = bnd[1][c]; GT = bnd[2][c]; }
      dx[0][c] = fmaf(-Au, Fu - Gu, fmaf(M.rc.cor_u_s, v, M.rc.cor_u_m));
      dx[1][c] = fmaf(-Av, Fv - Gv, -fmaf(M.rc.cor_v_s, u, M.rc.cor_v_m));
      dx[2][c] = -AT * (FT - GT);
    }
  };

  if (a.rhs_only) {
    float dx[3][NC];
    rhs(a.t_rhs, dx);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c0 + c < a.ncol)
#pragma unroll
        for (int q = 0; q < 3; ++q) a.dxdt[(size_t)(c0 + c) * 96 + 32 * q + lane] = dx[q][c];
    return;
  }

  const float h = tm.dt / (float)tm.n_substeps;
  const int ns = tab.n_stages;
  const size_t traj_stride = (size_t)a.n_saved * 96;
  int frame = 0, ci = 0;
  auto save_frame = [&](int fr) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c0 + c < a.ncol)
#pragma unroll
        for (int q = 0; q < 3; ++q) a.traj[(size_t)(c0 + c) * traj_stride + (size_t)fr * 96 + 32 * q + lane] = x[q][c];
  };
  auto save_ckpt = [&](int k) {  // tile-native layout of the adjoint kernel: [tile of 32 columns][n_ckpt][S][32]
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c0 + c, tile = col >> 5, ct = col & 31;
#pragma unroll
      for (int q = 0; q < 3; ++q) a.ckpt[(((size_t)tile * a.n_ckpt + k) * 96 + 32 * q + lane) * 32 + ct] = x[q][c];
    }
  };
  if (a.traj != nullptr && tm.save_stride > 0 && !a.skip_frame0) { save_frame(0); frame = 1; }
  if (a.ckpt != nullptr) { save_ckpt(0); ci = 1; }
  for (int n = 0; n < tm.n_steps; ++n) {
    for (int sub = 0; sub < tm.n_substeps; ++sub) {
      const float tb = tm.t0 + (float)(tm.step0 + n) * tm.dt + (float)sub * h;
      if (implicit) implicit_step(h);
#pragma unroll 1
      for (int i = 0; i < ns; ++i) {
        float dx[3][NC];
        rhs(tb + tab.c[i] * h, dx);
        const bool last = (i + 1 == ns);
        float acc[3][NC];
        const float ci_ = last ? tab.b[i] : tab.a[(i + 1) % CPZ_MAX_STAGES][i];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[q][c] = ci_ * dx[q][c];
#pragma unroll 1
        for (int j = 0; j < i; ++j) {
          const float cj = last ? tab.b[j] : tab.a[(i + 1) % CPZ_MAX_STAGES][j];
          const float* kj = ks + (size_t)j * ks_stride;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[q][c] = fmaf(cj, kj[q * NC + c], acc[q][c]);
        }
        if (!last) {
          float* ki = ks + (size_t)i * ks_stride;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int c = 0; c < NC; ++c) { ki[q * NC + c] = dx[q][c]; X[q][c] = fmaf(h, acc[q][c], x[q][c]); }
        } else {
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int c = 0; c < NC; ++c) { x[q][c] = fmaf(h, acc[q][c], x[q][c]); X[q][c] = x[q][c]; }
        }
      }
    }
    const int step = n + 1;
    const bool do_save = a.traj != nullptr && ((tm.save_stride > 0 && step % tm.save_stride == 0) ||
                                                (tm.save_stride <= 0 && step == tm.n_steps));
    if (do_save) { save_frame(frame); ++frame; }
    if (a.ckpt != nullptr && (step % tm.ckpt_stride == 0 || step == tm.n_steps)) { save_ckpt(ci); ++ci; }
  }
}

}  // namespace cpz

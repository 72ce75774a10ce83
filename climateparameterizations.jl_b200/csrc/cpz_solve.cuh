// Persistent forward-solve kernel: one CTA integrates CT columns through all time steps with the state,
// the Runge–Kutta stage derivatives, the MLP weights and activations resident in shared memory.
#pragma once
#include "cpz_device.cuh"

namespace cpz {

struct SolveArgs {
  const float* theta;  // [P] destructure order
  const float* x0;     // [ncol][S]
  const float* bcs;    // [ncol][nbc]
  const float* Q;      // [ncol] diurnal amplitudes or null
  float* traj;         // [ncol][n_saved][S] or null
  float* ckpt;         // [n_tiles][n_ckpt][S][CT] (tile-native layout) or null
  float* dxdt;         // rhs_only: [ncol][S]
  int ncol;
  int n_saved;
  int n_ckpt;
  int rhs_only;
  float t_rhs;
};

// shared-memory carve-up (floats) used by solve_kernel and by the host to size the launch
struct SolveSmem {
  int w, xs, xa, ks, arena, bcf, qs, total_floats;
};
__host__ __device__ inline SolveSmem solve_smem_layout(const ModelD& M, int CT, int n_stages) {
  SolveSmem L;
  int o = 0;
  L.w = o; o += M.w_in_smem ? M.smem_w_floats : 0;
  L.xs = o; o += M.S * CT;
  L.xa = o; o += CT * (M.S + 4);  // stage input for stages >= 1, doubles as the [CT][S+4] transpose staging buffer
  L.ks = o; o += n_stages * M.S * CT;
  L.arena = o; o += M.arena_floats * CT;
  L.bcf = o; o += M.nbc * CT;
  L.qs = o; o += CT;
  L.total_floats = o + 4;  // + mbarrier (8 B) and padding
  return L;
}

// One RHS evaluation for the tile: MLP phases, face fluxes; leaves E in the arena. `in` is the stage input [S][CT].
template <int CT, int NT, bool WS>
__device__ __forceinline__ void rhs_eval(const ModelD& M, const float* __restrict__ in, float* __restrict__ arena,
                                         const float* __restrict__ wsm, const float* __restrict__ theta,
                                         float* __restrict__ bcf, const float* __restrict__ qs, float t) {
  if (M.flags & F_DIURNAL) {
    if (threadIdx.x < CT) bcf[(M.nbc - 1) * CT + threadIdx.x] = diurnal_top_eff(M, qs[threadIdx.x], t);
  }
  for (int p = 0; p < M.n_phase; ++p) {
    run_phase<WS, CT, NT, false>(M, p, in, arena, nullptr, wsm, theta);
    __syncthreads();
  }
  if (M.n_phase == 0) __syncthreads();
  faces_phase<CT, NT>(M, in, arena, arena + M.flux_off * CT, bcf);
  __syncthreads();
}

template <int CT, int NT, bool WS>
__global__ void __launch_bounds__(NT, 1) solve_kernel(const __grid_constant__ ModelD M, const __grid_constant__ TableauD tab,
                                                      const TimeD tm, const SolveArgs a) {
  extern __shared__ __align__(16) float smem[];
  const SolveSmem L = solve_smem_layout(M, CT, tab.n_stages);
  float* wsm = smem + L.w;
  float* xs = smem + L.xs;
  float* xa = smem + L.xa;
  float* ks = smem + L.ks;
  float* arena = smem + L.arena;
  float* bcf = smem + L.bcf;
  float* qs = smem + L.qs;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((L.total_floats - 4 + 1) & ~1));
  const int S = M.S, N = M.Nz;
  const int SC = S * CT;
  const int tile = blockIdx.x;
  const int col0 = tile * CT;
  uint32_t parity = 0;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (WS) load_weights_smem<NT>(M, wsm, a.theta);
  // boundary fluxes: raw -> effective, stored [nbc][CT]
  if (threadIdx.x < CT) {
    const int col = min(col0 + (int)threadIdx.x, a.ncol - 1);
    float raw[6], eff[6];
    for (int j = 0; j < M.nbc; ++j) raw[j] = __ldg(a.bcs + (size_t)col * M.nbc + j);
    bc_effective(M, raw, eff);
    for (int j = 0; j < M.nbc; ++j) bcf[j * CT + threadIdx.x] = eff[j];
    qs[threadIdx.x] = (a.Q != nullptr) ? __ldg(a.Q + col) : 0.f;
  }
  __syncthreads();
  load_tile<CT, NT>(xs, xa, bar, parity, a.x0, (size_t)S, S, col0, a.ncol);
  __syncthreads();

  const int ns = tab.n_stages;

  if (a.rhs_only) {
    rhs_eval<CT, NT, WS>(M, xs, arena, wsm, a.theta, bcf, qs, a.t_rhs);
    const float* E = arena + M.flux_off * CT;
    for (int i = threadIdx.x; i < N * CT; i += NT) {
      const int k = i / CT, c = i - k * CT;
      for (int q = 0; q < M.nf; ++q) ks[(q * N + k) * CT + c] = tendency(M, E, xs, q, k, c, CT);
    }
    __syncthreads();
    store_tile<CT, NT>(ks, xa, a.dxdt, (size_t)S, S, col0, a.ncol);
    if (threadIdx.x < CT) bulk_wait0();
    return;
  }

  const float h = tm.dt / (float)tm.n_substeps;
  int frame = 0, ci = 0;
  const size_t traj_stride = (size_t)a.n_saved * S;
  // frame 0 = initial condition; checkpoint 0
  if (a.traj != nullptr && tm.save_stride > 0) {
    store_tile<CT, NT>(xs, xa, a.traj, traj_stride, S, col0, a.ncol);
    if (threadIdx.x < CT) bulk_wait_read0();
    frame = 1;
  }
  if (a.ckpt != nullptr) {
    float4* dst = reinterpret_cast<float4*>(a.ckpt + ((size_t)tile * a.n_ckpt + 0) * SC);
    for (int i = threadIdx.x; i < SC / 4; i += NT) dst[i] = reinterpret_cast<const float4*>(xs)[i];
    ci = 1;
  }
  __syncthreads();

  for (int n = 0; n < tm.n_steps; ++n) {
    for (int sub = 0; sub < tm.n_substeps; ++sub) {
      const float tb = tm.t0 + (float)n * tm.dt + (float)sub * h;
      for (int i = 0; i < ns; ++i) {
        const float* in = (i == 0) ? xs : xa;
        rhs_eval<CT, NT, WS>(M, in, arena, wsm, a.theta, bcf, qs, tb + tab.c[i] * h);
        const float* E = arena + M.flux_off * CT;
        const bool last = (i + 1 == ns);
        for (int it = threadIdx.x; it < N * CT; it += NT) {
          const int k = it / CT, c = it - k * CT;
          float dx[3];
          for (int q = 0; q < M.nf; ++q) dx[q] = tendency(M, E, in, q, k, c, CT);
          for (int q = 0; q < M.nf; ++q) {
            const int e = (q * N + k) * CT + c;
            ks[i * SC + e] = dx[q];
            if (!last) {
              float acc = tab.a[i + 1][i] * dx[q];
              for (int j = 0; j < i; ++j) acc = fmaf(tab.a[i + 1][j], ks[j * SC + e], acc);
              xa[e] = fmaf(h, acc, xs[e]);
            } else {
              float acc = tab.b[i] * dx[q];
              for (int j = 0; j < i; ++j) acc = fmaf(tab.b[j], ks[j * SC + e], acc);
              xs[e] = fmaf(h, acc, xs[e]);
            }
          }
        }
        __syncthreads();
      }
    }
    const int step = n + 1;
    const bool do_save = a.traj != nullptr && ((tm.save_stride > 0 && step % tm.save_stride == 0) ||
                                                (tm.save_stride <= 0 && step == tm.n_steps));
    if (do_save) {
      float* dst = a.traj + (size_t)frame * S;
      store_tile<CT, NT>(xs, xa, dst, traj_stride, S, col0, a.ncol);
      if (threadIdx.x < CT) bulk_wait_read0();  // xa is rewritten by the next stage
      ++frame;
    }
    if (a.ckpt != nullptr && (step % tm.ckpt_stride == 0 || step == tm.n_steps)) {
      float4* dst = reinterpret_cast<float4*>(a.ckpt + ((size_t)tile * a.n_ckpt + ci) * SC);
      for (int i = threadIdx.x; i < SC / 4; i += NT) dst[i] = reinterpret_cast<const float4*>(xs)[i];
      ++ci;
    }
  }
  if (threadIdx.x < CT) bulk_wait0();
}

}  // namespace cpz

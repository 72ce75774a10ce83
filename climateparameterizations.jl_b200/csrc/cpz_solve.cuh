// Persistent forward-solve kernel: one CTA integrates CT columns through all time steps with the state,
// the Runge–Kutta stage derivatives, the MLP weights and activations resident in shared memory.
#pragma once
#include "cpz_device.cuh"

namespace cpz {

// Per-stage records the tensor-core adjoint exchanges through HBM (cpz_adjoint_tc.cuh). One record per (32-column tile,
// stage evaluation e of the launch): `rows` x 32 columns stored as [column quad 0..8)[row][4] floats — the canonical
// un-swizzled K-major operand image of tcgen05.mma with K = column (LBO = 16*rows bytes, SBO = 128), so the weight-gradient
// contraction reads a record as an MMA operand without any re-layout.
//   x : stage input X_i (row = 32*field + level)           rx rows   written by the forward (segment) pass
//   z1: layer-1 pre-activations (row = net*h1 + o)          r1 rows     "
//   z2: layer-2 pre-activations (row = net*h2 + o)          r2 rows     "
//   d1, d2: cotangents of z1, z2; d3: cotangent of the NN fluxes (row = 32*net + j, r3 = 96)   written by the reverse sweep
struct AuxD {
  float *x, *z1, *z2, *d1, *d2, *d3;
  float* xp;     // implicit diffusion: the state BEFORE the implicit step of every sub-step (rx rows; record of tile t, sub-step s at t*n_eval/n_stages + s)
  int n_stages;  // stage evaluations per sub-step (xp index = evaluation index / n_stages)
  int rx, r1, r2, r3;
  int n_eval;    // stage evaluations per tile the x / z1 / z2 buffers hold (record of tile t, evaluation e at t*n_eval + e)
  int n_eval_d;  // the same for the d1 / d2 / d3 buffers (they only live from a reverse launch to its weight-gradient launch)
  int ev_skip;   // forward pass: leading stage evaluations of the launch that are NOT stored (record e holds evaluation ev_skip + e)
  int ev0;       // reverse / weight-gradient launch: x / z record of its evaluation e is ev0 + e
};

struct SolveArgs {
  const float* theta;  // [P] destructure order
  const float* x0;     // [ncol][S]
  const float* bcs;    // [ncol][nbc]
  const float* Q;      // [ncol] diurnal amplitudes or null
  float* traj;         // [ncol][n_saved][S] or null
  float* ckpt;         // [n_tiles32][n_ckpt][S][32] (32-column-tile layout whatever the kernel's tile width) or null
  float* dxdt;         // rhs_only == 1: [ncol][S] tendencies; rhs_only == 2: [ncol][nf][Nz+1] total face fluxes (predict_flux)
  int ncol;
  int n_saved;
  int n_ckpt;
  int rhs_only;
  float t_rhs;
  unsigned long long* prof;  // optional [8] cycle counters (CTA 0, thread 0): set CPZ_PROF=1 in the environment
  size_t x0_stride;          // floats between consecutive columns of x0 (0 = S); a trajectory frame can be the start state
  int skip_frame0;           // continuation of a chunked solve: the start state is already in the trajectory, do not store it
  int small_tiles;           // host hint: a checkpointing solve whose adjoint runs on tiles of this many (4, 8, 16) columns; 0 = 32
  float* kstore;             // optional [n_tiles][n_rk_steps][n_stages][S][32]: every stage tendency k_i, for the adjoint (tcgen05 solve only)
  const float* x0_tile;      // tcgen05 solve only: start state in the checkpoint layout, tile t at x0_tile + t*x0_tile_stride ([S][32]); overrides x0
  size_t x0_tile_stride;
  AuxD aux;                  // tcgen05 solve only: aux.x != null stores X_i, z1, z2 of every stage evaluation (segment pass of the adjoint)
  int split;                 // tcgen05 segment pass only: two CTAs per 32-column tile, CTA b integrates column group b & 1 of tile b >> 1
};

#define CPZ_PROF_BEGIN() const long long prof_t0__ = (a.prof && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0
#define CPZ_PROF_END(slot)                                                                     \
  do {                                                                                         \
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[slot] += (unsigned long long)(clock64() - prof_t0__); \
  } while (0)

// shared-memory carve-up (floats) used by solve_kernel and by the host to size the launch
struct SolveSmem {
  int w, buf, bufsz, ks, arena, bcf, qs, model, total_floats;
};
__host__ __device__ inline SolveSmem solve_smem_layout(const ModelD& M, int CT, int n_stages) {
  SolveSmem L;
  int o = 0;
  L.w = o; o += M.w_in_smem ? M.smem_w_floats : 0;
  // three rotating [S][CT] buffers (step state, two alternating stage inputs); each is large enough to double as the
  // [CT][S+4] transpose staging buffer of the bulk copies
  L.bufsz = CT * (M.S + 4);
  L.buf = o; o += 3 * L.bufsz;
  // Runge–Kutta stage slots; the implicit-diffusion step borrows two of them as scratch at the start of a sub-step
  L.ks = o; o += (((M.flags & F_IMPLICIT) && n_stages < 2) ? 2 : n_stages) * M.S * CT;
  L.arena = o; o += M.arena_floats * CT;
  L.bcf = o; o += M.nbc * CT;
  L.qs = o; o += CT;
  L.model = o; o += (int)((sizeof(ModelD) + 15) / 16) * 4;
  L.total_floats = o + 4;  // + mbarrier (8 B) and padding
  return L;
}

struct StageCoef { float c[CPZ_MAX_STAGES]; };

// MLP part of one RHS evaluation (each phase ends with a block barrier); also refreshes the diurnal top flux.
template <int CT, int NT, bool WS, bool SAVE_Z>
__device__ __forceinline__ void run_phase_cached(const ModelD& M, int p, const PhaseCache& pc, const float* __restrict__ in,
                                                 float* __restrict__ arena, float* __restrict__ zarena,
                                                 const float* __restrict__ wsm, const float* __restrict__ theta) {
  switch (p) {  // static indices keep the cache in registers
    case 0: run_phase<WS, CT, NT, SAVE_Z>(M, 0, pc.tp[0], pc.shape[0], pc.rounds[0], in, arena, zarena, wsm, theta); break;
    case 1: run_phase<WS, CT, NT, SAVE_Z>(M, 1, pc.tp[1], pc.shape[1], pc.rounds[1], in, arena, zarena, wsm, theta); break;
    case 2: run_phase<WS, CT, NT, SAVE_Z>(M, 2, pc.tp[2], pc.shape[2], pc.rounds[2], in, arena, zarena, wsm, theta); break;
    case 3: run_phase<WS, CT, NT, SAVE_Z>(M, 3, pc.tp[3], pc.shape[3], pc.rounds[3], in, arena, zarena, wsm, theta); break;
    default: run_phase<WS, CT, NT, SAVE_Z>(M, p, pc.tp[0], 0, -1, in, arena, zarena, wsm, theta); break;
  }
}

template <int CT, int NT, bool WS>
__device__ __forceinline__ void rhs_mlp(const ModelD& M, const PhaseCache& pc, const float* __restrict__ in,
                                        float* __restrict__ arena, const float* __restrict__ wsm,
                                        const float* __restrict__ theta, float* __restrict__ bcf,
                                        const float* __restrict__ qs, float t) {
  if (M.flags & F_DIURNAL) {
    if (threadIdx.x < CT) bcf[(M.nbc - 1) * CT + threadIdx.x] = diurnal_top_eff(M, qs[threadIdx.x], t);
  }
  for (int p = 0; p < M.n_phase; ++p) {
    run_phase_cached<CT, NT, WS, false>(M, p, pc, in, arena, nullptr, wsm, theta);
    __syncthreads();
  }
  if (M.n_phase == 0 && (M.flags & F_DIURNAL)) __syncthreads();
}

// Tendencies of one RHS evaluation handed to sink(k0, c, dx[NF][4]) (fused stencil), or — for the smoothing variants —
// computed through the two-phase faces/centres path with the flux scratch in the arena.
template <int CT, int NT, int NF, class Sink>
__device__ __forceinline__ void rhs_tendencies(const ModelD& M, const float* __restrict__ in, float* __restrict__ arena,
                                               const float* __restrict__ bcf, Sink sink) {
  const bool smoothing = M.variant == RHS_TRAIN && (M.flags & (F_SMOOTH_NN | F_SMOOTH_RI));
  if (!smoothing) {
    stencil_fused<CT, NT, NF>(M, in, arena, bcf, sink);
    return;
  }
  float* E = arena + M.flux_off * CT;
  faces_phase<CT, NT>(M, in, arena, E, bcf);
  __syncthreads();
  const int N = M.Nz;
  for (int it = threadIdx.x; it < (N / 4) * CT; it += NT) {
    const int kg = it / CT, c = it - kg * CT;
    float dx[NF][4];
#pragma unroll
    for (int q = 0; q < NF; ++q)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) dx[q][kk] = tendency(M, E, in, q, 4 * kg + kk, c, CT);
    sink(4 * kg, c, dx);
  }
}

// M: shared-memory copy of the model (for the non-inlined phase functions); Mp: the kernel parameter itself, whose
// fields the inlined stencil reads as constant-bank operands.
template <int CT, int NT, bool WS, int NF>
__device__ __forceinline__ void solve_body(const ModelD& M, const ModelD& Mp, const TableauD& tab, const TimeD& tm,
                                           const SolveArgs& a, float* smem) {
  const SolveSmem L = solve_smem_layout(M, CT, tab.n_stages);
  float* wsm = smem + L.w;
  float* xs = smem + L.buf;                 // step state
  float* xa = smem + L.buf + L.bufsz;       // stage input (alternates with xb)
  float* xb = smem + L.buf + 2 * L.bufsz;
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
#pragma unroll
    for (int j = 0; j < 6; ++j) raw[j] = j < M.nbc ? __ldg(a.bcs + (size_t)col * M.nbc + j) : 0.f;
    bc_effective(M, raw, eff);
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j < M.nbc) bcf[j * CT + threadIdx.x] = eff[j];
    qs[threadIdx.x] = (a.Q != nullptr) ? __ldg(a.Q + col) : 0.f;
  }
  __syncthreads();
  load_tile<CT, NT>(xs, xa, bar, parity, a.x0, a.x0_stride ? a.x0_stride : (size_t)S, S, col0, a.ncol);
  __syncthreads();

  const int ns = tab.n_stages;
  PhaseCache pc;
  build_phase_cache<WS, CT, NT>(M, pc);

  if (a.rhs_only == 2) {
    // predict_flux (NDE_training.jl:83-147; the wT reconstruction of free_convection/src/solve.jl:35-48): the total face
    // fluxes E_q[0..Nz] whose cell difference gives the tendency. Needs a plan with the face-flux scratch rows (the adjoint plan).
    rhs_mlp<CT, NT, WS>(M, pc, xs, arena, wsm, a.theta, bcf, qs, a.t_rhs);
    float* E = arena + M.flux_off * CT;
    faces_phase<CT, NT>(M, xs, arena, E, bcf);
    __syncthreads();
    const int rows = M.nf * (N + 1);
    for (int i = threadIdx.x; i < rows * CT; i += NT) {
      const int c = i % CT, r = i / CT;
      if (col0 + c < a.ncol) a.dxdt[(size_t)(col0 + c) * rows + r] = E[r * CT + c];
    }
    return;
  }
  if (a.rhs_only) {
    rhs_mlp<CT, NT, WS>(M, pc, xs, arena, wsm, a.theta, bcf, qs, a.t_rhs);
    rhs_tendencies<CT, NT, NF>(Mp, xs, arena, bcf, [=](int k0, int c, const float (&dx)[NF][4]) {
#pragma unroll
      for (int q = 0; q < NF; ++q)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ks[(q * N + k0 + kk) * CT + c] = dx[q][kk];
    });
    __syncthreads();
    store_tile<CT, NT>(ks, xa, a.dxdt, (size_t)S, S, col0, a.ncol);
    if (threadIdx.x < CT) bulk_wait0();
    return;
  }

  const float h = tm.dt / (float)tm.n_substeps;
  int frame = 0, ci = 0;
  const size_t traj_stride = (size_t)a.n_saved * S;
  // frame 0 = initial condition; checkpoint 0
  if (a.traj != nullptr && tm.save_stride > 0 && !a.skip_frame0) {
    store_tile<CT, NT>(xs, xa, a.traj, traj_stride, S, col0, a.ncol);
    if (threadIdx.x < CT) bulk_wait_read0();
    frame = 1;
  }
  // checkpoints go to the global 32-column-tile layout [tile32][n_ckpt][S][32] shared with the adjoint kernel
  auto save_ckpt = [&](int c, const float* src) {
    constexpr int Q = CT / 4;
    float* base = a.ckpt + ((size_t)(col0 / 32) * a.n_ckpt + c) * (size_t)S * 32 + (col0 % 32);
    for (int i = threadIdx.x; i < SC / 4; i += NT)
      *(reinterpret_cast<float4*>(base + (size_t)(i / Q) * 32) + (i % Q)) = reinterpret_cast<const float4*>(src)[i];
  };
  if (a.ckpt != nullptr) { save_ckpt(0, xs); ci = 1; }
  __syncthreads();

  const long long prof_all0 = (a.prof && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
  for (int n = 0; n < tm.n_steps; ++n) {
    for (int sub = 0; sub < tm.n_substeps; ++sub) {
      const float tb = tm.t0 + (float)(tm.step0 + n) * tm.dt + (float)sub * h;
      if (M.flags & F_IMPLICIT) {  // backward-Euler diffusion with the incoming state's diffusivities, then the explicit step
        implicit_diffusion_tile<CT, NT>(M, xs, ks, h);
        __syncthreads();
      }
      const float* in = xs;
      for (int i = 0; i < ns; ++i) {
        if (a.prof) {
          for (int p = 0; p < M.n_phase; ++p) {
            CPZ_PROF_BEGIN();
            run_phase_cached<CT, NT, WS, false>(M, p, pc, in, arena, nullptr, wsm, a.theta);
            __syncthreads();
            CPZ_PROF_END(p < 3 ? p : 3);
          }
        } else {
          rhs_mlp<CT, NT, WS>(M, pc, in, arena, wsm, a.theta, bcf, qs, tb + tab.c[i] * h);
        }
        const bool last = (i + 1 == ns);
        CPZ_PROF_BEGIN();
        // the stencil of other threads still reads `in`, so the next stage input / new state goes to a different buffer
        float* out = last ? ((in == xs) ? xa : xs) : ((in == xa) ? xb : xa);
        StageCoef sc;
#pragma unroll
        for (int j = 0; j < CPZ_MAX_STAGES; ++j) sc.c[j] = last ? tab.b[j] : tab.a[(i + 1) % CPZ_MAX_STAGES][j];
        const float coef_i = last ? tab.b[i] : tab.a[(i + 1) % CPZ_MAX_STAGES][i];
        const float* xs_c = xs;
        rhs_tendencies<CT, NT, NF>(Mp, in, arena, bcf, [=](int k0, int c, const float (&dx)[NF][4]) {
#pragma unroll
          for (int q = 0; q < NF; ++q)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int e = (q * N + k0 + kk) * CT + c;
              float acc = coef_i * dx[q][kk];
#pragma unroll
              for (int j = 0; j < CPZ_MAX_STAGES - 1; ++j)
                if (j < i) acc = fmaf(sc.c[j], ks[j * SC + e], acc);
              if (!last) ks[i * SC + e] = dx[q][kk];
              out[e] = fmaf(h, acc, xs_c[e]);
            }
        });
        __syncthreads();
        CPZ_PROF_END(4);
        if (last && out != xs) {  // single-stage integrator: the new state was written next to the old one
          float* t = xs; xs = out; xa = t;
        }
        in = out;
      }
    }
    const int step = n + 1;
    const bool do_save = a.traj != nullptr && ((tm.save_stride > 0 && step % tm.save_stride == 0) ||
                                                (tm.save_stride <= 0 && step == tm.n_steps));
    if (do_save) {
      CPZ_PROF_BEGIN();
      float* dst = a.traj + (size_t)frame * S;
      store_tile<CT, NT>(xs, xb, dst, traj_stride, S, col0, a.ncol);
      if (threadIdx.x < CT) bulk_wait_read0();  // xb is rewritten by a later stage
      ++frame;
      CPZ_PROF_END(5);
    }
    if (a.ckpt != nullptr && (step % tm.ckpt_stride == 0 || step == tm.n_steps)) { save_ckpt(ci, xs); ++ci; }
  }
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[7] += (unsigned long long)(clock64() - prof_all0);
  if (threadIdx.x < CT) bulk_wait0();
}

// The model description is copied from the kernel parameters into shared memory once: the (non-inlined) phase functions
// take it by reference, and a reference into parameter space would turn every field access into a generic global load.
template <int NT>
__device__ __forceinline__ const ModelD& model_to_smem(const ModelD& Mp, float* dst) {
  const int* src = reinterpret_cast<const int*>(&Mp);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(ModelD) / 4); i += NT) d[i] = src[i];
  __syncthreads();
  return *reinterpret_cast<const ModelD*>(dst);
}

template <int CT, int NT, bool WS>
__global__ void __launch_bounds__(NT, 1) solve_kernel(const __grid_constant__ ModelD Mp, const __grid_constant__ TableauD tab,
                                                      const TimeD tm, const SolveArgs a) {
  extern __shared__ __align__(16) float smem[];
  const SolveSmem L = solve_smem_layout(Mp, CT, tab.n_stages);
  const ModelD& M = model_to_smem<NT>(Mp, smem + L.model);
  if (Mp.nf == 3) solve_body<CT, NT, WS, 3>(M, Mp, tab, tm, a, smem);
  else solve_body<CT, NT, WS, 1>(M, Mp, tab, tm, a, smem);
}

}  // namespace cpz

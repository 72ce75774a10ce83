// tcgen05 forward solve of the T-only free-convection NDEs (FreeConvectionNDE / ConvectiveAdjustmentNDE,
// free_convection/src/free_convection_nde.jl:29-38, convective_adjustment_nde.jl:33-48; BASELINE configs 1 and 4; with
// CPZ_FLAG_MPP also the mPP base diffusivity of wind_mixing/src/NDE_training.jl:114-139 at u = v = 0).
// Same orientation as the closure kernel (cpz_closure_tc.cuh): a CTA owns 128 columns for the whole integration,
// thread t <-> column t <-> TMEM lane t holds the 32-level profile in registers; per RHS evaluation the scaled profile
// goes to TMEM as the A operand (tcgen05.st), the three Dense layers run as 3xTF32 MMAs against the shared-memory
// weight image, and the flux divergence + Runge-Kutta combination happen in the column's own registers. The stage
// derivatives k_j live in an L2-resident global scratch ([tile][stage][level][128 columns], coalesced).
#pragma once
#include "cpz_closure_tc.cuh"

namespace cpz {

// Two 128-column tiles per CTA, one per warp set (4 warps each), run half a phase apart: while one set's tile is in its MLP
// section (the other set takes every other chunk of its hidden-layer epilogues) (operands and accumulators in TMEM: hidden activations hi [0,128) lo [128,256), accumulators [256,384); the
// stage input X sits in the first 32 columns of the activation planes, the layer-3 accumulator in the first 32
// accumulator columns), the other set does its column work in registers (flux divergence, Runge-Kutta combination
// against the global stage slots, frame stores). Phases are separated by CTA barriers; TMEM is owned by one set per phase.
// MPP: the mPP base diffusivity term (BASELINE config 1) is compiled only into its own instantiation, so that the plain
// FreeConvectionNDE / ConvectiveAdjustmentNDE kernel (config 4) carries none of its registers or instructions.
template <int ACT, bool MPP>
__global__ void __launch_bounds__(CTC_NT, 1) solve_fc_tc_kernel(const __grid_constant__ ClosureTcD C, const __grid_constant__ ModelD M,
                                                                const __grid_constant__ TableauD tab, const TimeD tm, const SolveArgs a,
                                                                const float* __restrict__ img, float* __restrict__ kscr, const int n_pair_ctas) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int set = warp >> 2;       // warp set = tile of the pair
  const int ltid = tid & 127;      // column / TMEM lane inside the tile
  constexpr int N = 32;
  uint64_t* bar_w = &bars[0];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar_w, (uint32_t)C.img_bytes);
    for (int o = 0; o < C.img_bytes; o += 32768) bulk_g2s(sm + o, reinterpret_cast<const char*>(img) + o, (uint32_t)min(32768, C.img_bytes - o), bar_w);
  }
  const uint32_t tl = tb + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t HAh = 0, HAl = 128, D12 = 256;
  const float* bias = reinterpret_cast<const float*>(sm + C.o_b);

  auto chain = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t bh_off, uint32_t bl_off, int K, int n) {
    const uint32_t id = tc_idesc(128, n);
    const uint32_t sbo = (uint32_t)(K / 4) * 128;
    const uint64_t bh = ctc_desc(smem_u32(sm) + bh_off, sbo), bl = ctc_desc(smem_u32(sm) + bl_off, sbo);
    const int steps = K / 8;
#pragma unroll 4
    for (int s = 0; s < steps; ++s) tc_mma_ts(d, a_hi + 8 * s, bl + 16 * s, id, s > 0);
#pragma unroll 4
    for (int s = 0; s < steps; ++s) tc_mma_ts(d, a_lo + 8 * s, bh + 16 * s, id, 1);
#pragma unroll 4
    for (int s = 0; s < steps; ++s) tc_mma_ts(d, a_hi + 8 * s, bh + 16 * s, id, 1);
  };
  // hidden-layer epilogue of the tile that owns TMEM: the owner set takes the 32-output chunks 0, 2, the other set 1, 3
  auto hidden = [&](int n_cols, int b_off, int act, int first) {
    for (int j = 32 * first; j < n_cols; j += 64) {
      float v[32], lo[32];
      tmem_ld32(tl + D12 + j, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float z = tc_act<ACT>(act, v[i] + bias[b_off + j + i]);
        const float hi = tf32_hi(z);
        lo[i] = z - hi;
        v[i] = hi;
      }
      tmem_st32(tl + HAh + j, v);
      tmem_st32(tl + HAl + j, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  };
  uint32_t par0 = 0u, par1 = 0u;  // phase parity of the two sets' MMA barriers (every thread follows both)
  auto issue = [&](int owner, int layer) {  // executed by all threads right after a CTA barrier
    if (warp == 4 * owner) {
      tc_fence_after();
      if (elect_one()) {
        if (layer == 0) chain(tb + D12, tb + HAh, tb + HAl, C.o_w1h, C.o_w1l, 32, C.n1);
        else if (layer == 1) chain(tb + D12, tb + HAh, tb + HAl, C.o_w2h, C.o_w2l, C.k2, C.n2);
        else chain(tb + D12, tb + HAh, tb + HAl, C.o_w3h, C.o_w3l, C.k3, C.n3);
        tc_commit(&bars[1 + owner]);
      }
      __syncwarp();
    }
  };
  auto wait_mma = [&](int owner) {
    if (owner == 0) { mbar_wait(&bars[1], par0); par0 ^= 1u; }
    else { mbar_wait(&bars[2], par1); par1 ^= 1u; }
    tc_fence_after();
  };
  auto cta_sync = [&]() { tc_fence_before(); __syncthreads(); tc_fence_after(); };

  // ---- this thread's column ----
  // CTAs [0, n_pair_ctas) own two tiles; the rest (a last partial wave) own one tile and skip the second set's MLP phases
  const bool two_tiles = (int)blockIdx.x < n_pair_ctas;
  const int tile = two_tiles ? 2 * blockIdx.x + set : 2 * n_pair_ctas + ((int)blockIdx.x - n_pair_ctas);
  const int col = tile * CTC_TILE + ltid;
  const int colc = min(col, a.ncol - 1);
  const bool active = two_tiles || set == 0;  // in a single-tile CTA the second warp set only helps with the epilogues
  const bool live = active && col < a.ncol;
  const size_t xs = a.x0_stride ? a.x0_stride : (size_t)N;
  float x[N], X[N], nn[32];
#pragma unroll
  for (int k4 = 0; k4 < N / 4; ++k4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.x0 + (size_t)colc * xs) + k4);
    x[4 * k4] = v.x; x[4 * k4 + 1] = v.y; x[4 * k4 + 2] = v.z; x[4 * k4 + 3] = v.w;
  }
  const float bc_b = __ldg(a.bcs + (size_t)colc * 2), bc_t = __ldg(a.bcs + (size_t)colc * 2 + 1);
#pragma unroll
  for (int k = 0; k < N; ++k) { X[k] = x[k]; nn[k] = 0.f; }
  const bool ca = (M.flags & F_CA) != 0;
  const float ANf = M.rc.A[2] * M.rc.Nf, Nf = M.rc.Nf, Kca = M.rc.K_ca;
  const float h = tm.dt / (float)tm.n_substeps;
  const int ns = tab.n_stages;
  float* ks = kscr + (size_t)tile * ns * N * CTC_TILE + ltid;  // k_j[level] at ks[(j*N + level)*128]
  const size_t traj_stride = (size_t)a.n_saved * N;
  int frame = 0, ci = 0;
  auto save_frame = [&](int fr) {
    if (live) {
      float4* dst = reinterpret_cast<float4*>(a.traj + (size_t)col * traj_stride + (size_t)fr * N);
#pragma unroll
      for (int k4 = 0; k4 < N / 4; ++k4) dst[k4] = make_float4(x[4 * k4], x[4 * k4 + 1], x[4 * k4 + 2], x[4 * k4 + 3]);
    }
  };
  auto save_ckpt = [&](int c) {  // adjoint tile layout: [tile of 32 columns][n_ckpt][S][32]
    if (active && col < ((a.ncol + 31) & ~31)) {
      const int t32 = col >> 5, ct = col & 31;
#pragma unroll
      for (int k = 0; k < N; ++k) a.ckpt[(((size_t)t32 * a.n_ckpt + c) * N + k) * 32 + ct] = x[k];
    }
  };
  if (!a.rhs_only) {
    if (a.traj != nullptr && tm.save_stride > 0 && !a.skip_frame0) { save_frame(0); frame = 1; }
    if (a.ckpt != nullptr) { save_ckpt(0); ci = 1; }
  }
  mbar_wait(bar_w, 0);

  // column work of this thread's tile after its MLP section: tendencies from the NN fluxes in `nn`, Runge-Kutta
  // combination, next stage input (state of the flattened step / sub-step / stage loop in n_, sub_, i_)
  int n_ = 0, sub_ = 0, i_ = 0;
  auto column_work = [&]() {
    float dx[N];
    {
      const float* b3 = bias + C.n1 + C.n2;
      float Elo = bc_b;  // dx[k] = -A Nz (E[k+1] - E[k]), E = [bottom; NN(X) - [CA] min(0, K dT/dz); top]
#pragma unroll
      for (int k = 0; k < N; ++k) {
        float Ehi;
        if (k == N - 1) Ehi = bc_t;
        else {
          Ehi = nn[k] + b3[k];
          if constexpr (MPP) {
            const float G = Nf * (X[k + 1] - X[k]);
            Ehi -= fc_mpp_cnu(M, G) * G;   // mPP base at u = v = 0 (BASELINE config 1)
          }
          if (ca) Ehi -= fminf(0.f, Kca * (Nf * (X[k + 1] - X[k])));
        }
        dx[k] = -ANf * (Ehi - Elo);
        Elo = Ehi;
      }
    }
    if (a.rhs_only) {
      if (live) {
#pragma unroll
        for (int k4 = 0; k4 < N / 4; ++k4)
          reinterpret_cast<float4*>(a.dxdt + (size_t)col * N)[k4] = make_float4(dx[4 * k4], dx[4 * k4 + 1], dx[4 * k4 + 2], dx[4 * k4 + 3]);
      }
      return;
    }
    const int i = i_;
    const bool last = (i + 1 == ns);
    float acc[N];
    const float ci_ = last ? tab.b[i] : tab.a[(i + 1) % CPZ_MAX_STAGES][i];
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = ci_ * dx[k];
#pragma unroll 1
    for (int j = 0; j < i; ++j) {
      const float cj = last ? tab.b[j] : tab.a[(i + 1) % CPZ_MAX_STAGES][j];
      const float* kj = ks + (size_t)j * N * CTC_TILE;
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] = fmaf(cj, kj[k * CTC_TILE], acc[k]);
    }
    if (!last) {
      float* ki = ks + (size_t)i * N * CTC_TILE;
#pragma unroll
      for (int k = 0; k < N; ++k) { ki[k * CTC_TILE] = dx[k]; X[k] = fmaf(h, acc[k], x[k]); }
      ++i_;
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) { x[k] = fmaf(h, acc[k], x[k]); X[k] = x[k]; }
      i_ = 0;
      if (++sub_ == tm.n_substeps) {
        sub_ = 0;
        const int step = ++n_;
        const bool do_save = a.traj != nullptr && ((tm.save_stride > 0 && step % tm.save_stride == 0) ||
                                                    (tm.save_stride <= 0 && step == tm.n_steps));
        if (do_save) { save_frame(frame); ++frame; }
        if (a.ckpt != nullptr && (step % tm.ckpt_stride == 0 || step == tm.n_steps)) { save_ckpt(ci); ++ci; }
      }
    }
  };

  // Phases: in phase p the tile of set p&1 owns TMEM and runs evaluation p>>1; the other set helps with every other
  // epilogue chunk and does the column work of ITS previous evaluation while layer 2 is in the tensor pipe.
  const int K = a.rhs_only ? 1 : tm.n_steps * tm.n_substeps * ns;
#pragma unroll 1
  for (int p = 0; p <= 2 * K; ++p) {
    const int owner = p & 1;
    const bool mine = owner == set;
    const bool has_mlp = (p >> 1) < K && (owner == 0 || two_tiles);  // the owner tile still has an evaluation to run
    const bool has_col = !mine && p >= 1 && ((p - 1) >> 1) < K && (set == 0 || two_tiles);  // column work pending from this set's last MLP section
    if (has_mlp) {
      if (mine) {
        float hi[N], lo[N];
#pragma unroll
        for (int k = 0; k < N; ++k) { hi[k] = tf32_hi(X[k]); lo[k] = X[k] - hi[k]; }
        tmem_st32(tl + HAh, hi);
        tmem_st32(tl + HAl, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      cta_sync();
      issue(owner, 0);
      wait_mma(owner);
      hidden(C.n1, 0, C.act1, mine ? 0 : 1);
      cta_sync();
      issue(owner, 1);
      if (has_col) column_work();
      wait_mma(owner);
      hidden(C.n2, C.n1, C.act2, mine ? 0 : 1);
      cta_sync();
      issue(owner, 2);
      if (mine) {
        wait_mma(owner);
        tmem_ld32(tl + D12, nn);
      } else {
        if (owner == 0) par0 ^= 1u; else par1 ^= 1u;  // the helper does not need the layer-3 result
      }
    } else if (has_col) {
      column_work();
    }
    cta_sync();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

}  // namespace cpz

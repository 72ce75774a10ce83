// tcgen05 version of the per-step NN closure (BASELINE config 5; same contract as closure_kernel in cpz_closure.cuh):
// implicit convective adjustment + neural-network temperature-flux forcing for every column of an (Nx,Ny,Nz) field,
// replacing convective_adjustment! and compute_neural_network_forcing! (free_convection/double_gyre_nn.jl:27-62,149-168).
//
// A field has tens of thousands of columns per GPU, so here the COLUMNS are the M side of the MMA: one CTA owns a tile of
// 128 consecutive columns, thread t (< 128) <-> column t <-> TMEM lane t; a second set of four warps shares the lanes and
// takes every other 32-output chunk of the hidden-layer epilogues. Everything a column needs stays
// in that thread's registers (the 32-level profile, the Thomas solve, the flux divergence); the three Dense layers are
//     D[128 columns][N outputs] = act[128][K] * W[K][N]      (tcgen05.mma kind::tf32, 3xTF32, FP32 accumulators in TMEM)
// with the activations as the A operand IN TENSOR MEMORY (written by tcgen05.st straight from the epilogue registers:
// no shared-memory round trip, no proxy fence) and the weights as the B operand resident in shared memory (hi and lo
// planes, canonical K-major layout, copied once per CTA with cp.async.bulk from a prebuilt image).
// TMEM map: X hi/lo [0,64) | accumulators of layers 1 and 2 [64,192) | hidden activations hi [192,320) lo [320,448) |
// layer-3 accumulator [448,480).
#pragma once
#include "cpz_closure.cuh"
#include "cpz_tc.cuh"

namespace cpz {

struct ClosureTcD {
  int h1, h2, nout;        // layer widths (Nz -> h1 -> h2 -> Nz-1)
  int n1, n2, n3;          // MMA N of the three layers (multiples of 16)
  int k2, k3;              // MMA K of layers 2, 3 (multiples of 8; layer 1 has K = 32)
  int act1, act2;
  int w_off[3], b_off[3];  // theta offsets
  int o_w1h, o_w1l, o_w2h, o_w2l, o_w3h, o_w3l, o_b, img_bytes;  // byte offsets inside the weight image / shared memory
};

// image element (layer l, output n, input k) of plane hi/lo at  (n/8)*SBO_l + (k/4)*128 + (n%8)*16 + (k%4)*4, SBO_l = (K_l/4)*128
static __global__ void closure_tc_image_kernel(const __grid_constant__ ClosureTcD C, const float* __restrict__ theta, float* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int Ks[3] = {32, C.k2, C.k3}, Ns[3] = {C.n1, C.n2, C.n3}, ins[3] = {32, C.h1, C.h2}, outs[3] = {C.h1, C.h2, C.nout};
  const int oh[3] = {C.o_w1h, C.o_w2h, C.o_w3h}, ol[3] = {C.o_w1l, C.o_w2l, C.o_w3l};
  int base = 0;
  for (int l = 0; l < 3; ++l) {
    const int cnt = Ns[l] * Ks[l];
    if (idx >= base && idx < base + cnt) {
      const int e = idx - base, n = e / Ks[l], k = e % Ks[l];
      float w = 0.f;
      if (n < outs[l] && k < ins[l]) w = theta[C.w_off[l] + k * outs[l] + n];
      const float hi = tf32_hi(w);
      const int off = (n / 8) * (Ks[l] / 4) * 128 + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4;
      *reinterpret_cast<float*>(reinterpret_cast<char*>(img) + oh[l] + off) = hi;
      *reinterpret_cast<float*>(reinterpret_cast<char*>(img) + ol[l] + off) = w - hi;
    }
    base += cnt;
  }
  // biases (zero padded): [n1 | n2 | n3]
  const int nb = C.n1 + C.n2 + C.n3;
  if (idx >= base && idx < base + nb) {
    const int j = idx - base;
    float b = 0.f;
    if (j < C.n1) { if (j < C.h1) b = theta[C.b_off[0] + j]; }
    else if (j < C.n1 + C.n2) { if (j - C.n1 < C.h2) b = theta[C.b_off[1] + j - C.n1]; }
    else if (j - C.n1 - C.n2 < C.nout) b = theta[C.b_off[2] + j - C.n1 - C.n2];
    *reinterpret_cast<float*>(reinterpret_cast<char*>(img) + C.o_b + 4 * j) = b;
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}

// K-major, no swizzle B descriptor: LBO = 128 B between the two 4-element K chunks of a K step, SBO between 8-row groups
__device__ __forceinline__ uint64_t ctc_desc(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

constexpr int CTC_TILE = 128;  // columns per tile = TMEM lanes
constexpr int CTC_NT = 256;    // two warps per TMEM lane quadrant: they split the output chunks of the hidden-layer epilogues

template <int ACT>
__global__ void __launch_bounds__(CTC_NT, 1) closure_tc_kernel(const __grid_constant__ ClosureTcD C, const ClosureD cd, const ClosureArgs a,
                                                               const float* __restrict__ img) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int hh = warp >> 2;        // 0: owns the columns (profile, Thomas solve, forcing); 1: epilogue helper
  const int ltid = tid & 127;      // column / TMEM lane inside the tile
  constexpr int N = 32;
  uint64_t* bar_w = &bars[0];    // weight image landed
  uint64_t* bar_mma = &bars[1];  // MMA chain complete
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
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
  if (tid == 0) {  // weights + biases: global image -> shared memory, 32 KB bulk copies
    mbar_expect_tx(bar_w, (uint32_t)C.img_bytes);
    for (int o = 0; o < C.img_bytes; o += 32768) bulk_g2s(sm + o, reinterpret_cast<const char*>(img) + o, (uint32_t)min(32768, C.img_bytes - o), bar_w);
  }
  const uint32_t tl = tb + ((uint32_t)(32 * (warp & 3)) << 16);  // this warp's TMEM lanes
  const uint32_t XA = 0, D12 = 64, HAh = 192, HAl = 320, D3 = 448;
  const float* bias = reinterpret_cast<const float*>(sm + C.o_b);
  uint32_t par = 0;

  // one MMA chain: D[128][n] (+)= A(TMEM, K = 8*steps) * B(smem planes): lo*hi + hi*lo + hi*hi
  auto chain = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t bh_off, uint32_t bl_off, int K, int n) {
    const uint32_t id = tc_idesc(128, n);
    const uint32_t sbo = (uint32_t)(K / 4) * 128;
    const uint64_t bh = ctc_desc(smem_u32(sm) + bh_off, sbo), bl = ctc_desc(smem_u32(sm) + bl_off, sbo);
    const int steps = K / 8;
#pragma unroll 4
    for (int s = 0; s < steps; ++s) tc_mma_ts(d, a_hi + 8 * s, bl + 16 * s, id, s > 0);  // descriptor += 256 B per K step
#pragma unroll 4
    for (int s = 0; s < steps; ++s) tc_mma_ts(d, a_lo + 8 * s, bh + 16 * s, id, 1);
#pragma unroll 4
    for (int s = 0; s < steps; ++s) tc_mma_ts(d, a_hi + 8 * s, bh + 16 * s, id, 1);
  };
  // hidden-layer epilogue: accumulator row -> bias, activation, hi/lo split -> A operand of the next layer
  auto hidden = [&](int n_cols, int b_off, int act) {
    for (int j = 32 * hh; j < n_cols; j += 64) {
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

  // per-tile column state of this warp set (set = hh): the two sets alternate tiles, so that the profile load and the
  // Thomas solve of tile i+1 run while layer 2 of tile i is in the tensor pipe
  float T_top = 0.f;
  int col = 0, colc = 0;
  bool live = false;
  // column part: profile load, implicit convective adjustment, adjusted T out, scaled NN input -> TMEM A operand
  auto column_part = [&](int tile) {
    col = tile * CTC_TILE + ltid;
    colc = min(col, a.ncol - 1);
    live = col < a.ncol;
    float T[N];
#pragma unroll
    for (int k = 0; k < N; ++k) T[k] = __ldg(a.T + (size_t)k * a.ncol + colc);
    // implicit convective adjustment (oceananigans_nn.jl:13-40; free_convection/convective_adjustment.jl:106-129)
    {
      float kap[N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float g0 = (k >= 1) ? (T[k] - T[k - 1]) * cd.inv_dz : 0.f;
        const float g1 = (k + 1 <= N - 1) ? (T[k + 1] - T[k]) * cd.inv_dz : 0.f;
        kap[k] = (0.5f * (g0 + g1) < 0.f) ? cd.K : 0.f;
      }
      float cp[N], dp[N];
      {
        const float inv = rcp_fast(1.f + cd.r * (kap[0] + kap[1]));
        cp[0] = -cd.r * kap[1] * inv;
        dp[0] = T[0] * inv;
      }
#pragma unroll
      for (int k = 1; k < N; ++k) {
        const float lo = -cd.r * kap[k];
        const float kn = (k + 1 < N) ? kap[k + 1] : 0.f;
        const float diag = (k < N - 1) ? 1.f + cd.r * (kap[k] + kn) : 1.f + cd.r * kap[k];
        const float up = (k < N - 1) ? -cd.r * kn : 0.f;
        const float inv = rcp_fast(diag - lo * cp[k - 1]);  // MUFU.RCP (2 ulp): diagonally dominant system, den >= 1
        cp[k] = up * inv;
        dp[k] = (T[k] - lo * dp[k - 1]) * inv;
      }
      T[N - 1] = dp[N - 1];
#pragma unroll
      for (int k = N - 2; k >= 0; --k) T[k] = dp[k] - cp[k] * T[k + 1];
    }
    // adjusted T out; NN input T_scaling(T_shift + T/T_div) (double_gyre_nn.jl:155-158) -> TMEM A operand
    float hi[N], lo[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (live) a.T_out[(size_t)k * a.ncol + col] = T[k];
      const float x = ((cd.T_shift + T[k] * cd.inv_T_div) - cd.mu_T) * cd.inv_sig_T;
      hi[k] = tf32_hi(x);
      lo[k] = x - hi[k];
    }
    tmem_st32(tl + XA, hi);
    tmem_st32(tl + XA + 32, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    T_top = T[N - 1];
  };

  int it = 0;
  if ((int)blockIdx.x < a.n_tiles && hh == 0) column_part(blockIdx.x);
  mbar_wait(bar_w, 0);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
    const int owner = it & 1;
    tc_fence_before();
    __syncthreads();
    // ---- layer 1 ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        chain(tb + D12, tb + XA, tb + XA + 32, C.o_w1h, C.o_w1l, 32, C.n1);
        tc_commit(bar_mma);
      }
      __syncwarp();
    }
    mbar_wait(bar_mma, par); par ^= 1u;
    tc_fence_after();
    hidden(C.n1, 0, C.act1);
    tc_fence_before();
    __syncthreads();
    // ---- layer 2 ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        chain(tb + D12, tb + HAh, tb + HAl, C.o_w2h, C.o_w2l, C.k2, C.n2);
        tc_commit(bar_mma);
      }
      __syncwarp();
    }
    // while layer 2 runs: the other warp set prepares the next tile (layer 1 has released the X operand)
    if (tile + (int)gridDim.x < a.n_tiles && hh == (owner ^ 1)) column_part(tile + gridDim.x);
    mbar_wait(bar_mma, par); par ^= 1u;
    tc_fence_after();
    hidden(C.n2, C.n1, C.act2);
    tc_fence_before();
    __syncthreads();
    // ---- layer 3 ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        chain(tb + D3, tb + HAh, tb + HAl, C.o_w3h, C.o_w3l, C.k3, C.n3);
        tc_commit(bar_mma);
      }
      __syncwarp();
    }
    mbar_wait(bar_mma, par); par ^= 1u;
    tc_fence_after();
    // ---- wT = [0; inv(wT_scaling)(NN); surface_flux]; forcing = d(wT)/dz at centres (double_gyre_nn.jl:159-166,140-147) ----
    if (hh == owner) {
      float nn[32];
      tmem_ld32(tl + D3, nn);
      const float* b3 = bias + C.n1 + C.n2;
      const int jy = colc / cd.Nx;
      const float T_ref = cd.T_mid + cd.dT_over_Ly * __ldg(a.y + jy);
      float lo = 0.f;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        float hi;
        if (k == N - 1) hi = -cd.mu_relax * (T_top - T_ref);
        else hi = cd.sig_wT * (nn[k] + b3[k]) + cd.mu_wT;
        if (live) a.forcing[(size_t)k * a.ncol + col] = (hi - lo) * cd.inv_dz;
        lo = hi;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

}  // namespace cpz

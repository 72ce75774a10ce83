// Tensor-core forward solve: tcgen05 (UTCHMMA) 3xTF32 MLP with TMEM-resident weights, sm_100a.
//
// Same contract as solve_kernel (cpz_solve.cuh) for the u/v/T NDE whose three nets are 96 -> h1 -> h2 -> 31 Dense chains
// (NDE_training.jl:83-165, training_postprocessing.jl:55-153). The MLP of one RHS evaluation is computed as
//     D[out rows][columns] = W^T[out rows][K] * act[K][columns]
// i.e. the WEIGHTS are the M side of the MMA (128 TMEM lanes) and the model COLUMNS are the N side (16 per column
// group), so a CTA needs only 32 columns to fill the tensor pipe and 4096 columns still spread over 128 SMs.
//
//   * precision: 3xTF32. Every operand is split x = hi + lo with hi = tf32(x) (round to nearest) and
//     D = Whi*Xlo + Wlo*Xhi + Whi*Xhi accumulated in FP32 in TMEM: error ~2^-21 per product (measured 3.4e-7 relative
//     on random data, tools/micro/umma_test.cu), inside the 1e-5 RHS tolerance.
//   * A operands (weights) live in TENSOR MEMORY for the whole solve (tcgen05.mma with A from TMEM): measured 10.7
//     cycles per M=128,N=16,K=8 MMA against 41 when A is read from shared memory (4 KB of A per MMA saturates the
//     shared-memory port). TMEM map (512 columns x 128 lanes):
//         [  0, 96) W1 rows   0..127 hi   [ 96,192) lo           (layer 1, the 3 nets stacked: row = net*h1 + o)
//         [192,288) W1 rows 128..     hi  [288,384) lo           in lanes 96.. only (quadrant 3)
//         lanes 0..95 of [192,384): layer-2 stack (row 32q+o = net q), then layer-3 stack, hi and lo each
//         [384,448) accumulators of column group 0, [448,512) of group 1
//     Rows of a stack that belong to other nets / are padding produce garbage accumulator rows that are never read.
//   * B operands (activations) are written by the epilogue threads into shared memory in the canonical no-swizzle
//     K-major layout (8 columns x 16 B core matrices, SBO = 128 B between column octets, LBO = 272 B between K chunks;
//     the 16 B pad makes the per-row scalar stores of a warp conflict free).
//   * two column groups of 16 columns per CTA run the same sequence half a phase apart on their own warps, named
//     barriers and mbarriers, so one group's epilogue / stencil overlaps the other's MMAs.
//   * thread <-> data: warp quadrant q (warp%4 — the TMEM lanes a warp may read) <-> field/net q, lane <-> level
//     (Nz = 32) resp. output row, 8 columns per thread in registers. Quadrant 3 owns the 22 extra layer-1 rows and
//     computes the Richardson-number diffusivities and Coriolis terms while layer 1 is in the tensor pipe.
#pragma once
#include "cpz_solve.cuh"

namespace cpz {

constexpr int TC_NG = 2;     // column groups per CTA
constexpr int TC_GN = 16;    // columns per group = MMA N
constexpr int TC_CT = TC_NG * TC_GN;
constexpr int TC_NT = 512;   // 2 groups x 8 warps
constexpr uint32_t TC_SBO = 128;
constexpr uint32_t TC_LBO = (TC_GN / 8) * 128 + 16;  // 272
constexpr int TC_WCOLS = 384;                         // TMEM columns holding weights
constexpr int TC_SIDE_ARRAYS = 8;                     // Du, Dv, DT, cor_u, cor_v, Xu, Xv, XT

struct TcD {
  int h1, h2, nout;      // per-net layer widths (identical for the three nets)
  int act1, act2, act3;
  int n1b;               // layer-1 rows in block 1 (= 3*h1 - 128, <= 32)
  int k2_start[3], k2_steps;  // K windows (rows of H1) of the three layer-2 MMAs
  int k3_start[3], k3_steps;  // K windows (rows of H2) of the three layer-3 MMAs
  int h1_rows, h2_rows;       // allocated rows (multiples of 4; rows past the real ones stay zero)
  int c_a2hi, c_a2lo, c_a3hi, c_a3lo;  // TMEM columns of the layer-2/3 stacks
  int w_off[3][3], b_off[3][3];        // theta offsets [net][layer]
};

struct TcSmem {  // byte offsets
  int xh, xl, h1h, h1l, h2h, h2l, side, bc, grp_bytes;  // per group (relative to the group's base)
  int ks, misc, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(const TcD& T, int n_stages) {
  TcSmem L;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) & ~127; return r; };
  L.xh = take(24 * TC_LBO); L.xl = take(24 * TC_LBO);
  L.h1h = take((T.h1_rows / 4) * TC_LBO); L.h1l = take((T.h1_rows / 4) * TC_LBO);
  L.h2h = take((T.h2_rows / 4) * TC_LBO); L.h2l = take((T.h2_rows / 4) * TC_LBO);
  L.side = take(TC_SIDE_ARRAYS * 4 * 32 * 16);
  L.bc = take(6 * TC_GN * 4 + TC_GN * 4);
  L.grp_bytes = o;
  o = TC_NG * L.grp_bytes;
  L.ks = take(n_stages * (TC_NG * 2 * 3 * 32) * 8 * 4);
  L.misc = take(128);
  L.total = o;
  return L;
}

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {  // K-major, no swizzle, LBO/SBO of the B layout
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {  // kind::tf32, FP32 accumulate, A and B K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t el = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(el));
  return el;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  return __uint_as_float(h);
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// split 8 values and store them as one row (k) of a B operand: columns 8h..8h+7 of the group
__device__ __forceinline__ void store_row_hilo(uint32_t hi_base, uint32_t lo_base, const float* v) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float hi = tf32_hi(v[r]);
    sts_f32(hi_base + 16 * r, hi);
    sts_f32(lo_base + 16 * r, v[r] - hi);
  }
}
// byte offset of (row k, column octet h) inside a B operand plane
__device__ __forceinline__ uint32_t tc_row_off(int k, int h) { return (uint32_t)(k >> 2) * TC_LBO + (uint32_t)(k & 3) * 4 + (uint32_t)h * TC_SBO; }

struct TcArgs {
  const float* wimg;  // [TC_WCOLS][128] TMEM image of the weights (hi/lo split), built by tc_image_kernel
};

// ---- weight image ---------------------------------------------------------------------------------------------
// wimg[c][l] = value of TMEM column c, lane l (see the map in the header comment). Flux stores W (out x in) column-major:
// element (o, k) of net/layer at theta[w_off + k*out + o].
__global__ void tc_image_kernel(const __grid_constant__ TcD T, const float* __restrict__ theta, float* __restrict__ wimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TC_WCOLS * 128) return;
  const int c = idx >> 7, l = idx & 127;
  float w = 0.f;
  bool lo = false;
  if (c < 192) {  // layer 1 block 0
    lo = c >= 96;
    const int k = lo ? c - 96 : c;
    const int f = l;
    if (f < 3 * T.h1) { const int q = f / T.h1, o = f - q * T.h1; w = theta[T.w_off[q][0] + k * T.h1 + o]; }
  } else if (l >= 96) {  // layer 1 block 1 (quadrant 3)
    lo = c >= 288;
    const int k = lo ? c - 288 : c - 192;
    const int f = 128 + (l - 96);
    if (f < 3 * T.h1) { const int q = f / T.h1, o = f - q * T.h1; w = theta[T.w_off[q][0] + k * T.h1 + o]; }
  } else {  // layer 2 / 3 stacks
    const int q = l >> 5, o = l & 31;
    const int K2 = 8 * T.k2_steps, K3 = 8 * T.k3_steps;
    if (c >= T.c_a2hi && c < T.c_a2hi + K2) {
      const int f = T.k2_start[q] + (c - T.c_a2hi) - T.h1 * q;
      if (o < T.h2 && f >= 0 && f < T.h1) w = theta[T.w_off[q][1] + f * T.h2 + o];
    } else if (c >= T.c_a2lo && c < T.c_a2lo + K2) {
      lo = true;
      const int f = T.k2_start[q] + (c - T.c_a2lo) - T.h1 * q;
      if (o < T.h2 && f >= 0 && f < T.h1) w = theta[T.w_off[q][1] + f * T.h2 + o];
    } else if (c >= T.c_a3hi && c < T.c_a3hi + K3) {
      const int f = T.k3_start[q] + (c - T.c_a3hi) - T.h2 * q;
      if (o < T.nout && f >= 0 && f < T.h2) w = theta[T.w_off[q][2] + f * T.nout + o];
    } else if (c >= T.c_a3lo && c < T.c_a3lo + K3) {
      lo = true;
      const int f = T.k3_start[q] + (c - T.c_a3lo) - T.h2 * q;
      if (o < T.nout && f >= 0 && f < T.h2) w = theta[T.w_off[q][2] + f * T.nout + o];
    }
  }
  const float hi = tf32_hi(w);
  wimg[idx] = lo ? (w - hi) : hi;
}

// ---- the solve kernel ---------------------------------------------------------------------------------------------
template <int ACT>  // hidden activation when both hidden layers share it, -1 = read T.act1/T.act2 at run time
__global__ void __launch_bounds__(TC_NT, 1) solve_tc_kernel(const __grid_constant__ ModelD M, const __grid_constant__ TcD T,
                                                            const __grid_constant__ TableauD tab, const TimeD tm,
                                                            const SolveArgs a, const TcArgs ta) {
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  const TcSmem L = tc_smem_layout(T, tab.n_stages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = warp >> 3, wg = warp & 7, qd = wg & 3, h = wg >> 2;
  const uint32_t sbase = smem_u32(smem_tc);
  const uint32_t gbase = sbase + g * L.grp_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_tc + L.misc) + g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_tc + L.misc + 64);
  const int N = 32;
  const int tile = blockIdx.x, col0 = tile * TC_CT;
  const int cg0 = TC_GN * g + 8 * h;  // first of this thread's 8 columns inside the tile

  // ---- prologue: zero shared memory, barriers, TMEM allocation, weights -> TMEM ----
  for (int i = tid; i < L.total / 16; i += TC_NT) reinterpret_cast<float4*>(smem_tc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;
  const uint32_t tlane = tb + ((uint32_t)(32 * qd) << 16);  // this warp's TMEM lane quadrant
  {
    // warps with the same quadrant split the 384 weight columns: 96 each
    const int part = warp >> 2;  // 0..3
    for (int c0 = 96 * part; c0 < 96 * part + 96; c0 += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(ta.wimg + (size_t)(c0 + i) * 128 + 32 * qd + lane);
      tmem_st8(tlane + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // per-thread biases
  float b1a = 0.f, b1b = 0.f, b2 = 0.f, b3 = 0.f;
  {
    const int f = 32 * qd + lane;
    if (f < 3 * T.h1) { const int q = f / T.h1; b1a = __ldg(a.theta + T.b_off[q][0] + (f - q * T.h1)); }
    const int f2 = 128 + lane;
    if (qd == 3 && f2 < 3 * T.h1) { const int q = f2 / T.h1; b1b = __ldg(a.theta + T.b_off[q][0] + (f2 - q * T.h1)); }
    if (qd < 3 && lane < T.h2) b2 = __ldg(a.theta + T.b_off[qd][1] + lane);
    if (qd < 3 && lane < T.nout) b3 = __ldg(a.theta + T.b_off[qd][2] + lane);
  }
  // boundary fluxes -> bc[q*2+tb][16] and diurnal amplitudes -> bc[6][16] of the group
  float* bcs_g = reinterpret_cast<float*>(smem_tc + g * L.grp_bytes + L.bc);
  if (wg == 0 && lane < TC_GN) {
    const int col = min(col0 + TC_GN * g + lane, a.ncol - 1);
    float raw[6], eff[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) raw[j] = __ldg(a.bcs + (size_t)col * 6 + j);
    bc_effective(M, raw, eff);
#pragma unroll
    for (int j = 0; j < 6; ++j) bcs_g[j * TC_GN + lane] = eff[j];
    bcs_g[6 * TC_GN + lane] = a.Q != nullptr ? __ldg(a.Q + col) : 0.f;
  }
  // state: x[r] = field qd, level lane, column cg0 + r
  float x[8], X[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int col = min(col0 + cg0 + r, a.ncol - 1);
    x[r] = qd < 3 ? __ldg(a.x0 + (size_t)col * 96 + 32 * qd + lane) : 0.f;
    X[r] = x[r];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // shared-memory addresses of this thread
  const uint32_t x_off = tc_row_off(32 * qd + lane, h);             // X row (qd < 3) / H1 row of block 0
  const uint32_t h1b_off = tc_row_off(128 + lane, h);               // H1 row of block 1 (qd == 3)
  const uint32_t h2_off = tc_row_off(T.h2 * qd + lane, h);          // H2 row (qd < 3, lane < h2)
  const uint32_t side = gbase + L.side;                             // [array][chunk 0..3][lane] float4
  auto side_addr = [&](int arr, int chunk) { return side + (uint32_t)(((arr * 4 + chunk) * 32 + lane) * 16); };
  const int tq = ((g * 2 + h) * 3 + qd) * 32 + lane;                // index among the stencil threads (qd < 3)
  const uint32_t ks_base = sbase + L.ks + (uint32_t)tq * 32;
  const uint32_t ks_stride = (uint32_t)(TC_NG * 2 * 3 * 32) * 32;
  const uint32_t dg = tb + 384 + 64 * g;                            // accumulator columns of this group
  const uint32_t id16 = tc_idesc(128, TC_GN);
  const bool issuer = (wg == 7);
  uint32_t parity = 0;
  const int bar_id = 1 + g;
  const bool mpp = (M.flags & F_MPP) || M.variant == RHS_INFER;
  const bool ca = (M.flags & F_CA) != 0;
  const float eps = M.variant == RHS_TRAIN ? M.rc.eps : 0.f;
  const float Nf = M.rc.Nf;
  const int act1 = ACT >= 0 ? ACT : T.act1, act2 = ACT >= 0 ? ACT : T.act2;

  auto write_X = [&]() {  // stage input -> B operand (hi/lo) + full-precision copy for quadrant 3
    if (qd < 3) {
      store_row_hilo(gbase + L.xh + x_off, gbase + L.xl + x_off, X);
      sts_v4(side_addr(5 + qd, 2 * h), X[0], X[1], X[2], X[3]);
      sts_v4(side_addr(5 + qd, 2 * h + 1), X[4], X[5], X[6], X[7]);
    }
  };

  // one RHS evaluation at stage input X (registers + shared memory); returns the tendencies in dx (qd < 3)
  auto rhs_eval = [&](float t_stage, float* dx) {
    // (1) operands visible to the async proxy, previous accumulator reads done
    fence_proxy_async();
    tc_fence_before();
    bar_sync_named(bar_id, 256);
    if (issuer) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bh = tc_desc(gbase + L.xh), bl = tc_desc(gbase + L.xl);
        const uint64_t kstep = (2 * TC_LBO) >> 4;
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          if (blk == 1 && T.n1b == 0) break;
          const uint32_t ahi = tb + 192 * blk, alo = ahi + 96, d = dg + 16 * blk;
#pragma unroll
          for (int s = 0; s < 12; ++s) tc_mma_ts(d, ahi + 8 * s, bl + s * kstep, id16, s > 0);  // Whi * Xlo
#pragma unroll
          for (int s = 0; s < 12; ++s) tc_mma_ts(d, alo + 8 * s, bh + s * kstep, id16, 1);      // Wlo * Xhi
#pragma unroll
          for (int s = 0; s < 12; ++s) tc_mma_ts(d, ahi + 8 * s, bh + s * kstep, id16, 1);      // Whi * Xhi
        }
        tc_commit(mbar);
      }
      __syncwarp();
    }
    if (qd == 3) {
      // diffusivities at face lane+1 and Coriolis terms at level lane, for this thread's 8 columns
      float u[8], v[8], Tt[8];
      {
        float4 p;
        p = lds_v4(side_addr(5, 2 * h)); u[0] = p.x; u[1] = p.y; u[2] = p.z; u[3] = p.w;
        p = lds_v4(side_addr(5, 2 * h + 1)); u[4] = p.x; u[5] = p.y; u[6] = p.z; u[7] = p.w;
        p = lds_v4(side_addr(6, 2 * h)); v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w;
        p = lds_v4(side_addr(6, 2 * h + 1)); v[4] = p.x; v[5] = p.y; v[6] = p.z; v[7] = p.w;
        p = lds_v4(side_addr(7, 2 * h)); Tt[0] = p.x; Tt[1] = p.y; Tt[2] = p.z; Tt[3] = p.w;
        p = lds_v4(side_addr(7, 2 * h + 1)); Tt[4] = p.x; Tt[5] = p.y; Tt[6] = p.z; Tt[7] = p.w;
      }
      float Du[8], Dv[8], DT[8], cu[8], cv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float Gu = Nf * (__shfl_down_sync(0xffffffffu, u[r], 1) - u[r]);
        const float Gv = Nf * (__shfl_down_sync(0xffffffffu, v[r], 1) - v[r]);
        const float GT = Nf * (__shfl_down_sync(0xffffffffu, Tt[r], 1) - Tt[r]);
        float nu = 0.f, nuT = 0.f;
        if (mpp) {
          const float su = M.rc.sig_u * (Gu + eps), sv = M.rc.sig_v * (Gv + eps);
          const float Ri = __fdividef(M.rc.BzC * (GT + eps), su * su + sv * sv);
          nu = nu_of_ri(M, Ri);
          nuT = nu * M.rc.inv_Pr;
          if (M.variant == RHS_INFER && ca) {
            const float test = (M.flags & F_CA_LITERAL_U) ? Gu : GT;
            nuT = test > 0.f ? nuT : M.rc.kappa;
          }
        } else if (ca) {
          nuT = GT < 0.f ? M.rc.kappa : 0.f;  // c2*kappa*min(0, G_T) (NDE_training.jl:140-143 with the inference kappa)
        }
        Du[r] = M.rc.c[0] * nu; Dv[r] = M.rc.c[1] * nu; DT[r] = M.rc.c[2] * nuT;
        cu[r] = M.rc.cor_u_s * v[r] + M.rc.cor_u_m;
        cv[r] = -(M.rc.cor_v_s * u[r] + M.rc.cor_v_m);
      }
      sts_v4(side_addr(0, 2 * h), Du[0], Du[1], Du[2], Du[3]); sts_v4(side_addr(0, 2 * h + 1), Du[4], Du[5], Du[6], Du[7]);
      sts_v4(side_addr(1, 2 * h), Dv[0], Dv[1], Dv[2], Dv[3]); sts_v4(side_addr(1, 2 * h + 1), Dv[4], Dv[5], Dv[6], Dv[7]);
      sts_v4(side_addr(2, 2 * h), DT[0], DT[1], DT[2], DT[3]); sts_v4(side_addr(2, 2 * h + 1), DT[4], DT[5], DT[6], DT[7]);
      sts_v4(side_addr(3, 2 * h), cu[0], cu[1], cu[2], cu[3]); sts_v4(side_addr(3, 2 * h + 1), cu[4], cu[5], cu[6], cu[7]);
      sts_v4(side_addr(4, 2 * h), cv[0], cv[1], cv[2], cv[3]); sts_v4(side_addr(4, 2 * h + 1), cv[4], cv[5], cv[6], cv[7]);
      if ((M.flags & F_DIURNAL) && h == 0 && lane < TC_GN) bcs_g[5 * TC_GN + lane] = diurnal_top_eff(M, bcs_g[6 * TC_GN + lane], t_stage);
    }
    // ---- layer 1 epilogue ----
    mbar_wait(mbar, parity); parity ^= 1u;
    tc_fence_after();
    {
      float v[8];
      tmem_ld8(dg + ((uint32_t)(32 * qd) << 16) + 8 * h, v);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = act_fwd(act1, v[r] + b1a);
      store_row_hilo(gbase + L.h1h + x_off, gbase + L.h1l + x_off, v);
      if (qd == 3 && T.n1b > 0) {
        tmem_ld8(dg + ((uint32_t)96 << 16) + 16 + 8 * h, v);
        if (lane < T.n1b) {
#pragma unroll
          for (int r = 0; r < 8; ++r) v[r] = act_fwd(act1, v[r] + b1b);
          store_row_hilo(gbase + L.h1h + h1b_off, gbase + L.h1l + h1b_off, v);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    bar_sync_named(bar_id, 256);
    // ---- layer 2 ----
    if (issuer) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kstep = (2 * TC_LBO) >> 4;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const uint64_t bh = tc_desc(gbase + L.h1h + (T.k2_start[q] >> 2) * TC_LBO);
          const uint64_t bl = tc_desc(gbase + L.h1l + (T.k2_start[q] >> 2) * TC_LBO);
          const uint32_t d = dg + 16 * q;
          const uint32_t ahi = tb + T.c_a2hi, alo = tb + T.c_a2lo;
#pragma unroll 1
          for (int s = 0; s < T.k2_steps; ++s) tc_mma_ts(d, ahi + 8 * s, bl + s * kstep, id16, s > 0);
#pragma unroll 1
          for (int s = 0; s < T.k2_steps; ++s) tc_mma_ts(d, alo + 8 * s, bh + s * kstep, id16, 1);
#pragma unroll 1
          for (int s = 0; s < T.k2_steps; ++s) tc_mma_ts(d, ahi + 8 * s, bh + s * kstep, id16, 1);
        }
        tc_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, parity); parity ^= 1u;
    tc_fence_after();
    if (qd < 3) {
      float v[8];
      tmem_ld8(dg + ((uint32_t)(32 * qd) << 16) + 16 * qd + 8 * h, v);
      if (lane < T.h2) {
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = act_fwd(act2, v[r] + b2);
        store_row_hilo(gbase + L.h2h + h2_off, gbase + L.h2l + h2_off, v);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    bar_sync_named(bar_id, 256);
    // ---- layer 3 ----
    if (issuer) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kstep = (2 * TC_LBO) >> 4;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const uint64_t bh = tc_desc(gbase + L.h2h + (T.k3_start[q] >> 2) * TC_LBO);
          const uint64_t bl = tc_desc(gbase + L.h2l + (T.k3_start[q] >> 2) * TC_LBO);
          const uint32_t d = dg + 16 * q;
          const uint32_t ahi = tb + T.c_a3hi, alo = tb + T.c_a3lo;
#pragma unroll 1
          for (int s = 0; s < T.k3_steps; ++s) tc_mma_ts(d, ahi + 8 * s, bl + s * kstep, id16, s > 0);
#pragma unroll 1
          for (int s = 0; s < T.k3_steps; ++s) tc_mma_ts(d, alo + 8 * s, bh + s * kstep, id16, 1);
#pragma unroll 1
          for (int s = 0; s < T.k3_steps; ++s) tc_mma_ts(d, ahi + 8 * s, bh + s * kstep, id16, 1);
        }
        tc_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, parity); parity ^= 1u;
    tc_fence_after();
    // ---- layer 3 epilogue + stencil (Appendix A of SURVEY.md) ----
    if (qd < 3) {
      float nn[8];
      tmem_ld8(dg + ((uint32_t)(32 * qd) << 16) + 16 * qd + 8 * h, nn);
      float D[8], cor[8], bt[8], bb[8];
      {
        float4 p;
        p = lds_v4(side_addr(qd, 2 * h)); D[0] = p.x; D[1] = p.y; D[2] = p.z; D[3] = p.w;
        p = lds_v4(side_addr(qd, 2 * h + 1)); D[4] = p.x; D[5] = p.y; D[6] = p.z; D[7] = p.w;
        if (qd < 2) {
          p = lds_v4(side_addr(3 + qd, 2 * h)); cor[0] = p.x; cor[1] = p.y; cor[2] = p.z; cor[3] = p.w;
          p = lds_v4(side_addr(3 + qd, 2 * h + 1)); cor[4] = p.x; cor[5] = p.y; cor[6] = p.z; cor[7] = p.w;
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) cor[r] = 0.f;
        }
        const uint32_t bca = gbase + L.bc + (uint32_t)((2 * qd) * TC_GN + 8 * h) * 4;
        p = lds_v4(bca); bb[0] = p.x; bb[1] = p.y; bb[2] = p.z; bb[3] = p.w;
        p = lds_v4(bca + 16); bb[4] = p.x; bb[5] = p.y; bb[6] = p.z; bb[7] = p.w;
        p = lds_v4(bca + TC_GN * 4); bt[0] = p.x; bt[1] = p.y; bt[2] = p.z; bt[3] = p.w;
        p = lds_v4(bca + TC_GN * 4 + 16); bt[4] = p.x; bt[5] = p.y; bt[6] = p.z; bt[7] = p.w;
      }
      const float Aq = M.rc.A[qd] * Nf;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float G = Nf * (__shfl_down_sync(0xffffffffu, X[r], 1) - X[r]);
        float Fup = (nn[r] + b3) - D[r] * G;       // face lane+1
        if (lane == N - 1) Fup = bt[r];
        float Fdn = __shfl_up_sync(0xffffffffu, Fup, 1);
        if (lane == 0) Fdn = bb[r];
        dx[r] = cor[r] - Aq * (Fup - Fdn);
      }
    }
  };

  if (a.rhs_only) {
    write_X();
    float dx[8];
    rhs_eval(a.t_rhs, dx);
    if (qd < 3) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int col = col0 + cg0 + r;
        if (col < a.ncol) a.dxdt[(size_t)col * 96 + 32 * qd + lane] = dx[r];
      }
    }
  } else {
    const float hstep = tm.dt / (float)tm.n_substeps;
    const int ns = tab.n_stages;
    int frame = 0, ci = 0;
    const size_t traj_stride = (size_t)a.n_saved * 96;
    auto save_frame = [&](int fr) {
      if (qd < 3) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int col = col0 + cg0 + r;
          if (col < a.ncol) a.traj[(size_t)col * traj_stride + (size_t)fr * 96 + 32 * qd + lane] = x[r];
        }
      }
    };
    auto save_ckpt = [&](int c) {  // tile-native layout of the adjoint kernel: [tile][n_ckpt][S][32]
      if (qd < 3) {
        float4* dst = reinterpret_cast<float4*>(a.ckpt + (((size_t)tile * a.n_ckpt + c) * 96 + 32 * qd + lane) * TC_CT + cg0);
        dst[0] = make_float4(x[0], x[1], x[2], x[3]);
        dst[1] = make_float4(x[4], x[5], x[6], x[7]);
      }
    };
    if (a.traj != nullptr && tm.save_stride > 0) { save_frame(0); frame = 1; }
    if (a.ckpt != nullptr) { save_ckpt(0); ci = 1; }
    write_X();
    for (int n = 0; n < tm.n_steps; ++n) {
      for (int sub = 0; sub < tm.n_substeps; ++sub) {
        const float tbase = tm.t0 + (float)n * tm.dt + (float)sub * hstep;
#pragma unroll 1
        for (int i = 0; i < ns; ++i) {
          float dx[8];
          rhs_eval(tbase + tab.c[i] * hstep, dx);
          if (qd < 3) {
            const bool last = (i + 1 == ns);
            float acc[8];
            const float ci_ = last ? tab.b[i] : tab.a[(i + 1) % CPZ_MAX_STAGES][i];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = ci_ * dx[r];
#pragma unroll 1
            for (int j = 0; j < i; ++j) {
              const float cj = last ? tab.b[j] : tab.a[(i + 1) % CPZ_MAX_STAGES][j];
              const float4 p0 = lds_v4(ks_base + j * ks_stride), p1 = lds_v4(ks_base + j * ks_stride + 16);
              acc[0] = fmaf(cj, p0.x, acc[0]); acc[1] = fmaf(cj, p0.y, acc[1]); acc[2] = fmaf(cj, p0.z, acc[2]); acc[3] = fmaf(cj, p0.w, acc[3]);
              acc[4] = fmaf(cj, p1.x, acc[4]); acc[5] = fmaf(cj, p1.y, acc[5]); acc[6] = fmaf(cj, p1.z, acc[6]); acc[7] = fmaf(cj, p1.w, acc[7]);
            }
            if (!last) {
              sts_v4(ks_base + i * ks_stride, dx[0], dx[1], dx[2], dx[3]);
              sts_v4(ks_base + i * ks_stride + 16, dx[4], dx[5], dx[6], dx[7]);
#pragma unroll
              for (int r = 0; r < 8; ++r) X[r] = fmaf(hstep, acc[r], x[r]);
            } else {
#pragma unroll
              for (int r = 0; r < 8; ++r) { x[r] = fmaf(hstep, acc[r], x[r]); X[r] = x[r]; }
            }
            write_X();
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
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

}  // namespace cpz

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
#include <type_traits>

#include "cpz_solve.cuh"

namespace cpz {

constexpr int TC_NG = 2;     // column groups per CTA
constexpr int TC_GN = 16;    // columns per group = MMA N
constexpr int TC_CT = TC_NG * TC_GN;
constexpr int TC_NT = 512;   // 2 groups x 8 warps
constexpr uint32_t TC_SBO = 128;                      // bytes between column octets of a B operand
constexpr uint32_t TC_LBO = (TC_GN / 8) * 128 + 16;  // 272: bytes between 4-row K chunks (16 B pad: conflict-free row stores)
constexpr int TC_K2S = 7;            // K steps of a layer-2 chain (covers h1 <= 53 plus window alignment)
constexpr int TC_WCOLS = 384;        // TMEM columns holding weights
constexpr int TC_SIDE_ARRAYS = 6;    // Du, Dv, DT (x Nz, face lane+1), Xu, Xv, XT (full-precision stage input)

// constants of the face-diffusivity computation, folded on the host
struct SideC {
  float e;          // eps / Nz (train variant) or 0
  float su2, sv2;   // sigma_u^2, sigma_v^2
  float k1, k2;     // exp2 argument: k1 * (dT+e)/den - k2, with the factors 2/dRi, BzC/Nz and log2(e) folded in
  float a0, a1;     // Nz*c_u*nu0, Nz*c_u*nu_m
  float b0, b1;     // v
  float t0, t1;     // T (includes 1/Pr)
  float kap;        // Nz*c_T*kappa
};

struct TcD {
  int h1, h2, nout;      // per-net layer widths (identical for the three nets)
  int act1, act2, act3;
  int n1b;               // layer-1 rows in block 1 (= 3*h1 - 128, <= 32)
  int k2_start[3], k3_start[3];  // first operand row (multiple of 4) of net q's K window in H1 / H2
  int k3_steps;                  // K steps (of 8 rows) of a layer-3 chain (3 or 4); layer 2 always runs TC_K2S
  int h1_rows, h2_rows;       // allocated operand rows (rows past the real ones stay zero)
  int c_a2hi, c_a2lo, c_a3hi, c_a3lo;  // TMEM columns of the layer-2/3 stacks
  int w_off[3][3], b_off[3][3];        // theta offsets [net][layer]
  int side_mode;
  SideC sc;
};

// B operand planes (activations) in shared memory: canonical K-major layout without swizzle (verified by the address
// probe in tools/micro/umma_test.cu):  byte(n, k) = (n/8)*SBO + (n%8)*16 + (k/4)*LBO + (k%4)*4.
// A thread owns one row k and 8 columns: 8 scalar stores 16 B apart; the 32 rows of a warp hit 32 distinct banks.
// (An MN-major operand would allow 16-byte stores, but kind::tf32 with an un-swizzled MN-major B returned zeros in the
// probe, so it is not used.)
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
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {  // kind::tf32, FP32 accumulate, A (TMEM) and B K-major
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
// round to nearest tf32 (ties away from zero, as cvt.rna.tf32.f32, without its Inf/NaN guard: 2 ALU instructions, not 4)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// fast activations on raw MUFU.EX2 / MUFU.RCP (ex2.approx.ftz, rcp.approx.ftz: no range fix-ups; arguments are clamped)
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ACT>
__device__ __forceinline__ float tc_act(int act_rt, float x) {
  constexpr float L2E = 1.4426950408889634f;
  if constexpr (ACT == ACT_MISH) {  // x*tanh(softplus(x)) = x*n/(n+2), n = e^x(e^x+2)      (NNlib 0.7.20 mish)
    const float e = ex2_fast(fminf(x, 20.f) * L2E);
    const float n = e * (e + 2.f);
    return x * n * rcp_fast(n + 2.f);
  } else if constexpr (ACT == ACT_RELU) {
    return x < 0.f ? 0.f : x;  // NaN propagates (fmaxf would swallow it)
  } else {
    return act_fwd(act_rt, x);
  }
}

// split 8 values (columns 8h..8h+7 of one operand row) into tf32 hi + residual lo and store them (16 B apart)
template <int CNT = 8>
__device__ __forceinline__ void store_row_hilo(uint32_t hi_addr, uint32_t lo_addr, const float* v) {
#pragma unroll
  for (int r = 0; r < CNT; ++r) {
    const float hi = tf32_hi(v[r]);
    sts_f32(hi_addr + 16 * r, hi);
    sts_f32(lo_addr + 16 * r, v[r] - hi);
  }
}
// byte offset of (row k, column octet h) inside a B operand plane
__device__ __forceinline__ uint32_t tc_row_off(int k, int h) { return (uint32_t)(k >> 2) * TC_LBO + (uint32_t)(k & 3) * 4 + (uint32_t)h * TC_SBO; }
__device__ __forceinline__ void lds8(uint32_t a0, uint32_t a1, float* v) {
  const float4 p = lds_v4(a0), q = lds_v4(a1);
  v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
}

// CLOSURE mode of solve_tc_kernel: per-step closure of the u/v/T NDE embedded in a host ocean model (cpz_closure_uvt.cuh;
// wind_mixing/src/NDE_oceananigans.jl:380-405) — dimensional (Nx,Ny,Nz) fields in, d/dz of the NN fluxes and the implicitly
// diffused state out, one MLP evaluation and one implicit step per column
struct TcClosure {
  const float* f[3];   // u, v, T: [Nz][ncol]
  float* dzf;          // [3][Nz][ncol]
  float* out;          // [3][Nz][ncol]
  int ncol, n_tiles;
  float inv_dz;
  float hsub;          // non-dimensional sub-step with hsub A_q N (N c_q nu) = dt nu / dz^2
  float top[3];        // uw, vw, wT at the surface face
  float mu[6], sig[6], inv_sig[3];
};

struct TcArgs {
  const float* wimg;  // [TC_WCOLS][128] TMEM image of the weights (hi/lo split), built by tc_image_kernel
  int stagger_ns;     // start delay of the second column group
  TcClosure cl;       // CLOSURE instantiation only
};

// ---- weight image ---------------------------------------------------------------------------------------------
// wimg[c][l] = value of TMEM column c, lane l (see the map in the header comment). Flux stores W (out x in) column-major:
// element (o, k) of net/layer at theta[w_off + k*out + o].
static __global__ void tc_image_kernel(const __grid_constant__ TcD T, const float* __restrict__ theta, float* __restrict__ wimg) {
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
  } else {  // layer 2 / 3 stacks: lane 32q+o = output o of net q, column = input index inside the net's K window
    const int q = l >> 5, o = l & 31;
    const int K2 = 8 * TC_K2S, K3 = 8 * T.k3_steps;
    int f = -1, layer = 0;
    if (c >= T.c_a2hi && c < T.c_a2hi + K2) { f = c - T.c_a2hi; layer = 1; }
    else if (c >= T.c_a2lo && c < T.c_a2lo + K2) { f = c - T.c_a2lo; layer = 1; lo = true; }
    else if (c >= T.c_a3hi && c < T.c_a3hi + K3) { f = c - T.c_a3hi; layer = 2; }
    else if (c >= T.c_a3lo && c < T.c_a3lo + K3) { f = c - T.c_a3lo; layer = 2; lo = true; }
    if (layer == 1) f += T.k2_start[q] - T.h1 * q;  // operand row -> input index of net q
    if (layer == 2) f += T.k3_start[q] - T.h2 * q;
    if (layer == 1 && o < T.h2 && f >= 0 && f < T.h1) w = theta[T.w_off[q][1] + f * T.h2 + o];
    if (layer == 2 && o < T.nout && f >= 0 && f < T.h2) w = theta[T.w_off[q][2] + f * T.nout + o];
  }
  const float hi = tf32_hi(w);
  wimg[idx] = lo ? (w - hi) : hi;
}

// diffusivity modes of the face computation (hoisted out of the column loop)
enum { SIDE_NONE = 0, SIDE_MPP = 1, SIDE_MPP_CA_T = 2, SIDE_MPP_CA_U = 3, SIDE_CA_ONLY = 4 };

// Nz * c_q * nu_q at face lane+1 from the level differences du, dv, dT (lane+1 minus lane) of one column.
// Ri = H g alpha sigma_T dT/dz / ((sigma_u du/dz)^2 + (sigma_v dv/dz)^2)      (NDE_training.jl:46-52,115-119)
// nu = nu0 + nu_m (1 - tanh((Ri - Ric)/dRi))/2 = nu0 + nu_m / (1 + exp(2 (Ri - Ric)/dRi))   (NDE_training.jl:54,125)
template <int MODE>
__device__ __forceinline__ void side_column(const SideC& C, float du, float dv, float dT, float& Du, float& Dv, float& DT) {
  if constexpr (MODE == SIDE_NONE) {
    Du = 0.f; Dv = 0.f; DT = 0.f;
  } else if constexpr (MODE == SIDE_CA_ONLY) {  // c_T*kappa*min(0, dT/dz): NDE_training.jl:140-143
    Du = 0.f; Dv = 0.f; DT = dT < 0.f ? C.kap : 0.f;
  } else {
    const float a = du + C.e, b = dv + C.e, c = dT + C.e;
    const float den = fmaf(C.sv2 * b, b, C.su2 * a * a);
    const float y = fmaf(C.k1 * c, rcp_fast(den), -C.k2);
    const float w = rcp_fast(1.f + ex2_fast(fminf(y, 126.f)));
    Du = fmaf(C.a1, w, C.a0); Dv = fmaf(C.b1, w, C.b0); DT = fmaf(C.t1, w, C.t0);
    if constexpr (MODE == SIDE_MPP_CA_T) DT = dT > 0.f ? DT : C.kap;   // training_postprocessing.jl:118-124
    if constexpr (MODE == SIDE_MPP_CA_U) DT = du > 0.f ? DT : C.kap;
  }
}

// Tridiagonal solve across the 32 lanes of a warp (lane = level): a_k x_{k-1} + b_k x_k + c_k x_{k+1} = d_k with a = 0 on lane 0
// and c = 0 on lane 31, by parallel cyclic reduction — five shuffle steps, no shared memory, no serial sweep. After the step
// with stride s every equation couples level k to k-2s and k+2s only; a coefficient that would point outside the column is
// zero by induction, so out-of-range shuffles (which return the lane's own value) are multiplied by zero.
// a / b from MUFU.RCP plus one Newton step on the reciprocal and one residual correction of the quotient: within an ulp of the
// IEEE quotient for the well-scaled, diagonally dominant pivots here (|b| >= 1), at a quarter of the instructions of the
// IEEE division sequence (which dominated the implicit step)
__device__ __forceinline__ float div_nr(float a, float b) {
  float r = rcp_fast(b);
  r = fmaf(r, fmaf(-b, r, 1.f), r);
  const float q = a * r;
  return fmaf(r, fmaf(-b, q, a), q);
}
__device__ __forceinline__ float pcr32(float a, float b, float c, float d) {
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const float am = __shfl_up_sync(0xffffffffu, a, s), bm = __shfl_up_sync(0xffffffffu, b, s);
    const float cm = __shfl_up_sync(0xffffffffu, c, s), dm = __shfl_up_sync(0xffffffffu, d, s);
    const float ap = __shfl_down_sync(0xffffffffu, a, s), bp = __shfl_down_sync(0xffffffffu, b, s);
    const float cp = __shfl_down_sync(0xffffffffu, c, s), dp = __shfl_down_sync(0xffffffffu, d, s);
    const float al = -div_nr(a, bm), ga = -div_nr(c, bp);  // a bare MUFU.RCP is not enough: the kappa = 10 convective-adjustment rows amplify its error
    b = fmaf(al, cm, fmaf(ga, ap, b));
    d = fmaf(al, dm, fmaf(ga, dp, d));
    a = al * am;
    c = ga * cp;
  }
  return div_nr(d, b);  // at kappa = 10 the off-diagonals are ~100 and a bare MUFU.RCP's error shows up in the profiles
}

// ---- the solve kernel ---------------------------------------------------------------------------------------------
// out of line: the full-range sinf keeps its slow path (and local-memory table) away from the hot loop
static __device__ __noinline__ float tc_diurnal_top(const ModelD& M, float Q, float t) { return diurnal_top_eff(M, Q, t); }

// ACT: shared hidden activation (-1: T.act1/T.act2 at run time); K3S: layer-3 K steps; RHS_ONLY: single evaluation (cpz_rhs)
// CPT: real columns per thread (8, or 7 so that 4 096 columns make 147 tiles of 28 and use every SM; the 8th slot is padding)
// AUX: the segment pass of the tensor-core adjoint — every stage evaluation also stores its input X_i and the pre-activations
//      z1, z2 as operand-image records (AuxD, cpz_solve.cuh); the start state may come from a checkpoint (a.x0_tile)
// IMPL: CPZ_FLAG_IMPLICIT_DIFFUSION solves — the backward-Euler step is compiled only into this instantiation (inlined into the
//       explicit kernel it cost 6 registers + a spill and 8 % of the config-2 time)
// CLOSURE: persistent over 32-column tiles of three (Nx,Ny,Nz) fields (TcArgs::cl): one MLP evaluation + one implicit step per tile
template <int ACT, int K3S, bool PROF = false, bool RHS_ONLY = false, int CPT = 8, bool AUX = false, bool IMPL = false, bool CLOSURE = false>
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
  // split mode (segment pass of the adjoint with few tiles): CTA b runs only column group b & 1 of tile b >> 1, so that a batch
  // of n tiles spreads over 2 n SMs and a group has the SM's pipes to itself; the other group's warps skip the time loop
  const bool split = AUX && a.split != 0;
  const bool active = !split || g == (int)(blockIdx.x & 1);
  const int tile = split ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, col0 = tile * (4 * CPT);
  const int cg0 = 2 * CPT * g + CPT * h;  // first of this thread's CPT columns inside the tile

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
  // state x[r] = field qd, level lane, column cg0 + r; bnd[r] = effective boundary flux of field qd: bottom (lane 0) /
  // top (lane 31) of column r; Qd[r] = diurnal amplitude (only read by lane 31 of the T warps)
  float x[8], X[8], bnd[8];
  const bool diurnal = (M.flags & F_DIURNAL) != 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) { x[r] = 0.f; X[r] = 0.f; bnd[r] = 0.f; }
#pragma unroll
  for (int r = 0; r < (CLOSURE ? 0 : CPT); ++r) {
    const int col = min(col0 + cg0 + r, a.ncol - 1);
    if (AUX && a.x0_tile != nullptr) x[r] = qd < 3 ? __ldcg(a.x0_tile + (size_t)tile * a.x0_tile_stride + (size_t)(32 * qd + lane) * TC_CT + cg0 + r) : 0.f;
    else x[r] = qd < 3 ? __ldg(a.x0 + (size_t)col * (a.x0_stride ? a.x0_stride : (size_t)96) + 32 * qd + lane) : 0.f;
    X[r] = x[r];
    float raw[6], eff[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) raw[j] = __ldg(a.bcs + (size_t)col * 6 + j);
    bc_effective(M, raw, eff);
    bnd[r] = 0.f;
    if (qd < 3) bnd[r] = lane == 0 ? eff[2 * qd] : eff[2 * qd + 1];
    if (diurnal && qd == 2 && lane == 31) reinterpret_cast<float*>(smem_tc + g * L.grp_bytes + L.bc)[8 * h + r] = a.Q != nullptr ? __ldg(a.Q + col) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // shared-memory addresses of this thread (operand planes: [column block][row][4])
  const uint32_t xb = tc_row_off(32 * qd + lane, h);            // X row (qd < 3) and H1 row of block 0
  const uint32_t h1a = xb;
  const uint32_t h1b = tc_row_off(128 + lane, h);               // H1 row of block 1 (qd == 3)
  const uint32_t h2a = tc_row_off(T.h2 * qd + lane, h);         // H2 row (qd < 3, lane < h2)
  const uint32_t side = gbase + L.side;                                                         // [array][column block][lane] float4
  auto side_addr = [&](int arr, int blk) { return side + (uint32_t)(((arr * 4 + blk) * 32 + lane) * 16); };
  const int tq = ((g * 2 + h) * 3 + qd) * 32 + lane;                // index among the stencil threads (qd < 3)
  // stage slots: [stage][column half][thread] float4 — a 16-byte thread stride keeps the 128-bit accesses free of bank conflicts
  // (config 2: 43.4 -> 42.8 ms). The implicit-diffusion forward kernel keeps [stage][thread][2 x float4]: measured 28.5 ms
  // against 30.0 ms with the split layout (its two column groups fall into a worse relative phase)
  constexpr bool KS_SPLIT = !(IMPL && !AUX);
  const uint32_t ks_base = sbase + L.ks + (uint32_t)tq * (KS_SPLIT ? 16 : 32);
  const uint32_t ks_stride = (uint32_t)(TC_NG * 2 * 3 * 32) * 32, ks_half = KS_SPLIT ? ks_stride / 2 : 16;
  const uint32_t dg = tb + 384 + 64 * g;                            // accumulator columns of this group
  const uint32_t id16 = tc_idesc(128, TC_GN);
  uint32_t parity = 0;
  const int bar_id = 1 + g;
  const int act1 = T.act1, act2 = T.act2;
  // per-thread stencil constants: dx = cs * X_other + cm - Aq * (F_up - F_down)
  const float Aq = qd < 3 ? M.rc.A[qd] * M.rc.Nf : 0.f;
  const float cs = qd == 0 ? M.rc.cor_u_s : (qd == 1 ? -M.rc.cor_v_s : 0.f);
  const float cm = qd == 0 ? M.rc.cor_u_m : (qd == 1 ? -M.rc.cor_v_m : 0.f);
  // CPZ_FLAG_IMPLICIT_DIFFUSION: the Runge–Kutta stages see no diffusive flux (mode NONE); the diffusivities act in the
  // backward-Euler solve at the start of every sub-step instead
  const bool implicit = (M.flags & F_IMPLICIT) != 0;
  const int side_mode = (implicit || CLOSURE) ? (int)SIDE_NONE : T.side_mode;

  auto write_X = [&]() {  // stage input -> B operand (hi/lo) + full-precision copy
    if (qd < 3) {
      store_row_hilo<CPT>(gbase + L.xh + xb, gbase + L.xl + xb, X);
      sts_v4(side_addr(3 + qd, 2 * h), X[0], X[1], X[2], X[3]);
      sts_v4(side_addr(3 + qd, 2 * h + 1), X[4], X[5], X[6], X[7]);
    }
  };

  // PROF: cycle counters of CTA 0 — slots 0..7 warp 0 (a stencil warp): bar1, wait1, E1+bar2, wait2, E2+bar3, wait3, E3+stencil, RK+write;
  // slots 8..11 warp 7 (quadrant 3 of group 0): issue L1, face diffusivities, wait L1, E1 (two blocks)
  long long pt = 0;
  int trace_n = 0;
  auto tick = [&](int slot) {
    if constexpr (PROF) {
      if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7)) {
        const long long now = clock64();
        if (slot >= 0 && ((warp == 0) == (slot < 8 || slot >= 12))) atomicAdd(a.prof + slot, (unsigned long long)(now - pt));
        pt = now;
      }
    }
  };

  // MMA chains: D (+)= Whi*Xlo + Wlo*Xhi + Whi*Xhi over `steps` K steps of 8 rows; a?? = TMEM columns, b? = descriptors
  auto chain = [&](uint32_t d, uint32_t ahi, uint32_t alo, uint64_t bh, uint64_t bl, auto STEPS) {
    constexpr int NS = decltype(STEPS)::value;
    constexpr uint64_t kstep = (2 * TC_LBO) >> 4;  // one K step = two 4-row chunks
#pragma unroll
    for (int s = 0; s < NS; ++s) tc_mma_ts(d, ahi + 8 * s, bl + s * kstep, id16, s > 0);
#pragma unroll
    for (int s = 0; s < NS; ++s) tc_mma_ts(d, alo + 8 * s, bh + s * kstep, id16, 1);
#pragma unroll
    for (int s = 0; s < NS; ++s) tc_mma_ts(d, ahi + 8 * s, bh + s * kstep, id16, 1);
  };

  // AUX records: this thread's two column quads (columns cg0..cg0+7) of row `row` of a record with `rows` rows
  int ev = AUX ? -a.aux.ev_skip : 0;  // record index of the current stage evaluation (negative: not stored)
  auto aux_store = [&](float* base, int rows, int row, const float* v) {
    if (ev < 0) return;
    float4* dst = reinterpret_cast<float4*>(base + ((size_t)tile * a.aux.n_eval + ev) * (size_t)(32 * rows)) + (size_t)(cg0 >> 2) * rows + row;
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[rows] = make_float4(v[4], v[5], v[6], v[7]);
  };

  // one RHS evaluation at stage input X (registers + shared memory); returns the tendencies in dx (qd < 3)
  auto rhs_eval = [&](float t_stage, float* dx) {
    // (1) operands visible to the async proxy, previous accumulator reads done
    tick(-1);
    if constexpr (AUX) { if (qd < 3) aux_store(a.aux.x, a.aux.rx, 32 * qd + lane, X); }
    fence_proxy_async();
    tc_fence_before();
    bar_sync_named(bar_id, 256);
    tick(0);
    if constexpr (PROF) {  // trace: clock at the start of the first 96 RHS evaluations of both groups (CTA 0)
      if (blockIdx.x == 0 && wg == 0 && lane == 0 && trace_n < 96) { a.prof[32 + 2 * trace_n + g] = (unsigned long long)clock64(); ++trace_n; }
    }
    if (wg == 7) {  // layer 1: rows 0..127 then the quadrant-3 block
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bh = tc_desc(gbase + L.xh), bl = tc_desc(gbase + L.xl);
        chain(dg, tb, tb + 96, bh, bl, std::integral_constant<int, 12>{});
        if (T.n1b > 0) chain(dg + 16, tb + 192, tb + 288, bh, bl, std::integral_constant<int, 12>{});
        tc_commit(mbar);
      }
      __syncwarp();
      tick(8);
    }
    {
      // face diffusivities (x Nz) at face lane+1, two columns per thread (columns 8h + 2qd, +1), while layer 1 runs
      const uint32_t off = (uint32_t)(8 * (qd & 1));
      const int blk = 2 * h + (qd >> 1);
      float2 u, v, Tt, Du, Dv, DT;
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(u.x), "=f"(u.y) : "r"(side_addr(3, blk) + off) : "memory");
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(side_addr(4, blk) + off) : "memory");
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(Tt.x), "=f"(Tt.y) : "r"(side_addr(5, blk) + off) : "memory");
      u.x = __shfl_down_sync(0xffffffffu, u.x, 1) - u.x; u.y = __shfl_down_sync(0xffffffffu, u.y, 1) - u.y;
      v.x = __shfl_down_sync(0xffffffffu, v.x, 1) - v.x; v.y = __shfl_down_sync(0xffffffffu, v.y, 1) - v.y;
      Tt.x = __shfl_down_sync(0xffffffffu, Tt.x, 1) - Tt.x; Tt.y = __shfl_down_sync(0xffffffffu, Tt.y, 1) - Tt.y;
      switch (side_mode) {
#define CPZ_SIDE_CASE(MODE)                                             \
  case MODE:                                                            \
    side_column<MODE>(T.sc, u.x, v.x, Tt.x, Du.x, Dv.x, DT.x);           \
    side_column<MODE>(T.sc, u.y, v.y, Tt.y, Du.y, Dv.y, DT.y);           \
    break;
        CPZ_SIDE_CASE(SIDE_MPP)
        CPZ_SIDE_CASE(SIDE_MPP_CA_T)
        CPZ_SIDE_CASE(SIDE_MPP_CA_U)
        CPZ_SIDE_CASE(SIDE_CA_ONLY)
        default:
          CPZ_SIDE_CASE(SIDE_NONE)
#undef CPZ_SIDE_CASE
      }
      asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(side_addr(0, blk) + off), "f"(Du.x), "f"(Du.y) : "memory");
      asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(side_addr(1, blk) + off), "f"(Dv.x), "f"(Dv.y) : "memory");
      asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(side_addr(2, blk) + off), "f"(DT.x), "f"(DT.y) : "memory");
      if (wg == 7) tick(9);
    }
    float Xo[8];  // the other horizontal velocity at this level (Coriolis), read while layer 1 runs
#pragma unroll
    for (int r = 0; r < 8; ++r) Xo[r] = 0.f;
    if (qd < 2) lds8(side_addr(3 + (1 - qd), 2 * h), side_addr(3 + (1 - qd), 2 * h + 1), Xo);
    if (diurnal && qd == 2 && lane == 31) {
      const float* Qs = reinterpret_cast<const float*>(smem_tc + g * L.grp_bytes + L.bc) + 8 * h;
#pragma unroll
      for (int r = 0; r < CPT; ++r) bnd[r] = tc_diurnal_top(M, Qs[r], t_stage);
    }
    // ---- layer 1 epilogue ----
    mbar_wait(mbar, parity); parity ^= 1u;
    tc_fence_after();
    tick(1);
    if (wg == 7) tick(10);
    {
      float v[8];
      tmem_ld8(dg + ((uint32_t)(32 * qd) << 16) + 8 * h, v);
      tick(12);
#pragma unroll
      for (int r = 0; r < CPT; ++r) v[r] += b1a;
      if constexpr (AUX) { if (32 * qd + lane < 3 * T.h1) aux_store(a.aux.z1, a.aux.r1, 32 * qd + lane, v); }
#pragma unroll
      for (int r = 0; r < CPT; ++r) v[r] = tc_act<ACT>(act1, v[r]);
      store_row_hilo<CPT>(gbase + L.h1h + h1a, gbase + L.h1l + h1a, v);
      tick(13);
      if (qd == 3 && T.n1b > 0) {
        tmem_ld8(dg + ((uint32_t)96 << 16) + 16 + 8 * h, v);
        if (lane < T.n1b) {
#pragma unroll
          for (int r = 0; r < CPT; ++r) v[r] += b1b;
          if constexpr (AUX) aux_store(a.aux.z1, a.aux.r1, 128 + lane, v);
#pragma unroll
          for (int r = 0; r < CPT; ++r) v[r] = tc_act<ACT>(act1, v[r]);
          store_row_hilo<CPT>(gbase + L.h1h + h1b, gbase + L.h1l + h1b, v);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    tick(14);
    if (wg == 7) tick(11);
    bar_sync_named(bar_id, 256);
    tick(2);
    // ---- layer 2: net q reads rows h1*q.. of H1, result in lanes 32q.. of accumulator block q ----
    if (wg == 3) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint64_t bh = tc_desc(gbase + L.h1h + (uint32_t)(T.k2_start[q] >> 2) * TC_LBO);
          const uint64_t bl = tc_desc(gbase + L.h1l + (uint32_t)(T.k2_start[q] >> 2) * TC_LBO);
          chain(dg + 16 * q, tb + T.c_a2hi, tb + T.c_a2lo, bh, bl, std::integral_constant<int, TC_K2S>{});
        }
        tc_commit(mbar);
      }
      __syncwarp();
    }
    if (qd == 3) {
      parity ^= 1u;  // quadrant 3 has no layer-2 / layer-3 epilogue: it skips those two waits
    } else {
      mbar_wait(mbar, parity); parity ^= 1u;
      tc_fence_after();
      tick(3);
      float v[8];
      tmem_ld8(dg + ((uint32_t)(32 * qd) << 16) + 16 * qd + 8 * h, v);
      if (lane < T.h2) {
#pragma unroll
        for (int r = 0; r < CPT; ++r) v[r] += b2;
        if constexpr (AUX) aux_store(a.aux.z2, a.aux.r2, T.h2 * qd + lane, v);
#pragma unroll
        for (int r = 0; r < CPT; ++r) v[r] = tc_act<ACT>(act2, v[r]);
        store_row_hilo<CPT>(gbase + L.h2h + h2a, gbase + L.h2l + h2a, v);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    bar_sync_named(bar_id, 256);
    tick(4);
    // ---- layer 3 ----
    if (wg == 3) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint64_t bh = tc_desc(gbase + L.h2h + (uint32_t)(T.k3_start[q] >> 2) * TC_LBO);
          const uint64_t bl = tc_desc(gbase + L.h2l + (uint32_t)(T.k3_start[q] >> 2) * TC_LBO);
          chain(dg + 16 * q, tb + T.c_a3hi, tb + T.c_a3lo, bh, bl, std::integral_constant<int, K3S>{});
        }
        tc_commit(mbar);
      }
      __syncwarp();
    }
    if (qd == 3) {
      parity ^= 1u;
    } else {
      mbar_wait(mbar, parity); parity ^= 1u;
      tc_fence_after();
      tick(5);
      // ---- layer 3 epilogue + stencil: flux at face lane+1, divergence at level lane ----
      float nn[8], D[8];
      tmem_ld8(dg + ((uint32_t)(32 * qd) << 16) + 16 * qd + 8 * h, nn);
      if constexpr (CLOSURE) {  // the caller wants the scaled NN flux at face lane+1 itself
#pragma unroll
        for (int r = 0; r < CPT; ++r) dx[r] = nn[r] + b3;
        return;
      }
      lds8(side_addr(qd, 2 * h), side_addr(qd, 2 * h + 1), D);
#pragma unroll
      for (int r = 0; r < CPT; ++r) {
        const float dq = __shfl_down_sync(0xffffffffu, X[r], 1) - X[r];
        float Fup = fmaf(-D[r], dq, nn[r] + b3);
        if (lane == 31) Fup = bnd[r];
        float Fdn = __shfl_up_sync(0xffffffffu, Fup, 1);
        if (lane == 0) Fdn = bnd[r];
        dx[r] = fmaf(-Aq, Fup - Fdn, fmaf(cs, Xo[r], cm));
      }
    }
    tick(6);
    ++ev;
  };

  // Backward-Euler vertical diffusion over one sub-step with the diffusivities of the incoming state (NDE_oceananigans.jl:
  // 61-101): lane <-> level, so the tridiagonal system of each of the thread's columns is solved across the warp by parallel
  // cyclic reduction. The other two fields' profiles come from the full-precision stage-input copies in shared memory.
  auto implicit_step = [&](float hsub) {
    bar_sync_named(bar_id, 256);  // every field's X (= x) of this group is in shared memory
    if constexpr (AUX) {  // the adjoint of this step needs the incoming state (its diffusivities are frozen there)
      if (qd < 3 && ev >= 0) {
        float4* dst = reinterpret_cast<float4*>(a.aux.xp + ((size_t)tile * (a.aux.n_eval / a.aux.n_stages) + ev / a.aux.n_stages) * (size_t)(32 * a.aux.rx)) +
                      (size_t)(cg0 >> 2) * a.aux.rx + 32 * qd + lane;
        dst[0] = make_float4(x[0], x[1], x[2], x[3]);
        dst[a.aux.rx] = make_float4(x[4], x[5], x[6], x[7]);
      }
    }
    if (qd < 3) {
      float Xq[3][8];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (q == qd) {
#pragma unroll
          for (int r = 0; r < 8; ++r) Xq[q][r] = x[r];
        } else {
          lds8(side_addr(3 + q, 2 * h), side_addr(3 + q, 2 * h + 1), Xq[q]);
        }
      }
      float rup[8], dqs[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float du = __shfl_down_sync(0xffffffffu, Xq[0][r], 1) - Xq[0][r];
        const float dv = __shfl_down_sync(0xffffffffu, Xq[1][r], 1) - Xq[1][r];
        const float dT = __shfl_down_sync(0xffffffffu, Xq[2][r], 1) - Xq[2][r];
        float Du = 0.f, Dv = 0.f, DT = 0.f;
        switch (T.side_mode) {
          case SIDE_MPP: side_column<SIDE_MPP>(T.sc, du, dv, dT, Du, Dv, DT); break;
          case SIDE_MPP_CA_T: side_column<SIDE_MPP_CA_T>(T.sc, du, dv, dT, Du, Dv, DT); break;
          case SIDE_MPP_CA_U: side_column<SIDE_MPP_CA_U>(T.sc, du, dv, dT, Du, Dv, DT); break;
          case SIDE_CA_ONLY: side_column<SIDE_CA_ONLY>(T.sc, du, dv, dT, Du, Dv, DT); break;
          default: break;
        }
        const float D = qd == 0 ? Du : (qd == 1 ? Dv : DT);   // Nz c_q nu_q at face lane+1
        rup[r] = lane == 31 ? 0.f : hsub * Aq * D;
        dqs[r] = qd == 0 ? du : (qd == 1 ? dv : dT);          // x[lane+1] - x[lane] of this thread's field
      }
      // incremental form: L (x' - x) = x - L x = r_up (x_up - x) - r_dn (x - x_dn); the small right-hand side keeps the
      // rounding of the solve relative to the increment, not to the state
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float rdn = __shfl_up_sync(0xffffffffu, rup[r], 1);
        const float dqdn = __shfl_up_sync(0xffffffffu, dqs[r], 1);
        if (lane == 0) rdn = 0.f;
        x[r] += pcr32(-rdn, 1.f + rdn + rup[r], -rup[r], rup[r] * dqs[r] - rdn * dqdn);
        X[r] = x[r];
      }
    }
    bar_sync_named(bar_id, 256);  // nobody still reads the old X when it is overwritten
    write_X();
  };

  if constexpr (CLOSURE) {
    const TcClosure& C = ta.cl;
    // 16-byte accesses when every row of every field is 16-byte aligned
    const bool vec4 = (C.ncol & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(C.f[0]) | reinterpret_cast<uintptr_t>(C.f[1]) | reinterpret_cast<uintptr_t>(C.f[2]) |
                        reinterpret_cast<uintptr_t>(C.dzf) | reinterpret_cast<uintptr_t>(C.out)) & 15) == 0;
    // this thread's 8 columns of field qd at level lane, tile tl2 (clamped past the last column)
    auto load_tile = [&](int tl2, float* dst) {
      const int c = tl2 * TC_CT + cg0;
      const float* src = C.f[qd < 3 ? qd : 0] + (size_t)lane * C.ncol;
      if (vec4 && c + 7 < C.ncol) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(src + c)), p1 = __ldg(reinterpret_cast<const float4*>(src + c) + 1);
        dst[0] = p0.x; dst[1] = p0.y; dst[2] = p0.z; dst[3] = p0.w; dst[4] = p1.x; dst[5] = p1.y; dst[6] = p1.z; dst[7] = p1.w;
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) dst[r] = __ldg(src + min(c + r, C.ncol - 1));
      }
    };
    auto store_tile = [&](float* base, int c, const float* v) {
      float* dst = base + ((size_t)qd * 32 + lane) * C.ncol + c;
      if (vec4 && c + 7 < C.ncol) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (c + r < C.ncol) dst[r] = v[r];
      }
    };
    float xnx[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) xnx[r] = 0.f;
    if (qd < 3 && (int)blockIdx.x < C.n_tiles) load_tile(blockIdx.x, xnx);
    // the second column group starts late: the groups loop independently, so one's implicit step then runs under the other's MLP
    if (g == 1 && ta.stagger_ns > 0) __nanosleep(ta.stagger_ns);
    for (int tl = blockIdx.x; tl < C.n_tiles; tl += gridDim.x) {
      const int c0 = tl * TC_CT + cg0;  // first of this thread's 8 columns
      float xin[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        xin[r] = xnx[r];
        if (qd < 3) { x[r] = (xin[r] - C.mu[qd]) * C.inv_sig[qd]; X[r] = x[r]; }
      }
      if (qd < 3 && tl + (int)gridDim.x < C.n_tiles) load_tile(tl + gridDim.x, xnx);  // in flight during this tile's evaluation
      write_X();
      float nnv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) nnv[r] = 0.f;
      rhs_eval(0.f, nnv);
      if (qd < 3) {
        // F = [0; unscaled NN - shift; top flux] on the faces, d/dz at the centres (NDE_oceananigans.jl:281-344): the momentum
        // chains subtract inv(scaling) of the already unscaled first output (:292,301), the temperature chain the first output
        const float sg = C.sig[3 + qd], mu = C.mu[3 + qd];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float un = fmaf(sg, nnv[r], mu);
          const float u0 = __shfl_sync(0xffffffffu, un, 0);
          const float shift = qd < 2 ? fmaf(sg, u0, mu) : u0;
          const float Fup = lane == 31 ? C.top[qd] : un - shift;
          float Fdn = __shfl_up_sync(0xffffffffu, Fup, 1);
          if (lane == 0) Fdn = 0.f;
          nnv[r] = (Fup - Fdn) * C.inv_dz;
        }
        store_tile(C.dzf, c0, nnv);
      }
      implicit_step(C.hsub);  // x <- L(nu(x))^-1 x in the scaled variables (same system as modified_pacanowski_philander!, :61-101)
      if (qd < 3) {
#pragma unroll
        float ov[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          // the dimensional input plus the unscaled increment (not sigma x' + mu: the scale / unscale round trip of the state
          // would cost an order of magnitude in accuracy)
          ov[r] = fmaf(C.sig[qd], x[r] - (xin[r] - C.mu[qd]) * C.inv_sig[qd], xin[r]);
          if (qd == 2 && lane == 0) ov[r] = xin[r];  // T'[bottom] = T_bottom (:93)
        }
        store_tile(C.out, c0, ov);
      }
    }
  } else if constexpr (RHS_ONLY) {
    write_X();
    float dx[8];
    dx[7] = 0.f;
    rhs_eval(a.t_rhs, dx);
    if (qd < 3) {
#pragma unroll
      for (int r = 0; r < CPT; ++r) {
        const int col = col0 + cg0 + r;
        if (col < a.ncol) a.dxdt[(size_t)col * 96 + 32 * qd + lane] = dx[r];
      }
    }
  } else if (active) {
    const float hstep = tm.dt / (float)tm.n_substeps;
    const int ns = tab.n_stages;
    int frame = 0, ci = 0;
    const size_t traj_stride = (size_t)a.n_saved * 96;
    auto save_frame = [&](int fr) {
      if (qd < 3) {
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
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
    if (a.traj != nullptr && tm.save_stride > 0 && !a.skip_frame0) { save_frame(0); frame = 1; }
    if (a.ckpt != nullptr) { save_ckpt(0); ci = 1; }
    write_X();
    // start the second column group roughly half an RHS evaluation late so that its epilogues overlap the first group's
    // MMAs; contention for the tensor pipe keeps the two apart afterwards
    if (g == 1 && ta.stagger_ns > 0 && !split) __nanosleep(ta.stagger_ns);
    for (int n = 0; n < tm.n_steps; ++n) {
      for (int sub = 0; sub < tm.n_substeps; ++sub) {
        const float tbase = tm.t0 + (float)(tm.step0 + n) * tm.dt + (float)sub * hstep;
        if constexpr (IMPL) implicit_step(hstep);
#pragma unroll 1
        for (int i = 0; i < ns; ++i) {
          float dx[8];
          dx[7] = 0.f;
          rhs_eval(tbase + tab.c[i] * hstep, dx);
          if (qd < 3) {
            const bool last = (i + 1 == ns);
            // k_i for the adjoint's reverse sweep (same tile-native layout as the checkpoints); checkpointing solves
            // always run the 32-column tiles, so the 28-column instantiation carries no trace of this
            if (CPT == 8 && a.kstore != nullptr) {
              const size_t rk = (size_t)n * tm.n_substeps + sub, nrk = (size_t)tm.n_steps * tm.n_substeps;
              float4* dst = reinterpret_cast<float4*>(a.kstore + ((((size_t)tile * nrk + rk) * ns + i) * 96 + 32 * qd + lane) * TC_CT + cg0);
              __stcs(dst, make_float4(dx[0], dx[1], dx[2], dx[3]));
              __stcs(dst + 1, make_float4(dx[4], dx[5], dx[6], dx[7]));
            }
            float acc[8];
            const float ci_ = last ? tab.b[i] : tab.a[(i + 1) % CPZ_MAX_STAGES][i];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = r < CPT ? ci_ * dx[r] : 0.f;
#pragma unroll 1
            for (int j = 0; j < i; ++j) {
              const float cj = last ? tab.b[j] : tab.a[(i + 1) % CPZ_MAX_STAGES][j];
              const float4 p0 = lds_v4(ks_base + j * ks_stride), p1 = lds_v4(ks_base + j * ks_stride + ks_half);
              acc[0] = fmaf(cj, p0.x, acc[0]); acc[1] = fmaf(cj, p0.y, acc[1]); acc[2] = fmaf(cj, p0.z, acc[2]); acc[3] = fmaf(cj, p0.w, acc[3]);
              acc[4] = fmaf(cj, p1.x, acc[4]); acc[5] = fmaf(cj, p1.y, acc[5]); acc[6] = fmaf(cj, p1.z, acc[6]); acc[7] = fmaf(cj, p1.w, acc[7]);
            }
            if (!last) {
              sts_v4(ks_base + i * ks_stride, dx[0], dx[1], dx[2], dx[3]);
              sts_v4(ks_base + i * ks_stride + ks_half, dx[4], dx[5], dx[6], dx[7]);
#pragma unroll
              for (int r = 0; r < CPT; ++r) X[r] = fmaf(hstep, acc[r], x[r]);
            } else {
#pragma unroll
              for (int r = 0; r < CPT; ++r) { x[r] = fmaf(hstep, acc[r], x[r]); X[r] = x[r]; }
            }
            write_X();
          }
          tick(7);
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

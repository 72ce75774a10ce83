// Dense-layer phases of the MLP: out[N][CT] = act(W^T in[K][CT] + b) as a register-tiled FP32 contraction.
//
// Measured on B200 (tools/micro/ffma2_bench.cu, profiles/r01_micro_ffma.txt): one SM sub-partition issues a scalar FFMA
// every ~1.05 cycles but an FFMA2 only every ~2.6-2.9 (so scalar FFMA is used), and shared-memory wavefronts — not
// instruction issue — bound a register-tiled GEMM: a thread tile of TC columns x TO outputs costs (TC+TO) wavefronts per
// 32*TC*TO FMAs. The tile is therefore large (8 x 10 for the 50-wide layers) and the parallelism comes from splitting K
// over KQ lanes of the SAME warp, whose partial sums are combined with a shuffle reduce-scatter (no shared memory, no
// block barrier); each lane then finishes (bias, activation, store) its 1/KQ share of the tile's output rows.
#pragma once

namespace cpz {

__device__ __forceinline__ float4 act_fwd4(int act, float4 z) {
  float4 r;
  switch (act) {
    case ACT_RELU: r = make_float4(act_fwd(ACT_RELU, z.x), act_fwd(ACT_RELU, z.y), act_fwd(ACT_RELU, z.z), act_fwd(ACT_RELU, z.w)); break;  // NaN-propagating
    case ACT_MISH: r = make_float4(act_fwd(ACT_MISH, z.x), act_fwd(ACT_MISH, z.y), act_fwd(ACT_MISH, z.z), act_fwd(ACT_MISH, z.w)); break;
    case ACT_SWISH: r = make_float4(act_fwd(ACT_SWISH, z.x), act_fwd(ACT_SWISH, z.y), act_fwd(ACT_SWISH, z.z), act_fwd(ACT_SWISH, z.w)); break;
    case ACT_LEAKY: r = make_float4(act_fwd(ACT_LEAKY, z.x), act_fwd(ACT_LEAKY, z.y), act_fwd(ACT_LEAKY, z.z), act_fwd(ACT_LEAKY, z.w)); break;
    case ACT_TANH: r = make_float4(act_fwd(ACT_TANH, z.x), act_fwd(ACT_TANH, z.y), act_fwd(ACT_TANH, z.z), act_fwd(ACT_TANH, z.w)); break;
    default: r = z; break;
  }
  return r;
}

// Per-thread description of its tile in one phase. It depends only on (threadIdx, plan), not on the data, so it is
// computed once per kernel and kept in registers: the phase functions are entered ~14,000 times per solve and the
// dependent chain "thread id -> tile -> gemm -> offsets" would otherwise be paid (at full latency, 2 warps per
// sub-partition) on every entry.
struct TileCtx {
  int in_off;     // float offset of the tile's input columns (arena-relative, or X-relative when from_x)
  int out_off;    // float offset of the tile's output columns: arena + out_off + row*CT
  int w_off;      // weight offset of W[0][j0] (shared arena when WS, theta otherwise)
  int b_off;      // bias offset of b[j0]
  int k0, k1;     // this lane's k range
  int ldw;        // weight row stride
  int nrow;       // valid output rows of this lane's share: rows [row0, row0+nrow) of the tile
  int row0;       // first tile row this lane finishes
  int act;        // activation, bit 8: input is the state tile X, bit 9: lane is active
  int valid;      // real output rows of the tile (<= TO)
};

// TileCtx packed into 5 registers (offsets are < 2^16 floats of shared memory or < 2^26 floats of theta).
struct TilePack { int a, b, c, d, e; };
__device__ __forceinline__ TilePack pack_ctx(const TileCtx& t) {
  TilePack q;
  q.a = t.in_off | (t.out_off << 16);
  q.b = t.w_off;
  q.c = t.b_off;
  q.d = t.k0 | (t.k1 << 12) | (t.nrow << 24) | (t.valid << 28);
  q.e = t.ldw | (t.row0 << 16) | (t.act << 20);
  return q;
}
__device__ __forceinline__ TileCtx unpack_ctx(const TilePack& q) {
  TileCtx t;
  t.in_off = q.a & 0xffff; t.out_off = (q.a >> 16) & 0xffff;
  t.w_off = q.b; t.b_off = q.c;
  t.k0 = q.d & 0xfff; t.k1 = (q.d >> 12) & 0xfff; t.nrow = (q.d >> 24) & 0xf; t.valid = (q.d >> 28) & 0xf;
  t.ldw = q.e & 0xffff; t.row0 = (q.e >> 16) & 0xf; t.act = (q.e >> 20) & 0xfff;
  return t;
}

template <int TC, int TO, int KQ, bool WS, int CT, int NT>
__device__ __forceinline__ TileCtx make_tile_ctx(const ModelD& M, int p, int round) {
  constexpr int NCG = CT / TC, TPW = 32 / KQ;
  constexpr int RP = ((TO + KQ - 1) / KQ) * KQ, RS = RP / KQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kq = lane / TPW, tl = lane - kq * TPW;
  const int n_tiles = M.phase[p].n_tiles;
  const int tile_raw = (warp + round * (NT / 32)) * TPW + tl;
  const bool active = tile_raw < n_tiles;
  const int tile = active ? tile_raw : n_tiles - 1;
  int gi = M.phase[p].g0;
  while (gi + 1 < M.phase[p].g1 && tile >= M.gemm[gi + 1].tile_begin) ++gi;
  const GemmD& g = M.gemm[gi];
  const int local = tile - g.tile_begin;
  const int cg = local % NCG, og = local / NCG, j0 = og * TO;
  TileCtx t;
  t.in_off = (g.in_off < 0 ? 0 : g.in_off * CT) + TC * cg;
  t.out_off = g.out_off * CT + TC * cg + j0 * CT;
  t.ldw = WS ? g.Npad : g.N;
  t.w_off = (WS ? g.sw_off : g.w_off) + j0;
  t.b_off = (WS ? g.sb_off : g.b_off) + j0;
  t.k0 = (g.K * kq) / KQ;
  t.k1 = (g.K * (kq + 1)) / KQ;
  t.row0 = kq * RS;
  const int valid = min(TO, g.N - j0);           // real rows of this tile
  t.valid = valid;
  t.nrow = max(0, min(RS, valid - t.row0));
  t.act = g.act | (g.in_off < 0 ? 256 : 0) | (active ? 512 : 0);
  if (!active) t.nrow = 0;
  return t;
}

// One tile: k loop (scalar FFMA register tile), shuffle reduce-scatter over the KQ k-split lanes, finish own rows.
template <int TC, int TO, int KQ, bool WS, int CT, bool SAVE_Z>
__device__ __noinline__ void run_tile_g2(const TileCtx t, const float* __restrict__ X, float* __restrict__ arena,
                                         float* __restrict__ zarena, const float* __restrict__ wbase) {
  static_assert(TC % 4 == 0 && TO % 2 == 0 && (KQ == 1 || KQ == 2 || KQ == 4), "tile shape");
  constexpr int RP = ((TO + KQ - 1) / KQ) * KQ;   // output rows padded to a multiple of KQ
  constexpr int RS = RP / KQ;                     // rows each lane finishes
  const bool lead = t.row0 == 0;                  // kq == 0 adds the bias
  const float* in = ((t.act & 256) ? X : arena) + t.in_off;
  const float* W = wbase + t.w_off;
  float acc[RP][TC];
#pragma unroll
  for (int r = 0; r < RP; ++r) {
    float bv = 0.f;
    if (r < TO && lead) bv = WS ? wbase[t.b_off + r] : __ldg(wbase + t.b_off + min(r, t.valid - 1));
#pragma unroll
    for (int c = 0; c < TC; ++c) acc[r][c] = bv;
  }
  const int ldw = t.ldw;
  const float* xp = in + t.k0 * CT;
  const float* wp = W + (size_t)t.k0 * ldw;
  if constexpr (WS) {
#pragma unroll 4
    for (int k = t.k0; k < t.k1; ++k) {
      float xv[TC], w[TO];
#pragma unroll
      for (int c = 0; c < TC; c += 4) {
        const float4 q = *reinterpret_cast<const float4*>(xp + c);
        xv[c] = q.x; xv[c + 1] = q.y; xv[c + 2] = q.z; xv[c + 3] = q.w;
      }
      if constexpr (TO % 4 == 0) {
#pragma unroll
        for (int o = 0; o < TO; o += 4) {
          const float4 q = *reinterpret_cast<const float4*>(wp + o);
          w[o] = q.x; w[o + 1] = q.y; w[o + 2] = q.z; w[o + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int o = 0; o < TO; o += 2) {
          const float2 q = *reinterpret_cast<const float2*>(wp + o);
          w[o] = q.x; w[o + 1] = q.y;
        }
      }
#pragma unroll
      for (int o = 0; o < TO; ++o)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[o][c] = fmaf(w[o], xv[c], acc[o][c]);
      xp += CT;
      wp += ldw;
    }
  } else {
    // weights streamed from global memory (L2): rows past the layer width are clamped (their results are discarded)
    int jo[TO];
#pragma unroll
    for (int o = 0; o < TO; ++o) jo[o] = min(o, t.valid - 1);
#pragma unroll 2
    for (int k = t.k0; k < t.k1; ++k) {
      float xv[TC], w[TO];
#pragma unroll
      for (int c = 0; c < TC; c += 4) {
        const float4 q = *reinterpret_cast<const float4*>(xp + c);
        xv[c] = q.x; xv[c + 1] = q.y; xv[c + 2] = q.z; xv[c + 3] = q.w;
      }
#pragma unroll
      for (int o = 0; o < TO; ++o) w[o] = __ldg(wp + jo[o]);
#pragma unroll
      for (int o = 0; o < TO; ++o)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[o][c] = fmaf(w[o], xv[c], acc[o][c]);
      xp += CT;
      wp += ldw;
    }
  }
  // ---- combine the KQ partial tiles: shuffle reduce-scatter over the k-split lanes of this warp
  float fin[RS][TC];
  if constexpr (KQ == 1) {
#pragma unroll
    for (int r = 0; r < RS; ++r)
#pragma unroll
      for (int c = 0; c < TC; ++c) fin[r][c] = acc[r][c];
  } else if constexpr (KQ == 2) {
    const bool hi = t.row0 != 0;
#pragma unroll
    for (int r = 0; r < RS; ++r)
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float send = hi ? acc[r][c] : acc[RS + r][c];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        fin[r][c] = (hi ? acc[RS + r][c] : acc[r][c]) + recv;
      }
  } else {
    constexpr int RH = RP / 2;
    const int kq = t.row0 / RS;
    const bool b1 = (kq & 2) != 0, b0 = (kq & 1) != 0;
    float half[RH][TC];
#pragma unroll
    for (int r = 0; r < RH; ++r)
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float send = b1 ? acc[r][c] : acc[RH + r][c];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        half[r][c] = (b1 ? acc[RH + r][c] : acc[r][c]) + recv;
      }
#pragma unroll
    for (int r = 0; r < RS; ++r)
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float send = b0 ? half[r][c] : half[RS + r][c];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        fin[r][c] = (b0 ? half[RS + r][c] : half[r][c]) + recv;
      }
  }
  // ---- finish this lane's rows
  float* out = arena + t.out_off + t.row0 * CT;
  float* zo = SAVE_Z ? zarena + t.out_off + t.row0 * CT : nullptr;
  const int act = t.act & 255;
#pragma unroll
  for (int r = 0; r < RS; ++r) {
    if (r < t.nrow) {
#pragma unroll
      for (int c = 0; c < TC; c += 4) {
        const float4 z = make_float4(fin[r][c], fin[r][c + 1], fin[r][c + 2], fin[r][c + 3]);
        if constexpr (SAVE_Z) *reinterpret_cast<float4*>(zo + r * CT + c) = z;
        *reinterpret_cast<float4*>(out + r * CT + c) = act_fwd4(act, z);
      }
    }
  }
}

// Shapes the planner may choose (keep in sync with kShapes in cpz_plan.cpp).
#define CPZ_SHAPES(X) X(8, 10, 4) X(8, 10, 2) X(8, 10, 1) X(8, 8, 4) X(8, 8, 2) X(8, 8, 1) X(8, 4, 4) X(8, 4, 2) X(4, 4, 4) X(4, 4, 2) X(4, 4, 1) X(4, 2, 2)

constexpr int CPZ_CTX_PHASES = 4;  // round-0 tile contexts of the first phases are cached in registers (5 each)

template <bool WS, int CT, int NT>
__device__ __forceinline__ TileCtx phase_ctx(const ModelD& M, int p, int round) {
  TileCtx t{};
#define CPZ_CASE(tc, to, kq) \
  case (tc * 10000 + to * 100 + kq): if constexpr (tc <= CT) t = make_tile_ctx<tc, to, kq, WS, CT, NT>(M, p, round); break;
  switch (M.phase[p].TC * 10000 + M.phase[p].TO * 100 + M.phase[p].ksplit) {
    CPZ_SHAPES(CPZ_CASE)
    default: break;
  }
#undef CPZ_CASE
  return t;
}

__device__ __forceinline__ int phase_rounds(const ModelD& M, int p, int NT) {
  const int tpw = 32 / M.phase[p].ksplit;
  const int per_round = (NT / 32) * tpw;
  return (M.phase[p].n_tiles + per_round - 1) / per_round;
}

struct PhaseCache {
  TilePack tp[CPZ_CTX_PHASES];
  int shape[CPZ_CTX_PHASES];   // TC*10000 + TO*100 + KQ
  int rounds[CPZ_CTX_PHASES];
};
template <bool WS, int CT, int NT>
__device__ __forceinline__ void build_phase_cache(const ModelD& M, PhaseCache& pc) {
#pragma unroll
  for (int p = 0; p < CPZ_CTX_PHASES; ++p) {
    if (p < M.n_phase) {
      pc.tp[p] = pack_ctx(phase_ctx<WS, CT, NT>(M, p, 0));
      pc.shape[p] = M.phase[p].TC * 10000 + M.phase[p].TO * 100 + M.phase[p].ksplit;
      pc.rounds[p] = phase_rounds(M, p, NT);
    } else {
      pc.tp[p] = TilePack{0, 0, 0, 0, 0}; pc.shape[p] = 0; pc.rounds[p] = 0;
    }
  }
}

// Runs phase p. `tp0`/`shape`/`rounds` come from the PhaseCache when p < CPZ_CTX_PHASES (pass rounds < 0 otherwise).
template <bool WS, int CT, int NT, bool SAVE_Z>
__device__ __forceinline__ void run_phase(const ModelD& M, int p, const TilePack& tp0, int shape, int rounds,
                                          const float* __restrict__ X, float* __restrict__ arena,
                                          float* __restrict__ zarena, const float* __restrict__ wsm,
                                          const float* __restrict__ theta) {
  const float* wbase = WS ? wsm : theta;
  const bool cached = rounds >= 0;
  if (!cached) {
    rounds = phase_rounds(M, p, NT);
    shape = M.phase[p].TC * 10000 + M.phase[p].TO * 100 + M.phase[p].ksplit;
  }
  for (int r = 0; r < rounds; ++r) {
    const TileCtx t = (cached && r == 0) ? unpack_ctx(tp0) : phase_ctx<WS, CT, NT>(M, p, r);
#define CPZ_CASE(tc, to, kq) \
  case (tc * 10000 + to * 100 + kq): if constexpr (tc <= CT) run_tile_g2<tc, to, kq, WS, CT, SAVE_Z>(t, X, arena, zarena, wbase); break;
    switch (shape) {
      CPZ_SHAPES(CPZ_CASE)
      default: break;
    }
#undef CPZ_CASE
  }
}

}  // namespace cpz

// Small-batch training pass of the T-only NDE (BASELINE config 1: ONE column, convective adjustment + mPP base, forward solve +
// loss gradient): one CTA per column, no column padding.
//
// Replaces, for a handful of columns, what the reference does per simulation in train_neural_differential_equation!
// (free_convection/src/training.jl:44-71: solve_nde + Flux.mse + Zygote gradient through InterpolatingAdjoint) for the RHS of
// ConvectiveAdjustmentNDE (free_convection/src/convective_adjustment_nde.jl:33-48) with the mPP base of
// wind_mixing/src/NDE_training.jl:114-139 at u = v = 0. Same discrete adjoint as adjoint_kernel (cpz_adjoint.cuh); the
// difference is the mapping: with one column there is no batch dimension to tile over, so
//   * the weights stay in shared memory as a plain copy of theta (Flux's [in][out] order) and every layer is a mat-vec whose
//     outputs are spread over the 512 threads (output o, K quarter kq), partial sums combined through shared memory;
//   * every thread owns a FIXED set of weight-gradient entries in registers for the whole reverse sweep
//     (dW2: 32, dW1: 8, dW3: 8, biases: 3) — no gradient slab is read-modify-written per stage;
//   * the transposed products of the delta propagation read the same weight copy with a lane skew that keeps the
//     shared-memory banks distinct.
// Step checkpoints every ckpt_stride steps; a segment's sub-step start states are re-integrated into shared memory, each
// step's stage records (stage input, z1, a1, z2, a2) are recomputed right before its reverse stages.
#pragma once
#include "cpz_device.cuh"

namespace cpz {

constexpr int FC1_NT = 512;

struct Fc1D {
  int h1, h2;        // hidden widths (<= 128); input 32, output 31
  int act1, act2;
  int w_off[3], b_off[3];
  int P;
};

struct Fc1Args {
  const float* theta;
  const float* x0;       // [ncol][32] (row stride x0_stride, 0 = 32)
  size_t x0_stride;
  const float* bcs;      // [ncol][2]
  const float* targets;  // [ncol][n_saved][32]
  float* ckpt;           // [ncol][n_seg + 1][32]: segment start states, then the final state
  float* gpart;          // [ncol][P]: d(unnormalised loss of this column)/dtheta
  float* lpart;          // [ncol][8]: squared-error sum of the T profiles at index 2
  int ncol, n_saved, n_seg;
  float wT, inv_prof;
};

struct Fc1Smem {
  int w, xs, xbar, ks, yb, ys, z1, a1, z2, a2, part, d1, d2, d3, nn, segx, total_floats;
};
__host__ __device__ inline Fc1Smem fc1_smem_layout(const Fc1D& F, int n_stages, int seg_states) {
  Fc1Smem L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  L.w = take(F.P);
  L.xs = take(32); L.xbar = take(32);
  L.ks = take(n_stages * 32);   // forward: stage tendencies k_i
  L.yb = take(n_stages * 32);   // reverse: cotangents of the stage inputs
  L.ys = take(n_stages * 32);   // stage inputs
  L.z1 = take(n_stages * 128); L.a1 = take(n_stages * 128);
  L.z2 = take(n_stages * 128); L.a2 = take(n_stages * 128);
  L.part = take(4 * 128);       // partial sums: [4][128] or [16][32]
  L.d1 = take(128); L.d2 = take(128); L.d3 = take(32);
  L.nn = take(32);
  L.segx = take(seg_states * 32);
  L.total_floats = o;
  return L;
}

__global__ void __launch_bounds__(FC1_NT, 1) fc1_train_kernel(const __grid_constant__ ModelD M, const __grid_constant__ Fc1D F,
                                                              const __grid_constant__ TableauD tab, const TimeD tm,
                                                              const __grid_constant__ Fc1Args a) {
  extern __shared__ __align__(16) float smem[];
  const int ns = tab.n_stages, nsub = tm.n_substeps, cs = tm.ckpt_stride;
  const Fc1Smem L = fc1_smem_layout(F, ns, cs * nsub);
  float* wsm = smem + L.w;
  float* xs = smem + L.xs;
  float* xbar = smem + L.xbar;
  float* ks = smem + L.ks;
  float* yb = smem + L.yb;
  float* ys = smem + L.ys;
  float* z1s = smem + L.z1;
  float* a1s = smem + L.a1;
  float* z2s = smem + L.z2;
  float* a2s = smem + L.a2;
  float* part = smem + L.part;
  float* d1s = smem + L.d1;
  float* d2s = smem + L.d2;
  float* d3s = smem + L.d3;
  float* nns = smem + L.nn;
  float* segx = smem + L.segx;

  const int t = threadIdx.x, lane = t & 31;
  const int o128 = t & 127, kq4 = t >> 7;   // (output, K quarter) of the 128-wide layers
  const int o32 = t & 31, kq16 = t >> 5;    // (output, K sixteenth) of the 31-wide output layer
  const int col = blockIdx.x;
  const int h1 = F.h1, h2 = F.h2;
  const float* W1 = wsm + F.w_off[0];  // [32][h1]
  const float* W2 = wsm + F.w_off[1];  // [h1][h2]
  const float* W3 = wsm + F.w_off[2];  // [h2][31]
  const float* B1 = wsm + F.b_off[0];
  const float* B2 = wsm + F.b_off[1];
  const float* B3 = wsm + F.b_off[2];
  const float Nf = M.rc.Nf, AN = M.rc.A[2] * M.rc.Nf;
  const bool mpp = (M.flags & F_MPP) != 0, ca = (M.flags & F_CA) != 0;
  const float hstep = tm.dt / (float)nsub;

  for (int i = t; i < F.P; i += FC1_NT) wsm[i] = __ldg(a.theta + i);
  if (t < 32) {
    xs[t] = __ldg(a.x0 + (size_t)col * (a.x0_stride ? a.x0_stride : (size_t)32) + t);
    xbar[t] = 0.f;
  }
  const float bc0 = __ldg(a.bcs + (size_t)col * 2), bc1 = __ldg(a.bcs + (size_t)col * 2 + 1);
  __syncthreads();

  // ---- MLP forward at the stage input y; the pre-activations / activations go to z1o, a1o, z2o, a2o; result in nns[0..30] ----
  auto mlp_forward = [&](const float* __restrict__ y, float* __restrict__ z1o, float* __restrict__ a1o, float* __restrict__ z2o,
                         float* __restrict__ a2o) {
    {  // layer 1: K = 32 split in four
      float acc = 0.f;
      if (o128 < h1) {
        const float* w = W1 + (8 * kq4) * h1 + o128;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(w[i * h1], y[8 * kq4 + i], acc);
      }
      part[kq4 * 128 + o128] = acc;
    }
    __syncthreads();
    if (t < 128) {
      float z = 0.f, av = 0.f;
      if (t < h1) {
        z = (part[t] + part[128 + t]) + (part[256 + t] + part[384 + t]) + B1[t];
        av = act_fwd(F.act1, z);
      }
      z1o[t] = z; a1o[t] = av;
    }
    __syncthreads();
    {  // layer 2: K = h1 (<= 128) split in four
      float acc0 = 0.f, acc1 = 0.f;
      if (o128 < h2) {
        const int k0 = 32 * kq4, kn = min(32, h1 - k0);
        const float* w = W2 + k0 * h2 + o128;
        int i = 0;
        for (; i + 1 < kn; i += 2) {
          acc0 = fmaf(w[i * h2], a1o[k0 + i], acc0);
          acc1 = fmaf(w[(i + 1) * h2], a1o[k0 + i + 1], acc1);
        }
        if (i < kn) acc0 = fmaf(w[i * h2], a1o[k0 + i], acc0);
      }
      part[kq4 * 128 + o128] = acc0 + acc1;
    }
    __syncthreads();
    if (t < 128) {
      float z = 0.f, av = 0.f;
      if (t < h2) {
        z = (part[t] + part[128 + t]) + (part[256 + t] + part[384 + t]) + B2[t];
        av = act_fwd(F.act2, z);
      }
      z2o[t] = z; a2o[t] = av;
    }
    __syncthreads();
    {  // layer 3: 31 outputs, K = h2 split in sixteen
      float acc = 0.f;
      if (o32 < 31) {
        const int k0 = 8 * kq16, kn = min(8, h2 - k0);
        const float* w = W3 + k0 * 31 + o32;
        for (int i = 0; i < kn; ++i) acc = fmaf(w[i * 31], a2o[k0 + i], acc);
      }
      part[kq16 * 32 + o32] = acc;
    }
    __syncthreads();
    if (t < 32) {
      float s = 0.f;
      if (t < 31) {
#pragma unroll
        for (int q = 0; q < 16; ++q) s += part[q * 32 + t];
        s += B3[t];
      }
      nns[t] = s;
    }
    __syncwarp();
  };

  // tendency of level `lane` from the stage input y and the NN fluxes in nns (warp 0 only)
  auto tendency = [&](const float* __restrict__ y) -> float {
    float Fl = bc0;  // flux at face lane (face 0: bottom boundary)
    if (lane >= 1) {
      const float G = Nf * (y[lane] - y[lane - 1]);
      Fl = nns[lane - 1];
      if (mpp) Fl -= fc_mpp_cnu(M, G) * G;
      if (ca) Fl -= fminf(0.f, M.rc.K_ca * G);
    }
    float Fu = __shfl_down_sync(0xffffffffu, Fl, 1);
    if (lane == 31) Fu = bc1;
    return -AN * (Fu - Fl);
  };

  // one Runge–Kutta step from xs (in place); with RECORD the stage records stay in ys / z1s / a1s / z2s / a2s
  auto rk_step = [&](bool record) {
    for (int i = 0; i < ns; ++i) {
      float* y = ys + (record ? i : 0) * 32;
      if (t < 32) {
        float yv = 0.f;
        for (int j = 0; j < i; ++j) yv = fmaf(tab.a[i][j], ks[j * 32 + t], yv);
        y[t] = fmaf(hstep, yv, xs[t]);
      }
      __syncthreads();
      const int r = record ? i : 0;
      mlp_forward(y, z1s + r * 128, a1s + r * 128, z2s + r * 128, a2s + r * 128);
      if (t < 32) ks[i * 32 + t] = tendency(y);
      __syncwarp();
    }
    if (t < 32) {
      float acc = 0.f;
      for (int i = 0; i < ns; ++i) acc = fmaf(tab.b[i], ks[i * 32 + t], acc);
      xs[t] = fmaf(hstep, acc, xs[t]);
    }
    __syncthreads();
  };

  // ---- forward pass with segment checkpoints ----------------------------------------------------------------------------
  float* ck = a.ckpt + (size_t)col * (a.n_seg + 1) * 32;
  for (int n = 0; n < tm.n_steps; ++n) {
    if (n % cs == 0 && t < 32) ck[(n / cs) * 32 + t] = xs[t];
    for (int sub = 0; sub < nsub; ++sub) rk_step(false);
  }
  if (t < 32) ck[a.n_seg * 32 + t] = xs[t];

  // ---- reverse sweep -------------------------------------------------------------------------------------------------------
  auto frame_of = [&](int step) -> int {
    if (tm.save_stride <= 0) return step == tm.n_steps ? 0 : -1;
    return (step % tm.save_stride == 0) ? step / tm.save_stride : -1;
  };
  float sse = 0.f;
  auto loss_frame = [&](const float* __restrict__ x, int fr) {  // warp 0
    const float d = x[lane] - __ldg(a.targets + ((size_t)col * a.n_saved + fr) * 32 + lane);
    sse = fmaf(d, d, sse);
    xbar[lane] += a.wT * 2.f * a.inv_prof * d;
  };
  float acc1[8], acc2[32], acc3[8], db1 = 0.f, db2 = 0.f, db3 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc1[i] = 0.f; acc3[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc2[i] = 0.f;
  const int skew2 = (h2 & 1) ? 0 : lane;  // lane skew of the transposed layer-2 reads: (row stride + skew step) must be odd

  if (t < 32) {
    const int fr = frame_of(tm.n_steps);
    if (fr >= 0) loss_frame(xs, fr);
  }
  for (int seg = a.n_seg - 1; seg >= 0; --seg) {
    const int step0 = seg * cs, steps = min(cs, tm.n_steps - step0), R = steps * nsub;
    __syncthreads();
    if (t < 32) xs[t] = ck[seg * 32 + t];
    __syncthreads();
    for (int r = 0; r < R; ++r) {  // sub-step start states of the segment
      if (t < 32) segx[r * 32 + t] = xs[t];
      if (r + 1 < R) rk_step(false);
    }
    for (int r = R - 1; r >= 0; --r) {
      __syncthreads();
      if (t < 32) xs[t] = segx[r * 32 + t];
      __syncthreads();
      rk_step(true);  // stage records of this step (xs moves on to the step's end state; its start stays in segx)
      for (int i = ns - 1; i >= 0; --i) {
        const float* y = ys + i * 32;
        // B0 (warp 0): kbar_i, cotangent of the face fluxes = delta3, direct part of Ybar_i through the diffusive flux
        if (t < 32) {
          float kb = tab.b[i] * xbar[t];
          for (int j = i + 1; j < ns; ++j) kb = fmaf(tab.a[j][i], yb[j * 32 + t], kb);
          kb *= hstep;
          const float kbm = __shfl_up_sync(0xffffffffu, kb, 1);
          float Fb = 0.f, Gb = 0.f;  // face lane (>= 1)
          if (lane >= 1) {
            Fb = -AN * (kbm - kb);
            const float G = Nf * (y[lane] - y[lane - 1]);
            float dFdG = 0.f;
            if (mpp) { float dc; const float c = fc_mpp_cnu(M, G, &dc); dFdG -= fmaf(G, dc, c); }
            if (ca && M.rc.K_ca * G < 0.f) dFdG -= M.rc.K_ca;
            Gb = Fb * dFdG;
            d3s[lane - 1] = Fb;
          } else {
            d3s[31] = 0.f;
          }
          float Gbu = __shfl_down_sync(0xffffffffu, Gb, 1);
          if (lane == 31) Gbu = 0.f;
          yb[i * 32 + t] = Nf * (Gb - Gbu);
        }
        __syncthreads();
        // B1: dW3 += a2 d3^T, db3; abar2 = W3 d3
        {
          const float d3 = d3s[o32];
          const float* av = a2s + i * 128 + 8 * kq16;
#pragma unroll
          for (int q = 0; q < 8; ++q) acc3[q] = fmaf(av[q], d3, acc3[q]);
          if (t < 31) db3 += d3s[t];
          float acc = 0.f;
          if (o128 < h2) {
            const int ob = 8 * kq4, on = min(8, 31 - ob);
            const float* w = W3 + o128 * 31 + ob;
            for (int q = 0; q < on; ++q) acc = fmaf(w[q], d3s[ob + q], acc);
          }
          part[kq4 * 128 + o128] = acc;
        }
        __syncthreads();
        if (t < 128) {
          float d = 0.f;
          if (t < h2) d = ((part[t] + part[128 + t]) + (part[256 + t] + part[384 + t])) * act_grad(F.act2, z2s[i * 128 + t]);
          d2s[t] = d;
          db2 += d;
        }
        __syncthreads();
        // B3: dW2 += a1 d2^T; abar1 = W2 d2
        {
          const float d2 = d2s[o128];
          const float4* av = reinterpret_cast<const float4*>(a1s + i * 128 + 32 * kq4);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = av[q];
            acc2[4 * q] = fmaf(v.x, d2, acc2[4 * q]); acc2[4 * q + 1] = fmaf(v.y, d2, acc2[4 * q + 1]);
            acc2[4 * q + 2] = fmaf(v.z, d2, acc2[4 * q + 2]); acc2[4 * q + 3] = fmaf(v.w, d2, acc2[4 * q + 3]);
          }
          float s0 = 0.f, s1 = 0.f;
          if (o128 < h1) {
            const int ob = 32 * kq4;
            const float* w = W2 + o128 * h2 + ob;
            const float* dv = d2s + ob;
#pragma unroll 4
            for (int q = 0; q < 32; q += 2) {
              const int oa = (q + skew2) & 31, oc = (q + 1 + skew2) & 31;
              if (ob + oa < h2) s0 = fmaf(w[oa], dv[oa], s0);
              if (ob + oc < h2) s1 = fmaf(w[oc], dv[oc], s1);
            }
          }
          part[kq4 * 128 + o128] = s0 + s1;
        }
        __syncthreads();
        if (t < 128) {
          float d = 0.f;
          if (t < h1) d = ((part[t] + part[128 + t]) + (part[256 + t] + part[384 + t])) * act_grad(F.act1, z1s[i * 128 + t]);
          d1s[t] = d;
          db1 += d;
        }
        __syncthreads();
        // B5: dW1 += y d1^T; Ybar_i += W1 d1 (thread (k = t / 16, oq = t % 16): outputs oq, oq + 16, ...)
        {
          const float d1 = d1s[o128];
          const float* yv = y + 8 * kq4;
#pragma unroll
          for (int q = 0; q < 8; ++q) acc1[q] = fmaf(yv[q], d1, acc1[q]);
          const int k = t >> 4, oq = t & 15;
          const float* w = W1 + k * h1;
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int o = oq + 16 * q;
            if (o < h1) s = fmaf(w[o], d1s[o], s);
          }
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          if (oq == 0) yb[i * 32 + k] += s;
        }
        __syncthreads();
      }
      if (t < 32) {
        float s = xbar[t];
        for (int i = 0; i < ns; ++i) s += yb[i * 32 + t];
        xbar[t] = s;
        if (r % nsub == 0) {
          const int fr = frame_of(step0 + r / nsub);
          if (fr >= 0) loss_frame(segx + r * 32, fr);
        }
      }
    }
  }

  // ---- results --------------------------------------------------------------------------------------------------------------
  float* gp = a.gpart + (size_t)col * F.P;
  if (o128 < h1) {
#pragma unroll
    for (int q = 0; q < 8; ++q) gp[F.w_off[0] + (8 * kq4 + q) * h1 + o128] = acc1[q];
  }
  if (o128 < h2) {
#pragma unroll
    for (int q = 0; q < 32; ++q)
      if (32 * kq4 + q < h1) gp[F.w_off[1] + (32 * kq4 + q) * h2 + o128] = acc2[q];
  }
  if (o32 < 31) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (8 * kq16 + q < h2) gp[F.w_off[2] + (8 * kq16 + q) * 31 + o32] = acc3[q];
  }
  if (t < h1) gp[F.b_off[0] + t] = db1;
  if (t < h2) gp[F.b_off[1] + t] = db2;
  if (t < 31) gp[F.b_off[2] + t] = db3;
  if (t < 32) {
    for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
    if (t < 8) a.lpart[(size_t)col * 8 + t] = t == 2 ? sse : 0.f;
  }
}

// out[p] = sum over columns of gpart[col][p] (fixed order)
static __global__ void fc1_reduce_kernel(const float* __restrict__ gpart, int ncol, int P, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int c = 0; c < ncol; ++c) s += gpart[(size_t)c * P + p];
  out[p] = s;
}

}  // namespace cpz

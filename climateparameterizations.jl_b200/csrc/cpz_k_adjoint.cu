// Adjoint kernel instantiations, small reduction / optimiser kernels and their launchers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "cpz_adjoint_launch.h"

namespace cpz {

// ---- small reductions -----------------------------------------------------------------------------------------------
// out[p] = sum over CTA slabs of part[slab][map[p]]  (fixed order: deterministic); map translates destructure order into
// the adjoint's 4x4-tile slab layout.
__global__ void reduce_slabs_kernel(const float* __restrict__ part, int n_slabs, int slab, const int* __restrict__ map, int P,
                                    float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int o = map[p];
  float s = 0.f;
  for (int t = 0; t < n_slabs; ++t) s += part[(size_t)t * slab + o];
  out[p] = s;
}

// pack[0..P) = grad (unnormalised here: normalisation is folded into the loss cotangent), pack[P..P+6) = raw squared-error
// sums, pack[P+6] = column count, pack[P+7] = 0, pack[P+8..P+13) = gradient sums wrt the five mPP parameters, rest 0.
// `stride` floats per partial row (LP from the adjoint kernel, 8 from the loss-only path).
__global__ void pack_loss_kernel(const float* __restrict__ lpart, int n_slabs, int stride, float ncol, float* __restrict__ pack_tail) {
  const int q = threadIdx.x;
  if (q < 6 || (q >= 8 && q < 13 && stride >= 16)) {
    float s = 0.f;
    for (int t = 0; t < n_slabs; ++t) s += lpart[(size_t)t * stride + q];
    pack_tail[q] = s;
  } else if (q == 6) {
    pack_tail[6] = ncol;
  } else if (q < 16) {
    pack_tail[q] = 0.f;
  }
}

// loss_out[0..6) = w_q * sum_q * inv_norm_q ; loss_out[6] = total
__global__ void finalize_loss_kernel(const float* __restrict__ pack_tail, const W6 w6, float inv_prof,
                                     float inv_grad, float* __restrict__ loss_out, unsigned int* __restrict__ nonfinite) {
  if (threadIdx.x == 0) {
    float tot = 0.f;
    const float inv_ncol = 1.f / pack_tail[6];
    for (int q = 0; q < 6; ++q) {
      const float v = w6.w[q] * pack_tail[q] * (q < 3 ? inv_prof : inv_grad) * inv_ncol;
      loss_out[q] = v;
      tot += v;
    }
    loss_out[6] = tot;
    if (tot - tot != 0.f) { atomicAdd(nonfinite, 1u); atomicAdd(nonfinite + 1, 1u); }  // NaN / Inf loss: reported as CPZ_ERR_NONFINITE by the host flavours
  }
}

// Loss-only path: six squared-error sums of a device trajectory against targets ([ncol][n_saved][S] both).
// One block per column; partial sums to lpart[col][8].
__global__ void loss_traj_kernel(const float* __restrict__ traj, const float* __restrict__ tgt, int n_saved, int S, int Nz,
                                 int nf, float Nf, float* __restrict__ lpart) {
  __shared__ float red[6][8];
  const size_t base = (size_t)blockIdx.x * n_saved * S;
  float ls[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < n_saved * S; i += blockDim.x) {
    const int s = i % S, q = s / Nz, k = s - q * Nz;
    const int wq = nf == 1 ? 2 : q;
    const float d = traj[base + i] - tgt[base + i];
    ls[wq] = fmaf(d, d, ls[wq]);
    if (nf == 3 && k >= 1) {
      const float g = Nf * (d - (traj[base + i - 1] - tgt[base + i - 1]));
      ls[3 + q] = fmaf(g, g, ls[3 + q]);
    }
  }
  for (int q = 0; q < 6; ++q) {
    float v = ls[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    lpart[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
  }
}

// non-finite values in a result (IEEE semantics are kept; this only reports: CPZ_ERR_NONFINITE)
__global__ void check_finite_kernel(const float* __restrict__ p, size_t stride, int ncol, int S, unsigned int* __restrict__ counter) {
  unsigned int bad = 0;
  for (int col = blockIdx.x; col < ncol; col += gridDim.x)
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      const float v = p[(size_t)col * stride + i];
      bad += (v - v != 0.f) ? 1u : 0u;  // NaN or +-Inf
    }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) { atomicAdd(counter, bad); atomicAdd(counter + 1, bad); }
}

// grad[p] *= 1/ncol_global, for the theta gradient [0, P) and the mPP-parameter gradient at [P+8, P+13)
__global__ void scale_kernel(float* __restrict__ g, int P, const float* __restrict__ pack_tail) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P || (p >= P + 8 && p < P + 13)) g[p] *= 1.f / pack_tail[6];
}

// Flux 0.11 ADAM: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; theta -= lr * m/(1-bp1) / (sqrt(v/(1-bp2)) + eps)
__global__ void adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                            int P, float lr, float b1, float b2, float eps, float bp1, float bp2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float gp = g[p];
  const float mp = b1 * m[p] + (1.f - b1) * gp;
  const float vp = b2 * v[p] + (1.f - b2) * gp * gp;
  m[p] = mp;
  v[p] = vp;
  theta[p] -= mp / (1.f - bp1) / (sqrtf(vp / (1.f - bp2)) + eps) * lr;
}


static int ensure_counters(cpz_ctx* c);

int launch_adjoint(cpz_model* m, const AdjArgs& a, int grid) {
  if (m->bwd.M.w_in_smem) return launch_adjoint_t<32, 256, true>(m, m->bwd, a, grid);
  return launch_adjoint_t<32, 256, false>(m, m->bwd, a, grid);
}

#define CPZ_LAUNCHED(m)            \
  CPZ_CUDA(cudaGetLastError());    \
  (m)->ctx->launches++;            \
  return CPZ_OK

int launch_reduce_slabs(cpz_model* m, const float* part, int n_slabs, int P, float* out) {
  reduce_slabs_kernel<<<(P + 255) / 256, 256, 0, m->ctx->stream>>>(part, n_slabs, m->bwd.M.slab, m->d_gmap, P, out);
  CPZ_LAUNCHED(m);
}
int launch_pack_loss(cpz_model* m, const float* lpart, int n_slabs, int stride, float ncol, float* pack_tail) {
  pack_loss_kernel<<<1, 32, 0, m->ctx->stream>>>(lpart, n_slabs, stride, ncol, pack_tail);
  CPZ_LAUNCHED(m);
}
int launch_finalize_loss(cpz_model* m, const float* pack_tail, const float* w6, float inv_prof, float inv_grad, float* loss_out) {
  W6 w;
  for (int q = 0; q < 6; ++q) w.w[q] = w6[q];
  { int rc = ensure_counters(m->ctx); if (rc) return rc; }
  finalize_loss_kernel<<<1, 32, 0, m->ctx->stream>>>(pack_tail, w, inv_prof, inv_grad, loss_out, m->ctx->d_nonfinite);
  CPZ_LAUNCHED(m);
}
int launch_loss_traj(cpz_model* m, const float* traj, const float* tgt, int ncol, int n_saved, int S, int Nz, int nf, float* lpart) {
  loss_traj_kernel<<<(unsigned)ncol, 256, 0, m->ctx->stream>>>(traj, tgt, n_saved, S, Nz, nf, (float)Nz, lpart);
  CPZ_LAUNCHED(m);
}
static int ensure_counters(cpz_ctx* c) {
  if (!c->d_nonfinite) {
    CPZ_CUDA(cudaMalloc(&c->d_nonfinite, 2 * sizeof(unsigned int)));
    CPZ_CUDA(cudaMemsetAsync(c->d_nonfinite, 0, 2 * sizeof(unsigned int), c->stream));
  }
  return CPZ_OK;
}
int begin_host_call(cpz_ctx* c) {
  int rc = ensure_counters(c);
  if (rc) return rc;
  CPZ_CUDA(cudaMemsetAsync(c->d_nonfinite + 1, 0, sizeof(unsigned int), c->stream));
  return CPZ_OK;
}
int launch_check_finite(cpz_ctx* c, const float* p, size_t stride, int ncol, int S) {
  int rc = ensure_counters(c);
  if (rc) return rc;
  check_finite_kernel<<<std::min(ncol, 592), 128, 0, c->stream>>>(p, stride, ncol, S, c->d_nonfinite);
  CPZ_CUDA(cudaGetLastError());
  c->launches++;
  return CPZ_OK;
}
int report_nonfinite(cpz_ctx* c, const char* what) {
  if (!c->d_nonfinite) return CPZ_OK;
  unsigned int fresh = 0;
  CPZ_CUDA(cudaMemcpyAsync(&fresh, c->d_nonfinite + 1, sizeof(fresh), cudaMemcpyDeviceToHost, c->stream));
  CPZ_CUDA(cudaStreamSynchronize(c->stream));
  if (fresh) return fail(CPZ_ERR_NONFINITE, "%u non-finite value(s) in %s (results are returned as computed)", fresh, what);
  return CPZ_OK;
}
int launch_scale(cpz_model* m, float* g, int P, const float* pack_tail) {
  scale_kernel<<<(P + 16 + 255) / 256, 256, 0, m->ctx->stream>>>(g, P, pack_tail);
  CPZ_LAUNCHED(m);
}
int launch_adam(cpz_model* m, const float* g, float lr, float b1, float b2, float eps) {
  const int P = (int)m->P;
  adam_kernel<<<(P + 255) / 256, 256, 0, m->ctx->stream>>>(m->d_theta, m->d_m, m->d_v, g, P, lr, b1, b2, eps, m->beta_pow[0], m->beta_pow[1]);
  CPZ_LAUNCHED(m);
}

}  // namespace cpz

// Device-side model layout shared by the host planner (cpz_plan.cpp) and the kernels.
// Everything here is POD and is passed to kernels by value (kernel parameter space = constant bank).
#pragma once
#include <stdint.h>

#define CPZ_MAX_GEMM 18   // CPZ_MAX_NETS * CPZ_MAX_LAYERS
#define CPZ_MAX_PHASE 18
#define CPZ_MAX_STAGES 6

// One Dense layer of one net: out[N][CT] = act(W^T in + b), W stored [K][N] (== Flux's column-major out x in).
struct GemmD {
  int K, N;          // fan-in, fan-out
  int Npad;          // row stride of the shared-memory copy of W (multiple of TO, zero padded)
  int act;           // cpz_activation
  int w_off, b_off;  // float offsets into theta (destructure order, global memory)
  int sw_off, sb_off;// float offsets into the shared-memory weight arena (valid when w_in_smem)
  int in_off;        // float offset of the input activations in the activation arena, -1 = the state tile X
  int out_off;       // float offset of the output activations in the activation arena
  int tile_begin;    // first flattened tile index of this gemm inside its phase
  int n_og;          // output groups = ceil(N/TO)
  int TO;            // outputs per thread tile
  int net, layer;
  int gw_off, gb_off;// adjoint: offsets of this gemm's weight-gradient tiles / bias gradients in a CTA's gradient slab
};

// Gemms [g0,g1) run concurrently between two block barriers over a flattened tile space.
struct PhaseD {
  int g0, g1;
  int n_tiles;
  int TC, TO;  // thread tile: TC columns x TO outputs, uniform over the phase's gemms
  int ksplit;  // KQ: lanes of one warp that split the k range of a tile (1, 2 or 4)
};

// Pre-folded RHS constants (computed in double on the host, rounded once to float).
struct RhsC {
  float A[3];     // tau/H * sigma_flux/sigma_q            (NDE_training.jl:160-162)
  float c[3];     // sigma_q/sigma_flux/H (c[2] without 1/Pr)  (NDE_training.jl:130-132)
  float z0[3];    // s_flux(0) = -mu/sigma                  (NDE_training.jl:130)
  float cor_u_s, cor_u_m, cor_v_s, cor_v_m;  // Coriolis: du += cor_u_s*v + cor_u_m ; dv -= cor_v_s*u + cor_v_m
  float BzC;      // H*g*alpha*sigma_T                      (NDE_training.jl:49)
  float sig_u, sig_v;
  float nu0, nu_m, Ric, inv_dRi, inv_Pr, kappa, eps;
  float fc_iS2;   // T-only model with the mPP base: 1 / ((sigma_u eps)^2 + (sigma_v eps)^2), the shear term at u = v = 0
  float K_ca;     // T-only convective adjustment K
  float Nf;       // float(Nz): 1/Delta with Delta = 1/Nz
  float di_w;     // 2*pi*tau/period
  float di_amp;   // 1/(alpha*g)
  float mu_wT, inv_sig_wT;
};

struct ModelD {
  int Nz, nf, S;       // levels, fields, S = nf*Nz
  int variant, flags;
  int n_nets, n_gemm, n_phase;
  int nbc;             // 6 or 2
  int P;               // total parameters
  int slab;            // adjoint: floats per CTA gradient slab (4x4 weight-gradient tiles, 16 contiguous floats each, then biases)
  int w_in_smem;       // 1: all weights resident in shared memory
  int smem_w_floats;   // size of the shared weight arena
  int arena_floats;    // per-column floats of the activation arena (rows; multiply by CT)
  int nn_off[3];       // arena row offsets of the final layer outputs per net (-1: no net)
  int flux_off;        // arena row offset of the face-flux scratch (3*(Nz+1) rows), may alias hidden activations
  GemmD gemm[CPZ_MAX_GEMM];
  PhaseD phase[CPZ_MAX_PHASE];
  RhsC rc;
};

// Butcher tableau in double on the host, float on the device.
struct TableauD {
  int n_stages;
  float a[CPZ_MAX_STAGES][CPZ_MAX_STAGES];
  float b[CPZ_MAX_STAGES];
  float c[CPZ_MAX_STAGES];
};

struct TimeD {
  float dt, t0;
  int n_steps, n_substeps, save_stride, ckpt_stride;
  int step0;  // index of this launch's first step inside the whole solve (chunked host solves): t = t0 + (step0 + n) dt
};

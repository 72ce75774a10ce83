"""Batched (torch, CPU) restatement of the NDE column path: RHS in stencil form, fixed-step integrators,
loss, exact discrete-adjoint gradient (autograd through the unrolled solve), Flux ADAM, gyre closure step.

ORACLE = TEST INFRASTRUCTURE (see oracle/__init__.py). Parity unpinned: nothing in the reference's tests fixes
these numbers; `literal.py` is the independent restatement this file is checked against (tests/test_oracle_*.py).

Reference lines followed (relative to the reference checkout):
    RHS train     wind_mixing/src/NDE_training.jl:83-165 (+ :56-81 BC unpacking / diurnal top flux)
    RHS infer     wind_mixing/src/training_postprocessing.jl:55-153
    NN-free DE    wind_mixing/src/diffusivity_parameter_optimisation.jl:1-33  (= train RHS with no nets)
    T-only RHS    free_convection/src/free_convection_nde.jl:29-38, convective_adjustment_nde.jl:33-48
    loss          wind_mixing/src/loss.jl:1-42, NDE_training.jl:290-323; free_convection/src/training.jl:55-62
    closure step  free_convection/double_gyre_nn.jl:27-62,110-111,149-168; oceananigans_nn.jl:13-40
    Tsit5 tableau OrdinaryDiffEq 5.55.1 (third-party, wind_mixing/Manifest.toml:1340), Tsitouras 2011; fixed-step
                  precedent free_convection/convective_adjustment.jl:137
    ADAM          Flux 0.11.6 (third-party, wind_mixing/Manifest.toml:523); call site wind_mixing/train_NDE.jl:141

Layout: x [ncol, nf*Nz] (bottom level first), bcs [ncol, 6] = (uw_b, uw_t, vw_b, vw_t, wT_b, wT_t) or [ncol, 2].
"""
import math

import numpy as np
import torch

from .flux_nn import chain_torch

FLAG_MPP, FLAG_CA, FLAG_ZERO_WEIGHTS, FLAG_SMOOTH_NN, FLAG_SMOOTH_RI = 1, 2, 4, 8, 16
FLAG_DIURNAL, FLAG_CA_LITERAL_U, FLAG_DIURNAL_UNSHIFTED = 32, 64, 128
FLAG_IMPLICIT = 256  # diffusive / convective-adjustment fluxes are left out of the RHS and applied by implicit_diffusion()

# ---- tableaus ---------------------------------------------------------------------------------------------------
TSIT5_C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0)
TSIT5_A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
)
TSIT5_B = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774)
# 7th row = b with b7 = 0 (FSAL): the first-same-as-last evaluation is an implementation detail, not part of the map.

TABLEAUS = {
    "euler": (((),), (1.0,), (0.0,)),
    "rk4": (((), (0.5,), (0.0, 0.5), (0.0, 0.0, 1.0)), (1 / 6, 1 / 3, 1 / 3, 1 / 6), (0.0, 0.5, 0.5, 1.0)),
    "tsit5": (TSIT5_A, TSIT5_B, TSIT5_C),
}


def _c32(v):
    return float(np.float32(v))


def _filter3(v):
    """filters.* * v for the width-3 running mean of filtering_operators.jl:1-15, along the last dim."""
    out = torch.empty_like(v)
    third = _c32(1.0 / 3.0)  # the reference builds the filter as a Float32 matrix (filtering_operators.jl:3)
    out[..., 1:-1] = (v[..., :-2] + v[..., 1:-1] + v[..., 2:]) * third
    out[..., 0] = (v[..., 0] + v[..., 1]) * 0.5
    out[..., -1] = (v[..., -2] + v[..., -1]) * 0.5
    return out


def _delta_f(q, N):
    """D_face * q with Δ = 1/N: faces 1 and N+1 (Julia) are zero.  [ncol, N] -> [ncol, N+1]"""
    z = torch.zeros_like(q[:, :1])
    return torch.cat([z, N * (q[:, 1:] - q[:, :-1]), z], dim=1)


def _delta_c(F, N):
    """D_cell * F with Δ = 1/N.  [ncol, N+1] -> [ncol, N]"""
    return N * (F[:, 1:] - F[:, :-1])


def split_thetas(desc, theta):
    out, off = [], 0
    for n in desc.nets:
        out.append(theta[off:off + n.n_params])
        off += n.n_params
    assert off == theta.numel()
    return out


def diurnal_top(desc, Q, t, dtype):
    """s_wT(Q sin(2π tτ/period)/(αg)) per column (NDE_training.jl:72-73, data_containers.jl:135)."""
    tau, alpha, g, period = _c32(desc.tau), _c32(desc.alpha), _c32(desc.g), _c32(desc.diurnal_period)
    flux = Q.to(dtype) * math.sin(2 * math.pi / period * (float(t) * tau)) / (alpha * g)
    return (flux - _c32(desc.mu[5])) / _c32(desc.sigma[5])


def diffusivities(desc, x, p_mpp=None):
    """Face diffusivities D_q [ncol, N-1] on the interior faces (None where a field has none) such that the diffusive part of
    the flux is -D_q G_q, G_q = N (q_k - q_{k-1}):  u, v: c_q nu;  T: c_T nu_T (mPP, with the inference RHS's convective-
    adjustment rule), c_T kappa [G_T < 0] (CA-only training branch);  T-only: [mPP] c_T nu/Pr + [CA] K [G < 0].
    Same statements as in the reference RHS functions cited at the top of this file."""
    N = desc.Nz
    dtype = x.dtype
    H = _c32(desc.H)
    sg = [_c32(v) for v in desc.sigma]
    c = [sg[0] / sg[3] / H, sg[1] / sg[4] / H, sg[2] / sg[5] / H]
    if p_mpp is not None:
        nu0_, num_, dRi_, Ric_, Pr = p_mpp[0], p_mpp[1], p_mpp[2], p_mpp[3], p_mpp[4]
    else:
        nu0_, num_, dRi_, Ric_, Pr = _c32(desc.nu0), _c32(desc.nu_m), _c32(desc.dRi), _c32(desc.Ric), _c32(desc.Pr)
    BzC = H * _c32(desc.g) * _c32(desc.alpha) * sg[2]
    if desc.variant == 2:  # T-only
        G = N * (x[:, 1:] - x[:, :-1])
        D = None
        if desc.flags & FLAG_MPP:
            # BASELINE config 1 ("convective adjustment + mPP base" on the T-only NDE): the mPP
            # rule of NDE_training.jl:114-139 at u = v = 0, i.e. both shear gradients are D_face*0 + eps
            eps = _c32(desc.eps)
            Ri = BzC * (G + eps) / ((sg[0] * eps) ** 2 + (sg[1] * eps) ** 2)
            nu = nu0_ + num_ * (1 - torch.tanh((Ri - Ric_) / dRi_)) / 2
            D = c[2] * nu / Pr
        if desc.flags & FLAG_CA:
            Dca = torch.where(G < 0, torch.full_like(G, _c32(desc.K_ca)), torch.zeros_like(G))
            D = Dca if D is None else D + Dca
        return [D]
    u, v, T = x[:, :N], x[:, N:2 * N], x[:, 2 * N:]
    Gi = [N * (q[:, 1:] - q[:, :-1]) for q in (u, v, T)]
    mpp = bool(desc.flags & FLAG_MPP) or desc.variant == 1  # the inference RHS always applies mPP
    if mpp:
        eps = _c32(desc.eps) if desc.variant == 0 else 0.0
        if desc.variant == 0 and (desc.flags & FLAG_SMOOTH_RI):
            # filters.face acts on all N+1 faces, so the boundary-face Ri (gradients = 0 + eps) takes part
            Gf = [_delta_f(q, N) for q in (u, v, T)]
            Ri = BzC * (Gf[2] + eps) / ((sg[0] * (Gf[0] + eps)) ** 2 + (sg[1] * (Gf[1] + eps)) ** 2)
            Ri = _filter3(Ri)[:, 1:-1]
        else:
            Ri = BzC * (Gi[2] + eps) / ((sg[0] * (Gi[0] + eps)) ** 2 + (sg[1] * (Gi[1] + eps)) ** 2)
        nu = nu0_ + num_ * (1 - torch.tanh((Ri - Ric_) / dRi_)) / 2
        if desc.variant == 1 and (desc.flags & FLAG_CA):
            test = Gi[0] if (desc.flags & FLAG_CA_LITERAL_U) else Gi[2]
            nu_T = torch.where(test > 0, nu / Pr, torch.full_like(nu, _c32(desc.kappa)))
        else:
            nu_T = nu / Pr
        return [c[0] * nu, c[1] * nu, c[2] * nu_T]
    if desc.flags & FLAG_CA:
        return [None, None, c[2] * _c32(desc.kappa) * (Gi[2] < 0).to(dtype)]
    return [None, None, None]


def implicit_diffusion(desc, x, h, p_mpp=None):
    r"""Backward-Euler step of the vertical-diffusion part of the RHS over a non-dimensional interval h, with the
    diffusivities frozen at the incoming state — the treatment the reference uses when the NDE runs inside Oceananigans
    (modified_pacanowski_philander!, wind_mixing/src/NDE_oceananigans.jl:61-101; convective_adjustment!,
    free_convection/src/oceananigans_nn.jl:13-40):  lower_k = -r_k, upper_k = -r_{k+1}, diag_k = 1 + r_k + r_{k+1},
    r_k = h A_q N^2 D_q,k on the interior faces and r = 0 on the two boundary faces (their fluxes are prescribed and stay in
    the explicit part); x_q <- L_q \ x_q by the Thomas algorithm. Here in the NDE's scaled, non-dimensional variables."""
    N = desc.Nz
    tau, H = _c32(desc.tau), _c32(desc.H)
    sg = [_c32(v) for v in desc.sigma]
    Ds = diffusivities(desc, x, p_mpp)
    nf = desc.n_fields
    out = []
    for q in range(nf):
        xq = x[:, q * N:(q + 1) * N]
        if Ds[q] is None:
            out.append(xq)
            continue
        qq = q if nf == 3 else 2
        A = tau / H * sg[3 + qq] / sg[qq]
        z = torch.zeros_like(xq[:, :1])
        r = torch.cat([z, (h * A * N * N) * Ds[q], z], dim=1)  # faces 0..N
        lower, upper, diag = -r[:, :-1], -r[:, 1:], 1 + r[:, :-1] + r[:, 1:]  # per level k: faces k (below) and k+1 (above)
        cp, dp = [None] * N, [None] * N
        cp[0] = upper[:, 0] / diag[:, 0]
        dp[0] = xq[:, 0] / diag[:, 0]
        for k in range(1, N):
            den = diag[:, k] - lower[:, k] * cp[k - 1]
            cp[k] = upper[:, k] / den
            dp[k] = (xq[:, k] - lower[:, k] * dp[k - 1]) / den
        y = [None] * N
        y[N - 1] = dp[N - 1]
        for k in range(N - 2, -1, -1):
            y[k] = dp[k] - cp[k] * y[k + 1]
        out.append(torch.stack(y, dim=1))
    return torch.cat(out, dim=1)


def rhs(desc, theta, x, bcs, t=0.0, Q=None, p_mpp=None):
    """dx/dt for a batch of columns. x [ncol, S] torch tensor; theta torch 1-D (may require grad).
    p_mpp: optional tensor (nu0, nu_m, dRi, Ric, Pr) overriding the description's constants — the differentiable
    parameter vector p of `DE` (diffusivity_parameter_optimisation.jl:2)."""
    N = desc.Nz
    dtype = x.dtype
    H, tau = _c32(desc.H), _c32(desc.tau)
    mu = [_c32(m) for m in desc.mu]
    sg = [_c32(s) for s in desc.sigma]
    thetas = split_thetas(desc, theta) if desc.nets else []

    implicit = bool(desc.flags & FLAG_IMPLICIT)
    if desc.variant == 2:  # T-only
        T = x
        ncol = T.shape[0]
        if desc.nets:
            nn = chain_torch(thetas[0], desc.nets[0].sizes, desc.nets[0].acts, T)
        else:
            nn = torch.zeros(ncol, N - 1, dtype=dtype)
        if not implicit:
            D = diffusivities(desc, x, p_mpp)[0]
            if D is not None:
                # D G = [mPP] sigma_T/(sigma_wT H) nu/Pr dT/dz + [CA] min(0, K dT/dz)   (convective_adjustment_nde.jl:44-47)
                nn = nn - D * (N * (T[:, 1:] - T[:, :-1]))
        F = torch.cat([bcs[:, 0:1], nn, bcs[:, 1:2]], dim=1)
        A = sg[5] / sg[2] * tau / H
        return -A * _delta_c(F, N)

    f = _c32(desc.f)
    u, v, T = x[:, :N], x[:, N:2 * N], x[:, 2 * N:]
    ncol = x.shape[0]
    if desc.nets:
        nn = [chain_torch(thetas[i], desc.nets[i].sizes, desc.nets[i].acts, x) for i in range(3)]
    else:
        nn = [torch.zeros(ncol, N - 1, dtype=dtype) for _ in range(3)]
    if desc.variant == 0 and (desc.flags & FLAG_SMOOTH_NN):
        nn = [_filter3(n) for n in nn]

    bc_b = [bcs[:, 0:1], bcs[:, 2:3], bcs[:, 4:5]]
    bc_t = [bcs[:, 1:2], bcs[:, 3:4], bcs[:, 5:6]]
    z0 = [-mu[3 + i] / sg[3 + i] for i in range(3)]  # s_q(0)
    unshifted_top_T = False
    if desc.flags & FLAG_DIURNAL:
        bc_t[2] = diurnal_top(desc, Q, t, dtype).reshape(-1, 1)
        unshifted_top_T = desc.variant == 1 and bool(desc.flags & FLAG_DIURNAL_UNSHIFTED)

    # interior-face gradients only (faces 2..N): boundary rows of D_face are zero and never reach the fluxes
    Gi = [N * (q[:, 1:] - q[:, :-1]) for q in (u, v, T)]
    c = [sg[0] / sg[3] / H, sg[1] / sg[4] / H, sg[2] / sg[5] / H]
    mpp = bool(desc.flags & FLAG_MPP) or desc.variant == 1  # the inference RHS always applies mPP
    shift = desc.variant == 1 or bool(desc.flags & FLAG_ZERO_WEIGHTS)

    D = [None, None, None] if implicit else diffusivities(desc, x, p_mpp)
    interior = [nn[i] if D[i] is None else nn[i] - D[i] * Gi[i] for i in range(3)]

    F = []
    for i in range(3):
        if mpp and shift:
            b = bc_b[i] - z0[i]
            tp = bc_t[i] if (i == 2 and unshifted_top_T) else bc_t[i] - z0[i]
        elif desc.flags & FLAG_ZERO_WEIGHTS:  # zero_weights without mPP: literal zeros (NDE_training.jl:104-107)
            b = torch.zeros(ncol, 1, dtype=dtype)
            tp = torch.zeros(ncol, 1, dtype=dtype)
        else:
            b, tp = bc_b[i], bc_t[i]
        F.append(torch.cat([b.expand(ncol, 1), interior[i], tp.expand(ncol, 1)], dim=1))

    A = [tau / H * sg[3] / sg[0], tau / H * sg[4] / sg[1], tau / H * sg[5] / sg[2]]
    dudt = -A[0] * _delta_c(F[0], N) + f * tau / sg[0] * (sg[1] * v + mu[1])
    dvdt = -A[1] * _delta_c(F[1], N) - f * tau / sg[1] * (sg[0] * u + mu[0])
    dTdt = -A[2] * _delta_c(F[2], N)
    return torch.cat([dudt, dvdt, dTdt], dim=1)


def rk_step(desc, theta, x, bcs, t, h, Q=None, mpp=None):
    a, b, c = TABLEAUS[desc.integrator]
    ks = []
    for i in range(len(b)):
        xi = x
        for j, aij in enumerate(a[i]):
            if aij != 0.0:
                xi = xi + (h * aij) * ks[j]
        ks.append(rhs(desc, theta, xi, bcs, t + c[i] * h, Q, mpp))
    xn = x
    for i, bi in enumerate(b):
        xn = xn + (h * bi) * ks[i]
    return xn


def solve(desc, theta, x0, bcs, Q=None, mpp=None):
    """Fixed-step solve; returns traj [ncol, n_saved, S]. Frame 0 is the initial condition (save_stride > 0)."""
    x = x0
    frames = [x0] if desc.save_stride > 0 else []
    h = _c32(desc.dt) / desc.n_substeps
    for n in range(desc.n_steps):
        t_n = _c32(desc.t0) + n * _c32(desc.dt)
        for s in range(desc.n_substeps):
            if desc.flags & FLAG_IMPLICIT:  # Lie splitting in the order of the Oceananigans embedding: implicit diffusion with
                x = implicit_diffusion(desc, x, h, mpp)  # the diffusivities of the incoming state, then the explicit step
            x = rk_step(desc, theta, x, bcs, t_n + s * h, h, Q, mpp)
        if desc.save_stride > 0 and (n + 1) % desc.save_stride == 0:
            frames.append(x)
    if desc.save_stride <= 0:
        frames = [x]
    return torch.stack(frames, dim=1)


def loss_components(desc, traj, targets):
    """Unweighted (u, v, T, ∂u∂z, ∂v∂z, ∂T∂z) losses: per-column Flux.mse then mean over columns
    (NDE_training.jl:303-318; loss.jl:1-9). ∂_∂z keeps the two zero boundary rows in the mean."""
    N = desc.Nz
    ncol, nt, S = traj.shape
    out = []
    zero = torch.zeros((), dtype=traj.dtype)
    if desc.n_fields == 1:
        L_T = ((traj - targets) ** 2).mean()
        return [zero, zero, L_T, zero, zero, zero]
    for i in range(3):
        d = traj[:, :, i * N:(i + 1) * N] - targets[:, :, i * N:(i + 1) * N]
        out.append((d ** 2).mean())
    for i in range(3):
        d = traj[:, :, i * N:(i + 1) * N] - targets[:, :, i * N:(i + 1) * N]
        g = _delta_f(d.reshape(ncol * nt, N), N)
        out.append((g ** 2).mean())
    return out


def loss_total(desc, theta, x0, bcs, Q, targets, w, mpp=None):
    traj = solve(desc, theta, x0, bcs, Q, mpp)
    comps = loss_components(desc, traj, targets)
    scaled = [float(w[i]) * comps[i] for i in range(6)]
    return sum(scaled), scaled


def loss_grad(desc, theta, x0, bcs, Q, targets, w):
    """(total, [6 weighted components], grad wrt theta) — exact discrete adjoint via autograd."""
    th = theta.clone().detach().requires_grad_(True)
    total, scaled = loss_total(desc, th, x0, bcs, Q, targets, w)
    (g,) = torch.autograd.grad(total, th)
    return total.detach(), [s.detach() for s in scaled], g


def loss_grad_mpp(desc, theta, x0, bcs, Q, targets, w):
    """(total, grad wrt p = (nu0, nu_m, dRi, Ric, Pr)) — loss_mpp / loss_gradient_mpp of
    diffusivity_parameter_optimisation.jl:150-197 in UNSCALED parameters, exact discrete adjoint via autograd."""
    p = torch.tensor([_c32(desc.nu0), _c32(desc.nu_m), _c32(desc.dRi), _c32(desc.Ric), _c32(desc.Pr)], dtype=x0.dtype,
                     requires_grad=True)
    total, scaled = loss_total(desc, theta, x0, bcs, Q, targets, w, mpp=p)
    (g,) = torch.autograd.grad(total, p)
    return total.detach(), g


def adam_step(theta, g, m, v, beta_pow, lr, b1=0.9, b2=0.999, eps=1e-8):
    """Flux 0.11 ADAM: mt = β1 mt + (1-β1)Δ; vt = β2 vt + (1-β2)Δ²; Δ = mt/(1-βp1)/(√(vt/(1-βp2))+ε)·η; βp *= β.
    beta_pow starts at (β1, β2). Returns new (theta, m, v, beta_pow)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    delta = m / (1 - beta_pow[0]) / (torch.sqrt(v / (1 - beta_pow[1])) + eps) * lr
    return theta - delta, m, v, (beta_pow[0] * b1, beta_pow[1] * b2)


def closure_step(desc, theta, cdesc, T, y):
    """One host-model step of the gyre closure on a field T [Nz, Ny, Nx] (double_gyre_nn.jl:211-217 order:
    convective_adjustment! first, then compute_neural_network_forcing! on the adjusted T).
    Returns (forcing = ∂z wT_NN [Nz,Ny,Nx], T_adjusted [Nz,Ny,Nx])."""
    Nz, Ny, Nx = T.shape
    dtype = T.dtype
    cols = T.reshape(Nz, Ny * Nx).t().contiguous()  # [ncol, Nz]
    dz, dt, K = _c32(cdesc.dz), _c32(cdesc.dt), _c32(cdesc.K)
    # implicit convective adjustment (oceananigans_nn.jl:13-40), Thomas algorithm batched over columns
    z = torch.zeros_like(cols[:, :1])
    G = torch.cat([z, (cols[:, 1:] - cols[:, :-1]) / dz, z], dim=1)
    gc = 0.5 * (G[:, :-1] + G[:, 1:])
    kap = torch.where(gc < 0, torch.full_like(gc, K), torch.zeros_like(gc))
    r = dt / dz ** 2
    lower = torch.cat([z, -r * kap[:, 1:]], dim=1)          # ld[k] multiplies T[k-1], k>=1
    upper = torch.cat([-r * kap[:, 1:], z], dim=1)          # ud[k] multiplies T[k+1], k<=N-2
    diag = 1 + r * torch.cat([kap[:, :-1] + kap[:, 1:], kap[:, -1:]], dim=1)
    cp = torch.zeros_like(cols)
    dp = torch.zeros_like(cols)
    cp[:, 0] = upper[:, 0] / diag[:, 0]
    dp[:, 0] = cols[:, 0] / diag[:, 0]
    for k in range(1, Nz):
        den = diag[:, k] - lower[:, k] * cp[:, k - 1]
        cp[:, k] = upper[:, k] / den
        dp[:, k] = (cols[:, k] - lower[:, k] * dp[:, k - 1]) / den
    Tn = torch.zeros_like(cols)
    Tn[:, -1] = dp[:, -1]
    for k in range(Nz - 2, -1, -1):
        Tn[:, k] = dp[:, k] - cp[:, k] * Tn[:, k + 1]
    # NN forcing (double_gyre_nn.jl:149-168)
    yy = y.to(dtype).reshape(Ny, 1).expand(Ny, Nx).reshape(-1)
    T_ref = _c32(cdesc.T_mid) + _c32(cdesc.dT) / _c32(cdesc.Ly) * yy
    surface_flux = -_c32(cdesc.mu_relax) * (Tn[:, -1] - T_ref)
    prof = _c32(cdesc.T_shift) + Tn / _c32(cdesc.T_div)
    xin = (prof - _c32(desc.mu[2])) / _c32(desc.sigma[2])
    nn = chain_torch(theta, desc.nets[0].sizes, desc.nets[0].acts, xin)
    nn = _c32(desc.sigma[5]) * nn + _c32(desc.mu[5])
    wT = torch.cat([z, nn, surface_flux.reshape(-1, 1)], dim=1)
    forcing = (wT[:, 1:] - wT[:, :-1]) / dz
    return forcing.t().reshape(Nz, Ny, Nx).contiguous(), Tn.t().reshape(Nz, Ny, Nx).contiguous()


def closure_step_uvt(desc, theta, cdesc, u, v, T):
    """One host-model step of the u/v/T NDE embedded in Oceananigans (progress_neural_network,
    wind_mixing/src/NDE_oceananigans.jl:380-405): the three NN forcing chains on the incoming state (:288-344), then the
    backward-Euler modified Pacanowski–Philander step (:17-101). Fields [Nz, Ny, Nx], dimensional. Returns
    (dz_flux [3, Nz, Ny, Nx] = dz_uw_NN, dz_vw_NN, dz_wT_NN;  state [3, Nz, Ny, Nx] = u', v', T').
    Boundary-face nu_T and the literal momentum shift: see oracle/literal.py."""
    Nz, Ny, Nx = T.shape
    dtype = T.dtype
    cols = [f.reshape(Nz, Ny * Nx).t().contiguous() for f in (u, v, T)]  # [ncol, Nz]
    mu = [_c32(m) for m in desc.mu]
    sg = [_c32(s) for s in desc.sigma]
    dz, dt = _c32(cdesc.dz), _c32(cdesc.dt)
    z = torch.zeros_like(cols[0][:, :1])
    # NN forcing chains
    x = torch.cat([(cols[q] - mu[q]) / sg[q] for q in range(3)], dim=1)
    thetas = split_thetas(desc, theta)
    tops = [_c32(cdesc.uw_top), _c32(cdesc.vw_top), _c32(cdesc.wT_top)]
    dzf = []
    for q in range(3):
        nn = chain_torch(thetas[q], desc.nets[q].sizes, desc.nets[q].acts, x)
        un = sg[3 + q] * nn + mu[3 + q]
        shift = (sg[3 + q] * un[:, :1] + mu[3 + q]) if q < 2 else un[:, :1]
        F = torch.cat([z, un - shift, torch.full_like(z, tops[q])], dim=1)
        dzf.append((F[:, 1:] - F[:, :-1]) / dz)
    # implicit mPP step
    G = [(c[:, 1:] - c[:, :-1]) / dz for c in cols]  # interior faces
    Ri = _c32(desc.g) * _c32(desc.alpha) * G[2] / (G[0] ** 2 + G[1] ** 2)
    nu = _c32(desc.nu0) + _c32(desc.nu_m) * (1 - torch.tanh((Ri - _c32(desc.Ric)) / _c32(desc.dRi))) / 2
    nu_T = nu / _c32(desc.Pr)
    if cdesc.convective_adjustment:
        nu_T = torch.where(Ri > 0, nu_T, torch.full_like(nu_T, _c32(cdesc.kappa_ca)))
    r = dt / dz ** 2
    out = []
    for q in range(3):
        nq = torch.cat([z, nu if q < 2 else nu_T, z], dim=1)  # faces 0..Nz (Julia 1..Nz+1); only faces 0..Nz-1 enter the matrix
        lower = -r * nq[:, :-1]                                # row k: multiplies x[k-1] (k >= 1)
        upper = -r * nq[:, 1:]                                 # row k: multiplies x[k+1] (k <= Nz-2)
        diag = 1 + r * (nq[:, :-1] + nq[:, 1:])                # the last row's nq[Nz] is 0: 1 + r nu[Nz-1], as :85-86
        cq = cols[q]
        cp, dp = [None] * Nz, [None] * Nz
        cp[0] = upper[:, 0] / diag[:, 0]
        dp[0] = cq[:, 0] / diag[:, 0]
        for k in range(1, Nz):
            den = diag[:, k] - lower[:, k] * cp[k - 1]
            cp[k] = upper[:, k] / den
            dp[k] = (cq[:, k] - lower[:, k] * dp[k - 1]) / den
        y = [None] * Nz
        y[Nz - 1] = dp[Nz - 1]
        for k in range(Nz - 2, -1, -1):
            y[k] = dp[k] - cp[k] * y[k + 1]
        if q == 2:
            y[0] = cq[:, 0]
        out.append(torch.stack(y, dim=1))
    to_field = lambda c: c.t().reshape(Nz, Ny, Nx).contiguous()
    return torch.stack([to_field(c) for c in dzf]), torch.stack([to_field(c) for c in out])

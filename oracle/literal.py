"""One-column, dense-matrix, line-by-line restatement of the reference RHS functions (numpy).

ORACLE = TEST INFRASTRUCTURE. This file mirrors the Julia statement by statement (dense Dᶜ/Dᶠ matrices,
concatenations, broadcasts) so that the vectorised stencil form in `nde.py` — which is what the CUDA kernels are
compared against at scale — can itself be checked against something that reads like the reference.

`desc` is any object with the attributes of `ModelDesc` (Nz, flags, constants, mu/sigma, nets ...).
All constants are rounded to float32 first (the engine sees float32 values), then promoted to `dtype`.
"""
import numpy as np

from .flux_nn import chain_numpy
from .operators import D_c, D_f, smoothing_filter

FLAG_MPP, FLAG_CA, FLAG_ZERO_WEIGHTS, FLAG_SMOOTH_NN, FLAG_SMOOTH_RI = 1, 2, 4, 8, 16
FLAG_DIURNAL, FLAG_CA_LITERAL_U, FLAG_DIURNAL_UNSHIFTED = 32, 64, 128


def _c(desc, name, dtype):
    return dtype(np.float32(getattr(desc, name)))


def _net_thetas(desc, theta):
    out, off = [], 0
    for n in desc.nets:
        out.append(theta[off:off + n.n_params])
        off += n.n_params
    assert off == len(theta)
    return out


def local_richardson(dudz, dvdz, dTdz, H, g, alpha, sig_u, sig_v, sig_T):
    """wind_mixing/src/NDE_training.jl:46-52"""
    Bz = H * g * alpha * sig_T * dTdz
    S2 = (sig_u * dudz) ** 2 + (sig_v * dvdz) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return Bz / S2


def tanh_step(x):
    """wind_mixing/src/NDE_training.jl:54"""
    return (1 - np.tanh(x)) / 2


def diurnal_wT_top(desc, Q, t, dtype):
    """scalings.wT(wT_top_function(t*tau));  NDE_training.jl:72-73, data_containers.jl:135"""
    tau, alpha, g = _c(desc, "tau", dtype), _c(desc, "alpha", dtype), _c(desc, "g", dtype)
    period = _c(desc, "diurnal_period", dtype)
    flux = dtype(Q) * np.sin(2 * np.pi / period * (dtype(t) * tau)) / (alpha * g)
    return (flux - dtype(np.float32(desc.mu[5]))) / dtype(np.float32(desc.sigma[5]))


def rhs_train(desc, theta, x, bcs, t=0.0, Q=None, dtype=np.float64):
    """predict_flux + predict_NDE: wind_mixing/src/NDE_training.jl:83-165 (and NDE :56-81 for the BC unpacking).
    With desc.nets == [] this is `DE` (diffusivity_parameter_optimisation.jl:1-33)."""
    Nz = desc.Nz
    H, tau, f = _c(desc, "H", dtype), _c(desc, "tau", dtype), _c(desc, "f", dtype)
    mu = [dtype(np.float32(m)) for m in desc.mu]
    sg = [dtype(np.float32(s)) for s in desc.sigma]
    mu_u, mu_v, sig_u, sig_v, sig_T = mu[0], mu[1], sg[0], sg[1], sg[2]
    sig_uw, sig_vw, sig_wT = sg[3], sg[4], sg[5]
    s0 = [(0 - mu[3 + i]) / sg[3 + i] for i in range(3)]  # scalings.uw(0f0) etc.
    D_cell = np.float32(D_c(Nz, 1 / Nz)).astype(dtype)  # NDE_training.jl:35
    D_face = np.float32(D_f(Nz, 1 / Nz)).astype(dtype)
    x = np.asarray(x, dtype=dtype)
    theta = np.asarray(theta, dtype=dtype)
    uw_b, uw_t, vw_b, vw_t, wT_b, wT_t = [dtype(b) for b in bcs]
    if desc.flags & FLAG_DIURNAL:
        wT_t = diurnal_wT_top(desc, Q, t, dtype)

    u, v, T = x[0:Nz], x[Nz:2 * Nz], x[2 * Nz:3 * Nz]

    if desc.nets:
        th = _net_thetas(desc, theta)
        uw_int, vw_int, wT_int = [chain_numpy(th[i], desc.nets[i].sizes, desc.nets[i].acts, x) for i in range(3)]
    else:
        uw_int = vw_int = wT_int = np.zeros(Nz - 1, dtype=dtype)

    if desc.flags & FLAG_SMOOTH_NN:  # :98-102
        F = smoothing_filter(Nz - 1, 3).astype(dtype)
        uw_int, vw_int, wT_int = F @ uw_int, F @ vw_int, F @ wT_int

    z = dtype(0)
    if desc.flags & FLAG_ZERO_WEIGHTS:  # :104-112
        uw = np.concatenate([[z], uw_int, [z]])
        vw = np.concatenate([[z], vw_int, [z]])
        wT = np.concatenate([[z], wT_int, [z]])
    else:
        uw = np.concatenate([[uw_b], uw_int, [uw_t]])
        vw = np.concatenate([[vw_b], vw_int, [vw_t]])
        wT = np.concatenate([[wT_b], wT_int, [wT_t]])

    if desc.flags & FLAG_MPP:  # :114-139
        eps = _c(desc, "eps", dtype)
        dudz, dvdz, dTdz = D_face @ u, D_face @ v, D_face @ T
        Ri = local_richardson(dudz + eps, dvdz + eps, dTdz + eps, H, _c(desc, "g", dtype), _c(desc, "alpha", dtype),
                              sig_u, sig_v, sig_T)
        if desc.flags & FLAG_SMOOTH_RI:
            Ri = smoothing_filter(Nz + 1, 3).astype(dtype) @ Ri
        nu = _c(desc, "nu0", dtype) + _c(desc, "nu_m", dtype) * tanh_step((Ri - _c(desc, "Ric", dtype)) / _c(desc, "dRi", dtype))
        Pr = _c(desc, "Pr", dtype)
        if desc.flags & FLAG_ZERO_WEIGHTS:
            nu_dudz = np.concatenate([[-(uw_b - s0[0])], sig_u / sig_uw / H * nu[1:-1] * dudz[1:-1], [-(uw_t - s0[0])]])
            nu_dvdz = np.concatenate([[-(vw_b - s0[1])], sig_v / sig_vw / H * nu[1:-1] * dvdz[1:-1], [-(vw_t - s0[1])]])
            nu_dTdz = np.concatenate([[-(wT_b - s0[2])], sig_T / sig_wT / H * nu[1:-1] / Pr * dTdz[1:-1], [-(wT_t - s0[2])]])
        else:
            nu_dudz = sig_u / sig_uw / H * nu * dudz
            nu_dvdz = sig_v / sig_vw / H * nu * dvdz
            nu_dTdz = sig_T / sig_wT / H * nu * dTdz / Pr
        uw, vw, wT = uw - nu_dudz, vw - nu_dvdz, wT - nu_dTdz
    elif desc.flags & FLAG_CA:  # :140-143 (reference reads an undefined κ here; desc.kappa is used)
        dTdz = D_face @ T
        wT = wT - sig_T / sig_wT / H * _c(desc, "kappa", dtype) * np.minimum(0, dTdz)

    # predict_NDE :160-164
    dudt = -tau / H * sig_uw / sig_u * (D_cell @ uw) + f * tau / sig_u * (sig_v * v + mu_v)
    dvdt = -tau / H * sig_vw / sig_v * (D_cell @ vw) - f * tau / sig_v * (sig_u * u + mu_u)
    dTdt = -tau / H * sig_wT / sig_T * (D_cell @ wT)
    return np.concatenate([dudt, dvdt, dTdt])


def rhs_infer(desc, theta, x, bcs, t=0.0, Q=None, dtype=np.float64):
    """NDE! / predict_flux! of solve_NDE_mutating: wind_mixing/src/training_postprocessing.jl:55-153."""
    Nz = desc.Nz
    H, tau, f = _c(desc, "H", dtype), _c(desc, "tau", dtype), _c(desc, "f", dtype)
    mu = [dtype(np.float32(m)) for m in desc.mu]
    sg = [dtype(np.float32(s)) for s in desc.sigma]
    mu_u, mu_v, sig_u, sig_v, sig_T = mu[0], mu[1], sg[0], sg[1], sg[2]
    sig_uw, sig_vw, sig_wT = sg[3], sg[4], sg[5]
    s0 = [(0 - mu[3 + i]) / sg[3 + i] for i in range(3)]
    nu0, nu_m, Ric, dRi, Pr = [_c(desc, k, dtype) for k in ("nu0", "nu_m", "Ric", "dRi", "Pr")]
    D_cell = np.float32(D_c(Nz, 1 / Nz)).astype(dtype)
    D_face = np.float32(D_f(Nz, 1 / Nz)).astype(dtype)
    x = np.asarray(x, dtype=dtype)
    theta = np.asarray(theta, dtype=dtype)
    uw_b, uw_t, vw_b, vw_t, wT_b, wT_t = [dtype(b) for b in bcs]

    uw = np.zeros(Nz + 1, dtype=dtype)
    vw = np.zeros(Nz + 1, dtype=dtype)
    wT = np.zeros(Nz + 1, dtype=dtype)
    uw[0], vw[0], wT[0] = uw_b - s0[0], vw_b - s0[1], wT_b - s0[2]  # :82-84
    uw[-1], vw[-1] = uw_t - s0[0], vw_t - s0[1]  # :86-87
    if desc.flags & FLAG_DIURNAL:  # :89-93 and :142-144
        top = diurnal_wT_top(desc, Q, t, dtype)
        # the code writes wT[end] = BCs.wT.top(t*τ) in every RHS call (unshifted, Q3); the shifted form is the default
        wT[-1] = top if (desc.flags & FLAG_DIURNAL_UNSHIFTED) else top - s0[2]
    else:
        wT[-1] = wT_t - s0[2]

    u, v, T = x[0:Nz], x[Nz:2 * Nz], x[2 * Nz:3 * Nz]
    if desc.nets:
        th = _net_thetas(desc, theta)
        uw[1:-1], vw[1:-1], wT[1:-1] = [chain_numpy(th[i], desc.nets[i].sizes, desc.nets[i].acts, x) for i in range(3)]  # :106-108

    dudz, dvdz, dTdz = D_face @ u, D_face @ v, D_face @ T  # :110-112
    Ri = local_richardson(dudz, dvdz, dTdz, H, _c(desc, "g", dtype), _c(desc, "alpha", dtype), sig_u, sig_v, sig_T)  # :114
    nu = nu0 + nu_m * tanh_step((Ri - Ric) / dRi)  # :116
    nu_T = np.zeros(Nz + 1, dtype=dtype)
    if desc.flags & FLAG_CA:  # :118-122
        kappa = _c(desc, "kappa", dtype)
        test = dudz if (desc.flags & FLAG_CA_LITERAL_U) else dTdz  # Q2: the code tests ∂u∂z; default here is ∂T∂z
        for i in range(1, Nz):
            nu_T[i] = nu[i] / Pr if test[i] > 0 else kappa
    else:
        nu_T[:] = nu / Pr  # :124
    uw[1:-1] -= sig_u / sig_uw / H * nu[1:-1] * dudz[1:-1]  # :126-128
    vw[1:-1] -= sig_v / sig_vw / H * nu[1:-1] * dvdz[1:-1]
    wT[1:-1] -= sig_T / sig_wT / H * nu_T[1:-1] * dTdz[1:-1]

    dudt = -tau / H * sig_uw / sig_u * (D_cell @ uw) + f * tau / sig_u * (sig_v * v + mu_v)  # :150-152
    dvdt = -tau / H * sig_vw / sig_v * (D_cell @ vw) - f * tau / sig_v * (sig_u * u + mu_u)
    dTdt = -tau / H * sig_wT / sig_T * (D_cell @ wT)
    return np.concatenate([dudt, dvdt, dTdt])


def rhs_free_convection(desc, theta, T, bcs, t=0.0, Q=None, dtype=np.float64):
    """∂T∂t of FreeConvectionNDE (free_convection/src/free_convection_nde.jl:29-38) and, with FLAG_CA,
    ConvectiveAdjustmentNDE (convective_adjustment_nde.jl:33-48).  p = [θ; bottom_flux, top_flux, σ_T, σ_wT, H, τ]."""
    Nz = desc.Nz
    H, tau = _c(desc, "H", dtype), _c(desc, "tau", dtype)
    sig_T, sig_wT = dtype(np.float32(desc.sigma[2])), dtype(np.float32(desc.sigma[5]))
    Dzc = D_c(Nz, 1 / Nz).astype(dtype)  # Δẑ = Δz/H = 1/Nz   free_convection_nde.jl:16-18
    Dzf = D_f(Nz, 1 / Nz).astype(dtype)
    T = np.asarray(T, dtype=dtype)
    bottom_flux, top_flux = dtype(bcs[0]), dtype(bcs[1])
    if desc.nets:
        wT_interior = chain_numpy(np.asarray(theta, dtype=dtype), desc.nets[0].sizes, desc.nets[0].acts, T)
    else:
        wT_interior = np.zeros(Nz - 1, dtype=dtype)
    wT = np.concatenate([[bottom_flux], wT_interior, [top_flux]])
    if desc.flags & FLAG_MPP:
        # "convective adjustment + mPP base" of BASELINE config 1: predict_flux's mPP block (NDE_training.jl:114-139,
        # not zero_weights) with u = v = 0, so that dudz = dvdz = D_face*0
        eps = _c(desc, "eps", dtype)
        sig_u, sig_v = dtype(np.float32(desc.sigma[0])), dtype(np.float32(desc.sigma[1]))
        zero = np.zeros(Nz + 1, dtype=dtype)
        dTdz = Dzf @ T
        Ri = local_richardson(zero + eps, zero + eps, dTdz + eps, H, _c(desc, "g", dtype), _c(desc, "alpha", dtype), sig_u, sig_v, sig_T)
        nu = _c(desc, "nu0", dtype) + _c(desc, "nu_m", dtype) * tanh_step((Ri - _c(desc, "Ric", dtype)) / _c(desc, "dRi", dtype))
        nu_dTdz = sig_T / sig_wT / H * nu * dTdz / _c(desc, "Pr", dtype)   # boundary rows of D_face are zero
        wT = wT - nu_dTdz
    if desc.flags & FLAG_CA:
        dz_wT = Dzc @ wT
        dTdz = Dzf @ T
        dz_KdTdz = Dzc @ np.minimum(0, _c(desc, "K_ca", dtype) * dTdz)
        return sig_wT / sig_T * tau / H * (-dz_wT + dz_KdTdz)
    dz_wT = Dzc @ (sig_wT / sig_T * tau / H * wT)
    return -dz_wT


def rhs(desc, theta, x, bcs, t=0.0, Q=None, dtype=np.float64):
    if desc.variant == 0:
        return rhs_train(desc, theta, x, bcs, t, Q, dtype)
    if desc.variant == 1:
        return rhs_infer(desc, theta, x, bcs, t, Q, dtype)
    return rhs_free_convection(desc, theta, x, bcs, t, Q, dtype)


# ---- loss (wind_mixing/src/loss.jl:1-42) ---------------------------------------------------------------------

def loss(a, b):
    """Flux.mse: mean of squared differences over all entries.  loss.jl:1-3"""
    return np.mean((np.asarray(a) - np.asarray(b)) ** 2)


def d_dz(profile, D_face):
    """∂_∂z: D_face applied column-wise to an (Nz x Nt) matrix.  loss.jl:9"""
    return np.stack([D_face @ profile[:, i] for i in range(profile.shape[1])], axis=1)


def calculate_loss_scalings(losses, fractions, train_gradient):
    """loss.jl:11-31.  losses/fractions are dicts keyed u,v,T,du,dv,dT / T,dT,profile."""
    velocity_scaling = (1 - fractions["T"]) / fractions["T"] * losses["T"] / (losses["u"] + losses["v"])
    profile_loss = velocity_scaling * (losses["u"] + losses["v"]) + losses["T"]
    if train_gradient:
        vgs = (1 - fractions["dT"]) / fractions["dT"] * losses["dT"] / (losses["du"] + losses["dv"])
        gradient_loss = vgs * (losses["du"] + losses["dv"]) + losses["dT"]
        tgs = (1 - fractions["profile"]) / fractions["profile"] * profile_loss / gradient_loss
    else:
        vgs = gradient_loss = tgs = 0
    return {"u": velocity_scaling, "v": velocity_scaling, "T": 1, "du": tgs * vgs, "dv": tgs * vgs, "dT": tgs}


def apply_loss_scalings(losses, scalings):
    """loss.jl:33-42"""
    return {k: scalings[k] * losses[k] for k in ("u", "v", "T", "du", "dv", "dT")}


# ---- implicit convective adjustment and gyre closure (free_convection) ------------------------------------------

def convective_adjustment_implicit(T, dt, dz, K):
    """convective_adjustment!: free_convection/src/oceananigans_nn.jl:13-40 (3-D twin double_gyre_nn.jl:27-62).
    The centre-located ∂T/∂z is defined as in the stand-alone script free_convection/convective_adjustment.jl:106-129:
    mean of the two adjacent face gradients with zero boundary-face gradients (the Oceananigans halo rule the
    embedded version relies on is third-party and not restated)."""
    T = np.asarray(T, dtype=np.float64)
    N = len(T)
    G = np.zeros(N + 1)
    G[1:N] = (T[1:] - T[:-1]) / dz
    dTdz_c = 0.5 * (G[:-1] + G[1:])
    kap = np.where(dTdz_c < 0, K, 0.0)
    r = dt / dz ** 2
    ld = np.array([-r * kap[k] for k in range(1, N)])
    ud = np.array([-r * kap[k + 1] for k in range(0, N - 1)])
    d = np.zeros(N)
    for k in range(N - 1):
        d[k] = 1 + r * (kap[k] + kap[k + 1])
    d[N - 1] = 1 + r * kap[N - 1]
    L = np.diag(d) + np.diag(ld, -1) + np.diag(ud, 1)
    return np.linalg.solve(L, T)


def gyre_closure_column(desc, theta, cdesc, T_col, y):
    """compute_neural_network_forcing! for one column: free_convection/double_gyre_nn.jl:110-111,149-168.
    Returns ∂z(wT) at the Nz centres (the Forcing kernel applies the minus sign, :135)."""
    T_col = np.asarray(T_col, dtype=np.float64)
    Nz = desc.Nz
    muT, sgT = np.float64(np.float32(desc.mu[2])), np.float64(np.float32(desc.sigma[2]))
    muw, sgw = np.float64(np.float32(desc.mu[5])), np.float64(np.float32(desc.sigma[5]))
    T_ref = np.float64(np.float32(cdesc.T_mid)) + np.float64(np.float32(cdesc.dT)) / np.float64(np.float32(cdesc.Ly)) * y
    surface_flux = -np.float64(np.float32(cdesc.mu_relax)) * (T_col[Nz - 1] - T_ref)
    T_profile = np.float64(np.float32(cdesc.T_shift)) + T_col / np.float64(np.float32(cdesc.T_div))
    wT_int = chain_numpy(np.asarray(theta, dtype=np.float64), desc.nets[0].sizes, desc.nets[0].acts, (T_profile - muT) / sgT)
    wT_int = sgw * wT_int + muw
    wT = np.concatenate([[0.0], wT_int, [surface_flux]])
    return (wT[1:] - wT[:-1]) / np.float64(np.float32(cdesc.dz))


# ---- u/v/T NDE embedded in Oceananigans (wind_mixing) -----------------------------------------------------------

def mpp_diffusivity_oceananigans(desc, cdesc, u, v, T):
    """modified_pacanowski_philander_diffusivity: wind_mixing/src/NDE_oceananigans.jl:17-58, one column, dimensional.
    Ri at the interior faces is Oceanostics' richardson_number_ccf! (dz b / (dz u^2 + dz v^2), b = g alpha T, third-party,
    restated). nu is filled on faces 2..Nz only (:45-47). With convective adjustment nu_T = Ri > 0 ? nu/Pr : kappa_ca
    (:50-53, kappa_ca hard-coded 1f0). On the two boundary faces the reference's Ri comes from Oceananigans halo values
    (third-party); nu is 0 there and nu_T is taken as 0 too (what a stably stratified gradient boundary condition gives:
    Ri = +Inf > 0 -> nu/Pr = 0). Returns (nu, nu_T) on the Nz+1 faces."""
    u, v, T = (np.asarray(a, dtype=np.float64) for a in (u, v, T))
    N = desc.Nz
    f64 = lambda x: np.float64(np.float32(x))
    g, alpha, dz = f64(desc.g), f64(desc.alpha), f64(cdesc.dz)
    nu0, nu_m, dRi, Ric, Pr = f64(desc.nu0), f64(desc.nu_m), f64(desc.dRi), f64(desc.Ric), f64(desc.Pr)
    nu = np.zeros(N + 1)
    nu_T = np.zeros(N + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(1, N):  # Julia faces 2..Nz
            dudz, dvdz, dTdz = (u[i] - u[i - 1]) / dz, (v[i] - v[i - 1]) / dz, (T[i] - T[i - 1]) / dz
            Ri = g * alpha * dTdz / (dudz ** 2 + dvdz ** 2)
            nu[i] = nu0 + nu_m * tanh_step((Ri - Ric) / dRi)
            if cdesc.convective_adjustment:
                nu_T[i] = nu[i] / Pr if Ri > 0 else f64(cdesc.kappa_ca)
            else:
                nu_T[i] = nu[i] / Pr
    return nu, nu_T


def modified_pacanowski_philander_step(desc, cdesc, u, v, T):
    """modified_pacanowski_philander!: wind_mixing/src/NDE_oceananigans.jl:61-101 — backward-Euler tridiagonal solve of the
    vertical diffusion of u, v (nu) and T (nu_T) over dt, then T'[1] = T_bottom. Dense solve, one column."""
    u, v, T = (np.asarray(a, dtype=np.float64) for a in (u, v, T))
    N = desc.Nz
    dz, dt = np.float64(np.float32(cdesc.dz)), np.float64(np.float32(cdesc.dt))
    nu, nu_T = mpp_diffusivity_oceananigans(desc, cdesc, u, v, T)

    def L(nuq):
        ld = np.array([-dt / dz ** 2 * nuq[i] for i in range(1, N)])       # Julia i in 2:Nz
        ud = np.array([-dt / dz ** 2 * nuq[i + 1] for i in range(0, N - 1)])  # Julia i in 1:Nz-1
        d = np.zeros(N)
        for i in range(N - 1):
            d[i] = 1 + dt / dz ** 2 * (nuq[i] + nuq[i + 1])
        d[N - 1] = 1 + dt / dz ** 2 * nuq[N - 1]
        return np.diag(d) + np.diag(ld, -1) + np.diag(ud, 1)

    u2 = np.linalg.solve(L(nu), u)
    v2 = np.linalg.solve(L(nu), v)
    T2 = np.linalg.solve(L(nu_T), T)
    T2[0] = T[0]
    return u2, v2, T2


def uvt_forcing_column(desc, theta, cdesc, u, v, T):
    """NN_uw_forcing / NN_vw_forcing / NN_wT_forcing: wind_mixing/src/NDE_oceananigans.jl:288-344 (and the callback that fills
    dz_uw_NN, dz_vw_NN, dz_wT_NN, :380-405), one column, dimensional in and out. The momentum chains subtract
    inv(scaling)(uw[1]) of the ALREADY unscaled first output (:292,301 — followed literally); the temperature chain
    subtracts the unscaled first output (:324). enforce_fluxes: [0; NN; top flux] (:281-285); dz at the centres."""
    u, v, T = (np.asarray(a, dtype=np.float64) for a in (u, v, T))
    f64 = lambda x: np.float64(np.float32(x))
    mu = [f64(m) for m in desc.mu]
    sg = [f64(s) for s in desc.sigma]
    th = _net_thetas(desc, np.asarray(theta, dtype=np.float64))
    uvT = np.concatenate([(u - mu[0]) / sg[0], (v - mu[1]) / sg[1], (T - mu[2]) / sg[2]])
    tops = [f64(cdesc.uw_top), f64(cdesc.vw_top), f64(cdesc.wT_top)]
    dz = f64(cdesc.dz)
    out = []
    for q in range(3):
        nn = chain_numpy(th[q], desc.nets[q].sizes, desc.nets[q].acts, uvT)
        un = sg[3 + q] * nn + mu[3 + q]
        if q < 2:
            un = un - (sg[3 + q] * un[0] + mu[3 + q])
        else:
            un = un - (sg[5] * nn[0] + mu[5])
        F = np.concatenate([[0.0], un, [tops[q]]])
        out.append((F[1:] - F[:-1]) / dz)
    return out

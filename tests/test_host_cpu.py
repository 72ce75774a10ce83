"""Host-side logic that needs no GPU: model description, synthetic inputs, Flux containers, loss weights, sharding."""
import numpy as np
import pytest

from cpz_b200 import flux, parallel, synthetic as syn, wind_mixing as wm
from cpz_b200.desc import FLAG_CA, FLAG_MPP, RHS_FREE_CONVECTION, RHS_INFER, RHS_TRAIN, ModelDesc, NetDesc
from cpz_b200.ocean_parameterizations import Dc, Df, MinMaxScaling, ZeroMeanUnitVarianceScaling
from oracle import flux_nn, operators


def test_reference_net_shapes_and_parameter_counts():
    assert syn.NET_SHAPES["uvT_small"]().n_params == 6521      # train_NDE.jl:103
    assert syn.NET_SHAPES["uvT_test"]().n_params == 51231      # test_train_NDE.jl:33
    assert syn.NET_SHAPES["uvT_large"]().n_params == 211631    # construct_NN.jl:32-34
    assert syn.NET_SHAPES["T_only"]().n_params == 24735        # train_free_convection_nde.jl:119-121
    assert syn.wind_mixing_desc().n_params == 19563
    assert syn.NET_SHAPES["uvT_small"]().macs == 6420


def test_model_desc_round_trip_and_validation():
    d = syn.wind_mixing_desc(variant=RHS_INFER, n_steps=1152)
    c = d.to_c()
    assert (c.Nz, c.n_fields, c.n_nets, c.n_steps) == (32, 3, 3, 1152)
    assert list(c.nets[2].sizes)[:4] == [96, 50, 20, 31]
    assert d.n_saved == 1153 and d.S == 96 and d.n_bc == 6
    assert d.n_substeps == 2 and d.rhs_evals_per_step == 12  # 0.1001*600/64 = 0.94 > 0.8
    bad = ModelDesc(Nz=32, n_fields=1, variant=RHS_TRAIN)
    with pytest.raises(AssertionError):
        bad.validate()


def test_substep_rule():
    assert syn.wind_mixing_desc(nu_m=0.01).n_substeps == 1
    assert syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_CA, kappa=1.0).n_substeps == 12  # kappa=1 needs >= 12
    assert syn.free_convection_desc(ca=False).n_substeps == 1
    assert syn.free_convection_desc(ca=True).n_substeps == 4


def test_synthetic_inputs_are_seeded_and_well_formed():
    d = syn.wind_mixing_desc()
    a, b = syn.columns(d, 50, seed=5)
    a2, b2 = syn.columns(d, 50, seed=5)
    np.testing.assert_array_equal(a, a2); np.testing.assert_array_equal(b, b2)
    assert a.shape == (50, 96) and b.shape == (50, 6) and a.dtype == np.float32
    # non-zero shear everywhere (quirk Q4) and bottoms at s_q(0)
    assert np.abs(np.diff(a[:, :32], axis=1)).min() > 0
    np.testing.assert_allclose(b[:, 0], (0 - d.mu[3]) / d.sigma[3], rtol=1e-6)
    assert len({tuple(r) for r in b.round(4)}) == 9  # nine forcing cases
    t = syn.free_convection_desc()
    x, bc = syn.columns(t, 7)
    assert x.shape == (7, 32) and bc.shape == (7, 2)
    th = syn.theta_init(d)
    assert th.shape == (19563,) and np.abs(th).max() < 1e-5 * np.sqrt(6 / 51) * 1.001


def test_flux_destructure_matches_the_oracle_convention():
    rng = np.random.default_rng(0)
    nn = flux.Chain(flux.Dense(5, 4, "mish", rng=rng), flux.Dense(4, 3, rng=rng))
    theta, re = flux.destructure(nn)
    assert theta.shape == (5 * 4 + 4 + 4 * 3 + 3,)
    layers = flux_nn.reconstruct(theta, nn.sizes)
    np.testing.assert_array_equal(layers[0][0], nn.layers[0].W)
    np.testing.assert_array_equal(layers[1][0], nn.layers[1].W)
    nn2 = re(theta * 2)
    np.testing.assert_array_equal(nn2.layers[0].W, 2 * nn.layers[0].W)
    assert nn.net_desc().sizes == [5, 4, 3] and nn.net_desc().acts == ["mish", "identity"]
    np.testing.assert_allclose(nn.scale(1e-5).layers[1].W, nn.layers[1].W * np.float32(1e-5))


def test_host_operators_match_the_oracle_restatement():
    np.testing.assert_array_equal(Dc(32, 1 / 32), operators.D_c(32, 1 / 32))
    np.testing.assert_array_equal(Df(32, 1 / 32), operators.D_f(32, 1 / 32))
    data = np.random.default_rng(1).random((4, 7))
    s = ZeroMeanUnitVarianceScaling(data)
    assert s.mu == data.mean() and s.sigma == data.std(ddof=1)
    np.testing.assert_allclose(s.unscale(s(data)), data)
    m = MinMaxScaling(data, a=-1, b=2)
    np.testing.assert_allclose(m.unscale(m(data)), data)


def test_loss_scalings_mirror():
    losses = dict(zip(wm.LOSS_KEYS, [0.3, 0.2, 0.7, 0.05, 0.04, 0.09]))
    fr = {"T": 0.8, "∂T∂z": 0.8, "profile": 0.5}   # train_NDE.jl:110
    s = wm.calculate_loss_scalings(losses, fr, True)
    L = wm.apply_loss_scalings(losses, s)
    np.testing.assert_allclose(L["T"] / (L["u"] + L["v"]), 0.8 / 0.2)
    np.testing.assert_allclose(L["∂T∂z"] / (L["∂u∂z"] + L["∂v∂z"]), 0.8 / 0.2)
    np.testing.assert_allclose((L["u"] + L["v"] + L["T"]) / (L["∂u∂z"] + L["∂v∂z"] + L["∂T∂z"]), 1.0)
    s0 = wm.calculate_loss_scalings(losses, fr, False)
    assert s0["∂u∂z"] == 0 and s0["∂T∂z"] == 0


def test_time_grid_and_integrator_names():
    t = np.arange(1153) * 600.0
    dt, t0, n, stride = wm._time_grid(t, range(0, 1153, 9), 691200.0)
    assert (n, stride) == (1152, 9) and abs(dt - 1 / 1152) < 1e-12 and t0 == 0
    with pytest.raises(ValueError):
        wm._time_grid(t, [0, 9, 20], 691200.0)
    assert wm._integrator("Tsit5") == "tsit5"
    with pytest.raises(ValueError, match="ROCK4"):
        wm._integrator("ROCK4")


def test_train_NDE_asserts_like_the_reference():
    data = wm.ProfileData(np.zeros((1, 3, 96), np.float32), np.arange(3) * 600.0, np.zeros((1, 6), np.float32),
                          {k: ZeroMeanUnitVarianceScaling(mu=0.0, sigma=1.0) for k in ("u", "v", "T", "uw", "vw", "wT")},
                          np.linspace(-256, 0, 33))
    nn = flux.Chain(flux.Dense(96, 31))
    with pytest.raises(AssertionError):   # NDE_training.jl:171
        wm.train_NDE(nn, nn, nn, data, [0, 1, 2], "Tsit5", [flux.ADAM()], 1, modified_pacanowski_philander=True, convective_adjustment=True)
    with pytest.raises(AssertionError):   # NDE_training.jl:192-194
        wm.train_NDE(nn, nn, nn, data, [0, 1, 2], "Tsit5", [flux.ADAM()], 1, zero_weights=True)


def test_shard_columns_partitions_exactly():
    for ncol, world in [(9216, 8), (4096, 3), (5, 8), (1048576, 8)]:
        spans = [parallel.shard_columns(ncol, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == ncol
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


# ---- data path: mirrors the reference's own test/test_coarse_graining.jl:1-36 ------------------------
def test_coarse_grain_center_matches_reference_test():
    from cpz_b200.ocean_parameterizations import Center, coarse_grain
    y = 0.5 * np.arange(1, 101) - 3.0
    yc = coarse_grain(y, 20, Center)
    assert len(yc) == 20
    assert np.allclose(np.diff(yc), np.diff(yc)[0])
    assert np.isclose(y.mean(), yc.mean())
    with pytest.raises(ValueError):
        coarse_grain(y, 30, Center)


def test_coarse_grain_linear_interpolation_face_matches_reference_test():
    from cpz_b200.ocean_parameterizations import Face, coarse_grain_linear_interpolation
    x = np.arange(1, 101)
    y_lin, y_quad = 0.5 * x - 3.0, 0.5 * x ** 2 - 3.0 * x + 5.0
    yc = coarse_grain_linear_interpolation(y_lin, 20, Face)
    yq = coarse_grain_linear_interpolation(y_quad, 20, Face)
    assert len(yc) == 20 and len(yq) == 20
    assert np.allclose(np.diff(yc), np.diff(yc)[0])
    assert np.isclose(y_lin.mean(), yc.mean())
    assert yc[0] == y_lin[0] and yc[-1] == y_lin[-1]
    # 129 -> 33 faces of the LES regridding (data_containers.jl:343-427): every 4th face exactly
    f = np.linspace(-256.0, 0.0, 129)
    assert np.allclose(coarse_grain_linear_interpolation(f, 33, Face), f[::4])


def test_coarse_grain_face_preserves_end_points_and_block_means():
    from cpz_b200.ocean_parameterizations import Face, coarse_grain
    y = np.arange(1.0, 131.0)            # N = 130: (N-2)/(n-2) = 128/32 = 4 exactly
    yc = coarse_grain(y, 34, Face)
    assert yc[0] == 1.0 and yc[-1] == 130.0
    assert np.allclose(yc[1:-1], y[1:-1].reshape(32, 4).mean(axis=1))
    z = np.arange(1.0, 102.0)            # N = 101, n = 12: non-integer ratio -> rounded windows
    zc = coarse_grain(z, 12, Face)
    assert zc[0] == 1.0 and zc[-1] == 101.0 and np.all(np.diff(zc) > 0)


def test_training_history_file_keeps_the_reference_layout(tmp_path):
    """data_writing.py mirrors wind_mixing/src/data_writing.jl:1-78: same group hierarchy and 1-based counts per stage."""
    from cpz_b200 import data_writing as dw
    from cpz_b200.flux import ADAM, Chain, Dense, destructure
    rng = np.random.default_rng(0)
    nets = [Chain(Dense(96, 5, "mish", rng=rng), Dense(5, 31, rng=rng)) for _ in range(3)]
    opts = [ADAM(3e-4), ADAM(1e-4)]
    path = str(tmp_path / "history.npz")
    dw.write_metadata_NDE_training(path, ["wind_-5e-4_cooling_3e-8_new"], [2, 2], [[0, 100]], {"ν₀": 1e-4, "Pr": 1.0}, opts, *nets)
    scal = {k: 1.0 for k in dw.LOSS_KEYS}
    for it in range(3):
        losses = {k: float(it + 1 + i) for i, k in enumerate(dw.LOSS_KEYS)}
        c = dw.write_data_NDE_training(path, losses, scal, *nets, 1, opts[0], state={"m": np.zeros(3), "v": np.ones(3), "beta_pow": [0.9, 0.999]})
        assert c == it + 1
    assert dw.write_data_NDE_training(path, losses, scal, *nets, 2, opts[1]) == 1  # a new stage starts at 1
    h = dw.read_training_history(path)
    assert float(h["training_data/loss/total/1/2"]) == sum(2 + i for i in range(6))
    assert float(h["training_data/loss/profile/1/1"]) == 1 + 2 + 3 and float(h["training_data/loss/gradient/1/1"]) == 4 + 5 + 6
    np.testing.assert_array_equal(h["training_data/neural_network/uw/1/3/theta"], destructure(nets[0])[0])
    assert list(h["training_info/uw_neural_network/sizes"]) == [96, 5, 31]
    assert float(h["training_data/optimizer/η/2/1"]) == 1e-4 and "training_info/loss_scalings/∂T∂z" in h
    np.testing.assert_allclose(dw.loss_series(path, "T", 1), [3.0, 4.0, 5.0])

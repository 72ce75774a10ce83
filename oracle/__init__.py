"""CPU oracle for the NDE column hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, on the CPU, the arithmetic of the reference's (CliMA/ClimateParameterizations.jl,
a.k.a. OceanParameterizations.jl) neural-differential-equation column solve and its gradient. It exists so the
CUDA engine (libcpz.so) can be checked for parity. Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product path
(`climateparameterizations.jl_b200/`) never does, and fails loudly when the CUDA library is missing.

PARITY UNPINNED. The reference is pure Julia and cannot be executed in this image (no `julia` binary, no LES
input files); the heavy arithmetic it runs lives in third-party Julia packages that are not vendored under
/root/reference:
    Flux 0.11.6 (Dense, Chain, destructure, ADAM)      wind_mixing/Manifest.toml:523
    NNlib 0.7.20 (relu, mish, swish, leakyrelu)        wind_mixing/Manifest.toml:1222
    OrdinaryDiffEq 5.55.1 (Tsit5 tableau)              wind_mixing/Manifest.toml:1340
    DiffEqSensitivity 6.45.0 / Zygote 0.6.11           wind_mixing/Manifest.toml:370,2118
    GalacticOptim 1.2.0                                wind_mixing/Manifest.toml:616
and the reference's own tests hold no golden vector, known-answer test or fixture for the NDE right-hand side,
the solver, the loss or the gradient — only scaler properties (test/test_feature_scaling.jl:1-30) and the
(stale) loss-fraction identities (wind_mixing/test/test_training_scaling.jl:17-19). Those are checked in
tests/test_oracle_*.py. Everything else is pinned only by (i) two independent restatements that must agree
(`literal.py`: a line-by-line dense-matrix translation of the Julia; `nde.py`: a batched stencil form),
(ii) analytic known answers (operators on ramps, pure diffusion decay, inertial oscillation, Tsit5 order
conditions), (iii) finite-difference checks of the autograd gradients.

The reference integrates with adaptive solvers and a continuous adjoint; the engine (per BASELINE.json's
north_star) uses fixed-step explicit integration with the Tsit5 tableau and the exact discrete adjoint, and so
does this oracle (gradients = torch autograd through the unrolled fixed-step solve).

Modules
    operators.py  Dᶜ/Dᶠ/smoothing_filter             src/differentiation_operators.jl:6-29, wind_mixing/src/filtering_operators.jl:1-15
    scaling.py    ZeroMeanUnitVarianceScaling, MinMax src/DataWrangling/feature_scaling.jl:7-54
    flux_nn.py    Dense/Chain/destructure/activations Flux 0.11.6 semantics (call sites NDE_training.jl:11-13,62-64,94-96)
    literal.py    one-column, dense-matrix restatement wind_mixing/src/NDE_training.jl:83-165, training_postprocessing.jl:55-153, ...
    nde.py        batched torch restatement + integrators + loss + ADAM + closure step
"""

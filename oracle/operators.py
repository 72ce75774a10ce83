"""Dense staggered-grid operators, restated from the reference (oracle: test infrastructure only).

Follows src/differentiation_operators.jl:6-29 and wind_mixing/src/filtering_operators.jl:1-15.
Julia indices are 1-based; here row/col 0 is Julia's 1.
"""
import numpy as np


def D_c(N: int, delta: float) -> np.ndarray:
    """Dᶜ(N, Δ): N x (N+1), face -> centre, (f[k+1]-f[k])/Δ.  differentiation_operators.jl:6-14"""
    D = np.zeros((N, N + 1))
    for k in range(N):
        D[k, k] = -1.0
        D[k, k + 1] = 1.0
    return (1.0 / delta) * D


def D_f(N: int, delta: float) -> np.ndarray:
    """Dᶠ(N, Δ): (N+1) x N, centre -> face, rows 1 and N+1 (Julia) are zero.  differentiation_operators.jl:21-29"""
    D = np.zeros((N + 1, N))
    for k in range(1, N):  # Julia k in 2:N
        D[k, k - 1] = -1.0
        D[k, k] = 1.0
    return (1.0 / delta) * D


def smoothing_filter(N: int, filter_width: int) -> np.ndarray:
    """N x N running mean of odd width, shrinking at the ends.  filtering_operators.jl:1-15"""
    assert N >= filter_width and filter_width % 2 == 1
    filt = np.zeros((N, N), dtype=np.float32)
    half = (filter_width - 1) // 2
    for i in range(1, half + 1):  # Julia 1-based i
        filt[i - 1, 0:half + i] = 1.0 / (half + i)
        filt[N - i, N - (half + i):N] = 1.0 / (half + i)
    for i in range(half + 1, N - half + 1):
        filt[i - 1, i - half - 1:i + half] = 1.0 / filter_width
    return filt

"""Feature scalers restated from src/DataWrangling/feature_scaling.jl:7-54 (oracle: test infrastructure only)."""
import numpy as np


class ZeroMeanUnitVarianceScaling:
    """mu = mean(data), sigma = std(data) with Julia's default corrected (n-1) estimator, over the WHOLE array.
    feature_scaling.jl:17-20"""

    def __init__(self, data=None, mu=None, sigma=None):
        if data is not None:
            data = np.asarray(data)
            mu, sigma = data.mean(), data.std(ddof=1)
        self.mu, self.sigma = mu, sigma

    def scale(self, x):  # feature_scaling.jl:22
        return (x - self.mu) / self.sigma

    def unscale(self, y):  # feature_scaling.jl:23
        return self.sigma * y + self.mu

    __call__ = scale  # feature_scaling.jl:53

    def inv(self):  # feature_scaling.jl:54
        return self.unscale


class MinMaxScaling:
    """feature_scaling.jl:29-47"""

    def __init__(self, data, a=0, b=1):
        data = np.asarray(data)
        self.a, self.b = a, b
        self.data_min, self.data_max = data.min(), data.max()

    def scale(self, x):
        return self.a + (x - self.data_min) * (self.b - self.a) / (self.data_max - self.data_min)

    def unscale(self, y):
        return self.data_min + (y - self.a) * (self.data_max - self.data_min) / (self.b - self.a)

    __call__ = scale

    def inv(self):
        return self.unscale

"""Running statistics with the interface of the reference's utils.Average / StandardDev /
StatMeter (reference utils.py:233-317), re-implemented for the receivers in this package."""
import json

import numpy as np


class Average:
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum = self.sum + val * n
        self.count += n
        self.avg = self.sum / self.count


class StandardDev:
    """Welford's online variance (sample variance, n-1 denominator)."""

    def __init__(self):
        self.n = 0
        self.mean = 0
        self.M2 = 0

    def update(self, x):
        self.n += 1
        delta = x - self.mean
        self.mean = self.mean + delta / self.n
        self.M2 = self.M2 + delta * (x - self.mean)

    def variance(self):
        return float("nan") if self.n < 2 else self.M2 / (self.n - 1)

    def stddev(self):
        return self.variance() ** 0.5


class StatMeter:
    """avg + std per (timestep, layer); `results['time_steps'][t][l]['avg'|'std']`."""

    def __init__(self, T, n_layers):
        self.T = T
        self.n_layers = n_layers
        self.results = {"time_steps": {t: {l: {"avg": Average(), "std": StandardDev()} for l in range(n_layers)}
                                       for t in range(T)}}

    def update(self, val, t, n_layer):
        cell = self.results["time_steps"][t][n_layer]
        cell["avg"].update(val)
        cell["std"].update(val)

    def save(self, path):
        out = {"time_steps": {}}
        for t in range(self.T):
            out["time_steps"][t] = {}
            for l in range(self.n_layers):
                cell = self.results["time_steps"][t][l]
                avg, std = cell["avg"].avg, cell["std"].stddev()
                out["time_steps"][t][l] = {
                    "avg": avg.tolist() if isinstance(avg, np.ndarray) else avg,
                    "std": std.tolist() if isinstance(std, np.ndarray) else std,
                }
        with open(path, "w") as f:
            json.dump(out, f)

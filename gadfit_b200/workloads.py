"""Synthetic workloads of BASELINE.json (configs C2-C5) and their set-up.

Data synthesis only: uniform x grids, y = model(truth) + N(0, sigma^2) noise,
explicit sigma array, start values = truth * 1.1 + 0.01.
The numpy model formulas below exist to *generate* y; nothing here is on the
solver path.  The same arrays are handed to the CUDA library and to the
oracle/reference, which is what makes the parity comparison meaningful.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# 15-point Kronrod rule (the rule's own constants), node order as the solver's
_K15_X = np.array([
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.0,
    0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
    0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
    0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
    0.207784955007898467600689403773245])
_K15_W = np.array([
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714,
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649])


def f_gauss_hermite3(p, x):
    y = p[18] + p[19] * x
    for k in range(3):
        A, mu, w, a, b, c = p[6 * k: 6 * k + 6]
        u = (x - mu) / w
        y = y + A * np.exp(-u * u) * (1.0 + u * (a + u * (b + u * c)))
    return y


def f_gauss_linear_bg(p, x):
    A, mu, w, s, b0, b1 = p
    u = (x - mu) / w
    return A * np.exp(-u * u) * (1.0 + s * u) + b0 + b1 * x


_integral_cache: dict = {}


def _cached(key, x, compute):
    """The quadrature term depends only on the shared parameters and the x
    grid, which all data sets of a workload have in common."""
    full = key + (x.shape[0], float(x[0]), float(x[-1]))
    if full not in _integral_cache:
        if len(_integral_cache) > 8:
            _integral_cache.clear()
        _integral_cache[full] = compute()
    return _integral_cache[full]


def f_conv_exp_gauss(p, x):
    A, tau, sg, t0, bgr = p

    def integral():
        t = 5.0 * _K15_X + 5.0  # one panel over [0, 10]
        z = (x[:, None] - t[None, :] - t0) / sg
        integrand = np.exp(-t[None, :] / tau) * np.exp(-z * z) / (sg * np.sqrt(np.pi))
        return 5.0 * (integrand @ _K15_W)
    return A * _cached(("conv", tau, sg, t0), x, integral) + bgr


def f_lorentz_double_int(p, x):
    A, x0, g, b, c, k1, k2 = p

    def integral():
        t = 0.5 * _K15_X + 0.5
        acc = np.zeros_like(x)
        for i in range(15):
            for j in range(15):
                s = x - t[i] - t[j]
                acc += (_K15_W[i] * _K15_W[j] * np.exp(-k1 * t[i] ** 2 - k2 * t[j] ** 2)
                        / (1.0 + s * s))
        return 0.25 * acc
    return A * g * g / ((x - x0) ** 2 + g * g) + b + c * _cached(("lor", k1, k2), x, integral)


def f_exponential(p, x):
    return p[0] * np.exp(-x / p[1]) + p[2]


MODEL_FUNCS = {
    "exponential": f_exponential,
    "gauss_hermite3": f_gauss_hermite3,
    "gauss_linear_bg": f_gauss_linear_bg,
    "conv_exp_gauss": f_conv_exp_gauss,
    "lorentz_double_int": f_lorentz_double_int,
}


@dataclass
class Workload:
    """A global fit: S data sets over one model."""

    name: str
    model: str
    n_pars: int
    shared: list            # parameter indices fitted as global
    local: list             # parameter indices fitted per data set
    n_sets: int
    n_points: int           # per data set
    x: np.ndarray           # [S * n_points] data-set major
    y: np.ndarray
    sigma: np.ndarray
    truth: np.ndarray       # [S][NP]
    start: np.ndarray       # [S][NP]
    lambda0: float = 10.0
    settings: dict = field(default_factory=dict)
    f_val: int = 0          # algorithmic flops of one passive evaluation
    first_set: int = 0      # x / y / sigma hold data sets [first_set, first_set + n_held)

    @property
    def n_total(self) -> int:
        return self.n_sets * self.n_points

    @property
    def nnz_jacobian(self) -> int:
        """Structural non-zeros of J = sum_s N_s (p_g + p_l)."""
        return self.n_total * (len(self.shared) + len(self.local))

    def k1_flops_per_point(self) -> int:
        """Cost table v1 of SURVEY.md 8(d): 4 F_val + (2 + p) + 2 [p(p+1)/2 + p + 1]."""
        p = len(self.shared) + len(self.local)
        return 4 * self.f_val + (2 + p) + 2 * (p * (p + 1) // 2 + p + 1)

    def slice(self, s: int):
        a = (s - self.first_set) * self.n_points
        b = a + self.n_points
        return self.x[a:b], self.y[a:b], self.sigma[a:b]

    def setup(self, solver, owned: range | None = None, pointers: bool = False):
        """addDataset / setPar / settings exactly as a user of the reference
        would write them.  ``owned``: data sets whose arrays this rank holds."""
        for s in range(self.n_sets):
            if owned is not None and s not in owned:
                solver.addDataset(None, n_datapoints=self.n_points)
            else:
                xs, ys, es = self.slice(s)
                if pointers:
                    solver.addDatasetPtr(self.n_points, xs.ctypes.data,
                                         ys.ctypes.data, es.ctypes.data)
                else:
                    solver.addDataset(xs, ys, es)
        for k in self.shared:
            solver.setPar(k, float(self.start[0, k]), True)
        for s in range(self.n_sets):
            for k in self.local:
                solver.setPar(k, float(self.start[s, k]), True, s)
        for k in range(self.n_pars):
            if k not in self.shared and k not in self.local:
                solver.setPar(k, float(self.start[0, k]), False)
        for key, val in self.settings.items():
            setattr(solver.settings, key, val)


def _make(name, model, n_sets, n_points, x_range, sigma, seed, truth, shared,
          local, settings, f_val, alloc=None, owned: range | None = None):
    n_pars = truth.shape[1]
    held = range(n_sets) if owned is None else owned
    total = len(held) * n_points
    if alloc is None:
        x = np.empty(total)
        y = np.empty(total)
        sg = np.empty(total)
    else:
        x, y, sg = alloc(total), alloc(total), alloc(total)
    grid = np.linspace(x_range[0], x_range[1], n_points)
    f = MODEL_FUNCS[model]
    for s in held:
        a = (s - held.start) * n_points
        b = a + n_points
        rng = np.random.default_rng(seed + s)
        x[a:b] = grid
        y[a:b] = f(truth[s], grid) + sigma * rng.standard_normal(n_points)
        sg[a:b] = sigma
    start = truth * 1.1 + 0.01
    return Workload(name, model, n_pars, shared, local, n_sets, n_points, x, y,
                    sg, truth, start, 10.0, settings, f_val, held.start)


def c2(n_sets=64, n_points=1_000_000, alloc=None, owned=None, iteration_limit=10):
    """64 data sets x 1M points, 3 distorted Gaussians, 20 shared parameters."""
    t = np.array([3.0, -2.0, 0.7, 0.05, -0.03, 0.02,
                  5.0, 0.5, 1.1, -0.04, 0.05, -0.01,
                  2.0, 2.5, 0.5, 0.03, 0.02, -0.05,
                  0.3, 0.05])
    truth = np.tile(t, (n_sets, 1))
    return _make("C2", "gauss_hermite3", n_sets, n_points, (-5.0, 5.0), 0.05,
                 20260201, truth, list(range(20)), [],
                 dict(iteration_limit=iteration_limit), 173, alloc, owned)


def c3(n_sets=1000, n_points=4096, alloc=None, owned=None, iteration_limit=10):
    """1000 x 4096, Gaussian + linear background: shared (w, s), local (A, mu, b0, b1)."""
    rng = np.random.default_rng(7)
    truth = np.empty((n_sets, 6))
    truth[:, 0] = rng.uniform(1, 5, n_sets)
    truth[:, 1] = rng.uniform(-1, 1, n_sets)
    truth[:, 2] = 0.8
    truth[:, 3] = 0.1
    truth[:, 4] = rng.uniform(0, 1, n_sets)
    truth[:, 5] = rng.uniform(-0.1, 0.1, n_sets)
    return _make("C3", "gauss_linear_bg", n_sets, n_points, (-4.0, 4.0), 0.02,
                 20260301, truth, [2, 3], [0, 1, 4, 5],
                 dict(iteration_limit=iteration_limit), 55, alloc, owned)


def c4(n_sets=256, n_points=65536, alloc=None, owned=None, iteration_limit=8):
    """256 x 65536, single-integral model + geodesic acceleration:
    shared (tau, sg, t0), local (A, bgr)."""
    rng = np.random.default_rng(11)
    truth = np.empty((n_sets, 5))
    truth[:, 0] = rng.uniform(1, 3, n_sets)
    truth[:, 1] = 2.0
    truth[:, 2] = 0.4
    truth[:, 3] = 0.5
    truth[:, 4] = rng.uniform(0, 0.2, n_sets)
    return _make("C4", "conv_exp_gauss", n_sets, n_points, (-2.0, 12.0), 0.01,
                 20260401, truth, [1, 2, 3], [0, 4],
                 dict(iteration_limit=iteration_limit,
                      acceleration_threshold=0.9), 1500, alloc, owned)


def c5(n_sets=100_000, n_points=256, alloc=None, owned=None, iteration_limit=6):
    """100000 x 256, Lorentzian + double-integral term: shared (gamma, c, k1,
    k2), local (A, x0, b)."""
    rng = np.random.default_rng(13)
    truth = np.empty((n_sets, 7))
    truth[:, 0] = rng.uniform(1, 4, n_sets)
    truth[:, 1] = rng.uniform(-0.5, 0.5, n_sets)
    truth[:, 2] = 0.5
    truth[:, 3] = rng.uniform(0, 0.5, n_sets)
    truth[:, 4] = 0.3
    truth[:, 5] = 1.0
    truth[:, 6] = 2.0
    return _make("C5", "lorentz_double_int", n_sets, n_points, (-3.0, 3.0), 0.02,
                 20260501, truth, [2, 4, 5, 6], [0, 1, 3],
                 dict(iteration_limit=iteration_limit), 12600, alloc, owned)


WORKLOADS = {"C2": c2, "C3": c3, "C4": c4, "C5": c5}

"""Distance laws V(d) for Slater-Koster hoppings (host-side setup).

The k-point hot path needs each law evaluated once per bond at the bond length
(reference: `System.get_hop_params` system.py:205-213 -> `ScalingLaw.__call__` scaling.py:106-113,
cosine cutoff scaling.py:60-75, `_raw_value` of each subclass).  The first derivative `deriv1` (scaling.py:115-128:
V'f + Vf') feeds the Hellmann-Feynman forces; second derivatives (phonons) are not provided.
"""
import numpy as np


class ScalingLaw:
    """V(d) = raw(d) * f_cut(d);  f_cut = 1 below rc-w, cosine ramp on [rc-w, rc], 0 above rc."""

    def __init__(self, V0, d0, cutoff=None, smooth_width=0.5):
        self.V0 = V0
        self.d0 = d0
        self.cutoff = cutoff
        self.smooth_width = smooth_width

    def _raw_value(self, d):  # pragma: no cover - abstract
        raise NotImplementedError

    def _cutoff_func(self, d):
        d = np.asarray(d)
        f = np.ones_like(d, dtype=float)
        if self.cutoff is None:
            return f
        rc = self.cutoff
        rs = rc - self.smooth_width
        ramp = (d >= rs) & (d <= rc)
        f[ramp] = 0.5 + 0.5 * np.cos(np.pi * (d[ramp] - rs) / (rc - rs))
        f[d > rc] = 0.0
        return f

    def __call__(self, d):
        d = np.asarray(d)
        scalar = d.ndim == 0
        d = np.atleast_1d(d)
        out = self._raw_value(d) * self._cutoff_func(d)
        return float(out[0]) if scalar else out

    def _raw_deriv1(self, d):  # pragma: no cover - abstract
        raise NotImplementedError

    def _cutoff_deriv1(self, d):
        d = np.asarray(d)
        df = np.zeros_like(d, dtype=float)
        if self.cutoff is None:
            return df
        rc = self.cutoff
        rs = rc - self.smooth_width
        ramp = (d >= rs) & (d <= rc)
        df[ramp] = -0.5 * np.pi / (rc - rs) * np.sin(np.pi * (d[ramp] - rs) / (rc - rs))
        return df

    def deriv1(self, d):
        """d[V f]/dd = V' f + V f' (scaling.py:115-128)."""
        d = np.asarray(d)
        scalar = d.ndim == 0
        d = np.atleast_1d(d)
        out = self._raw_deriv1(d) * self._cutoff_func(d) + self._raw_value(d) * self._cutoff_deriv1(d)
        return float(out[0]) if scalar else out

    def deriv2(self, *a, **k):
        raise NotImplementedError("second derivatives of scaling laws (phonons) are outside the k-point hot path")


class Constant(ScalingLaw):
    def __init__(self, value):
        super().__init__(V0=value, d0=1.0)
        self.value = value

    def _raw_value(self, d):
        return np.full_like(d, self.value, dtype=float)

    def _raw_deriv1(self, d):
        return np.zeros_like(d, dtype=float)

    def __repr__(self):
        return f"Constant({self.value})"


class Harrison(ScalingLaw):
    def _raw_value(self, d):
        return self.V0 * (self.d0 / d) ** 2

    def _raw_deriv1(self, d):
        return -2 * self.V0 * self.d0 ** 2 / d ** 3


class PowerLaw(ScalingLaw):
    def __init__(self, V0, d0, eta=2.0, cutoff=None, smooth_width=0.5):
        super().__init__(V0, d0, cutoff, smooth_width)
        self.eta = eta

    def _raw_value(self, d):
        return self.V0 * (self.d0 / d) ** self.eta

    def _raw_deriv1(self, d):
        return -self.eta * self.V0 * self.d0 ** self.eta / d ** (self.eta + 1)


class Exponential(ScalingLaw):
    def __init__(self, V0, d0, alpha=1.0, cutoff=None, smooth_width=0.5):
        super().__init__(V0, d0, cutoff, smooth_width)
        self.alpha = alpha

    def _raw_value(self, d):
        return self.V0 * np.exp(-self.alpha * (d - self.d0))

    def _raw_deriv1(self, d):
        return -self.alpha * self.V0 * np.exp(-self.alpha * (d - self.d0))


class GSP(ScalingLaw):
    def __init__(self, V0, d0, n=2.0, nc=4.0, dc=3.5, cutoff=None, smooth_width=0.5):
        super().__init__(V0, d0, cutoff, smooth_width)
        self.n, self.nc, self.dc = n, nc, dc

    def _raw_value(self, d):
        arg = self.n * (-((d / self.dc) ** self.nc) + (self.d0 / self.dc) ** self.nc)
        return self.V0 * (self.d0 / d) ** self.n * np.exp(arg)

    def _raw_deriv1(self, d):
        dlnV = -self.n / d - self.n * self.nc / self.dc * (d / self.dc) ** (self.nc - 1)
        return self._raw_value(d) * dlnV


class Polynomial(ScalingLaw):
    def __init__(self, coeffs, d0, cutoff=None, smooth_width=0.5):
        super().__init__(V0=coeffs[0] if coeffs else 0.0, d0=d0, cutoff=cutoff, smooth_width=smooth_width)
        self.coeffs = np.asarray(coeffs, dtype=float)

    def _raw_value(self, d):
        return np.polyval(self.coeffs[::-1], d - self.d0)

    def _raw_deriv1(self, d):
        if len(self.coeffs) < 2:
            return np.zeros_like(d, dtype=float)
        dc = self.coeffs[1:] * np.arange(1, len(self.coeffs))
        return np.polyval(dc[::-1], d - self.d0)


class Tabulated(ScalingLaw):
    def __init__(self, distances, values, cutoff=None, smooth_width=0.5):
        from scipy.interpolate import CubicSpline

        distances = np.asarray(distances)
        values = np.asarray(values)
        super().__init__(values[0], distances[0], cutoff, smooth_width)
        self.distances, self.values = distances, values
        self._spline = CubicSpline(distances, values, extrapolate=True)
        self._spline_d1 = self._spline.derivative(1)

    def _raw_value(self, d):
        return self._spline(d)

    def _raw_deriv1(self, d):
        return self._spline_d1(d)


class Custom(ScalingLaw):
    def __init__(self, func, d0, deriv1=None, deriv2=None, cutoff=None, smooth_width=0.5):
        super().__init__(func(d0), d0, cutoff, smooth_width)
        self._func = func
        self._deriv1_func = deriv1
        self._h = 1e-5                      # central finite-difference step when no derivative is supplied (scaling.py:421)

    def _raw_value(self, d):
        return np.vectorize(self._func)(d)

    def _raw_deriv1(self, d):
        if self._deriv1_func is not None:
            return np.vectorize(self._deriv1_func)(d)
        d = np.asarray(d)
        return (self._raw_value(d + self._h) - self._raw_value(d - self._h)) / (2 * self._h)


def ensure_scaling(value):
    return value if isinstance(value, ScalingLaw) else Constant(float(value))

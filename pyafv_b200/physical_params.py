"""``PhysicalParams`` -- host-side mirror of the reference's parameter record.

Mirrors ``pyafv/physical_params.py:54-158`` (frozen, validated scalars; ``delta=None`` resolves to ``0.45 r``;
``replace``; the cell-doublet helpers ``get_steady_state`` / ``with_optimal_radius``) and ``target_delta``
(``:201-260``).  The doublet helpers are scalar closed forms that run on the host; they are not part of the
GPU hot path.
"""
from __future__ import annotations

import dataclasses
import math
import numbers

import numpy as np


def _real(value, name: str) -> float:
    """float(value) for finite real scalars; bool / non-numbers are rejected like the reference does."""
    if isinstance(value, (bool, np.bool_)):
        raise TypeError(f"{name} must be a real scalar (float-like), got bool")
    if not isinstance(value, (numbers.Real, np.integer, np.floating)):
        raise TypeError(f"{name} must be a real scalar (float-like), got {type(value).__name__}")
    out = float(value)
    if not math.isfinite(out):
        raise ValueError(f"{name} must be finite, got {value}")
    return out


def _seed(value, name: str):
    if value is None:
        return None
    if isinstance(value, (bool, np.bool_)):
        raise TypeError(f"{name} must be a non-negative integer scalar, got bool")
    if not isinstance(value, (numbers.Integral, np.integer)):
        raise TypeError(f"{name} must be a non-negative integer scalar, got {type(value).__name__}")
    if int(value) < 0:
        raise ValueError(f"{name} must be non-negative, got {value}")
    return int(value)


def _doublet_measures(ell: float, half_dist):
    """Area, perimeter and free-arc length of ONE cell of a symmetric doublet.

    Each cell is a disk of radius ``ell`` cut by the bisector at distance ``half_dist`` from its centre:
    chord half-length c = sqrt(ell^2 - d^2), free arc angle = 2 pi - 2 atan2(c, d).
    """
    c = np.sqrt(np.maximum(ell * ell - half_dist * half_dist, 0.0))
    arc_angle = 2.0 * np.pi - 2.0 * np.arctan2(c, half_dist)
    area = half_dist * c + 0.5 * ell * ell * arc_angle
    arc = ell * arc_angle
    return area, 2.0 * c + arc, arc, c


@dataclasses.dataclass(frozen=True)
class PhysicalParams:
    """Physical parameters of the AFV model (same fields, defaults and validation as the reference).

    r: maximal cell radius (l); A0, P0: preferred area / perimeter; KA, KP: elastic constants; Lambda: tension
    of non-contacting edges; delta: contact truncation threshold (None -> 0.45 r); seed: RNG seed (doublet
    optimisation restarts here; default ``step()`` noise stream in the simulator).
    """

    r: float = 1.0
    A0: float = np.pi
    P0: float = 4.8
    KA: float = 1.0
    KP: float = 1.0
    Lambda: float = 0.2
    delta: float | None = None
    seed: int | None = None

    def __post_init__(self):
        for name in ("r", "A0", "P0", "KA", "KP", "Lambda"):
            object.__setattr__(self, name, _real(getattr(self, name), name))
        object.__setattr__(self, "seed", _seed(self.seed, "seed"))
        if self.delta is None:
            object.__setattr__(self, "delta", 0.45 * self.r)
        else:
            try:
                object.__setattr__(self, "delta", _real(self.delta, "delta"))
            except TypeError:
                raise TypeError(f"delta must be a real scalar (float-like) or None, got {type(self.delta).__name__}") from None

    # -- convenience -------------------------------------------------------------------------------------
    def replace(self, **changes) -> "PhysicalParams":
        """New instance with the given fields replaced (``dataclasses.replace``)."""
        return dataclasses.replace(self, **changes)

    # -- cell-doublet closed forms -------------------------------------------------------------------------
    def _doublet_energy(self, ell: float, half_dist: float) -> float:
        area, perim, arc, _ = _doublet_measures(ell, half_dist)
        return float(self.KA * (area - self.A0) ** 2 + self.KP * (perim - self.P0) ** 2 + self.Lambda * arc)

    def get_steady_state(self) -> tuple[float, float]:
        """Energy-minimising (l, d) of a cell doublet; 2d is the centre-to-centre distance."""
        from scipy.optimize import minimize

        rng = np.random.default_rng(self.seed)

        def cost(q):
            ell = math.exp(q[0])
            frac = 1.0 / (1.0 + math.exp(-q[1])) if q[1] >= 0 else math.exp(q[1]) / (1.0 + math.exp(q[1]))
            return self._doublet_energy(ell, ell * math.sin(0.5 * math.pi * frac))

        best_q, best_val = None, math.inf
        for _ in range(10):
            q0 = rng.normal(size=2)
            with np.errstate(over="ignore", invalid="ignore"):
                res = minimize(cost, q0, method="BFGS", options={"gtol": 1e-8, "maxiter": 10_000})
            if res.fun < best_val:
                best_q, best_val = res.x, res.fun
        ell = math.exp(best_q[0])
        u = best_q[1]
        frac = 1.0 / (1.0 + math.exp(-u)) if u >= 0 else math.exp(u) / (1.0 + math.exp(u))
        return ell, ell * math.sin(0.5 * math.pi * frac)

    def with_optimal_radius(self, digits: int | None = None, delta: float | None = None) -> "PhysicalParams":
        """Copy with ``r`` set to the doublet steady-state radius (``delta`` -> 0.45 r unless given)."""
        ell, _ = self.get_steady_state()
        if digits is not None:
            ell = round(ell, digits)
        return dataclasses.replace(self, r=ell, delta=(0.45 * ell if delta is None else delta))


def target_delta(params: PhysicalParams, target_force: float) -> float:
    """delta at which the doublet detachment force equals ``target_force`` (largest matching separation).

    Scans separations from 1e-6 l to (2 - 1e-6) l in 10^6 steps like the reference (physical_params.py:201-260)
    and interpolates linearly inside the last sign change.
    """
    if not isinstance(params, PhysicalParams):
        raise TypeError("params must be an instance of PhysicalParams")
    ell = params.r
    dist = np.linspace(1e-6, 2.0 - 1e-6, 10**6) * ell
    d = 0.5 * dist
    area, perim, _, c = _doublet_measures(ell, d)
    force = 4.0 * c * (params.KA * (area - params.A0) + params.KP * (perim - params.P0) / (ell + d)
                       + 0.5 * params.Lambda * ell / (c * c))
    f = force - target_force
    hit = (f[:-1] == 0) | (np.signbit(f[:-1]) != np.signbit(f[1:]))
    if not np.any(hit):
        raise ValueError("No valid delta found for the given target force.")
    i = int(np.flatnonzero(hit)[-1])
    x = dist[i] if f[i + 1] == f[i] else dist[i] - f[i] * (dist[i + 1] - dist[i]) / (f[i + 1] - f[i])
    return float(np.sqrt(4.0 * ell * ell - x * x))


__all__ = ["PhysicalParams", "target_delta"]

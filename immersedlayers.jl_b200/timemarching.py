"""IF-HERK time marching of the constrained heat equation on the B200 operators
(BASELINE config C3).

The reference hands its problem functions to ConstrainedSystems.jl (un-vendored): `ode_rhs` =
surface_divergence!(dT, -kappa [T], sys), `constraint_force` = regularize!(dT, -sigma, sys),
`bc_op` = interpolate!, `bc_rhs` = average of the prescribed surface values, `lin_op` = kappa L
(test/literate/heatconduction.jl:87-129, 160-169; src/timemarching.jl:86-107), and the default
integrator is LiskaIFHERK: a 3-stage half-explicit Runge-Kutta scheme with the integrating factor
exp(kappa L dt) (`plan_intfact`).  This module restates that scheme on the accelerated operators:

    dT/dt = kappa L T + r(t) - R sigma,      E T = b(t)
    stage i = 1..s (c_0 = 0, U_0 = T_n), H_i = exp(kappa L (c_i - c_{i-1}) dt):
        w_j <- H_i w_j (j < i),  q <- H_i q,  w_i = H_i r(t_{i-1})
        U*  = q + dt sum_{j<=i} a_ij w_j
        S_i sigma^ = b(t_i) - E U*,   S_i = -E H_i R            (dense LU on the GPU)
        w_i <- w_i - H_i R sigma^ / (dt a_ii),   U_i = U* - H_i R sigma^
    T_{n+1} = U_s,  tableau a, c = Liska & Colonius (2017): c = (1/2, 1, 1), H_3 = I.

For a moving body the plan's tables are refreshed every step (`update_points`, the analogue of
update_system!, src/system.jl:26-50) and the stage complements S_i are rebuilt -- from the compact
integrating-factor table (`create_RTHR_direct`, milliseconds) or column by column through the
convolution engine (`create_RTHR`, the reference's way).  The operators of a step are those of the
body position at the start of the step.

PARITY: the integrator lives in ConstrainedSystems.jl (compat 0.3.8, not in the container); the tableau and
the stage recursion follow the published IF-HERK scheme, the oracle restates the same recursion on the CPU
(oracle/ilm_oracle.py:heat_ifherk_step) and reproduces the temperatures printed by the reference's executed
notebook examples/heatconduction.ipynb (cells 58, 62: 51 and 54 steps) to 2e-14; the GPU path is compared with
the oracle step by step and with the notebook values directly (tests/test_golden.py)."""
from __future__ import annotations

import numpy as np

from . import api
from . import lgf as _lgf

_SQ3 = np.sqrt(3.0)
LISKA_IFHERK = {
    "a": ((0.5, 0.0, 0.0),
          (_SQ3 / 3.0, (3.0 - _SQ3) / 3.0, 0.0),
          ((3.0 + _SQ3) / 6.0, -_SQ3 / 3.0, (3.0 + _SQ3) / 6.0)),
    "c": (0.5, 1.0, 1.0),
}
IF_EULER = {"a": ((1.0,),), "c": (1.0,)}


def timestep_fourier(g, kappa, fourier):
    """dt = Fo dx^2 / kappa (test/literate/heatconduction.jl:182-189)."""
    return fourier * g.dx ** 2 / kappa



def _compact_intfact_table(a, nmax):
    """The plan_intfact table exp(L a) cut where it has decayed to rounding level, for the direct-table form of the stage
    complements (k_schur_direct treats entries beyond the table as zero).  The support grows with the Fourier number a
    (31 entries for a <= 5, 64 at a = 30): the table is doubled until its last entry of the first column is below
    1e-17 of the peak, or the grid size is reached (then nothing is cut)."""
    n = 64
    while True:
        n = min(n, nmax)
        T = _lgf.intfact_table(a, n)
        if n >= nmax or abs(T[n - 1, 0]) <= 1e-17 * abs(T[0, 0]):
            return T
        n *= 2


class DirichletHeatConduction:
    """Unsteady heat conduction with prescribed surface temperatures T+ (outside) / T- (inside) on a
    possibly moving body (test/literate/heatconduction.jl; moving case :386-401).

    body_at(t) -> (x, y, nx, ny, ds); Tplus / Tminus: constants or callables (x, y, t) -> array."""

    def __init__(self, g, body_at, kappa=1.0, fourier=1.0, Tplus=0.0, Tminus=1.0, moving=None, tableau=LISKA_IFHERK,
                 direct_schur=True, lgf_table=None, device=True, ddftype="yang3"):
        self.g, self.body_at, self.kappa = g, body_at, float(kappa)
        self.dt = timestep_fourier(g, kappa, fourier)
        self.Tplus, self.Tminus = Tplus, Tminus
        self.tab_a, self.tab_c = tableau["a"], tableau["c"]
        self.direct_schur, self.device = bool(direct_schur), bool(device)
        self.t, self.nstep = 0.0, 0
        body0 = body_at(0.0)
        self.moving = bool(moving) if moving is not None else False
        self.cache = api.SurfaceScalarCache(body0, g, ddftype=ddftype, lgf_table=lgf_table, device=device)
        self.T = self.cache.zeros_grid()
        # integrating factors H_i = exp(kappa L dc_i dt): argument a_i = kappa/dx^2 * dc_i * dt = Fo * dc_i
        self.stage_a = []
        prev = 0.0
        for c in self.tab_c:
            self.stage_a.append(self.kappa / g.dx ** 2 * (c - prev) * self.dt)
            prev = c
        self.kernel_id, self.table = {}, {}
        for a in sorted(set(self.stage_a)):
            if a > 0.0:
                self.kernel_id[a] = self.cache.add_kernel(_lgf.intfact_table(a, max(g.NX, g.NY)))
                self.table[a] = _compact_intfact_table(a, max(g.NX, g.NY))
            else:
                self.table[a] = np.array([[1.0]])
                self.kernel_id[a] = None                              # H = I
        self._lu = None
        self.stats = {"schur_builds": 0, "plan_refreshes": 0}

    # -- problem functions (heatconduction.jl:87-129) ------------------------------------------------
    def _surface_values(self, f, t):
        x, y = self.cache.points()
        v = f(x, y, t) if callable(f) else np.full(self.cache.N, float(f))
        return np.asarray(v, dtype=np.float64)

    def ode_rhs(self, t):
        """surface_divergence!(dT, -kappa [T], sys) with [T] = T+ - T- (prescribed_surface_jump!)."""
        jump = self._surface_values(self.Tplus, t) - self._surface_values(self.Tminus, t)
        d = self.cache.zeros_surface().set(-self.kappa * jump)
        out = self.cache.zeros_grid()
        api.surface_divergence(out, d, self.cache)
        return out

    def bc_rhs(self, t):
        """prescribed_surface_average!: (T+ + T-)/2."""
        return 0.5 * (self._surface_values(self.Tplus, t) + self._surface_values(self.Tminus, t))

    # -- operators --------------------------------------------------------------------------------
    def _apply_H(self, w, a):
        if self.kernel_id[a] is not None:
            api.convolve(w, self.cache, self.kernel_id[a])
        return w

    def _stage_schur(self, a):
        if self.direct_schur:
            return api.create_RTHR_direct(self.cache, self.table[a])
        if self.kernel_id[a] is None:
            if None not in self.kernel_id:
                self.kernel_id[None] = self.cache.add_kernel(_lgf.intfact_table(0.0, max(self.g.NX, self.g.NY)))
            return api.create_RTHR(self.cache, self.kernel_id[None])
        return api.create_RTHR(self.cache, self.kernel_id[a])

    def _refresh(self, t):
        if self._lu is not None and not self.moving:
            return
        if self._lu is not None:
            self.cache.update_points(self.body_at(t))
            self.stats["plan_refreshes"] += 1
        self._lu = {}
        for a in set(self.stage_a):
            self._lu[a] = api.LU(self._stage_schur(a))
            self.stats["schur_builds"] += 1

    # -- one step ----------------------------------------------------------------------------------
    def step(self):
        cache, dt, t0 = self.cache, self.dt, self.t
        self._refresh(t0)
        s = len(self.tab_c)
        q = self.T                                   # propagated T_n (overwritten stage by stage)
        w = []
        U = None
        c_prev = 0.0
        for i in range(s):
            a_i = self.stage_a[i]
            aii = self.tab_a[i][i]
            for wj in w:
                self._apply_H(wj, a_i)
            self._apply_H(q, a_i)
            wi = self._apply_H(self.ode_rhs(t0 + c_prev * dt), a_i)
            w.append(wi)
            U = cache.zeros_grid()
            U.data[...] = q.data
            for j in range(i + 1):
                if self.tab_a[i][j] != 0.0:
                    U.data += (dt * self.tab_a[i][j]) * w[j].data
            # constraint: S_i sigma^ = b(t_i) - E U*
            EU = cache.zeros_surface()
            api.interpolate(EU, U, cache)
            b = self.bc_rhs(t0 + self.tab_c[i] * dt)
            rhs = (api._dev(b) - EU.data) if cache.device else (b - EU.data)
            sig = self._lu[a_i].solve(rhs)
            corr = cache.zeros_grid()
            api.regularize(corr, api.ScalarData(cache.N, data=sig), cache)
            self._apply_H(corr, a_i)                 # H_i R sigma^
            U.data -= corr.data
            wi.data -= corr.data * (1.0 / (dt * aii))
            c_prev = self.tab_c[i]
            self.sigma = sig
        self.T = U
        self.t = t0 + dt
        self.nstep += 1
        return self.T

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self.T


class UnboundedHeatConduction:
    """Transient heat conduction in an unbounded domain with area / line / point heating and convection by a
    prescribed velocity field (test/literate/heatconduction-unbounded.jl):

        dT/dt = kappa L T - N(v, T) + q(T, t)

    with N = convective_derivative! (src/grid_operators.jl:258-264) and q = apply_forcing! (src/forcing.jl:375-396).
    There is no constraint, so LiskaIFHERK reduces to the integrating-factor Runge-Kutta recursion

        w_j <- H_i w_j (j < i),  q <- H_i q,  w_i = H_i r(U_{i-1}, t_{i-1}),  U_i = q + dt sum_{j<=i} a_ij w_j

    (same tableau and integrating factors H_i = exp(kappa L (c_i - c_{i-1}) dt) as DirichletHeatConduction).
    forcing_models: list of Area/Line/PointForcingModel; velocity_model(vel, t, cache, phys_params) fills the Edges
    `vel` (the "convection velocity model" of the reference's forcing Dict, :128-134); dt = min(Fo dx^2/kappa,
    CFL dx/Umax) as the reference's timestep function (:212-222) when `umax` is given."""

    def __init__(self, g, kappa, forcing_models=None, velocity_model=None, phys_params=None, fourier=1.0, cfl=None, umax=None,
                 tableau=LISKA_IFHERK, lgf_table=None, device=True, ddftype="yang3"):
        from . import forcing as F
        self.g, self.kappa, self.phys_params = g, float(kappa), phys_params
        self.dt = timestep_fourier(g, kappa, fourier)
        if cfl is not None and umax:
            self.dt = min(self.dt, cfl * g.dx / umax)
        self.tab_a, self.tab_c = tableau["a"], tableau["c"]
        z = np.zeros(0)
        self.cache = api.SurfaceScalarCache((z, z, z, z, z), g, ddftype=ddftype, lgf_table=lgf_table, device=device)
        self.fcache = F.ForcingModelAndRegion(forcing_models, self.cache)
        self.velocity_model = velocity_model
        self.vel = self.cache.zeros_gridgrad()
        self.T = self.cache.zeros_grid()
        self.t, self.nstep = 0.0, 0
        self.stage_a, self.kernel_id = [], {}
        prev = 0.0
        for c in self.tab_c:
            self.stage_a.append(self.kappa / g.dx ** 2 * (c - prev) * self.dt)
            prev = c
        for a in sorted(set(self.stage_a)):
            self.kernel_id[a] = self.cache.add_kernel(_lgf.intfact_table(a, max(g.NX, g.NY))) if a > 0.0 else None

    def _apply_H(self, w, a):
        if self.kernel_id[a] is not None:
            api.convolve(w, self.cache, self.kernel_id[a])
        return w

    def ode_rhs(self, T, t):
        """heatconduction_rhs! (heatconduction-unbounded.jl:143-158): dT = -N(v, T) + forcing."""
        from . import forcing as F
        dT = self.cache.zeros_grid()
        if self.velocity_model is not None:
            self.velocity_model(self.vel, t, self.cache, self.phys_params)
            api.convective_derivative(dT, self.vel, T, self.cache)
            api._iscale(dT, -1.0)
        tmp = self.cache.zeros_grid()
        F.apply_forcing(tmp, T, None, t, self.fcache, self.phys_params, None, self.cache)
        api._iadd(dT, tmp)
        return dT

    def step(self):
        dt, t0 = self.dt, self.t
        q, w, U = self.T, [], self.T
        c_prev = 0.0
        for i, c in enumerate(self.tab_c):
            a_i = self.stage_a[i]
            r = self.ode_rhs(U, t0 + c_prev * dt)          # before q (= U_0 = T_n at the first stage) is propagated in place
            for wj in w:
                self._apply_H(wj, a_i)
            self._apply_H(q, a_i)
            w.append(self._apply_H(r, a_i))
            U = self.cache.zeros_grid()
            U.data[...] = q.data
            for j in range(i + 1):
                if self.tab_a[i][j] != 0.0:
                    U.data += (dt * self.tab_a[i][j]) * w[j].data
            c_prev = c
        self.T = U
        self.t = t0 + dt
        self.nstep += 1
        return self.T

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self.T

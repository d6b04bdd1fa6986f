"""CPU restatement ("port") of the reference's ray-tracing algorithm in numpy.

TEST INFRASTRUCTURE -- never imported by the product (graph_framework_b200/).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.

What it restates, with the reference lines it follows:
  * table index rule                    piecewise.hpp:26-65
  * cubic spline coefficient folding    equilibrium.hpp:1121-1133 (build_1D_spline)
  * EFIT psi(R, Z), profiles, B field   equilibrium.hpp:1279-1313 (build_psi), :1324-1384 (set_cache),
                                        incl. the quirks ne_c0/ne_c1 := te_c0/te_c1 (:1478) and ni := te (:1361)
  * analytic equilibria                 equilibrium.hpp:482, 611, 735, 864, 991
  * dispersion functions                dispersion.hpp:449-477 (simple), :510-575 (bohm_gross),
                                        :785-829 (ordinary_wave), :838-895 (extra_ordinary_wave), :903-1008 (cold_plasma)
  * ray equations                       dispersion.hpp:1387-1433
  * RK2 / RK4 step                      solver.hpp:638-665, :811-869
  * Newton initial solve                newton.hpp:42-50 + workflow.hpp:179-205
  * Boris push                          graph_korc/xkorc.cpp:66-103
  * deposition histogram                utilities/bin.py:53-106
  * weak damping k_amp                  absorption.hpp:395-412 with dispersion.hpp:1010-1098 (cold expansion),
                                        :1209-1290 (hot expansion), :289-305 (Z through erfi)
  * power stage                         graph_driver/xrays.cpp:693-736

The reference differentiates D symbolically; this restatement differentiates the SAME function D by
the complex-step method (f(x + ih).imag/h, h = 1e-30), which is exact to rounding for the analytic
expressions involved and is independent of both the reference's reducer and the product's.

Pinning (tests/test_oracle.py): B, ne, te against the reference's golden file efit_gold.nc
(graph_tests/efit_test.cpp:131-187) and every right-hand-side component / trajectory against
outputs of the reference itself (oracle/_ref, fixtures in tests/golden/ref_*.npz).

Known reference defect (documented in DESIGN.md, evidence in tests/test_oracle.py::test_reference_dkz_defect):
for cold_plasma in a z-dependent field the reference's SYMBOLIC dD/dz is wrong (it disagrees with a
finite difference of the reference's own D by orders of magnitude) because of its reduction of
`b_hat->cross(n)->length()`.  This port and the product agree with the finite difference.
"""
import numpy as np

H = 1.0e-30

# dispersion.hpp:494-502
EPSILON0 = 8.8541878138E-12
MU0 = np.pi*4.0E-7
Q = 1.602176634E-19
ME = 9.1093837015E-31
C = 1.0/np.sqrt(EPSILON0*MU0)
MI = 3.34449469E-27        # equilibrium.hpp:1475 deuterium


def table_index(x, scale, offset, n):
    """piecewise.hpp:26-65."""
    u = np.minimum(np.maximum((np.real(x) - offset)/scale, 0.0), n - 1.0)
    u = np.where(np.isnan(u), 0.0, u)
    return u.astype(np.int64)


def fold_spline(c, scale, offset):
    """equilibrium.hpp:1121-1133, same operation order."""
    s2 = scale*scale
    s3 = scale*scale*scale
    c3 = c[3]/s3
    c2 = c[2]/s2 - (3.0*offset)*c[3]/s3
    c1 = c[1]/scale - (2.0*offset)*c[2]/s2 + (3.0*offset*offset)*c[3]/s3
    c0 = c[0] - offset*c[1]/scale + (offset*offset)*c[2]/s2 - (offset*offset*offset)*c[3]/s3
    return c0, c1, c2, c3


def horner(c, x):
    return ((c[3]*x + c[2])*x + c[1])*x + c[0]


def horner_d(c, x):
    return (3.0*c[3]*x + 2.0*c[2])*x + c[1]


class Efit:
    def __init__(self, tables):
        """tables: dict as read from the GFBT file (graph_framework_b200.tools.gfbt.read_gfbt)."""
        t = tables
        self.rmin, self.dr = float(np.ravel(t["rmin"])[0]), float(np.ravel(t["dr"])[0])
        self.zmin, self.dz = float(np.ravel(t["zmin"])[0]), float(np.ravel(t["dz"])[0])
        self.psimin, self.dpsi = float(np.ravel(t["psimin"])[0]), float(np.ravel(t["dpsi"])[0])
        self.ne_scale, self.te_scale, self.pres_scale = float(np.ravel(t["ne_scale"])[0]), float(np.ravel(t["te_scale"])[0]), float(np.ravel(t["pres_scale"])[0])
        self.numr, self.numz = t["psi_c00"].shape
        self.psi = [fold_spline([t["psi_c%d%d" % (i, j)].ravel() for j in range(4)], self.dz, self.zmin)
                    for i in range(4)]
        te = [t["te_c%d" % i] for i in range(4)]
        ne = [te[0], te[1], t["ne_c2"], t["ne_c3"]]                 # equilibrium.hpp:1478
        self.te = fold_spline(te, self.dpsi, self.psimin)
        self.ne = fold_spline(ne, self.dpsi, self.psimin)
        self.pres = fold_spline([t["pressure_c%d" % i] for i in range(4)], self.dpsi, self.psimin)
        self.fpol = fold_spline([t["fpol_c%d" % i] for i in range(4)], self.dpsi, self.psimin)
        self.npsi = te[0].size

    def _psi(self, r, z):
        cell = table_index(r, self.dr, self.rmin, self.numr)*self.numz + table_index(z, self.dz, self.zmin, self.numz)
        ci = [[c[cell] for c in self.psi[i]] for i in range(4)]
        cz = [horner(ci[i], z) for i in range(4)]
        cz_z = [horner_d(ci[i], z) for i in range(4)]
        rn = (r - self.rmin)/self.dr
        psi = ((cz[3]*rn + cz[2])*rn + cz[1])*rn + cz[0]
        psi_z = ((cz_z[3]*rn + cz_z[2])*rn + cz_z[1])*rn + cz_z[0]
        psi_r = ((3.0*cz[3]*rn + 2.0*cz[2])*rn + cz[1])/self.dr
        return psi, psi_r, psi_z

    def _profile(self, c, psi):
        k = table_index(psi, self.dpsi, self.psimin, self.npsi)
        return horner([a[k] for a in c], psi)

    def fields(self, x, y, z):
        r = np.sqrt(x*x + y*y)
        psi, psi_r, psi_z = self._psi(r, z)
        ne = self.ne_scale*self._profile(self.ne, psi)
        te = self.te_scale*self._profile(self.te, psi)
        pressure = self.pres_scale*self._profile(self.pres, psi)
        q = 1.60218E-19                                             # equilibrium.hpp:1359
        ni = te                                                     # equilibrium.hpp:1361
        ti = (pressure - ne*te*q)/(ni*q)
        br = psi_z/r
        bp = self._profile(self.fpol, psi)/r
        bz = -psi_r/r
        cos, sin = x/r, y/r                                         # trigonometry.hpp:85-91, 342-348
        b = (br*cos - bp*sin, br*sin + bp*cos, bz)
        return {"ne": ne, "ni": ni, "te": te, "ti": ti, "b": b, "psi": psi, "pressure": pressure}


class Analytic:
    def __init__(self, name):
        self.name = name

    def fields(self, x, y, z):
        one = np.ones_like(x)
        zero = np.zeros_like(x)
        n = self.name
        if n == "no_magnetic_field":
            ne, te, b = 1.0E19*(0.1*x + 1.0), 1000.0*one, (zero, zero, zero)
        elif n == "slab":
            ne, te, b = 1.0E19*one, 1000.0*one, (zero, zero, 0.1*x + 1.0)
        elif n == "slab_density":
            ne, te, b = 1.0E19*(0.1*x + 1.0), 1000.0*one, (zero, zero, one)
        elif n == "slab_field":
            ne, te, b = 1.0E19*(0.01*x + 1.0), 2000.0*(0.01*x + 1.0), (zero, zero, 0.01*x + 1.0)
        elif n == "gaussian_density":
            ne, te, b = 1.0E19*np.exp((x*x + y*y)/-0.2), 1000.0*one, (one, zero, zero)
        else:
            raise ValueError(n)
        return {"ne": ne, "ni": ne, "te": te, "ti": te, "b": b}


def make_equilibrium(name, tables=None):
    return Efit(tables) if name == "efit" else Analytic(name)


def _dot(a, b):
    return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]


def _cross(a, b):
    return (a[1]*b[2] - a[2]*b[1], a[2]*b[0] - a[0]*b[2], a[0]*b[1] - a[1]*b[0])


def dispersion(name, eq, w, kx, ky, kz, x, y, z, t):
    """D(w, k, x, t) for complex or real array arguments."""
    f = None if name in ("simple",) else eq.fields(x, y, z)
    k = (kx, ky, kz)
    if name == "simple":
        return kz*kz/(w*w) + (kx*kx + ky*ky)/(w*w) - 1.0
    wpe2 = f["ne"]*Q*Q/(EPSILON0*ME*C*C)
    b = f["b"]
    if name == "bohm_gross":
        vterm2 = 2.0*Q*f["te"]/(ME*C*C)
        if eq.__class__ is Analytic and eq.name == "no_magnetic_field":
            kpara2 = _dot(k, k)
        else:
            bl = np.sqrt(_dot(b, b))
            kpara = _dot((b[0]/bl, b[1]/bl, b[2]/bl), k)
            kpara2 = kpara*kpara
        return wpe2 + 3.0/2.0*kpara2*vterm2 - w*w
    bl = np.sqrt(_dot(b, b))
    bh = (b[0]/bl, b[1]/bl, b[2]/bl)
    n = (kx/w, ky/w, kz/w)
    w2 = w*w
    if name == "ordinary_wave":
        nperp = _cross(bh, n)
        return 1.0 - wpe2/w2 - _dot(nperp, nperp)
    if name == "extra_ordinary_wave":
        wec = -Q*bl/(ME*C)
        nperp = _cross(bh, n)
        wh = wpe2 + wec*wec
        return 1.0 - wpe2/w2*(w2 - wpe2)/(w2 - wh) - _dot(nperp, nperp)
    if name == "cold_plasma":
        ec = -Q*bl/(ME*C)
        denome = 1.0 - ec*ec/w2
        e11 = 1.0 - (wpe2/w2)/denome
        e12 = ((ec/w)*(wpe2/w2))/denome
        e33 = wpe2
        wpi2 = f["ni"]*Q*Q/(EPSILON0*MI*C*C)
        ic = Q*bl/(MI*C)
        denomi = 1.0 - ic*ic/w2
        e11 = e11 - (wpi2/w2)/denomi
        e12 = e12 + ((ic/w)*(wpi2/w2))/denomi
        e33 = e33 + wpi2
        e12 = -1.0*e12
        e33 = 1.0 - e33/w2
        npara = _dot(bh, n)
        npara2 = npara*npara
        cr = _cross(bh, n)
        nperp = np.sqrt(_dot(cr, cr))
        nperp2 = nperp*nperp
        m11 = e11 - npara2
        m12 = e12
        m13 = npara*nperp
        m22 = e11 - npara2 - nperp2
        m33 = e33 - nperp2
        return (m11*m22 - m12*m12)*m33 - m22*(m13*m13)
    raise ValueError(name)


ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")


def cold_plasma_reference_defect(eq, s):
    """What the reference's symbolic dD/dz carries on top of the true derivative for cold_plasma in
    the EFIT field.  Its reducer rewrites ((a b)^2 c)/(d^2 b^4) as a^2 c/d^4 (arithmetic.hpp divide /
    multiply reductions; reproduced with four plain variables by `ref_driver reducer`), which hits the
    quotient-rule term of d(n_par^2)/dz and d(n_perp^2)/dz inside m11 and m22 (dispersion.hpp:994-1003):
        extra = K ((p + q) m13^2 - (p m22 + m11 (p + q)) m33),   p = n_par^2, q = n_perp^2,
        K = -(d(B.B)/dz / B.B) (w^2/((B.B)^2 R^8) - 1).
    Pinned to the reference's own kernels in tests/test_oracle.py."""
    def elements(x, y, z, w, kx, ky, kz):
        f = eq.fields(x, y, z)
        b = f["b"]
        g = _dot(b, b)
        bl = np.sqrt(g)
        bh = (b[0]/bl, b[1]/bl, b[2]/bl)
        n = (kx/w, ky/w, kz/w)
        w2 = w*w
        wpe2 = f["ne"]*Q*Q/(EPSILON0*ME*C*C)
        ec = -Q*bl/(ME*C)
        denome = 1.0 - ec*ec/w2
        e11 = 1.0 - (wpe2/w2)/denome
        e33 = wpe2
        wpi2 = f["ni"]*Q*Q/(EPSILON0*MI*C*C)
        ic = Q*bl/(MI*C)
        denomi = 1.0 - ic*ic/w2
        e11 = e11 - (wpi2/w2)/denomi
        e33 = 1.0 - (e33 + wpi2)/w2
        npara = _dot(bh, n)
        cr = _cross(bh, n)
        p, q = npara*npara, _dot(cr, cr)
        return g, p, q, e11 - p, e11 - p - q, e33 - q
    c = {k: np.asarray(s[k], dtype=np.complex128) for k in ORDER}
    g_z = elements(c["x"], c["y"], c["z"] + 1j*H, c["w"], c["kx"], c["ky"], c["kz"])[0].imag/H
    g, p, q, m11, m22, m33 = (v.real for v in elements(c["x"], c["y"], c["z"], c["w"], c["kx"], c["ky"], c["kz"]))
    r2 = np.asarray(s["x"], dtype=np.float64)**2 + np.asarray(s["y"], dtype=np.float64)**2
    w = np.asarray(s["w"], dtype=np.float64)
    K = -(g_z/g)*(w*w/(g*g*r2**4) - 1.0)
    return K*((p + q)*(p*q) - (p*m22 + m11*(p + q))*m33)


def rhs(name, eq, s, reference_defects=True):
    """dispersion.hpp:1387-1433 for Cartesian equilibria: returns dict with dxdt..dkzdt and D.
    reference_defects: reproduce the reference's effective dD/dz for cold_plasma + EFIT (see
    cold_plasma_reference_defect) instead of the true derivative."""
    base = {k: np.asarray(s[k], dtype=np.complex128) for k in ORDER}

    def d(var):
        p = dict(base)
        p[var] = base[var] + 1j*H
        return dispersion(name, eq, p["w"], p["kx"], p["ky"], p["kz"], p["x"], p["y"], p["z"], p["t"]).imag/H

    D = dispersion(name, eq, base["w"], base["kx"], base["ky"], base["kz"], base["x"], base["y"], base["z"], base["t"]).real
    dDdw = d("w")
    dDdz = d("z")
    if reference_defects and name == "cold_plasma" and isinstance(eq, Efit):
        dDdz = dDdz + cold_plasma_reference_defect(eq, s)
    return {"dxdt": -d("kx")/dDdw, "dydt": -d("ky")/dDdw, "dzdt": -d("kz")/dDdw,
            "dkxdt": d("x")/dDdw, "dkydt": d("y")/dDdw, "dkzdt": dDdz/dDdw, "D": D}


EVOLVED = (("kx", "dkxdt"), ("ky", "dkydt"), ("kz", "dkzdt"), ("x", "dxdt"), ("y", "dydt"), ("z", "dzdt"))


def rk4_step(name, eq, s, dt, reference_defects=True):
    """solver.hpp:811-869.  Returns (new state, residual = D^2 at the old state)."""
    s = {k: np.asarray(s[k], dtype=np.float64) for k in ORDER}
    f1 = rhs(name, eq, s, reference_defects)
    k1 = {v: dt*f1[r] for v, r in EVOLVED}
    s2 = dict(s, t=s["t"] + dt/2.0, **{v: s[v] + k1[v]/2.0 for v, _ in EVOLVED})
    f2 = rhs(name, eq, s2, reference_defects)
    k2 = {v: dt*f2[r] for v, r in EVOLVED}
    s3 = dict(s, t=s["t"] + dt/2.0, **{v: s[v] + k2[v]/2.0 for v, _ in EVOLVED})
    f3 = rhs(name, eq, s3, reference_defects)
    k3 = {v: dt*f3[r] for v, r in EVOLVED}
    s4 = dict(s, t=s["t"] + dt, **{v: s[v] + k3[v] for v, _ in EVOLVED})
    f4 = rhs(name, eq, s4, reference_defects)
    k4 = {v: dt*f4[r] for v, r in EVOLVED}
    out = dict(s, t=s["t"] + dt)
    for v, _ in EVOLVED:
        out[v] = s[v] + (k1[v] + 2.0*(k2[v] + k3[v]) + k4[v])/6.0
    return out, f1["D"]**2


def rk2_step(name, eq, s, dt):
    """solver.hpp:638-665."""
    s = {k: np.asarray(s[k], dtype=np.float64) for k in ORDER}
    f1 = rhs(name, eq, s)
    k1 = {v: dt*f1[r] for v, r in EVOLVED}
    s2 = dict(s, t=s["t"] + dt, **{v: s[v] + k1[v] for v, _ in EVOLVED})
    f2 = rhs(name, eq, s2)
    k2 = {v: dt*f2[r] for v, r in EVOLVED}
    out = dict(s, t=s["t"] + dt)
    for v, _ in EVOLVED:
        out[v] = s[v] + (k1[v] + k2[v])/2.0
    return out, f1["D"]**2


def trace(name, eq, s, dt, nsteps, order=4, reference_defects=True):
    res = None
    for _ in range(nsteps):
        s, res = rk4_step(name, eq, s, dt, reference_defects) if order == 4 else rk2_step(name, eq, s, dt)
    return s, res


def newton(name, eq, s, var="kx", tolerance=1.0e-30, max_iterations=1000, per_ray=True):
    """newton.hpp:42-50 driven by workflow.hpp:179-205.  per_ray=False is the reference's
    ensemble-maximum stopping rule; per_ray=True applies the same rule to each ray."""
    s = {k: np.array(s[k], dtype=np.float64) for k in ORDER}
    n = s[var].size
    active = np.ones(n, dtype=bool)

    def one_iteration(mask):
        base = {k: s[k][mask].astype(np.complex128) for k in ORDER}
        p = dict(base)
        p[var] = base[var] + 1j*H
        Dc = dispersion(name, eq, p["w"], p["kx"], p["ky"], p["kz"], p["x"], p["y"], p["z"], p["t"])
        D = dispersion(name, eq, base["w"], base["kx"], base["ky"], base["kz"], base["x"], base["y"], base["z"], base["t"]).real
        s[var][mask] = s[var][mask] - D/(Dc.imag/H)
        return D*D

    big = np.finfo(np.float64).max
    if not per_ray:
        it = 0
        res = one_iteration(active).max()
        last = off = big
        while abs(res) > abs(tolerance) and abs(last - res) > abs(tolerance) and abs(off - res) > abs(tolerance) and it < max_iterations:
            it += 1
            last = res
            if not it % 2:
                off = res
            res = one_iteration(active).max()
        return s
    it = np.zeros(n, dtype=np.int64)
    res = one_iteration(active)
    last = np.full(n, big)
    off = np.full(n, big)
    while True:
        cont = (np.abs(res) > abs(tolerance)) & (np.abs(last - res) > abs(tolerance)) & \
               (np.abs(off - res) > abs(tolerance)) & (it < max_iterations) & active
        it = np.where(cont, it + 1, it)
        active = cont
        if not active.any():
            break
        last = np.where(active, res, last)
        off = np.where(active & (it % 2 == 0), res, off)
        res_new = one_iteration(active)
        res = res.copy()
        res[active] = res_new
    return s


def boris_step(eq, p, dt, b0, larmor_radius):
    """graph_korc/xkorc.cpp:87-103 (normalised relativistic Boris push)."""
    x, y, z, ux, uy, uz, gamma = p
    f = eq.fields(x, y, z)
    b = tuple(c/b0 for c in f["b"])
    u = (ux, uy, uz)
    cr = _cross(u, b)
    up = tuple(u[i] - dt*cr[i]/(2.0*gamma) for i in range(3))
    tau = tuple(-0.5*dt*b[i] for i in range(3))
    tau_sq = _dot(tau, tau)
    sigma = 1.0 + _dot(up, up) - tau_sq
    ustar = _dot(up, tau)
    gamma_next = np.sqrt(0.5*(sigma + np.sqrt(sigma*sigma + 4.0*(tau_sq + ustar*ustar))))
    t = tuple(c/gamma_next for c in tau)
    s = 1.0 + _dot(t, t)
    upt = _dot(up, t)
    cr2 = _cross(up, t)
    un = tuple((up[i] + upt*t[i] + cr2[i])/s for i in range(3))
    pos = (x, y, z)
    pn = tuple(pos[i] + larmor_radius*dt*un[i]/gamma_next for i in range(3))
    return pn + un + (gamma_next,)


def deposit(x, y, z, weight, lo, hi, bins):
    """utilities/bin.py:53-106: half-open uniform bins, weights summed per bin."""
    hist = np.zeros(bins, dtype=np.float64)
    fx = np.floor((x - lo[0])*(bins[0]/(hi[0] - lo[0])))
    fy = np.floor((y - lo[1])*(bins[1]/(hi[1] - lo[1])))
    fz = np.floor((z - lo[2])*(bins[2]/(hi[2] - lo[2])))
    ok = (fx >= 0) & (fx < bins[0]) & (fy >= 0) & (fy < bins[1]) & (fz >= 0) & (fz < bins[2])
    np.add.at(hist, (fx[ok].astype(int), fy[ok].astype(int), fz[ok].astype(int)), weight[ok])
    return hist


def _expansion_terms(eq, w, k, x, y, z):
    f = eq.fields(x, y, z)
    b = f["b"]
    bl = np.sqrt(_dot(b, b))
    bh = (b[0]/bl, b[1]/bl, b[2]/bl)
    ec = Q*bl/(ME*C)
    wpe2 = f["ne"]*Q*Q/(EPSILON0*ME*C*C)
    P = wpe2/(w*w)
    q = P/(2.0*(1.0 + ec/w))
    n = (k[0]/w, k[1]/w, k[2]/w)
    n2 = _dot(n, n)
    npara = _dot(n, bh)
    npara2 = npara*npara
    cr = _cross(bh, n)
    nperp2 = _dot(cr, cr)
    q_func = 1.0 - 2.0*q
    n_func = n2 + npara2
    p_func = 1.0 - P
    gamma1 = (1.0 - q)*n2*nperp2 + p_func*(n2*npara2 - (1.0 - q)*n_func) + q_func*(p_func - nperp2)
    return dict(f=f, ec=ec, P=P, q=q, n2=n2, npara=npara, npara2=npara2, nperp2=nperp2, q_func=q_func,
                n_func=n_func, p_func=p_func, gamma1=gamma1)


def cold_plasma_expansion(eq, w, kx, ky, kz, x, y, z):
    """dispersion.hpp:1010-1098."""
    e = _expansion_terms(eq, w, (kx, ky, kz), x, y, z)
    gamma0 = e["nperp2"]*(e["n2"] - 2.0*e["q_func"]) + e["p_func"]*(2.0*e["q_func"] - e["n_func"])
    return -e["P"]/2.0*(1.0 + e["ec"]/w)*gamma0 + (1.0 - e["ec"]*e["ec"]/(w*w))*e["gamma1"]


def erfi_real(x):
    """erfi for real x as the reference evaluates it (special_functions.hpp:1504-1512): overflow is
    replaced by +-DBL_MAX above x^2 = 720.  scipy's erfi (an independent Faddeeva build) supplies the
    values."""
    from scipy.special import erfi
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore"):
        return np.where(x*x > 720.0, np.copysign(np.finfo(np.float64).max, x), erfi(x))


def hot_plasma_expansion(eq, w, kx, ky, kz, x, y, z):
    """dispersion.hpp:1209-1290 with Z from z_erfi (:289-305); real arguments, complex result."""
    e = _expansion_terms(eq, w, (kx, ky, kz), x, y, z)
    ve = np.sqrt(2.0*Q*e["f"]["te"]/ME)
    vtnorm = ve/C
    ec, P, npara = e["ec"], e["P"], e["npara"]
    with np.errstate(all="ignore"):
        zeta = (1.0 - ec/w)/(npara*vtnorm)
        Z = -np.sqrt(np.pi)*np.exp(-zeta*zeta)*(erfi_real(zeta) - 1j)
        gamma5 = P*(e["n2"]*e["npara2"] - (1.0 - e["q"])*e["n_func"] + e["q_func"])
        gamma2 = P*w/ec*e["nperp2"]*(e["n2"] - e["q_func"]) + \
                 P*P*w*w/(4.0*ec*ec)*(e["n_func"] - 2.0*e["q_func"])*e["nperp2"]/e["npara2"]
        D = -(1.0 + ec/w)*npara*vtnorm*(e["gamma1"] + gamma2 +
                                         e["nperp2"]/(2.0*npara)*(w*w/(ec*ec))*vtnorm*zeta*gamma5)*(1.0/Z + zeta)
    # Z underflows to 0 for |zeta| > 27.2 and 1/Z is 0/0; the reference's SAFE_MATH kernels deliver
    # D = 0 there (cpu_context.hpp:533-544 stores NaN as 0; measured with oracle/_ref).
    return np.where(np.isnan(D.real), 0.0, D.real) + 1j*np.where(np.isnan(D.imag), 0.0, D.imag)


def weak_damping(eq, s):
    """absorption.hpp:395-412 for Cartesian equilibria: complex k_amp per ray."""
    base = {k: np.asarray(s[k], dtype=np.complex128) for k in ORDER}

    def d(var):
        p = dict(base)
        p[var] = base[var] + 1j*H
        return cold_plasma_expansion(eq, p["w"], p["kx"], p["ky"], p["kz"], p["x"], p["y"], p["z"]).imag/H

    r = {k: np.asarray(s[k], dtype=np.float64) for k in ORDER}
    klen = np.sqrt(r["kx"]**2 + r["ky"]**2 + r["kz"]**2)
    with np.errstate(all="ignore"):
        slope = (r["kx"]*d("kx") + r["ky"]*d("ky") + r["kz"]*d("kz"))/klen
        Dw = hot_plasma_expansion(eq, r["w"], r["kx"], r["ky"], r["kz"], r["x"], r["y"], r["z"])
        return klen - Dw/slope


def power_stage(records_xyz, kamp_im):
    """xrays.cpp:693-776: records_xyz [nrec, 3, n], kamp_im [nrec, n] -> power, d_power [nrec, n].
    p_next uses k_sum before the current segment is added; record 0 is the initial state."""
    nrec, _, n = records_xyz.shape
    power = np.ones((nrec, n))
    d_power = np.zeros((nrec, n))
    k_sum = np.zeros(n)
    with np.errstate(all="ignore"):
        for j in range(1, nrec):
            dl = np.sqrt(((records_xyz[j] - records_xyz[j - 1])**2).sum(axis=0))
            p_next = np.exp(-2.0*k_sum)
            d_power[j] = np.sqrt((p_next - power[j - 1])**2)
            k_sum = kamp_im[j]*dl + k_sum
            power[j] = p_next
    return power, d_power

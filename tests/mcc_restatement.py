"""TEST INFRASTRUCTURE: plain-Python restatement of MC_MEX_Ionization (ch4/v3/src/Interactions.cpp:476-762, 845-947) with the
random draws passed in, so that (1) its deterministic parts can be pinned against the compiled reference on CPU
(tests/test_oracle_vs_reference.py) and (2) the device kernel can be checked pair by pair by replaying its Philox stream
(tests/test_gpu_stochastic.py).  Libm's log / exp / pow / tan / atan / sin / cos stand in for the device's (1-2 ulp apart).
"""
import math

import numpy as np

QE = 1.602176565e-19      # all.h:16
ME = 9.10938215e-31       # all.h:18
PI = 3.141592653          # all.h:20


class MccModel:
    def __init__(self, m_n, m_e, E_ion_J, tab_E, tab_s, dv, mpw0_n, mpw0_e):
        self.m_n, self.m_e, self.sum_mass = m_n, m_e, m_n + m_e
        m_r = m_n * m_e / (m_n + m_e)                                  # :503
        self.E_rel_eV = 0.5 * m_r / QE                                 # :509
        self.E_ele_eV = ME * 0.5 / QE                                  # Interactions.h:123
        self.two_qe_me = 2 * QE / ME                                   # Interactions.h:124
        self.c0, self.c1, self.c2, self.B_inc = 1.015e-18, 9.793e+00, 6.181e+01, 10.0      # :513-518
        self.E_ion_eV = E_ion_J / QE                                   # :499
        self.inv_dv = 1 / dv                                           # :500-501
        order = np.argsort(tab_E, kind="stable")
        self.tab_E, self.tab_s = [float(x) for x in np.asarray(tab_E)[order]], [float(x) for x in np.asarray(tab_s)[order]]
        self.w_max0 = 1e-14 * max(mpw0_e, mpw0_n)                      # Interactions.h:129, .cpp:534

    def sigma_coll(self, E):                                           # evaluateSigmaColl :541-558 (std::map lower_bound + interpolation)
        lo, hi = 0, len(self.tab_E)
        while lo < hi:
            mid = (lo + hi) >> 1
            if self.tab_E[mid] < E:
                lo = mid + 1
            else:
                hi = mid
        if lo == 0:
            return self.tab_s[0]
        if lo == len(self.tab_E):
            return self.tab_s[-1]
        x1, x2, y1, y2 = self.tab_E[lo - 1], self.tab_E[lo], self.tab_s[lo - 1], self.tab_s[lo]
        return y1 + (E - x1) * (y2 - y1) / (x2 - x1)

    def sigma_ion(self, E):                                            # evaluateSigmaIon :559-566
        if E <= self.E_ion_eV:
            return 0.0
        return self.c0 * math.log(E / self.c1) / E * math.exp(-self.c2 / E)

    def new_velocity_electron(self, r, E, u):                          # newVelocityElecton :845-862 (IONIZE_1, LAB frame)
        cos_ksi = (2 + E - 2 * math.pow(1 + E, next(r))) / E
        sq = 1 - cos_ksi * cos_ksi
        sin_ksi = math.sqrt(sq) if sq >= 0 else float("nan")
        phi = 2 * PI * next(r)
        v_mag = math.sqrt(E * self.two_qe_me)
        ixu = [0.0 * u[2] - 0.0 * u[1], 0.0 * u[0] - 1.0 * u[2], 1.0 * u[1] - 0.0 * u[0]]          # (1,0,0) x u, not normalised
        uxi = [u[1] * ixu[2] - u[2] * ixu[1], u[2] * ixu[0] - u[0] * ixu[2], u[0] * ixu[1] - u[1] * ixu[0]]
        sp, cp = math.sin(phi), math.cos(phi)
        return [(cos_ksi * u[c] + ixu[c] * sin_ksi * sp + uxi[c] * sin_ksi * cp) * v_mag for c in range(3)]

    def collide(self, r, vn, ve, s_coll):
        """collide :885-947.  Returns (ionised, new electron velocity, velocity of the created electron); vn never changes."""
        g = [vn[c] - ve[c] for c in range(3)]
        g_mag = math.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2])
        m_r = self.m_n * self.m_e / self.sum_mass
        E_rel_J = 0.5 * m_r * g_mag * g_mag
        p_ion = self.sigma_ion(E_rel_J / QE) / s_coll
        if next(r) <= p_ion:
            ve_mag = math.sqrt(ve[0] * ve[0] + ve[1] * ve[1] + ve[2] * ve[2])
            E_inc = ve_mag * ve_mag * self.E_ele_eV
            if E_inc < self.E_ion_eV:                                  # :900-905
                return False, list(ve), [0.0, 0.0, 0.0]
            E_ej = 10.0 * math.tan(next(r) * math.atan((E_inc - self.E_ion_eV) / (2 * self.B_inc)))
            E_sc = E_inc - self.E_ion_eV - E_ej
            if E_sc < 0:
                E_sc = 0.000001
            inv = 1.0 / ve_mag
            u = [ve[0] * inv, ve[1] * inv, ve[2] * inv]
            ve_new = self.new_velocity_electron(r, E_sc, u)
            v_new = self.new_velocity_electron(r, E_ej, u)
            return True, ve_new, v_new
        inv_sum = 1.0 / self.sum_mass                                  # elastic, isotropic in the centre of mass (:933-947)
        cm = [(self.m_n * vn[c] + self.m_e * ve[c]) * inv_sum for c in range(3)]
        cos_ksi = 2 * next(r) - 1
        sin_ksi = math.sqrt(1 - cos_ksi * cos_ksi)
        eps = 2 * PI * next(r)
        g = [g_mag * cos_ksi, g_mag * sin_ksi * math.cos(eps), g_mag * sin_ksi * math.sin(eps)]
        f = self.m_n / self.sum_mass
        return False, [cm[c] - f * g[c] for c in range(3)], [0.0, 0.0, 0.0]

    def apply_cell(self, r, neu, ele, dt, w_max):
        """apply_vector_indexes :600-762 for ONE cell holding all of neu / ele (lists of [x y z u v w mpw], modified in place).
        Returns (candidates, collisions, new ions, new electrons, split-off neutrals, largest W*sigma*v_rel sampled)."""
        np_n0, np_e = len(neu), len(ele)
        np_n = np_n0
        frac = np_n * np_e * w_max * dt * self.inv_dv                  # :646
        n_groups = int(frac + 0.5)
        if n_groups > np_n:
            n_groups = np_n - 1                                        # :649-653
        ions, new_ele, split, extra = [], [], [], []
        n_coll, step_max = 0, 0.0
        for _ in range(max(n_groups, 0)):
            a = int(next(r) * np_n)
            b = int(next(r) * np_e)
            pn = neu[a] if a < np_n0 else extra[a - np_n0]
            pe = ele[b]
            vn, ve = pn[3:6], pe[3:6]
            d = [vn[c] - ve[c] for c in range(3)]
            v_rel = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
            s_coll = self.sigma_coll(self.E_rel_eV * v_rel * v_rel)
            Wn, We = pn[6], pe[6]
            Wg, Wl = (We, Wn) if Wn < We else (Wn, We)
            wsv = Wg * s_coll * v_rel
            step_max = max(step_max, wsv)
            if next(r) < wsv / w_max:
                n_coll += 1
                if Wn > We:                                            # split the neutral (:684-703)
                    ionised, ve_new, v_new = self.collide(r, vn, ve, s_coll)
                    pn[6] = Wn - We
                    pe[3:6] = ve_new
                    if ionised:
                        ions.append(list(pn[0:3]) + list(vn) + [Wl])
                        new_ele.append(list(pn[0:3]) + list(v_new) + [Wl])
                    else:
                        q = list(pn[0:3]) + list(vn) + [We]
                        split.append(q)
                        if len(extra) < 16:                            # MCC_EXTRA: split-offs of this call that remain selectable
                            extra.append(q); np_n += 1
        return max(n_groups, 0), n_coll, ions, new_ele, split, step_max


def _sqrt(x):
    return math.sqrt(x) if x >= 0 else float("nan")


def _pow(b, e):
    try:
        return math.pow(b, e)
    except ValueError:                                                 # negative base, non-integer exponent: std::pow returns NaN
        return float("nan")


class MccModelV2(MccModel):
    """The fixed-weight ancestor: MC_MEX_Ionization of ch4/v2 (ch4/v2/Interactions.cpp:476-735).  Differences from v3: Bird's NTC
    candidate count with neutrals.mpw0 (:591), unweighted acceptance (:608-615), evaluateSigmaIon without the threshold guard
    (:560-563), collide without the `E_inc < E_ion` early-out (:678-735: the ejected electron of a sub-threshold ionisation gets a
    NaN velocity), products through Species::addParticle (:624-628), neutrals never depleted (:631)."""

    def __init__(self, m_n, m_e, E_ion_J, tab_E, tab_s, dv, mpw0_n, mpw0_i, mpw0_e):
        super().__init__(m_n, m_e, E_ion_J, tab_E, tab_s, dv, mpw0_n, mpw0_e)
        self.w_max0 = 1e-14                                            # ch4/v2/Interactions.h:129
        self.mpw0_n, self.mpw0_i, self.mpw0_e = mpw0_n, mpw0_i, mpw0_e
        self.ions_to_create = int(mpw0_e / mpw0_i + 0.5)               # :623

    def sigma_ion(self, E):                                            # :560-563, no guard
        return self.c0 * math.log(E / self.c1) / E * math.exp(-self.c2 / E) if E > 0 else float("nan")

    def new_velocity_electron(self, r, E, u):                          # :643-662
        cos_ksi = (2 + E - 2 * _pow(1 + E, next(r))) / E
        sin_ksi = _sqrt(1 - cos_ksi * cos_ksi)
        phi = 2 * PI * next(r)
        v_mag = _sqrt(E * self.two_qe_me)
        ixu = [0.0 * u[2] - 0.0 * u[1], 0.0 * u[0] - 1.0 * u[2], 1.0 * u[1] - 0.0 * u[0]]
        uxi = [u[1] * ixu[2] - u[2] * ixu[1], u[2] * ixu[0] - u[0] * ixu[2], u[0] * ixu[1] - u[1] * ixu[0]]
        sp, cp = math.sin(phi), math.cos(phi)
        return [(cos_ksi * u[c] + ixu[c] * sin_ksi * sp + uxi[c] * sin_ksi * cp) * v_mag for c in range(3)]

    def collide(self, r, vn, ve, s_coll):                              # :678-735 (IONIZE_1)
        g = [vn[c] - ve[c] for c in range(3)]
        g_mag = math.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2])
        m_r = self.m_n * self.m_e / self.sum_mass
        E_rel_J = 0.5 * m_r * g_mag * g_mag
        p_ion = self.sigma_ion(E_rel_J / QE) / s_coll
        if next(r) <= p_ion:
            ve_mag = math.sqrt(ve[0] * ve[0] + ve[1] * ve[1] + ve[2] * ve[2])
            E_inc = ve_mag * ve_mag * self.E_ele_eV
            E_ej = 10.0 * math.tan(next(r) * math.atan((E_inc - self.E_ion_eV) / (2 * self.B_inc)))
            E_sc = E_inc - self.E_ion_eV - E_ej
            if E_sc < 0:
                E_sc = 0.000001
            inv = 1.0 / ve_mag
            u = [ve[0] * inv, ve[1] * inv, ve[2] * inv]
            return True, self.new_velocity_electron(r, E_sc, u), self.new_velocity_electron(r, E_ej, u)
        inv_sum = 1.0 / self.sum_mass
        cm = [(self.m_n * vn[c] + self.m_e * ve[c]) * inv_sum for c in range(3)]
        cos_ksi = 2 * next(r) - 1
        sin_ksi = math.sqrt(1 - cos_ksi * cos_ksi)
        eps = 2 * PI * next(r)
        g = [g_mag * cos_ksi, g_mag * sin_ksi * math.cos(eps), g_mag * sin_ksi * math.sin(eps)]
        f = self.m_n / self.sum_mass
        return False, [cm[c] - f * g[c] for c in range(3)], [0.0, 0.0, 0.0]

    def apply_cell(self, r, neu, ele, dt, sv_max, add_ion, add_ele):
        """apply :566-641 for ONE cell holding all of neu / ele.  add_ion / add_ele(pos, vel, mpw) stand for Species::addParticle of the two
        product species (bounds / object filter + half-step rewind); they return the stored particle or None.
        Returns (candidates, collisions, ionisations, new ions, new electrons, largest sigma*v_rel sampled)."""
        np_n, np_e = len(neu), len(ele)
        frac = 0.5 * np_n * np_e * self.mpw0_n * sv_max * dt * self.inv_dv     # :591
        n_groups = int(frac + 0.5)
        if n_groups > np_n:
            n_groups = np_n - 1                                        # :598-600
        ions, new_ele = [], []
        n_coll, n_ion, step_max = 0, 0, 0.0
        for _ in range(max(n_groups, 0)):
            pn = neu[int(next(r) * np_n)]
            pe = ele[int(next(r) * np_e)]
            vn, ve = pn[3:6], pe[3:6]
            d = [vn[c] - ve[c] for c in range(3)]
            v_rel = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
            s_coll = self.sigma_coll(self.E_rel_eV * v_rel * v_rel)
            sv = s_coll * v_rel
            step_max = max(step_max, sv)
            if sv / sv_max > next(r):
                n_coll += 1
                ionised, ve_new, v_new = self.collide(r, vn, ve, s_coll)
                pe[3:6] = ve_new
                if ionised:
                    n_ion += 1
                    for _k in range(self.ions_to_create):
                        q = add_ion(list(pn[0:3]), list(vn), self.mpw0_i)
                        if q is not None:
                            ions.append(q)
                    q = add_ele(list(pn[0:3]), list(v_new), self.mpw0_e)
                    if q is not None:
                        new_ele.append(q)
        return max(n_groups, 0), n_coll, n_ion, ions, new_ele, step_max

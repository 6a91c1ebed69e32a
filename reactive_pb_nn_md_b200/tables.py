"""Look-up tables of the reference, built with the reference's own arithmetic (REAL*4 where the
Fortran source is REAL*4).  SURVEY.md Appendix A.3, A.6, A.8, A.9.

  B_spline            src/pme.f90:505-532       (REAL*4 recursion storage `real,dimension(n-1,n-1)::mn`)
  B6/B5 tables        src/initialize_routines.f90:212-217
  erfc / ewaldscale   src/initialize_routines.f90:229-242
  Tang-Toennies       src/initialize_routines.f90:248-264, src/general_routines.f90:1782-1815
  CB_array / bm_sq    src/pme.f90:537-595       (pi = 3.14159265 REAL*4; cmplx(cos,sin) default kind)
"""
import math

import numpy as np
from scipy.special import erfc as _erfc

PI = 3.141592654                       # constants%pi        glob_v.f90:386
PI_SQRT = 1.772453851                  # constants%pi_sqrt   glob_v.f90:387
CONV_E2A_KJMOL = float(np.float32(1389.35465))   # REAL*4 literal, glob_v.f90:389
CONV_KJMOL = 100.0                     # glob_v.f90:388
BOLTZMANN = 0.008314462                # glob_v.f90:390
SAFE_VERLET = float(np.float32(1.2))   # glob_v.f90:393
VERLET_THRESH = 1.2                    # glob_v.f90:394
SPLINE_GRID = 100000                   # glob_v.f90:397
ERFC_GRID = 100000                     # glob_v.f90:398
TT_MAX = 50.0                          # glob_v.f90:347
TT_GRID = 1000                         # glob_v.f90:348
PI32 = float(np.float32(3.14159265))   # `real*8, parameter :: pi=3.14159265` pme.f90:542


def b_spline(u, n):
    """B_spline(u,n) for an array of arguments u (float64); mn is stored in float32."""
    u = np.asarray(u, np.float64)
    mn = np.zeros((n - 1, n - 1) + u.shape, np.float32)
    for i in range(1, n):
        ui = u - float(i - 1)
        val = 1.0 - np.abs(ui - 1.0)
        mn[0, i - 1] = np.where((ui < 0.0) | (ui > 2.0), 0.0, val).astype(np.float32)
    for j in range(2, n):
        for i in range(1, n - j + 1):
            ui = u - float(i - 1)
            rhs = (ui / float(j)) * mn[j - 2, i - 1].astype(np.float64) + \
                  ((float(j + 1) - ui) / float(j)) * mn[j - 2, i].astype(np.float64)
            mn[j - 1, i - 1] = rhs.astype(np.float32)
    return mn[n - 2, 0].astype(np.float64)


def spline_tables(spline_grid=SPLINE_GRID):
    i = np.arange(1, spline_grid + 1, dtype=np.float64)
    B6 = b_spline(6.0 / float(spline_grid) * i, 6)
    B5 = b_spline(5.0 / float(spline_grid) * i, 5)
    return B6, B5


def ewald_tables(real_space_cutoff, alpha_sqrt, erfc_grid=ERFC_GRID):
    erfc_dx = real_space_cutoff / float(erfc_grid)
    i = np.arange(1, erfc_grid + 2, dtype=np.float64)
    r = i * erfc_dx
    x = r * alpha_sqrt
    erfc_table = _erfc(x) * CONV_E2A_KJMOL
    ewaldscale = erfc_table + x * 2.0 / PI_SQRT * np.exp(-x * x) * CONV_E2A_KJMOL
    return erfc_dx, erfc_table, ewaldscale


def tang_toennies_tables(tt_max=TT_MAX, tt_grid=TT_GRID):
    factorial = np.ones(13)
    for k in range(2, 13):
        factorial[k] = float(k) * factorial[k - 1]
    x = tt_max / float(tt_grid) * np.arange(1, tt_grid + 1, dtype=np.float64)
    tt = np.zeros((4, tt_grid), order="F")
    dtt = np.zeros((4, tt_grid), order="F")
    for q, n in enumerate((6, 8, 10, 12)):
        s = np.ones_like(x)
        xn = np.ones_like(x)
        for k in range(1, n + 1):
            xn = xn * x
            s = s + xn / factorial[k]
        tt[q] = 1.0 - s * np.exp(-x)
        dtt[q] = np.exp(-x) * x ** n / factorial[n]
    return tt, dtt


def bm_sq(K, n=6):
    """|b(m)|^2, m=0..K-1 (pme.f90:578-595): float32 phase factors, complex128 accumulation."""
    bs = b_spline(np.arange(1, n, dtype=np.float64), n)        # B_spline(dble(i+1),n), i=0..n-2
    out = np.zeros(K)
    for m in range(K):
        s = 0.0 + 0.0j
        for i in range(n - 1):
            tmp = 2.0 * PI * float(m * i) / float(K)
            ph = complex(float(np.float32(math.cos(tmp))), float(np.float32(math.sin(tmp))))
            s = s + bs[i] * ph
        bm = 1.0 / s
        out[m] = bm.real * bm.real + bm.imag * bm.imag
    return out


def cb_array(box_length, K, alpha_sqrt, n=6):
    """CB(K,K,K) column-major (pme.f90:537-573) for the cubic box the reference assumes (pme.f90:646)."""
    vol = box_length * box_length * box_length        # box(1,1)**3
    kkd = (box_length * box_length) / vol         # kk(l,l) = (b x c)_l / volume
    idx = np.arange(K)
    m = np.where(idx > K // 2, idx - K, idx).astype(np.float64)
    mm = m * kkd
    b = bm_sq(K, n)
    m1 = mm[:, None, None]; m2 = mm[None, :, None]; m3 = mm[None, None, :]
    mag = m1 * m1 + m2 * m2 + m3 * m3
    with np.errstate(divide="ignore", invalid="ignore"):
        CB = 1.0 / (vol * PI32) * np.exp(-(PI32 * PI32) * mag / (alpha_sqrt * alpha_sqrt)) / mag
        CB = CB * b[:, None, None] * b[None, :, None] * b[None, None, :]
    CB[0, 0, 0] = 0.0
    return np.asfortranarray(CB)


def ewald_self(charge, alpha_sqrt):
    """update_Ewald_self pme.f90:608-625 (sequential accumulation)."""
    s = 0.0
    for q in np.asarray(charge, np.float64):
        s = s - q * q
    return s * alpha_sqrt / PI_SQRT * CONV_E2A_KJMOL

"""ORACLE (test infrastructure, never imported by the product path): numpy restatements of the
reference's spherical-harmonics and frequency encoders and of ``trunc_exp``.

SH (shencoder/src/shencoder.cu:43-121 values, :130-350 Jacobian).  The reference hard-codes 64
polynomials in x, y, z.  They are the real spherical harmonics with the Condon-Shortley phase,
written as  N_l^|m| * Q_l^|m|(z) * {Re, Im}(x + i y)^|m|  with  Q_l^m = d^m P_l / dz^m  — e.g.
``outputs[6] = 0.946 z^2 - 0.315`` is N_2^0 * P_2(z) (shencoder.cu:58).  This oracle evaluates that
closed form in fp64 with numpy's Legendre class (a different route than the product kernel's
recurrences); ``oracle/make_golden.py`` checks it — values AND the x/y/z partial derivatives —
against the polynomials parsed from the reference source text, and stores those as golden vectors.

Freq (freqencoder/src/freqencoder.cu:30-94), trunc_exp (activation.py:5-18).
"""
from __future__ import annotations

import math

import numpy as np
from numpy.polynomial import legendre as _leg
from numpy.polynomial import polynomial as _poly


def sh_norm(l, m):
    """sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) times sqrt(2)(-1)^m for m > 0."""
    k = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))
    return k * (math.sqrt(2) * (-1) ** m if m > 0 else 1.0)


def _q_poly(l, m):
    """Power-series coefficients of Q_l^m(z) = d^m/dz^m P_l(z)."""
    c = _leg.leg2poly([0] * l + [1])
    for _ in range(m):
        c = _poly.polyder(c)
    return c


def sh_encode(dirs, degree, want_jacobian=False):
    """dirs [B,3] (already normalised by SHEncoder.forward, sphere_harmonics.py:79-82) ->
    out [B, degree^2] fp64; optional jacobian [B, 3, degree^2] treating x,y,z as independent
    variables, like the reference's dx/dy/dz tables."""
    d = np.asarray(dirs, dtype=np.float64)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    B = d.shape[0]
    n = degree * degree
    out = np.zeros((B, n))
    jac = np.zeros((B, 3, n)) if want_jacobian else None
    xy = x + 1j * y
    for l in range(degree):
        centre = l * l + l
        for m in range(0, l + 1):
            q = _poly.polyval(z, _q_poly(l, m))
            q1 = _poly.polyval(z, _q_poly(l, m + 1)) if m + 1 <= l else np.zeros_like(z)
            N = sh_norm(l, m)
            if m == 0:
                out[:, centre] = N * q
                if want_jacobian:
                    jac[:, 2, centre] = N * q1
                continue
            p = xy ** m
            pm1 = xy ** (m - 1)
            A, Bm = p.real, p.imag
            out[:, centre + m] = N * q * A
            out[:, centre - m] = N * q * Bm
            if want_jacobian:
                jac[:, 0, centre + m] = N * q * m * pm1.real
                jac[:, 1, centre + m] = -N * q * m * pm1.imag
                jac[:, 2, centre + m] = N * q1 * A
                jac[:, 0, centre - m] = N * q * m * pm1.imag
                jac[:, 1, centre - m] = N * q * m * pm1.real
                jac[:, 2, centre - m] = N * q1 * Bm
    if want_jacobian:
        return out, jac
    return out


def sh_backward(grad, jac):
    """kernel_sh_backward (shencoder.cu:358-382): grad_in[b,d] = sum_ch grad[b,ch] * dy_dx[b,d,ch]."""
    return np.einsum("bc,bdc->bd", np.asarray(grad, np.float64), np.asarray(jac, np.float64))


def freq_encode(x, degree):
    """kernel_freq (freqencoder.cu:30-58): [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...], each block D wide.
    cos is evaluated as sin(. + pi/2) on the device (fast-math __sinf): compare with abs+rel tolerance."""
    x = np.asarray(x, dtype=np.float64)
    parts = [x]
    for f in range(degree):
        parts.append(np.sin(x * 2.0 ** f))
        parts.append(np.sin(x * 2.0 ** f + np.float32(np.pi / 2).astype(np.float64)))
    return np.concatenate(parts, axis=-1)


def freq_backward(grad, outputs, D, degree):
    """kernel_freq_backward (freqencoder.cu:63-94), from the saved outputs."""
    g = np.asarray(grad, np.float64)
    o = np.asarray(outputs, np.float64)
    r = g[:, :D].copy()
    for f in range(degree):
        s = D + 2 * f * D
        c = s + D
        r += 2.0 ** f * (g[:, s:s + D] * o[:, c:c + D] - g[:, c:c + D] * o[:, s:s + D])
    return r


def trunc_exp_forward(x):
    """activation.py:10."""
    return np.exp(np.asarray(x, np.float32)).astype(np.float32)


def trunc_exp_backward(g, x):
    """activation.py:16."""
    x = np.asarray(x, np.float32)
    return (np.asarray(g, np.float32) * np.exp(np.clip(x, -15, 15))).astype(np.float32)

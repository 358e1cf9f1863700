"""Seeded synthetic initial conditions for the hot-path tests and bench.

The reference ships no IC file (parameter.txt:5 names a URL); these generators make the
configurations BASELINE.json names: isolated Hernquist / NFW haloes (C1-C3, C5) and a
periodic box (C4).  Units are the sample parameter file's (parameter.txt:147-151):
kpc, 1e10 Msun, km/s, so G = 43007.1.

Velocities come from the isotropic Jeans equation (Gaussian with the local radial
dispersion, clipped below the escape speed): close enough to equilibrium for a per-step
throughput/parity workload; nothing on the hot path depends on exact equilibrium.
"""
from __future__ import annotations

import numpy as np

G_INTERNAL = 43007.1  # kpc (km/s)^2 / 1e10 Msun


def _halo(rho_fn, rmin, rmax, n, seed, G=G_INTERNAL, total_mass=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    ngrid = 4096
    r = np.geomspace(rmin, rmax, ngrid)
    rho = rho_fn(r)
    # enclosed mass (trapezoid in ln r), plus the analytic r->0 piece assuming rho ~ 1/r
    integrand = 4.0 * np.pi * rho * r ** 3
    lnr = np.log(r)
    m_enc = np.concatenate(([0.0], np.cumsum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(lnr))))
    m_enc += 2.0 * np.pi * rho[0] * r[0] ** 3
    mtot = m_enc[-1]
    # Jeans: rho sigma^2 (r) = int_r^rmax rho G M / r'^2 dr'
    f = rho * G * m_enc / r ** 2 * r  # d ln r integrand
    tail = np.concatenate((np.cumsum((0.5 * (f[1:] + f[:-1]) * np.diff(lnr))[::-1])[::-1], [0.0]))
    sig2 = tail / rho
    # potential for the escape speed: phi(r) = -G M/r - int_r^rmax G dM/r'
    g = 4.0 * np.pi * G * rho * r ** 2  # d ln r integrand of int G dM / r'
    outer = np.concatenate((np.cumsum((0.5 * (g[1:] + g[:-1]) * np.diff(lnr))[::-1])[::-1], [0.0]))
    phi = -G * m_enc / r - outer
    vesc = np.sqrt(np.maximum(-2.0 * phi, 0.0))

    u = rng.random(n)
    rr = np.interp(u * mtot, m_enc, r)
    cth = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    sth = np.sqrt(1.0 - cth * cth)
    pos = np.stack((rr * sth * np.cos(ph), rr * sth * np.sin(ph), rr * cth), axis=1)

    sig = np.sqrt(np.interp(rr, r, sig2))
    ve = np.interp(rr, r, vesc)
    vel = rng.standard_normal((n, 3)) * sig[:, None]
    for _ in range(64):  # redraw the few that exceed 0.95 v_esc
        vmag = np.sqrt((vel * vel).sum(axis=1))
        bad = vmag > 0.95 * ve
        if not bad.any():
            break
        vel[bad] = rng.standard_normal((int(bad.sum()), 3)) * sig[bad, None]
    vmag = np.sqrt((vel * vel).sum(axis=1))
    bad = vmag > 0.95 * ve
    vel[bad] *= (0.9 * ve[bad] / vmag[bad])[:, None]

    m_part = (total_mass if total_mass is not None else mtot) / n
    mass = np.full(n, m_part, dtype=np.float32)
    ids = np.arange(1, n + 1, dtype=np.int32)
    return pos.astype(np.float32), vel.astype(np.float32), mass, ids


def hernquist(n, seed=1, M=1.0, a=10.0, rmax_over_a=100.0, G=G_INTERNAL):
    """C1: Hernquist halo, M in 1e10 Msun, scale a in kpc, truncated at rmax_over_a * a."""
    rho = lambda r: M * a / (2.0 * np.pi * r * (r + a) ** 3)
    return _halo(rho, 1e-4 * a, rmax_over_a * a, n, seed, G)


def nfw(n, seed=2, rho0=1.49e-4, rs=11.14356, rmax_over_rs=100.0, G=G_INTERNAL):
    """C2/C3/C5: the NFW halo of parameter.txt:6-11."""
    rho = lambda r: rho0 / ((r / rs) * (1.0 + r / rs) ** 2)
    return _halo(rho, 1e-4 * rs, rmax_over_rs * rs, n, seed, G)


def periodic_box(ngrid, seed=4, box=100.0, sigma_disp_cells=0.2, total_mass=1.0, vel_sigma=0.0):
    """C4: ngrid^3 equal-mass particles on a grid + Gaussian displacements, wrapped to [0, box)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = ngrid ** 3
    cell = box / ngrid
    g = (np.arange(ngrid, dtype=np.float64) + 0.5) * cell
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(n, 3)
    pos += rng.standard_normal((n, 3)) * (sigma_disp_cells * cell)
    pos = np.mod(pos, box)
    pos32 = pos.astype(np.float32)
    pos32[pos32 >= np.float32(box)] = 0.0
    vel = (rng.standard_normal((n, 3)) * vel_sigma).astype(np.float32)
    mass = np.full(n, total_mass / n, dtype=np.float32)
    ids = np.arange(1, n + 1, dtype=np.int32)
    return pos32, vel, mass, ids


def cross_section_internal(sigma_cm2_per_g, unit_mass_g=1.989e43, unit_length_cm=3.085678e21):
    """sigma/m in cm^2/g -> internal units (begrun.c set_units: sigma * UnitMass / UnitLength^2)."""
    return sigma_cm2_per_g * unit_mass_g / unit_length_cm ** 2

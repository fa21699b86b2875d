"""Synthetic inputs in the reference's .state layout (plane[i, j]: i = x index, j = y index, j contiguous).

`orszag_tang` is the "OT-N" workload of SURVEY.md 8(d): a doubly periodic Orszag-Tang vortex in coronal
CGS units on a NON-uniform rectilinear grid, analytic (no RNG).  `zfull=True` switches on non-zero
z-components / external field / gravity so that every force term of the ideal-MHD right-hand side
(reference: source/equationsets/idealmhd.cpp:51-86) is exercised by the parity tests.

`stratified_loop` is a small stand-in for the reference's example.state: gravity-stratified atmosphere with
a bipolar external field, used with non-periodic boundaries (fixed / reflect / open).
"""
from __future__ import annotations

import numpy as np

K_B = 1.3807e-16      # reference source/constants.hpp:8
M_I = 1.6726e-24
GAMMA = 5.0 / 3.0


def stretched_spacing(n: int, length: float, amp: float = 0.2) -> np.ndarray:
    i = np.arange(n, dtype=np.float64)
    d = (length / n) * (1.0 + amp * np.sin(2.0 * np.pi * (i + 0.5) / n))
    return d * (length / d.sum())


def centres(d: np.ndarray) -> np.ndarray:
    # same recurrence as the reference's convertCellSizesToCellPositions (plasmadomain.cpp:169-181)
    pos = np.empty_like(d)
    pos[0] = 0.5 * d[0]
    for k in range(1, d.size):
        pos[k] = pos[k - 1] + 0.5 * d[k - 1] + 0.5 * d[k]
    return pos


def orszag_tang(nx: int, ny: int | None = None, *, length: float = 1.0e9, zfull: bool = False,
                temp_mod: float = 0.0, stretch: float = 0.2, rows: tuple | None = None) -> dict:
    """Returns dict(planes=..., ion_mass=..., adiabatic_index=..., dx=, dy=) with all planes the reference needs
    (7 domain grids + IdealMHD state variables).  rows=(row0, n): only those x rows of every plane (a slab of a
    decomposed domain; values identical to the same rows of the full planes); dx, dy are always the global 1-D sizes."""
    ny = nx if ny is None else ny
    dx = stretched_spacing(nx, length, stretch)
    dy = stretched_spacing(ny, length, stretch)
    px, py = centres(dx), centres(dy)
    dx_g, dy_g = dx, dy
    if rows is not None:
        px, dx = px[rows[0]:rows[0] + rows[1]], dx[rows[0]:rows[0] + rows[1]]
        nx = rows[1]
    X = np.repeat(px[:, None], ny, axis=1)
    Y = np.repeat(py[None, :], nx, axis=0)
    rho0, T0 = 1.0e-15, 1.0e6
    n0 = rho0 / M_I
    p0 = 2.0 * n0 * K_B * T0
    cs = np.sqrt(GAMMA * p0 / rho0)
    v0 = cs
    B0 = 0.6 * v0 * np.sqrt(4.0 * np.pi * rho0)
    k = 2.0 * np.pi / length
    rho = np.full((nx, ny), rho0)
    temp = T0 * (1.0 + temp_mod * np.sin(k * X) * np.sin(k * Y))
    vx = -v0 * np.sin(k * Y)
    vy = v0 * np.sin(k * X)
    z = np.zeros((nx, ny))
    P = {
        "d_x": np.repeat(dx[:, None], ny, axis=1), "d_y": np.repeat(dy[None, :], nx, axis=0),
        "pos_x": X, "pos_y": Y,
        "be_x": z.copy(), "be_y": z.copy(), "be_z": z.copy(),
        "rho": rho, "temp": temp,
        "mom_x": rho * vx, "mom_y": rho * vy, "mom_z": z.copy(),
        "bi_x": -B0 * np.sin(k * Y), "bi_y": B0 * np.sin(2.0 * k * X), "bi_z": z.copy(),
        "grav_x": z.copy(), "grav_y": z.copy(),
    }
    if zfull:
        P["rho"] = rho0 * (1.0 + 0.3 * np.cos(k * X) * np.sin(2 * k * Y))
        P["mom_x"] = P["rho"] * vx
        P["mom_y"] = P["rho"] * vy
        P["mom_z"] = P["rho"] * (0.2 * v0) * np.cos(k * X + 0.3)
        P["bi_z"] = 0.3 * B0 * np.sin(k * X + k * Y)
        P["be_x"] = 0.25 * B0 * (1.0 + 0.5 * np.cos(k * Y))
        P["be_y"] = -0.15 * B0 * (1.0 + 0.5 * np.sin(k * X))
        P["be_z"] = 0.1 * B0 * np.cos(2 * k * X) * np.cos(k * Y)
        g0 = 0.05 * cs * cs / (length / 10.0)
        P["grav_x"] = g0 * np.sin(k * X)
        P["grav_y"] = -g0 * (1.0 + 0.2 * np.cos(k * Y))
    return dict(planes=P, ion_mass=M_I, adiabatic_index=GAMMA, dx=dx_g, dy=dy_g)


def stratified_loop(nx: int, ny: int, *, length: float = 4.0e9, zfull: bool = True, bump: float = 0.0) -> dict:
    """Gravity-stratified isothermal atmosphere (y = height) with a bipolar external field and a velocity /
    field perturbation; meant for non-periodic boundaries.  `bump` adds a Gaussian temperature excess."""
    dx = stretched_spacing(nx, length, 0.15)
    dy = stretched_spacing(ny, length, 0.10)
    px, py = centres(dx), centres(dy)
    X = np.repeat(px[:, None], ny, axis=1)
    Y = np.repeat(py[None, :], nx, axis=0)
    T0 = 1.0e6
    g = 2.748e4
    H = 2.0 * K_B * T0 / (M_I * g)
    rho = 1.0e-14 * np.exp(-Y / H) + 2.0e-16
    temp = T0 * (1.0 + bump * np.exp(-(((X - 0.5 * length) / (0.12 * length)) ** 2 + ((Y - 0.4 * length) / (0.12 * length)) ** 2)))
    cs = np.sqrt(GAMMA * 2.0 * K_B * T0 / M_I)
    k = 2.0 * np.pi / length
    vx = 0.05 * cs * np.sin(k * X) * np.cos(k * Y)
    vy = 0.04 * cs * np.cos(2 * k * X) * np.sin(k * Y)
    B0 = 5.0
    l = length / 4.0
    be_x = B0 * np.cos(np.pi * (X - 0.5 * length) / (2 * l)) * np.exp(-np.pi * Y / (2 * l))
    be_y = -B0 * np.sin(np.pi * (X - 0.5 * length) / (2 * l)) * np.exp(-np.pi * Y / (2 * l))
    z = np.zeros((nx, ny))
    P = {
        "d_x": np.repeat(dx[:, None], ny, axis=1), "d_y": np.repeat(dy[None, :], nx, axis=0),
        "pos_x": X, "pos_y": Y,
        "be_x": be_x, "be_y": be_y, "be_z": z.copy(),
        "rho": rho, "temp": temp,
        "mom_x": rho * vx, "mom_y": rho * vy, "mom_z": z.copy(),
        "bi_x": 0.02 * B0 * np.sin(k * Y), "bi_y": 0.02 * B0 * np.sin(k * X), "bi_z": z.copy(),
        "grav_x": z.copy(), "grav_y": np.full((nx, ny), -g),
    }
    if zfull:
        P["mom_z"] = rho * 0.03 * cs * np.sin(k * X + 2 * k * Y)
        P["bi_z"] = 0.01 * B0 * np.cos(k * X) * np.sin(k * Y)
        P["be_z"] = 0.05 * B0 * np.exp(-Y / (2 * H)) * np.cos(k * X)
        P["grav_x"] = 0.02 * g * np.sin(k * X)
    return dict(planes=P, ion_mass=M_I, adiabatic_index=GAMMA)


M_SR = 1.455e-22          # strontium ion mass, reference source/constants.hpp:17
M_E = 9.1094e-28          # electron mass, reference source/constants.hpp:9
E_CGS = 4.80320425e-10    # reference source/constants.hpp:19


def ucnp_cloud(nx: int, ny: int, *, length: float = 1.0, n0: float = 1.0e9, sigma: float = 0.1, Te: float = 20.0, Ti: float = 1.0,
               drift: float = 0.0, bfield: float = 0.0) -> dict:
    """Two-fluid ultracold-neutral-plasma expansion probe: Gaussian Sr+ / electron cloud on a non-uniform
    grid, quasi-neutral, optional initial drift and a weak magnetic field so that every two-fluid term is exercised.
    Planes = the 7 domain grids + Ideal2F::state_variables() (reference source/equationsets/ideal2F.hpp:40-42)."""
    dx = stretched_spacing(nx, length, 0.15)
    dy = stretched_spacing(ny, length, 0.10)
    px, py = centres(dx) - 0.5 * length, centres(dy) - 0.5 * length
    X = np.repeat(px[:, None], ny, axis=1)
    Y = np.repeat(py[None, :], nx, axis=0)
    n = n0 * np.exp(-(X * X + Y * Y) / (2.0 * sigma * sigma)) + 1.0e-4 * n0
    z = np.zeros((nx, ny))
    vx = drift * X / sigma
    vy = -0.5 * drift * Y / sigma
    P = {
        "d_x": np.repeat(dx[:, None], ny, axis=1), "d_y": np.repeat(dy[None, :], nx, axis=0), "pos_x": X, "pos_y": Y,
        "be_x": z + bfield, "be_y": z - 0.5 * bfield, "be_z": z.copy(),
        "i_rho": n * M_SR, "e_rho": n * (1.0 + 1.0e-3 * np.sin(6.0 * X / length) * np.cos(5.0 * Y / length)) * M_E,
        "i_mom_x": n * M_SR * vx, "i_mom_y": n * M_SR * vy, "e_mom_x": n * M_E * vx * 1.1, "e_mom_y": n * M_E * vy * 0.9,
        "i_temp": z + Ti, "e_temp": z + Te,
        "bi_x": z + 0.0, "bi_y": z + 0.0, "bi_z": z + 0.3 * bfield * np.cos(3.0 * X / length),
        "E_x": 1.0e-6 * X / sigma * np.exp(-(X * X + Y * Y) / (2.0 * sigma * sigma)), "E_y": 1.0e-6 * Y / sigma * np.exp(-(X * X + Y * Y) / (2.0 * sigma * sigma)), "E_z": z.copy(),
        "grav_x": z.copy(), "grav_y": z.copy(),
    }
    return dict(planes=P, ion_mass=M_SR, adiabatic_index=GAMMA)


def two_energy(nx: int, ny: int, loop: bool = True, bump: float = 0.4):
    """An IdealMHD2E state (state variables of idealmhd2E.hpp:27-29) from the ideal-MHD generators: the same density, momenta and field with
    unequal, spatially varying ion / electron temperatures."""
    s = stratified_loop(nx, ny, bump=bump) if loop else orszag_tang(nx, ny, zfull=False)
    P = s["planes"]
    pl = {k: P[k] for k in ("d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z")}
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    pl["rho"] = P["rho"]
    pl["i_temp"] = P["temp"] * (1.0 + 0.2 * np.sin(2 * np.pi * X / nx))
    pl["e_temp"] = P["temp"] * (0.7 + 0.1 * np.cos(2 * np.pi * Y / ny))
    for k in ("mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y"):
        pl[k] = P[k]
    return dict(planes=pl, ion_mass=s["ion_mass"], adiabatic_index=s["adiabatic_index"])


def ucnp_cloud_mhd(nx: int, ny: int, *, length: float = 1.0, n0: float = 1.0e9, sigma: float = 0.1, T: float = 20.0, drift: float = 0.0) -> dict:
    """One-fluid (ideal_mhd) ultracold-plasma cloud centred on the origin, for the UCNP modules (coulomb_explosion reads the radius sqrt(x^2 + y^2),
    global_temperature diffuses temp): Gaussian Sr+ density on a non-uniform grid, smooth temperature variation, optional radial drift."""
    dx = stretched_spacing(nx, length, 0.15)
    dy = stretched_spacing(ny, length, 0.10)
    px, py = centres(dx) - 0.5 * length, centres(dy) - 0.5 * length
    X = np.repeat(px[:, None], ny, axis=1)
    Y = np.repeat(py[None, :], nx, axis=0)
    n = n0 * np.exp(-(X * X + Y * Y) / (2.0 * sigma * sigma)) + 1.0e-3 * n0
    z = np.zeros((nx, ny))
    P = {
        "d_x": np.repeat(dx[:, None], ny, axis=1), "d_y": np.repeat(dy[None, :], nx, axis=0), "pos_x": X, "pos_y": Y,
        "be_x": z.copy(), "be_y": z.copy(), "be_z": z.copy(),
        "rho": n * M_SR, "temp": T * (1.0 + 0.3 * np.cos(5.0 * X / length) * np.sin(4.0 * Y / length + 0.3)),
        "mom_x": n * M_SR * drift * X / sigma, "mom_y": -0.5 * n * M_SR * drift * Y / sigma, "mom_z": z.copy(),
        "bi_x": z.copy(), "bi_y": z.copy(), "bi_z": z.copy(), "grav_x": z.copy(), "grav_y": z.copy(),
    }
    return dict(planes=P, ion_mass=M_SR, adiabatic_index=GAMMA)


def ucnp_cloud_2e(nx: int, ny: int, *, length: float = 1.0, n0: float = 1.0e9, sigma: float = 0.1, Te: float = 20.0, Ti: float = 1.0, drift: float = 0.0,
                  bfield: float = 0.0) -> dict:
    """The UCNP configuration of the one-fluid set with two temperatures (`ideal_mhd_2E` + `eic_thermalization`): the Gaussian Sr+ cloud of
    ucnp_cloud_mhd with hot electrons and cold ions, both temperatures smoothly varying so that the collisional exchange differs from cell to cell.
    Planes = the 7 domain grids + IdealMHD2E::state_variables() (reference source/equationsets/idealmhd2E.hpp:27-29)."""
    s = ucnp_cloud_mhd(nx, ny, length=length, n0=n0, sigma=sigma, drift=drift)
    P = s["planes"]
    X, Y = P["pos_x"], P["pos_y"]
    pl = {k: P[k] for k in ("d_x", "d_y", "pos_x", "pos_y")}
    z = np.zeros((nx, ny))
    pl["be_x"], pl["be_y"], pl["be_z"] = z + bfield, z - 0.5 * bfield, z.copy()
    pl["rho"] = P["rho"]
    pl["i_temp"] = Ti * (1.0 + 0.2 * np.sin(4.0 * X / length + 0.1) * np.cos(3.0 * Y / length))
    pl["e_temp"] = Te * (1.0 + 0.3 * np.cos(5.0 * X / length) * np.sin(4.0 * Y / length + 0.3))
    pl["mom_x"], pl["mom_y"] = P["mom_x"], P["mom_y"]
    pl["bi_x"], pl["bi_y"] = z + 0.0, 0.2 * bfield * np.cos(3.0 * X / length)
    pl["grav_x"], pl["grav_y"] = z.copy(), z.copy()
    return dict(planes=pl, ion_mass=M_SR, adiabatic_index=GAMMA)

"""Seeded synthetic workloads (SURVEY 8d / BASELINE.md): S200 and scaled variants.

S200: nx=ny=202, nz=9 (0..80 km), origin (35 N, 100 E), 0.05 deg spacing, sublayers 3,
periods 5..40 s step 5 (8 periods); Vs = depth ramp 3.0->4.5 km/s x (1 + 6% checkerboard of
8x8 cells, sign alternating with depth) + 1% Gaussian random field (corr. length 10 nodes,
default_rng(20260101)), clipped to [2.5, 4.8]; 1000 stations uniform in the inner 90% of the
box (default_rng(12345)); every station is a source at every period, its receivers are the 32
stations that follow it in a per-period permutation (default_rng(777+period)): 8000 eikonal
solves, 256 000 rays.  Gc, Gs = 0.03 sin/cos checkerboards.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .formats import F32, PI32, Survey, delsph32


@dataclasses.dataclass
class Workload:
    name: str
    nx: int
    ny: int
    nz: int
    goxd: float
    gozd: float
    dvxd: float
    dvzd: float
    sublayers: float
    depz: np.ndarray
    tRc: np.ndarray
    vs: np.ndarray        # (nx,ny,nz) F-order float32
    gc: np.ndarray        # (nx-2,ny-2,nz-1)
    gs: np.ndarray
    sv: Survey

    @property
    def n_solves(self) -> int:
        return int(self.sv.nsrcsurf1.sum())

    @property
    def n_rays(self) -> int:
        return int(self.sv.dall)

    @property
    def nodes_coarse(self) -> int:
        return ((self.nx - 3) * 5 + 1) * ((self.ny - 3) * 5 + 1)


def _grf(shape, corr, rng):
    w = rng.standard_normal(shape)
    kx = np.fft.fftfreq(shape[0])[:, None]
    ky = np.fft.fftfreq(shape[1])[None, :]
    filt = np.exp(-0.5 * (2 * np.pi * corr) ** 2 * (kx ** 2 + ky ** 2) / 4.0)
    f = np.fft.ifft2(np.fft.fft2(w) * filt).real
    return f / f.std()


def make_model(nx, ny, nz, dz_km=10.0, cells=8, seed=20260101):
    rng = np.random.default_rng(seed)
    depz = (np.arange(nz) * dz_km).astype(F32)
    ramp = np.linspace(3.0, 4.5, nz)
    ii = (np.arange(nx) * cells // nx)[:, None]
    jj = (np.arange(ny) * cells // ny)[None, :]
    chk = np.where((ii + jj) % 2 == 0, 1.0, -1.0)
    vs = np.zeros((nx, ny, nz), np.float64)
    for k in range(nz):
        g = _grf((nx, ny), 10.0, rng)
        vs[:, :, k] = ramp[k] * (1.0 + 0.06 * chk * (1 if k % 2 == 0 else -1)) * (1.0 + 0.01 * g)
    vs = np.clip(vs, 2.5, 4.8)
    x = np.arange(nx - 2)[:, None, None]
    y = np.arange(ny - 2)[None, :, None]
    k = np.arange(nz - 1)[None, None, :]
    gc = 0.03 * np.sin(2 * np.pi * x * cells / (2.0 * (nx - 2))) * np.cos(2 * np.pi * y * cells / (2.0 * (ny - 2))) * (1 - 2 * (k % 2))
    gs = 0.03 * np.cos(2 * np.pi * x * cells / (2.0 * (nx - 2))) * np.sin(2 * np.pi * y * cells / (2.0 * (ny - 2))) * (1 - 2 * (k % 2))
    return depz, np.asfortranarray(vs.astype(F32)), np.asfortranarray(gc.astype(F32)), np.asfortranarray(gs.astype(F32))


def make_survey(nx, ny, goxd, gozd, dvxd, dvzd, kmax, nsta, nrec, src_per_period=None, seed=12345):
    rng = np.random.default_rng(seed)
    # the propagation grid spans the interior control nodes: (goxd, gozd) is interior node 1
    # (FwdAzimuthalAniMap.f90:79), so lat in [goxd-(nx-3)*dvxd, goxd], lon in [gozd, gozd+(ny-3)*dvzd]
    lat_hi = goxd; lat_lo = goxd - (nx - 3) * dvxd
    lon_lo = gozd; lon_hi = gozd + (ny - 3) * dvzd
    m_lat = 0.05 * (lat_hi - lat_lo); m_lon = 0.05 * (lon_hi - lon_lo)
    lat = rng.uniform(lat_lo + m_lat, lat_hi - m_lat, nsta).astype(F32)
    lon = rng.uniform(lon_lo + m_lon, lon_hi - m_lon, nsta).astype(F32)
    colat = ((F32(90.0) - lat) * PI32 / F32(180.0)).astype(F32)
    lonr = (lon * PI32 / F32(180.0)).astype(F32)
    ns = nsta if src_per_period is None else min(nsta, src_per_period)
    periods = np.zeros((ns, kmax), np.int32, order="F"); nrc1 = np.zeros((ns, kmax), np.int32, order="F")
    scxf = np.zeros((ns, kmax), F32, order="F"); sczf = np.zeros((ns, kmax), F32, order="F")
    rcxf = np.zeros((nrec, ns, kmax), F32, order="F"); rczf = np.zeros((nrec, ns, kmax), F32, order="F")
    nsrcsurf1 = np.full(kmax, ns, np.int32)
    for k in range(kmax):
        perm = np.random.default_rng(777 + k + 1).permutation(nsta)
        pos = np.empty(nsta, np.int64); pos[perm] = np.arange(nsta)
        for s in range(ns):
            periods[s, k] = k + 1
            nrc1[s, k] = nrec
            scxf[s, k] = colat[s]; sczf[s, k] = lonr[s]
            idx = perm[(pos[s] + 1 + np.arange(nrec)) % nsta]
            rcxf[:, s, k] = colat[idx]; rczf[:, s, k] = lonr[idx]
    dall = int(nrc1.sum())
    wave = np.full((ns, kmax), 2, np.int32, order="F"); igrt = np.zeros((ns, kmax), np.int32, order="F")
    sv = Survey(kmax, ns, nrec, periods, nrc1, nsrcsurf1, scxf, sczf, rcxf, rczf, wave, igrt,
                np.zeros(dall, F32), np.zeros(dall, F32), dall)
    return sv


def s200(src_per_period=None, kmax=8, n=202, nz=9, nsta=1000, nrec=32) -> Workload:
    goxd, gozd, dv = 35.0, 100.0, 0.05
    depz, vs, gc, gs = make_model(n, n, nz)
    tRc = (5.0 * (1 + np.arange(kmax))).astype(np.float64)
    sv = make_survey(n, n, goxd, gozd, dv, dv, kmax, nsta, nrec, src_per_period)
    return Workload("S200" if src_per_period is None and n == 202 else f"S{n-2}-{src_per_period or nsta}src", n, n, nz, goxd,
                    gozd, dv, dv, 3.0, depz, tRc, vs, gc, gs, sv)


def proxy_tables(w: Workload, seed=7):
    """Smooth stand-in depth-kernel tables with realistic magnitudes (used only to exercise the
    eikonal/ray/assembly stages in isolation, e.g. by profiles; never by parity tests)."""
    nxy = w.nx * w.ny
    k = len(w.tRc)
    z = np.asarray(w.depz, np.float64)
    pv = np.zeros((nxy, k), np.float64, order="F")
    vsf = np.asarray(w.vs, np.float64).reshape(nxy, w.nz, order="F")
    for ip, T in enumerate(w.tRc):
        wgt = np.exp(-0.5 * ((z - 1.1 * T) / (0.6 * T + 5.0)) ** 2) + 1e-3
        wgt /= wgt.sum()
        pv[:, ip] = np.float32(0.92 * (vsf * wgt[None, :]).sum(1))
    rng = np.random.default_rng(seed)
    sen = [np.asfortranarray(0.05 + 0.2 * rng.random((nxy, k, w.nz))) for _ in range(3)]
    L = np.asfortranarray((0.3 * rng.random((nxy, k, w.nz - 1)) + 0.05).astype(F32))
    return dict(pvRc=pv, sen_vs=sen[0], sen_vp=sen[1], sen_rho=sen[2], Lsen_Gsc=L)


def yunnan_shaped(nsta=300, src_per_period=None, nrec=None, kmax=36, seed=4242) -> Workload:
    """BASELINE config 5 shape: the grid, depth nodes, spacing, periods and sub-layering of
    example/test4_Yunnan (para.in: 38 42 18; 29.0 98.0; 0.25 0.25; sublayers 4; 36 periods 5..40 s;
    MOD line 1 depths) with a seeded synthetic model and station geometry (the real data file is not
    redistributed).  Every station is a source at every period; receivers = the stations that
    follow it in the station list (all pairs once), optionally capped at nrec."""
    nx, ny, nz = 38, 42, 18
    goxd, gozd, dv = 29.0, 98.0, 0.25
    depz = np.array([0, 5, 10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 70, 80, 90, 100, 120], F32)
    _, vs, gc, gs = make_model(nx, ny, nz, cells=6, seed=seed)
    # crustal/upper-mantle ramp on the real depth nodes (make_model's ramp is per index)
    tRc = (5.0 + np.arange(kmax)).astype(np.float64)
    rng = np.random.default_rng(seed)
    lat_hi = goxd; lat_lo = goxd - (nx - 3) * dv
    lon_lo = gozd; lon_hi = gozd + (ny - 3) * dv
    lat = rng.uniform(lat_lo + 0.4, lat_hi - 0.4, nsta).astype(F32)
    lon = rng.uniform(lon_lo + 0.4, lon_hi - 0.4, nsta).astype(F32)
    colat = ((F32(90.0) - lat) * PI32 / F32(180.0)).astype(F32)
    lonr = (lon * PI32 / F32(180.0)).astype(F32)
    ns = nsta - 1 if src_per_period is None else min(nsta - 1, src_per_period)
    nrcf = (nsta - 1) if nrec is None else nrec
    periods = np.zeros((ns, kmax), np.int32, order="F"); nrc1 = np.zeros((ns, kmax), np.int32, order="F")
    scxf = np.zeros((ns, kmax), F32, order="F"); sczf = np.zeros((ns, kmax), F32, order="F")
    rcxf = np.zeros((nrcf, ns, kmax), F32, order="F"); rczf = np.zeros((nrcf, ns, kmax), F32, order="F")
    for k in range(kmax):
        for s in range(ns):
            n = min(nsta - 1 - s, nrcf)
            periods[s, k] = k + 1; nrc1[s, k] = n
            scxf[s, k] = colat[s]; sczf[s, k] = lonr[s]
            rcxf[:n, s, k] = colat[s + 1:s + 1 + n]; rczf[:n, s, k] = lonr[s + 1:s + 1 + n]
    dall = int(nrc1.sum())
    wave = np.full((ns, kmax), 2, np.int32, order="F"); igrt = np.zeros((ns, kmax), np.int32, order="F")
    sv = Survey(kmax, ns, nrcf, periods, nrc1, np.full(kmax, ns, np.int32), scxf, sczf, rcxf, rczf, wave, igrt,
                np.zeros(dall, F32), np.zeros(dall, F32), dall)
    return Workload(f"YN-{nsta}sta" + ("" if src_per_period is None else f"-{src_per_period}src"), nx, ny, nz, goxd, gozd,
                    dv, dv, 4.0, depz, tRc, vs, gc, gs, sv)


def t1_shaped(nsta=121, kmax=36, seed=99) -> Workload:
    """BASELINE config 1 shape (example/test1_syn_foward): 17 x 17 x 4 model (71 x 71 propagation grid, refined
    source boxes that cover most of it), 36 periods 5..40 s, sublayers 2, ~120 sources per period with all other
    stations as receivers -- with a seeded synthetic model and geometry (the shipped files stay in /root/reference)."""
    nx = ny = 17; nz = 4
    goxd, gozd, dv = 26.5, 101.0, 0.25          # any origin works; spacing like the example
    depz = np.array([0, 10, 35, 60], F32)
    _, vs, gc, gs = make_model(nx, ny, nz, cells=4, seed=seed)
    rng = np.random.default_rng(seed)
    lat = rng.uniform(goxd - (nx - 3) * dv + 0.2, goxd - 0.2, nsta).astype(F32)
    lon = rng.uniform(gozd + 0.2, gozd + (ny - 3) * dv - 0.2, nsta).astype(F32)
    colat = ((F32(90.0) - lat) * PI32 / F32(180.0)).astype(F32)
    lonr = (lon * PI32 / F32(180.0)).astype(F32)
    ns = nsta - 1
    nrcf = nsta - 1
    periods = np.zeros((ns, kmax), np.int32, order="F"); nrc1 = np.zeros((ns, kmax), np.int32, order="F")
    scxf = np.zeros((ns, kmax), F32, order="F"); sczf = np.zeros((ns, kmax), F32, order="F")
    rcxf = np.zeros((nrcf, ns, kmax), F32, order="F"); rczf = np.zeros((nrcf, ns, kmax), F32, order="F")
    for k in range(kmax):
        for s in range(ns):
            n = nsta - 1 - s
            periods[s, k] = k + 1; nrc1[s, k] = n
            scxf[s, k] = colat[s]; sczf[s, k] = lonr[s]
            rcxf[:n, s, k] = colat[s + 1:]; rczf[:n, s, k] = lonr[s + 1:]
    dall = int(nrc1.sum())
    wave = np.full((ns, kmax), 2, np.int32, order="F"); igrt = np.zeros((ns, kmax), np.int32, order="F")
    sv = Survey(kmax, ns, nrcf, periods, nrc1, np.full(kmax, ns, np.int32), scxf, sczf, rcxf, rczf, wave, igrt,
                np.zeros(dall, F32), np.zeros(dall, F32), dall)
    tRc = (5.0 + np.arange(kmax)).astype(np.float64)
    return Workload(f"T1-{nsta}sta", nx, ny, nz, goxd, gozd, dv, dv, 2.0, depz, tRc, vs, gc, gs, sv)

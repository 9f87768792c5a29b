"""Readers/writers for the reference's text formats (host logic, no compute).

Reference parsers mirrored here (all paths relative to /root/reference):
  * forward ``para.in``      src/src_forward/MainForward.f90:147-159,188,330
  * inversion ``para.in``    src/src_inv_iso_joint/Main_Jt.f90:158-211
  * ``MOD`` / ``MODVs.true`` src/src_forward/MainForward.f90:338-343
  * ``MODGc/Gs.true``        src/src_forward/MainForward.f90:351-356
  * surface-wave data file   src/src_forward/MainForward.f90:239-281
  * ``surfphase_forward.dat``src/src_forward/MainForward.f90:403-429
  * ``period_Azm_tomo.*``    src/src_forward/FwdAzimuthalAniMap.f90:41-79

Every REAL of the reference is float32 here, including the float32 value of
pi (``real,parameter :: pi=3.1415926535898``, MainForward.f90:47): that is why
an input latitude of 23.3 is echoed as 23.300011 in the reference's output.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional

import numpy as np

F32 = np.float32
PI32 = F32(3.1415926535898)


@dataclasses.dataclass
class ForwardPara:
    datafile: str
    nx: int
    ny: int
    nz: int
    goxd: float
    gozd: float
    dvxd: float
    dvzd: float
    nsrc: int
    sublayers: float
    spfra: float
    writepath: bool
    kmaxRc: int
    tRc: np.ndarray  # float64 (kmaxRc,)
    noiselevel: float = 0.0


@dataclasses.dataclass
class InvPara:
    datafile: str
    nx: int
    ny: int
    nz: int
    goxd: float
    gozd: float
    dvxd: float
    dvzd: float
    sublayers: float
    minvel: float
    maxvel: float
    nsrc: int
    spfra: float
    maxiter: int
    iso_mod: bool
    weightVs: float
    weightGcs: float
    damp: float
    kmaxRc: int
    tRc: np.ndarray


def _tok(line: str) -> List[str]:
    """List-directed read of one record: blank/comma separated, stops at 'c:' comments."""
    return line.replace(",", " ").split()


def _logical(tok: str) -> bool:
    t = tok.strip().strip(".").upper()
    return t.startswith("T")


def read_para_forward(path: str) -> ForwardPara:
    with open(path) as f:
        lines = f.read().splitlines()
    it = iter(lines[3:])  # three header lines are skipped (MainForward.f90:147-149)
    datafile = _tok(next(it))[0].strip("'\"")
    nx, ny, nz = (int(t) for t in _tok(next(it))[:3])
    goxd, gozd = (float(t) for t in _tok(next(it))[:2])
    dvxd, dvzd = (float(t) for t in _tok(next(it))[:2])
    nsrc = int(_tok(next(it))[0])
    sublayers = float(_tok(next(it))[0])
    spfra = float(_tok(next(it))[0])
    writepath = _logical(_tok(next(it))[0])
    kmaxRc = int(_tok(next(it))[0])
    toks = _tok(next(it))
    tRc = np.array([float(t) for t in toks[:kmaxRc]], dtype=np.float64)
    try:
        noise = float(_tok(next(it))[0])
    except StopIteration:
        noise = 0.0
    return ForwardPara(datafile, nx, ny, nz, goxd, gozd, dvxd, dvzd, nsrc, sublayers, spfra,
                       writepath, kmaxRc, tRc, noise)


def read_para_inv(path: str) -> InvPara:
    with open(path) as f:
        lines = f.read().splitlines()
    it = iter(lines[3:])
    datafile = _tok(next(it))[0].strip("'\"")
    nx, ny, nz = (int(t) for t in _tok(next(it))[:3])
    goxd, gozd = (float(t) for t in _tok(next(it))[:2])
    dvxd, dvzd = (float(t) for t in _tok(next(it))[:2])
    sublayers = float(_tok(next(it))[0])
    minvel, maxvel = (float(t) for t in _tok(next(it))[:2])
    nsrc = int(_tok(next(it))[0])
    spfra = float(_tok(next(it))[0])
    maxiter = int(_tok(next(it))[0])
    iso_mod = _logical(_tok(next(it))[0])
    next(it)  # comment line
    weightVs = float(_tok(next(it))[0])
    weightGcs = float(_tok(next(it))[0])
    damp = float(_tok(next(it))[0])
    next(it)  # comment line
    kmaxRc = int(_tok(next(it))[0])
    tRc = np.array([float(t) for t in _tok(next(it))[:kmaxRc]], dtype=np.float64)
    return InvPara(datafile, nx, ny, nz, goxd, gozd, dvxd, dvzd, sublayers, minvel, maxvel, nsrc,
                   spfra, maxiter, iso_mod, weightVs, weightGcs, damp, kmaxRc, tRc)


def read_model(path: str, nx: int, ny: int, nz: int):
    """MOD / MODVs.true -> (depz float32 (nz,), vs float32 Fortran-ordered (nx,ny,nz))."""
    with open(path) as f:
        vals = f.read().split()
    depz = np.array(vals[:nz], dtype=F32)
    body = np.array(vals[nz:nz + nx * ny * nz], dtype=F32)
    if body.size != nx * ny * nz:
        raise ValueError(f"{path}: expected {nx*ny*nz} velocities, found {body.size}")
    # file order: k outer, j, then i fastest == Fortran (nx,ny,nz) memory order
    vs = np.asfortranarray(body.reshape((nz, ny, nx)).transpose(2, 1, 0))
    return depz, vs


def read_gcgs(path: str, nx: int, ny: int, nz: int) -> np.ndarray:
    """MODGc.true / MODGs.true -> float32 Fortran-ordered (nx-2,ny-2,nz-1)."""
    with open(path) as f:
        vals = f.read().split()
    n = (nx - 2) * (ny - 2) * (nz - 1)
    body = np.array(vals[:n], dtype=F32)
    if body.size != n:
        raise ValueError(f"{path}: expected {n} values, found {body.size}")
    return np.asfortranarray(body.reshape((nz - 1, ny - 2, nx - 2)).transpose(2, 1, 0))


def write_model(path: str, depz: np.ndarray, vs: np.ndarray) -> None:
    nx, ny, nz = vs.shape
    with open(path, "w") as f:
        f.write(" ".join(f"{float(d):6.1f}" for d in depz) + "\n")
        for k in range(nz):
            for j in range(ny):
                f.write(" ".join(f"{float(vs[i, j, k]):6.3f}" for i in range(nx)) + "\n")


@dataclasses.dataclass
class Survey:
    """Station tables in the layout the Fortran drivers pass down (column-major).

    scxf/sczf are colatitude/longitude in radians (float32), built with the
    reference's float32 arithmetic.  Leading dimensions are compacted to the
    sizes actually used (the reference allocates nsrc x nsrc x kmax).
    """
    kmax: int
    nsrc: int                # leading dim of the (nsrc,kmax) tables
    nrcf: int                # leading dim of the receiver tables
    periods: np.ndarray      # int32 (nsrc,kmax) F-order
    nrc1: np.ndarray         # int32 (nsrc,kmax)
    nsrcsurf1: np.ndarray    # int32 (kmax,)
    scxf: np.ndarray         # float32 (nsrc,kmax)
    sczf: np.ndarray
    rcxf: np.ndarray         # float32 (nrcf,nsrc,kmax)
    rczf: np.ndarray
    wavetype: np.ndarray
    igrt: np.ndarray
    dist: np.ndarray         # float32 (dall,) delsph distances in file order
    obsvel: np.ndarray       # float32 (dall,) velocities in file order
    dall: int

    def row_offsets(self) -> np.ndarray:
        """First global row id (0-based) of every (period, source) in loop order."""
        out = []
        c = 0
        for k in range(self.kmax):
            for s in range(int(self.nsrcsurf1[k])):
                out.append(c)
                c += int(self.nrc1[s, k])
        return np.array(out + [c], dtype=np.int64)


def delsph32(flat1, flon1, flat2, flon2):
    """delsph.f90:1-28 in float32 (vectorised)."""
    R = F32(6371.0)
    flat1 = np.asarray(flat1, F32); flon1 = np.asarray(flon1, F32)
    flat2 = np.asarray(flat2, F32); flon2 = np.asarray(flon2, F32)
    dlat = flat2 - flat1
    dlon = flon2 - flon1
    lat1 = PI32 / F32(2) - flat1
    lat2 = PI32 / F32(2) - flat2
    sa = np.sin(dlat / F32(2)).astype(F32)
    so = np.sin(dlon / F32(2)).astype(F32)
    a = sa * sa + so * so * np.cos(lat1).astype(F32) * np.cos(lat2).astype(F32)
    c = F32(2) * np.arctan2(np.sqrt(a).astype(F32), np.sqrt(F32(1) - a).astype(F32)).astype(F32)
    return (R * c).astype(F32)


def read_surfdata(path: str, kmax: int) -> Survey:
    """Parse the '#'-block data file (MainForward.f90:239-281 / Main_Jt.f90:274-315)."""
    src_lat: List[List[float]] = [[] for _ in range(kmax)]
    src_lon: List[List[float]] = [[] for _ in range(kmax)]
    src_per: List[List[int]] = [[] for _ in range(kmax)]
    src_wt: List[List[int]] = [[] for _ in range(kmax)]
    src_vt: List[List[int]] = [[] for _ in range(kmax)]
    rec: List[List[List[tuple]]] = [[] for _ in range(kmax)]
    order = []  # (knum, istep) per data line, file order
    vels = []
    knum = 0
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            if line[0] == "#":
                t = line[1:].split()
                lat, lon, period, wavetp, veltp = float(t[0]), float(t[1]), int(t[2]), int(t[3]), int(t[4])
                if wavetp == 2 and veltp == 0:
                    new_knum = period
                else:
                    raise ValueError("can only deal with Rayleigh wave phase velocity data")
                if not 1 <= new_knum <= kmax:
                    raise ValueError("period index %d of source block '%s' is outside 1..%d (kmaxRc of para.in)"
                                     % (new_knum, line.strip(), kmax))
                # The reference resets its source counter when the period index changes (Main_Jt.f90:283-286,
                # MainForward.f90:248-251): a period that re-appears later in the file silently OVERWRITES its earlier
                # sources while the ray counter keeps running.  That is a corrupt input, not a feature: refuse it.
                if new_knum != knum and (src_lat[new_knum - 1]):
                    raise ValueError("period index %d appears in two separate groups of the data file; the reference "
                                     "would overwrite the first group (Main_Jt.f90:283-286): sort the file by period" % new_knum)
                knum = new_knum
                k = knum - 1
                src_lat[k].append(lat); src_lon[k].append(lon)
                src_per[k].append(period); src_wt[k].append(wavetp); src_vt[k].append(veltp)
                rec[k].append([])
            else:
                if knum == 0:
                    raise ValueError("data line before the first '#' source line")
                t = line.split()
                rec[knum - 1][-1].append((float(t[0]), float(t[1])))
                vels.append(float(t[2]))
                order.append((knum - 1, len(rec[knum - 1]) - 1))
    nsrc = max(1, max(len(x) for x in src_lat))
    nrcf = max(1, max((len(r) for k in range(kmax) for r in rec[k]), default=1))
    periods = np.zeros((nsrc, kmax), np.int32, order="F")
    nrc1 = np.zeros((nsrc, kmax), np.int32, order="F")
    wavetype = np.zeros((nsrc, kmax), np.int32, order="F")
    igrt = np.zeros((nsrc, kmax), np.int32, order="F")
    nsrcsurf1 = np.zeros((kmax,), np.int32)
    scxf = np.zeros((nsrc, kmax), F32, order="F")
    sczf = np.zeros((nsrc, kmax), F32, order="F")
    rcxf = np.zeros((nrcf, nsrc, kmax), F32, order="F")
    rczf = np.zeros((nrcf, nsrc, kmax), F32, order="F")
    c180 = F32(180.0)
    for k in range(kmax):
        ns = len(src_lat[k])
        nsrcsurf1[k] = ns
        if ns == 0:
            continue
        lat = np.array(src_lat[k], F32); lon = np.array(src_lon[k], F32)
        scxf[:ns, k] = (F32(90.0) - lat) * PI32 / c180
        sczf[:ns, k] = lon * PI32 / c180
        periods[:ns, k] = src_per[k]
        wavetype[:ns, k] = src_wt[k]
        igrt[:ns, k] = src_vt[k]
        for s in range(ns):
            r = rec[k][s]
            nrc1[s, k] = len(r)
            if r:
                rl = np.array([x[0] for x in r], F32); ro = np.array([x[1] for x in r], F32)
                rcxf[:len(r), s, k] = (F32(90.0) - rl) * PI32 / c180
                rczf[:len(r), s, k] = ro * PI32 / c180
    dall = len(vels)
    # distances in file order
    ks = np.array([o[0] for o in order], np.int64); ss = np.array([o[1] for o in order], np.int64)
    counters = {}
    ridx = np.zeros(dall, np.int64)
    for i, o in enumerate(order):
        ridx[i] = counters.get(o, 0)
        counters[o] = ridx[i] + 1
    dist = delsph32(scxf[ss, ks], sczf[ss, ks], rcxf[ridx, ss, ks], rczf[ridx, ss, ks]) if dall else np.zeros(0, F32)
    return Survey(kmax, nsrc, nrcf, periods, nrc1, nsrcsurf1, scxf, sczf, rcxf, rczf, wavetype, igrt,
                  dist, np.array(vels, F32), dall)


def survey_loop_coords(sv: Survey):
    """(scx, scz, rcx, rcz) float32 per ray in the reference's (period, source, receiver) loop order."""
    scx, scz, rcx, rcz = [], [], [], []
    for k in range(sv.kmax):
        for s in range(int(sv.nsrcsurf1[k])):
            n = int(sv.nrc1[s, k])
            scx.append(np.full(n, sv.scxf[s, k], F32)); scz.append(np.full(n, sv.sczf[s, k], F32))
            rcx.append(sv.rcxf[:n, s, k]); rcz.append(sv.rczf[:n, s, k])
    cat = lambda x: np.concatenate(x) if x else np.zeros(0, F32)
    return cat(scx), cat(scz), cat(rcx), cat(rcz)


def forward_velocities(sv: Survey, tsyn: np.ndarray) -> np.ndarray:
    """c = delsph/T per ray in loop order, float32 (MainForward.f90:413-424)."""
    scx, scz, rcx, rcz = survey_loop_coords(sv)
    d = delsph32(scx, scz, rcx, rcz)
    return (d / np.asarray(tsyn, F32)).astype(F32)


def write_surfphase_forward(path: str, sv: Survey, tsyn: np.ndarray) -> None:
    """surfphase_forward.dat: '(a,2f11.6,3I3)' headers, '(2f11.6,f9.5)' rows (MainForward.f90:403-429)."""
    vel = forward_velocities(sv, tsyn)
    c180 = F32(180.0)
    i = 0
    with open(path, "w") as f:
        for k in range(sv.kmax):
            for s in range(int(sv.nsrcsurf1[k])):
                latd = F32(90.0) - sv.scxf[s, k] * c180 / PI32
                lond = sv.sczf[s, k] * c180 / PI32
                f.write("#%11.6f%11.6f%3d%3d%3d\n" % (latd, lond, sv.periods[s, k], sv.wavetype[s, k], sv.igrt[s, k]))
                for r in range(int(sv.nrc1[s, k])):
                    lat2 = F32(90.0) - sv.rcxf[r, s, k] * c180 / PI32
                    lon2 = sv.rczf[r, s, k] * c180 / PI32
                    f.write("%11.6f%11.6f%9.5f\n" % (lat2, lon2, vel[i]))
                    i += 1


def read_surfphase_velocities(path: str) -> np.ndarray:
    """Velocity column of a surfphase file (data lines only), float64."""
    out = []
    with open(path) as f:
        for line in f:
            if line and line[0] != "#" and line.strip():
                out.append(float(line.split()[2]))
    return np.array(out)


def azim_map(nx, ny, nz, goxd, gozd, dvxd, dvzd, tRc, gcf, gsf, lsen_gsc, tRcV):
    """FwdAzimuthalAniMap.f90:41-79 -> array (kmax*(ny-2)*(nx-2), 9) as printed by '(10f10.5)'.
    Vectorised over (period, jj, ii); the float32 accumulation over layers keeps the reference's order."""
    kmax = len(tRc)
    nvx, nvz = nx - 2, ny - 2
    L = np.asarray(lsen_gsc, F32).reshape((nx, ny, kmax, nz - 1), order="F")[1:-1, 1:-1]      # node = jj*nx + ii
    tv = np.asarray(tRcV).reshape((nvx, nvz, kmax), order="F")
    gc = np.asarray(gcf, F32); gs = np.asarray(gsf, F32)
    ct = np.zeros((nvx, nvz, kmax), F32); st = np.zeros((nvx, nvz, kmax), F32)
    for kk in range(nz - 1):
        ct = (ct + L[:, :, :, kk] * gc[:, :, kk, None]).astype(F32)
        st = (st + L[:, :, :, kk] * gs[:, :, kk, None]).astype(F32)
    amp = np.sqrt((ct * ct + st * st).astype(F32)).astype(F32)
    isoc = tv.astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(isoc != 0, (amp / isoc).astype(F32), F32(0)).astype(F32)
    ang = (np.arctan2(st, ct).astype(np.float64) / np.float64(F32(3.1415926535898)) * 180).astype(F32)
    ang = np.where(ang < 0, (ang + F32(360)).astype(F32), ang).astype(F32)
    ang = (F32(0.5) * ang).astype(F32)
    lon = (F32(gozd) + np.arange(nvz, dtype=F32) * F32(dvzd)).astype(F32)          # gozd+(jj-1)*dvzd
    lat = (F32(goxd) - np.arange(nvx, dtype=F32) * F32(dvxd)).astype(F32)
    out = np.zeros((kmax, nvz, nvx, 9), np.float64)                                  # period outer, jj, ii fastest
    out[..., 0] = lon[None, :, None]
    out[..., 1] = lat[None, None, :]
    out[..., 2] = np.asarray(tRc, np.float64)[:, None, None]
    for c, a in ((3, isoc), (4, ang), (5, rel), (6, amp), (7, ct), (8, st)):
        out[..., c] = np.transpose(a, (2, 1, 0))
    return out.reshape(-1, 9)


# ---- writers of the inversion driver (Main_Jt.f90:752-908) ----------------------------------------------------

def fortran_e(x: float, w: int = 12, d: int = 3) -> str:
    """Fortran Ew.d edit descriptor (mantissa 0.ddd, two-digit exponent), e.g. e12.3 -> '   0.123E+01'."""
    x = float(x)
    if x == 0.0 or not np.isfinite(x):
        body = "0." + "0" * d + "E+00" if x == 0.0 else ("NaN" if x != x else "Infinity")
        s = ("-" if (x < 0 or (x == 0.0 and np.signbit(x))) else "") + body
        return s.rjust(w)
    m, e = ("%.*e" % (d - 1, abs(x))).split("e")          # d significant digits: m = 'D.DD..'
    e = int(e) + 1
    digits = m.replace(".", "")
    s = ("-" if x < 0 else "") + "0." + digits + "E%+03d" % e
    return s.rjust(w) if len(s) <= w else "*" * w


def write_vs_model(path: str, nx, ny, nz, gozd, goxd, dvzd, dvxd, depz, vsf) -> None:
    """writeVsmodel (Main_Jt.f90:838-856): '(5f8.4)' lon lat depth Vs, k outer, j, i fastest."""
    gozd, goxd, dvzd, dvxd = F32(gozd), F32(goxd), F32(dvzd), F32(dvxd)
    with open(path, "w") as f:
        for k in range(nz):
            for j in range(1, ny + 1):
                lon = gozd + F32(j - 2) * dvzd
                for i in range(1, nx + 1):
                    lat = goxd - F32(i - 2) * dvxd
                    f.write("%8.4f%8.4f%8.4f%8.4f\n" % (lon, lat, depz[k], vsf[i - 1, j - 1, k]))


def azimuthal_rows(nx, ny, nz, gozd, goxd, dvzd, dvxd, depz, gcf, gsf, vsf) -> np.ndarray:
    """The 8 columns of writeAzimuthal (Main_Jt.f90:859-887): lon lat depth(lower) Vs(mid) angle amp Gc% Gs%."""
    gozd, goxd, dvzd, dvxd = F32(gozd), F32(goxd), F32(dvzd), F32(dvxd)
    pi8 = np.float64(3.1415926535898)
    rows = []
    for k in range(nz - 1):
        for j in range(1, ny - 1):
            for i in range(1, nx - 1):
                c = F32(gcf[i - 1, j - 1, k]); s = F32(gsf[i - 1, j - 1, k])
                amp = F32(0.5) * F32(np.sqrt(F32(c * c + s * s)))
                ang = F32(np.arctan2(s, c).astype(np.float64) / pi8 * 180)
                if ang < 0:
                    ang = F32(ang + F32(360))
                ang = F32(0.5) * ang
                vsref = F32(F32(vsf[i, j, k] + vsf[i, j, k + 1]) / F32(2))
                rows.append((gozd + F32(j - 1) * dvzd, goxd - F32(i - 1) * dvxd, depz[k + 1], vsref, ang, amp,
                             c * F32(100), s * F32(100)))
    return np.array(rows, dtype=np.float64)


def write_azimuthal(path: str, nx, ny, nz, gozd, goxd, dvzd, dvxd, depz, gcf, gsf, vsf) -> None:
    """Gc_Gs_model.inv, '(8f10.4)'."""
    with open(path, "w") as f:
        for r in azimuthal_rows(nx, ny, nz, gozd, goxd, dvzd, dvxd, depz, gcf, gsf, vsf):
            f.write("".join("%10.4f" % x for x in r) + "\n")


def write_period_phasev(path: str, nx, ny, gozd, goxd, dvzd, dvxd, tRc, tRcV) -> None:
    """WTPeriodPhaseV (Main_Jt.f90:889-908): '(5f10.4)' lon lat period c over the interior nodes."""
    gozd, goxd, dvzd, dvxd = F32(gozd), F32(goxd), F32(dvzd), F32(dvxd)
    tv = np.asarray(tRcV).reshape(((nx - 2) * (ny - 2), len(tRc)), order="F")
    with open(path, "w") as f:
        for tt in range(len(tRc)):
            for jj in range(1, ny - 1):
                for ii in range(1, nx - 1):
                    f.write("%10.4f%10.4f%10.4f%10.4f\n" % (gozd + F32(jj - 1) * dvzd, goxd - F32(ii - 1) * dvxd, tRc[tt],
                                                           tv[(jj - 1) * (nx - 2) + ii - 1, tt]))


def write_mod_ref(path: str, depz, vsf) -> None:
    """MOD_Ref (Main_Jt.f90:753-769): depths f7.1 on the first record, then one record of nx f8.4 per (k, j)."""
    nx, ny, nz = vsf.shape
    with open(path, "w") as f:
        f.write("".join("%7.1f" % float(d) for d in depz))
        for k in range(nz):
            for j in range(ny):
                f.write("\n" + "".join("%8.4f" % float(vsf[i, j, k]) for i in range(nx)))
        f.write("\n")


def interior_phase_velocity(pvRc, nx, ny) -> np.ndarray:
    """tRcV((nx-2)(ny-2),kmax) from pvRc(nx*ny,kmax) (FwdTraveltimeCPS.f90:771-778)."""
    pv = np.asarray(pvRc)
    k = pv.shape[1]
    out = np.zeros(((nx - 2) * (ny - 2), k), np.float64, order="F")
    for jj in range(1, ny - 1):
        out[(jj - 1) * (nx - 2):(jj) * (nx - 2), :] = pv[jj * nx + 1:jj * nx + nx - 1, :]
    return out


def stage_reference_example(fixture_dir: str, tag: str, outdir: str) -> str:
    """Unpack the fixture copy of one of the reference's inversion examples (tests/golden/inv: test2 = test2_syn_iso_inv,
    test3 = test3_syn_joint_inv, test4 = test4_Yunnan) into outdir as para.in / MOD / <data file>; returns para.in's path."""
    import lzma
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "para.in"), "w") as f:
        f.write(open(os.path.join(fixture_dir, "%s_para.in" % tag)).read())
    datafile = read_para_inv(os.path.join(outdir, "para.in")).datafile
    mod = os.path.join(fixture_dir, "%s_MOD" % tag)
    with open(os.path.join(outdir, "MOD"), "wb") as f:
        f.write(lzma.open(mod + ".xz", "rb").read() if os.path.exists(mod + ".xz") else open(mod, "rb").read())
    data = "test4_data.dat.xz" if tag == "test4" else "surfphase_forward_RV3th.dat.xz"
    with open(os.path.join(outdir, datafile), "wb") as f:
        f.write(lzma.open(os.path.join(fixture_dir, data), "rb").read())
    return os.path.join(outdir, "para.in")

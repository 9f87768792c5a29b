"""Thin forward driver with the reference's command line and file formats:

    python -m dazimsurftomo_b200.forward para.in        (reference: SurfAAForward para.in)

Reads para.in (forward layout, MainForward.f90:147-159,188,330), the '#'-block data file
(:239-281), MODVs.true / MODGc.true / MODGs.true (:334-356) from the directory of para.in, runs
the forward-modelling path on the GPU through the same routine the Fortran driver would call
(FwdObsTraveltimeCPS, MainForward.f90:372), and writes

    surfphase_forward.dat      '(a,2f11.6,3I3)' / '(2f11.6,f9.5)'      (MainForward.f90:403-429)
    period_Azm_tomo.real       '(10f10.5)' azimuthal-anisotropy map    (FwdAzimuthalAniMap.f90:79)

Noise (MainForward.f90:390-401) is applied only when para.in asks for it; the reference's generator
is unseeded (gaussian.f90), so noisy outputs are not reproducible on either side.
Everything numerical happens in libdazim_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import api, formats as fm


def run(para_path: str, outdir: str | None = None, handle=None) -> dict:
    t0 = time.time()
    base = os.path.dirname(os.path.abspath(para_path))
    outdir = outdir or base
    p = fm.read_para_forward(para_path)
    depz, vs = fm.read_model(os.path.join(base, "MODVs.true"), p.nx, p.ny, p.nz)
    gc = fm.read_gcgs(os.path.join(base, "MODGc.true"), p.nx, p.ny, p.nz)
    gs = fm.read_gcgs(os.path.join(base, "MODGs.true"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(base, p.datafile), p.kmaxRc)
    r = api.FwdObsTraveltimeCPS(vs, gc, gs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, handle=handle)
    print("  DepthkernelTI time cost= %13.4f s" % (r["times"]["kernels_ms"] * 1e-3))
    tsyn = (r["dsurf"] + r["obsTaa"]).astype(np.float32)          # T = T_iso + T_aa (MainForward.f90:390)
    if p.noiselevel > 0:
        rng = np.random.default_rng()
        tsyn = (tsyn + np.float32(p.noiselevel) * rng.standard_normal(len(tsyn)).astype(np.float32)).astype(np.float32)
    os.makedirs(outdir, exist_ok=True)
    fm.write_surfphase_forward(os.path.join(outdir, "surfphase_forward.dat"), sv, tsyn)
    tab = fm.azim_map(p.nx, p.ny, p.nz, p.goxd, p.gozd, p.dvxd, p.dvzd, p.tRc, gc, gs, r["Lsen_Gsc"], r["tRcV"])
    with open(os.path.join(outdir, "period_Azm_tomo.real"), "w") as f:
        for row in tab:
            f.write("".join("%10.5f" % x for x in row) + "\n")
    print("  All time cost= %13.4f s" % (time.time() - t0))
    return dict(para=p, survey=sv, tsyn=tsyn, azim=tab, times=r["times"])


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 1:
        print("usage: python -m dazimsurftomo_b200.forward para.in [outdir]", file=sys.stderr)
        return 2
    run(argv[0], argv[1] if len(argv) > 1 else None)
    return 0


if __name__ == "__main__":
    sys.exit(main())

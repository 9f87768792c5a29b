"""Thin inversion driver with the reference's command line and file formats:

    python -m dazimsurftomo_b200.invert para.in [outdir]       (reference: DAzimSurfTomo para.in)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 -m dazimsurftomo_b200.invert para.in

Reads para.in (inversion layout, Main_Jt.f90:158-211), the '#'-block data file (:274-315) and MOD (:345-353) from
the directory of para.in and runs the outer loop of Main_Jt.f90 (:364-750) with every numerical stage on the GPU:

    depthkernel / depthkernelTI  ->  plan.run (dice, eikonal, rays, G rows; CalSurfG / CalSurfGAnisoJoint)
    ->  plan.iterate (residual, CalDdatSigma, weights, Tikhonov rows, LSMR, model update, norms)

With N ranks the depth kernels are cut into strips of grid rows, the G build into period-aligned (period, source)
ranges, the row blocks are all-gathered over NCCL from HBM to HBM and every rank solves the gathered system.
G stays in HBM for the whole iteration; the host sees the model (nx*ny*nz floats), the solution vector and the
per-row travel-time columns.  Files written (same names and formats as the reference):

    DSurfTomo.inv  Gc_Gs_model.inv  MOD_Ref  period_phaseVMOD.dat  phaseV_FWD.dat  period_Azm_tomo.inv
    IterVel.out  Traveltime_statis_00th.dat  <para.in>_inv.log  lsmr.txt (one summary line per outer iteration)

Deliberate deviations in ISOTROPIC mode (iso_mod = T), where the reference's own outputs are degenerate: CalSurfG has no
tRcV argument, so Main_Jt.f90 writes zeros into period_phaseVMOD.dat / phaseV_FWD.dat and an all-zero
period_Azm_tomo.inv (:784-787).  This driver writes the real phase velocities of the final model into the first two and
does not write period_Azm_tomo.inv (there is no anisotropy to map).  Everything else, and every file in joint mode, has the
reference's names, column layout and formats.

There is no CPU fallback: everything numerical happens in libdazim_b200.so.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import api, formats as fm, partition


def loop_order_obst(sv: fm.Survey) -> np.ndarray:
    """obst = dist/velocity per row (Main_Jt.f90:306-308).  The reference stores it in FILE order while rows are
    numbered in (period, source, receiver) loop order; the two agree only for period-sorted files, which
    is also the only kind its reader parses correctly -- anything else is refused here."""
    scx, scz, rcx, rcz = fm.survey_loop_coords(sv)
    d = fm.delsph32(scx, scz, rcx, rcz)
    if d.shape != sv.dist.shape or not np.array_equal(d, sv.dist):
        raise ValueError("data file is not sorted by period: row order and file order differ (Main_Jt.f90:283-297)")
    return (sv.dist / sv.obsvel).astype(np.float32)


def _ranks():
    """(rank, world, torch device or None).  Under torchrun (WORLD_SIZE > 1) one process per GPU over NCCL: the depth
    kernels are cut into strips of grid rows, the G build into contiguous (period, source) ranges aligned to periods
    ("period-sharded", BASELINE config 4), and the row blocks are all-gathered before the solve -- the one place the
    inversion needs the full system."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), world, torch.device("cuda", local)


def tables_for(iso_inv: bool, vsf, depz, tRc, minthk, handle, rank=0, world=1, device=None):
    """The depth-kernel tables one G build needs (CalSurfG.f90:1004, CalSurfGAniso_Joint.f90:321-333); with several
    ranks each computes a strip of grid rows and one all-gather hands every rank the full tables."""
    nx, ny, nz = vsf.shape
    strips = partition.node_strips(ny, world)
    part = vsf if world == 1 else partition.strip_model(vsf, strips[rank], strips[rank + 1])
    t = {}
    if not iso_inv:
        _, t["Lsen_Gsc"] = api.depthkernelTI(part, depz, tRc, minthk, handle=handle)
    t["pvRc"], t["sen_vs"], t["sen_vp"], t["sen_rho"] = api.depthkernel(part, depz, tRc, minthk, handle=handle)
    if world > 1:
        t = partition.gather_tables(t, nx, ny, strips, rank, device=device)
    return t


def run(para_path: str, outdir: str | None = None, handle=None, maxiter: int | None = None, sv=None,
        write_files: bool = True, log_stream=None, rows: bool | None = None) -> dict:
    """rows (several ranks only; default: environment DAZIM_ROWS=1): leave every rank's rows of G where they were built
    and run the tail row-distributed (Plan.iterate_rows: no all-gather of G, LSMR with one n-vector all-reduce per
    iteration) instead of gathering the system on every rank."""
    t_start = time.time()
    if rows is None:
        rows = os.environ.get("DAZIM_ROWS", "0") not in ("", "0")
    base = os.path.dirname(os.path.abspath(para_path))
    outdir = outdir or base
    p = fm.read_para_inv(para_path)
    depz, vsf = fm.read_model(os.path.join(base, "MOD"), p.nx, p.ny, p.nz)
    if sv is None:
        sv = fm.read_surfdata(os.path.join(base, p.datafile), p.kmaxRc)
    obst = loop_order_obst(sv)
    rank, world, device = _ranks()
    write_files = write_files and rank == 0
    if rank != 0 and log_stream is None:
        log_stream = open(os.devnull, "w")
    h = handle or api.default_handle()
    nx, ny, nz = p.nx, p.ny, p.nz
    maxvp = (nx - 2) * (ny - 2) * (nz - 1)
    niter = p.maxiter if maxiter is None else maxiter
    iso_inv = bool(p.iso_mod)
    gcf = np.zeros((nx - 2, ny - 2, nz - 1), np.float32, order="F"); gsf = np.zeros_like(gcf)
    os.makedirs(outdir, exist_ok=True)
    logf = open(os.path.join(outdir, os.path.basename(para_path) + "_inv.log"), "w") if write_files else None
    iterf = open(os.path.join(outdir, "IterVel.out"), "w") if write_files else None
    lsmrf = open(os.path.join(outdir, "lsmr.txt"), "w") if write_files else None

    def say(s=""):
        print(s, file=log_stream or sys.stdout)
        if logf:
            logf.write(s + "\n")

    say("")
    say("                  DAzimSurfTomo  (dazim_b200: GPU path)")
    say("")
    say(" model origin:latitude,longitue")
    say("%10.4f%10.4f" % (p.goxd, p.gozd))
    say(" grid spacing:latitude,longitue")
    say("%10.4f%10.4f" % (p.dvxd, p.dvzd))
    say(" model dimension:nx,ny,nz")
    say("%5d%5d%5d" % (nx, ny, nz))
    say(" Rayleigh wave phase velocity used,periods:(s)")
    say("".join("%6.1f" % t for t in p.tRc))
    say(" Number of all measurements%7d" % sv.dall)

    plan = None
    comm = None
    if world > 1 and rows:
        comm = api.Comm.from_torch(device.index)
    history = []
    tRcV_first = tRcV = None
    tables = None
    rows_out = None
    gpu_ms = dict(kernels=0.0, gbuild=0.0, iterate=0.0, lsmr=0.0)
    try:
        for it in range(1, niter + 1):
            say(" -----------------------------------------------------------")
            say("%12d%s" % (it, "th iteration, invert for isotropic Vs para." if iso_inv else
                            "th iteration, invert for dVs, Gc, Gs "))
            say(" -----------------------------------------------------------")
            tables = tables_for(iso_inv, vsf, depz, p.tRc, p.sublayers, h, rank, world, device)
            gpu_ms["kernels"] += h.times["kernels_ms"]
            if plan is None:
                sb, se = 0, -1
                if world > 1:
                    bounds = partition.split_units(sv, world, align_periods=True)
                    sb, se = bounds[rank], bounds[rank + 1]
                plan = api.Plan(1 if iso_inv else 2, vsf, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv,
                                tables, src_begin=sb, src_end=se, handle=h)
            else:
                plan.update_model(vsf, tables)
            tm = plan.run()
            gpu_ms["gbuild"] += tm["total_ms"]
            system = None
            nnz_all = plan.nnz
            if world > 1 and comm is not None:
                # rows stay where they are: only the non-zero count of the whole system is needed (spfra test below)
                import torch
                import torch.distributed as dist
                cnt = torch.tensor([plan.nnz], dtype=torch.int64, device=device)
                dist.all_reduce(cnt)
                nnz_all = int(cnt.item())
            elif world > 1:
                # the exchange step: all-gather-v of the CSR row blocks and dsurf over NCCL, straight from HBM
                import torch
                t = plan.device_tensors()
                t0 = time.perf_counter()
                full = partition.gather_rows(dict(dsurf=t["dsurf"], nnz_row=t["rowptr"][1:] - t["rowptr"][:-1], col=t["col"],
                                                  val=t["val"]), plan.row0)
                system = partition.assemble_system(full, nx, ny, nz, joint=not iso_inv)
                torch.cuda.synchronize()
                gpu_ms["gather"] = gpu_ms.get("gather", 0.0) + (time.perf_counter() - t0) * 1e3
                nnz_all = system["nnz"]
            # maxnar = spfra*dall*nx*ny*nz*3 in default-real arithmetic (Main_Jt.f90:325); the reference tests nar AFTER
            # the regularisation rows are appended (Main_Jt.f90:513-523)
            maxnar = int(np.float32(np.float32(np.float32(np.float32(np.float32(p.spfra) * np.float32(sv.dall)) * np.float32(nx))
                                               * np.float32(ny)) * np.float32(nz)) * np.float32(3))
            nar_with_reg = nnz_all + (1 if iso_inv else 3) * partition.tikh_block_entries(nx, ny, nz)
            if nar_with_reg > maxnar:
                raise api.DazimError(3, "increase sparsity fraction(spfra)")
            tRcV = fm.interior_phase_velocity(tables["pvRc"], nx, ny)
            if it == 1:
                tRcV_first = tRcV
            last = it == niter
            if world == 1:
                r = plan.iterate(obst, vsf, iso_inv, p.weightVs, p.weightGcs, p.damp, p.minvel, p.maxvel,
                                 want_rows=(it == 1 or last))
            elif comm is not None:
                r = plan.iterate_rows(comm, sv.dall, obst, vsf, iso_inv, p.weightVs, p.weightGcs, p.damp, p.minvel, p.maxvel,
                                      want_rows=(it == 1 or last))
            else:
                # every rank solves the gathered system (same data, same kernels: identical models, no broadcast needed)
                r = api.iterate_device((nx, ny, nz), system, obst, vsf, iso_inv, p.weightVs, p.weightGcs, p.damp, p.minvel,
                                       p.maxvel, want_rows=(it == 1 or last), handle=h)
            s = r["stats"]
            gpu_ms["iterate"] += s["step_ms"]; gpu_ms["lsmr"] += s["lsmr"]["solve_ms"] + s["lsmr"]["setup_ms"]
            vsf = r["vsf"]
            dv = r["dv"]
            if not iso_inv:
                gcf, gsf = r["gcf"], r["gsf"]
            b, a = s["before"], s["after"]
            say("  Before Inversion: abs mean, std, RMS of Res:%12.4f s %10.2f s %10.2f s" % (b["meanabs"], b["std"], b["rms"]))
            say("  mean data weight:%8.3f |  abs data mean with weight:%8.3fs  |  dt/t0:%7.3f %%" %
                (s["mean_weight"], s["meanabs_weighted"], s["meandeltaT"] * 100))
            say("  damp,  lamebda Gsc, lamebda Vs: %8.2f%8.2f%8.2f" % (p.damp, p.weightGcs, p.weightVs))
            if s["lsmr"]["istop"] == 3:
                say("  istop = 3, large condition number, LSMR failed")
            say("  itn=               %7d" % s["lsmr"]["itn"])
            say("  Condition NO. of A=%7.1f" % s["lsmr"]["condA"])
            say("  min  max and abs mean  dVs (km/s)%10.4f%10.4f%10.4f" %
                (dv[:maxvp].min(), dv[:maxvp].max(), np.abs(dv[:maxvp]).sum(dtype=np.float32) / maxvp))
            if not iso_inv:
                for name, blk in (("Gc", dv[maxvp:2 * maxvp]), ("Gs", dv[2 * maxvp:])):
                    say("  min  max and abs mean   %s/L (%%) %10.4f%10.4f%10.4f" %
                        (name, blk.min() * 100, blk.max() * 100, np.abs(blk).sum(dtype=np.float32) / maxvp * 100))
            per = (nx - 2) * (ny - 2)
            for k in range(nz - 1):
                vv = np.abs(dv[k * per:(k + 1) * per]).sum(dtype=np.float32) / per
                if iso_inv:
                    say("  Z %5.1f - %5.1f km  abs mean dVs (km/s)%10.4f" % (depz[k], depz[k + 1], vv))
                else:
                    say("  Z %5.1f - %5.1f km  Abs Mean Gc (%%)  Gs (%%)   dVs (km/s)%10.3f%10.3f%9.4f" %
                        (depz[k], depz[k + 1], np.abs(gcf[:, :, k]).sum(dtype=np.float32) / per * 100,
                         np.abs(gsf[:, :, k]).sum(dtype=np.float32) / per * 100, vv))
            nm = s["norms"]
            if iso_inv:
                say("  dVs:  ||Lm||^2      and ||wLm||^2    : %12.3f%12.3f" % (nm["Mnorm2"], nm["MwNorm2"]))
                say("  dVs:  ||(Gm-d)||^2  and ||W(Gm-d)||^2: %12.3f%12.3f" % (s["res2Nm"], s["resW2Nm"]))
            else:
                say("  dVs:  ||Lm||^2   and   ||wLm||^2     : %12.3f%12.3f" % (nm["VsNorm2"], nm["VswNorm2"]))
                say("  Gcs:  ||Lm||^2   and   ||wLm||^2     : %12.3f%12.3f" % (nm["GcsNorm2"], nm["GcswNorm2"]))
                say("  All:  ||Lm||^2   and   ||wLm||^2     : %12.3f%12.3f" % (nm["Mnorm2"], nm["MwNorm2"]))
                say("  All:  ||(Gm-d)||^2  and ||W(Gm-d)||^2: %12.3f%12.3f" % (s["res2Nm"], s["resW2Nm"]))
                say("  ABS Mean T(AA): %12.4fs" % s["meanabs_Taa"])
                say("  ABS Mean T(dVs):%12.4fs" % s["meanabs_Tvs"])
            say("  After Inversion: abs mean, std, RMS of Res :%12.4f s %10.2f s %10.2f s" % (a["meanabs"], a["std"], a["rms"]))
            say("")
            if lsmrf:
                lsmrf.write("outer %d: istop %d itn %d normA %.4e condA %.4e normr %.4e normAr %.4e normx %.4e\n" %
                            (it, s["lsmr"]["istop"], s["lsmr"]["itn"], s["lsmr"]["normA"], s["lsmr"]["condA"],
                             s["lsmr"]["normr"], s["lsmr"]["normAr"], s["lsmr"]["normx"]))
            if iterf and p.iso_mod:                   # Main_Jt.f90:733-746
                iterf.write(" ,OUTPUT S VELOCITY AT ITERATION%12d\n" % it)
                for k in range(nz):
                    for j in range(ny):
                        iterf.write("".join("%7.3f" % vsf[i, j, k] for i in range(nx)) + "\n")
                iterf.write(" ,OUTPUT DWS AT ITERATION%12d\n" % it)
                dws = r["dws"].reshape((nx - 2, ny - 2, nz - 1), order="F")
                for k in range(nz - 1):
                    for j in range(ny - 2):
                        iterf.write("".join("%10.3f" % dws[i, j, k] for i in range(nx - 2)) + "\n")
            if r.get("resbst") is not None:
                if world == 1:
                    dsyn = plan.fetch(csr=False)["dsurf"]
                elif comm is not None:            # the synthetic times of every rank's rays (zero-padded sum)
                    import torch
                    import torch.distributed as dist
                    loc = plan.fetch(csr=False)["dsurf"]
                    full = torch.zeros(sv.dall, dtype=torch.float32, device=device)
                    full[plan.row0:plan.row0 + len(loc)] = torch.from_numpy(np.ascontiguousarray(loc)).to(device)
                    dist.all_reduce(full)
                    dsyn = full.cpu().numpy()
                else:
                    dsyn = system["dsurf"].cpu().numpy()
                rows_out = dict(dsyn=dsyn, Tdata=(obst - dsyn).astype(np.float32), fwdTvs=r["fwdTvs"], fwdTaa=r["fwdTaa"],
                                resbst=r["resbst"], sigmaT=r["sigmaT"])
                # Main_Jt.f90:701-717: written at iteration 1 and at the last one under the same name (id stays '00': its
                # write is commented out), so only the last one survives -- write that one
                if write_files and last:
                    with open(os.path.join(outdir, "Traveltime_statis_00th.dat"), "w") as f:
                        if p.iso_mod:
                            f.write("   Dist(km)   T_obs(s)  T_ref_iso   Res(in)   dT(dvs)   Res(out)\n")
                            for i in range(sv.dall):
                                f.write("%10.3f%10.3f%10.3f" % (sv.dist[i], obst[i], rows_out["dsyn"][i]) +
                                        "".join(fm.fortran_e(x) for x in (rows_out["Tdata"][i], r["fwdTvs"][i], r["resbst"][i])) + "\n")
                        else:
                            f.write("          Dist(km)       T_obs(s)        T_ref-iso        Res(in)   dT(aa)        dT(dvs)        Res(out)\n")
                            for i in range(sv.dall):
                                f.write("%10.4f%10.4f%10.4f" % (sv.dist[i], obst[i], rows_out["dsyn"][i]) +
                                        "".join(fm.fortran_e(x) for x in (rows_out["Tdata"][i], r["fwdTvs"][i], r["fwdTaa"][i],
                                                                          r["resbst"][i])) + "\n")
            history.append(s)
    finally:
        if plan is not None:
            plan.close()
    if write_files:
        fm.write_mod_ref(os.path.join(outdir, "MOD_Ref"), depz, vsf)
        fm.write_vs_model(os.path.join(outdir, "DSurfTomo.inv"), nx, ny, nz, p.gozd, p.goxd, p.dvzd, p.dvxd, depz, vsf)
        fm.write_azimuthal(os.path.join(outdir, "Gc_Gs_model.inv"), nx, ny, nz, p.gozd, p.goxd, p.dvzd, p.dvxd, depz, gcf, gsf, vsf)
        if tRcV_first is not None:
            fm.write_period_phasev(os.path.join(outdir, "period_phaseVMOD.dat"), nx, ny, p.gozd, p.goxd, p.dvzd, p.dvxd, p.tRc, tRcV_first)
            fm.write_period_phasev(os.path.join(outdir, "phaseV_FWD.dat"), nx, ny, p.gozd, p.goxd, p.dvzd, p.dvxd, p.tRc, tRcV)
        if tables is not None and "Lsen_Gsc" in tables:     # FwdAzimuthalAniMap (Main_Jt.f90:784-787); Lsen_Gsc is zero in iso mode
            tab = fm.azim_map(nx, ny, nz, p.goxd, p.gozd, p.dvxd, p.dvzd, p.tRc, gcf, gsf, tables["Lsen_Gsc"], tRcV)
            with open(os.path.join(outdir, "period_Azm_tomo.inv"), "w") as f:
                for row in tab:
                    f.write("".join("%10.5f" % x for x in row) + "\n")
    say("  -----------------------------------------------------------")
    say("   Program finishes successfully")
    say("   All time cost= %13.1fs   (GPU: depth kernels %.0f ms, G build %.0f ms, iteration tail %.0f ms of which LSMR %.0f ms%s)" %
        (time.time() - t_start, gpu_ms["kernels"], gpu_ms["gbuild"], gpu_ms["iterate"], gpu_ms["lsmr"],
         "" if world == 1 else ("; %d ranks, rows of G left in place, row-distributed LSMR" % world if comm is not None else
                                "; %d ranks, all-gather of the row blocks %.0f ms" % (world, gpu_ms.get("gather", 0.0)))))
    if comm is not None:
        comm.close()
    for f in (logf, iterf, lsmrf):
        if f:
            f.close()
    return dict(para=p, survey=sv, vsf=vsf, gcf=gcf, gsf=gsf, history=history, rows=rows_out, gpu_ms=gpu_ms, depz=depz,
                rank=rank, world=world)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 1:
        print("usage: python -m dazimsurftomo_b200.invert para.in [outdir] [--rows]", file=sys.stderr)
        return 2
    rows = True if "--rows" in argv else None
    argv = [a for a in argv if a != "--rows"]
    run(argv[0], argv[1] if len(argv) > 1 else None, rows=rows)
    return 0


if __name__ == "__main__":
    sys.exit(main())

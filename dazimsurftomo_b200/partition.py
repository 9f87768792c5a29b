"""(period x source) partition of the forward-modelling path across ranks and the one
exchange step before the solver.

Units are the (period, source) pairs in the reference's loop order
(FwdTraveltimeCPS.f90:464-465).  Every rank owns a contiguous range of units, hence a
contiguous range of global row ids (`count1`, CalSurfGAniso_Joint.f90:693): no data-path
collective is needed while G is being built.  Only "where the inversion step needs the full
system" are the row blocks exchanged: an all-gather of per-rank sizes, then an all-gather of
equal-size padded blocks (NCCL has no gather-v), trimmed on arrival.  Works on any
torch.distributed backend (NCCL with CUDA tensors on the B200 box, gloo with CPU tensors in
the CPU tests).
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional

import numpy as np

from .formats import Survey


def split_units(sv: Survey, world: int, align_periods: bool = False) -> List[int]:
    """Unit boundaries [b_0=0, ..., b_world=nunit], contiguous, balanced by ray count.

    align_periods=True snaps every boundary to a period boundary (BASELINE config 4:
    "period-sharded")."""
    offs = sv.row_offsets()
    nunit = len(offs) - 1
    pstart = np.concatenate([[0], np.cumsum(sv.nsrcsurf1.astype(np.int64))])
    bounds = [0]
    for r in range(1, world):
        target = offs[-1] * r / world
        b = int(np.searchsorted(offs, target))
        if align_periods:
            b = int(pstart[np.argmin(np.abs(offs[pstart] - target))])
        b = min(max(b, bounds[-1]), nunit)
        bounds.append(b)
    bounds.append(nunit)
    return bounds


def sub_survey(sv: Survey, ub: int, ue: int):
    """Survey restricted to units [ub, ue) (same loop order) and the global id of its first row."""
    offs = sv.row_offsets()
    out = copy.copy(sv)
    kmax = sv.kmax
    ns1 = np.zeros(kmax, np.int32)
    sel = []   # (k, s) kept
    u = 0
    for k in range(kmax):
        for s in range(int(sv.nsrcsurf1[k])):
            if ub <= u < ue:
                sel.append((k, s))
                ns1[k] += 1
            u += 1
    nsrc = max(1, int(ns1.max()) if len(sel) else 1)
    z2 = lambda a: np.zeros((nsrc, kmax), a.dtype, order="F")
    periods, nrc1, scxf, sczf = z2(sv.periods), z2(sv.nrc1), z2(sv.scxf), z2(sv.sczf)
    wavetype, igrt = z2(sv.wavetype), z2(sv.igrt)
    rcxf = np.zeros((sv.nrcf, nsrc, kmax), sv.rcxf.dtype, order="F")
    rczf = np.zeros((sv.nrcf, nsrc, kmax), sv.rczf.dtype, order="F")
    pos = np.zeros(kmax, np.int64)
    for k, s in sel:
        j = int(pos[k]); pos[k] += 1
        periods[j, k] = sv.periods[s, k]; nrc1[j, k] = sv.nrc1[s, k]
        scxf[j, k] = sv.scxf[s, k]; sczf[j, k] = sv.sczf[s, k]
        wavetype[j, k] = sv.wavetype[s, k]; igrt[j, k] = sv.igrt[s, k]
        rcxf[:, j, k] = sv.rcxf[:, s, k]; rczf[:, j, k] = sv.rczf[:, s, k]
    out.nsrc = nsrc
    out.periods, out.nrc1, out.nsrcsurf1, out.scxf, out.sczf = periods, nrc1, ns1, scxf, sczf
    out.wavetype, out.igrt, out.rcxf, out.rczf = wavetype, igrt, rcxf, rczf
    row0, row1 = int(offs[min(ub, len(offs) - 1)]), int(offs[min(ue, len(offs) - 1)])
    out.dall = row1 - row0
    # dist/obsvel are in FILE order == loop order for period-sorted files
    out.dist = sv.dist[row0:row1] if len(sv.dist) == sv.dall else sv.dist
    out.obsvel = sv.obsvel[row0:row1] if len(sv.obsvel) == sv.dall else sv.obsvel
    return out, row0


def _all_gather_v(t, dist, group=None):
    """All-gather of 1-D tensors of different length (same dtype/device) in rank order.
    NCCL: one collective straight into exact-size slices of a single output buffer (uneven all-gather,
    no padding, no concatenation copy).  gloo (CPU tests): sizes first, then padded equal blocks."""
    import torch
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    allsz = torch.zeros(world, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(allsz, n, group=group)
    sizes = [int(x) for x in allsz.tolist()]
    if t.is_cuda:
        out = torch.empty(sum(sizes), dtype=t.dtype, device=t.device)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        views = [out[int(offs[r]):int(offs[r + 1])] for r in range(world)]
        dist.all_gather(views, t.contiguous(), group=group)
        return out, sizes
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(blocks, sizes)]), sizes


def gather_rows(block: Dict[str, "object"], row0: int, group=None) -> Dict[str, "object"]:
    """Exchange step: every rank contributes its CSR row block
    {dsurf (rows,), nnz_row (rows,), col (nnz,), val (nnz,)} (torch tensors on one device) whose first
    global row is row0; every rank receives the full system in the reference's COO convention
    (rw, row 1-based = iw(2:nar+1), col 1-based) in global row order.  Ranks hold contiguous,
    ascending row ranges, so concatenation in rank order IS global row order (bit-identical
    for any world size)."""
    import torch
    import torch.distributed as dist
    dsurf, _ = _all_gather_v(block["dsurf"], dist, group)
    nnz_row, rsizes = _all_gather_v(block["nnz_row"].to(torch.int64), dist, group)
    col, _ = _all_gather_v(block["col"], dist, group)
    val, _ = _all_gather_v(block["val"], dist, group)
    r0 = torch.tensor([row0], dtype=torch.int64, device=dsurf.device)
    r0s = [torch.zeros_like(r0) for _ in range(dist.get_world_size(group))]
    dist.all_gather(r0s, r0, group=group)
    exp = 0
    for r, n in zip(r0s, rsizes):
        if int(r.item()) != exp:
            raise RuntimeError("row blocks are not contiguous in rank order")
        exp += n
    rows = torch.repeat_interleave(torch.arange(1, nnz_row.numel() + 1, device=nnz_row.device, dtype=torch.int32), nnz_row)
    return dict(dsurf=dsurf, rw=val, col=col, row=rows, nar=int(val.numel()), nnz_row=nnz_row)


def tikh_block_entries(nx: int, ny: int, nz: int) -> int:
    """Entries of one Tikhonov block (TikhRegul.f90:22-56): one per cell + six more per interior cell."""
    nvx, nvz, nl = nx - 2, ny - 2, nz - 1
    return nvx * nvz * nl + 6 * max(nvx - 2, 0) * max(nvz - 2, 0) * max(nl - 2, 0)


def assemble_system(full: Dict[str, "object"], nx: int, ny: int, nz: int, joint: bool) -> Dict[str, "object"]:
    """What the solver stage needs from the all-gathered row blocks (gather_rows output): CSR row pointers and the
    val / col / row arrays re-housed with room for the regularisation rows that dazim_iterate_device appends
    (1 block in iso mode, 3 in joint mode).  Pure torch: runs on CUDA tensors on the box, on CPU tensors in the tests."""
    import torch
    nnz = int(full["rw"].numel())
    nrow = int(full["dsurf"].numel())
    cap = nnz + (3 if joint else 1) * tikh_block_entries(nx, ny, nz)
    dev = full["rw"].device
    rowptr = torch.zeros(nrow + 1, dtype=torch.int64, device=dev)
    torch.cumsum(full["nnz_row"].to(torch.int64), 0, out=rowptr[1:])
    if int(rowptr[-1]) != nnz:
        raise RuntimeError("row counts and triplets disagree")

    def housed(t):
        out = torch.empty(cap, dtype=t.dtype, device=dev)
        out[:nnz].copy_(t)
        return out

    return dict(nrow=nrow, nnz=nnz, cap=cap, rowptr=rowptr, col=housed(full["col"]), val=housed(full["rw"]),
                row=housed(full["row"]), dsurf=full["dsurf"].contiguous())


def misfit_sums(obst, dsurf, group=None):
    """mean / std / rms inputs of Main_Jt.f90:432-434 as three float64 partial sums reduced over ranks:
    returns (n, sum(r), sum(r^2)) of r = obst - dsurf over all ranks."""
    import torch
    import torch.distributed as dist
    r = (obst.to(torch.float64) - dsurf.to(torch.float64))
    s = torch.stack([torch.tensor(float(r.numel()), dtype=torch.float64, device=r.device), r.sum(), (r * r).sum()])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        # fixed rank order: gather then sum on every rank identically (deterministic for any backend)
        parts = [torch.zeros_like(s) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, s, group=group)
        s = torch.stack(parts).sum(0)
    return s


# ---------------------------------------------------------------------------------------------
# Stage A (Thomson-Haskell depth kernels): nodes are independent, so the model is cut into strips
# of grid rows (the jj index of depthkernel's outer loop, CalSurfG.f90:39-43), each rank computes
# its strip, and ONE exchange (all-gather of the strip tables) gives every rank the full
# phase-velocity map it needs to dice any period (SURVEY 8e stage A).

def node_strips(ny: int, world: int) -> List[int]:
    """Row boundaries [0, ..., ny] of the per-rank strips (contiguous, near-equal)."""
    return [int(round(ny * r / world)) for r in range(world + 1)]


def strip_model(vels: np.ndarray, j0: int, j1: int) -> np.ndarray:
    """vels(nx, j0:j1, nz) as its own Fortran-ordered model (a strip is a valid model for the depth kernels)."""
    return np.asfortranarray(vels[:, j0:j1, :])


def place_strip(full: np.ndarray, part: np.ndarray, nx: int, j0: int, j1: int) -> None:
    """Write a strip table (nx*(j1-j0), k[, nz]) into the full table (nx*ny, k[, nz]): node = jj*nx + ii."""
    full[j0 * nx:j1 * nx, ...] = part


def gather_tables(local: Dict[str, np.ndarray], nx: int, ny: int, strips: List[int], rank: int, group=None,
                  device=None) -> Dict[str, np.ndarray]:
    """All-gather the depth-kernel tables computed on the strips.  `local` maps table name ->
    array (nx*(j1-j0), k[, nz]) for this rank's strip; returns the full (nx*ny, k[, nz]) tables
    (numpy, Fortran order).  Bit-identical for any world size: no arithmetic crosses ranks."""
    import torch
    import torch.distributed as dist
    out = {}
    for name in sorted(local):
        part = local[name]
        # move the node axis last so that a strip is a contiguous block: shape (k[, nz], nodes)
        flat = torch.from_numpy(np.ascontiguousarray(np.moveaxis(part, 0, -1)).reshape(-1))
        if device is not None:
            flat = flat.to(device)
        allv, sizes = _all_gather_v(flat, dist, group)
        allv = allv.cpu().numpy()
        tail = part.shape[1:]
        full = np.zeros((nx * ny,) + tail, part.dtype, order="F")
        o = 0
        for r, n in enumerate(sizes):
            j0, j1 = strips[r], strips[r + 1]
            blk = allv[o:o + n].reshape(tail + (nx * (j1 - j0),))
            place_strip(full, np.moveaxis(blk, -1, 0), nx, j0, j1)
            o += n
        out[name] = full
    return out

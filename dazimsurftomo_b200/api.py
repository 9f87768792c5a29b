"""Host-side mirror of the reference's forward-modelling subroutines over the C ABI
(include/dazim_b200.h -> libdazim_b200.so).  Names and argument meaning follow the
reference:

  depthkernel          src/src_inv_iso_joint/CalSurfG.f90:1
  depthkernelTI        src/src_forward/depthkernelTI.f90:2
  FwdObsTraveltimeCPS  src/src_forward/FwdTraveltimeCPS.f90:208
  CalSurfG             src/src_inv_iso_joint/CalSurfG.f90:909
  CalSurfGAnisoJoint   src/src_inv_iso_joint/CalSurfGAniso_Joint.f90:209

There is no CPU fallback: if the CUDA library is missing or no device is
present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build
from .formats import Survey

_lib = None


class DazimError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dazim_b200 error {code}: {msg}")
        self.code = code


class Problem(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("vels", C.c_void_p),
                ("goxd", C.c_float), ("gozd", C.c_float), ("dvxd", C.c_float), ("dvzd", C.c_float),
                ("kmaxRc", C.c_int), ("tRc", C.c_void_p), ("depz", C.c_void_p), ("minthk", C.c_float),
                ("kmax", C.c_int), ("nsrc", C.c_int), ("nrcf", C.c_int),
                ("periods", C.c_void_p), ("nrc1", C.c_void_p), ("nsrcsurf1", C.c_void_p),
                ("scxf", C.c_void_p), ("sczf", C.c_void_p), ("rcxf", C.c_void_p), ("rczf", C.c_void_p)]


class Tables(C.Structure):
    _fields_ = [("pvRc", C.c_void_p), ("sen_vs", C.c_void_p), ("sen_vp", C.c_void_p), ("sen_rho", C.c_void_p),
                ("Lsen_Gsc", C.c_void_p)]


class Coo(C.Structure):
    _fields_ = [("rw", C.c_void_p), ("iw_row", C.c_void_p), ("col", C.c_void_p), ("maxnar", C.c_longlong),
                ("nar", C.c_longlong)]


class Times(C.Structure):
    _fields_ = [("kernels_ms", C.c_float), ("dice_ms", C.c_float), ("fmm_ms", C.c_float), ("trace_ms", C.c_float),
                ("assemble_ms", C.c_float), ("total_ms", C.c_float), ("n_accept", C.c_longlong),
                ("n_steps", C.c_longlong), ("n_fmm_launch", C.c_longlong), ("n_trace_launch", C.c_longlong),
                ("n_launch", C.c_longlong), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
                ("rbint", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def library_path() -> str:
    return _build.LIB


def load():
    """Load libdazim_b200.so (must have been built: __graft_entry__.build())."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise DazimError(-1, f"{path} is missing: run `python -m dazimsurftomo_b200.build` (no CPU fallback)")
        lib = C.CDLL(path)
        lib.dazim_strerror.restype = C.c_char_p
        lib.dazim_last_times.restype = C.POINTER(Times)
        lib.dazim_last_times.argtypes = [C.c_void_p]
        lib.dazim_plan_rows.restype = C.c_longlong
        lib.dazim_plan_rows.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        lib.dazim_plan_nnz.restype = C.c_longlong
        lib.dazim_plan_nnz.argtypes = [C.c_void_p]
        lib.dazim_plan_destroy.argtypes = [C.c_void_p]
        lib.dazim_plan_destroy.restype = None
        lib.dazim_destroy.argtypes = [C.c_void_p]
        lib.dazim_destroy.restype = None
        lib.dazim_host_free.argtypes = [C.c_void_p]
        lib.dazim_host_free.restype = None
        lib.dazim_comm_destroy.argtypes = [C.c_void_p]
        lib.dazim_comm_destroy.restype = None
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Pinned:
    """Cache of page-locked host buffers (dazim_host_alloc) handed out as numpy arrays; the multi-GB COO
    outputs are written into these so the device->host copy runs at PCIe speed and repeated calls
    (outer iterations) do not re-fault fresh pages."""

    def __init__(self):
        self.bufs = {}

    def get(self, tag: str, n: int, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        nbytes = max(int(n), 1) * dtype.itemsize
        ptr, cap = self.bufs.get(tag, (None, 0))
        if cap < nbytes:
            if ptr:
                load().dazim_host_free(C.c_void_p(ptr))
            vp = C.c_void_p()
            cap = nbytes + nbytes // 8
            _chk(load().dazim_host_alloc(C.byref(vp), C.c_ulonglong(cap)))
            ptr = vp.value
            self.bufs[tag] = (ptr, cap)
        arr = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(ptr)).view(dtype)
        return arr[:n]


_pinned = _Pinned()


def _chk(code: int):
    if code:
        raise DazimError(code, load().dazim_strerror(C.c_int(code)).decode())


class Handle:
    """One CUDA device + stream (dazim_create)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _chk(load().dazim_create(C.byref(self._h), C.c_int(device)))
        self.device = device

    def close(self):
        if self._h:
            load().dazim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def times(self) -> dict:
        return load().dazim_last_times(self._h).contents.as_dict()


_default: Optional[Handle] = None


def default_handle() -> Handle:
    global _default
    if _default is None:
        dev = int(os.environ.get("LOCAL_RANK", os.environ.get("DAZIM_DEVICE", "0")))
        _default = Handle(dev)
    return _default


class _Prob:
    """Keeps the numpy arrays behind a dazim_problem alive."""

    def __init__(self, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv: Survey):
        self.vels = np.asfortranarray(vels, np.float32)
        nx, ny, nz = self.vels.shape
        self.depz = np.ascontiguousarray(depz, np.float32)
        self.tRc = np.ascontiguousarray(tRc, np.float64)
        self.keep = [np.asfortranarray(x) for x in (sv.periods.astype(np.int32), sv.nrc1.astype(np.int32),
                                                    sv.nsrcsurf1.astype(np.int32), sv.scxf.astype(np.float32),
                                                    sv.sczf.astype(np.float32), sv.rcxf.astype(np.float32),
                                                    sv.rczf.astype(np.float32))]
        p = Problem()
        p.nx, p.ny, p.nz = nx, ny, nz
        p.vels = _p(self.vels)
        p.goxd, p.gozd, p.dvxd, p.dvzd = goxd, gozd, dvxd, dvzd
        p.kmaxRc = len(self.tRc); p.tRc = _p(self.tRc); p.depz = _p(self.depz); p.minthk = minthk
        p.kmax, p.nsrc, p.nrcf = sv.kmax, sv.nsrc, sv.nrcf
        p.periods, p.nrc1, p.nsrcsurf1, p.scxf, p.sczf, p.rcxf, p.rczf = (_p(x) for x in self.keep)
        self.c = p
        self.shape = (nx, ny, nz)
        self.dall = int(sv.dall)


class _Tab:
    def __init__(self, shape, k, tables: Optional[dict]):
        nx, ny, nz = shape
        tables = tables or {}
        g = tables.get
        self.pvRc = np.zeros((nx * ny, k), np.float64, order="F") if g("pvRc") is None else np.asfortranarray(g("pvRc"), np.float64)
        self.sen = [np.zeros((nx * ny, k, nz), np.float64, order="F") if g(n) is None else np.asfortranarray(g(n), np.float64)
                    for n in ("sen_vs", "sen_vp", "sen_rho")]
        self.L = np.zeros((nx * ny, k, nz - 1), np.float32, order="F") if g("Lsen_Gsc") is None else np.asfortranarray(g("Lsen_Gsc"), np.float32)
        t = Tables()
        t.pvRc = _p(self.pvRc); t.sen_vs, t.sen_vp, t.sen_rho = (_p(s) for s in self.sen); t.Lsen_Gsc = _p(self.L)
        self.c = t


def depthkernel(vel, depz, tRc, minthk, handle: Optional[Handle] = None):
    """CalSurfG.f90:1 -> (pvRc, sen_vs, sen_vp, sen_rho) Fortran-ordered float64."""
    h = handle or default_handle()
    vel = np.asfortranarray(vel, np.float32); nx, ny, nz = vel.shape
    depz = np.ascontiguousarray(depz, np.float32); tRc = np.ascontiguousarray(tRc, np.float64); k = len(tRc)
    pv = np.zeros((nx * ny, k), np.float64, order="F")
    s = [np.zeros((nx * ny, k, nz), np.float64, order="F") for _ in range(3)]
    _chk(load().dazim_depthkernel(h._h, C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(vel), _p(pv), _p(s[0]), _p(s[1]),
                                  _p(s[2]), C.c_int(k), _p(tRc), _p(depz), C.c_float(minthk)))
    return pv, s[0], s[1], s[2]


def depthkernelTI(vel, depz, tRc, minthk, handle: Optional[Handle] = None):
    """depthkernelTI.f90:2 -> (pvRc float64 (nx*ny,k), Lsen_Gsc float32 (nx*ny,k,nz-1))."""
    h = handle or default_handle()
    vel = np.asfortranarray(vel, np.float32); nx, ny, nz = vel.shape
    depz = np.ascontiguousarray(depz, np.float32); tRc = np.ascontiguousarray(tRc, np.float64); k = len(tRc)
    pv = np.zeros((nx * ny, k), np.float64, order="F")
    L = np.zeros((nx * ny, k, nz - 1), np.float32, order="F")
    _chk(load().dazim_depthkernel_ti(h._h, C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(vel), _p(pv), C.c_int(k),
                                     _p(tRc), _p(depz), C.c_float(minthk), _p(L)))
    return pv, L


def surfdisp96(thk, vp, vs, rho, periods, handle: Optional[Handle] = None):
    """surfdisp96.f:52 for nprof profiles: inputs (nlayer,nprof) or (nlayer,), returns cg (kmax,nprof)."""
    h = handle or default_handle()
    arrs = []
    for x in (thk, vp, vs, rho):
        x = np.asarray(x, np.float32)
        if x.ndim == 1:
            x = x[:, None]
        arrs.append(np.asfortranarray(x))
    nl, npf = arrs[0].shape
    t = np.ascontiguousarray(periods, np.float64)
    cg = np.zeros((len(t), npf), np.float64, order="F")
    _chk(load().dazim_surfdisp96(h._h, C.c_int(npf), C.c_int(nl), *[_p(a) for a in arrs], C.c_int(len(t)), _p(t), _p(cg)))
    return cg


def _gbuild(mode, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, gc, gs, tables, maxnar, handle, devices=None):
    h = None if devices is not None else (handle or default_handle())
    pr = _Prob(vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv)
    nx, ny, nz = pr.shape
    k = len(pr.tRc)
    tb = _Tab(pr.shape, k, tables)
    dall = pr.dall
    dsurf = np.zeros(dall, np.float32); taa = np.zeros(dall, np.float32)
    tRcV = np.zeros(((nx - 2) * (ny - 2), k), np.float64, order="F")
    if gc is not None:
        gc = np.asfortranarray(gc, np.float32); gs = np.asfortranarray(gs, np.float32)
    coo = Coo()
    if mode != 0:
        if maxnar is None:
            maxnar = max(1024, dall * 400 * (3 if mode == 2 else 1))
        # outputs land in cached page-locked buffers: valid until the next call of the same routine
        rw = _pinned.get("rw", maxnar, np.float32); iw = _pinned.get("iw", maxnar, np.int32)
        col = _pinned.get("col", maxnar, np.int32)
        coo.rw, coo.iw_row, coo.col, coo.maxnar = _p(rw), _p(iw), _p(col), maxnar
    if devices is not None:
        # single process, several GPUs behind ONE C-ABI call (what a Fortran / C host gets): dazim_gbuild_multi
        dv = np.ascontiguousarray(devices, np.int32)
        tm = Times()
        _chk(load().dazim_gbuild_multi(C.c_int(len(dv)), _p(dv), C.c_int(mode), C.byref(pr.c), C.byref(tb.c),
                                       C.c_int(0 if tables is None else 1), _p(gc), _p(gs), _p(dsurf), _p(taa), _p(tRcV),
                                       C.byref(coo) if mode != 0 else None, C.byref(tm)))
        times = tm.as_dict()
    else:
        _chk(load().dazim_gbuild(h._h, C.c_int(mode), C.byref(pr.c), C.byref(tb.c), C.c_int(0 if tables is None else 1),
                                 _p(gc), _p(gs), _p(dsurf), _p(taa), _p(tRcV), C.byref(coo) if mode != 0 else None))
        times = h.times
    out = dict(dsurf=dsurf, obsTaa=taa, tRcV=tRcV, pvRc=tb.pvRc, sen_vs=tb.sen[0], sen_vp=tb.sen[1], sen_rho=tb.sen[2],
               Lsen_Gsc=tb.L, times=times)
    if mode != 0:
        n = coo.nar
        out.update(rw=rw[:n], row=iw[:n], col=col[:n], nar=n)
    return out


def FwdObsTraveltimeCPS(vels, Gctrue, Gstrue, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv: Survey, tables=None,
                        handle: Optional[Handle] = None, devices=None):
    """FwdTraveltimeCPS.f90:208: returns dict(dsurf=T_iso, obsTaa=T_aa, tRcV, Lsen_Gsc, pvRc).
    devices=[0, 1, ...]: one call, several GPUs (dazim_gbuild_multi)."""
    return _gbuild(0, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, Gctrue, Gstrue, tables, None, handle, devices)


def CalSurfG(vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv: Survey, tables=None, maxnar=None,
             handle: Optional[Handle] = None, devices=None):
    """CalSurfG.f90:909: returns dict(dsurf, rw, row (1-based, = iw(2:nar+1)), col (1-based), nar)."""
    return _gbuild(1, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, None, None, tables, maxnar, handle, devices)


def CalSurfGAnisoJoint(vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv: Survey, tables=None, maxnar=None,
                       handle: Optional[Handle] = None, devices=None):
    """CalSurfGAniso_Joint.f90:209: COO over [dVs | Gc | Gs] (3*nparpi columns)."""
    return _gbuild(2, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, None, None, tables, maxnar, handle, devices)


class LsmrInfo(C.Structure):
    _fields_ = [("istop", C.c_int), ("itn", C.c_int), ("normA", C.c_float), ("condA", C.c_float), ("normr", C.c_float),
                ("normAr", C.c_float), ("normx", C.c_float), ("setup_ms", C.c_float), ("solve_ms", C.c_float)]


def LSMR(m, n, row, col, rw, b, damp=0.0, atol=1e-5, btol=1e-4, conlim=200.0, itnlim=500, localSize=10,
         handle: Optional[Handle] = None):
    """lsmrModule.f90:36 on the reference's COO system (row = iw(2:nar+1), col = iw(nar+2:), rw; 1-based).
    Defaults are the joint-inversion controls of Main_Jt.f90:548-553.  Returns (x, info dict)."""
    h = handle or default_handle()
    row = np.ascontiguousarray(row, np.int32); col = np.ascontiguousarray(col, np.int32)
    rw = np.ascontiguousarray(rw, np.float32); b = np.ascontiguousarray(b, np.float32)
    x = np.zeros(n, np.float32)
    info = LsmrInfo()
    _chk(load().dazim_lsmr(h._h, C.c_int(m), C.c_int(n), C.c_longlong(len(rw)), _p(row), _p(col), _p(rw), _p(b),
                           C.c_float(damp), C.c_float(atol), C.c_float(btol), C.c_float(conlim), C.c_int(itnlim),
                           C.c_int(localSize), _p(x), C.byref(info)))
    return x, {k: getattr(info, k) for k, _ in info._fields_}


COMM_ID_BYTES = 128


class Comm:
    """Communicator of the row-distributed solver (dazim_comm_create): one process per GPU.  `Comm.from_torch()` takes
    rank / world size from an initialised torch.distributed group (any backend) and broadcasts rank 0's NCCL id over
    it; `Comm(device, id_bytes, rank, nranks)` is the plain form for callers with their own launcher."""

    def __init__(self, device: int, id_bytes: bytes, rank: int, nranks: int):
        if len(id_bytes) != COMM_ID_BYTES:
            raise ValueError("communicator id must have %d bytes" % COMM_ID_BYTES)
        self._c = C.c_void_p()
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(id_bytes)
        _chk(load().dazim_comm_create(C.c_int(device), buf, C.c_int(rank), C.c_int(nranks), C.byref(self._c)))
        self.rank, self.nranks, self.device = rank, nranks, device

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * COMM_ID_BYTES)()
        _chk(load().dazim_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def broadcast_id(cls) -> bytes:
        """Rank 0's NCCL id on every rank of the initialised torch.distributed group (works over gloo and nccl)."""
        import torch.distributed as dist
        box = [cls.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    @classmethod
    def from_torch(cls, device: int):
        import torch.distributed as dist
        return cls(device, cls.broadcast_id(), dist.get_rank(), dist.get_world_size())

    def close(self):
        if getattr(self, "_c", None) is not None and self._c:
            load().dazim_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def split_rows(m: int, nranks: int):
    """Contiguous row blocks of an m-row system for the row-distributed solve: [(first, count)] per rank (0-based)."""
    base, extra = divmod(m, nranks)
    out, first = [], 0
    for r in range(nranks):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def row_block(row, col, rw, b, first: int, count: int):
    """The triplets and right-hand side of rows first+1 .. first+count (1-based ids in `row`), re-numbered from 1:
    what one rank passes to LSMR_rows."""
    row = np.asarray(row); sel = (row > first) & (row <= first + count)
    return (np.ascontiguousarray(row[sel] - first, np.int32), np.ascontiguousarray(np.asarray(col)[sel], np.int32),
            np.ascontiguousarray(np.asarray(rw)[sel], np.float32), np.ascontiguousarray(np.asarray(b)[first:first + count], np.float32))


def LSMR_rows(comm: Comm, m_local, m_total, n, row, col, rw, b, damp=0.0, atol=1e-5, btol=1e-4, conlim=200.0, itnlim=500,
              localSize=10, handle: Optional[Handle] = None):
    """lsmrModule.f90:36 with the rows of A spread over the ranks of `comm` (dazim_lsmr_rows): this rank's m_local rows
    (local 1-based ids) and b entries in, the replicated solution out.  Collective."""
    h = handle or default_handle()
    row = np.ascontiguousarray(row, np.int32); col = np.ascontiguousarray(col, np.int32)
    rw = np.ascontiguousarray(rw, np.float32); b = np.ascontiguousarray(b, np.float32)
    x = np.zeros(n, np.float32)
    info = LsmrInfo()
    _chk(load().dazim_lsmr_rows(h._h, comm._c, C.c_int(m_local), C.c_longlong(m_total), C.c_int(n), C.c_longlong(len(rw)),
                                _p(row), _p(col), _p(rw), _p(b), C.c_float(damp), C.c_float(atol), C.c_float(btol),
                                C.c_float(conlim), C.c_int(itnlim), C.c_int(localSize), _p(x), C.byref(info)))
    return x, {k: getattr(info, k) for k, _ in info._fields_}


def fmm_solve(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, handle: Optional[Handle] = None):
    """Test seam: eikonal fields of n sources on one phase-velocity map."""
    h = handle or default_handle()
    scx = np.ascontiguousarray(scx, np.float32); scz = np.ascontiguousarray(scz, np.float32)
    n = len(scx)
    nnx = (nx - 3) * 5 + 1; nnz = (ny - 3) * 5 + 1
    pv = np.ascontiguousarray(pv, np.float64)
    veln = np.zeros((nnz, nnx), np.float32, order="F")
    ttn = np.zeros((nnz, nnx, n), np.float32, order="F"); nsts = np.zeros((nnz, nnx, n), np.int32, order="F")
    ttnr = np.zeros((129, 129, n), np.float32, order="F"); nstsr = np.zeros((129, 129, n), np.int32, order="F")
    geom = np.zeros((8, n), np.int32, order="F")
    _chk(load().dazim_fmm_solve(h._h, C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                                C.c_float(dvzd), _p(pv), C.c_int(n), _p(scx), _p(scz), _p(veln), _p(ttn), _p(nsts),
                                _p(ttnr), _p(nstsr), _p(geom)))
    return dict(veln=veln, ttn=ttn, nsts=nsts, ttnr=ttnr, nstsr=nstsr, geom=geom, times=h.times)


def fmm_host_twin(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, hcap=448, hspill=None, ahead=None):
    """TEST SEAM (no device needed): the thread-per-solve eikonal code of csrc/dazim_tps.h -- the functions the CUDA kernel
    k_fmm_tps runs, compiled __host__ __device__ -- executed on the CPU for one source.  Same outputs as fmm_solve(n=1).
    Only tests call this: it is how the kernel's logic is checked against the oracle without a GPU.
    ahead = 0 / 1: replay the cohort kernel's "records computed ahead" protocol (records of the predicted next node gathered
    before / after the current updates are applied); the result then also holds `ahead_stats` = (rounds predicted, not
    predicted, records patched)."""
    nnx = (nx - 3) * 5 + 1; nnz = (ny - 3) * 5 + 1
    if hspill is None:
        hspill = (8 * (nnx + nnz) + 1024 + 16) & ~1
    pv = np.ascontiguousarray(pv, np.float64)
    ttn = np.zeros((nnz, nnx), np.float32, order="F"); nsts = np.zeros((nnz, nnx), np.int32, order="F")
    ttnr = np.zeros((129, 129), np.float32, order="F"); nstsr = np.zeros((129, 129), np.int32, order="F")
    geom = np.zeros(8, np.int32)
    nacc = C.c_longlong(0)
    if ahead is None:
        _chk(load().dazim_debug_fmm_host_twin(C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                                              C.c_float(dvzd), _p(pv), C.c_float(scx), C.c_float(scz), C.c_int(hcap),
                                              C.c_int(hspill), _p(ttn), _p(nsts), _p(ttnr), _p(nstsr), _p(geom), C.byref(nacc)))
        return dict(ttn=ttn, nsts=nsts, ttnr=ttnr, nstsr=nstsr, geom=geom, n_accept=nacc.value)
    stats = (C.c_longlong * 3)()
    _chk(load().dazim_debug_fmm_host_twin_ahead(C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                                                C.c_float(dvzd), _p(pv), C.c_float(scx), C.c_float(scz), C.c_int(hcap),
                                                C.c_int(hspill), C.c_int(int(ahead)), _p(ttn), _p(nsts), _p(ttnr), _p(nstsr),
                                                _p(geom), C.byref(nacc), stats))
    return dict(ttn=ttn, nsts=nsts, ttnr=ttnr, nstsr=nstsr, geom=geom, n_accept=nacc.value, ahead_stats=tuple(stats))


def raytrace(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, rcx, rcz, azim=True, handle: Optional[Handle] = None):
    """Test seam: travel time + dense Frechet maps of n (source, receiver) pairs."""
    h = handle or default_handle()
    scx, scz, rcx, rcz = (np.ascontiguousarray(x, np.float32) for x in (scx, scz, rcx, rcz))
    n = len(scx)
    pv = np.ascontiguousarray(pv, np.float64)
    shp = (ny, nx, n)
    tt = np.zeros(n, np.float32)
    fdm = np.zeros(shp, np.float32, order="F"); fdmc = np.zeros(shp, np.float32, order="F"); fdms = np.zeros(shp, np.float32, order="F")
    _chk(load().dazim_raytrace(h._h, C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                               C.c_float(dvzd), _p(pv), C.c_int(n), _p(scx), _p(scz), _p(rcx), _p(rcz),
                               C.c_int(int(azim)), _p(tt), _p(fdm), _p(fdmc), _p(fdms)))
    return tt, fdm, fdmc, fdms


class IterParams(C.Structure):
    _fields_ = [("iso_inv", C.c_int), ("weightVs", C.c_float), ("weightGcs", C.c_float), ("damp", C.c_float),
                ("minvel", C.c_float), ("maxvel", C.c_float), ("use_ref_controls", C.c_int), ("atol", C.c_float),
                ("btol", C.c_float), ("conlim", C.c_float), ("itnlim", C.c_int), ("localSize", C.c_int)]


class IterStats(C.Structure):
    _fields_ = [("before", C.c_float * 4), ("after", C.c_float * 4), ("meandeltaT", C.c_float),
                ("mean_weight", C.c_float), ("meanabs_weighted", C.c_float), ("norms", C.c_float * 6),
                ("res2Nm", C.c_float), ("resW2Nm", C.c_float), ("meanabs_Taa", C.c_float), ("meanabs_Tvs", C.c_float),
                ("nar1", C.c_longlong), ("nar", C.c_longlong), ("count3", C.c_int), ("lsmr", LsmrInfo),
                ("step_ms", C.c_float), ("scale_ms", C.c_float)]

    def as_dict(self):
        st = ("meanabs", "std", "rms", "mean")
        d = dict(before=dict(zip(st, map(float, self.before))), after=dict(zip(st, map(float, self.after))),
                 norms=dict(zip(("VsNorm2", "VswNorm2", "GcsNorm2", "GcswNorm2", "Mnorm2", "MwNorm2"),
                                map(float, self.norms))),
                 lsmr={k: getattr(self.lsmr, k) for k, _ in self.lsmr._fields_})
        for k in ("meandeltaT", "mean_weight", "meanabs_weighted", "res2Nm", "resW2Nm", "meanabs_Taa", "meanabs_Tvs",
                  "nar1", "nar", "count3", "step_ms", "scale_ms"):
            d[k] = getattr(self, k)
        return d


def CalDdatSigma(obst, cbst, handle: Optional[Handle] = None):
    """CalSigamNorm.f90:2 -> (sigmaT, meandeltaT)."""
    h = handle or default_handle()
    obst = np.ascontiguousarray(obst, np.float32); cbst = np.ascontiguousarray(cbst, np.float32)
    sig = np.zeros(len(obst), np.float32)
    mean = C.c_float(0)
    _chk(load().dazim_cal_ddat_sigma(h._h, C.c_int(len(obst)), _p(obst), _p(cbst), _p(sig), C.byref(mean)))
    return sig, mean.value


def _tikh(joint, nx, ny, nz, dall, iso_inv, weightGcs, weightVs, handle):
    h = handle or default_handle()
    maxvp = (nx - 2) * (ny - 2) * (nz - 1)
    cap = 7 * 3 * maxvp
    rw = np.zeros(cap, np.float32); row = np.zeros(cap, np.int32); col = np.zeros(cap, np.int32)
    nar = C.c_longlong(0); narvs = C.c_longlong(-1); cnt = C.c_int(0)
    _chk(load().dazim_tikhonov(h._h, C.c_int(joint), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(maxvp), C.c_int(dall),
                               C.byref(nar), _p(rw), _p(row), _p(col), C.byref(narvs), C.byref(cnt), C.c_int(int(iso_inv)),
                               C.c_float(weightGcs), C.c_float(weightVs)))
    n = nar.value
    return dict(rw=rw[:n], row=row[:n], col=col[:n], count3=cnt.value, narVs=narvs.value if joint else None)


def TikhonovRegularization(nx, ny, nz, dall, iso_inv, weightGcs, weightVs, handle: Optional[Handle] = None):
    """TikhRegul.f90:2: the appended triplets (rw, row = iw(2:), col; 1-based) and count3."""
    return _tikh(0, nx, ny, nz, dall, iso_inv, weightGcs, weightVs, handle)


def TikhRegul_joint(nx, ny, nz, dall, weightGcs, weightVs, handle: Optional[Handle] = None):
    """TikhRegul.f90:108: dVs block then Gc and Gs blocks; narVs = entries after the dVs block."""
    return _tikh(1, nx, ny, nz, dall, False, weightGcs, weightVs, handle)


def _iterate(call, shape, nrow, obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel, controls, want_rows):
    """Host-side marshalling shared by Plan.iterate and iterate_device: `call` receives the common argument tail
    (obst, prm, vsf, dv, gcf, gsf, dws, sigmaT, resbst, fwdTvs, fwdTaa, stats)."""
    nx, ny, nz = shape
    maxvp = (nx - 2) * (ny - 2) * (nz - 1)
    n = maxvp if iso_inv else 3 * maxvp
    obst = np.ascontiguousarray(obst, np.float32)
    if len(obst) != nrow:
        raise ValueError("obst must have one entry per row of the system")
    v = np.array(vsf, np.float32, order="F")
    dv = np.zeros(n, np.float32)
    g = (nx - 2, ny - 2, nz - 1)
    gcf = None if iso_inv else np.zeros(g, np.float32, order="F")
    gsf = None if iso_inv else np.zeros(g, np.float32, order="F")
    dws = np.zeros(maxvp, np.float32) if iso_inv else None
    rows = [np.zeros(nrow, np.float32) if want_rows else None for _ in range(4)]
    prm = IterParams()
    prm.iso_inv = int(bool(iso_inv)); prm.weightVs = weightVs; prm.weightGcs = weightGcs; prm.damp = damp
    prm.minvel = minvel; prm.maxvel = maxvel
    prm.use_ref_controls = 1 if controls is None else 0
    if controls is not None:
        prm.atol, prm.btol, prm.conlim = controls["atol"], controls["btol"], controls["conlim"]
        prm.itnlim, prm.localSize = controls["itnlim"], controls["localSize"]
    st = IterStats()
    _chk(call(_p(obst), C.byref(prm), _p(v), _p(dv), _p(gcf), _p(gsf), _p(dws), _p(rows[0]), _p(rows[1]), _p(rows[2]),
              _p(rows[3]), C.byref(st)))
    out = dict(vsf=v, dv=dv, gcf=gcf, gsf=gsf, dws=dws, stats=st.as_dict())
    if want_rows:
        out.update(sigmaT=rows[0], resbst=rows[1], fwdTvs=rows[2], fwdTaa=rows[3])
    return out


def iterate_device(shape, system: dict, obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel,
                   controls: Optional[dict] = None, want_rows=False, handle: Optional[Handle] = None):
    """dazim_iterate_device: the iteration tail on a system held in HBM by the caller -- `system` is what
    partition.assemble_system() makes of the all-gathered row blocks: torch CUDA tensors rowptr (int64, rows+1),
    col / val / row (capacity >= nnz + regularisation entries), dsurf, and the ints nrow, nnz, cap.  The tensors'
    producing stream is synchronised here; val / col / row are modified (rows weighted, regularisation appended)."""
    import torch
    h = handle or default_handle()
    torch.cuda.synchronize(system["val"].device)
    ptr = lambda name: C.c_void_p(system[name].data_ptr())
    nx, ny, nz = shape
    return _iterate(lambda *tail: load().dazim_iterate_device(
        h._h, C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_longlong(system["nrow"]), C.c_longlong(system["nnz"]),
        C.c_longlong(system["cap"]), ptr("rowptr"), ptr("col"), ptr("val"), ptr("row"), ptr("dsurf"), *tail),
        shape, int(system["nrow"]), obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel, controls, want_rows)


class Plan:
    """Device-resident plan (dazim_plan_*): inputs uploaded once, run() leaves results in HBM."""

    def __init__(self, mode, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv: Survey, tables: dict, gc=None,
                 gs=None, src_begin=0, src_end=-1, handle: Optional[Handle] = None):
        self.h = handle or default_handle()
        self.mode = mode
        self._pr = _Prob(vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv)
        self._tb = _Tab(self._pr.shape, len(self._pr.tRc), tables)
        self._gc = None if gc is None else np.asfortranarray(gc, np.float32)
        self._gs = None if gs is None else np.asfortranarray(gs, np.float32)
        self._plan = C.c_void_p()
        _chk(load().dazim_plan_create(self.h._h, C.c_int(mode), C.byref(self._pr.c), C.byref(self._tb.c),
                                      _p(self._gc), _p(self._gs), C.c_longlong(src_begin), C.c_longlong(src_end),
                                      C.byref(self._plan)))
        r0 = C.c_longlong(0)
        self.rows = load().dazim_plan_rows(self._plan, C.byref(r0))
        self.row0 = r0.value
        self.h2d_bytes = self.h.times["h2d_bytes"]

    def run(self) -> dict:
        _chk(load().dazim_plan_run(self._plan))
        return self.h.times

    @property
    def nnz(self) -> int:
        return load().dazim_plan_nnz(self._plan)

    @property
    def eikonal_kernel(self) -> str:
        f = load().dazim_plan_eikonal_kernel
        f.restype = C.c_char_p
        f.argtypes = [C.c_void_p]
        return f(self._plan).decode()

    def fetch(self, csr=True, pinned=False):
        """Row block of this plan on the host: dsurf (+ obsTaa) and the CSR arrays.  pinned=True: the arrays are views of
        cached page-locked buffers (valid until the next pinned fetch), so the copy runs at PCIe speed."""
        n = self.rows
        alloc = (lambda tag, cnt, dt: _pinned.get("plan_" + tag, cnt, dt)) if pinned else (lambda tag, cnt, dt: np.zeros(cnt, dt))
        dsurf = alloc("dsurf", n, np.float32)
        taa = alloc("taa", n, np.float32) if self.mode == 0 else None
        rowptr = col = val = None
        if self.mode != 0 and csr == "rowptr":          # row pointers only (the triplets of a big job are tens of GB)
            rowptr = alloc("rowptr", n + 1, np.int64)
        elif self.mode != 0 and csr:
            rowptr = alloc("rowptr", n + 1, np.int64); col = alloc("col", self.nnz, np.int32); val = alloc("val", self.nnz, np.float32)
        _chk(load().dazim_plan_fetch(self._plan, _p(dsurf), _p(taa), _p(rowptr), _p(col), _p(val)))
        return dict(dsurf=dsurf, obsTaa=taa, rowptr=rowptr, col=col, val=val)

    def device_ptrs(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        _chk(load().dazim_plan_device_ptrs(self._plan, *[C.byref(p) for p in ptrs]))
        return dict(zip(("dsurf", "obsTaa", "rowptr", "col", "val"), [p.value for p in ptrs]))

    def lsmr(self, b, damp=0.0, atol=1e-5, btol=1e-4, conlim=200.0, itnlim=500, localSize=10):
        """LSMR on the G row block of the last run, resident in HBM (dazim_plan_lsmr): returns (x, info)."""
        b = np.ascontiguousarray(b, np.float32)
        nx, ny, nz = self._pr.shape
        n = (3 if self.mode == 2 else 1) * (nx - 2) * (ny - 2) * (nz - 1)
        x = np.zeros(n, np.float32)
        info = LsmrInfo()
        _chk(load().dazim_plan_lsmr(self._plan, _p(b), C.c_float(damp), C.c_float(atol), C.c_float(btol), C.c_float(conlim),
                                    C.c_int(itnlim), C.c_int(localSize), _p(x), C.byref(info)))
        return x, {k: getattr(info, k) for k, _ in info._fields_}

    def lsmr_rows(self, comm, m_total, b, damp=0.0, atol=1e-5, btol=1e-4, conlim=200.0, itnlim=500, localSize=10):
        """Row-distributed LSMR over the G row blocks the ranks' plans built for their shares of the sources
        (dazim_plan_lsmr_rows): no gather of G, one n-vector all-reduce per iteration.  b = this rank's rows."""
        b = np.ascontiguousarray(b, np.float32)
        nx, ny, nz = self._pr.shape
        n = (3 if self.mode == 2 else 1) * (nx - 2) * (ny - 2) * (nz - 1)
        x = np.zeros(n, np.float32)
        info = LsmrInfo()
        _chk(load().dazim_plan_lsmr_rows(self._plan, comm._c, C.c_longlong(m_total), _p(b), C.c_float(damp), C.c_float(atol),
                                         C.c_float(btol), C.c_float(conlim), C.c_int(itnlim), C.c_int(localSize), _p(x),
                                         C.byref(info)))
        return x, {k: getattr(info, k) for k, _ in info._fields_}

    def iterate_rows(self, comm, dall_total, obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel,
                     controls: Optional[dict] = None, want_rows=False):
        """The iteration tail with G left on the ranks that built it (dazim_plan_iterate_rows): this rank's plan covers
        its share of the sources; obst and the per-row results have dall_total entries.  Collective; every rank gets the
        same result."""
        return _iterate(lambda *tail: load().dazim_plan_iterate_rows(self._plan, comm._c, C.c_longlong(dall_total), *tail),
                        self._pr.shape, int(dall_total), obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel,
                        controls, want_rows)

    def update_model(self, vels, tables: dict):
        """New model, same geometry (dazim_plan_update_model): re-uploads vels and the depth-kernel tables."""
        nx, ny, nz = self._pr.shape
        self._vels = np.asfortranarray(vels, np.float32)
        self._tb = _Tab(self._pr.shape, len(self._pr.tRc), tables)
        _chk(load().dazim_plan_update_model(self._plan, _p(self._vels), C.byref(self._tb.c)))

    def iterate(self, obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel, controls: Optional[dict] = None,
                want_rows=False):
        """The rest of one outer iteration of Main_Jt.f90 (:416-727) on the G of the last run, resident in HBM
        (dazim_plan_iterate).  Returns dict(vsf, dv, gcf, gsf, dws, stats[, sigmaT, resbst, fwdTvs, fwdTaa])."""
        return _iterate(lambda *tail: load().dazim_plan_iterate(self._plan, *tail), self._pr.shape, self.rows, obst, vsf,
                        iso_inv, weightVs, weightGcs, damp, minvel, maxvel, controls, want_rows)

    def device_tensors(self):
        """Zero-copy torch views of the last run's outputs in HBM (for NCCL exchanges without a host hop).
        plan.run() has synchronised the library stream, so the views are safe on torch's current stream."""
        import torch
        p = self.device_ptrs()
        dev = "cuda:%d" % self.h.device

        class _View:
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

        def view(name, n, ts, dt):
            if not p[name] or n == 0:
                return torch.zeros(0, dtype=dt, device=dev)
            return torch.as_tensor(_View(p[name], n, ts), device=dev)

        n, nnz = self.rows, self.nnz
        out = dict(dsurf=view("dsurf", n, "<f4", torch.float32))
        if self.mode == 0:
            out["obsTaa"] = view("obsTaa", n, "<f4", torch.float32)
        else:
            out.update(rowptr=view("rowptr", n + 1, "<i8", torch.int64), col=view("col", nnz, "<i4", torch.int32),
                       val=view("val", nnz, "<f4", torch.float32))
        return out

    def close(self):
        if self._plan:
            # a plan frees its buffers on its handle's stream: if the handle is already gone (interpreter shutdown
            # destroys objects in no particular order) the plan is abandoned rather than freed through a dead stream
            if self.h is not None and self.h._h:
                load().dazim_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

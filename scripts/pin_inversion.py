"""Pin the ORACLE's outer inversion loop (oracle/pyoracle.py::invert = Main_Jt.f90:364-750 restated) on the inversion
results the reference ships (run in the build container, where /root/reference exists; minutes of CPU):

  example/test2_syn_iso_inv/plot_script/DSurfTomo.inv      isotropic Vsv after 20 outer iterations
  example/test3_syn_joint_inv/plot_script/Gc_Gs_model.inv  joint dVs + Gc + Gs after 5 outer iterations

and write the small fixtures tests/golden/inv/{test2_iter.npz,test3_iter.npz}: the oracle's model after the first
two iterations (what the GPU parity test reproduces) and its final model next to the reference's shipped one.

    python scripts/pin_inversion.py [test2|test3|both] [maxiter-override]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dazimsurftomo_b200 import formats as fm   # noqa: E402
from oracle import pyoracle as po              # noqa: E402

REF = "/root/reference/example"
OUT = os.path.join(ROOT, "tests", "golden", "inv")


def fixed(path, w, ncol):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.strip():
                rows.append([float(line[i * w:(i + 1) * w]) for i in range(ncol)])
    return np.array(rows)


def run(case, maxiter=None, track=None):
    base = os.path.join(REF, case)
    p = fm.read_para_inv(os.path.join(base, "para.in"))
    depz, vs = fm.read_model(os.path.join(base, "MOD"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(base, p.datafile), p.kmaxRc)
    obst = (sv.dist / sv.obsvel).astype(np.float32)        # Main_Jt.f90:308
    snaps = {}

    def on_iter(it, vsf, gcf, gsf, rec):
        if it <= 2:
            snaps["vsf%d" % it] = vsf.copy(); snaps["gcf%d" % it] = gcf.copy(); snaps["gsf%d" % it] = gsf.copy()
        if track is not None:                      # distance to the shipped model after every outer iteration
            f = lambda a: a.ravel(order="F")
            vsm = f((vsf[1:-1, 1:-1, :-1] + vsf[1:-1, 1:-1, 1:]) / 2)
            print("  [%s] after iteration %d: vs shipped  Vs_mid max %.4f rms %.5f | Gc max %.4f rms %.5f %% | Gs max %.4f rms %.5f %%"
                  % (case, it, np.abs(vsm - track[:, 3]).max(), np.sqrt(((vsm - track[:, 3]) ** 2).mean()),
                     np.abs(f(gcf) * 100 - track[:, 6]).max(), np.sqrt(((f(gcf) * 100 - track[:, 6]) ** 2).mean()),
                     np.abs(f(gsf) * 100 - track[:, 7]).max(), np.sqrt(((f(gsf) * 100 - track[:, 7]) ** 2).mean())), flush=True)

    t0 = time.time()
    r = po.invert(vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, obst, p.iso_mod, p.weightVs,
                  p.weightGcs, p.damp, p.minvel, p.maxvel, maxiter or p.maxiter, spfra=p.spfra,
                  nthreads=os.cpu_count() or 1, log=lambda s: print("  [%s %.0fs] %s" % (case, time.time() - t0, s), flush=True),
                  on_iter=on_iter)
    return p, depz, r, snaps


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    mi = int(sys.argv[2]) if len(sys.argv) > 2 else None
    os.makedirs(OUT, exist_ok=True)
    if which in ("test2", "both"):
        p, depz, r, snaps = run("test2_syn_iso_inv", mi)
        ref = fixed(os.path.join(REF, "test2_syn_iso_inv/plot_script/DSurfTomo.inv"), 8, 4)
        true = np.loadtxt(os.path.join(REF, "test2_syn_iso_inv/plot_script/DSurfTomo.true"))
        ours = np.array([r["vsf"][i, j, k] for k in range(p.nz) for j in range(p.ny) for i in range(p.nx)])
        d = ours - ref[:, 3]
        print("test2: oracle vs shipped DSurfTomo.inv: max |dVs| %.4f  rms %.5f  (model perturbation rms vs start %.4f; "
              "shipped-vs-true rms %.4f, oracle-vs-true rms %.4f)" %
              (np.abs(d).max(), np.sqrt((d ** 2).mean()), np.sqrt(((ref[:, 3] - 3.2) ** 2).mean()) if False else
               np.sqrt(((ref[:, 3] - ours.mean()) ** 2).mean()), np.sqrt(((ref[:, 3] - true[:, 3]) ** 2).mean()),
               np.sqrt(((ours - true[:, 3]) ** 2).mean())))
        np.savez_compressed(os.path.join(OUT, "test2_iter.npz"), final=r["vsf"], shipped=ref[:, 3].astype(np.float32),
                            hist=np.array([[h["before"]["rms"], h["after"]["rms"], h["lsmr"]["itn"], h["lsmr"]["istop"]]
                                           for h in r["history"]], np.float32), **snaps)
    if which in ("test3", "both"):
        p, depz, r, snaps = run("test3_syn_joint_inv", mi)
        ref = np.loadtxt(os.path.join(REF, "test3_syn_joint_inv/plot_script/Gc_Gs_model.inv"))
        real = np.loadtxt(os.path.join(REF, "test3_syn_joint_inv/plot_script/Gc_Gs_model.real"))
        nvx, nvz = p.nx - 2, p.ny - 2
        gc = np.array([r["gcf"][i, j, k] for k in range(p.nz - 1) for j in range(nvz) for i in range(nvx)]) * 100
        gs = np.array([r["gsf"][i, j, k] for k in range(p.nz - 1) for j in range(nvz) for i in range(nvx)]) * 100
        vsm = np.array([(r["vsf"][i + 1, j + 1, k] + r["vsf"][i + 1, j + 1, k + 1]) / 2
                        for k in range(p.nz - 1) for j in range(nvz) for i in range(nvx)])
        print("test3: oracle vs shipped Gc_Gs_model.inv: max |dGc| %.4f %%  max |dGs| %.4f %%  max |dVs_mid| %.4f km/s "
              "(shipped amplitude: max |Gc| %.3f %%, max |Gs| %.3f %%; shipped-vs-real rms Gc %.3f, oracle-vs-real rms Gc %.3f)"
              % (np.abs(gc - ref[:, 6]).max(), np.abs(gs - ref[:, 7]).max(), np.abs(vsm - ref[:, 3]).max(),
                 np.abs(ref[:, 6]).max(), np.abs(ref[:, 7]).max(), np.sqrt(((ref[:, 6] - real[:, 6]) ** 2).mean()),
                 np.sqrt(((gc - real[:, 6]) ** 2).mean())))
        np.savez_compressed(os.path.join(OUT, "test3_iter.npz"), final_vsf=r["vsf"], final_gcf=r["gcf"],
                            final_gsf=r["gsf"], shipped=ref[:, [3, 6, 7]].astype(np.float32),
                            hist=np.array([[h["before"]["rms"], h["after"]["rms"], h["lsmr"]["itn"], h["lsmr"]["istop"]]
                                           for h in r["history"]], np.float32), **snaps)


def main_test4(mi=None):
    """example/test4_Yunnan (real data, 38x42x18, 86 refined layers, joint, 5 outer iterations) vs the shipped
    plot_script/Gc_Gs_model.inv and period_Azm_tomo.inv."""
    ref = np.loadtxt(os.path.join(REF, "test4_Yunnan/plot_script/Gc_Gs_model.inv"))
    p, depz, r, snaps = run("test4_Yunnan", mi, track=ref)
    nvx, nvz = p.nx - 2, p.ny - 2
    f = lambda a: a.ravel(order="F")
    gc = f(r["gcf"]) * 100; gs = f(r["gsf"]) * 100
    v = r["vsf"]
    vsm = f((v[1:-1, 1:-1, :-1] + v[1:-1, 1:-1, 1:]) / 2)
    print("test4: oracle vs shipped Gc_Gs_model.inv: max |dVs_mid| %.4f km/s (rms %.5f), max |dGc| %.4f %% (rms %.5f), "
          "max |dGs| %.4f %% (rms %.5f); shipped max |Gc| %.3f %%, |Gs| %.3f %%" %
          (np.abs(vsm - ref[:, 3]).max(), np.sqrt(((vsm - ref[:, 3]) ** 2).mean()), np.abs(gc - ref[:, 6]).max(),
           np.sqrt(((gc - ref[:, 6]) ** 2).mean()), np.abs(gs - ref[:, 7]).max(), np.sqrt(((gs - ref[:, 7]) ** 2).mean()),
           np.abs(ref[:, 6]).max(), np.abs(ref[:, 7]).max()))
    np.savez_compressed(os.path.join(OUT, "test4_iter.npz"), final_vsf=r["vsf"], final_gcf=r["gcf"], final_gsf=r["gsf"],
                        shipped=ref[:, [3, 6, 7]].astype(np.float32),
                        hist=np.array([[h["before"]["rms"], h["after"]["rms"], h["lsmr"]["itn"], h["lsmr"]["istop"]]
                                       for h in r["history"]], np.float32))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "test4":
        main_test4(int(sys.argv[2]) if len(sys.argv) > 2 else None)
        sys.exit(0)
    main()

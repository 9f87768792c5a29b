"""Generate tests/golden/ fixtures from the reference's shipped example (run in the build
container, where /root/reference exists).  The fixtures are the reference's OWN inputs and
outputs (example/test1_syn_foward), cut down so they stay small:

  test1/para.in, MODVs.true, MODGc.true, MODGs.true      verbatim inputs
  test1/surfdata_subset.dat                               '#' blocks of a few (period, source) units
  test1/surfphase_subset.dat                              the matching blocks of the reference's
                                                          output/surfphase_forward_RV3th.dat
  test1/period_Azm_tomo.npz                               output/period_Azm_tomo.real as an array
  inv/test2_para.in, test2_MOD, test3_para.in, test3_MOD  verbatim inputs of the two inversion examples
  inv/surfphase_forward_RV3th.dat.xz                      the data file of test2 and test3 (identical), xz -9
  inv/test4_para.in, test4_MOD.xz, test4_data.dat.xz      example/test4_Yunnan inputs (real data), xz -9
  (inv/test2_iter.npz, test3_iter.npz are written by scripts/pin_inversion.py)
"""
import os
import shutil
import sys

import numpy as np

REF = "/root/reference/example/test1_syn_foward"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "test1")
PERIODS = (1, 2, 3, 4)          # T = 5..8 s: the span on which the shipped outputs are mutually consistent
SOURCES_PER_PERIOD = 5      # first, then every 29th source


def blocks(path):
    cur = None
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                if cur:
                    yield cur
                cur = [line]
            elif line.strip():
                cur.append(line)
    if cur:
        yield cur


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in ("para.in", "MODVs.true", "MODGc.true", "MODGs.true"):
        shutil.copy(os.path.join(REF, f), os.path.join(OUT, f))
        os.chmod(os.path.join(OUT, f), 0o644)
    inp = list(blocks(os.path.join(REF, "Surfphase_RV3_5_40s_1s.dat")))
    gold = list(blocks(os.path.join(REF, "output", "surfphase_forward_RV3th.dat")))
    assert len(inp) == len(gold) == 4320
    seen = {}
    keep = []
    for i, b in enumerate(inp):
        per = int(b[0].split()[3])
        k = seen.get(per, 0)
        seen[per] = k + 1
        if per in PERIODS and k % 29 == 0 and k // 29 < SOURCES_PER_PERIOD:
            keep.append(i)
    with open(os.path.join(OUT, "surfdata_subset.dat"), "w") as fi, open(os.path.join(OUT, "surfphase_subset.dat"), "w") as fg:
        for i in keep:
            assert len(inp[i]) == len(gold[i])
            fi.writelines(inp[i])
            fg.writelines(gold[i])
    g = np.loadtxt(os.path.join(REF, "output", "period_Azm_tomo.real"))
    np.savez_compressed(os.path.join(OUT, "period_Azm_tomo.npz"), table=g.astype(np.float32))
    # inversion examples: control files and start models only (the data file of test2/test3 is test1's output)
    inv = os.path.join(OUT, "..", "inv")
    os.makedirs(inv, exist_ok=True)
    for case, tag in (("test2_syn_iso_inv", "test2"), ("test3_syn_joint_inv", "test3")):
        for f in ("para.in", "MOD"):
            dst = os.path.join(inv, "%s_%s" % (tag, f))
            shutil.copy(os.path.join(REF, "..", case, f), dst)
            os.chmod(dst, 0o644)
    # the data file of both inversion examples (byte-identical in test2 and test3; = test1's shipped output), xz'd:
    # with it the GPU box can run BASELINE configs 3 and 4 for real and compare with the shipped models
    import filecmp
    import lzma
    d2 = os.path.join(REF, "..", "test2_syn_iso_inv", "surfphase_forward_RV3th.dat")
    assert filecmp.cmp(d2, os.path.join(REF, "..", "test3_syn_joint_inv", "surfphase_forward_RV3th.dat"), shallow=False)
    with open(d2, "rb") as f, lzma.open(os.path.join(inv, "surfphase_forward_RV3th.dat.xz"), "wb", preset=9) as g:
        g.write(f.read())
    # example/test4_Yunnan (real data; BASELINE config 5's real counterpart): control file verbatim, MOD and data xz'd
    t4 = os.path.join(REF, "..", "test4_Yunnan")
    shutil.copy(os.path.join(t4, "para.in"), os.path.join(inv, "test4_para.in"))
    os.chmod(os.path.join(inv, "test4_para.in"), 0o644)
    for src, dst in (("MOD", "test4_MOD.xz"), ("China_YN_Rayleigh_RS_5-40s.dat", "test4_data.dat.xz")):
        with open(os.path.join(t4, src), "rb") as f, lzma.open(os.path.join(inv, dst), "wb", preset=9) as g:
            g.write(f.read())
    print("units kept:", len(keep), "rays:", sum(len(inp[i]) - 1 for i in keep))


if __name__ == "__main__":
    sys.exit(main())

import os

import numpy as np

from dazimsurftomo_b200 import formats as fm
from conftest import GOLD


def test_para_forward():
    p = fm.read_para_forward(os.path.join(GOLD, "para.in"))
    assert (p.nx, p.ny, p.nz) == (17, 17, 4)
    assert p.kmaxRc == 36 and p.tRc[0] == 5.0 and p.tRc[-1] == 40.0
    assert p.sublayers == 2.0 and p.writepath is False and p.noiselevel == 0.0
    assert p.datafile == "Surfphase_RV3_5_40s_1s.dat"


def test_model_layout(test1):
    vs = test1["vs"]
    assert vs.shape == (17, 17, 4) and vs.flags["F_CONTIGUOUS"]
    assert np.allclose(test1["depz"], [0, 10, 35, 60])
    # MODVs.true is 3.2 / 3.4 / 3.8 / 4.2 +- checkerboard
    assert abs(float(vs[1, 1, 0]) - 3.2) < 1e-6
    assert test1["gc"].shape == (15, 15, 3)


def test_float32_colatitude_round_trip(test1):
    """MainForward.f90:254-255,410-411: 23.3 is echoed as 23.300011 by the reference."""
    sv = test1["sv"]
    lat = np.float32(90.0) - sv.scxf[0, 0] * np.float32(180.0) / fm.PI32
    lon = sv.sczf[0, 0] * np.float32(180.0) / fm.PI32
    assert "%11.6f" % lat == "  23.300011"
    assert "%11.6f" % lon == " 101.550003"


def test_survey_tables(test1):
    sv = test1["sv"]
    assert sv.kmax == 36 and sv.dall == sum(int(x) for x in sv.nrc1.ravel())
    assert int(sv.nsrcsurf1[0]) == 5 and int(sv.nsrcsurf1[4]) == 0
    offs = sv.row_offsets()
    assert offs[-1] == sv.dall and len(offs) == 21


def test_writer_round_trip(tmp_path, test1):
    sv = test1["sv"]
    t = (sv.dist / np.float32(3.3)).astype(np.float32)   # any travel times, file order == loop order here
    out = tmp_path / "o.dat"
    fm.write_surfphase_forward(str(out), sv, t)
    c = fm.read_surfphase_velocities(str(out))
    assert len(c) == sv.dall and np.abs(c - 3.3).max() < 2e-5
    first = open(out).readline()
    assert first.startswith("#  23.300011 101.550003  1  2  0")


def test_surfdata_rejects_corrupt_period_blocks(tmp_path):
    """period index outside 1..kmaxRc, a period that re-appears later in the file (the reference would silently
    overwrite the first group, Main_Jt.f90:283-286), a data line before any source line, a non-Rayleigh-phase block."""
    import pytest
    good = "# 23.3 101.55 1 2 0\n23.3 101.84 3.1\n# 23.4 101.55 2 2 0\n23.3 101.84 3.2\n"
    p = tmp_path / "d.dat"
    p.write_text(good)
    sv = fm.read_surfdata(str(p), 3)
    assert sv.dall == 2 and list(sv.nsrcsurf1) == [1, 1, 0]
    for bad, msg in ((good.replace(" 2 2 0", " 0 2 0"), "outside 1..3"), (good.replace(" 2 2 0", " 4 2 0"), "outside 1..3"),
                     (good + "# 23.5 101.55 1 2 0\n23.3 101.84 3.3\n", "two separate groups"),
                     ("23.3 101.84 3.1\n" + good, "before the first"), (good.replace(" 1 2 0", " 1 1 0"), "Rayleigh")):
        p.write_text(bad)
        with pytest.raises(ValueError, match=msg):
            fm.read_surfdata(str(p), 3)


def test_bench_slice_is_a_proportional_slice():
    """bench.py's CPU legs time a slice that covers EVERY period with the same fraction of sources (no extrapolation
    from one period), rotated from step to step."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from dazimsurftomo_b200 import synthetic
    w = synthetic.s200(src_per_period=40, n=22, nz=4, nsta=60, nrec=6, kmax=5)
    a = bench.subset_survey(w.sv, 8, 0)
    b = bench.subset_survey(w.sv, 8, 7)
    assert list(a.nsrcsurf1) == [8] * 5 and a.dall == 8 * 5 * 6 and a.rcxf.shape == (6, 8, 5)
    assert not np.array_equal(a.scxf, b.scxf)                       # another step, other sources
    src = {(float(x), float(z)) for x, z in zip(w.sv.scxf[:, 2], w.sv.sczf[:, 2])}
    assert all((float(x), float(z)) in src for x, z in zip(a.scxf[:, 2], a.sczf[:, 2]))
    cfg1 = bench.config_of(w, 1); cfg2 = bench.config_of(w, 1)
    assert cfg1 == cfg2 and cfg1["solves"] == w.n_solves and cfg1["workload"] == w.name

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

"""EXPERIMENT kept honest (oracle/fim_experiment.cpp, scripts/fim_vs_fmm.py; CPU only): an order-free fixed point of the
reference's local eikonal solver -- what a fast-iterative-method kernel converges to -- agrees with the reference's heap
fast-marching field to a few float32 ulps, but not bit for bit; DESIGN.md section 5 quotes what that does to G."""
import numpy as np


def test_fixed_point_is_close_to_but_not_the_heap_march(oracle, test1, test1_tables):
    p = test1["para"]; sv = test1["sv"]
    pv = np.ascontiguousarray(test1_tables["pvRc"][:, 0])
    srcs = [(float(sv.scxf[s, 0]), float(sv.sczf[s, 0])) for s in range(int(sv.nsrcsurf1[0]))]
    differ = 0
    for scx, scz in srcs:
        a = oracle.fmm_source(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, scx, scz)["ttn"]
        b, passes = oracle.fmm_source_fim(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, scx, scz)
        assert passes > 0                                    # the overwrite phase reached a bitwise fixed point
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-6)
        assert rel.max() < 1e-5                              # travel times: inside the north star's tolerance ...
        differ += int((a != b).any())
    assert differ >= 1                                       # ... but not the reference's field (so not its G pattern)


def test_default_path_ignores_the_experiment_switch(oracle, test1, test1_tables, monkeypatch):
    """The library the tests check against and bench.py times (liboracle.so) has no experiment code in it at all:
    ORC_FIM_EXPERIMENT=1 changes nothing there; it only acts in liboracle_experiments.so."""
    p = test1["para"]
    import copy
    sv = copy.copy(test1["sv"])
    ns = np.zeros_like(sv.nsrcsurf1); ns[0] = 2
    sv.nsrcsurf1 = ns; sv.dall = int(sv.nrc1[:2, 0].sum())
    args = (0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, test1["gc"], test1["gs"])
    monkeypatch.delenv("ORC_FIM_EXPERIMENT", raising=False)
    a = oracle.gbuild(*args, tables=test1_tables)
    monkeypatch.setenv("ORC_FIM_EXPERIMENT", "1")
    b = oracle.gbuild(*args, tables=test1_tables)
    assert np.array_equal(a["dsurf"], b["dsurf"]) and np.array_equal(a["obsTaa"], b["obsTaa"])
    c = oracle.gbuild(*args, tables=test1_tables, experiments=True)
    assert not np.array_equal(a["dsurf"], c["dsurf"])             # the switch is live only in the experiments build
    assert np.abs(a["dsurf"] - c["dsurf"]).max() < 1e-5 * a["dsurf"].max()


def test_values_are_a_local_function_of_the_acceptance_order(oracle, test1, test1_tables):
    """ORDER EXPERIMENT (scripts/order_experiment.py): the rule 'a node's value = the quadrant solver at the acceptance
    of its last direct neighbour before its own pop, with the nodes accepted up to then alive' reproduces the
    reference's coarse march bit for bit, a replay fed with ranks predicted from the order-free fixed point does too on
    this model, and the local hazard checks never pass a wrong replay."""
    p = test1["para"]; sv = test1["sv"]
    for k in (0, 20):
        pv = np.ascontiguousarray(test1_tables["pvRc"][:, k])
        for s in range(int(sv.nsrcsurf1[0])):
            r = oracle.fmm_order_stats(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, float(sv.scxf[s, 0]), float(sv.sczf[s, 0]))
            assert r["popped"] > 4000 and r["rule_mismatch"] == 0
            assert r["sorted_exact_mismatch"] == 0 and r["sorted_fim_mismatch"] == 0 and r["sorted_fim_rank_errors"] == 0
            assert 50 < r["dag_levels"] < 200                    # ~40-60 nodes per wavefront on a 71 x 71 grid
            flagged = r["verify_order_flags"] + r["verify_key_increase_flags"] > 0
            assert flagged or r["sorted_fim_mismatch"] == 0      # soundness: wrong => flagged


def test_order_rule_holds_on_the_refined_source_box(oracle, test1, test1_tables):
    """Same rule on the refined box, with its stopping rule (the node the march stops at is alive but never updates its
    neighbours) and the trial values / statuses of the close nodes it hands to the coarse grid: bit for bit."""
    p = test1["para"]; sv = test1["sv"]
    pv = np.ascontiguousarray(test1_tables["pvRc"][:, 7])
    for s in range(int(sv.nsrcsurf1[0])):
        for prefix in (0, 4):
            r = oracle.fmm_order_stats(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, float(sv.scxf[s, 0]), float(sv.sczf[s, 0]),
                                       refined=True, prefix=prefix)
            assert r["popped"] > 3000 and r["rule_mismatch"] == 0
            flagged = r["verify_order_flags"] + r["verify_key_increase_flags"] > 0
            assert flagged or r["sorted_fim_mismatch"] == 0      # soundness of the local checks


def test_rank_iteration_from_a_distance_guess_reaches_the_reference_field(oracle, test1, test1_tables):
    """RANK ITERATION (scripts/rank_iteration_experiment.py): starting from the order of plain geometric distance to the
    source, ranks -> replay -> sort -> ranks settles within a few rounds on the reference's coarse field, bit for bit."""
    p = test1["para"]; sv = test1["sv"]
    pv = np.ascontiguousarray(test1_tables["pvRc"][:, 10])
    nnx = (p.nx - 3) * 5 + 1; nnz = (p.ny - 3) * 5 + 1
    gox = (90 - p.goxd) * np.pi / 180; goz = p.gozd * np.pi / 180
    x = gox + np.arange(nnx) * (p.dvxd * np.pi / 180 / 5); z = goz + np.arange(nnz) * (p.dvzd * np.pi / 180 / 5)
    X, Z = np.meshgrid(x, z)
    for s in range(int(sv.nsrcsurf1[0])):
        scx, scz = float(sv.scxf[s, 0]), float(sv.sczf[s, 0])
        guess = np.hypot(X - scx, (Z - scz) * np.sin(X)).astype(np.float32)
        r = oracle.fmm_rank_iteration(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, scx, scz, guess, prefix=64)
        assert r["popped"] > 4000 and r["mismatch"] == 0 and r["flags"] == 0 and r["rounds"] <= 12
        rr = oracle.fmm_order_stats(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, scx, scz, refined=True, prefix=1004)
        assert rr["rule_mismatch"] == 0 and rr["fim_passes"] <= 12
        flagged = rr["verify_order_flags"] + rr["verify_key_increase_flags"] > 0
        assert flagged or rr["sorted_fim_mismatch"] == 0

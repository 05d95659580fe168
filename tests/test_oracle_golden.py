"""Known-answer tests that pin the oracle to what the reference's own tests assert.

The reference holds no unit vectors for this path; these are its integration scenarios restated as event traces.
Everything the scenarios do not reach (tie-break chain, preferred cases (a)/(b), rpm filter, UpgradeTracker, inferred
preferences, reaper selection) is "parity unpinned by reference tests" and is covered by brute-force cross-checks in
test_oracle_properties.py instead.
"""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob

MIB = 1024 * 1024
UNIT = 8192


def ev(op, key, weight=0, last_used=0):
    return (op, key, weight, last_used)


def events(lst):
    a = np.zeros(len(lst), dtype=ob.LRU_EVENT)
    for i, e in enumerate(lst):
        a[i] = e
    return a


INSERT, TOUCH, RESIZE, REMOVE, SETCAP = 0, 1, 2, 3, 4


def load_model(key, size_units, t):
    """loadLocal: 1-unit placeholder (INSERTION_WEIGHT, MM:5011,5061) then inflate to the predicted size (MM:2094-2100)."""
    return [ev(INSERT, key, 1, t), ev(RESIZE, key, size_units)]


def test_capacity_constants_evictions_model_mesh_test(oracle_lib):
    # T/EvictionsModelMeshTest.java:30-33: 1024 MiB capacity, 50 MiB default size, 6 loading threads
    #   => unload buffer 75 MiB, effective capacity 949 MiB
    cap_units = 1024 * MIB // UNIT
    default_units = 50 * MIB // UNIT
    reserve = oracle_lib.orc_unload_reserve_units(cap_units, 6, default_units)
    assert reserve == 75 * MIB // UNIT == 9600
    assert (cap_units - reserve) * UNIT == 949 * MIB
    # isFull threshold MM:767-769 with an unload manager: max(6400, min(6400*6, 131072/20)) = 6553
    assert oracle_lib.orc_min_space_units(cap_units, 6, default_units, 1) == 6553
    # DummyModelMesh (T/DummyModelMesh.java:39): 10 x 20 MiB, 20 MiB models, reserve clamps to cap/10 => 9 models fit
    assert oracle_lib.orc_unload_reserve_units(25600, 8, 2560) == 2560


def test_basic_eviction_test_trace(oracle_lib):
    """T/EvictionsModelMeshTest.java:36-128 basicEvictionTest as an LRU event trace."""
    cap = 949 * MIB // UNIT  # effective capacity once the unload buffer entry is set aside
    m50, m160 = 50 * MIB // UNIT, 160 * MIB // UNIT
    lru = ob.OracleLru(cap)
    t = 1_000_000
    trace = []
    for i in range(18):  # :51-62 eighteen 50 MiB models fit
        trace += load_model(i, m50, t + 10 * i)
    assert len(lru.apply(events(trace), now_ms=t)) == 0 and lru.size() == 18
    # :64-83 the 19th takes the total to 950 MiB > 949 and evicts exactly myModel0 (the oldest)
    e = lru.apply(events(load_model(18, m50, t + 200)), now_ms=t)
    assert [int(x["key"]) for x in e] == [0]
    # :85-105 the 20th and 21st evict myModel1 then myModel2
    e = lru.apply(events(load_model(19, m50, t + 210)), now_ms=t)
    assert [int(x["key"]) for x in e] == [1]
    e = lru.apply(events(load_model(20, m50, t + 220)), now_ms=t)
    assert [int(x["key"]) for x in e] == [2]
    # :107-109 re-ensureLoaded(myModel0) evicts myModel3
    e = lru.apply(events(load_model(0, m50, t + 230)), now_ms=t)
    assert [int(x["key"]) for x in e] == [3]
    # :111-123 a 160 MiB model: predicted 50 MiB evicts myModel4, post-load sizing (UpdateTask) evicts 5 and 6, keeps 7
    e = lru.apply(events(load_model(21, m50, t + 240)), now_ms=t)
    assert [int(x["key"]) for x in e] == [4]
    e = lru.apply(events([ev(RESIZE, 21, m160)]), now_ms=t)
    assert [int(x["key"]) for x in e] == [5, 6]
    keys, _, _ = lru.dump()
    assert 7 in keys and 21 in keys and lru.weighted_size() <= cap
    assert lru.oldest_time() == t + 70


def test_concurrent_eviction_exactly_the_ten_oldest(oracle_lib):
    """T/EvictionsModelMeshTest.java:136-200: ten inserts into a full cache evict exactly the ten oldest, no cascade."""
    cap = 949 * MIB // UNIT
    m50 = 50 * MIB // UNIT
    lru = ob.OracleLru(cap)
    t = 5_000_000
    tr = []
    for i in range(18):
        tr += load_model(i, m50, t + i)
    lru.apply(events(tr), now_ms=t)
    tr = []
    for i in range(18, 28):
        tr += load_model(i, m50, t + 100 + i)
    e = lru.apply(events(tr), now_ms=t)
    assert sorted(int(x["key"]) for x in e) == list(range(10))
    assert [int(x["key"]) for x in e] == list(range(10))  # oldest first


def test_multi_load_with_eviction_standalone(oracle_lib):
    """T/ModelMeshEvictionsTest.java:156-187: 10 x 2560-unit capacity, 0.9 factor => 9 fit; loading 9+3 keeps the last 9."""
    lru = ob.OracleLru(25600 - 2560)
    tr = []
    for i in range(12):
        tr += load_model(i, 2560, 1000 + i)
    e = lru.apply(events(tr), now_ms=999)
    assert [int(x["key"]) for x in e] == [0, 1, 2]
    keys, _, _ = lru.dump()
    assert sorted(keys.tolist()) == list(range(3, 12))


def test_multi_load_with_big_eviction_standalone(oracle_lib):
    """:190-228 one 4x-size model displaces four: survivors are ids[6..] of the twelve."""
    lru = ob.OracleLru(25600 - 2560)
    tr = []
    for i in range(11):
        tr += load_model(i, 2560, 1000 + i)
    lru.apply(events(tr), now_ms=999)
    lru.apply(events(load_model(11, 4 * 2560, 2000)), now_ms=999)  # size hint = 4x (KNOWN_SIZE, MM:5160-5163)
    keys, _, _ = lru.dump()
    assert sorted(keys.tolist()) == list(range(6, 12))


def test_multi_load_with_eviction_standalone_reuse(oracle_lib):
    """:240-281 touching the first three (internalOperation(load=true,lastUsed=0) -> now) protects them."""
    lru = ob.OracleLru(25600 - 2560)
    tr = []
    for i in range(9):
        tr += load_model(i, 2560, 1000 + i)
    lru.apply(events(tr), now_ms=1500)
    lru.apply(events([ev(TOUCH, 0, 0, 0), ev(TOUCH, 1, 0, 0), ev(TOUCH, 2, 0, 0)]), now_ms=2000)
    tr = []
    for i in range(9, 12):
        tr += load_model(i, 2560, 3000 + i)
    e = lru.apply(events(tr), now_ms=3000)
    assert [int(x["key"]) for x in e] == [3, 4, 5]
    keys, _, _ = lru.dump()
    assert {0, 1, 2, 9, 10, 11} <= set(keys.tolist())


def _row(**kw):
    r = np.zeros(1, dtype=ob.INST)
    r["capacity"], r["lru_time"], r["l_threads"], r["active"], r["vers"] = 131072, (1 << 63) - 1, 6, 1, 1
    for k, v in kw.items():
        r[k] = v
    return r[0]


def test_error_propagation_test_type_constraint(oracle_lib):
    """T/ModelMeshErrorPropagationTest.java:52-62,89-121: type my-type-1 requires my-label-1; only pod 9000 carries it,
    so the single copy always lands there, whichever pod the request enters through."""
    o = ob.OracleFleet(6553, 600_000, 6400)
    o.types_set({"my-type-1": {"required": ["my-label-1"]}})
    o.instance_event(ob.ADDED, 0, _row(), "inst-9000", labels=["my-label-1"], now_ms=1)
    o.instance_event(ob.ADDED, 1, _row(), "inst-9004", labels=[], now_ms=2)
    allowed, pref = o.type_sets("my-type-1", 2)
    assert allowed.tolist() == [True, False]
    dec = np.zeros(4, dtype=ob.DECISION)
    dec["type_idx"] = 0
    dec["self"] = [0, 1, 0, 1]
    dec["fresh_idx"] = -1
    dec["favour_self"] = [0, 0, 1, 1]
    dec["decision_id"] = np.arange(4)
    off = np.zeros(5, dtype=np.int64)
    res = o.get_next_batch(dec, ["my-type-1"], off, np.zeros(0, dtype=np.int32), 10, 1)
    # entering at 9000: itself (ABORT_REQUEST); entering at 9004: forwarded to 9000
    assert res["target"].tolist() == [ob.SELF, 0, ob.SELF, 0]
    # once loaded there it is excluded, nothing else qualifies -> null ("Nowhere available to load")
    off = np.asarray([0, 1, 2, 3, 4], dtype=np.int64)
    res = o.get_next_batch(dec, ["my-type-1"], off, np.zeros(4, dtype=np.int32), 10, 1)
    assert res["target"].tolist() == [ob.NONE] * 4


def test_documented_constraint_shapes(oracle_lib):
    """config/examples/type-constraints-example/README.md:21-39 — the only documented JSON shapes."""
    cfg = {"type-name1": {"required": ["label1", "label2"]}, "type-name2": {"preferred": ["label2"]},
           "type-name3": {"required": ["label1", "label3"], "preferred": ["label4"]},
           "_default": {"required": ["_unrecognized"]}}
    o = ob.OracleFleet(6553, 600_000, 6400)
    o.types_set(cfg)
    labels = [["label1", "label2"], ["label2"], ["label1", "label3", "label4"], ["label1", "label3"], []]
    for i, l in enumerate(labels):
        o.instance_event(ob.ADDED, i, _row(), f"pod-{i:04d}", labels=l, now_ms=i)
    o.tc_converge()
    a1, _ = o.type_sets("type-name1", 5)
    a3, _ = o.type_sets("type-name3", 5)
    ad, _ = o.type_sets("some-unknown-type", 5)
    a2, p2 = o.type_sets("type-name2", 5)
    assert a1.tolist() == [True, False, False, False, False]
    assert a3.tolist() == [False, False, True, True, False]
    assert ad.tolist() == [False] * 5  # _default: nowhere
    assert a2 is None  # no required labels: all instances are candidates


def _dummy_cluster(n_inst, now):
    """DummyModelMesh cluster (T/DummyModelMesh.java:39, T/ModelMeshEvictionsTest.java:558): n pods of 10 x 20 MiB, the
    unload reserve (2560 units, MM:749-755) published out of the capacity (MM:5373-5378)."""
    o = ob.OracleFleet(2560, 0, 2560)  # minChurnAgeMs 0: the churn guard is off in the test harness
    o.types_set(None)
    rows = np.zeros(n_inst, dtype=ob.INST)
    rows["capacity"] = 25600 - 2560
    rows["lru_time"] = (1 << 63) - 1
    rows["l_threads"] = 8
    rows["active"] = 1
    for i in range(n_inst):
        o.instance_event(ob.ADDED, i, rows[i], f"pod-{i}", now_ms=now)
    return o, rows


def test_multi_load_with_eviction_cluster_closed_loop(oracle_lib):
    """T/ModelMeshEvictionsTest.java:324-357: a 3-instance cluster holds int(0.9 * 30) = 27 models of
    20 MiB; loading 27 + 3 in sequence must keep every model but the oldest 3 + 2 x clusterSize ("wiggle room because
    there is intentionally some thresholds around instance selection", :338-340).  Driven through the closed loop
    (placement -> loadLocal -> LRU -> eviction -> republish), one request per republish window."""
    now = 1_760_000_000_000
    n_inst, size = 3, 2560
    o, rows = _dummy_cluster(n_inst, now)
    n_models = 30
    models = np.zeros(n_models, dtype=ob.SIM_MODEL)
    models["type_idx"], models["size_units"] = 0, size
    sim = ob.OracleSim(o, models, ["ExampleType"], np.zeros(n_models + 1, dtype=np.int64), np.zeros(0, dtype=np.int32),
                       np.zeros(n_models, dtype=np.int32), rows["capacity"], load_timeout_ms=30_000, last_published_ms=now - 60_000)
    evicted = []
    for m in range(n_models):
        e = np.zeros(1, dtype=ob.SIM_EVENT)
        e["type"], e["model"], e["caller"], e["u"], e["t"] = ob.SIM_REQUEST, m, m % n_inst, m, now + 2000 * m + 1
        dec, evi, _, _, _ = sim.step(e, now + 2000 * m, now + 2000 * (m + 1), seed=m)
        assert len(dec) == 1 and dec["status"][0] == ob.SIM_ACCEPTED, (m, dec)
        evicted += [int(x) for x in evi["model"]]
    loaded = [m for m in range(n_models) if len(sim.model_copies(m)[0]) > 0]
    # :347-354 idsForModelsWhichShouldBeLoaded = everything but the oldest 3 + wiggle room; the overflow itself need not be
    # spread evenly over the pods (which pod evicts follows the LRU-distance shortlist of the full case, MM:4913-4917)
    assert set(range(3 + 2 * n_inst, n_models)) <= set(loaded)
    assert 3 <= len(evicted) <= 3 + 2 * n_inst and all(m < 3 + 2 * n_inst for m in evicted)
    assert len(loaded) == n_models - len(evicted)
    assert all(sim.lru_state(i)[1] <= 9 * size for i in range(n_inst))


def test_multi_load_with_eviction_cluster_reuse(oracle_lib):
    """T/ModelMeshEvictionsTest.java:371-410: fill the cluster (27), use the first five again, load three more: the three
    new ones and the five reused ones must all be loaded afterwards (touch protects an entry from eviction)."""
    now = 1_760_000_000_000
    n_inst, size = 3, 2560
    o, rows = _dummy_cluster(n_inst, now)
    n_models = 30
    models = np.zeros(n_models, dtype=ob.SIM_MODEL)
    models["type_idx"], models["size_units"] = 0, size
    sim = ob.OracleSim(o, models, ["ExampleType"], np.zeros(n_models + 1, dtype=np.int64), np.zeros(0, dtype=np.int32),
                       np.zeros(n_models, dtype=np.int32), rows["capacity"], load_timeout_ms=30_000, last_published_ms=now - 60_000)
    step = 0

    def request(m):
        nonlocal step
        e = np.zeros(1, dtype=ob.SIM_EVENT)
        e["type"], e["model"], e["caller"], e["u"], e["t"] = ob.SIM_REQUEST, m, step % n_inst, step, now + 2000 * step + 1
        out = sim.step(e, now + 2000 * step, now + 2000 * (step + 1), seed=step)
        step += 1
        return out

    for m in range(27):
        request(m)
    # (the pick among the shortlist is a draw, MM:4981: a pod may already have evicted during the fill -- the wiggle room of
    # :338-340 -- so "reused" = the five oldest models that are loaded now)
    held = [m for m in range(27) if len(sim.model_copies(m)[0]) == 1]
    assert len(held) >= 27 - 2 * n_inst
    reused = held[:5]
    for m in reused:
        dec, evi, _, _, _ = request(m)   # a cache hit: runtimeCache.get(model, now)
        assert len(dec) == 0 and len(evi) == 0
    for m in range(27, 30):
        request(m)
    loaded = {m for m in range(n_models) if len(sim.model_copies(m)[0]) > 0}
    assert {27, 28, 29} <= loaded and set(reused) <= loaded   # :397-404


def test_second_copy_trigger_timing(oracle_lib):
    """T/ModelMeshEvictionsTest.java:412-446 testSecondCopyTrigger with its settings (:98-102: max age 10 s, min age 4 s, rate
    task every 100 ms => [40, 100] iterations): uses at 0.06 s, 1.06 s, 12.56 s, 17.56 s; only the last one -- another
    use 5 s earlier -- adds the second copy (pins MM:5726-5758)."""
    import ctypes as C
    i1, i2 = C.c_int32(-(1 << 31)), C.c_int32(-(1 << 31))  # CacheEntry initial values MM:1648
    now = 1_760_000_000_000
    fired = []
    for t_ms in (60, 1060, 12560, 17560):
        it = t_ms // 100 + 1  # the first run of the task after the use
        fired.append(oracle_lib.orc_second_copy_trigger(C.byref(i1), C.byref(i2), it, 40, 100, 50_000, 69_120, now - 3_600_000,
                                                        now + t_ms, max(10 * 3 * 10 * 1000, 6 * 3_600_000)))
    assert fired == [0, 0, 0, 1]
    # > 90 % full and a young cache: the trigger is suppressed (MM:5750-5752)
    i1, i2 = C.c_int32(120), C.c_int32(126)
    assert oracle_lib.orc_second_copy_trigger(C.byref(i1), C.byref(i2), 176, 40, 100, 1000, 69_120, now - 3_600_000, now, 6 * 3_600_000) == 0
    i1, i2 = C.c_int32(120), C.c_int32(126)
    assert oracle_lib.orc_second_copy_trigger(C.byref(i1), C.byref(i2), 176, 40, 100, 1000, 69_120, now - 7 * 3_600_000, now, 6 * 3_600_000) == 1


def test_scaleup_copies_arithmetic(oracle_lib):
    """MM:5718, 5760-5795 on hand-computed cases (parity unpinned by reference tests)."""
    import ctypes as C
    rpm = C.c_int32()
    f = oracle_lib.orc_scaleup_copies
    # 5000 requests in 10 s = 30 000 rpm, threshold 2000: min(15, candidates 10 - 1 - 0 = 9) = 9 -> capped at 10 // 3 = 3
    assert f(5000, 10_000, 2000, 1, 0, 10, 0, 0, 0, C.byref(rpm)) == 3 and rpm.value == 30_000
    assert f(500, 10_000, 2000, 1, 0, 10, 0, 0, 0, C.byref(rpm)) == 1 and rpm.value == 3000       # 1.5 x threshold: one copy
    assert f(300, 10_000, 2000, 1, 0, 10, 0, 0, 0, C.byref(rpm)) == 0                              # below the threshold
    assert f(5000, 10_000, 2000, 1, 0, 10, 0, 0, 1, None) == 0                                     # a copy was loaded too recently
    assert f(5000, 10_000, 2000, 4, 2, 6, 0, 0, 0, None) == 0                                      # nowhere left to load
    assert f(5000, 10_000, 2000, 2, 0, 10, 3, 2, 0, None) == 3                                     # 10 - 2 = 8, - 2 - 3 = 3: min(15, 3) = 3, cap 10 // 3 = 3


def test_java_golden_vectors(oracle_lib):
    """Fixtures produced by the REFERENCE ITSELF (oracle/java/GetNextHarness.java on a box with JDK 21 + the reference's jars):
    for every tests/golden/fleet_*.json that has a *.expected.json beside it, the oracle must reproduce the PLACEMENT_ORDER
    order, and per decision the ordered shortlist, instReqLoad and the result up to the random draw (N4).  The input files
    are committed (tools/golden_fleets.py); no expected file can be produced in this image (no JDK) -> "parity unpinned"."""
    import glob
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    inputs = sorted(p for p in glob.glob(os.path.join(gdir, "fleet_*.json")) if not p.endswith(".expected.json"))
    assert inputs, "tests/golden/fleet_*.json missing: run tools/golden_fleets.py"
    pairs = [(p, p.replace(".json", ".expected.json")) for p in inputs if os.path.exists(p.replace(".json", ".expected.json"))]
    for p in inputs:  # the inputs themselves must load into the oracle (format check of what the harness will be fed)
        doc = json.load(open(p))
        o = ob.OracleFleet(doc["minSpaceUnits"], doc["minChurnAgeMs"], 2560)
        o.types_set(doc["typeConstraints"])
        ids = [x["id"] for x in doc["instances"]]
        for i, x in enumerate(doc["instances"]):
            r = x["record"]
            row = np.zeros(1, dtype=ob.INST)
            row["lru_time"], row["count"], row["capacity"], row["used"] = r["lruTime"], r["count"], r["cap"], r["used"]
            row["l_threads"], row["l_in_prog"], row["rpm"], row["shutting_down"] = r["lThreads"], r["lInProg"], r["rpm"], int(r["shutdown"])
            row["start_time"], row["vers"], row["active"] = r["startTime"], r["vers"], int(x["active"])
            o.instance_event(ob.ADDED, i, row[0], x["id"], r["loc"], r["zone"], r["labels"] or (), now_ms=doc["now"])
        if doc["typeConstraints"] is not None:
            o.tc_converge()
        o.set_replaced_replicasets(doc["replaced"])
        assert len(o.cluster_order()) == sum(1 for x in doc["instances"] if not x["record"]["shutdown"])
        exp_path = p.replace(".json", ".expected.json")
        if not os.path.exists(exp_path):
            continue
        exp = json.load(open(exp_path))
        assert [ids[i] for i in o.cluster_order()] == exp["order"], p
        types = sorted({d["type"] for d in doc["decisions"]})
        for k, (d, e) in enumerate(zip(doc["decisions"], exp["decisions"])):
            od = np.zeros(1, dtype=ob.DECISION)
            od["type_idx"], od["self"], od["favour_self"], od["last_used"] = types.index(d["type"]), ids.index(d["self"]), int(d["favourSelf"]), d["lastUsed"]
            fresh = None
            od["fresh_idx"] = -1
            if d["fresh"] is not None:
                r = d["fresh"]
                fresh = np.zeros(1, dtype=ob.INST)
                fresh["lru_time"], fresh["count"], fresh["capacity"], fresh["used"], fresh["rpm"] = r["lruTime"], r["count"], r["cap"], r["used"], r["rpm"]
                od["fresh_idx"] = 0
            ex = np.asarray([ids.index(x) for x in d["excluded"]], dtype=np.int32)
            res, coff, cidx, cload, ckeep = o.get_next_batch(od, types, np.asarray([0, len(ex)], dtype=np.int64), ex, doc["now"], 1, fresh=fresh,
                                                            want_candidates=True)
            assert [ids[i] for i in cidx] == e["candidates"], (p, k)
            assert [int(x) for x in cload] == e["instReqLoad"], (p, k)
            if e["result"] in ("null", "SELF") and len(e["candidates"]) <= 1:
                assert int(res["target"][0]) == (ob.NONE if e["result"] == "null" else ob.SELF), (p, k)
    if not pairs:
        pytest.skip("parity unpinned: no tests/golden/*.expected.json (needs JDK 21 + the reference's jars, see oracle/java/README.md)")

"""CPU tests of the host logic (ordering, type-constraint masks, JSON reader, C ABI surface) and oracle properties.
"parity unpinned by reference tests": PLACEMENT_ORDER's tie-break chain and the TypeConstraintManager set algebra are
pinned by two independently structured implementations (oracle: literal Java shape; product: key sort + bitmasks)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from modelmesh_b200 import _lib as L
from modelmesh_b200.fleet import Fleet, MmpError
from modelmesh_b200.synth import SplitMix, make_fleet
from oracle import binding as ob

from helpers import oracle_from_synth, solver_from_synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_bound():
    """Every function include/mmplace.h declares is in the ctypes table (and vice versa)."""
    hdr = open(os.path.join(ROOT, "include", "mmplace.h")).read()
    declared = set(re.findall(r"\b(mmp_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in L.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)


def test_product_library_exports_every_symbol():
    """The built CUDA library loads without a GPU and exports the whole C ABI (no compute calls here)."""
    so = L.PRODUCT_SO
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(so)
    for name, _, _ in L.SYMBOLS:
        assert hasattr(lib, name), name
    assert lib.mmp_abi_version() == 2  # MMP_ABI_VERSION (include/mmplace.h): 2 = round 2 (checked loads, closed loop, commit info)


def test_product_has_no_cpu_fallback():
    """Without a CUDA device mmp_fleet_create must fail loudly with MMP_E_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = L.load_product()
    with pytest.raises(MmpError) as e:
        Fleet(100, 1000, 10, 16, 16, lib=lib)
    assert e.value.code == L.E_CUDA


@pytest.mark.parametrize("config,ni,seed", [("C2", 300, 2), ("C3", 500, 3), ("C5", 400, 5), ("MIX", 200, 7), ("MIX", 97, 11)])
def test_placement_order_is_a_strict_weak_order_and_matches(emul_lib, oracle_lib, config, ni, seed):
    fl = make_fleet(config, 10, ni, seed)
    o = oracle_from_synth(fl)
    s = solver_from_synth(fl, emul_lib)
    order = o.cluster_order()
    assert np.array_equal(order, s.cluster_order())
    rng = SplitMix(seed)
    live = order
    idx = rng.randint(3000, 0, len(live)).reshape(-1, 3)
    pos = {int(v): k for k, v in enumerate(order)}
    for a, b, c in idx:
        a, b, c = int(live[a]), int(live[b]), int(live[c])
        ab, ba = o.compare(a, b), o.compare(b, a)
        assert np.sign(ab) == -np.sign(ba)
        if a != b:
            assert ab != 0 and (ab < 0) == (pos[a] < pos[b])
        if o.compare(a, b) < 0 and o.compare(b, c) < 0:
            assert o.compare(a, c) < 0


def test_tie_breaks_down_to_strings(emul_lib, oracle_lib):
    """Identical numeric columns: order falls to instanceId, then location / zone (nulls last), then labels
    (MM:4697-4700, Utils.java:25-36); ids compare as UTF-16 code units, not UTF-8 bytes."""
    base = np.zeros(1, dtype=L.INSTANCE_ROW)
    base["capacity"], base["used"], base["lru_time"], base["l_threads"], base["active"] = 1000, 100, 5000, 4, 1
    ids = ["b", "a", "a\U00010000", "a￿", "a", "aa", "B", "", "é", "z"]
    o = ob.OracleFleet(10, 600_000, 10)
    s = Fleet(10, 600_000, 10, 32, 4, lib=emul_lib)
    for i, iid in enumerate(ids):
        o.instance_event(ob.ADDED, i, base[0], iid, now_ms=1)
        s.instance_upsert(i, base[0], iid)
    s.commit()
    want = [ids[i] for i in o.cluster_order()]
    assert [ids[i] for i in s.cluster_order()] == want
    # UTF-16: the surrogate pair for U+10000 (D800 DC00) sorts before U+E000 and U+FFFF
    assert want.index("a\U00010000") < want.index("a") < want.index("a￿")


def test_type_masks_match_oracle_sets(emul_lib, oracle_lib):
    for config, ni, seed in [("C3", 400, 3), ("C5", 300, 5), ("MIX", 160, 2), ("MIX", 97, 6), ("MIX", 64, 10)]:
        fl = make_fleet(config, 10, ni, seed)
        if fl.type_config is None:
            continue
        o = oracle_from_synth(fl)
        s = solver_from_synth(fl, emul_lib)
        live = fl.inst_rows["shutting_down"] == 0
        active = fl.inst_rows["active"] != 0
        for t in fl.type_names + ["never-configured-type"]:
            a, p = o.type_sets(t, ni)
            a2, p2 = s.type_sets(s.type_id(t) if t in fl.type_names else 0, ni)
            assert (a is None) == (a2 is None), t
            assert (p is None) == (p2 is None), t
            if a is not None:
                assert np.array_equal(a & live & active, a2), t  # product folds the siMap bit into the candidate mask
            if p is not None:
                assert np.array_equal(p & live, p2), t


def test_deferred_refresh_reaches_the_same_state(oracle_lib):
    fl = make_fleet("C5", 10, 250, 5)
    outs = []
    for defer in (False, True):
        o = ob.OracleFleet(fl.min_space_units, fl.min_churn_age_ms, fl.default_model_size_units)
        o.types_set(fl.type_config)
        o.tc_defer_refresh(defer)
        for i in range(fl.n_instances):
            o.instance_event(ob.ADDED, i, fl.inst_rows[i], fl.inst_ids[i], fl.inst_locs[i], fl.inst_zones[i], fl.inst_labels[i], 1)
        o.tc_converge()
        outs.append([o.type_sets(t, fl.n_instances) for t in fl.type_names])
    for (a1, p1), (a2, p2) in zip(*outs):
        assert (a1 is None) == (a2 is None) and (p1 is None) == (p2 is None)
        assert a1 is None or np.array_equal(a1, a2)
        assert p1 is None or np.array_equal(p1, p2)


def test_json_reader_matches_python_json(emul_lib, oracle_lib):
    doc = r'''
    { "té-1" : {"required": ["a", "b", "a"], "preferred": ["c", "b"], "ignored": {"x": [1, 2, {"y": null}]}},
      "t2": {"preferred": ["😀", "a"]}, "t3": {}, "t4": {"required": null},
      "_default": {"required": ["zz"]} }'''
    cfg = json.loads(doc)
    cfg["t4"] = {}
    base = np.zeros(1, dtype=L.INSTANCE_ROW)
    base["capacity"], base["lru_time"], base["active"] = 1000, 5000, 1
    labels = [["a", "b"], ["a"], ["c"], ["\U0001F600"], [], ["zz"], ["b", "c", "a"]]
    o = ob.OracleFleet(10, 600_000, 10)
    o.types_set(cfg)
    s = Fleet(10, 600_000, 10, 16, 4, lib=emul_lib)
    s.types_set_json(doc)
    for i, l in enumerate(labels):
        o.instance_event(ob.ADDED, i, base[0], f"pod-{i:04d}", labels=l, now_ms=1)
        s.instance_upsert(i, base[0], f"pod-{i:04d}", labels=l)
    o.tc_converge()
    s.commit()
    for t in list(cfg.keys()) + ["unknown"]:
        a, p = o.type_sets(t, len(labels))
        a2, p2 = s.type_sets(s.type_id(t), len(labels))
        assert (a is None) == (a2 is None) and (p is None) == (p2 is None), t
        assert a is None or np.array_equal(a, a2), t
        assert p is None or np.array_equal(p, p2), t
    for bad in ['{"t": {"required": ["a"}', '[1]', '{"t": 3}', '{"t": {"required": "a"}}', '{"t": {}} x']:
        with pytest.raises(MmpError):
            s.types_set_json(bad)


def test_ingest_validation(emul_lib):
    s = Fleet(10, 600_000, 10, 8, 4, lib=emul_lib)
    row = np.zeros(1, dtype=L.INSTANCE_ROW)
    row["lru_time"] = -5
    with pytest.raises(MmpError):
        s.instance_upsert(0, row[0], "x")
    row["lru_time"], row["rpm"] = 1, 600_000_000
    with pytest.raises(MmpError):
        s.instance_upsert(0, row[0], "x")
    row["rpm"] = 1
    with pytest.raises(MmpError):
        s.instance_upsert(99, row[0], "x")
    s.instance_upsert(0, row[0], "x")
    dec = np.zeros(1, dtype=L.DECISION_IN)
    with pytest.raises(MmpError) as e:
        s.place_batch(dec, 1, 1)
    assert e.value.code == L.E_EPOCH
    m = np.zeros(1, dtype=L.MODEL_ROW)
    m["type_id"] = 77
    with pytest.raises(MmpError):
        s.model_upsert(0, m[0], [0])


def test_invalid_decisions_are_flagged(emul_lib):
    fl = make_fleet("C2", 50, 40, 2)
    s = solver_from_synth(fl, emul_lib)
    dec = np.zeros(3, dtype=L.DECISION_IN)
    dec["model"] = [0, 9999, 0]
    dec["self"] = [0, 0, -4]
    dec["fresh"] = -1
    out = s.place_batch(dec, fl.now_ms, 1)
    assert out["target"][1] == L.TARGET_INVALID and out["target"][2] == L.TARGET_INVALID and out["target"][0] != L.TARGET_INVALID


def test_upgrade_tracker_rolling_update(oracle_lib):
    """UT:120-187: a new replicaset with the same (empty) labels appears; the old one, whose pods are being removed,
    becomes likely-replaced; entries expire 15 min after the last change (UT:192-201).  N9: only label-less records
    share a PerTypeLabelStats."""
    o = ob.OracleFleet(10, 600_000, 10)
    base = np.zeros(1, dtype=ob.INST)
    base["capacity"], base["lru_time"], base["active"] = 1000, 5000, 1
    t0 = 1_000_000_000
    def ev(kind, idx, iid, start, now, labels=()):
        r = base[0].copy()
        r["start_time"] = start
        r["l_in_prog"] = 1 if kind == ob.UPDATED else 0  # an identical record is ignored (MM:1497-1502)
        o.instance_event(kind, idx, r, iid, labels=labels, now_ms=now)
    # NB MM:1552 only feeds the tracker on UPDATE-of-existing (existingWasRemoved); so add then update each pod
    for k in range(3):
        ev(ob.ADDED, k, f"oldrs1-{k}", t0 + k, t0 + k)
        ev(ob.UPDATED, k, f"oldrs1-{k}", t0 + k, t0 + 10 + k)
    assert o.get_replaced_replicasets() == []
    t1 = t0 + 3_600_000
    ev(ob.ADDED, 10, "newrs2-0", t1, t1)
    ev(ob.UPDATED, 10, "newrs2-0", t1, t1 + 5)
    assert o.get_replaced_replicasets() == ["oldrs1"]
    # labelled pods never share a tracker entry (identity-hashed String[] key)
    o2 = ob.OracleFleet(10, 600_000, 10)
    o = o2
    for k in range(3):
        ev(ob.ADDED, k, f"oldrs1-{k}", t0 + k, t0 + k, labels=["x"])
        ev(ob.UPDATED, k, f"oldrs1-{k}", t0 + k, t0 + 10 + k, labels=["x"])
    ev(ob.ADDED, 10, "newrs2-0", t1, t1, labels=["x"])
    ev(ob.UPDATED, 10, "newrs2-0", t1, t1 + 5, labels=["x"])
    assert o.get_replaced_replicasets() == []


def test_admission_rules(oracle_lib):
    """MM:3872-3884 churn guard and MM:5185-5190 early reject."""
    L_ = oracle_lib
    assert L_.orc_churn_reject(1000, 990, 5_000, 50, 600_000, 100_000) == 1   # full and LRU entry only 95 s old
    assert L_.orc_churn_reject(1000, 990, 5_000, 50, 600_000, 700_000) == 0   # old enough
    assert L_.orc_churn_reject(1000, 100, 5_000, 50, 600_000, 100_000) == 0   # not full
    assert L_.orc_churn_reject(1000, 990, -1, 50, 600_000, 100_000) == 0      # empty cache
    assert L_.orc_early_reject(2000, 1000, 0, -1, 0) == 1                      # bigger than the whole cache
    assert L_.orc_early_reject(200, 1000, 900, 5000, 4000) == 1                # no room and older than everything
    assert L_.orc_early_reject(200, 1000, 900, 5000, 6000) == 0
    assert L_.orc_early_reject(200, 1000, 900, 5000, 0) == 0                   # lastUsed 0 = "now"


@pytest.mark.parametrize("config,ni,seed", [("C3", 600, 3), ("C5", 500, 5), ("MIX", 300, 2), ("MIX", 160, 9), ("C2", 400, 2)])
def test_oracle_bulk_load_equals_event_driven(oracle_lib, config, ni, seed):
    """orc_bulk_add is a set-up shortcut for big fleets; it must reach exactly the event-driven state."""
    fl = make_fleet(config, 10, ni, seed)
    a = oracle_from_synth(fl, bulk=False)
    b = oracle_from_synth(fl, bulk=True)
    assert np.array_equal(a.cluster_order(), b.cluster_order())
    for k in ("total_capacity", "total_free", "global_lru", "instance_count", "model_copy_count"):
        assert int(a.cluster_stats()[k]) == int(b.cluster_stats()[k])
    for t in fl.type_names + ["zzz"]:
        (a1, p1), (a2, p2) = a.type_sets(t, ni), b.type_sets(t, ni)
        assert (a1 is None) == (a2 is None) and (p1 is None) == (p2 is None)
        assert a1 is None or np.array_equal(a1, a2)
        assert p1 is None or np.array_equal(p1, p2)
    if fl.type_config is not None:
        sa, ia = a.partition_stats()
        sb, ib = b.partition_stats()
        key = lambda s_: sorted((int(x["total_capacity"]), int(x["total_free"]), int(x["instance_count"]), int(x["model_copy_count"])) for x in s_ if x["instance_count"] > 0)
        assert key(sa) == key(sb)

"""BASELINE.json configs[0] (C1): 1k models x 16 instances, uniform sizes, sequential ensureLoaded of every model with the
state updated after each step (placement -> local LRU insert -> evictions -> republish).  The same closed loop is driven
through the oracle (OracleFleet + one OracleLru per instance) and through the C ABI (mmp_place_one + mmp_lru_apply +
mmp_instance_update + mmp_model_upsert + mmp_fleet_commit); every placement, eviction and published record must agree."""
import numpy as np
import pytest

from modelmesh_b200 import _lib as L
from modelmesh_b200.fleet import Fleet
from modelmesh_b200.synth import LONG_MAX, NOW_MS, make_fleet
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def test_c1_sequential_ensure_loaded_closed_loop(product_lib, oracle_lib):
    n_models, n_inst = 1000, 16
    fl = make_fleet("C1", n_models, n_inst, 1)
    size, cap_pub = 2560, 25600 - 2560  # 20 MiB models; 10 x 20 MiB minus the unload reserve -> 9 models per instance
    now = NOW_MS

    o = ob.OracleFleet(fl.min_space_units, fl.min_churn_age_ms, fl.default_model_size_units)
    o.types_set(None)
    s = Fleet(fl.min_space_units, fl.min_churn_age_ms, fl.default_model_size_units, n_inst, n_models, lib=product_lib)
    s.types_set_json(None)
    rows = fl.inst_rows.copy()
    for i in range(n_inst):
        o.instance_event(ob.ADDED, i, rows[i], fl.inst_ids[i], now_ms=now)
        s.instance_upsert(i, rows[i], fl.inst_ids[i])
    mrow = np.zeros(1, dtype=L.MODEL_ROW)
    tid = s.type_id("ExampleType")
    for m in range(n_models):
        mrow["last_used"], mrow["size_units"], mrow["type_id"] = fl.model_last_used[m], size, tid
        s.model_upsert(m, mrow[0], [])
    s.commit()
    s.lru_init(np.full(n_inst, cap_pub, dtype=np.int64), 64)
    olru = [ob.OracleLru(cap_pub) for _ in range(n_inst)]
    loaded = [set() for _ in range(n_models)]     # model -> instances (oracle side == product side by construction)
    placements, evictions_total = [], 0

    for m in range(n_models):
        self_idx = m % n_inst
        t = now + m  # one request per millisecond
        last_used = int(fl.model_last_used[m])
        # ---- placement: oracle ----
        od = np.zeros(1, dtype=ob.DECISION)
        od["type_idx"], od["self"], od["fresh_idx"], od["favour_self"], od["last_used"], od["decision_id"] = 0, self_idx, -1, 1, last_used, 0
        excl = np.asarray(sorted(loaded[m]), dtype=np.int32)
        ores = o.get_next_batch(od, ["ExampleType"], np.asarray([0, len(excl)], dtype=np.int64), excl, t, seed=m)
        # ---- placement: product ----
        d = np.zeros(1, dtype=L.DECISION_IN)
        d["model"], d["self"], d["last_used"], d["flags"], d["fresh"] = m, self_idx, last_used, L.DF_FAVOUR_SELF, -1
        res = s.place_one(d, t, seed=m)
        assert int(res["target"]) == int(ores["target"][0]), (m, res, ores)
        tgt = int(res["target"])
        tgt = self_idx if tgt == L.TARGET_SELF else tgt
        assert tgt >= 0
        placements.append(tgt)
        # ---- loadLocal on the target: 1-unit placeholder then the predicted size (MM:5061, 2094-2100) ----
        ev = np.zeros(2, dtype=L.LRU_EVENT)
        ev["op"] = [0, 2]
        ev["instance"], ev["model"] = tgt, m
        ev["weight"] = [1, size]
        ev["last_used"] = [last_used, 0]
        got = s.lru_apply(ev, t)
        oe = np.zeros(2, dtype=ob.LRU_EVENT)
        oe["op"], oe["key"], oe["weight"], oe["last_used"] = [0, 2], m, [1, size], [last_used, 0]
        want = olru[tgt].apply(oe, t)
        assert [(int(x["model"]), int(x["last_used"])) for x in got] == [(int(x["key"]), int(x["last_used"])) for x in want], m
        evictions_total += len(got)
        changed_models = {m}
        loaded[m].add(tgt)
        for x in got:
            loaded[int(x["model"])].discard(tgt)
            changed_models.add(int(x["model"]))
        if m in [int(x["model"]) for x in got]:
            loaded[m].discard(tgt)  # evicted immediately: older than everything else and no room (MM:5145-5148)
        # ---- republish the target's instance record (getFreshInstanceRecord MM:5369-5386) ----
        keys, ts, ws = olru[tgt].dump()
        rows[tgt]["used"] = int(ws.sum())
        rows[tgt]["count"] = len(keys)
        rows[tgt]["lru_time"] = int(ts[0]) if len(ts) else LONG_MAX
        o.instance_event(ob.UPDATED, tgt, rows[tgt], fl.inst_ids[tgt], now_ms=t)
        s.instance_update(tgt, rows[tgt])
        for cm in changed_models:
            mrow["last_used"], mrow["size_units"], mrow["type_id"] = fl.model_last_used[cm], size, tid
            s.model_upsert(cm, mrow[0], sorted(loaded[cm]))
        s.commit()
        oldest, weighted, count = s.lru_state()
        assert int(weighted[tgt]) == olru[tgt].weighted_size() and int(count[tgt]) == olru[tgt].size()
    # 16 instances x 9 slots = 144 resident copies; everything else was evicted, oldest (= highest index, lastUsed = now - i s) first
    assert sum(len(x) for x in loaded) == 144
    assert evictions_total == n_models - 144
    assert np.array_equal(s.cluster_order(), o.cluster_order())
    assert len(set(placements[:16])) > 1  # the first sixteen requests spread over the fleet

"""The threading contract of include/mmplace.h, exercised: many request threads place concurrently (mmp_place_one and the
micro-batcher's mmp_place_submit, the shape of litelinks' request pool calling getNext, MM:918-925, 1107-1110) while one
writer thread keeps ingesting instance updates and committing new epochs.  Every result must be the oracle's answer under ONE
of the epochs that were live -- never a mixture."""
import ctypes as C
import threading

import numpy as np
import pytest

from helpers import oracle_from_synth, oracle_inputs
from modelmesh_b200 import _lib as L
from modelmesh_b200.fleet import Fleet
from modelmesh_b200.synth import load_into_fleet, make_decisions, make_fleet
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def _oracle_answers(fl, rows, sd, seed, ids):
    """Per decision the oracle's target when decision i is drawn with (seed, ids[i])."""
    import copy
    fl2 = copy.copy(fl)
    fl2.inst_rows = rows
    o = oracle_from_synth(fl2)
    od, off, idx = oracle_inputs(fl2, sd)
    od["decision_id"] = ids
    res = o.get_next_batch(od, fl.type_names, off, idx, fl.now_ms, seed, fresh=sd.fresh if len(sd.fresh) else None)
    return res["target"].copy(), res["n_candidates"].copy()


@pytest.mark.parametrize("mode", ["place_one", "place_one_server", "submit"])
def test_concurrent_placement_with_ingest_and_commit(product_lib, oracle_lib, mode):
    lib = product_lib
    fl = make_fleet("C3", 3000, 700, 3)
    n_dec, n_threads, seed = 1536, 16, 77
    sd = make_decisions(fl, n_dec, 9)
    # two fleet states A / B: B = A with 150 instances' numeric columns changed (a non-structural, device-path commit)
    rows_a = fl.inst_rows.copy()
    rows_b = fl.inst_rows.copy()
    rng = np.random.default_rng(1)
    live = np.nonzero(rows_a["shutting_down"] == 0)[0]
    changed = rng.choice(live, size=150, replace=False)
    for i in changed:
        rows_b[i]["used"] = int(rng.integers(0, rows_b[i]["capacity"] + 1))
        rows_b[i]["count"] = int(rng.integers(0, 300))
        rows_b[i]["lru_time"] = int(fl.now_ms - rng.integers(1, 5_000_000))
    s = Fleet(fl.min_space_units, fl.min_churn_age_ms, fl.default_model_size_units, fl.n_instances, fl.n_models, lib=lib)
    load_into_fleet(fl, s)
    if mode == "place_one_server":  # the resident server answers one caller at a time, the others take the graph path; a short
        s._ck(lib.mmp_tune(s.h, b"one_mode", 3))  # lifetime makes restarts (and restarts across commits) frequent
        s._ck(lib.mmp_tune(s.h, b"server_life_us", 200))
    batcher = C.c_void_p()
    if mode == "submit":
        s._ck(lib.mmp_batcher_create(s.h, 256, 50, seed, C.byref(batcher)))
    results = np.zeros(n_dec, dtype=L.DECISION_OUT)
    dec_ids = np.zeros(n_dec, dtype=np.uint64)
    errors = []
    stop = threading.Event()

    def writer():
        k = 0
        while not stop.is_set():
            rows = rows_b if k % 2 == 0 else rows_a
            try:
                for i in changed:
                    s.instance_update(int(i), rows[i])
                s.commit()
            except Exception as e:  # pragma: no cover
                errors.append(e)
                return
            k += 1

    def reader(t):
        fresh = np.ascontiguousarray(sd.fresh, dtype=L.INSTANCE_ROW)
        extra = np.ascontiguousarray(sd.extra, dtype=np.int32)
        out = np.zeros(1, dtype=L.DECISION_OUT)
        did = C.c_uint32()
        try:
            for rep in range(3):
                for i in range(t, n_dec, n_threads):
                    d = sd.dec[i:i + 1]
                    fp = fresh.ctypes.data_as(C.c_void_p) if len(fresh) else None
                    ep = extra.ctypes.data_as(C.c_void_p) if len(extra) else None
                    if mode == "submit":
                        rc = lib.mmp_place_submit(batcher, d.ctypes.data_as(C.c_void_p), fp, ep, fl.now_ms, out.ctypes.data_as(C.c_void_p), C.byref(did))
                        dec_ids[i] = did.value
                    else:
                        rc = lib.mmp_place_one(s.h, d.ctypes.data_as(C.c_void_p), fp, ep, out.ctypes.data_as(C.c_void_p), fl.now_ms, seed)
                        dec_ids[i] = 0  # a batch of one: position 0
                    if rc < 0:
                        raise RuntimeError(lib.mmp_last_error(s.h))
                    results[i] = out[0]
        except Exception as e:  # pragma: no cover
            errors.append(e)

    wt = threading.Thread(target=writer)
    wt.start()
    readers = [threading.Thread(target=reader, args=(t,)) for t in range(n_threads)]
    for r in readers:
        r.start()
    for r in readers:
        r.join()
    stop.set()
    wt.join()
    assert not errors, errors[:2]
    if mode == "submit":
        nb, nd = C.c_int64(), C.c_int64()
        s._ck(lib.mmp_batcher_stats(batcher, C.byref(nb), C.byref(nd)))
        assert nd.value == 3 * n_dec and nb.value < nd.value  # callers were coalesced into shared launches
        lib.mmp_batcher_destroy(batcher)
    ta, ca = _oracle_answers(fl, rows_a, sd, seed, dec_ids)
    tb, cb = _oracle_answers(fl, rows_b, sd, seed, dec_ids)
    ok_a = (results["target"] == ta) & (results["n_candidates"] == ca)
    ok_b = (results["target"] == tb) & (results["n_candidates"] == cb)
    bad = np.nonzero(~(ok_a | ok_b))[0]
    assert len(bad) == 0, (len(bad), bad[:5], results[bad[:5]], ta[bad[:5]], tb[bad[:5]])
    assert np.count_nonzero(ta != tb) > 20  # the two epochs really differ

#!/usr/bin/env python
"""Writes tests/golden/fleet_*.json: small deterministic fleets + decisions in the input format of
oracle/java/GetNextHarness.java (the reference-side fixture generator, runnable only on a box with JDK 21 and the reference
built).  The matching *.expected.json files -- outputs of the reference's own getNext -- are consumed by
tests/test_oracle_golden.py::test_java_golden_vectors when present."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from modelmesh_b200 import _lib as L  # noqa: E402
from modelmesh_b200.synth import make_decisions, make_fleet  # noqa: E402


def record(row, loc, zone, labels):
    return {"lruTime": int(row["lru_time"]), "count": int(row["count"]), "cap": int(row["capacity"]), "used": int(row["used"]),
            "lThreads": int(row["l_threads"]), "lInProg": int(row["l_in_prog"]), "rpm": int(row["rpm"]),
            "shutdown": bool(row["shutting_down"]), "startTime": int(row["start_time"]), "vers": int(row["vers"]), "loc": loc,
            "zone": zone, "labels": list(labels) or None}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for config, nm, ni, seed in (("C2", 60, 24, 2), ("C3", 80, 40, 3), ("C5", 80, 40, 5), ("MIX", 60, 33, 14), ("MIX", 60, 48, 41)):
        fl = make_fleet(config, nm, ni, seed)
        sd = make_decisions(fl, 40, seed)
        doc = {"minSpaceUnits": fl.min_space_units, "minChurnAgeMs": fl.min_churn_age_ms, "now": fl.now_ms,
               "typeConstraints": fl.type_config, "replaced": list(fl.replaced_replicasets),
               "instances": [{"id": fl.inst_ids[i], "active": bool(fl.inst_rows[i]["active"]),
                              "record": record(fl.inst_rows[i], fl.inst_locs[i], fl.inst_zones[i], fl.inst_labels[i])} for i in range(ni)],
               "decisions": []}
        for d in sd.dec:
            m = int(d["model"])
            ex = [int(x) for x in fl.edge_inst[fl.edge_off[m]:fl.edge_off[m + 1]]] + [int(x) for x in sd.extra[d["extra_off"]:d["extra_off"] + d["extra_n"]]]
            lu = int(fl.model_last_used[m]) if d["flags"] & L.DF_MODEL_LAST_USED else int(d["last_used"])
            fresh = None
            if d["fresh"] >= 0:
                s = int(d["self"])
                fresh = record(sd.fresh[int(d["fresh"])], fl.inst_locs[s], fl.inst_zones[s], fl.inst_labels[s])
            doc["decisions"].append({"type": fl.type_names[int(fl.model_type[m])], "self": fl.inst_ids[int(d["self"])], "fresh": fresh,
                                     "favourSelf": bool(d["flags"] & L.DF_FAVOUR_SELF), "lastUsed": lu,
                                     "excluded": [fl.inst_ids[i] for i in ex]})
        with open(os.path.join(out_dir, f"fleet_{config.lower()}_{seed}.json"), "w") as f:
            json.dump(doc, f, indent=0, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()

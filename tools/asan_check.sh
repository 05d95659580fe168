#!/bin/bash
# Host code + decision routines under AddressSanitizer / UBSan (CPU harness build): JSON readers on garbage, every
# cooperative shape, the lane routine on exact-size window copies, instance-shard ranges.  No GPU involved.
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -Wall -Wl,-Bsymbolic -shared \
    -o /tmp/libmmplace_emul_asan.so tests/emul/emul.cpp
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python - <<'PY'
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from modelmesh_b200 import _lib
from modelmesh_b200.fleet import Fleet
from modelmesh_b200.synth import make_decisions, make_fleet, load_into_fleet
from helpers import solver_from_synth
import test_record_codec as t
lib = _lib.load("/tmp/libmmplace_emul_asan.so", require_all=False)
t.test_json_readers_survive_garbage(lib)
t.test_record_defaults_and_errors(lib)
for cfg, nm, ni in (("MIX", 400, 300), ("C5", 800, 700), ("C3", 800, 1300)):
    fl = make_fleet(cfg, nm, ni, 14)
    s = solver_from_synth(fl, lib)
    for shape in (32, 16, 8, 1):
        lib.mmp_emul_set_window(shape)
        sd = make_decisions(fl, 800, 3)
        s.place_batch(sd.dec, fl.now_ms, 5, fresh=sd.fresh, extra=sd.extra)
    for world, rank in ((1, 0), (3, 0), (3, 1), (3, 2), (8, 5)):
        f = Fleet(fl.min_space_units, fl.min_churn_age_ms, fl.default_model_size_units, fl.n_instances, fl.n_models,
                  shard_rank=rank, shard_count=world, lib=lib)
        load_into_fleet(fl, f)
        for win in (1, 2, 5, 14):
            lib.mmp_emul_set_window(2); lib.mmp_emul_set_lane_window(win)
            sd = make_decisions(fl, 800, 3)
            f.place_batch(sd.dec, fl.now_ms, 5, fresh=sd.fresh, extra=sd.extra)
        f.close()
# long rows: the chunk machinery of decide_stream (8-step gathers, chunks taken at once), a full best, non-simple (b)
for cfg, nm, ni in (("C5", 600, 10000), ("C5", 600, 5000), ("C3", 600, 10000)):
    fl = make_fleet(cfg, nm, ni, 5)
    s = solver_from_synth(fl, lib)
    for win, budget in ((12, 192), (0, 1000), (3, 400), (12, 7)):
        lib.mmp_emul_set_window(2); lib.mmp_emul_set_lane_window(win); lib.mmp_emul_set_lane_budget(budget)
        for plain in (True, False):
            sd = make_decisions(fl, 600, 5, sweep=plain, plain=plain)
            s.place_batch(sd.dec, fl.now_ms, 5, fresh=sd.fresh if len(sd.fresh) else None, extra=sd.extra if len(sd.extra) else None)
    s.close()
print("asan/ubsan: clean")
PY

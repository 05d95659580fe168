"""GPU parity tests proper: the CUDA library (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

from modelmesh_b200.synth import make_decisions, make_fleet

from helpers import compare_decisions, oracle_from_synth, solver_from_synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,nm,ni,seed", [
    ("C1", 1000, 16, 1), ("C2", 5000, 1000, 2), ("C2", 3000, 2000, 12), ("C3", 5000, 3000, 3), ("C5", 5000, 4000, 5),
    ("C3", 3000, 10000, 33), ("C5", 2000, 10000, 55), ("C3", 1000, 7000, 7), ("C5", 1000, 12345, 9),
    ("C3", 500, 20000, 11), ("C5", 300, 40000, 13), ("C3", 300, 65536, 17),
])
def test_decisions_match_oracle(product_lib, oracle_lib, config, nm, ni, seed):
    fl = make_fleet(config, nm, ni, seed)
    o = oracle_from_synth(fl)
    s = solver_from_synth(fl, product_lib)
    assert np.array_equal(s.cluster_order(), o.cluster_order())
    nd = 3000 if ni <= 4000 else 1200
    sd = make_decisions(fl, nd, seed)
    compare_decisions(fl, sd, o, s, seed=seed * 7919)
    sd = make_decisions(fl, 1000, seed + 1, sweep=True, plain=True)
    compare_decisions(fl, sd, o, s, seed=seed)


@pytest.mark.parametrize("seed", range(60))
def test_mixed_regimes_match_oracle(product_lib, oracle_lib, seed):
    ni = [33, 64, 97, 160, 300, 1000, 1025, 2100, 4200][seed % 9]
    fl = make_fleet("MIX", 600, ni, seed)
    o = oracle_from_synth(fl)
    s = solver_from_synth(fl, product_lib)
    assert np.array_equal(s.cluster_order(), o.cluster_order())
    sd = make_decisions(fl, 1500, seed)
    compare_decisions(fl, sd, o, s, seed=seed + 99)


@pytest.mark.parametrize("kernel,tile,ring_k", [("lanes", 16, 4), ("tile", 8, 4), ("tile", 16, 2), ("tile", 16, 4), ("tile", 32, 4)])
@pytest.mark.parametrize("config,nm,ni,seed", [("C3", 4000, 10000, 3), ("C5", 3000, 5000, 5), ("MIX", 600, 300, 14), ("C2", 3000, 1000, 2),
                                               ("C3", 2000, 16000, 7)])
def test_kernel_variants_match_oracle(product_lib, oracle_lib, monkeypatch, kernel, tile, ring_k, config, nm, ni, seed):
    """k_place_lanes (one decision per lane, the default) and every tile width / ring depth of the cooperative k_place
    give the same, oracle-identical answers."""
    monkeypatch.setenv("MMP_KERNEL", kernel)
    monkeypatch.setenv("MMP_TILE", str(tile))
    monkeypatch.setenv("MMP_RING_K", str(ring_k))
    fl = make_fleet(config, nm, ni, seed)
    o = oracle_from_synth(fl)
    s = solver_from_synth(fl, product_lib)
    sd = make_decisions(fl, 2500, seed)
    compare_decisions(fl, sd, o, s, seed=seed * 17, full_lists=False)
    sd = make_decisions(fl, 2500, seed + 1, sweep=True, plain=True)
    compare_decisions(fl, sd, o, s, seed=seed, full_lists=False)

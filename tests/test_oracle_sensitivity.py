"""What the reference's algorithm itself does when only its summation order changes (CPU, oracle only).

VERDICT r1 asked whether the GPU-vs-oracle differences of the chained tests (solve -> marginalize -> next window) are
the eps = 1e-30 pseudo-inverse's noise amplification (marginalization_factor.h:129, .cpp:240-263) or a GPU bug.  Here
the ORACLE runs the same chains against ITSELF with the factors handed over in shuffled order: a mathematically
identical problem whose J'J / J'r sums round differently in the last bit.  The measured drift is what any second
implementation of the reference's algorithm - including the reference on another compiler - can be expected to
show; the GPU parity tests use 4x this envelope (or the north-star tolerance, whichever is larger)."""
import numpy as np

from helpers import chain_difference, order_sensitivity, run_c3_sequence, run_c5


def test_oracle_c3_chain_sensitivity_to_summation_order(oracle_lib):
    base = run_c3_sequence(oracle_lib)
    env = order_sensitivity([base] + [run_c3_sequence(oracle_lib, perm_seed=s) for s in (1, 2, 3)])
    print("C3 oracle-vs-oracle (shuffled factor order):", env)
    # window A (no prior) is insensitive; the chain through the pseudo-inverse is not bit-stable but stays far below
    # anything that would hide a real defect
    for s in (1, 2):
        r = run_c3_sequence(oracle_lib, perm_seed=s)
        assert np.isclose(r["costs"][0], base["costs"][0], rtol=1e-11)
        assert r["iterations"] == base["iterations"]
    assert env["trans_rel"] < 1e-4 and env["rot_rad"] < 1e-4 and env["cost_rel"] < 5e-3


def test_oracle_c5_chain_sensitivity_to_summation_order(oracle_lib):
    n = 5
    base = run_c5(oracle_lib, n)
    env = order_sensitivity([base] + [run_c5(oracle_lib, n, perm_seed=s) for s in (1, 2, 3)])
    print("C5 oracle-vs-oracle (shuffled factor order):", env)
    assert env["trans_rel"] < 1e-4 and env["rot_rad"] < 1e-4 and env["cost_rel"] < 5e-3
    # the MARGIN_SECOND_NEW branch keeps the prior untouched and drops the second-newest frame
    r = run_c5(oracle_lib, n, second_new_every=3)
    assert r["marg_flags"].count(1) >= 1 and r["marg_flags"].count(0) >= 1
    assert all(p == r["prior_dims"][0] for p in r["prior_dims"])

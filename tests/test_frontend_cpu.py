"""Front-end formats, CPU side (no GPU): the oracle's DLT triangulation is PINNED against numpy.linalg.svd (LAPACK),
an implementation neither the oracle nor the product shares code with."""
import ctypes as C

import numpy as np

from helpers import pkg, triangulate_numpy, triangulation_case


def _oracle_triangulate(oracle_lib, c, init_depth=5.0):
    est = pkg.Estimator.__new__(pkg.Estimator)  # the oracle's triangulate needs no engine state
    est.lib, est.h = oracle_lib, C.c_void_p()
    return pkg.Estimator.Triangulate(est, c["Rs"], c["Ps"], c["ric"], c["tic"], c["start_frame"], c["obs_offset"],
                                     c["obs_point"], c["depth0"], c["window_size"], init_depth)


def test_oracle_triangulation_matches_lapack(oracle_lib):
    for seed in (3, 4, 5):
        c = triangulation_case(seed=seed)
        d_np = triangulate_numpy(c)
        d_or = _oracle_triangulate(oracle_lib, c)
        used = np.diff(c["obs_offset"])
        cand = (used >= 2) & (c["start_frame"] < c["window_size"] - 2) & (c["depth0"] <= 0)
        # non-candidates and initialised depths are untouched (feature_manager.cpp:236-240)
        assert np.array_equal(d_or[~cand], c["depth0"][~cand])
        # the fallback set is identical and the depths agree to SVD accuracy (the zero-parallax tracks have a
        # near-degenerate smallest singular pair: they are compared on the fallback decision only)
        fb_np, fb_or = d_np[cand] == 5.0, d_or[cand] == 5.0
        degenerate = (np.arange(len(used)) % 29 == 0)[cand]
        assert np.array_equal(fb_np[~degenerate], fb_or[~degenerate])
        ok = ~fb_np & ~degenerate
        assert np.allclose(d_or[cand][ok], d_np[cand][ok], rtol=1e-9)
        # well-conditioned tracks recover the generating depth
        good = cand & (c["truth"] > 0) & (used >= 4) & (np.arange(len(used)) % 29 != 0)
        assert np.median(np.abs(d_or[good] / c["truth"][good] - 1)) < 0.05
        assert (fb_or.sum() > 0) and (~fb_or).sum() > 50

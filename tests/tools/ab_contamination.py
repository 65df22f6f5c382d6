"""A/B of the contamination likelihood kernel's launch geometry (VLR_CONTAM_GEOM = threads,tile,rows_per_sm) under
`ncu --metrics gpu__time_duration.sum -k regex:vlr_contam_likelihood`: per geometry one parity call on 3000 observations
(incl. one AFD larger than the shared-memory tile) checked against the oracle, then two calls on 100 000 observations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tests.test_contamination import assert_same, make_observations, oracle_posterior  # noqa: E402
from varlociraptor_b200 import contamination as ct  # noqa: E402

GEOMS = sys.argv[1:] or ["128,2048,1", "416,2048,2", "416,1024,4", "128,1024,3", "224,1024,4"]
obs = make_observations(100000, seed=1)
small = make_observations(3000, seed=2)
grid = np.unique(np.concatenate([[0.0, small[5].max_posterior_vaf, 1.0],
                                 np.round(np.random.Generator(np.random.PCG64(1)).uniform(0, 1, 6000), 5)]))
small[5].vaf_dist = list(zip(grid.tolist(), (-0.5 * ((grid - small[5].max_posterior_vaf) / 0.1) ** 2).tolist()))
want_post, _, want_marg, _ = oracle_posterior(small)
for g in GEOMS:
    os.environ["VLR_CONTAM_GEOM"] = g
    got = ct.contamination_posterior(small)
    assert_same(got.ln_posterior, got.ln_marginal, want_post, want_marg)
    for _ in range(2):
        big = ct.contamination_posterior(obs)
    print("geometry %s: parity ok (max |d ln posterior| %.2e), ln marginal of the 100k call %.9f"
          % (g, np.nanmax(np.abs(np.where(np.isfinite(want_post), got.ln_posterior - want_post, 0.0))), big.ln_marginal))

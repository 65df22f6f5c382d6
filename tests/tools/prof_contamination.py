"""Timing of the contamination model (vlr_contamination_posterior, host buffers in/out) against the oracle's
sequential restatement on the same synthetic VariantObservations. Run under ncu for per-kernel durations
(scripts/gpu_profile.sh)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle  # noqa: E402
from tests.test_contamination import make_observations  # noqa: E402
from varlociraptor_b200 import contamination as ct  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
obs = make_observations(n, seed=1)
points = sum(len(o.vaf_dist) for o in obs)
ct.contamination_posterior(obs[:10])  # context creation
best = 1e30
for _ in range(3):
    t = time.perf_counter()
    got = ct.contamination_posterior(obs)
    best = min(best, time.perf_counter() - t)
sub = obs[:max(1, n // 10)]
pd, mpv, off, vaf, logp = ct.pack_observations(sub)
t = time.perf_counter()
oracle.contamination_posterior(pd, mpv, off, vaf, logp, ct.Prior(None).table())
t_cpu = (time.perf_counter() - t) * (n / len(sub))
print("contamination model: %d observations, %d AFD points: GPU call (pack + H2D + 3 kernels + D2H) %.2f ms = %.1f M "
      "terms/s; oracle (1 thread, extrapolated from %d observations) %.0f ms; marginal %.6f"
      % (n, points, best * 1e3, n * 404 / best / 1e6, len(sub), t_cpu * 1e3, got.ln_marginal))

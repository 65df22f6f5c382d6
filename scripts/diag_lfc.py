"""Diagnostic (GPU): how often do log2-fold-change threshold ties differ between CUDA's and glibc's log2, and what does
that do to the LFC parity scenario. Not part of the product; output feeds DESIGN.md."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from oracle import oracle
from tests.test_emu_parity import LFC_YAML
from tests.util import max_abs_delta
from varlociraptor_b200 import Scenario, synth, engine

rng = np.random.default_rng(1)
a = rng.uniform(0.02, 1.0, 200000)
cpu = np.log2(a) - np.log2(a / 2)
ta = torch.tensor(a, device="cuda")
gpu = (torch.log2(ta) - torch.log2(ta / 2)).cpu().numpy()
print("glibc: lfc == 1 exactly: %.4f  >= 1: %.4f" % ((cpu == 1.0).mean(), (cpu >= 1.0).mean()))
print("cuda : lfc == 1 exactly: %.4f  >= 1: %.4f" % ((gpu == 1.0).mean(), (gpu >= 1.0).mean()))
print("decisions (>= 1) that agree: %.4f" % ((cpu >= 1.0) == (gpu >= 1.0)).mean())

flat = Scenario.from_yaml(LFC_YAML).flatten()
_, b = synth.tumor_normal(40, seed=21, depth=30)
o = oracle.call_batch(flat, b, afd_capacity=128, n_threads=8)
g = engine.PosteriorEngine(flat).call_batch(b, afd_capacity=128)
d = np.array([max_abs_delta(o.log_posteriors[i], g.log_posteriors[i]) for i in range(b.n_loci)])
print("ties", o.lfc_threshold_ties().sum(), "of", b.n_loci)
print("per-locus max |dlogpost|:", np.sort(d)[::-1][:12])
print("same n_base:", (o.n_base_events == g.n_base_events).mean(), "same best:", (o.best_event == g.best_event).mean(), "same map:", np.mean([np.array_equal(o.map_vaf[i], g.map_vaf[i], equal_nan=True) for i in range(b.n_loci)]))

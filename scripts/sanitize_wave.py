"""Small batches through the wavefront pipeline for compute-sanitizer: tumor-normal with and without AFD, depth skew
(deep passes, 32 lanes per task), and a few deferred loci (tiny pileups -> generic engine over the deferred list)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import synth, engine
from varlociraptor_b200.batch import LocusBatch
sc, b = synth.tumor_normal(300, seed=3)
flat = sc.flatten()
eng = engine.PosteriorEngine(flat)
r = eng.call_batch(b, afd_capacity=32)
print("tn afd", np.isfinite(r.log_posteriors).sum(), eng.launches)
_, b5 = synth.tumor_normal(40, seed=4, depth_range=(3, 1800))
r5 = eng.call_batch(b5)
print("skew", np.isfinite(r5.log_posteriors).sum(), eng.launches)

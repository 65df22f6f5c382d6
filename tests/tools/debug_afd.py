import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle
from varlociraptor_b200 import synth, engine
sc, b = synth.tumor_normal(1500, seed=synth.SEED_BASE + 2)
flat = sc.flatten()
o = oracle.call_batch(flat, b, afd_capacity=96, n_threads=os.cpu_count())
eng = engine.PosteriorEngine(flat)
g = eng.call_batch(b, afd_capacity=96)
ke = o.knife_edge()
print("knife", ke.sum(), "count mismatch loci", np.where((o.afd_count != g.afd_count).any(axis=1))[0][:10])
with np.errstate(invalid="ignore"):
    dv = np.abs(o.afd_vaf - g.afd_vaf)
dv[np.isnan(o.afd_vaf) & np.isnan(g.afd_vaf)] = 0
bad = np.where((dv > 0).any(axis=(1, 2)) | np.isnan(dv).any(axis=(1,2)))[0]
print("vaf mismatch loci", bad[:20], "knife-edge among them", ke[bad][:20])
for i in bad[:4]:
    for s in range(2):
        ov, op = o.afd(i, s); gv, gp = g.afd(i, s)
        if len(ov) != len(gv) or np.any(ov != gv):
            k = np.where(ov != gv)[0] if len(ov) == len(gv) else []
            print("locus", i, "sample", s, "n", len(ov), len(gv), "diff idx", k[:5], [(repr(ov[j]), repr(gv[j])) for j in k[:3]])
print("max dlogpost", np.nanmax(np.abs(np.where(o.log_posteriors == g.log_posteriors, 0, o.log_posteriors - g.log_posteriors))))

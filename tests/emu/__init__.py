"""ctypes binding of the single-lane host build of the engine core (tests/emu/emu_driver.cpp). TEST ONLY."""
import ctypes as C
import os
import subprocess

from varlociraptor_b200 import abi
from varlociraptor_b200.batch import CallResults

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvlr_engine_emu.so")
_SRC = [os.path.join(_HERE, "emu_driver.cpp"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "engine_core.cuh"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "engine_wave.cuh"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "engine_resident.cuh"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "engine_sets.cuh"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "scenario_prep.h"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "engine_types.cuh"),
        os.path.join(_HERE, "..", "..", "varlociraptor_b200", "csrc", "contamination.cuh"),
        os.path.join(_HERE, "..", "..", "include", "vlr_engine.h")]
_lib = None


def build(force=False):
    stale = not os.path.exists(_LIB) or max(os.path.getmtime(p) for p in _SRC) > os.path.getmtime(_LIB)
    if force or stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DVLR_HOST_EMU", "-ffp-contract=off",
                               "-Wno-unknown-pragmas", "-o", _LIB, _SRC[0]], cwd=_HERE)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        for fn in (_lib.vlr_emu_call_batch, _lib.vlr_emu_wave_call_batch, _lib.vlr_emu_sets_call_batch):
            fn.restype = C.c_int32
            fn.argtypes = [C.POINTER(abi.Scenario), C.POINTER(abi.Batch), C.POINTER(abi.Results)]
        _lib.vlr_emu_contamination_posterior.restype = C.c_int32
        _lib.vlr_emu_contamination_posterior.argtypes = [C.POINTER(abi.ContaminationInput),
                                                         C.POINTER(abi.ContaminationOutput), C.c_int64]
    return _lib


def wave_call_batch(flat_scenario, batch, afd_capacity=0):
    """The wavefront pipeline run sequentially on the host. Returns (results, number of loci deferred to the generic
    engine); raises LookupError when the scenario does not have the two-level chain shape."""
    lib = _load()
    out = CallResults(batch.n_loci, batch.n_samples, flat_scenario.n_events, afd_capacity)
    cb, cr = batch.as_c(), out.as_c()
    rc = lib.vlr_emu_wave_call_batch(C.byref(flat_scenario.c), C.byref(cb), C.byref(cr))
    if rc == -100:
        raise LookupError("scenario is not served by the wavefront pipeline")
    if rc < 0:
        raise RuntimeError("emu failed with status %d" % -rc)
    return out, rc


def sets_call_batch(flat_scenario, batch, afd_capacity=0):
    """The all-Set pipeline (engine_sets.cuh) run sequentially on the host. Returns (results, number of loci deferred
    to the generic engine); raises LookupError when the scenario is not an all-Set one."""
    lib = _load()
    out = CallResults(batch.n_loci, batch.n_samples, flat_scenario.n_events, afd_capacity)
    cb, cr = batch.as_c(), out.as_c()
    rc = lib.vlr_emu_sets_call_batch(C.byref(flat_scenario.c), C.byref(cb), C.byref(cr))
    if rc == -100:
        raise LookupError("scenario is not served by the all-Set pipeline")
    if rc < 0:
        raise RuntimeError("emu failed with status %d" % -rc)
    return out, rc


def call_batch(flat_scenario, batch, afd_capacity=0) -> CallResults:
    _load()
    out = CallResults(batch.n_loci, batch.n_samples, flat_scenario.n_events, afd_capacity)
    cb, cr = batch.as_c(), out.as_c()
    rc = _lib.vlr_emu_call_batch(C.byref(flat_scenario.c), C.byref(cb), C.byref(cr))
    if rc != 0:
        raise RuntimeError("emu failed with status %d" % rc)
    return out


def contamination_posterior(observations, prior_estimate=None, device=0, chunk=37):
    """`varlociraptor_b200.contamination.contamination_posterior` with the device functions of csrc/contamination.cuh
    run on the host (chunked partial sums like the kernels)."""
    import numpy as np
    from varlociraptor_b200 import contamination as ct
    lib = _load()
    prob_denovo, mpv, offsets, vaf, logp = ct.pack_observations(observations)
    emsv = np.asarray(ct.EXPECTED_MAX_SOMATIC_VAFS, dtype=np.float64)
    ln_prior = ct.Prior(prior_estimate).table(ct.N_GRID)
    post = np.empty((len(emsv), ct.N_GRID))
    lik = np.empty_like(post)
    marginal, max_vaf = np.zeros(1), np.zeros(1)
    cin = abi.ContaminationInput(len(observations), abi.ptr(prob_denovo, C.c_double), abi.ptr(mpv, C.c_double),
                                 abi.ptr(offsets, C.c_int64), abi.ptr(vaf, C.c_double), abi.ptr(logp, C.c_double),
                                 ct.N_GRID, len(emsv), abi.ptr(emsv, C.c_double), abi.ptr(ln_prior, C.c_double))
    cout = abi.ContaminationOutput(abi.ptr(post, C.c_double), abi.ptr(lik, C.c_double), abi.ptr(marginal, C.c_double),
                                   abi.ptr(max_vaf, C.c_double))
    rc = lib.vlr_emu_contamination_posterior(C.byref(cin), C.byref(cout), chunk)
    if rc != 0:
        raise RuntimeError("emu contamination model failed with status %d" % rc)
    return ct.ContaminationPosterior(post, lik, float(marginal[0]), float(max_vaf[0]), tuple(emsv.tolist()))

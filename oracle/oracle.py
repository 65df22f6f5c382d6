"""ctypes binding of the CPU oracle (oracle/vlr_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from varlociraptor_b200 import abi
from varlociraptor_b200.batch import CallResults, LocusBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvlr_oracle.so")


class OracleDiag(C.Structure):
    _fields_ = [("margin_bias", C.POINTER(C.c_double)), ("margin_adaptive", C.POINTER(C.c_double)),
                ("n_pileup_evals", C.POINTER(C.c_uint64)), ("n_read_evals", C.POINTER(C.c_uint64)),
                ("lfc_tie", C.POINTER(C.c_uint8))]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vlr_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "vlr_engine.h")
    stale = (not os.path.exists(_LIB_PATH)) or (
        os.path.exists(src) and max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(_LIB_PATH))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libvlr_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.vlr_oracle_call_batch.restype = C.c_int32
        _lib.vlr_oracle_call_batch.argtypes = [C.POINTER(abi.Scenario), C.POINTER(abi.Batch), C.POINTER(abi.Results),
                                               C.POINTER(OracleDiag), C.c_int32]
        _lib.vlr_oracle_pileup_likelihood.restype = C.c_double
        _lib.vlr_oracle_pileup_likelihood.argtypes = [C.POINTER(abi.Batch), C.c_int64, C.c_int64, C.c_double,
                                                      C.c_double, C.c_double, C.c_int32]
        _lib.vlr_oracle_contamination_posterior.restype = C.c_int32
        _lib.vlr_oracle_contamination_posterior.argtypes = [C.POINTER(abi.ContaminationInput),
                                                            C.POINTER(abi.ContaminationOutput)]
    return _lib


def contamination_posterior(prob_denovo, max_posterior_vaf, afd_offsets, afd_vaf, afd_logp, ln_prior,
                            expected_max_somatic_vafs=(0.25, 0.5, 0.75, 1.0)):
    """estimation/contamination.rs' model on CSR-packed observations: (ln_posterior[rows][n_grid], ln_likelihood,
    ln_marginal, max_vaf)."""
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    prob_denovo, mpv, vaf, logp, ln_prior = map(f64, (prob_denovo, max_posterior_vaf, afd_vaf, afd_logp, ln_prior))
    offsets = np.ascontiguousarray(afd_offsets, dtype=np.int64)
    emsv = f64(expected_max_somatic_vafs)
    post = np.empty((len(emsv), len(ln_prior)))
    lik = np.empty_like(post)
    marginal, max_vaf = np.zeros(1), np.zeros(1)
    cin = abi.ContaminationInput(len(prob_denovo), abi.ptr(prob_denovo, C.c_double), abi.ptr(mpv, C.c_double),
                                 abi.ptr(offsets, C.c_int64), abi.ptr(vaf, C.c_double), abi.ptr(logp, C.c_double),
                                 len(ln_prior), len(emsv), abi.ptr(emsv, C.c_double), abi.ptr(ln_prior, C.c_double))
    cout = abi.ContaminationOutput(abi.ptr(post, C.c_double), abi.ptr(lik, C.c_double), abi.ptr(marginal, C.c_double),
                                   abi.ptr(max_vaf, C.c_double))
    rc = lib().vlr_oracle_contamination_posterior(C.byref(cin), C.byref(cout))
    if rc != 0:
        raise RuntimeError("oracle contamination model failed with status %d" % rc)
    return post, lik, float(marginal[0]), float(max_vaf[0])


class OracleOutput(CallResults):
    def __init__(self, n_loci, n_samples, n_events, afd_capacity):
        super().__init__(n_loci, n_samples, n_events, afd_capacity)
        self.margin_bias = np.full(n_loci, np.inf)
        self.margin_adaptive = np.full(n_loci, np.inf)
        self.n_pileup_evals = np.zeros(n_loci, dtype=np.uint64)
        self.n_read_evals = np.zeros(n_loci, dtype=np.uint64)
        self.lfc_tie = np.zeros(n_loci, dtype=np.uint8)

    def knife_edge(self, bias_tol: float = 1e-9, adaptive_tol: float = 1e-7) -> np.ndarray:
        """Loci whose result hinges on a discrete decision that is within rounding noise of flipping in the
        reference algorithm itself ("knife-edge loci")."""
        return (self.margin_bias < bias_tol) | (self.margin_adaptive < adaptive_tol)

    def lfc_threshold_ties(self) -> np.ndarray:
        """Loci whose result changes when the log2-fold-change predicates that were evaluated exactly ON their
        threshold are decided the other way. The integration limits the reference infers from a predicate
        (generic.rs:148-174) put abscissae there, where `log2(a) - log2(b) >= v` hangs on the last bit of the
        platform's log2: identical between this oracle and the reference (both glibc), not across libms (CUDA)."""
        return self.lfc_tie != 0


def call_batch(flat_scenario, batch: LocusBatch, afd_capacity: int = 0, n_threads: int = 1) -> OracleOutput:
    out = OracleOutput(batch.n_loci, batch.n_samples, flat_scenario.n_events, afd_capacity)
    cb = batch.as_c()
    cr = out.as_c()
    d = OracleDiag(abi.ptr(out.margin_bias, C.c_double), abi.ptr(out.margin_adaptive, C.c_double),
                   abi.ptr(out.n_pileup_evals, C.c_uint64), abi.ptr(out.n_read_evals, C.c_uint64),
                   abi.ptr(out.lfc_tie, C.c_uint8))
    rc = lib().vlr_oracle_call_batch(C.byref(flat_scenario.c), C.byref(cb), C.byref(cr), C.byref(d), n_threads)
    if rc != 0:
        raise RuntimeError("oracle failed with status %d" % rc)
    return out


def set_legacy_is_likely(on: bool) -> None:
    """Golden-pair pinning only (see vlr_oracle.cpp, g_legacy_is_likely)."""
    lib().vlr_oracle_set_legacy_is_likely(1 if on else 0)


def pileup_likelihood(batch: LocusBatch, lo: int, hi: int, vaf: float, vaf_secondary: float = 0.0,
                      purity: float = 1.0, contaminated: bool = False) -> float:
    cb = batch.as_c()
    return float(lib().vlr_oracle_pileup_likelihood(C.byref(cb), lo, hi, vaf, vaf_secondary, purity,
                                                    1 if contaminated else 0))

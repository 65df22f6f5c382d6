"""Host side of `varlociraptor estimate contamination` (src/estimation/contamination.rs) over the CUDA engine.

Same names and roles as the reference:

  * `estimate_contamination(sample, contaminant, output, output_plot, output_max_vaf_variants, prior_estimate)`
    (contamination.rs:421-473): the fixed two-sample scenario, run through `call_generic` with
  * `ContaminationCandidateFilter` (contamination.rs:397-419) and
  * `ContaminationEstimator`, a `CallProcessor` (contamination.rs:282-395) that keeps a `VariantObservation`
    (contamination.rs:35-116) per confidently de-novo call and, in `finalize`, evaluates the second Bayesian model.

That second model (4 x 101 events x all observations) runs on the GPU through `vlr_contamination_posterior`
(include/vlr_engine.h, csrc/contamination.cuh); this module packs the observations (CSR over the AFD points), computes
the 101 prior values (`Prior::prob`, a binomial pdf - contamination.rs:137-147) and writes the tables.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import sys
from dataclasses import dataclass
from decimal import Decimal
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import abi
from .calling import Call, CallProcessor, CandidateFilter, WorkItem, call_generic
from .scenario import Scenario

EXPECTED_MAX_SOMATIC_VAFS = (0.25, 0.5, 0.75, 1.0)  # Marginal::compute (contamination.rs:222)
N_GRID = 101                                        # ln_simpsons_integrate_exp(density, 0.0, 1.0, 101) (:233)

CONTAMINATION_SCENARIO = """
samples:
  sample:
    resolution: 0.01
    universe: "[0.0,1.0]"
  contaminant:
    resolution: 0.01
    universe: "[0.0,1.0]"
events:
  denovo:  "sample:]0.0,1.0] & contaminant:0.0"
  other: "sample:[0.0,1.0] & contaminant:]0.0,1.0]"
"""  # contamination.rs:436-449


@dataclass(frozen=True)
class PriorEstimate:
    """contamination.rs:276-280"""
    contamination: float
    n_observed_cells: int


class VariantObservation:
    """contamination.rs:35-42. The AFD (`vaf_dist`, a BTreeMap upstream) is kept as two ascending numpy columns so that
    packing 10^5 observations for the device is a concatenation, not a Python loop over points."""
    __slots__ = ("prob_denovo", "vafs", "densities", "max_posterior_vaf", "chrom", "pos")

    def __init__(self, prob_denovo: float, vaf_dist, max_posterior_vaf: float, chrom: str, pos: int):
        self.prob_denovo = float(prob_denovo)
        self.max_posterior_vaf = float(max_posterior_vaf)
        self.chrom, self.pos = chrom, pos
        self.vaf_dist = vaf_dist

    @property
    def vaf_dist(self) -> List[Tuple[float, float]]:
        """[(vaf, ln posterior density)] in ascending VAF order."""
        return list(zip(self.vafs.tolist(), self.densities.tolist()))

    @vaf_dist.setter
    def vaf_dist(self, pairs) -> None:
        a = np.asarray(list(pairs), dtype=np.float64).reshape(-1, 2)
        self.vafs, self.densities = np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])

    @classmethod
    def new(cls, call: Call, sample_names: Sequence[str]) -> Optional["VariantObservation"]:
        """VariantObservation::new (contamination.rs:44-82): None unless P(denovo) >= 0.95 and an AFD exists."""
        info = call.sample_info[list(sample_names).index("sample")]
        prob_denovo = call.event_probs["denovo"]
        if info is None or info.vaf_dist is None or math.exp(prob_denovo) < 0.95:
            return None
        dist: Dict[float, float] = {}
        for vaf, density in info.vaf_dist:  # collect() into a BTreeMap: a repeated key keeps the last value
            dist[float(vaf)] = float(density)
        return cls(float(prob_denovo), sorted(dist.items()), float(info.allelefreq_estimate), call.chrom, call.pos)


def binomial_pdf(k: int, p: float, n: int) -> float:
    """GSL `gsl_ran_binomial_pdf(k, p, n)` (rgsl, Cargo.toml; library not vendored), restated from its published
    algorithm: exp(lnchoose(n, k) + k ln p + (n - k) log1p(-p)) with the p = 0 / p = 1 corner cases."""
    if k > n:
        return 0.0
    if p == 0.0:
        return 1.0 if k == 0 else 0.0
    if p == 1.0:
        return 1.0 if k == n else 0.0
    ln_cnk = math.lgamma(n + 1.0) - math.lgamma(k + 1.0) - math.lgamma(n - k + 1.0)
    return math.exp(ln_cnk + k * math.log(p) + (n - k) * math.log1p(-p))


def grid_contamination(i: int, n_grid: int = N_GRID) -> float:
    """itertools-num linspace(0.0, 1.0, n_grid) as rust-bio's Simpson rule walks it."""
    return 0.0 + float(i) * ((1.0 - 0.0) / float(n_grid - 1))


class Prior:
    """contamination.rs:118-157"""

    def __init__(self, prior_estimate: Optional[PriorEstimate]):
        self.prior_estimate = prior_estimate

    def prob(self, contamination: float) -> float:
        if self.prior_estimate is None:
            return 0.0  # LogProb::ln_one()
        n = self.prior_estimate.n_observed_cells
        k = int(_rust_round(self.prior_estimate.contamination * n))
        pdf = binomial_pdf(k, contamination, n)
        return math.log(pdf) if pdf > 0.0 else -math.inf

    def table(self, n_grid: int = N_GRID) -> np.ndarray:
        return np.array([self.prob(grid_contamination(i, n_grid)) for i in range(n_grid)], dtype=np.float64)


def _rust_round(x: float) -> float:
    """f64::round: half away from zero (Python's round() is half to even)."""
    return math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5)


def _rust_f64(x: float) -> str:
    """`format!("{}", f64)`: shortest round-trip digits, never an exponent."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    if x == int(x) and abs(x) < 1e16:
        return "%d" % int(x) if x != 0 or math.copysign(1.0, x) > 0 else "-0"
    return format(Decimal(repr(x)), "f")


def pack_observations(observations: Sequence[VariantObservation]):
    """CSR columns of `vlr_contamination_input_t`."""
    n = len(observations)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(o.vafs) for o in observations], out=offsets[1:])
    vaf = np.concatenate([o.vafs for o in observations]) if n else np.empty(0, dtype=np.float64)
    logp = np.concatenate([o.densities for o in observations]) if n else np.empty(0, dtype=np.float64)
    prob_denovo = np.array([o.prob_denovo for o in observations], dtype=np.float64)
    mpv = np.array([o.max_posterior_vaf for o in observations], dtype=np.float64)
    return prob_denovo, mpv, offsets, vaf, logp


@dataclass
class ContaminationPosterior:
    """`ModelInstance` of the second model: `ln_posterior[k][i]` for expected maximum somatic VAF k and
    contamination grid point i."""
    ln_posterior: np.ndarray
    ln_likelihood: np.ndarray
    ln_marginal: float
    max_vaf: float
    expected_max_somatic_vafs: Tuple[float, ...]

    def event_posteriors(self) -> List[Tuple[float, float, float]]:
        """(expected_max_somatic_vaf, contamination, ln posterior), highest posterior first (ties: grid order;
        the reference's HashMap leaves tie order open)."""
        n_grid = self.ln_posterior.shape[1]
        rows = [(self.expected_max_somatic_vafs[k], grid_contamination(i, n_grid), float(self.ln_posterior[k, i]))
                for k in range(self.ln_posterior.shape[0]) for i in range(n_grid)]
        rows.sort(key=lambda r: (math.isnan(r[2]), -r[2] if not math.isnan(r[2]) else 0.0))
        return rows


def contamination_posterior(observations: Sequence[VariantObservation], prior_estimate: Optional[PriorEstimate] = None,
                            device: int = 0, n_grid: int = N_GRID,
                            expected_max_somatic_vafs: Sequence[float] = EXPECTED_MAX_SOMATIC_VAFS
                            ) -> ContaminationPosterior:
    """`Model::compute_from_marginal(&Marginal, &variant_observations)` (contamination.rs:312-320) on the GPU."""
    from .engine import EngineError, lib
    prob_denovo, mpv, offsets, vaf, logp = pack_observations(observations)
    emsv = np.asarray(expected_max_somatic_vafs, dtype=np.float64)
    ln_prior = Prior(prior_estimate).table(n_grid)
    post = np.empty((len(emsv), n_grid), dtype=np.float64)
    lik = np.empty_like(post)
    marginal = np.zeros(1, dtype=np.float64)
    max_vaf = np.zeros(1, dtype=np.float64)
    cin = abi.ContaminationInput(len(observations), abi.ptr(prob_denovo, C.c_double), abi.ptr(mpv, C.c_double),
                                 abi.ptr(offsets, C.c_int64), abi.ptr(vaf, C.c_double), abi.ptr(logp, C.c_double),
                                 n_grid, len(emsv), abi.ptr(emsv, C.c_double), abi.ptr(ln_prior, C.c_double))
    cout = abi.ContaminationOutput(abi.ptr(post, C.c_double), abi.ptr(lik, C.c_double), abi.ptr(marginal, C.c_double),
                                   abi.ptr(max_vaf, C.c_double))
    rc = lib().vlr_contamination_posterior(device, C.byref(cin), C.byref(cout))
    if rc != 0:
        raise EngineError("vlr_contamination_posterior failed: %s" % lib().vlr_status_string(rc).decode())
    return ContaminationPosterior(post, lik, float(marginal[0]), float(max_vaf[0]), tuple(float(x) for x in emsv))


class DeviceObservations:
    """The VariantObservations of a batch of calls, selected and packed on the device from device-resident results
    (`vlr_contamination_gather_device`: VariantObservation::new, contamination.rs:44-82, for every call at once). torch
    tensors are only the allocator."""

    def __init__(self, dev_results, sample: int, denovo_event: int, min_prob: float = 0.95, device: int = 0, stream: int = 0):
        import torch
        from .engine import EngineError, lib
        n, S, cap = dev_results.n_loci, dev_results.n_samples, dev_results.afd_capacity
        if cap < 1:
            raise ValueError("the calls carry no allele frequency distributions (afd_capacity = 0)")
        dev = dev_results.log_posteriors.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.prob_denovo = torch.empty(max(n, 1), **f64)
        self.max_posterior_vaf = torch.empty(max(n, 1), **f64)
        self.afd_offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
        self.afd_vaf = torch.empty(max(n * cap, 1), **f64)
        self.afd_logp = torch.empty(max(n * cap, 1), **f64)
        self.kept_loci = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        n_obs = C.c_int64(0)
        cr = dev_results.as_c()
        rc = lib().vlr_contamination_gather_device(
            device, C.byref(cr), n, S, dev_results.n_events, sample, denovo_event, float(min_prob),
            C.c_void_p(self.prob_denovo.data_ptr()), C.c_void_p(self.max_posterior_vaf.data_ptr()),
            C.c_void_p(self.afd_offsets.data_ptr()), C.c_void_p(self.afd_vaf.data_ptr()), C.c_void_p(self.afd_logp.data_ptr()),
            C.c_void_p(self.kept_loci.data_ptr()), C.byref(n_obs), C.c_void_p(stream or None))
        if rc != 0:
            raise EngineError("vlr_contamination_gather_device failed: %s" % lib().vlr_status_string(rc).decode())
        self.n_obs = int(n_obs.value)
        self.device = device


def contamination_posterior_device(obs: DeviceObservations, prior_estimate: Optional[PriorEstimate] = None, n_grid: int = N_GRID,
                                   expected_max_somatic_vafs: Sequence[float] = EXPECTED_MAX_SOMATIC_VAFS,
                                   stream: int = 0) -> ContaminationPosterior:
    """The second model over observations that never left the device (`vlr_contamination_posterior_device`)."""
    import torch
    from .engine import EngineError, lib
    dev = obs.prob_denovo.device
    emsv = torch.tensor(list(expected_max_somatic_vafs), dtype=torch.float64, device=dev)
    ln_prior = torch.from_numpy(Prior(prior_estimate).table(n_grid)).to(dev)
    post = torch.empty((len(emsv), n_grid), dtype=torch.float64, device=dev)
    lik = torch.empty_like(post)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)  # marginal, max_vaf
    p = lambda t, ct: abi.devptr(t.data_ptr(), ct)  # noqa: E731
    cin = abi.ContaminationInput(obs.n_obs, p(obs.prob_denovo, C.c_double), p(obs.max_posterior_vaf, C.c_double),
                                 p(obs.afd_offsets, C.c_int64), p(obs.afd_vaf, C.c_double), p(obs.afd_logp, C.c_double),
                                 n_grid, len(emsv), p(emsv, C.c_double), p(ln_prior, C.c_double))
    cout = abi.ContaminationOutput(p(post, C.c_double), p(lik, C.c_double), abi.devptr(scal.data_ptr(), C.c_double),
                                   abi.devptr(scal.data_ptr() + 8, C.c_double))
    rc = lib().vlr_contamination_posterior_device(obs.device, C.byref(cin), C.byref(cout), C.c_void_p(stream or None))
    if rc != 0:
        raise EngineError("vlr_contamination_posterior_device failed: %s" % lib().vlr_status_string(rc).decode())
    torch.cuda.synchronize()
    sc = scal.cpu().numpy()
    return ContaminationPosterior(post.cpu().numpy(), lik.cpu().numpy(), float(sc[0]), float(sc[1]),
                                  tuple(float(x) for x in expected_max_somatic_vafs))


class ContaminationEstimator(CallProcessor):
    """contamination.rs:282-395. `posterior_fn` exists for the host tests (an emulated engine); the product default
    is the CUDA path."""

    def __init__(self, output: Optional[str] = None, output_plot: Optional[str] = None,
                 output_max_vaf_variants: Optional[str] = None, prior_estimate: Optional[PriorEstimate] = None,
                 device: int = 0, posterior_fn=None):
        self.output = output
        self.output_plot = output_plot
        self.output_max_vaf_variants = output_max_vaf_variants
        self.prior_estimate = prior_estimate
        self.device = device
        self.variant_observations: List[VariantObservation] = []
        self.posterior: Optional[ContaminationPosterior] = None
        self._posterior_fn = posterior_fn or contamination_posterior

    def process_call(self, call: Call, sample_names: Sequence[str]) -> None:
        obs = VariantObservation.new(call, sample_names)
        if obs is not None:
            self.variant_observations.append(obs)

    def histogram(self) -> List[Tuple[float, int]]:
        """VAFDist::new's histogram (contamination.rs:249-254): bins of floor(vaf * 100) / 100."""
        hist: Dict[float, int] = {}
        for o in self.variant_observations:
            b = math.floor(o.max_posterior_vaf * 100.0) / 100.0
            hist[b] = hist.get(b, 0) + 1
        return sorted(hist.items())

    def calc_posterior(self, writer) -> None:
        """contamination.rs:308-377: the posterior table (TSV) plus the optional plot data and max-VAF variants."""
        self.posterior = m = self._posterior_fn(self.variant_observations, self.prior_estimate, device=self.device)
        if self.output_plot:
            prior = Prior(self.prior_estimate)
            n_grid = m.ln_posterior.shape[1]
            densities = [{"purity": 1.0 - grid_contamination(i, n_grid),
                          "density": math.exp(prior.prob(grid_contamination(i, n_grid))), "category": "prior"}
                         for i in range(n_grid)]
            densities += [{"purity": 1.0 - c, "density": math.exp(p),
                           "category": "posterior, max VAF=%s" % _rust_f64(v)} for v, c, p in m.event_posteriors()]
            # the reference embeds these two datasets into its vega-lite template (templates/plots/
            # contamination_estimation.json, not reproduced here)
            spec = {"datasets": {"empirical_vaf_dist": [{"vaf": v, "count": c} for v, c in self.histogram()],
                                 "densities": densities}}
            with open(self.output_plot, "w") as f:
                json.dump(spec, f, indent=2)
        if self.output_max_vaf_variants:
            with open(self.output_max_vaf_variants, "w") as f:
                f.write("chrom,pos\n")
                for o in self.variant_observations:
                    if o.max_posterior_vaf == m.max_vaf:
                        f.write("%s,%d\n" % (o.chrom, o.pos))
        writer.write("maximum somatic VAF\tcontamination\tposterior density\n")
        for v, c, p in m.event_posteriors():
            writer.write("%s\t%s\t%s\n" % (_rust_f64(v), _rust_f64(c), _rust_f64(math.exp(p))))

    def finalize(self) -> None:
        if self.output:
            with open(self.output, "w") as f:
                self.calc_posterior(f)
        else:
            self.calc_posterior(sys.stdout)


class ContaminationCandidateFilter(CandidateFilter):
    """contamination.rs:397-419: SNVs whose contaminant reads (>= 10) all support the reference and whose sample
    pileup (>= 10 reads) has at least one strong alt read."""

    def filter(self, work_item: WorkItem, sample_names: Sequence[str]) -> bool:  # noqa: A003
        names = list(sample_names)
        b = work_item.pileups

        def reads(name):
            s = names.index(name)
            lo, hi = int(b.read_offsets[s]), int(b.read_offsets[s + 1])
            return (b.columns["prob_alt"][lo:hi].astype(np.float64), b.columns["prob_ref"][lo:hi].astype(np.float64))

        if not (work_item.locus_flags & abi.LF_HAS_SNV):
            return False
        c_alt, c_ref = reads("contaminant")
        if len(c_alt) < 10 or not bool(np.all(c_ref > c_alt)):  # is_ref_support (read_observation.rs:439-441)
            return False
        s_alt, s_ref = reads("sample")
        if len(s_alt) < 10:
            return False
        with np.errstate(over="ignore"):  # is_strong_alt_support: Kass-Raftery >= Strong, i.e. exp(pa - pr) > 20
            return bool(np.any(np.exp(s_alt - s_ref) > 20.0))


def estimate_contamination(sample: Iterable[dict], contaminant: Iterable[dict], output: Optional[str] = None,
                           output_plot: Optional[str] = None, output_max_vaf_variants: Optional[str] = None,
                           prior_estimate: Optional[PriorEstimate] = None, **engine_kwargs) -> ContaminationEstimator:
    """contamination.rs:421-473. `sample` / `contaminant` are the records of the two observation files."""
    estimator = ContaminationEstimator(output, output_plot, output_max_vaf_variants, prior_estimate,
                                       device=engine_kwargs.get("device", 0))
    call_generic(Scenario.from_yaml(CONTAMINATION_SCENARIO), {"sample": sample, "contaminant": contaminant},
                 call_processor=estimator, candidate_filter=ContaminationCandidateFilter(), **engine_kwargs)
    return estimator

"""Independent restatement of the whole per-locus path for tumor-normal loci WITHOUT a surviving artifact config:
GenericPosterior::density (generic.rs:191-422), the adaptive integration (adaptive_integration.rs:25-141), the
observable range limits (formula.rs:1172-1224), the contaminated / single sample likelihoods (likelihood.rs) with the
learned forward-strand rate (strand_bias.rs:79-123) and the posterior normalisation, written here straight from the
reference's files in Python: abscissae in f64 exactly as the reference computes them, likelihood VALUES with 50
significant digits in linear space (mpmath). The oracle (and with it the engines) must reproduce event posteriors
within 1e-9 and the same MAP allele frequencies on loci whose decisions are not knife-edge.

What this adds to the golden pair (single sample, printed precision): nested integration over two samples, the
contaminated sample model, Set and Range nodes under clear-reference pruning - at full precision. What it does not
cover: artifact events (loci are chosen so that no artifact config is possible, informative and likely)."""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import oracle
from varlociraptor_b200 import Scenario, abi, synth

mp.mp.dps = 50
HALF = mp.mpf("0.5")


def _e(x):
    x = float(x)
    return mp.mpf(0) if x == -math.inf else mp.e ** mp.mpf(x)


class Pileup:
    """Per-read linear coefficients of one sample's pileup under Artifacts::none(): value(x) = c0 + c1 x, x the
    effective alt-sampling probability (every read of these batches has prob_sample_alt = 0: likelihood.rs:43-53)."""

    def __init__(self, b, lo, hi, forward_rate):
        c = b.columns
        self.n = hi - lo
        self.c0, self.c1, self.pos_ref = [], [], True
        for r in range(lo, hi):
            assert float(c["prob_sample_alt"][r]) == 0.0
            f = int(b.read_flags[r])
            strand = (f >> abi.RF_STRAND_SHIFT) & 3
            pdo = _e(c["prob_double_overlap"][r])
            phb = _e(c["prob_hit_base"][r])
            rpb = phb if f & abi.RF_READPOS_MAJOR else 1 - phb
            sb_alt = {0: forward_rate * (1 - pdo), 1: (1 - forward_rate) * (1 - pdo), 2: pdo, 3: mp.mpf(1)}[strand]
            b_alt = sb_alt * HALF * rpb * HALF   # strand, orientation, position, softclip = homopolymer = 1, alt locus
            b_ref = HALF * HALF * rpb * HALF
            b_any = HALF * HALF * rpb * HALF
            pm = _e(c["prob_mapping"][r])
            a_term = pm * b_alt * _e(c["prob_alt"][r])
            r_term = pm * b_ref * _e(c["prob_ref"][r])
            self.c0.append(r_term + (1 - pm) * _e(c["prob_missed_allele"][r]) * b_any)
            self.c1.append(a_term - r_term)
            # is_positive_ref_support (read_observation.rs:443-446): Kass-Raftery of exp(prob_ref - prob_alt) > 3
            if not math.exp(float(c["prob_ref"][r]) - float(c["prob_alt"][r])) > 3.0:
                self.pos_ref = False
        self.clear_ref = self.n > 10 and self.pos_ref  # generic.rs:270-291

    def value(self, x):
        x = mp.mpf(x)
        p = mp.mpf(1)
        for c0, c1 in zip(self.c0, self.c1):
            p *= c0 + c1 * x
        return p


def _forward_rate(b, rows):
    """StrandBias::estimate_forward_rate over all samples' reads (strand_bias.rs:79-123), else 0.5 (:69-77)."""
    c = b.columns
    strong_all = strong_fwd = 0.0
    for r in rows:
        strand = (int(b.read_flags[r]) >> abi.RF_STRAND_SHIFT) & 3
        strong_ref = math.exp(float(c["prob_ref"][r]) - float(c["prob_alt"][r])) > 20.0  # >= KassRaftery::Strong
        if strong_ref and strand != 2:
            strong_all += math.exp(float(c["prob_mapping"][r]))
        if strong_ref and strand == 0:
            strong_fwd += math.exp(float(c["prob_mapping"][r]))
    if strong_all > 2.0:
        frac = strong_fwd / strong_all
        if strong_all > 100.0 and 0.0 < frac < 1.0:
            return mp.mpf(frac)
        if 0.4 <= frac <= 0.6:
            return HALF
    return HALF


def _observable_max(rg, n_obs):
    if n_obs < 10 or not (n_obs * (rg.end - rg.start) > 1.0):
        return rg.end
    cnt = n_obs * rg.end
    if rg.right_exclusive and cnt % 1.0 == 0.0:
        cnt -= 1.0
    cnt = math.floor(cnt)
    return rg.end if cnt == 0.0 else math.floor(cnt) / n_obs


def _observable_min(rg, n_obs):
    if n_obs < 10 or not (n_obs * (rg.end - rg.start) > 1.0):
        mn = rg.start
    else:
        cnt = n_obs * rg.start
        adjust = lambda k: math.ceil(k) / n_obs  # noqa: E731
        if rg.left_exclusive and cnt % 1.0 == 0.0:
            end = _observable_max(rg, n_obs)
            for off in (1.0, 0.0):
                s = adjust(cnt + off)
                if s <= 1.0 and s <= end:
                    return s  # (a `return` inside the block upstream: the order check below is not applied to it)
        mn = adjust(cnt)
    return rg.start if mn >= _observable_max(rg, n_obs) else mn


def _integrate(density, lo, hi, res):
    """adaptive_integration.rs:25-141; density returns a LINEAR value; the integral is returned linear."""
    probs = {}

    def gp(x):
        probs[x] = density(x)
        return x
    left, right = gp(lo), gp(hi)
    first_middle = middle = None
    while ((right - left) >= res and left < right) or middle is None:
        middle = gp((right + left) / 2.0)
        m1 = gp((middle + left) / 2.0)
        m2 = gp((right + middle) / 2.0)
        if first_middle is None:
            first_middle = middle
        xs = [left, m1, m2, right]
        best = 0
        for k in range(4):
            if probs[xs[k]] > probs[xs[best]]:
                best = k
        left, right = (xs[best - 1] if best > 0 else xs[best]), (xs[best + 1] if best < 3 else xs[best])
    gp((first_middle + hi) / 2.0 if middle < first_middle else (lo + first_middle) / 2.0)
    a = max(middle - res * 3.0, lo)
    step = (middle - a) / 3.0
    for i in range(3):
        gp(a + step * i)
    z = min(middle + res * 3.0, hi)
    step = (z - middle) / 3.0
    for i in range(1, 4):
        gp(middle + step * i)
    xs = sorted(probs)
    return sum(((probs[x0] + probs[x1]) / 2 * (mp.mpf(x1) - mp.mpf(x0)) for x0, x1 in zip(xs, xs[1:])), mp.mpf(0))


class Locus:
    def __init__(self, b, locus, flat):
        S = 2
        offs = [int(b.read_offsets[locus * S + k]) for k in range(S + 1)]
        fr = _forward_rate(b, range(offs[0], offs[2]))
        self.normal, self.tumor = Pileup(b, offs[0], offs[1], fr), Pileup(b, offs[1], offs[2], fr)
        self.res = [0.1, 0.01]                      # cli.rs:1151-1173: normal 0.1, tumor 0.01
        self.purity = mp.mpf(1) - mp.mpf(0.25)      # tumor contaminated by normal, fraction = 1 - purity
        self.base = {}                              # (vn, vt) -> joint, in evaluation order (first entry kept)
        self.base_full, self._disc = {}, {}
        self.n_joint = 0
        self._ln = {}

    def joint(self, vn, vt):
        """Prior (flat inside the universes: prior.rs:398-406) x normal pileup x contaminated tumor pileup."""
        self.n_joint += 1
        if vn not in self._ln:
            self._ln[vn] = self.normal.value(vn)
        x_eff = self.purity * mp.mpf(vt) + (1 - self.purity) * mp.mpf(vn)   # likelihood.rs:108-135, s_r = 1
        j = self._ln[vn] * self.tumor.value(x_eff)
        self.base.setdefault((vn, vt), j)
        # base events as the reference keys them (LikelihoodOperands: allele frequency AND is_discrete per sample)
        self.base_full.setdefault((vn, self._disc.get(0, True), vt, self._disc.get(1, True)), j)
        return j

    def node(self, nd, fixed):
        """density() of a Sample node; `fixed` = allele frequency of the enclosing (normal) node or None."""
        pile, res = (self.normal, self.res[0]) if nd.sample == 0 else (self.tumor, self.res[1])

        def below(v):
            if nd.children:
                return self.node(nd.children[0], v)
            return self.joint(fixed, v)
        self._disc[nd.sample] = nd.kind == 0  # push_base_event(.., is_discrete): Set values true, integration points false
        if nd.kind == 0:  # Set
            vafs = sorted(nd.vafs)
            if pile.clear_ref and all(v > 0.0 for v in vafs):
                return mp.mpf(0)
            return sum((below(v) for v in vafs), mp.mpf(0))
        rg = nd.vafs
        if pile.clear_ref and rg.start > 0.0:
            return mp.mpf(0)
        mn, mx = _observable_min(rg, pile.n), _observable_max(rg, pile.n)
        assert mn <= mx
        if (mx - mn) < res:       # generic.rs:357-394
            return _simpson(below, mn, mx, 3)
        if pile.n < 5:
            return _simpson(below, mn, mx, 11)
        return _integrate(below, mn, mx, res)


def _chosen_loci(o, b, want=9):
    """Loci the restatement covers: no artifact event, nothing adjusted or filtered, decisions not knife-edge; a mix of
    best events."""
    ok = np.isneginf(o.log_posteriors[:, -1]) & (o.status == 0) & ~o.knife_edge()
    picked, seen = [], {}
    for i in np.nonzero(ok)[0]:
        k = int(o.best_event[i]) // 2
        if seen.get(k, 0) < 2:
            seen[k] = seen.get(k, 0) + 1
            picked.append(int(i))
        if len(picked) >= want:
            break
    return picked


def _results(source, flat, b, afd_capacity):
    """What is checked against the restatement: the oracle, or the engine itself (the single-lane host build of the
    kernel source, tests/emu) - the loci are chosen by the oracle's diagnostics either way."""
    o = oracle.call_batch(flat, b, afd_capacity=afd_capacity, n_threads=4)
    if source == "oracle":
        return o, o
    from tests import emu
    if source == "pipeline":  # the wavefront pipeline (engine_wave.cuh + engine_resident.cuh) instead of the generic engine
        return o, emu.wave_call_batch(flat, b, afd_capacity=afd_capacity)[0]
    return o, emu.call_batch(flat, b, afd_capacity=afd_capacity)


@pytest.mark.parametrize("n_loci,seed,source", [(80, 21, "oracle"), (120, 22, "oracle"), (120, 22, "engine"),
                                                (120, 22, "pipeline")])
def test_tumor_normal_posteriors_against_the_high_precision_restatement(n_loci, seed, source):
    sc, b = synth.tumor_normal(n_loci, seed=seed)
    flat = sc.flatten()
    diag, o = _results(source, flat, b, 192)
    trees = dict(sc.event_trees())
    names = list(flat.event_names)
    loci = _chosen_loci(diag, b)
    assert len(loci) >= 5
    worst = 0.0
    for i in loci:
        # the restated selection rules (further down) agree that no artifact config survives here: >= 10 strong alt reads
        # without a two-thirds majority for any bias, or nothing but reference support
        offs = [int(b.read_offsets[i * 2 + k]) for k in range(3)]
        piles = [Reads(b, offs[0], offs[1]), Reads(b, offs[1], offs[2])]
        assert _surviving_configs(piles, _forward_rate_opt([d for p in piles for d in p.rows]) is not None) == []
        L = Locus(b, i, flat)
        dens = []
        for name in names:
            roots = trees[name]
            dens.append(sum((L.node(r, None) for r in roots), mp.mpf(0)))
        total = sum(dens, mp.mpf(0))
        for k, d in enumerate(dens):
            got = float(o.log_posteriors[i, k])
            if d == 0:
                assert got == -math.inf, (i, names[k], got)
                continue
            want = mp.log(d / total)
            delta = abs(float(mp.mpf(got) - want))
            worst = max(worst, delta)
            assert delta <= 1e-9, (i, names[k], got, float(want), delta)
        assert L.n_joint == int(o.n_base_events[i])  # the same number of joint evaluations: identical adaptive grids
        best = max(range(len(dens)), key=lambda k: dens[k])
        assert 2 * best == int(o.best_event[i])  # index into the event universe: 2 e (plain), 2 e + 1 (artifact twin)
        # MAP: the strongest base event the best event contains (calling.rs:851-864); here: of its own tree
        L2 = Locus(b, i, flat)
        for r in trees[names[best]]:
            L2.node(r, None)
        (vn, vt), _ = max(L2.base.items(), key=lambda kv: kv[1])
        assert (vn, vt) == (float(o.map_vaf[i, 0]), float(o.map_vaf[i, 1])), (i, names[best], vn, vt, o.map_vaf[i])
        # allele frequency distributions (calling.rs:891-928): per sample, the base events of ALL events that the best
        # event contains when that sample is ignored and whose other sample equals the MAP (frequency and discreteness)
        (mvn, mdn, mvt, mdt), _ = max(L2.base_full.items(), key=lambda kv: kv[1])
        assert (mvn, mvt) == (vn, vt)

        def inside(nd, v):
            if nd.kind == 0:
                return v in nd.vafs
            r = nd.vafs
            return (r.start < v or (v == r.start and not r.left_exclusive)) and (v < r.end or (v == r.end and not r.right_exclusive))
        for sample in (0, 1):
            want_afd = {}
            for (bn, dn, bt, dt), j in L.base_full.items():
                ok_other = (bt, dt) == (mvt, mdt) if sample == 0 else (bn, dn) == (mvn, mdn)
                compatible = any(inside(r.children[0], bt) if sample == 0 else inside(r, bn) for r in trees[names[best]])
                if ok_other and compatible:
                    # posterior of a base event = joint - marginal; the joint carries no bias prior, the marginal is over
                    # the events' ln 0.5 + density (rust-bio Model::compute, generic.rs:430-460)
                    want_afd.setdefault(bn if sample == 0 else bt, []).append(
                        mp.log(j / (HALF * total)) if j > 0 else mp.mpf("-inf"))
            vaf, logp = o.afd(i, sample)
            assert sorted(want_afd) == sorted(set(float(v) for v in vaf)), (i, sample, sorted(want_afd)[:5], sorted(vaf)[:5])
            assert len(vaf) == sum(len(v) for v in want_afd.values())
            for v, lp in zip(vaf, logp):
                cands = want_afd[float(v)]
                if lp == -math.inf:
                    assert any(c == mp.mpf("-inf") for c in cands)
                else:
                    assert min(abs(float(mp.mpf(float(lp)) - c)) for c in cands if c != mp.mpf("-inf")) <= 1e-9, (i, sample, v, lp)
    assert worst > 0.0
    if seed == 22:  # this batch has somatic_normal calls without artifact events: the nested Range x Range integration
        assert any(int(o.best_event[i]) // 2 == names.index("somatic_normal") for i in loci)


# ---------------------------------------------------------------------------------------------- pedigree (config 3)
class GeneralPileup(Pileup):
    """Single-sample pileup value at an allele frequency with per-read prob_sample_alt (likelihood.rs:43-53, :198-220)."""

    def __init__(self, b, lo, hi, forward_rate):
        c = b.columns
        self.n = hi - lo
        self.rows, self.pos_ref = [], True
        for r in range(lo, hi):
            f = int(b.read_flags[r])
            strand = (f >> abi.RF_STRAND_SHIFT) & 3
            pdo = _e(c["prob_double_overlap"][r])
            phb = _e(c["prob_hit_base"][r])
            rpb = phb if f & abi.RF_READPOS_MAJOR else 1 - phb
            sb_alt = {0: forward_rate * (1 - pdo), 1: (1 - forward_rate) * (1 - pdo), 2: pdo, 3: mp.mpf(1)}[strand]
            pm = _e(c["prob_mapping"][r])
            self.rows.append((pm * sb_alt * HALF * rpb * HALF * _e(c["prob_alt"][r]),
                              pm * HALF * HALF * rpb * HALF * _e(c["prob_ref"][r]),
                              (1 - pm) * _e(c["prob_missed_allele"][r]) * HALF * HALF * rpb * HALF,
                              _e(c["prob_sample_alt"][r])))
            if not math.exp(float(c["prob_ref"][r]) - float(c["prob_alt"][r])) > 3.0:
                self.pos_ref = False
        self.clear_ref = self.n > 10 and self.pos_ref

    def value(self, vaf):
        v = mp.mpf(vaf)
        p = mp.mpf(1)
        for a, r, m, sa in self.rows:
            s = mp.mpf(1) if v == 1 else min(v * sa, mp.mpf(1))
            p *= s * a + (1 - s) * r + m
        return p


def _pedigree_prior(vafs, names, het):
    """Prior::compute in its default absent-only mode (prior.rs:737-754) for the simple pedigree (mother, father
    founders; child Mendelian; ploidy 2): P(all absent) = population term with m = 0 over the founders' four alleles
    (prior.rs:554-582) x Mendelian term 1; any other possible combination gets 1 - P(all absent); impossible ones (the
    child carries fewer alt alleles than a homozygous parent must pass on: prior.rs:600-678) get 0."""
    n_alt = {nm: int(round(2 * v)) for nm, v in zip(names, vafs)}
    p_absent = 1 - sum(het / m for m in range(1, 5))
    if all(k == 0 for k in n_alt.values()):
        return p_absent
    must = (n_alt["mother"] == 2) + (n_alt["father"] == 2)
    return mp.mpf(0) if n_alt["child"] < must else 1 - p_absent


@pytest.mark.parametrize("source", ["oracle", "engine", "all-Set pipeline"])
def test_pedigree_posteriors_against_the_high_precision_restatement(source):
    sc, b = synth.pedigree(150, seed=31)
    flat = sc.flatten()
    names = list(sc.sample_names)
    events = list(flat.event_names)
    S = len(names)
    diag = o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    if source != "oracle":
        from tests import emu
        o = emu.call_batch(flat, b) if source == "engine" else emu.sets_call_batch(flat, b)[0]
    trees = dict(sc.event_trees())
    ok = np.isneginf(diag.log_posteriors[:, -1]) & (diag.status == 0) & ~diag.knife_edge()
    picked, seen = [], {}
    for i in np.nonzero(ok)[0]:
        key = (int(o.best_event[i]) // 2, int((b.locus_flags[i] >> abi.LF_VARTYPE_SHIFT) & 3))
        if seen.get(key, 0) < 2:
            seen[key] = seen.get(key, 0) + 1
            picked.append(int(i))
    assert len(picked) >= 6 and len({k[1] for k in seen}) >= 2  # SNV and indel loci (their heterozygosities differ)
    worst = 0.0
    for i in picked:
        offs = [int(b.read_offsets[i * S + k]) for k in range(S + 1)]
        fr = _forward_rate(b, range(offs[0], offs[S]))
        piles = [GeneralPileup(b, offs[k], offs[k + 1], fr) for k in range(S)]
        vartype = int((b.locus_flags[i] >> abi.LF_VARTYPE_SHIFT) & 3)
        het = mp.mpf("0.001") * (1 if vartype == 0 else mp.mpf("0.0125"))  # grammar/mod.rs:375-415, prior.rs:243-271
        cache = {}
        n_joint = 0

        def lh(s, v):
            if (s, v) not in cache:
                cache[(s, v)] = piles[s].value(v)
            return cache[(s, v)]

        def walk(nd, vafs):
            nonlocal n_joint
            assert nd.kind == 0
            vs = sorted(nd.vafs)
            if piles[nd.sample].clear_ref and all(v > 0.0 for v in vs):  # generic.rs:294-299
                return mp.mpf(0)
            tot = mp.mpf(0)
            for v in vs:
                cur = dict(vafs)
                cur[nd.sample] = v
                if nd.children:
                    tot += sum((walk(ch, cur) for ch in nd.children), mp.mpf(0))
                else:
                    n_joint += 1
                    vv = [cur[k] for k in range(S)]
                    j = _pedigree_prior(vv, names, het)
                    for k in range(S):  # sample-index order (generic.rs:511-551)
                        j *= lh(k, vv[k])
                    tot += j
            return tot
        dens = [sum((walk(r, {}) for r in trees[e]), mp.mpf(0)) for e in events]
        total = sum(dens, mp.mpf(0))
        for k, d in enumerate(dens):
            got = float(o.log_posteriors[i, k])
            if d == 0:
                assert got == -math.inf, (i, events[k], got)
                continue
            delta = abs(float(mp.mpf(got) - mp.log(d / total)))
            worst = max(worst, delta)
            assert delta <= 1e-9, (i, events[k], got, delta)
        assert n_joint == int(o.n_base_events[i])
        assert 2 * max(range(len(dens)), key=lambda k: dens[k]) == int(o.best_event[i])
    assert worst > 0.0


# ---------------------------------------------------------------------------------------------- artifact events (a3)
# The artifact configs of an SNV record with all six checks on (bias/mod.rs:128-216: exactly one bias active; the
# homopolymer error only exists for homopolymer indels), their per-read terms (strand_bias.rs:28-53,
# read_orientation_bias.rs:17-35, read_position_bias.rs:17-45, softclip_bias.rs:14-27, alt_locus_bias.rs:62-112) and the
# three selection rules (bias/mod.rs:37-104 + the per-bias is_informative) restated from those files.
ARTIFACT_CONFIGS = ("sb_fwd", "sb_rev", "rob_f1r2", "rob_f2r1", "rpb", "scb", "alb")


class Reads:
    def __init__(self, b, lo, hi, filter_nonstandard=True):
        c = b.columns
        self.rows = []

        def opt(col, r):  # Option<LogProb> columns: NaN = None
            if col is None or np.isnan(col[r]):
                return None
            return _e(col[r])
        for r in range(lo, hi):
            f = int(b.read_flags[r])
            if filter_nonstandard and ((f >> abi.RF_ORIENT_SHIFT) & 15) not in (abi.ORIENT_F1R2, abi.ORIENT_F2R1, abi.ORIENT_NONE):
                continue  # Pileup::remove_nonstandard_alignments (pileup.rs:26-43; SNV / MNV records, calling.rs:596-603)
            hlen = ((f >> abi.RF_HOMOPOLYMER_LEN_SHIFT) & 0xff) if f & abi.RF_HAS_HOMOPOLYMER_LEN else 0
            self.rows.append(dict(
                hlen=hlen - 256 if hlen > 127 else hlen, hart=opt(b.prob_homopolymer_artifact, r),
                hvar=opt(b.prob_homopolymer_variant, r),
                strand=(f >> abi.RF_STRAND_SHIFT) & 3, orient=(f >> abi.RF_ORIENT_SHIFT) & 15,
                major=bool(f & abi.RF_READPOS_MAJOR), softclip=bool(f & abi.RF_SOFTCLIPPED), maxq=bool(f & abi.RF_MAX_MAPQ),
                altlocus=(f >> abi.RF_ALTLOCUS_SHIFT) & 3,
                pm=float(c["prob_mapping"][r]), pa=float(c["prob_alt"][r]), pr=float(c["prob_ref"][r]),
                e_pm=_e(c["prob_mapping"][r]), e_pa=_e(c["prob_alt"][r]), e_pr=_e(c["prob_ref"][r]),
                e_miss=_e(c["prob_missed_allele"][r]), e_pdo=_e(c["prob_double_overlap"][r]), e_phb=_e(c["prob_hit_base"][r]),
                psa=float(c["prob_sample_alt"][r])))
        for d in self.rows:
            bf_ref, bf_alt = math.exp(d["pr"] - d["pa"]), math.exp(d["pa"] - d["pr"])
            d["strong_ref"], d["strong_alt"] = bf_ref > 20.0, bf_alt > 20.0      # >= KassRaftery::Strong
            d["pos_ref"] = bf_ref > 3.0                                          # >= KassRaftery::Positive
            d["ref_support"] = d["pr"] > d["pa"]
            d["unique"] = d["pm"] >= math.log(0.95)


def _adjust_singleton_evidence(piles):
    """read_observation.rs:548-562: exactly one read over all pileups favours the alt allele -> its prob_alt() and
    prob_ref() become ln 0.5 (the likelihood reads those; the Kass-Raftery predicates keep reading the raw fields)."""
    alts = [d for p in piles for d in p.rows if d["pa"] > d["pr"]]
    if len(alts) == 1:
        alts[0]["e_pa"] = alts[0]["e_pr"] = HALF
    return len(alts) == 1


def _simpson(density, lo, hi, n):
    """rust-bio LogProb::ln_simpsons_integrate_exp in linear space: n odd points, weights 1 4 2 4 ... 4 1."""
    step = (hi - lo) / (n - 1)
    xs = [lo + step * k for k in range(n)]  # itertools-num linspace
    tot = density(xs[0]) + density(xs[-1])
    for k in range(1, n - 1):
        tot += (4 if k % 2 else 2) * density(xs[k])
    return tot * (mp.mpf(hi) - mp.mpf(lo)) / (n - 1) / 3


def _bias_alt(cfg, d, fr, has_alt_loci=False):
    """(strand, orientation, position, softclip, alt locus) factors of Artifacts::prob_alt for config `cfg` (None = none)."""
    if cfg in ("sb_fwd", "sb_rev"):
        want = 0 if cfg == "sb_fwd" else 1
        sb = mp.mpf(1) if d["strand"] == 3 else (mp.mpf(1) if d["strand"] == want else mp.mpf(0))
    else:
        sb = {0: fr * (1 - d["e_pdo"]), 1: (1 - fr) * (1 - d["e_pdo"]), 2: d["e_pdo"], 3: mp.mpf(1)}[d["strand"]]
    if cfg in ("rob_f1r2", "rob_f2r1"):
        want = abi.ORIENT_F1R2 if cfg == "rob_f1r2" else abi.ORIENT_F2R1
        other = abi.ORIENT_F2R1 if cfg == "rob_f1r2" else abi.ORIENT_F1R2
        rob = mp.mpf(1) if d["orient"] == want else (mp.mpf(0) if d["orient"] == other else HALF)
    else:
        rob = HALF
    rpb_any = d["e_phb"] if d["major"] else 1 - d["e_phb"]
    rpb = (mp.mpf(1) if d["major"] else mp.mpf(0)) if cfg == "rpb" else rpb_any
    scb = (mp.mpf(1) if d["softclip"] else mp.mpf(0)) if cfg == "scb" else mp.mpf(1)
    d["_he"] = (d["hart"] if cfg == "he" else d["hvar"])  # homopolymer_error.rs:22-40: the same term for alt and ref
    if d["_he"] is None:
        d["_he"] = mp.mpf(1)
    if cfg != "alb":
        alb = HALF
    elif has_alt_loci:
        alb = mp.mpf(1) if d["altlocus"] == abi.ALTLOCUS_MAJOR else mp.mpf(0)
    else:
        alb = mp.mpf(0) if d["maxq"] else mp.mpf(1)
    return sb, rob, rpb, scb, alb, rpb_any


class ConfigPileup(Pileup):
    """c0 + c1 x per read under one artifact config (prob_sample_alt = 0 for every read)."""

    def __init__(self, reads, cfg, fr, has_alt_loci=False):
        self.n = len(reads.rows)
        self.c0, self.c1 = [], []
        for d in reads.rows:
            assert d["psa"] == 0.0
            sb, rob, rpb, scb, alb, rpb_any = _bias_alt(cfg, d, fr, has_alt_loci)
            b_alt = sb * rob * rpb * scb * d["_he"] * alb
            # prob_ref = prob_any for every bias, except the homopolymer error (= its prob_alt) and the alt locus bias with
            # real alt loci (alt_locus_bias.rs:85-104)
            alb_ref = (mp.mpf(0) if d["altlocus"] == abi.ALTLOCUS_MAJOR else mp.mpf(1)) if cfg == "alb" and has_alt_loci else HALF
            b_ref = HALF * HALF * rpb_any * 1 * d["_he"] * alb_ref
            b_any = HALF * HALF * rpb_any * 1 * HALF
            a_term, r_term = d["e_pm"] * b_alt * d["e_pa"], d["e_pm"] * b_ref * d["e_pr"]
            self.c0.append(r_term + (1 - d["e_pm"]) * d["e_miss"] * b_any)
            self.c1.append(a_term - r_term)
        self.clear_ref = self.n > 10 and all(d["pos_ref"] for d in reads.rows)


def _surviving_configs(piles, fr_estimated, has_alt_loci=False, configs=ARTIFACT_CONFIGS):
    """is_possible && is_informative && is_likely of the artifact configs over the pileups (bias/mod.rs:37-104)."""
    every = [d for p in piles for d in p.rows]

    def evidence(cfg, d):  # prob_alt of the active bias != ln 0
        if cfg in ("sb_fwd", "sb_rev"):
            return d["strand"] == 3 or d["strand"] == (0 if cfg == "sb_fwd" else 1)
        if cfg in ("rob_f1r2", "rob_f2r1"):
            return d["orient"] != (abi.ORIENT_F2R1 if cfg == "rob_f1r2" else abi.ORIENT_F1R2)
        if cfg == "rpb":
            return d["major"]
        if cfg == "scb":
            return d["softclip"]
        return d["altlocus"] == abi.ALTLOCUS_MAJOR if has_alt_loci else not d["maxq"]

    def informative(cfg):
        if cfg in ("sb_fwd", "sb_rev"):
            return fr_estimated
        if cfg in ("rob_f1r2", "rob_f2r1"):
            std = (abi.ORIENT_F1R2, abi.ORIENT_F2R1)
            n_uncertain = sum(d["orient"] not in std for d in every)
            strong = [d for d in every if d["strong_ref"] and d["orient"] in std]
            uniform = len(strong) > 2 and 0.3 <= sum(d["orient"] == abi.ORIENT_F1R2 for d in strong) / len(strong) <= 0.7
            return n_uncertain < len(every) / 2.0 and uniform
        if cfg == "rpb":
            for p in piles:
                exp_all = sum(math.exp(d["pm"]) for d in p.rows if d["strong_ref"])
                if exp_all > 10.0:
                    exp_major = sum(math.exp(d["pm"]) for d in p.rows if d["strong_ref"] and d["major"])
                    exp_rate = sum(math.exp(d["pm"]) * float(d["e_phb"]) for d in p.rows if d["strong_ref"])
                    if exp_major > 0.0 and abs(exp_major / exp_all - exp_rate) < 0.05:
                        return True
            return False
        if cfg == "scb":
            return any(d["softclip"] for d in every)
        n_alt = sum(d["strong_alt"] for d in every)
        nm_alt = sum(d["strong_alt"] and not d["maxq"] for d in every)
        n_ref = sum(d["strong_ref"] for d in every)
        nm_ref = sum(d["strong_ref"] and not d["maxq"] for d in every)
        enough_alt = n_alt > 0 and nm_alt > n_alt * 0.1 and (n_alt - nm_alt) < 10
        return enough_alt and (has_alt_loci or (n_ref > 0 and nm_ref < n_ref * 0.9))

    def likely(cfg):
        for p in piles:
            strong = [d for d in p.rows if d["unique"] and d["strong_alt"]]
            if len(strong) >= 10:
                if sum(evidence(cfg, d) for d in strong) / len(strong) >= 0.66666:
                    return True
            elif all(d["ref_support"] for d in p.rows):
                continue
            elif not p.rows:
                continue
            else:
                return True
        return False
    def homopolymer_ok():  # HomopolymerError::is_informative = is_possible = is_likely (homopolymer_error.rs:46-75)
        return all(not any(d["strong_alt"] for d in p.rows)
                   or (any(d["hlen"] > 0 for d in p.rows) and any(d["hlen"] < 0 for d in p.rows)) for p in piles)
    out = []
    for c in configs:
        if c == "he":
            if homopolymer_ok():
                out.append(c)
        elif any(evidence(c, d) for d in every) and informative(c) and likely(c):
            out.append(c)
    return out


def _forward_rate_opt(every):
    strong_all = sum(math.exp(d["pm"]) for d in every if d["strong_ref"] and d["strand"] != 2)
    strong_fwd = sum(math.exp(d["pm"]) for d in every if d["strong_ref"] and d["strand"] == 0)
    if strong_all > 2.0:
        frac = strong_fwd / strong_all
        if strong_all > 100.0 and 0.0 < frac < 1.0:
            return mp.mpf(frac)
        if 0.4 <= frac <= 0.6:
            return HALF
    return None


class ConfigLocus(Locus):
    def __init__(self, normal, tumor):
        self.normal, self.tumor = normal, tumor
        self.res = [0.1, 0.01]
        self.purity = mp.mpf(1) - mp.mpf(0.25)
        self.base, self.n_joint, self._ln = {}, 0, {}
        self.base_full, self._disc = {}, {}


@pytest.mark.parametrize("source", ["oracle", "pipeline"])
def test_tumor_normal_artifact_events_against_the_high_precision_restatement(source):
    sc, b = synth.tumor_normal(120, seed=22)
    flat = sc.flatten()
    diag, o = _results(source, flat, b, 0)
    trees = dict(sc.event_trees())
    names = list(flat.event_names)
    E = len(names)
    ok = np.isfinite(diag.log_posteriors[:, -1]) & ((diag.status & np.uint32(0xffffffff ^ abi.ST_IS_ARTIFACT)) == 0) & ~diag.knife_edge()
    order = np.argsort(-diag.log_posteriors[:, -1], kind="stable")       # the strongest artifact posteriors first
    loci = [int(i) for i in order if ok[i]][:3] + [int(i) for i in np.nonzero(ok)[0][:3]]
    loci = list(dict.fromkeys(loci))
    assert len(loci) >= 4
    worst, n_cfg_seen, n_artifact_calls = 0.0, set(), 0
    for i in loci:
        offs = [int(b.read_offsets[i * 2 + k]) for k in range(3)]
        piles = [Reads(b, offs[0], offs[1]), Reads(b, offs[1], offs[2])]
        fr_opt = _forward_rate_opt([d for p in piles for d in p.rows])
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        assert surviving, i
        n_cfg_seen.update(surviving)
        n_joint = 0
        dens = {}
        for cfg in [None] + surviving:
            L = ConfigLocus(ConfigPileup(piles[0], cfg, fr), ConfigPileup(piles[1], cfg, fr))
            for name in names:
                if cfg is not None and name == "absent":
                    continue  # the absent event has no artifact twin (calling.rs:654-659)
                dens[(cfg, name)] = sum((L.node(r, None) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in names]
        twin = [sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving), mp.mpf(0)) if n != "absent" else mp.mpf(0)
                for n in names]
        total = sum(plain, mp.mpf(0)) + sum(twin, mp.mpf(0))
        want = [p / total for p in plain] + [sum(twin, mp.mpf(0)) / total]
        for k in range(E + 1):
            got = float(o.log_posteriors[i, k])
            if want[k] == 0:
                assert got == -math.inf, (i, k, got)
                continue
            delta = abs(float(mp.mpf(got) - mp.log(want[k])))
            worst = max(worst, delta)
            assert delta <= 1e-9, (i, k, surviving, got, float(mp.log(want[k])), delta)
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))
        # the call is an artifact iff the artifact event beats every other event (calling.rs:800-818); its samples then
        # report an allele frequency of 0 and the bias (calling.rs:866-876)
        is_artifact = all(want[E] > w for w in want[:E])
        assert is_artifact == bool(int(o.status[i]) & abi.ST_IS_ARTIFACT), (i, [float(w) for w in want])
        if is_artifact:
            n_artifact_calls += 1
            assert not o.map_vaf[i].any() and int(o.map_config[i]) != 0
    assert worst > 0.0 and len(n_cfg_seen) >= 2 and n_artifact_calls >= 1


def _artifact_posteriors(b, i, names, trees):
    """The E + 1 posteriors of tumor-normal locus i (plain events + the artifact event) and the number of joint evaluations."""
    offs = [int(b.read_offsets[i * 2 + k]) for k in range(3)]
    lf = int(b.locus_flags[i])
    configs = [c for c, bit in (("sb_fwd", abi.LF_CHECK_SB), ("sb_rev", abi.LF_CHECK_SB), ("rob_f1r2", abi.LF_CHECK_ROB),
                                ("rob_f2r1", abi.LF_CHECK_ROB), ("rpb", abi.LF_CHECK_RPB), ("scb", abi.LF_CHECK_SCB),
                                ("he", abi.LF_CHECK_HE), ("alb", abi.LF_CHECK_ALB)) if lf & bit]  # calling.rs:559-566
    flt = bool(lf & abi.LF_FILTER_NONSTANDARD)
    piles = [Reads(b, offs[0], offs[1], flt), Reads(b, offs[1], offs[2], flt)]
    _adjust_singleton_evidence(piles)
    every = [d for p in piles for d in p.rows]
    fr_opt = _forward_rate_opt(every)
    fr = fr_opt if fr_opt is not None else HALF
    has_alt_loci = any(d["altlocus"] != abi.ALTLOCUS_NONE for d in every)   # AltLocusBias::learn_parameters
    surviving = _surviving_configs(piles, fr_opt is not None, has_alt_loci, configs)
    n_joint, dens = 0, {}
    for cfg in [None] + surviving:
        L = ConfigLocus(ConfigPileup(piles[0], cfg, fr, has_alt_loci), ConfigPileup(piles[1], cfg, fr, has_alt_loci))
        for name in names:
            if cfg is None or name != "absent":
                dens[(cfg, name)] = sum((L.node(r, None) for r in trees[name]), mp.mpf(0))
        n_joint += L.n_joint
    plain = [HALF * dens[(None, n)] for n in names]
    twin = sum((HALF / len(configs) * dens[(c, n)] for c in surviving for n in names if n != "absent"), mp.mpf(0))
    total = sum(plain, mp.mpf(0)) + twin
    return [p / total for p in plain] + [twin / total], n_joint, surviving


def test_constructed_pileups_for_the_other_artifact_configs():
    """Softclip, read position and alt locus (with real alt loci) configs, and the two-thirds rule of is_likely."""
    from tests.util import batch_from_reads, read
    q = dict(prob_mapping=np.log1p(-1e-6), prob_double_overlap=-np.inf)
    hi, lo = np.log1p(-1e-3), np.log(1e-3 / 3)

    def ref(k, **kw):
        return read(prob_ref=hi, prob_alt=lo, strand=k % 2, orientation=(k // 2) % 2, **{**q, **kw})

    def alt(k, **kw):
        return read(prob_ref=lo, prob_alt=hi, strand=k % 2, orientation=(k // 2) % 2, **{**q, **kw})
    # has_valid_major_rate (read_position_bias.rs:62-122) compares the major FRACTION of the strong reference reads with
    # the SUM of their e^(prob_mapping + prob_hit_base) (upstream does not normalise it): 40 reads x 0.005 = 0.2 = 8 / 40
    phb = dict(prob_hit_base=math.log(0.005))
    loci = [
        # softclip bias: every alt read is softclipped, a few reference reads too
        [[ref(k, softclipped=k < 3) for k in range(30)],
         [ref(k) for k in range(24)] + [alt(k, softclipped=True) for k in range(6)]],
        # read position bias: alt reads at the major position; a fifth of the reference reads too (= prob_hit_base)
        [[ref(k, major=k % 5 == 0, **phb) for k in range(40)],
         [ref(k, major=k % 5 == 0, **phb) for k in range(30)] + [alt(k, major=True, **phb) for k in range(7)]],
        # alt locus bias with alt loci: alt reads point to the major alt locus and have low MAPQs
        [[ref(k) for k in range(30)],
         [ref(k) for k in range(25)] + [alt(k, alt_locus=abi.ALTLOCUS_MAJOR, max_mapq=False,
                                           prob_mapping=np.log1p(-1e-2)) for k in range(6)]],
        # is_likely by the two-thirds rule: 12 strong alt reads, 10 of them on the forward strand; normal all reference
        [[ref(k) for k in range(30)],
         [ref(k) for k in range(28)] + [alt(2 * (k % 2)) for k in range(10)] + [alt(1), alt(3)]],  # (orientations balanced)
        # ... and 12 strong alt reads spread evenly: no config is likely
        [[ref(k) for k in range(30)], [ref(k) for k in range(28)] + [alt(k) for k in range(12)]],
    ]
    b = batch_from_reads(loci)
    sc = Scenario.tumor_normal(0.75)
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    o = oracle.call_batch(flat, b, afd_capacity=0)
    expect = ["scb", "rpb", "alb", "sb_fwd", None]
    for i, must in enumerate(expect):
        assert not o.knife_edge()[i] and (int(o.status[i]) & ~abi.ST_IS_ARTIFACT) == 0, (i, int(o.status[i]))
        want, n_joint, surviving = _artifact_posteriors(b, i, names, trees)
        assert (must in surviving) if must else surviving == [], (i, surviving)
        if must == "sb_fwd":
            assert surviving == ["sb_fwd"]  # (the other configs fail the two-thirds rule)
        for k, w in enumerate(want):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))


def test_small_pileups_singleton_evidence_and_filtered_reads():
    """Simpson fallbacks (fewer than 5 reads: 11 points; fewer than 10: the range bounds themselves are the limits),
    the singleton-evidence adjustment and the orientation filter of SNV records."""
    from tests.util import batch_from_reads, read
    q = dict(prob_mapping=np.log1p(-1e-5), prob_double_overlap=-np.inf)

    def err(k):  # base error rates differ from read to read: no exact ties between abscissae
        return 10.0 ** -(2.0 + 0.13 * (k % 11))

    def ref(k, **kw):
        return read(**{**dict(prob_ref=np.log1p(-err(k)), prob_alt=np.log(err(k) / 3), strand=k % 2,
                              orientation=(k // 2) % 2), **q, **kw})

    def alt(k, **kw):
        return read(**{**dict(prob_ref=np.log(err(k + 5) / 3), prob_alt=np.log1p(-err(k + 5)), strand=k % 2,
                              orientation=(k // 2) % 2), **q, **kw})
    loci = [
        [[ref(0), ref(1), ref(2)], [ref(0), alt(1), alt(2), ref(3)]],                  # 3 and 4 reads: Simpson 11
        [[ref(k) for k in range(7)], [ref(k) for k in range(5)] + [alt(0), alt(3)]],   # 7 reads each: adaptive, raw bounds
        [[ref(k) for k in range(30)], [ref(k) for k in range(29)] + [alt(0)]],         # singleton evidence
        [[ref(k) for k in range(20)] + [alt(0, orientation=abi.ORIENT_F1F2)],          # filtered: leaves no alt read there
         [ref(k) for k in range(20)] + [alt(1), alt(2), ref(0, orientation=abi.ORIENT_R1R2)]],
        [[], [ref(0), alt(1)]],                                                         # an empty pileup
    ]
    b = batch_from_reads(loci)
    sc = Scenario.tumor_normal(0.75)
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    o = oracle.call_batch(flat, b, afd_capacity=0)
    assert int(o.status[2]) & abi.ST_SINGLETON_ADJUSTED and int(o.status[3]) & abi.ST_FILTERED_NONSTANDARD
    assert o.knife_edge().sum() <= 1
    for i in np.nonzero(~o.knife_edge())[0]:
        want, n_joint, surviving = _artifact_posteriors(b, i, names, trees)
        for k, w in enumerate(want):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))


def test_homopolymer_indel_records():
    """Indel records: no read-orientation / read-position / softclip configs and no orientation filter
    (calling.rs:559-566, :596-603); the homopolymer error model (homopolymer_error.rs) for homopolymer indels."""
    from tests.util import batch_from_reads, read
    from varlociraptor_b200 import obs_codec
    q = dict(prob_mapping=np.log1p(-1e-5), prob_double_overlap=-np.inf)

    def err(k):
        return 10.0 ** -(2.0 + 0.13 * (k % 11))

    def ref(k, **kw):
        return read(**{**dict(prob_ref=np.log1p(-err(k)), prob_alt=np.log(err(k) / 3), strand=k % 2,
                              orientation=(k // 2) % 2), **q, **kw})

    def alt(k, **kw):
        return read(**{**dict(prob_ref=np.log(err(k + 5) / 3), prob_alt=np.log1p(-err(k + 5)), strand=k % 2,
                              orientation=(k // 2) % 2), **q, **kw})
    h = lambda k, n: dict(hlen=n, hart=math.log(0.05 + 0.01 * (k % 5)), hvar=math.log(0.6 + 0.02 * (k % 7)))  # noqa: E731
    loci = [
        # homopolymer indel, insertions and deletions among the reads: the homopolymer config is considered
        [[ref(k, **h(k, 0)) for k in range(25)] + [alt(0, **h(0, 1)), ref(2, **h(2, -2))],
         [ref(k, **h(k, 0)) for k in range(20)] + [alt(k, **h(k, 1)) for k in range(6)] + [alt(7, **h(7, -1)), ref(3, **h(3, -1))]],
        # ... insertions only in the sample with strong alt reads: it is not
        [[ref(k, **h(k, 0)) for k in range(25)],
         [ref(k, **h(k, 0)) for k in range(20)] + [alt(k, **h(k, 1)) for k in range(6)]],
        # an ordinary indel with reads in non-standard orientations (kept: the filter is for SNVs and MNVs)
        [[ref(k) for k in range(25)] + [ref(0, orientation=abi.ORIENT_F1F2)],
         [ref(k) for k in range(20)] + [alt(k) for k in range(5)] + [alt(1, orientation=abi.ORIENT_R1R2)]],
    ]
    hom = obs_codec.locus_flags_for("A", "AT", True)
    plain = obs_codec.locus_flags_for("A", "AT", False)
    b = batch_from_reads(loci, locus_flags=[hom, hom, plain])
    sc = Scenario.tumor_normal(0.75)
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    o = oracle.call_batch(flat, b, afd_capacity=0)
    expect = [True, False, None]
    for i in range(len(loci)):
        assert not o.knife_edge()[i]
        assert not int(o.status[i]) & abi.ST_FILTERED_NONSTANDARD
        want, n_joint, surviving = _artifact_posteriors(b, i, names, trees)
        if expect[i] is not None:
            assert ("he" in surviving) == expect[i], (i, surviving)
        assert not {"rob_f1r2", "rob_f2r1", "rpb", "scb"} & set(surviving)
        for k, w in enumerate(want):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))


# ---------------------------------------------------------------------------------------------- priors with somatic rates
import collections  # noqa: E402

_Rg = collections.namedtuple("_Rg", "start end left_exclusive right_exclusive")


def _log2(x):
    return math.log2(x) if x > 0.0 else -math.inf


def _lfc_true(cmp_, v, lfc):  # Log2FoldChangePredicate::is_true (log2_fold_change.rs:40-52)
    if lfc != lfc:
        raise AssertionError("NaN log2 fold change")
    return {abi.CMP_EQ: abs(lfc - v) <= 2.220446049250313e-16 * max(abs(lfc), abs(v), 1.0) if False else lfc == v,
            abi.CMP_GT: lfc > v, abi.CMP_GE: lfc >= v, abi.CMP_LT: lfc < v, abi.CMP_LE: lfc <= v, abi.CMP_NE: lfc != v}[cmp_]


def _rg_empty(r):
    return r.start == r.end and (r.left_exclusive or r.right_exclusive)


def _rg_contains(r, v):
    return (r.start < v if r.left_exclusive else r.start <= v) and (r.end > v if r.right_exclusive else r.end >= v)


def _rg_intersect(a, o):  # VAFRange::intersect (formula.rs:1230-1262) with overlap() == None -> empty (:1140-1160)
    if a != o and ((a.end < o.start or a.start > o.end)
                   or (a.end <= o.start and (a.right_exclusive or o.left_exclusive))
                   or (a.start >= o.end and (a.left_exclusive or o.right_exclusive))):
        return _Rg(0.0, 0.0, True, True)
    lex = a.left_exclusive if a.start > o.start else (o.left_exclusive if a.start < o.start else a.left_exclusive or o.left_exclusive)
    rex = a.right_exclusive if a.end < o.end else (o.right_exclusive if a.end > o.end else a.right_exclusive or o.right_exclusive)
    return _Rg(max(a.start, o.start), min(a.end, o.end), lex, rex)


def _infer_vaf_bounds(cmp_, value, vaf):  # Log2FoldChangePredicate::infer_vaf_bounds (log2_fold_change.rs:54-93)
    proj = vaf / 2.0 ** value
    if proj < 0.0 or proj > 1.0:
        return _Rg(0.0, 0.0, True, True)
    return {abi.CMP_EQ: _Rg(proj, proj, False, False), abi.CMP_GT: _Rg(0.0, proj, False, True),
            abi.CMP_GE: _Rg(0.0, proj, False, False), abi.CMP_LT: _Rg(proj, 1.0, True, False),
            abi.CMP_LE: _Rg(proj, 1.0, False, False), abi.CMP_NE: _Rg(0.0, 1.0, False, False)}[cmp_]


_INVERT = {abi.CMP_EQ: (abi.CMP_EQ, 1), abi.CMP_GT: (abi.CMP_LE, -1), abi.CMP_GE: (abi.CMP_LT, -1),
           abi.CMP_LT: (abi.CMP_GE, -1), abi.CMP_LE: (abi.CMP_GT, -1), abi.CMP_NE: (abi.CMP_NE, 1)}  # invert() (:95-124)


def _lfc_bounds(vafs, sample):
    acc = None
    for a, b, cmp_, value in vafs.get("lfcs", ()):
        bounds = None
        if a == sample and b in vafs:
            c2, sign = _INVERT[cmp_]
            bounds = _infer_vaf_bounds(c2, sign * value, vafs[b])
        elif b == sample and a in vafs:
            bounds = _infer_vaf_bounds(cmp_, value, vafs[a])
        if bounds is not None:
            acc = bounds if acc is None else _rg_intersect(acc, bounds)
    return acc


class TreeLocus:
    """density() over arbitrary Set / Range trees of S uncontaminated samples with a prior function of the VAF vector."""

    def __init__(self, piles, res, prior, snv=None):
        self.piles, self.res, self.prior, self.snv = piles, res, prior, snv
        self.n_joint, self._lh = 0, {}

    def joint(self, vafs):
        self.n_joint += 1
        for a, bb, cmp_, value in vafs.get("lfcs", ()):  # GenericLikelihood::compute step 1 (generic.rs:503-509)
            va, vb = vafs[a], vafs[bb]
            lfc = 0.0 if va == 0.0 and vb == 0.0 else _log2(va) - _log2(vb)    # log2_fold_change.rs:17-27, in f64
            if not _lfc_true(cmp_, value, lfc):
                return mp.mpf(0)
        j = self.prior([vafs[k] for k in range(len(self.piles))])
        for k in range(len(self.piles)):  # sample-index order (generic.rs:511-551)
            if (k, vafs[k]) not in self._lh:
                self._lh[(k, vafs[k])] = self.piles[k].value(vafs[k])
            j *= self._lh[(k, vafs[k])]
        return j

    def node(self, nd, vafs):
        if nd.kind == abi.NODE_VARIANT:  # generic.rs:398-420: gate on the record's SNV bases (IUPAC masks), then skip the node
            if self.snv is not None:
                mask = {"A": 1, "C": 2, "G": 4, "T": 8}
                contains = bool(nd.refmask & mask[self.snv[0]]) and bool(nd.altmask & mask[self.snv[1]])
                if nd.positive != contains:
                    return mp.mpf(0)
            elif nd.positive:
                return mp.mpf(0)
            if nd.children:
                return sum((self.node(ch, vafs) for ch in nd.children), mp.mpf(0))
            return self.joint(vafs)
        if nd.kind == abi.NODE_LFC:  # generic.rs:233-244: remember the predicate, go on below
            cur = dict(vafs)
            cur["lfcs"] = tuple(vafs.get("lfcs", ())) + ((nd.sample, nd.sample_b, nd.cmp, nd.lfc_value),)
            if nd.children:
                return sum((self.node(ch, cur) for ch in nd.children), mp.mpf(0))
            return self.joint(cur)
        pile, res = self.piles[nd.sample], self.res[nd.sample]

        def below(v):
            cur = dict(vafs)
            cur[nd.sample] = v
            if nd.children:  # one child: recurse; several: ln_sum_exp over them (generic.rs:199-226)
                return sum((self.node(ch, cur) for ch in nd.children), mp.mpf(0))
            return self.joint(cur)
        bounds = _lfc_bounds(vafs, nd.sample)           # generic.rs:148-174
        if bounds is not None and _rg_empty(bounds):
            return mp.mpf(0)
        if nd.kind == 0:
            vs = sorted(nd.vafs)
            if pile.clear_ref and all(v > 0.0 for v in vs):
                return mp.mpf(0)
            if bounds is not None:
                vs = [v for v in vs if _rg_contains(bounds, v)]
            return sum((below(v) for v in vs), mp.mpf(0))
        rg = nd.vafs
        if bounds is not None:
            rg = _rg_intersect(_Rg(rg.start, rg.end, rg.left_exclusive, rg.right_exclusive), bounds)
        if _rg_empty(rg):
            return mp.mpf(0)
        if pile.clear_ref and rg.start > 0.0:
            return mp.mpf(0)
        if rg.start == rg.end:  # singleton: a point event (generic.rs:349-355)
            return below(rg.start)
        mn, mx = _observable_min(rg, pile.n), _observable_max(rg, pile.n)
        assert mn <= mx
        if (mx - mn) < res:
            return _simpson(below, mn, mx, 3)
        if pile.n < 5:
            return _simpson(below, mn, mx, 11)
        return _integrate(below, mn, mx, res)


def _tumor_normal_prior(full_prior):
    """Prior::compute (prior.rs:718-761) for the reference's tests/resources/prior/scenarios/tumor-normal scenario:
    normal = germline only (heterozygosity 0.001, ploidy 2), tumor = clonal from normal (somatic: false) with a somatic
    effective mutation rate of 1e-6. calc_prob (:298-438): the normal's allele frequency must be a germline one
    (:237-241); population term over the normal (:554-582); the tumor's germline equals the normal's (:458-470) and what
    is left of its allele frequency is somatic: rate if != 0 else 1 - rate (:440-456). Samples sorted: normal, tumor."""
    het, rate = mp.mpf("0.001"), mp.mpf("1e-6")

    def full(vafs):
        vn, vt = vafs
        n_alt = 2 * vn
        if n_alt != round(n_alt):
            return mp.mpf(0)
        m = int(round(n_alt))
        pop = het / m if m > 0 else 1 - (het / 1 + het / 2)
        return pop * ((1 - rate) if vt - vn == 0.0 else rate)

    def absent_only(vafs):  # the default mode (:737-754)
        if all(v == 0.0 for v in vafs):
            return full(vafs)
        return mp.mpf(0) if full(vafs) == 0 else 1 - full([0.0, 0.0])
    return full if full_prior else absent_only


@pytest.mark.parametrize("full_prior", [False, True])
def test_somatic_rate_and_clonal_inheritance_prior(golden_dir, full_prior):
    import json
    import os
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"]["tumor-normal"]
    sc = Scenario.from_yaml(text, full_prior=full_prior).for_contig("all")
    assert list(sc.sample_names) == ["normal", "tumor"]
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    b = synth.tumor_normal(60, seed=41, depth=30)[1]
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    prior = _tumor_normal_prior(full_prior)
    checked, worst = 0, 0.0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT:
            continue
        offs = [int(b.read_offsets[i * 2 + k]) for k in range(3)]
        piles = [Reads(b, offs[0], offs[1]), Reads(b, offs[1], offs[2])]
        every = [d for p in piles for d in p.rows]
        fr_opt = _forward_rate_opt(every)
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        if len(surviving) > 2 and checked >= 4:
            continue  # (every surviving config is another pass over all trees: keep the test short)
        n_joint, dens = 0, {}
        for cfg in [None] + surviving:
            L = TreeLocus([ConfigPileup(p, cfg, fr) for p in piles], [0.01, 0.01], prior)
            for name in names:
                if cfg is None or name != "absent":
                    dens[(cfg, name)] = sum((L.node(r, {}) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in names]
        twin = sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving for n in names if n != "absent"), mp.mpf(0))
        total = sum(plain, mp.mpf(0)) + twin
        for k, w in enumerate([p / total for p in plain] + [twin / total]):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                delta = abs(float(mp.mpf(got) - mp.log(w)))
                worst = max(worst, delta)
                assert delta <= 1e-9, (i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))
        checked += 1
        if checked >= 8:
            break
    assert checked >= 6 and worst > 0.0


@pytest.mark.parametrize("full_prior", [False, True])
def test_population_prior_of_two_founders(golden_dir, full_prior):
    """tests/resources/prior/scenarios/population: two unrelated diploid samples, heterozygosity 0.001: the population
    term counts the alt alleles of BOTH (prior.rs:554-582): het / m for m > 0, 1 - sum_{m=1..4} het / m for none."""
    import json
    import os
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"]["population"]
    sc = Scenario.from_yaml(text, full_prior=full_prior).for_contig("all")
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    het = mp.mpf("0.001")
    p_absent = 1 - sum(het / m for m in range(1, 5))

    def prior(vafs):
        m = sum(int(round(2 * v)) for v in vafs)
        if m == 0:
            return p_absent
        return het / m if full_prior else 1 - p_absent
    b = synth.tumor_normal(40, seed=43, depth=30)[1]
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    checked = 0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT:
            continue
        offs = [int(b.read_offsets[i * 2 + k]) for k in range(3)]
        piles = [Reads(b, offs[0], offs[1]), Reads(b, offs[1], offs[2])]
        every = [d for p in piles for d in p.rows]
        fr_opt = _forward_rate_opt(every)
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        n_joint, dens = 0, {}
        for cfg in [None] + surviving:
            L = TreeLocus([ConfigPileup(p, cfg, fr) for p in piles], [0.01, 0.01], prior)
            for name in names:
                if cfg is None or name != "absent":
                    dens[(cfg, name)] = sum((L.node(r, {}) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in names]
        twin = sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving for n in names if n != "absent"), mp.mpf(0))
        total = sum(plain, mp.mpf(0)) + twin
        for k, w in enumerate([p / total for p in plain] + [twin / total]):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i])
        checked += 1
    assert checked >= 20


def test_variant_nodes_gate_on_the_snv_bases():
    """`C>T & s:]0,1]` / `!C>T & s:]0,1]` (generic.rs:398-420, formula.rs:23-43) on single-sample SNV loci."""
    from varlociraptor_b200 import LocusBatch
    sc = Scenario.from_yaml("""
samples:
  s:
    universe: "[0.0,1.0]"
events:
  ct: "C>T & s:]0.0,1.0]"
  other: "!C>T & s:]0.0,1.0]"
""")
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    _, b2 = synth.tumor_normal(24, seed=22, depth=25)
    b = LocusBatch(1, b2.read_offsets[::2].copy(), dict(b2.columns), b2.read_flags, b2.locus_flags)  # one sample of 50 reads
    o = oracle.call_batch(flat, b, afd_capacity=0)
    res = float(flat.c.samples[0].resolution)
    n_ct = checked = 0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT:
            continue
        lf = int(b.locus_flags[i])
        snv = (chr((lf >> abi.LF_REFBASE_SHIFT) & 0xff), chr((lf >> abi.LF_ALTBASE_SHIFT) & 0xff)) if lf & abi.LF_HAS_SNV else None
        piles = [Reads(b, int(b.read_offsets[i]), int(b.read_offsets[i + 1]))]
        fr_opt = _forward_rate_opt(piles[0].rows)
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        n_joint, dens = 0, {}
        for cfg in [None] + surviving:
            L = TreeLocus([ConfigPileup(piles[0], cfg, fr)], [res], lambda v: mp.mpf(1), snv)
            for name in names:
                if cfg is None or name != "absent":
                    dens[(cfg, name)] = sum((L.node(r, {}) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in names]
        twin = sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving for n in names if n != "absent"), mp.mpf(0))
        total = sum(plain, mp.mpf(0)) + twin
        for k, w in enumerate([p / total for p in plain] + [twin / total]):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, k, snv, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, n_joint, int(o.n_base_events[i]))
        is_ct = snv == ("C", "T")
        assert (dens[(None, "ct")] > 0) == is_ct or piles[0].clear_ref
        n_ct += is_ct
        checked += 1
    assert checked >= 12 and 0 < n_ct < checked


def test_log2_fold_change_nodes():
    """l2fc(a, b) >= 1 / < 1 over two full ranges (generic.rs:148-174, :233-244, :503-509; log2_fold_change.rs): the
    limits inferred from the predicate, their intersection with the node's range, the predicate check of every joint
    evaluation. The predicate is evaluated in f64 with the platform's log2 like the reference does (abscissae ON the
    threshold hang on its last bit: the same libm here, in the oracle and in the emulation)."""
    from tests.test_emu_parity import LFC_YAML
    sc = Scenario.from_yaml(LFC_YAML)
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    _, b = synth.tumor_normal(30, seed=21, depth=30)
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    res = [float(flat.c.samples[k].resolution) for k in range(2)]
    checked = 0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT:
            continue
        offs = [int(b.read_offsets[i * 2 + k]) for k in range(3)]
        piles = [Reads(b, offs[0], offs[1]), Reads(b, offs[1], offs[2])]
        every = [d for p in piles for d in p.rows]
        fr_opt = _forward_rate_opt(every)
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        if len(surviving) > 2:
            continue
        n_joint, dens = 0, {}
        for cfg in [None] + surviving:
            L = TreeLocus([ConfigPileup(p, cfg, fr) for p in piles], res, lambda v: mp.mpf(1))
            for name in names:
                if cfg is None or name != "absent":
                    dens[(cfg, name)] = sum((L.node(r, {}) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in names]
        twin = sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving for n in names if n != "absent"), mp.mpf(0))
        total = sum(plain, mp.mpf(0)) + twin
        for k, w in enumerate([p / total for p in plain] + [twin / total]):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, names, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))
        checked += 1
        if checked >= 8:
            break
    assert checked >= 5


def _hypergeom_pmf(N, K, n, k):
    """statrs Hypergeometric::new(N, K, n).pmf(k): k successes in n draws without replacement from N with K successes."""
    if k > K or n - k > N - K or k < 0 or n - k < 0:
        return mp.mpf(0)
    return mp.binomial(K, k) * mp.binomial(N - K, n - k) / mp.binomial(N, n)


def _mendelian(pl_parents, pl_child, alt_parents, alt_child, mu):
    """prob_mendelian_alt_counts (prior.rs:600-678): meiotic splits of both parents, de novo mutations for the rest."""
    def cases(p):
        return [p // 2] if p % 2 == 0 else [p // 2, p // 2 + 1]
    total, valid = mp.mpf(0), False
    for p1 in cases(pl_parents[0]):
        for p2 in cases(pl_parents[1]):
            if p1 + p2 != pl_child:
                continue
            valid = True
            for a1 in range(0, min(alt_parents[0], p1) + 1):
                for a2 in range(0, min(alt_parents[1], p2) + 1):
                    if a1 + a2 <= alt_child:
                        total += (_hypergeom_pmf(pl_parents[0], alt_parents[0], p1, a1)
                                  * _hypergeom_pmf(pl_parents[1], alt_parents[1], p2, a2) * mu ** (alt_child - a1 - a2))
    assert valid
    return total


@pytest.mark.parametrize("full_prior", [False, True])
@pytest.mark.parametrize("contig", ["all", "X", "Y"])
def test_pedigree_with_sex_chromosome_ploidies(golden_dir, contig, full_prior):
    """tests/resources/prior/scenarios/pedigree (mother, father, son, daughter) on an autosome, X and Y: per-sample
    ploidies 2/2/2/2, 2/1/1/2, 0/1/1/0; population term over the founders, Mendelian terms with odd and zero ploidies."""
    import json
    import os
    from tests.util import four_sample_batch
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"]["pedigree"]
    sc = Scenario.from_yaml(text, full_prior=full_prior).for_contig(contig)
    flat = sc.flatten()
    names, events, trees = list(sc.sample_names), list(flat.event_names), dict(sc.event_trees())
    S = len(names)
    ploidy = {n: sc.ploidy(n) for n in names}
    het, mu = mp.mpf("0.001"), mp.mpf("1e-3")

    def full(vafs):
        v = dict(zip(names, vafs))
        alts = {}
        for n in names:
            if ploidy[n] == 0 and v[n] != 0.0:
                return mp.mpf(0)
            k = ploidy[n] * v[n]
            if k != round(k):
                return mp.mpf(0)
            alts[n] = int(round(k))
        m = alts["mother"] + alts["father"]
        p = het / m if m > 0 else 1 - sum(het / k for k in range(1, ploidy["mother"] + ploidy["father"] + 1))
        for kid in ("child", "sibling"):
            p *= _mendelian((ploidy["mother"], ploidy["father"]), ploidy[kid], (alts["mother"], alts["father"]), alts[kid], mu)
        return p

    def prior(vafs):
        if full_prior or all(x == 0.0 for x in vafs):
            return full(vafs)
        return mp.mpf(0) if full(vafs) == 0 else 1 - full([0.0] * S)
    b = four_sample_batch(24, seed=51, depth=16)
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    checked = 0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT:
            continue
        offs = [int(b.read_offsets[i * S + k]) for k in range(S + 1)]
        piles = [Reads(b, offs[k], offs[k + 1]) for k in range(S)]
        fr_opt = _forward_rate_opt([d for p in piles for d in p.rows])
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        n_joint, dens = 0, {}
        for cfg in [None] + surviving:
            L = TreeLocus([ConfigPileup(p, cfg, fr) for p in piles], [0.01] * S, prior)
            for name in events:
                if cfg is None or name != "absent":
                    dens[(cfg, name)] = sum((L.node(r, {}) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in events]
        twin = sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving for n in events if n != "absent"), mp.mpf(0))
        total = sum(plain, mp.mpf(0)) + twin
        for k, w in enumerate([p / total for p in plain] + [twin / total]):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (contig, i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i]), (i, surviving, n_joint, int(o.n_base_events[i]))
        checked += 1
        if checked >= 6:
            break
    assert checked >= 4


@pytest.mark.parametrize("full_prior", [False, True])
def test_tumor_relapse_scenario_with_a_uniform_prior_sample(golden_dir, full_prior):
    """tests/resources/prior/scenarios/tumor-relapse: normal (germline), tumor (clonal from normal, somatic rate 1e-6) and
    a relapse sample that only declares a universe: it gets a flat prior inside it and its germline counts as 0
    (prior.rs:398-406), the others keep theirs - two full ranges nested around a three-valued set."""
    import json
    import os
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"]["tumor-relapse"]
    sc = Scenario.from_yaml(text, full_prior=full_prior).for_contig("all")
    assert list(sc.sample_names) == ["normal", "relapse", "tumor"]
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    het, rate = mp.mpf("0.001"), mp.mpf("1e-6")

    def full(vafs):
        vn, vr, vt = vafs
        if not 0.0 <= vr <= 1.0:
            return mp.mpf(0)
        k = 2 * vn
        if k != round(k):
            return mp.mpf(0)
        m = int(round(k))
        pop = het / m if m > 0 else 1 - (het / 1 + het / 2)
        return pop * ((1 - rate) if vt - vn == 0.0 else rate)

    def prior(vafs):
        if full_prior or all(v == 0.0 for v in vafs):
            return full(vafs)
        return mp.mpf(0) if full(vafs) == 0 else 1 - full([0.0, 0.0, 0.0])
    b = synth.pedigree(30, seed=61, depth=30)[1]
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    checked = 0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT or (int(b.locus_flags[i]) >> abi.LF_VARTYPE_SHIFT) & 3:
            continue  # (SNV loci: the variant-type fractions are covered by the pedigree test)
        offs = [int(b.read_offsets[i * 3 + k]) for k in range(4)]
        piles = [Reads(b, offs[k], offs[k + 1]) for k in range(3)]
        fr_opt = _forward_rate_opt([d for p in piles for d in p.rows])
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        if surviving:
            continue  # (every config is another pass over ~10^4 joint evaluations)
        L = TreeLocus([ConfigPileup(p, None, fr) for p in piles], [0.01] * 3, prior)
        dens = [sum((L.node(r, {}) for r in trees[n]), mp.mpf(0)) for n in names]
        total = sum(dens, mp.mpf(0))
        for k, d in enumerate(dens):
            got = float(o.log_posteriors[i, k])
            if d == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(d / total))) <= 1e-9, (i, names[k], got, float(mp.log(d / total)))
        assert o.log_posteriors[i, -1] == -math.inf and L.n_joint == int(o.n_base_events[i]), (i, L.n_joint, int(o.n_base_events[i]))
        checked += 1
        if checked >= 3:
            break
    assert checked >= 2


SUBCLONAL_YAML = """
species:
  heterozygosity: 0.001
  ploidy: 2
  genome-size: 3.5e9

samples:
  tumor:
    somatic-effective-mutation-rate: 1e-6
    inheritance:
      clonal:
        from: normal
        somatic: false
  relapse:
    somatic-effective-mutation-rate: 1e-5
    inheritance:
      subclonal:
        from: tumor
  normal:
    somatic-effective-mutation-rate: 1e-10

events:
  tumor_only: "normal:0.0 & tumor:0.25 & relapse:0.0"
  relapse_only: "normal:0.0 & tumor:0.0 & relapse:0.25"
  kept: "normal:0.0 & tumor:0.25 & relapse:0.5"
  germline: "normal:0.5 & tumor:0.5 & relapse:0.5"
  loh: "normal:0.5 & tumor:1.0 & relapse:]0.5,1.0]"
"""


@pytest.mark.parametrize("full_prior", [False, True])
def test_subclonal_inheritance_prior(full_prior):
    """The samples of the reference's tumor-normal-relapse scenario (somatic rates everywhere: the germline odometer of
    calc_prob, prior.rs:400-421; clonal and subclonal inheritance, :458-552) with events on few allele frequencies, so
    that a locus stays at a few hundred joint evaluations."""
    sc = Scenario.from_yaml(SUBCLONAL_YAML, full_prior=full_prior).for_contig("all")
    assert list(sc.sample_names) == ["normal", "relapse", "tumor"]
    flat = sc.flatten()
    trees, names = dict(sc.event_trees()), list(flat.event_names)
    het = mp.mpf("0.001")
    rate = {"normal": mp.mpf("1e-10"), "tumor": mp.mpf("1e-6"), "relapse": mp.mpf("1e-5")}

    def som(who, d):  # prob_somatic_mutation (:440-456); relative_eq!(d, 0.0) with epsilon = f64::EPSILON
        return 1 - rate[who] if abs(d) <= 2.220446049250313e-16 else rate[who]

    def full(vafs):
        vn, vr, vt = vafs
        total = mp.mpf(0)
        for g in (0.0, 0.5, 1.0):  # every sample has somatic variation: germline alt counts 0..ploidy each; clonal and
            m = int(round(2 * g))  # subclonal inheritance leave only equal germlines (:466, :521)
            pop = het / m if m > 0 else 1 - (het / 1 + het / 2)
            sub = som("relapse", vr) if (vt == 0.0 and g == 0.0) else mp.mpf(1)
            total += pop * som("normal", vn - g) * som("tumor", vt - g) * sub
        return total

    def prior(vafs):
        if full_prior or all(v == 0.0 for v in vafs):
            return full(vafs)
        return mp.mpf(0) if full(vafs) == 0 else 1 - full([0.0, 0.0, 0.0])
    b = synth.pedigree(40, seed=71, depth=30)[1]
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=4)
    checked = 0
    for i in range(b.n_loci):
        if o.knife_edge()[i] or int(o.status[i]) & ~abi.ST_IS_ARTIFACT or (int(b.locus_flags[i]) >> abi.LF_VARTYPE_SHIFT) & 3:
            continue
        offs = [int(b.read_offsets[i * 3 + k]) for k in range(4)]
        piles = [Reads(b, offs[k], offs[k + 1]) for k in range(3)]
        fr_opt = _forward_rate_opt([d for p in piles for d in p.rows])
        fr = fr_opt if fr_opt is not None else HALF
        surviving = _surviving_configs(piles, fr_opt is not None)
        n_joint, dens = 0, {}
        for cfg in [None] + surviving:
            L = TreeLocus([ConfigPileup(p, cfg, fr) for p in piles], [0.01] * 3, prior)
            for name in names:
                if cfg is None or name != "absent":
                    dens[(cfg, name)] = sum((L.node(r, {}) for r in trees[name]), mp.mpf(0))
            n_joint += L.n_joint
        plain = [HALF * dens[(None, n)] for n in names]
        twin = sum((HALF / len(ARTIFACT_CONFIGS) * dens[(c, n)] for c in surviving for n in names if n != "absent"), mp.mpf(0))
        total = sum(plain, mp.mpf(0)) + twin
        for k, w in enumerate([p / total for p in plain] + [twin / total]):
            got = float(o.log_posteriors[i, k])
            if w == 0:
                assert got == -math.inf, (i, k, got)
            else:
                assert abs(float(mp.mpf(got) - mp.log(w))) <= 1e-9, (i, k, surviving, got, float(mp.log(w)))
        assert n_joint == int(o.n_base_events[i])
        checked += 1
        if checked >= 10:
            break
    assert checked >= 6

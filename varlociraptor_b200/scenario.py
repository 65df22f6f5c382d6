"""Scenario front-end for the harness: YAML scenario -> normalised formulas -> VAF trees ->
the flattened `vlr_scenario_t` the C-ABI takes.

In a deployment the Rust host keeps varlociraptor's own grammar (src/grammar/*) and only
flattens its `VAFTree`s; this module exists so tests and benchmarks can build the same
trees without it. It mirrors

  * `Scenario` / `Sample` (src/grammar/mod.rs:129-144, 469-600): samples indexed in name order
    (BTreeMap), `contig_universe`, ploidy, resolution (default 0.01), contamination;
  * the formula grammar (src/grammar/formula.pest) and `Formula::normalize`
    (src/grammar/formula.rs:473-485): expression expansion, negation push-down against the
    sample universe (:717-865), merging of same-sample atoms (:575-708), operand sorting (:455-471);
  * `VAFTree::new` incl. `add_missing_samples` and `VAFTree::absent` (src/grammar/vaftree.rs:18-40,168-305);
  * the tumor-normal scenario of `call variants tumor-normal` (src/cli.rs:1151-1173).

`Formula::simplify` (formula.rs:710-714) goes through `boolean_expression::Expr::simplify_via_bdd`
(crate `boolean_expression 0.4`, Cargo.toml; not vendored). After `apply_negations` every literal is
positive, so the formula is a monotone function of its distinct terminals, and the reference panics on
any negative literal that survives (`to_normalized_formula`, formula.rs:553-555): whenever the reference
works at all, the BDD round trip returns a positive, irredundant sum of products, and for a monotone
function that cover is unique - the set of its prime implicants. `Scenario._simplify` computes exactly
that set (distribution with absorption); pinned by the reference's own normalisation tests
(formula.rs:1621-1735, tests/test_scenario_normalize.py).
"""
from __future__ import annotations

import ctypes as C
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple, Union

from . import abi

NAN = float("nan")


def _opt_float(x):
    # PyYAML reads "1e-3" as a string (YAML 1.1 floats need a dot); serde_yaml reads a float
    return None if x is None else float(x)


# ----------------------------------------------------------------------------- spectra
@dataclass(frozen=True)
class VAFRange:
    start: float
    end: float
    left_exclusive: bool
    right_exclusive: bool

    def contains(self, v: float) -> bool:
        l = self.start < v if self.left_exclusive else self.start <= v
        r = self.end > v if self.right_exclusive else self.end >= v
        return l and r

    def is_empty(self) -> bool:
        return self.start == self.end and (self.left_exclusive or self.right_exclusive)

    def is_singleton(self) -> bool:
        return self.start == self.end and not (self.left_exclusive or self.right_exclusive)

    def no_overlap(self, o: "VAFRange") -> bool:
        if self == o:
            return False
        return ((self.end < o.start or self.start > o.end)
                or (self.end <= o.start and (self.right_exclusive or o.left_exclusive))
                or (self.start >= o.end and (self.left_exclusive or o.right_exclusive)))

    def is_complete(self) -> bool:
        return self.start == 0.0 and self.end == 1.0 and not self.left_exclusive and not self.right_exclusive

    def overlap(self, o: "VAFRange") -> str:
        """VAFRange::overlap (formula.rs:1137-1170): how `self` lies relative to `o`."""
        if self == o:
            return "equal"
        if self.no_overlap(o):
            return "none"
        start_right = (self.start >= o.start) if (self.left_exclusive and not o.left_exclusive) \
            else (self.start > o.start)
        end_left = (self.end <= o.end) if (self.right_exclusive and not o.right_exclusive) \
            else (self.end < o.end)
        if start_right:
            return "contained" if end_left else "start"
        return "end" if end_left else "contains"

    def intersect(self, o: "VAFRange") -> "VAFRange":
        """`&a & &b` (formula.rs:1264-1282)."""
        ov = self.overlap(o)
        if ov in ("contained", "equal"):
            return self
        if ov == "contains":
            return o
        if ov == "start":
            return VAFRange(self.start, o.end, self.left_exclusive, o.right_exclusive)
        if ov == "end":
            return VAFRange(o.start, self.end, o.left_exclusive, self.right_exclusive)
        return VAFRange(0.0, 0.0, True, True)

    def union(self, o: "VAFRange") -> Optional["VAFRange"]:
        """First component of `&a | &b` (formula.rs:1284-1302); None when the ranges do not overlap."""
        ov = self.overlap(o)
        if ov == "contained":
            return o
        if ov in ("contains", "equal"):
            return self
        if ov == "start":
            return VAFRange(o.start, self.end, o.left_exclusive, self.right_exclusive)
        if ov == "end":
            return VAFRange(self.start, o.end, self.left_exclusive, o.right_exclusive)
        return None

    def split_at(self, vaf: float):
        """formula.rs:1105-1135; returns (left, right) spectra or None."""
        left = VAFRange(self.start, vaf, self.left_exclusive, True)
        right = VAFRange(vaf, self.end, True, self.right_exclusive)

        def to_spec(r: VAFRange):
            if r.start == r.end:
                if not (r.left_exclusive and self.right_exclusive):
                    return frozenset([r.start])
                return None
            return r
        return to_spec(left), to_spec(right)


Spectrum = Union[VAFRange, frozenset]  # frozenset of floats = VAFSpectrum::Set


def spectrum_contains(sp: Spectrum, v: float) -> bool:
    return sp.contains(v) if isinstance(sp, VAFRange) else (v in sp)


def spectrum_is_empty(sp: Spectrum) -> bool:
    return sp.is_empty() if isinstance(sp, VAFRange) else len(sp) == 0


_VAF_RE = r"(?:0\.\d+|1\.0)"


def parse_vafdef(text: str) -> Spectrum:
    text = text.strip()
    m = re.fullmatch(r"([\[\]])\s*(%s)\s*,\s*(%s)\s*([\[\]])" % (_VAF_RE, _VAF_RE), text)
    if m:
        return VAFRange(float(m.group(2)), float(m.group(3)), m.group(1) == "]", m.group(4) == "[")
    m = re.fullmatch(r"\{(.*)\}", text)
    if m:
        return frozenset(float(x) for x in m.group(1).split(","))
    if re.fullmatch(_VAF_RE, text):
        return frozenset([float(text)])
    raise ValueError("invalid VAF definition: %r" % text)


def parse_universe(text: str) -> List[Spectrum]:
    return [parse_vafdef(t) for t in text.split("|")]


# ----------------------------------------------------------------------------- formulas
@dataclass(frozen=True)
class Atom:
    sample: str
    vafs: Spectrum


@dataclass(frozen=True)
class Variant:
    positive: bool
    refbase: str
    altbase: str


@dataclass(frozen=True)
class Expression:
    identifier: str
    negated: bool = False


@dataclass(frozen=True)
class L2FC:
    sample_a: str
    sample_b: str
    cmp: int
    value: float


@dataclass(frozen=True)
class Const:
    value: bool


@dataclass(frozen=True)
class And:
    operands: tuple


@dataclass(frozen=True)
class Or:
    operands: tuple


@dataclass(frozen=True)
class Not:
    operand: object


_CMP = {"==": abi.CMP_EQ, ">": abi.CMP_GT, ">=": abi.CMP_GE, "<": abi.CMP_LT, "<=": abi.CMP_LE, "!=": abi.CMP_NE}
_CMP_NOT = {abi.CMP_EQ: abi.CMP_NE, abi.CMP_NE: abi.CMP_EQ, abi.CMP_GT: abi.CMP_LE, abi.CMP_LE: abi.CMP_GT,
            abi.CMP_GE: abi.CMP_LT, abi.CMP_LT: abi.CMP_GE}
_IUPAC = {"A": 1, "C": 2, "G": 4, "T": 8, "R": 1 | 4, "Y": 2 | 8, "S": 4 | 2, "W": 1 | 8, "K": 4 | 8, "M": 1 | 2,
          "B": 2 | 4 | 8, "D": 1 | 4 | 8, "H": 1 | 2 | 8, "V": 1 | 2 | 4, "N": 15}


class _Parser:
    """Recursive-descent parser for src/grammar/formula.pest."""

    def __init__(self, text: str):
        self.t = text
        self.i = 0

    def ws(self):
        while self.i < len(self.t) and self.t[self.i] == " ":
            self.i += 1

    def peek(self, s: str) -> bool:
        self.ws()
        return self.t.startswith(s, self.i)

    def eat(self, s: str):
        self.ws()
        if not self.t.startswith(s, self.i):
            raise ValueError("expected %r at %d in %r" % (s, self.i, self.t))
        self.i += len(s)

    def parse(self):
        f = self.formula()
        self.ws()
        if self.i != len(self.t):
            raise ValueError("trailing input at %d in %r" % (self.i, self.t))
        return f

    def formula(self):
        first = self.sub()
        self.ws()
        if self.peek("&"):
            ops = [first]
            while self.peek("&"):
                self.eat("&")
                ops.append(self.sub())
            return And(tuple(ops))
        if self.peek("|"):
            ops = [first]
            while self.peek("|"):
                self.eat("|")
                ops.append(self.sub())
            return Or(tuple(ops))
        return first

    def sub(self):
        self.ws()
        if self.peek("("):
            self.eat("(")
            f = self.formula()
            self.eat(")")
            return f
        if self.peek("!"):
            self.eat("!")
            return Not(self.sub())
        if self.peek("$"):
            self.eat("$")
            return Expression(self.identifier())
        if self.peek("l2fc("):
            self.eat("l2fc(")
            a = self.identifier()
            self.eat(",")
            b = self.identifier()
            self.eat(")")
            op = self.cmp_op()
            self.ws()
            m = re.match(r"-?(?:0|[1-9]\d*)(?:\.\d*)?(?:[eE][+-]?\d+)?", self.t[self.i:])
            if not m:
                raise ValueError("expected number in %r" % self.t)
            self.i += m.end()
            return L2FC(a, b, op, float(m.group(0)))
        if self.peek("false"):
            self.eat("false")
            return Const(False)
        if self.peek("true"):
            self.eat("true")
            return Const(True)
        m = re.match(r"([ACGTRYSWKMBDHVN])\s*>\s*([ACGTRYSWKMBDHVN])(?![\w.:-])", self.t[self.i:])
        if m:
            self.i += m.end()
            return Variant(True, m.group(1), m.group(2))
        ident = self.identifier()
        self.ws()
        if self.peek(":"):
            self.eat(":")
            self.ws()
            m = re.match(r"[\[\]][^\[\]]*[\[\]]|\{[^}]*\}|%s" % _VAF_RE, self.t[self.i:])
            if not m:
                raise ValueError("expected VAF definition at %d in %r" % (self.i, self.t))
            self.i += m.end()
            return Atom(ident, parse_vafdef(m.group(0)))
        # cmp: a <op> b  ==  l2fc(a,b) <op> 0 (formula.rs, `cmp` rule)
        op = self.cmp_op()
        other = self.identifier()
        return L2FC(ident, other, op, 0.0)

    def cmp_op(self) -> int:
        self.ws()
        for s in ("<=", ">=", "!=", "==", "<", ">"):
            if self.t.startswith(s, self.i):
                self.i += len(s)
                return _CMP[s]
        raise ValueError("expected comparison operator at %d in %r" % (self.i, self.t))

    def identifier(self) -> str:
        self.ws()
        m = re.match(r"[A-Za-z0-9_.\-]+", self.t[self.i:])
        if not m:
            raise ValueError("expected identifier at %d in %r" % (self.i, self.t))
        self.i += m.end()
        return m.group(0)


def parse_formula(text: str):
    return _Parser(text).parse()


# ----------------------------------------------------------------------------- scenario
@dataclass
class Contamination:
    by: str
    fraction: float


@dataclass
class Inheritance:
    kind: int
    parents: Tuple[str, ...]
    somatic: bool = False


@dataclass
class SampleDef:
    resolution: float = 0.01
    universe: Union[None, List[Spectrum], Dict[str, List[Spectrum]]] = None  # per contig when a dict
    contamination: Optional[Contamination] = None
    somatic_effective_mutation_rate: Optional[float] = None
    germline_mutation_rate: Optional[float] = None
    ploidy: Optional[int] = None
    inheritance: Optional[Inheritance] = None
    sex: Optional[str] = None


@dataclass
class Species:
    heterozygosity: Optional[float] = None
    germline_mutation_rate: Optional[float] = None
    somatic_effective_mutation_rate: Optional[float] = None
    ploidy: Union[None, int, dict] = None
    vtf_indel: float = 0.0125
    vtf_mnv: float = 0.001
    vtf_sv: float = 0.01


@dataclass
class TreeNode:
    kind: int
    sample: int = 0
    sample_b: int = 0
    cmp: int = 0
    vafs: Optional[Spectrum] = None
    lfc_value: float = 0.0
    positive: bool = True
    refmask: int = 0
    altmask: int = 0
    children: List["TreeNode"] = field(default_factory=list)

    def clone(self) -> "TreeNode":
        return TreeNode(self.kind, self.sample, self.sample_b, self.cmp, self.vafs, self.lfc_value, self.positive,
                        self.refmask, self.altmask, [c.clone() for c in self.children])

    def leafs(self) -> List["TreeNode"]:
        if not self.children:
            return [self]
        out = []
        for c in self.children:
            out.extend(c.leafs())
        return out


class Scenario:
    """Mirror of grammar::Scenario restricted to what the posterior engine consumes."""

    def __init__(self, samples: Dict[str, SampleDef], events: Dict[str, str],
                 species: Optional[Species] = None, expressions: Optional[Dict[str, str]] = None,
                 full_prior: bool = False):
        self.samples = dict(sorted(samples.items()))  # BTreeMap order -> sample indices
        self.sample_names = list(self.samples.keys())
        self.events = dict(sorted(events.items()))  # BTreeMap<String, Formula>
        self.species = species
        self.expressions = {k: parse_formula(v) for k, v in (expressions or {}).items()}
        self.event_formulas = {k: parse_formula(v) for k, v in self.events.items()}
        for k, f in self.event_formulas.items():  # events are registered as expressions (mod.rs:155-162)
            self.expressions.setdefault(k, f)
        self.full_prior = full_prior
        self.contig = "all"  # the reference's tests normalise on "all" (formula.rs:1663)
        self._keepalive = None

    # -- construction helpers
    @classmethod
    def from_yaml(cls, text: str, full_prior: bool = False) -> "Scenario":
        import yaml
        doc = yaml.safe_load(text)
        sp = None
        if doc.get("species"):
            d = doc["species"]
            vtf = d.get("variant-type-fractions", {}) or {}
            ploidy = d.get("ploidy")  # u32 | {contig: u32} | {sex: u32 | {contig: u32}} (mod.rs:288-342)
            sp = Species(_opt_float(d.get("heterozygosity")), _opt_float(d.get("germline-mutation-rate")),
                         _opt_float(d.get("somatic-effective-mutation-rate")), ploidy,
                         float(vtf.get("indel", 0.0125)), float(vtf.get("mnv", 0.001)), float(vtf.get("sv", 0.01)))
        samples = {}
        for name, d in doc["samples"].items():
            d = d or {}
            s = SampleDef()
            if "resolution" in d:
                s.resolution = float(d["resolution"])
            if "universe" in d:  # UniverseDefinition::{Simple, Map} (mod.rs:651-654)
                u = d["universe"]
                s.universe = {c: parse_universe(t) for c, t in u.items()} if isinstance(u, dict) else parse_universe(u)
            if "contamination" in d:
                s.contamination = Contamination(d["contamination"]["by"], float(d["contamination"]["fraction"]))
            s.somatic_effective_mutation_rate = _opt_float(d.get("somatic-effective-mutation-rate"))
            s.germline_mutation_rate = _opt_float(d.get("germline-mutation-rate"))
            s.ploidy = d.get("ploidy")
            s.sex = d.get("sex")
            if "inheritance" in d:
                inh = d["inheritance"]
                if "mendelian" in inh:
                    s.inheritance = Inheritance(abi.INHERIT_MENDELIAN, tuple(inh["mendelian"]["from"]))
                elif "clonal" in inh:
                    s.inheritance = Inheritance(abi.INHERIT_CLONAL, (inh["clonal"]["from"],),
                                                bool(inh["clonal"].get("somatic", True)))
                elif "subclonal" in inh:
                    s.inheritance = Inheritance(abi.INHERIT_SUBCLONAL, (inh["subclonal"]["from"],))
            samples[name] = s
        return cls(samples, doc["events"], sp, doc.get("expressions"), full_prior)

    @classmethod
    def tumor_normal(cls, purity: float = 1.0, full_prior: bool = False) -> "Scenario":
        """The scenario `call variants tumor-normal` builds (src/cli.rs:1151-1173)."""
        samples = {
            "tumor": SampleDef(resolution=0.01, universe=parse_universe("[0.0,1.0]"),
                               contamination=Contamination("normal", 1.0 - purity)),
            "normal": SampleDef(resolution=0.1, universe=parse_universe("[0.0,0.5[ | 0.5 | 1.0")),
        }
        events = {
            "somatic_tumor": "tumor:]0.0,1.0] & normal:0.0",
            "somatic_normal": "tumor:]0.0,1.0] & normal:]0.0,0.5[",
            "germline_het": "tumor:]0.0,1.0] & normal:0.5",
            "germline_hom": "tumor:]0.0,1.0] & normal:1.0",
        }
        return cls(samples, events, None, None, full_prior)

    # -- Sample accessors (grammar/mod.rs:499-600)
    def idx(self, name: str) -> int:
        return self.sample_names.index(name)

    @staticmethod
    def _contig_ploidy(definition, contig: str) -> int:
        """PloidyDefinition::contig_ploidy (mod.rs:296-312)."""
        if isinstance(definition, dict):
            if contig in definition:
                return int(definition[contig])
            if "all" not in definition:
                raise ValueError("ploidy definition for contig %s not found" % contig)
            return int(definition["all"])
        return int(definition)

    def ploidy(self, name: str) -> Optional[int]:
        """Sample::contig_ploidy (mod.rs:581-593) on `self.contig`."""
        s = self.samples[name]
        if s.ploidy is not None:
            return self._contig_ploidy(s.ploidy, self.contig)
        if self.species is None or self.species.ploidy is None:
            return None
        d = self.species.ploidy
        if isinstance(d, dict) and (any(isinstance(v, dict) for v in d.values()) or set(d) <= {"male", "female"}):
            # SexPloidyDefinition::Specific (mod.rs:321-342)
            if s.sex is None:
                raise ValueError("sex specific ploidy definition found but no sex specified in sample")
            if s.sex not in d:
                raise ValueError("ploidy definition for %s not found" % s.sex)
            return self._contig_ploidy(d[s.sex], self.contig)
        return self._contig_ploidy(d, self.contig)

    def is_contig_dependent(self) -> bool:
        """True if some ploidy or universe is given per contig (PloidyDefinition::Map, UniverseDefinition::Map)."""
        sp = self.species.ploidy if self.species else None
        if isinstance(sp, dict) and (set(sp) - {"male", "female"} or any(isinstance(v, dict) for v in sp.values())):
            return True
        return any(isinstance(s.ploidy, dict) or isinstance(s.universe, dict) for s in self.samples.values())

    def contig_signature(self) -> tuple:
        """What of the scenario depends on the contig: per sample its universe and ploidy on `self.contig`."""
        return tuple((n, repr(self.universe(n)), self.ploidy(n)) for n in self.sample_names)

    def for_contig(self, contig: str) -> "Scenario":
        """The scenario as `Caller::configure_model` sees it on `contig` (calling.rs:632-718): universes,
        and with them the event trees, follow the contig's ploidy."""
        import copy
        sc = copy.copy(self)
        sc.contig = contig
        sc._keepalive = None
        return sc

    def somatic_rate(self, name: str) -> Optional[float]:
        s = self.samples[name]
        if s.somatic_effective_mutation_rate is not None:
            return s.somatic_effective_mutation_rate
        return self.species.somatic_effective_mutation_rate if self.species else None

    def germline_rate(self, name: str) -> Optional[float]:
        s = self.samples[name]
        if s.germline_mutation_rate is not None:
            return s.germline_mutation_rate
        return self.species.germline_mutation_rate if self.species else None

    def universe(self, name: str) -> List[Spectrum]:
        s = self.samples[name]
        if isinstance(s.universe, dict):  # per-contig universes fall back to "all" (mod.rs:511-519)
            if self.contig in s.universe:
                return list(s.universe[self.contig])
            if "all" not in s.universe:
                raise ValueError("universe definition for contig %s not found" % self.contig)
            return list(s.universe["all"])
        if s.universe is not None:
            return list(s.universe)
        ploidy = self.ploidy(name)
        # NOTE contig_universe looks at the sample's own somatic rate only (mod.rs:524-528)
        has_somatic = s.somatic_effective_mutation_rate is not None
        if ploidy is not None:
            spectrum = sorted({(n / ploidy if ploidy > 0 else 0.0) for n in range(ploidy + 1)})
            if not has_somatic:
                return [frozenset(spectrum)]
            out: List[Spectrum] = []
            for a, b in zip(spectrum[:-1], spectrum[1:]):
                out.append(VAFRange(a, b, True, True))
            out.append(frozenset(spectrum))
            return out
        if has_somatic:
            return [VAFRange(0.0, 1.0, False, False)]
        raise ValueError("sample needs to define either universe, ploidy or somatic_mutation_rate")

    # -- Formula::normalize
    def _expand(self, f):
        if isinstance(f, And):
            return And(tuple(self._expand(o) for o in f.operands))
        if isinstance(f, Or):
            return Or(tuple(self._expand(o) for o in f.operands))
        if isinstance(f, Not):
            return Not(self._expand(f.operand))
        if isinstance(f, Expression):
            if f.identifier == "absent" and "absent" not in self.expressions:
                g = And(tuple(Atom(n, frozenset([0.0])) for n in self.sample_names))
            else:
                g = self._expand(self.expressions[f.identifier])
            return Not(g) if f.negated else g
        return f

    def _negate(self, f):
        if isinstance(f, Const):
            return Const(not f.value)
        if isinstance(f, And):
            return Or(tuple(self._negate(o) for o in f.operands))
        if isinstance(f, Or):
            return And(tuple(self._negate(o) for o in f.operands))
        if isinstance(f, Not):
            return self._apply_negations(f.operand)
        if isinstance(f, Variant):
            return Variant(not f.positive, f.refbase, f.altbase)
        if isinstance(f, L2FC):
            return L2FC(f.sample_a, f.sample_b, _CMP_NOT[f.cmp], f.value)
        if isinstance(f, Atom):
            universe = self.universe(f.sample)
            disj: List[Spectrum] = []
            if isinstance(f.vafs, frozenset):
                stack = list(universe)
                while stack:
                    u = stack.pop(0)
                    if isinstance(u, frozenset):
                        diff = frozenset(u - f.vafs)
                        if diff:
                            disj.append(diff)
                    else:
                        for vaf in sorted(f.vafs):
                            if u.contains(vaf):
                                left, right = u.split_at(vaf)
                                if right is not None:
                                    stack.append(right)
                                if left is not None:
                                    disj.append(left)
                            else:
                                disj.append(u)
            else:
                rng = f.vafs
                for u in universe:
                    if isinstance(u, frozenset):
                        rest = frozenset(v for v in u if not rng.contains(v))
                        if rest:
                            disj.append(rest)
                    else:
                        if rng == u:
                            continue
                        if rng.no_overlap(u):
                            disj.append(u)
                            continue
                        # Contained / Start / End / Contains (formula.rs:817-838)
                        start_right = (rng.start >= u.start) if (rng.left_exclusive and not u.left_exclusive) \
                            else (rng.start > u.start)
                        end_left = (rng.end <= u.end) if (rng.right_exclusive and not u.right_exclusive) \
                            else (rng.end < u.end)
                        if start_right:
                            left = u.split_at(rng.start)[0]
                            if left is not None:
                                disj.append(left)
                        if end_left:
                            right = u.split_at(rng.end)[1]
                            if right is not None:
                                disj.append(right)
            if not disj:
                return Atom(f.sample, frozenset())
            return Or(tuple(Atom(f.sample, d) for d in disj))
        raise TypeError(f)

    def _apply_negations(self, f):
        """Move negations into the atoms (formula.rs:868-921); the result is negation-free."""
        if isinstance(f, Not):
            return self._negate(f.operand)
        if isinstance(f, And):
            return And(tuple(self._apply_negations(o) for o in f.operands))
        if isinstance(f, Or):
            return Or(tuple(self._apply_negations(o) for o in f.operands))
        return f

    @staticmethod
    def _flatten(f):
        """Flatten nested And/And and Or/Or and unwrap single-operand nodes."""
        if isinstance(f, (And, Or)):
            ops = []
            for o in f.operands:
                o = Scenario._flatten(o)
                if type(o) is type(f):
                    ops.extend(o.operands)
                else:
                    ops.append(o)
            if len(ops) == 1:
                return ops[0]
            return type(f)(tuple(ops))
        return f

    @staticmethod
    def _merge_atoms(f):
        if isinstance(f, And):
            groups: Dict[Optional[str], list] = {}
            for o in f.operands:
                key = o.sample if isinstance(o, Atom) else None
                groups.setdefault(key, []).append(o)
            ops = []
            for key, stmts in groups.items():
                if key is None:
                    ops.extend(Scenario._merge_atoms(s) for s in stmts)
                    continue
                cur = stmts[0].vafs
                for s in stmts[1:]:
                    o = s.vafs
                    if isinstance(cur, VAFRange) and isinstance(o, VAFRange):
                        inter = cur.intersect(o)
                        cur = frozenset([inter.start]) if inter.is_singleton() else inter
                    elif isinstance(cur, VAFRange):
                        cur = frozenset(v for v in o if cur.contains(v))
                    elif isinstance(o, VAFRange):
                        cur = frozenset(v for v in cur if o.contains(v))
                    else:
                        cur = frozenset(cur & o)
                if spectrum_is_empty(cur):
                    return Const(False)
                ops.append(Atom(key, cur))
            return And(tuple(ops)) if len(ops) > 1 else ops[0]
        if isinstance(f, Or):
            groups = {}
            for o in f.operands:
                key = o.sample if isinstance(o, Atom) else None
                groups.setdefault(key, []).append(o)
            ops = []
            for key, stmts in groups.items():
                if key is None:
                    ops.extend(Scenario._merge_atoms(s) for s in stmts)
                    continue
                stmts = sorted(stmts, key=lambda a: a.vafs.start if isinstance(a.vafs, VAFRange) else min(a.vafs))
                cur = stmts[0].vafs
                for s in stmts[1:]:
                    o = s.vafs
                    merged = None
                    if isinstance(cur, frozenset) and isinstance(o, frozenset):
                        merged = frozenset(cur | o)
                    elif isinstance(cur, VAFRange) and isinstance(o, frozenset):
                        merged = cur if all(cur.contains(v) for v in o) else None
                    elif isinstance(cur, frozenset) and isinstance(o, VAFRange):
                        merged = o if all(o.contains(v) for v in cur) else None
                    elif isinstance(cur, VAFRange) and isinstance(o, VAFRange):
                        merged = cur.union(o)  # None when they do not overlap (formula.rs:157-160)
                    if merged is not None:
                        cur = merged
                    else:
                        ops.append(Atom(key, cur))
                        cur = o
                ops.append(Atom(key, cur))
            return Or(tuple(ops)) if len(ops) > 1 else ops[0]
        return f

    @staticmethod
    def _sort_key(f):
        # derived Ord: Conjunction < Disjunction < Negation < Terminal; LFC terminals first (formula.rs:455-471)
        if isinstance(f, L2FC):
            return (0, 3, f.sample_a, f.sample_b)
        if isinstance(f, (And, Or)):  # Vec<Formula> compares lexicographically
            return (1, 0 if isinstance(f, And) else 1, tuple(Scenario._sort_key(o) for o in f.operands), "")
        if isinstance(f, Atom):
            v = f.vafs
            sub = (0, tuple(sorted(v))) if isinstance(v, frozenset) else (1, (v.start, v.end))
            return (1, 2, f.sample, sub)
        if isinstance(f, Variant):
            return (1, 3, (f.positive, f.refbase, f.altbase), "")
        return (1, 4, "", "")

    @staticmethod
    def _sort(f):
        if isinstance(f, (And, Or)):
            ops = [Scenario._sort(o) for o in f.operands]
            ops.sort(key=Scenario._sort_key)
            return type(f)(tuple(ops))
        return f

    @staticmethod
    def _absorb(cubes):
        """Drop every cube that is a superset of another one (x | x & y = x)."""
        out = []
        for c in sorted(set(cubes), key=len):
            if not any(k <= c for k in out):
                out.append(c)
        return out

    @classmethod
    def _prime_implicants(cls, f):
        """Cubes (frozensets of terminals) of the minimal positive DNF of the negation-free formula `f`."""
        if isinstance(f, Const):
            return [frozenset()] if f.value else []
        if isinstance(f, Atom):  # From<Formula> for Expr (formula.rs:369-379): empty -> false, complete -> true
            if spectrum_is_empty(f.vafs):
                return []
            if isinstance(f.vafs, VAFRange) and f.vafs.is_complete():
                return [frozenset()]
            return [frozenset([f])]
        if isinstance(f, (Variant, L2FC)):
            return [frozenset([f])]
        if isinstance(f, Or):
            cubes = []
            for o in f.operands:
                cubes.extend(cls._prime_implicants(o))
            return cls._absorb(cubes)
        if isinstance(f, And):
            cubes = [frozenset()]
            for o in f.operands:
                sub = cls._prime_implicants(o)
                cubes = cls._absorb([a | b for a in cubes for b in sub])
            return cubes
        raise TypeError("negations and expressions must be resolved before simplification: %r" % (f,))

    @classmethod
    def _simplify(cls, f):
        """Formula::simplify (formula.rs:710-714), see the module docstring."""
        cubes = cls._prime_implicants(f)
        if not cubes:
            return Const(False)
        if any(len(c) == 0 for c in cubes):
            return Const(True)
        terms = []
        for c in cubes:
            lits = sorted(c, key=cls._sort_key)
            terms.append(lits[0] if len(lits) == 1 else And(tuple(lits)))
        terms.sort(key=cls._sort_key)  # cubes come out of sets: fix the order
        return terms[0] if len(terms) == 1 else Or(tuple(terms))

    def normalize(self, f):
        """Formula::normalize (formula.rs:473-485)."""
        g = self._simplify(self._apply_negations(self._expand(f)))
        g = self._simplify(self._flatten(self._merge_atoms(g)))
        if isinstance(g, Or):  # strip_false
            keep = [o for o in g.operands if not (isinstance(o, Const) and not o.value)
                    and not (isinstance(o, And) and any(isinstance(x, Const) and not x.value for x in o.operands))]
            g = Or(tuple(keep)) if len(keep) > 1 else (keep[0] if keep else Const(False))
        return self._sort(g)

    # -- VAFTree::new (vaftree.rs:168-305)
    def _subtrees(self, f) -> List[TreeNode]:
        if isinstance(f, Atom):
            kind = abi.NODE_RANGE if isinstance(f.vafs, VAFRange) else abi.NODE_SET
            return [TreeNode(kind, sample=self.idx(f.sample), vafs=f.vafs)]
        if isinstance(f, Or):
            out = []
            for o in f.operands:
                out.extend(self._subtrees(o))
            return out
        if isinstance(f, And):
            ops = sorted(f.operands, key=lambda o: 1 if isinstance(o, Or) else 0)  # stable
            roots = self._subtrees(ops[0])
            for o in ops[1:]:
                sub = self._subtrees(o)
                for r in roots:
                    for leaf in r.leafs():
                        leaf.children = [s.clone() for s in sub]
            return roots
        if isinstance(f, Variant):
            return [TreeNode(abi.NODE_VARIANT, positive=f.positive, refmask=_IUPAC[f.refbase],
                             altmask=_IUPAC[f.altbase])]
        if isinstance(f, Const):
            return [TreeNode(abi.NODE_TRUE if f.value else abi.NODE_FALSE)]
        if isinstance(f, L2FC):
            return [TreeNode(abi.NODE_LFC, sample=self.idx(f.sample_a), sample_b=self.idx(f.sample_b), cmp=f.cmp,
                             lfc_value=f.value)]
        raise TypeError(f)

    def _add_missing(self, node: TreeNode, seen: set):
        if node.kind == abi.NODE_FALSE:
            return
        if node.kind in (abi.NODE_SET, abi.NODE_RANGE):
            seen.add(node.sample)
        if not node.children:
            for name in self.sample_names:
                i = self.idx(name)
                if i not in seen:
                    seen.add(i)
                    node.children = [
                        TreeNode(abi.NODE_RANGE if isinstance(u, VAFRange) else abi.NODE_SET, sample=i, vafs=u)
                        for u in self._universe_order(name)]
                    self._add_missing(node, seen)
                    break
        else:
            for c in node.children[1:]:
                self._add_missing(c, set(seen))
            self._add_missing(node.children[0], seen)

    def _universe_order(self, name: str) -> List[Spectrum]:
        # VAFUniverse is a HashSet in the reference (iteration order random); fixed here: ranges by start, then sets
        u = self.universe(name)
        return sorted(u, key=lambda sp: (0, sp.start, sp.end) if isinstance(sp, VAFRange) else (1, min(sp), 0))

    def vaftree(self, formula) -> List[TreeNode]:
        roots = self._subtrees(self.normalize(formula))
        for r in roots:
            self._add_missing(r, set())
        return roots

    def absent_tree(self) -> List[TreeNode]:
        n = len(self.sample_names)
        node = None
        for s in reversed(range(n)):
            node = TreeNode(abi.NODE_SET, sample=s, vafs=frozenset([0.0]), children=[node] if node else [])
        return [node]

    def validate(self) -> None:
        """Scenario::validate (grammar/mod.rs:223-278): two events are rejected when their disjunction normalises to
        one of the scenario's events, i.e. when one contains the other ("the following events are not disjunct")."""
        by_formula: Dict[str, List[str]] = {}
        normal = {}
        for name, f in self.event_formulas.items():
            if name == "absent":
                continue
            g = self.normalize(f)
            normal[repr(g)] = g
            by_formula.setdefault(repr(g), []).append(name)
        keys = sorted(normal)
        overlapping = []
        for i, k1 in enumerate(keys):
            for k2 in keys[i + 1:]:
                e1, e2 = normal[k1], normal[k2]
                if any(isinstance(e, Const) and not e.value for e in (e1, e2)):
                    continue  # a terminal `false` overlaps with nothing
                d = repr(self.normalize(Or((e1, e2))))
                if d in normal:
                    overlapping.append("(%s | %s) = %s" % (by_formula[k1], by_formula[k2], by_formula[d]))
        if overlapping:
            raise ValueError("the following events are not disjunct: " + ", ".join(overlapping))

    def event_trees(self) -> List[Tuple[str, List[TreeNode]]]:
        """[absent] + scenario events (calling.rs:655-687); scenario events in name order. Validates the events like
        `Scenario::vaftrees` does (grammar/mod.rs:206-221)."""
        self.validate()
        out = [("absent", self.absent_tree())]
        for name, f in self.event_formulas.items():
            if name == "absent":
                continue
            out.append((name, self.vaftree(f)))
        return out

    # -- flatten to the C-ABI
    def flatten(self) -> "FlatScenario":
        nodes: List[abi.Node] = []
        set_vafs: List[float] = []
        events = []

        def add_vafs(vs) -> Tuple[int, int]:
            off = len(set_vafs)
            set_vafs.extend(sorted(vs))
            return off, len(vs)

        def place(children: List[TreeNode]) -> int:
            """Reserve a contiguous block for `children`, fill recursively, return first index."""
            first = len(nodes)
            for _ in children:
                nodes.append(abi.Node())
            for k, ch in enumerate(children):
                n = nodes[first + k]
                n.kind = ch.kind
                n.sample, n.sample_b, n.cmp = ch.sample, ch.sample_b, ch.cmp
                n.lfc_value = ch.lfc_value
                n.variant_positive = 1 if ch.positive else 0
                n.refmask, n.altmask = ch.refmask, ch.altmask
                if ch.kind == abi.NODE_SET:
                    n.vaf_offset, n.n_vafs = add_vafs(ch.vafs)
                elif ch.kind == abi.NODE_RANGE:
                    n.start, n.end = ch.vafs.start, ch.vafs.end
                    n.left_exclusive = 1 if ch.vafs.left_exclusive else 0
                    n.right_exclusive = 1 if ch.vafs.right_exclusive else 0
                n.n_children = len(ch.children)
                n.first_child = place(ch.children) if ch.children else 0
            return first

        for name, roots in self.event_trees():
            ev = abi.Event()
            ev.name = name.encode()[:63]
            ev.n_roots = len(roots)
            ev.first_root = place(roots)
            ev.has_artifact_twin = 0 if name == "absent" else 1
            events.append(ev)

        spectra: List[abi.Spectrum] = []
        samples = []
        for name in self.sample_names:
            sd = self.samples[name]
            sm = abi.Sample()
            sm.resolution = sd.resolution
            if sd.contamination is not None:
                sm.contamination_by = self.idx(sd.contamination.by)
                sm.contamination_fraction = sd.contamination.fraction
            else:
                sm.contamination_by = -1
                sm.contamination_fraction = 0.0
            g, so = self.germline_rate(name), self.somatic_rate(name)
            sm.germline_mutation_rate = NAN if g is None else g
            sm.somatic_effective_mutation_rate = NAN if so is None else so
            sm.uniform_prior = 1 if sd.universe is not None else 0
            pl = self.ploidy(name)
            sm.ploidy = -1 if pl is None else pl
            sm.inheritance = abi.INHERIT_NONE
            sm.parent_a = sm.parent_b = -1
            if sd.inheritance is not None:
                sm.inheritance = sd.inheritance.kind
                sm.parent_a = self.idx(sd.inheritance.parents[0])
                if len(sd.inheritance.parents) > 1:
                    sm.parent_b = self.idx(sd.inheritance.parents[1])
                sm.clonal_somatic = 1 if sd.inheritance.somatic else 0
            sm.universe_offset = len(spectra)
            for u in self._universe_order(name):
                sp = abi.Spectrum()
                if isinstance(u, VAFRange):
                    sp.kind = abi.SPECTRUM_RANGE
                    sp.start, sp.end = u.start, u.end
                    sp.left_exclusive, sp.right_exclusive = int(u.left_exclusive), int(u.right_exclusive)
                else:
                    sp.kind = abi.SPECTRUM_SET
                    sp.vaf_offset, sp.n_vafs = add_vafs(u)
                spectra.append(sp)
            sm.n_universe = len(spectra) - sm.universe_offset
            samples.append(sm)
        return FlatScenario(self, samples, events, nodes, set_vafs, spectra)


class FlatScenario:
    """Owns the ctypes arrays behind a `vlr_scenario_t`."""

    def __init__(self, scenario: Scenario, samples, events, nodes, set_vafs, spectra):
        self.scenario = scenario
        self.n_samples = len(samples)
        self.n_events = len(events)
        self.event_names = [e.name.decode() for e in events]
        self.sample_names = list(scenario.sample_names)
        self._samples = (abi.Sample * max(1, len(samples)))(*samples)
        self._events = (abi.Event * max(1, len(events)))(*events)
        self._nodes = (abi.Node * max(1, len(nodes)))(*nodes)
        self._set_vafs = (C.c_double * max(1, len(set_vafs)))(*set_vafs)
        self._spectra = (abi.Spectrum * max(1, len(spectra)))(*spectra)
        sp = scenario.species
        c = abi.Scenario()
        c.abi_version = abi.VLR_ABI_VERSION
        c.n_samples, c.n_events, c.n_nodes = len(samples), len(events), len(nodes)
        c.n_set_vafs, c.n_spectra = len(set_vafs), len(spectra)
        c.samples = C.cast(self._samples, C.POINTER(abi.Sample))
        c.events = C.cast(self._events, C.POINTER(abi.Event))
        c.nodes = C.cast(self._nodes, C.POINTER(abi.Node))
        c.set_vafs = C.cast(self._set_vafs, C.POINTER(C.c_double))
        c.spectra = C.cast(self._spectra, C.POINTER(abi.Spectrum))
        c.heterozygosity = NAN if (sp is None or sp.heterozygosity is None) else sp.heterozygosity
        c.vtf_indel = sp.vtf_indel if sp else 0.0125
        c.vtf_mnv = sp.vtf_mnv if sp else 0.001
        c.vtf_sv = sp.vtf_sv if sp else 0.01
        c.full_prior = 1 if scenario.full_prior else 0
        self.c = c

    def describe(self) -> str:
        """Human-readable dump of the flattened trees (for DESIGN.md / debugging)."""
        lines = []

        def rec(i, depth):
            n = self._nodes[i]
            if n.kind == abi.NODE_SET:
                vs = [self._set_vafs[n.vaf_offset + k] for k in range(n.n_vafs)]
                d = "%s:{%s}" % (self.sample_names[n.sample], ",".join("%g" % v for v in vs))
            elif n.kind == abi.NODE_RANGE:
                d = "%s:%s%g,%g%s" % (self.sample_names[n.sample], "]" if n.left_exclusive else "[", n.start, n.end,
                                      "[" if n.right_exclusive else "]")
            elif n.kind == abi.NODE_LFC:
                d = "l2fc(%s,%s) op%d %g" % (self.sample_names[n.sample], self.sample_names[n.sample_b], n.cmp,
                                              n.lfc_value)
            elif n.kind == abi.NODE_VARIANT:
                d = "%svariant(%d>%d)" % ("" if n.variant_positive else "!", n.refmask, n.altmask)
            else:
                d = "true" if n.kind == abi.NODE_TRUE else "false"
            lines.append("  " * depth + d)
            for k in range(n.n_children):
                rec(n.first_child + k, depth + 1)

        for e in range(self.n_events):
            ev = self._events[e]
            lines.append("%s:" % ev.name.decode())
            for r in range(ev.n_roots):
                rec(ev.first_root + r, 1)
        return "\n".join(lines)

"""Formula::normalize (src/grammar/formula.rs:473-485) against the reference's own normalisation tests
(formula.rs:1621-1735), re-expressed on varlociraptor_b200.scenario, plus the effects of the BDD
simplification step (formula.rs:710-714) on the VAF trees the engine receives."""
import os

import pytest

from varlociraptor_b200 import abi
from varlociraptor_b200.scenario import And, Atom, Const, Or, Scenario, VAFRange, parse_formula

MERGE_ATOMS_YAML = """
species:
  heterozygosity: 0.001
  germline-mutation-rate: 1e-3
  ploidy:
    male:
        all: 2
        X: 1
        Y: 1
    female:
        all: 2
        X: 2
        Y: 0
  genome-size: 3.5e9

samples:
  tumor:
    sex: female
    somatic-effective-mutation-rate: 1e-6
    inheritance:
      clonal:
        from: normal
        somatic: false
    contamination:
      by: normal
      fraction: 0.11
  normal:
    sex: female
    somatic-effective-mutation-rate: 1e-10

expressions:
  loh: "normal:0.5 & tumor:1.0"
  loh_or_amplification: "normal:0.5 & tumor:[0.9,1.0["
events:
  germline: "(normal:0.5 | normal:1.0) & !($loh | $loh_or_amplification)"
  expected: "(normal:0.5 & tumor:{0.0, 0.5}) | (normal:0.5 & tumor:]0.0,0.5[) | (normal:0.5 & tumor:]0.5,0.9[) | normal:1.0"
"""


def _single(events):
    text = 'samples:\n  normal:\n    resolution: 0.01\n    universe: "[0.0,1.0]"\nevents:\n'
    text += "".join('  %s: "%s"\n' % kv for kv in events.items())
    return Scenario.from_yaml(text)


def _norm(sc, name):
    return sc.normalize(sc.event_formulas[name])


def test_merge_atoms():  # formula.rs:1621-1671
    sc = Scenario.from_yaml(MERGE_ATOMS_YAML)
    germline = _norm(sc, "germline")
    assert germline == _norm(sc, "expected")
    assert isinstance(germline, Or) and len(germline.operands) == 4
    assert germline.operands[-1] == Atom("normal", frozenset([1.0]))  # conjunctions sort before terminals


def test_range_conjunction():  # formula.rs:1673-1698
    sc = _single({"full": "normal:[0.0,1.0]", "part1": "normal:[0.0,0.7]", "part2": "normal:[0.3,1.0]",
                  "expected": "normal:[0.3,0.7]"})
    conj = sc.normalize(And((sc.event_formulas["part1"], sc.event_formulas["part2"])))
    assert conj == _norm(sc, "expected") == Atom("normal", VAFRange(0.3, 0.7, False, False))
    assert conj != _norm(sc, "full")


def test_nested_range_disjunction():  # formula.rs:1700-1716
    sc = _single({"full": "(normal:[0.0, 0.25] | normal:[0.5,0.75]) | (normal:[0.25,0.5] | normal:[0.75,1.0]) "
                          "| normal:[0.1,0.4] | normal:0.1",
                  "expected": "normal:[0.0,1.0]"})
    assert _norm(sc, "full") == _norm(sc, "expected")
    # a complete range is the constant `true` for the BDD (formula.rs:369-379)
    assert _norm(sc, "expected") == Const(True)


def test_two_separate_range_disjunctions():  # formula.rs:1718-1734
    sc = _single({"full": "(normal:[0.0, 0.25] | normal:[0.5,0.6]) | ((normal:[0.25,0.5] | normal:[0.7,0.9]) "
                          "| normal:[0.9,1.0]) | normal:]0.8,0.9[ | normal:0.75",
                  "expected": "normal:[0.0,0.6] | normal:[0.7,1.0]"})
    full = _norm(sc, "full")
    assert full == _norm(sc, "expected")
    assert full == Or((Atom("normal", VAFRange(0.0, 0.6, False, False)),
                       Atom("normal", VAFRange(0.7, 1.0, False, False))))


def test_range_overlap_classes():  # formula.rs:1137-1170, 1264-1302
    a, b = VAFRange(0.0, 0.7, False, True), VAFRange(0.3, 1.0, False, False)
    assert a.overlap(b) == "end" and b.overlap(a) == "start"
    assert a.intersect(b) == b.intersect(a) == VAFRange(0.3, 0.7, False, True)  # formula.rs:1601-1619
    assert a.union(b) == b.union(a) == VAFRange(0.0, 1.0, False, False)
    assert VAFRange(0.0, 0.5, False, True).union(VAFRange(0.5, 1.0, True, False)) is None
    assert VAFRange(0.2, 0.3, False, False).overlap(VAFRange(0.0, 1.0, False, False)) == "contained"
    assert VAFRange(0.0, 1.0, False, False).overlap(VAFRange(0.2, 0.3, True, True)) == "contains"
    # equal ends are "not left of" each other (formula.rs:1151-1156): ]0,0.5] starts inside [0,0.5] and ends with it
    assert VAFRange(0.0, 0.5, True, False).overlap(VAFRange(0.0, 0.5, False, False)) == "start"


def test_simplification_distributes_and_absorbs():
    sc = Scenario.from_yaml(MERGE_ATOMS_YAML)
    a, b, c = (Atom("normal", frozenset([v])) for v in (0.0, 0.5, 1.0))
    t = Atom("tumor", VAFRange(0.0, 0.5, True, True))
    # (a | b) & t  ->  a & t | b & t ; x | x & y -> x ; false operands vanish, true operands vanish from products
    assert Scenario._simplify(And((Or((a, b)), t))) == Or((And((a, t)), And((b, t))))
    assert Scenario._simplify(Or((a, And((a, t)), Const(False)))) == a
    assert Scenario._simplify(And((a, Const(True), Atom("tumor", VAFRange(0.0, 1.0, False, False))))) == a
    assert Scenario._simplify(And((a, Atom("tumor", frozenset())))) == Const(False)
    assert sc.normalize(And((a, c))) == Const(False)  # contradictory atoms of one sample (formula.rs:597-603)
    with pytest.raises(TypeError):
        Scenario._simplify(parse_formula("!normal:0.5"))


def test_pedigree_event_becomes_one_branch_per_child_genotype():
    from varlociraptor_b200 import synth
    sc = Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML)
    f = _norm(sc, "denovo_child")
    assert isinstance(f, Or) and [len(o.operands) for o in f.operands] == [3, 3]
    roots = sc.vaftree(sc.event_formulas["denovo_child"])
    assert [(r.kind, sorted(r.vafs)) for r in roots] == [(abi.NODE_SET, [0.5]), (abi.NODE_SET, [1.0])]


def test_complete_atom_is_dropped_and_re_added_as_missing_sample():
    """`other` of the contamination scenario (estimation/contamination.rs:446-447): `sample:[0.0,1.0]` is the
    constant true, so the tree is rooted at the contaminant and the sample's universe is appended below it."""
    sc = Scenario.from_yaml("""
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
""")
    assert _norm(sc, "other") == Atom("contaminant", VAFRange(0.0, 1.0, True, False))
    (root,) = sc.vaftree(sc.event_formulas["other"])
    assert root.sample == sc.idx("contaminant") and root.children[0].sample == sc.idx("sample")
    assert root.children[0].vafs == VAFRange(0.0, 1.0, False, False)


def test_ploidy_follows_contig_and_sex():  # grammar/mod.rs:288-342, 581-593
    sc = Scenario.from_yaml(MERGE_ATOMS_YAML.replace("sex: female\n    somatic-effective-mutation-rate: 1e-10",
                                                     "sex: male\n    somatic-effective-mutation-rate: 1e-10"))
    assert sc.ploidy("normal") == 2 and sc.ploidy("tumor") == 2
    x = sc.for_contig("X")
    assert x.ploidy("normal") == 1 and x.ploidy("tumor") == 2
    assert frozenset([0.0, 1.0]) in x.universe("normal") and frozenset([0.0, 0.5, 1.0]) in x.universe("tumor")
    y = sc.for_contig("Y")
    assert y.ploidy("tumor") == 0 and y.universe("tumor") == [frozenset([0.0])]
    with pytest.raises(ValueError):
        Scenario.from_yaml(MERGE_ATOMS_YAML.replace("        all: 2\n", "")).ploidy("normal")


def test_universe_follows_contig():  # UniverseDefinition::Map (grammar/mod.rs:508-520, 651-654)
    sc = Scenario.from_yaml("""
samples:
  a:
    universe: "[0.0,1.0]"
  b:
    universe:
      all: "[0.0,1.0]"
      Y: "0.0 | 1.0"
events:
  both: "a:]0.0,1.0] & !b:0.0"
  not_half: "!b:0.5"
""")
    assert sc.universe("b") == [VAFRange(0.0, 1.0, False, False)]
    assert _norm(sc, "not_half") == Or((Atom("b", VAFRange(0.0, 0.5, False, True)),
                                        Atom("b", VAFRange(0.5, 1.0, True, False))))
    # upstream quirk, mirrored: splitting [0,1] at its own inclusive start keeps the singleton {0} on the left
    # (`split_at`'s emptiness test looks at the wrong bounds, formula.rs:1120-1131), so !b:0.0 still contains b:0.0
    a = Atom("a", VAFRange(0.0, 1.0, True, False))
    assert _norm(sc, "both") == Or((And((a, Atom("b", frozenset([0.0])))), And((a, Atom("b", VAFRange(0.0, 1.0, True, False))))))
    y = sc.for_contig("Y")
    assert y.universe("b") == [frozenset([0.0]), frozenset([1.0])]
    assert _norm(y, "both") == And((a, Atom("b", frozenset([1.0]))))
    assert _norm(y, "not_half") == Atom("b", frozenset([0.0, 1.0]))  # set differences, merged by the disjunction
    assert y.flatten().c.samples[1].uniform_prior == 1
    with pytest.raises(ValueError):
        Scenario.from_yaml('samples:\n  a:\n    universe:\n      X: "0.0"\nevents:\n  e: "a:0.0"\n').universe("a")


REFERENCE_RESOURCES = "/root/reference/tests/resources"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_RESOURCES), reason="reference checkout not present (GPU box)")
def test_every_scenario_of_the_reference_test_suite_normalises_and_flattens():
    """All scenario files under the reference's tests/resources go through parse -> normalize -> VAF trees -> C-ABI
    structs; events of one scenario must stay pairwise distinct after normalisation."""
    import glob
    files = sorted(glob.glob(os.path.join(REFERENCE_RESOURCES, "**", "*scenario*.y*ml"), recursive=True))
    assert len(files) >= 100
    for path in files:
        sc = Scenario.from_yaml(open(path).read())
        if os.path.basename(os.path.dirname(path)) == "test_overlapping_events":  # testcase_should_panic! upstream
            with pytest.raises(ValueError, match="not disjunct"):
                sc.flatten()
            continue
        flat = sc.flatten()
        assert flat.n_events == len(sc.events) + (0 if "absent" in sc.events else 1), path
        normal = [repr(sc.normalize(f)) for f in sc.event_formulas.values()]
        assert len(set(normal)) == len(normal), path


def test_overlapping_events_are_rejected():
    """Scenario::validate (grammar/mod.rs:223-278); the reference's `testcase_should_panic!(test_overlapping_events)`
    (tests/lib.rs:160): `germline` contains `germline_het` and `germline_hom`, `somatic` contains its two parts."""
    sc = Scenario.from_yaml("""
samples:
  tumor:
    contamination:
      by: normal
      fraction: 0.25
    resolution: 0.01
    universe: "[0.0,1.0]"
  normal:
    resolution: 0.1
    universe: "[0.0,0.5[ | 0.5 | 1.0"
events:
  somatic: "tumor:]0.0,1.0] & normal:[0.0,0.5["
  somatic_tumor:  "tumor:]0.0,1.0] & normal:0.0"
  germline: "normal:0.5 | normal:1.0"
  germline_het:   "tumor:[0.0,1.0] & normal:0.5"
""")
    with pytest.raises(ValueError, match="the following events are not disjunct") as err:
        sc.flatten()
    assert "['germline'] | ['germline_het']) = ['germline']" in str(err.value)
    # containment that needs interval reasoning inside a conjunction (`somatic` covers `somatic_tumor`) is not found:
    # atoms are only merged at the top level of a disjunction, upstream as well (formula.rs:636-706)
    assert "somatic" not in str(err.value)
    Scenario.tumor_normal(0.75).validate()  # the CLI's own scenario is disjunct


# ------------------------------------------------------------------------------------------ property: meaning is preserved
def _holds(f, vafs):
    """Truth of a (normalised or raw) formula at one VAF assignment {sample: vaf}."""
    from varlociraptor_b200.scenario import Not, spectrum_contains
    if isinstance(f, Const):
        return f.value
    if isinstance(f, Atom):
        return spectrum_contains(f.vafs, vafs[f.sample])
    if isinstance(f, And):
        return all(_holds(o, vafs) for o in f.operands)
    if isinstance(f, Or):
        return any(_holds(o, vafs) for o in f.operands)
    if isinstance(f, Not):
        return not _holds(f.operand, vafs)
    raise TypeError(f)


def test_normalisation_preserves_the_meaning_of_random_formulas():
    """expand -> negations -> simplify -> merge atoms -> simplify must not change which VAF assignments satisfy an
    event. Random formulas over two samples with discrete universes (negation against a range universe is excluded:
    upstream's `split_at` keeps an inclusive boundary point on the wrong side, mirrored and tested above)."""
    from hypothesis import given, settings, strategies as st
    from varlociraptor_b200.scenario import Not
    sc = Scenario.from_yaml("""
samples:
  a:
    universe: "0.0 | 0.25 | 0.5 | 0.75 | 1.0"
  b:
    universe: "{0.0,0.5} | 1.0"
events:
  e: "a:0.5"
""")
    points = [0.0, 0.25, 0.5, 0.75, 1.0]
    spectra = st.one_of(
        st.frozensets(st.sampled_from(points), min_size=1, max_size=3),
        st.builds(lambda lo, hi, le, re: VAFRange(min(lo, hi), max(lo, hi), le, re), st.sampled_from(points),
                  st.sampled_from(points), st.booleans(), st.booleans()).filter(lambda r: r.start < r.end))
    atoms = st.builds(Atom, st.sampled_from(["a", "b"]), spectra)
    set_atoms = st.builds(Atom, st.sampled_from(["a", "b"]), st.frozensets(st.sampled_from(points), min_size=1, max_size=3))
    formulas = st.recursive(st.one_of(atoms, st.builds(Not, set_atoms)),
                            lambda inner: st.one_of(st.builds(lambda xs: And(tuple(xs)), st.lists(inner, min_size=2, max_size=3)),
                                                    st.builds(lambda xs: Or(tuple(xs)), st.lists(inner, min_size=2, max_size=3))),
                            max_leaves=8)

    @settings(max_examples=300, deadline=None)
    @given(formulas)
    def check(f):
        g = sc.normalize(f)
        for va in points:
            for vb in (0.0, 0.5, 1.0):  # b's universe
                assert _holds(f, {"a": va, "b": vb}) == _holds(g, {"a": va, "b": vb}), (f, g, va, vb)
    check()

"""Generates the committed golden fixtures from the reference's own test resources.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py

Outputs (committed):
  flamegraph_obs.npz        the 11 preprocessed loci of tests/resources/flamegraph_profiling/normal.vcf
                            decoded into the SoA batch layout (inputs of `call variants`)
  flamegraph_expected.json  what the reference printed for them (calls.vcf): PROB_* (PHRED f32),
                            AF, AFD text, DP, SAOBS/SROBS/OBS/OOBS, the raw
                            THIRD_ALLELE_EVIDENCE INFO arrays, plus the scenario
  raw_info_records.json     verbatim INFO integer arrays of three reference records (codec round trip)
  prior_scenarios.json      the six scenario files of tests/resources/prior/scenarios, verbatim
  real_pileups.npz/.json    single-sample observation-format-15 pileups embedded in
                            tests/resources/testcases/*/candidates.vcf (inputs only) with their scenario text
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from varlociraptor_b200 import obs_codec  # noqa: E402
from varlociraptor_b200.batch import LocusBatch  # noqa: E402

REF = "/root/reference/tests/resources"


def flamegraph():
    d = os.path.join(REF, "flamegraph_profiling")
    recs = obs_codec.parse_observation_vcf(os.path.join(d, "normal.vcf"))
    batch = obs_codec.batch_from_records([recs])
    batch.save(os.path.join(HERE, "flamegraph_obs.npz"))
    expected = []
    with open(os.path.join(d, "calls.vcf")) as f:
        for line in f:
            if line.startswith("#"):
                continue
            t = line.rstrip("\n").split("\t")
            info = dict(kv.split("=", 1) for kv in t[7].split(";") if "=" in kv)
            fmt = dict(zip(t[8].split(":"), t[9].split(":")))
            afd = [(float(a), float(b)) for a, b in (x.split("=") for x in fmt["AFD"].split(","))]
            expected.append({"chrom": t[0], "pos": int(t[1]), "info": {k: v for k, v in info.items()
                                                                       if k.startswith("PROB_")},
                             "DP": int(fmt["DP"]), "AF": float(fmt["AF"]), "SAOBS": fmt["SAOBS"],
                             "SROBS": fmt["SROBS"], "OBS": fmt["OBS"], "OOBS": fmt["OOBS"], "AFD": afd})
    with open(os.path.join(d, "scenario.yaml")) as f:
        scenario = f.read()
    assert [r["pos"] for r in recs] == [e["pos"] for e in expected]
    for r, e in zip(recs, expected):  # output-only column the SoA batch does not carry (feeds FORMAT/OBS)
        e["THIRD_ALLELE_EVIDENCE"] = r["info"]["THIRD_ALLELE_EVIDENCE"]
    with open(os.path.join(HERE, "flamegraph_expected.json"), "w") as f:
        json.dump({"source": "tests/resources/flamegraph_profiling/{normal.vcf,calls.vcf,scenario.yaml}",
                   "scenario_yaml": scenario, "records": expected}, f, indent=1)
    print("flamegraph: %d loci, %d reads" % (batch.n_loci, batch.n_reads))


def prior_scenarios():
    """The six scenario files of tests/resources/prior/scenarios (every prior branch: population, Mendelian with
    sex/contig ploidy maps, clonal and subclonal inheritance, somatic rates), verbatim, keyed by file name."""
    d = os.path.join(REF, "prior", "scenarios")
    out = {name[:-len(".scenario.yaml")]: open(os.path.join(d, name)).read() for name in sorted(os.listdir(d))
           if name.endswith(".scenario.yaml")}
    with open(os.path.join(HERE, "prior_scenarios.json"), "w") as f:
        json.dump({"source": "tests/resources/prior/scenarios/*.scenario.yaml", "scenarios": out}, f, indent=1)
    print("prior scenarios: %s" % ", ".join(out))


def real_pileups():
    root = os.path.join(REF, "testcases")
    batches, meta = [], []
    for name in sorted(os.listdir(root)):
        cand = os.path.join(root, name, "candidates.vcf")
        scen = os.path.join(root, name, "scenario.yaml")
        if not (os.path.exists(cand) and os.path.exists(scen)):
            continue
        with open(cand, "rb") as f:
            raw = f.read(20000)
        if raw[:2] == b"\x1f\x8b" or raw[:3] == b"BCF":
            continue
        head = raw.decode("utf-8", "replace")
        if "##varlociraptor_observation_format_version=15" not in head:
            continue
        recs = [r for r in obs_codec.parse_observation_vcf(cand) if "PROB_MAPPING" in r["info"]]
        if not recs:
            continue
        try:
            b = obs_codec.batch_from_records([recs])
        except Exception as e:  # noqa: BLE001
            print("skip %s: %s" % (name, e))
            continue
        with open(scen) as f:
            scenario = f.read()
        # the expectations the reference's own test of this testcase checks on the call (testcase.yaml `expected:`,
        # evaluated by src/testcase/runner/common/mod.rs:330-395), verbatim with the comments next to them
        import yaml
        with open(os.path.join(root, name, "testcase.yaml")) as f:
            text = f.read()
        tc_yaml = yaml.safe_load(text)
        expected = tc_yaml.get("expected") or {}
        exp_lines = []
        for line in text.split("\n"):
            if line.startswith("expected:"):
                exp_lines = [line]
            elif exp_lines and (line.startswith(" ") or line.startswith("#")):
                exp_lines.append(line)
            elif exp_lines:
                break
        batches.append(b)
        meta.append({"testcase": name, "n_loci": b.n_loci, "n_reads": b.n_reads,
                     "expected": {"allelefreqs": expected.get("allelefreqs") or [],
                                  "posteriors": expected.get("posteriors") or [],
                                  "omit": sorted(k for k, v in tc_yaml.items() if k.startswith("omit_") and v),
                                  "yaml_text": "\n".join(exp_lines)},
                     "records": [{"chrom": r["chrom"], "pos": r["pos"], "ref": r["ref"], "alt": r["alt"]}
                                 for r in recs], "scenario_yaml": scenario})
        print("%s: %d loci, %d reads" % (name, b.n_loci, b.n_reads))
    LocusBatch.concat(batches).save(os.path.join(HERE, "real_pileups.npz"))
    with open(os.path.join(HERE, "real_pileups.json"), "w") as f:
        json.dump({"source": "tests/resources/testcases/*/candidates.vcf (observation format 15)",
                   "testcases": meta}, f, indent=1)


def raw_info_records():
    """Verbatim INFO integer arrays of a few reference records (bit-exact decode -> encode round trip test)."""
    recs = obs_codec.parse_observation_vcf(os.path.join(REF, "flamegraph_profiling", "normal.vcf"))[:2]
    extra = obs_codec.parse_observation_vcf(os.path.join(REF, "testcases", "test_false_negative_indel_call",
                                                         "candidates.vcf"))
    extra = [r for r in extra if "PROB_MAPPING" in r["info"]][:1]
    out = [{"chrom": r["chrom"], "pos": r["pos"], "ref": r["ref"], "alt": r["alt"],
            "info": {k: v for k, v in r["info"].items() if isinstance(v, list)}} for r in recs + extra]
    with open(os.path.join(HERE, "raw_info_records.json"), "w") as f:
        json.dump({"source": "raw INFO integer arrays of tests/resources/flamegraph_profiling/normal.vcf (2 records) "
                             "and testcases/test_false_negative_indel_call/candidates.vcf (1 record)",
                   "records": out}, f)


if __name__ == "__main__":
    flamegraph()
    prior_scenarios()
    real_pileups()
    raw_info_records()

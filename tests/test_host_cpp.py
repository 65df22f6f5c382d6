"""C++ host mirror (host/vlr_caller.hpp): compiled here with g++ against a recording mock engine (no GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_caller_with_mock_engine(tmp_path):
    exe = str(tmp_path / "test_caller_mock")
    src = os.path.join(ROOT, "tests", "host", "test_caller_mock.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe, src])
    out = subprocess.check_output([exe], text=True)
    assert "host caller mock test: ok" in out


def test_cpp_contamination_estimator_with_mock_model(tmp_path):
    """host/vlr_contamination.hpp: selection, CSR packing, prior, table and filter against a mock of the C-ABI entry."""
    exe = str(tmp_path / "test_contamination_mock")
    src = os.path.join(ROOT, "tests", "host", "test_contamination_mock.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe, src])
    assert "host contamination mock test: ok" in subprocess.check_output([exe], text=True)


def _bits(a, dtype):
    import numpy as np
    return ["%08x" % x for x in np.asarray(a, dtype=dtype).view(np.uint32).tolist()]


def test_cpp_observation_decoder_matches_the_python_codec(tmp_path, golden_dir):
    """host/vlr_obs_codec.hpp against varlociraptor_b200.obs_codec on the reference's own records (verbatim INFO arrays
    of two methylation records and one indel record), a synthetic record with homopolymer columns, and a truncated
    array."""
    import json
    import numpy as np
    from varlociraptor_b200 import abi, obs_codec
    from varlociraptor_b200.batch import mini_logprob
    exe = str(tmp_path / "test_obs_codec")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "test_obs_codec.cpp")])
    infos = [r["info"] for r in json.load(open(os.path.join(golden_dir, "raw_info_records.json")))["records"]]
    rng = np.random.Generator(np.random.PCG64(3))
    n = 37
    cols = {k: mini_logprob(-rng.exponential(8.0, n)) for k in abi.BATCH_F32_COLUMNS}
    cols["prob_double_overlap"][::3] = -np.inf
    flags = (rng.integers(0, 4, n) << abi.RF_STRAND_SHIFT) | (rng.choice([0, 1, 8, 5], n) << abi.RF_ORIENT_SHIFT) | \
        (rng.integers(0, 3, n) << abi.RF_ALTLOCUS_SHIFT) | (rng.integers(0, 2, n) * abi.RF_SOFTCLIPPED) | \
        (rng.integers(0, 2, n) * abi.RF_PAIRED) | (rng.integers(0, 2, n) * abi.RF_MAX_MAPQ) | \
        (rng.integers(0, 2, n) * abi.RF_READPOS_MAJOR)
    hl = rng.integers(-3, 4, n)
    has = rng.random(n) < 0.7
    flags = flags.astype(np.uint32) | np.where(has, abi.RF_HAS_HOMOPOLYMER_LEN | ((hl & 0xff) << abi.RF_HOMOPOLYMER_LEN_SHIFT), 0).astype(np.uint32)
    hart = np.where(has, mini_logprob(-rng.exponential(3.0, n)), np.nan).astype(np.float32)
    hvar = np.where(has, mini_logprob(-rng.exponential(3.0, n)), np.nan).astype(np.float32)
    infos.append(obs_codec.encode_record(cols, flags, hart, hvar))
    broken = dict(infos[0])
    broken["PROB_ALT"] = broken["PROB_ALT"][:-3]
    infos.append(broken)
    path = tmp_path / "records.txt"
    with open(path, "w") as f:
        for info in infos:
            for tag, vals in info.items():
                if isinstance(vals, list):
                    f.write(tag + " " + " ".join(str(v) for v in vals) + "\n")
            f.write("--\n")
    blocks = subprocess.check_output([exe, str(path)], text=True).strip().split("--\n")
    blocks = [b for b in (x.strip() for x in blocks) if b]
    assert len(blocks) == len(infos)
    for info, block in zip(infos[:-1], blocks):
        got = {ln.split()[0]: ln.split()[1:] for ln in block.splitlines()}
        want_cols, want_flags, want_hart, want_hvar = obs_codec.decode_record(info)
        for k in abi.BATCH_F32_COLUMNS:
            assert got[k] == _bits(want_cols[k], np.float32), k
        assert got["read_flags"] == _bits(want_flags, np.uint32)
        assert got["hart"] == ([] if want_hart is None else _bits(want_hart, np.float32))
        assert got["hvar"] == ([] if want_hvar is None else _bits(want_hvar, np.float32))
        if "THIRD_ALLELE_EVIDENCE" in info:
            h, v = obs_codec.decode_optional_u32(info["THIRD_ALLELE_EVIDENCE"])
            assert got["third"] == [str(int(x)) if y else "." for x, y in zip(v, h)]
    assert blocks[-1].startswith("ERROR truncated observation INFO array")


def _build_contamination_link(tmp_path):
    from varlociraptor_b200 import build as vbuild
    lib = vbuild.build()
    exe = str(tmp_path / "test_contamination_link")
    src = os.path.join(ROOT, "tests", "host", "test_contamination_link.cpp")
    libdir = os.path.dirname(lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe, src, "-L" + libdir,
                           "-lvlr_engine", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_contamination_estimator_refuses_without_device(tmp_path):
    """The real library behind the C++ estimator: no device, no result (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = subprocess.check_output([_build_contamination_link(tmp_path)], text=True)
    assert "refused: vlr_contamination_posterior failed with status 5" in out


import numpy as np
import pytest


@pytest.mark.gpu
def test_cpp_contamination_estimator_on_gpu(tmp_path):
    out = subprocess.check_output([_build_contamination_link(tmp_path)], text=True)
    assert "posterior integrates to 1.0000000" in out or "posterior integrates to 0.9999999" in out


@pytest.mark.gpu
def test_cpp_call_generic_on_gpu_matches_oracle(tmp_path):
    """The C++ `call_generic` over the real engine library == oracle on the same records."""
    from oracle import oracle
    from varlociraptor_b200 import build as vbuild, synth
    lib = vbuild.build()
    sc, b = synth.tumor_normal(10, seed=41, depth=24)
    flat = sc.flatten()
    for name, arr in (("samples", flat._samples), ("events", flat._events), ("nodes", flat._nodes),
                      ("set_vafs", flat._set_vafs), ("spectra", flat._spectra)):
        n = {"samples": flat.c.n_samples, "events": flat.c.n_events, "nodes": flat.c.n_nodes,
             "set_vafs": flat.c.n_set_vafs, "spectra": flat.c.n_spectra}[name]
        raw = bytes(arr)
        (tmp_path / (name + ".bin")).write_bytes(raw[: len(raw) // len(arr) * n])
    from varlociraptor_b200 import abi
    for s, name in enumerate(["normal", "tumor"]):
        out = []
        for i in range(b.n_loci):
            lo, hi = b.read_offsets[i * 2 + s], b.read_offsets[i * 2 + s + 1]
            out.append(np.array([hi - lo], dtype=np.float32))
            for k in abi.BATCH_F32_COLUMNS:
                out.append(b.columns[k][lo:hi])
            out.append(b.read_flags[lo:hi].astype(np.float32))
        (tmp_path / (name + ".bin")).write_bytes(np.concatenate(out).astype(np.float32).tobytes())
    exe = str(tmp_path / "test_caller_gpu")
    libdir = os.path.dirname(lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host", "test_caller_gpu.cpp"),
                           "-L" + libdir, "-lvlr_engine", "-Wl,-rpath," + libdir])
    text = subprocess.check_output([exe, str(tmp_path)], text=True)
    want = oracle.call_batch(flat, b, afd_capacity=128)
    lines = [ln.split() for ln in text.splitlines() if ln.startswith("CALL")]
    assert len(lines) == b.n_loci
    E = flat.n_events
    for i, t in enumerate(lines):
        assert int(t[1]) == i and int(t[2]) == int(want.status[i])
        got = np.array([float(x) for x in t[3:3 + E + 1]])
        ref = want.log_posteriors[i]
        same = (got == ref) | (np.abs(got - ref) <= 1e-9)
        assert same.all(), (i, got, ref)
        assert float(t[3 + E + 1]) == want.map_vaf[i, 0] and float(t[3 + E + 3]) == want.map_vaf[i, 1]
        assert int(t[3 + E + 2]) == want.afd_count[i, 0] and int(t[3 + E + 4]) == want.afd_count[i, 1]

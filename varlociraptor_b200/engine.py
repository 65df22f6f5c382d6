"""Host-side handle of the CUDA posterior engine (ctypes over the C-ABI in include/vlr_engine.h).

`PosteriorEngine` plays the role of the configured `bio::stats::bayesian::Model` that `Caller` holds in the
reference (src/calling/variants/calling.rs:48-49, 632-718): built once per contig from a scenario, then fed
batches of records. There is deliberately no CPU path: if the CUDA library is missing or no device is usable,
construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import abi
from .batch import CallResults, LocusBatch

# The host entries keep eight chunks in flight on eight streams, next to the context's own ones. With the default of 8
# hardware connections some of those streams share a queue and a chunk waits for the whole chunk in front of it on
# that queue (seen in the VLR_CHUNK_TIMING timeline: chunks 4..7 started when chunks 0..3 had finished). Has to be set
# before the process creates its CUDA context; a host application sets it in its environment.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_LIB_PATH = os.environ.get("VLR_ENGINE_LIB",  # developer override for tuning experiments (another CUDA build)
                           os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libvlr_engine.so"))
_lib = None


class EngineError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise EngineError("CUDA engine library not built: %s (run `python -m varlociraptor_b200.build`); "
                              "there is no CPU fallback" % _LIB_PATH)
        l = C.CDLL(_LIB_PATH)
        l.vlr_abi_version.restype = C.c_int32
        l.vlr_ctx_create.restype = C.c_int32
        l.vlr_ctx_create.argtypes = [C.POINTER(abi.Scenario), C.c_int32, C.POINTER(C.c_void_p)]
        l.vlr_ctx_destroy.restype = None
        l.vlr_ctx_destroy.argtypes = [C.c_void_p]
        l.vlr_call_batch.restype = C.c_int32
        l.vlr_call_batch.argtypes = [C.c_void_p, C.POINTER(abi.Batch), C.POINTER(abi.Results)]
        l.vlr_call_batch_device.restype = C.c_int32
        l.vlr_call_batch_device.argtypes = [C.c_void_p, C.POINTER(abi.Batch), C.POINTER(abi.Results), C.c_void_p]
        l.vlr_ctx_reserve.restype = C.c_int32
        l.vlr_ctx_reserve.argtypes = [C.c_void_p, C.c_int64]
        l.vlr_host_alloc.restype = C.c_void_p
        l.vlr_host_alloc.argtypes = [C.c_size_t]
        l.vlr_host_free.restype = None
        l.vlr_host_free.argtypes = [C.c_void_p]
        l.vlr_last_launch_count.restype = C.c_int64
        l.vlr_last_launch_count.argtypes = [C.c_void_p]
        l.vlr_ctx_stream.restype = C.c_void_p
        l.vlr_ctx_stream.argtypes = [C.c_void_p]
        l.vlr_last_error.restype = C.c_char_p
        l.vlr_last_error.argtypes = [C.c_void_p]
        l.vlr_status_string.restype = C.c_char_p
        l.vlr_status_string.argtypes = [C.c_int32]
        l.vlr_measure_fp64_peak.restype = C.c_int32
        l.vlr_measure_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
        l.vlr_contamination_posterior.restype = C.c_int32
        l.vlr_contamination_posterior.argtypes = [C.c_int32, C.POINTER(abi.ContaminationInput),
                                                  C.POINTER(abi.ContaminationOutput)]
        l.vlr_contamination_posterior_device.restype = C.c_int32
        l.vlr_contamination_posterior_device.argtypes = [C.c_int32, C.POINTER(abi.ContaminationInput),
                                                         C.POINTER(abi.ContaminationOutput), C.c_void_p]
        l.vlr_contamination_gather_device.restype = C.c_int32
        l.vlr_contamination_gather_device.argtypes = [C.c_int32, C.POINTER(abi.Results), C.c_int64, C.c_int32, C.c_int32,
                                                      C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 6 + \
            [C.POINTER(C.c_int64), C.c_void_p]
        l.vlr_pack_batch.restype = C.c_int32
        l.vlr_pack_batch.argtypes = [C.POINTER(abi.Batch), C.c_int32, C.POINTER(C.POINTER(abi.PackedBatch))]
        l.vlr_packed_batch_free.restype = None
        l.vlr_packed_batch_free.argtypes = [C.POINTER(abi.PackedBatch)]
        l.vlr_packed_batch_bytes.restype = C.c_int64
        l.vlr_packed_batch_bytes.argtypes = [C.POINTER(abi.PackedBatch), C.c_int32]
        l.vlr_call_batch_packed.restype = C.c_int32
        l.vlr_call_batch_packed.argtypes = [C.c_void_p, C.POINTER(abi.PackedBatch), C.POINTER(abi.Results)]
        if l.vlr_abi_version() != abi.VLR_ABI_VERSION:
            raise EngineError("ABI version mismatch between abi.py and libvlr_engine.so")
        _lib = l
    return _lib


EXPORTED_SYMBOLS = ["vlr_ctx_create", "vlr_ctx_destroy", "vlr_call_batch", "vlr_call_batch_device", "vlr_ctx_reserve",
                    "vlr_host_alloc", "vlr_host_free", "vlr_last_launch_count", "vlr_ctx_stream", "vlr_last_error",
                    "vlr_status_string", "vlr_abi_version", "vlr_measure_fp64_peak", "vlr_contamination_posterior",
                    "vlr_contamination_posterior_device", "vlr_contamination_gather_device", "vlr_pack_batch", "vlr_packed_batch_free",
                    "vlr_packed_batch_bytes", "vlr_call_batch_packed"]


def measure_fp64_peak(device: int = 0) -> float:
    """Measured fp64 FMA throughput of the device in TFLOP/s (register-resident DFMA microbenchmark)."""
    out = C.c_double(0.0)
    rc = lib().vlr_measure_fp64_peak(device, C.byref(out))
    if rc != 0:
        raise EngineError("vlr_measure_fp64_peak failed: %s" % lib().vlr_status_string(rc).decode())
    return float(out.value)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked memory from vlr_host_alloc; freed when the last view dies."""
    import weakref
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(1, n * dtype.itemsize)
    ptr = lib().vlr_host_alloc(nbytes)
    if not ptr:
        raise EngineError("vlr_host_alloc(%d) failed" % nbytes)
    buf = (C.c_char * nbytes).from_address(ptr)
    weakref.finalize(buf, lib().vlr_host_free, ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pin_batch(batch: LocusBatch) -> LocusBatch:
    """Copy of a batch whose columns live in page-locked memory."""
    def pin(a):
        if a is None:
            return None
        out = pinned_empty(a.shape, a.dtype)
        out[...] = a
        return out
    nb = LocusBatch.__new__(LocusBatch)
    nb.n_samples, nb.n_loci, nb.n_reads = batch.n_samples, batch.n_loci, batch.n_reads
    nb.read_offsets = pin(batch.read_offsets)
    nb.columns = {k: pin(v) for k, v in batch.columns.items()}
    nb.read_flags = pin(batch.read_flags)
    nb.locus_flags = pin(batch.locus_flags)
    nb.prob_homopolymer_artifact = pin(batch.prob_homopolymer_artifact)
    nb.prob_homopolymer_variant = pin(batch.prob_homopolymer_variant)
    nb.locus_heterozygosity_phred = pin(batch.locus_heterozygosity_phred)
    nb.locus_semr_phred = pin(batch.locus_semr_phred)
    return nb


def pinned_results(n_loci, n_samples, n_events, afd_capacity=0) -> CallResults:
    r = CallResults.__new__(CallResults)
    r.n_loci, r.n_samples, r.n_events, r.afd_capacity = n_loci, n_samples, n_events, afd_capacity
    r.log_posteriors = pinned_empty((n_loci, n_events + 1), np.float64)
    r.log_marginal = pinned_empty(n_loci, np.float64)
    r.map_vaf = pinned_empty((n_loci, n_samples), np.float64)
    r.map_config = pinned_empty(n_loci, np.int32)
    r.best_event = pinned_empty(n_loci, np.int32)
    r.status = pinned_empty(n_loci, np.uint32)
    r.n_base_events = pinned_empty(n_loci, np.uint32)
    if afd_capacity > 0:
        r.afd_count = pinned_empty((n_loci, n_samples), np.int32)
        r.afd_vaf = pinned_empty((n_loci, n_samples, afd_capacity), np.float64)
        r.afd_logp = pinned_empty((n_loci, n_samples, afd_capacity), np.float64)
    else:
        r.afd_count = r.afd_vaf = r.afd_logp = None
    return r


class PackedBatch:
    """A LocusBatch losslessly encoded for the host -> device link (vlr_pack_batch): every column in the smallest of
    f32 / f16 / 16-bit dictionary / 8-bit dictionary / constant that reproduces its f32 bit patterns. Keeps the source
    batch alive (offsets and locus columns are borrowed)."""

    def __init__(self, batch: LocusBatch, n_threads: int = 0):
        self.batch = batch
        self.n_loci, self.n_reads, self.n_samples = batch.n_loci, batch.n_reads, batch.n_samples
        self._cb = batch.as_c()
        self._p = C.POINTER(abi.PackedBatch)()
        rc = lib().vlr_pack_batch(C.byref(self._cb), int(n_threads), C.byref(self._p))
        if rc != 0:
            self._p = C.POINTER(abi.PackedBatch)()
            raise EngineError("vlr_pack_batch failed: %s" % lib().vlr_status_string(rc).decode())

    @property
    def c(self):
        return self._p

    @property
    def encodings(self) -> dict:
        """column -> (encoding name, dictionary size)"""
        cols = self._p.contents.columns
        return {name: (abi.ENC_NAMES[cols[i].encoding], int(cols[i].n_dict)) for i, name in enumerate(abi.PACKED_COLUMNS)}

    def nbytes(self) -> int:
        """Bytes vlr_call_batch_packed moves host -> device."""
        return int(lib().vlr_packed_batch_bytes(self._p, self.n_samples))

    def decode(self, column: str) -> np.ndarray:
        """The column widened on the host (tests): uint32 bit patterns."""
        col = self._p.contents.columns[abi.PACKED_COLUMNS.index(column)]
        n = self.n_reads
        d = np.ctypeslib.as_array(col.dict, shape=(col.n_dict,)).copy() if col.n_dict else None
        if col.encoding == abi.ENC_CONST:
            return np.full(n, d[0], dtype=np.uint32)
        dt = {abi.ENC_F32: np.uint32, abi.ENC_F16: np.float16, abi.ENC_DICT16: np.uint16, abi.ENC_DICT8: np.uint8}[col.encoding]
        raw = np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(col.data), dtype=dt, count=n)
        if col.encoding == abi.ENC_F32:
            return raw.copy()
        if col.encoding == abi.ENC_F16:
            return raw.astype(np.float32).view(np.uint32)
        return d[raw]

    def close(self):
        if self._p:
            lib().vlr_packed_batch_free(self._p)
            self._p = C.POINTER(abi.PackedBatch)()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class PosteriorEngine:
    def __init__(self, flat_scenario, device: int = 0):
        self.flat = flat_scenario
        self.n_samples = flat_scenario.n_samples
        self.n_events = flat_scenario.n_events
        self._ctx = C.c_void_p()
        rc = lib().vlr_ctx_create(C.byref(flat_scenario.c), device, C.byref(self._ctx))
        if rc != 0:
            self._ctx = C.c_void_p()
            raise EngineError("vlr_ctx_create failed: %s" % lib().vlr_status_string(rc).decode())
        self.device = device

    def close(self):
        if self._ctx:
            lib().vlr_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError("%s: %s" % (lib().vlr_status_string(rc).decode(),
                                          (lib().vlr_last_error(self._ctx) or b"").decode()))

    @property
    def launches(self) -> int:
        return int(lib().vlr_last_launch_count(self._ctx))

    @property
    def stream(self) -> int:
        return int(lib().vlr_ctx_stream(self._ctx) or 0)

    def reserve(self, max_reads_per_locus: int):
        self._check(lib().vlr_ctx_reserve(self._ctx, int(max_reads_per_locus)))

    def call_batch(self, batch: LocusBatch, afd_capacity: int = 0, out: Optional[CallResults] = None) -> CallResults:
        """Host buffers in, host buffers out (H2D, kernels, D2H inside)."""
        assert batch.n_samples == self.n_samples
        if out is None:
            out = CallResults(batch.n_loci, self.n_samples, self.n_events, afd_capacity)
        cb, cr = batch.as_c(), out.as_c()
        self._check(lib().vlr_call_batch(self._ctx, C.byref(cb), C.byref(cr)))
        return out

    def call_batch_packed(self, packed: PackedBatch, afd_capacity: int = 0, out: Optional[CallResults] = None) -> CallResults:
        """call_batch on a PackedBatch: 9..20 instead of 32 bytes per read over the link, results bit for bit the same."""
        assert packed.n_samples == self.n_samples
        if out is None:
            out = CallResults(packed.n_loci, self.n_samples, self.n_events, afd_capacity)
        cr = out.as_c()
        self._check(lib().vlr_call_batch_packed(self._ctx, packed.c, C.byref(cr)))
        return out

    def call_batch_device(self, dev_batch: "DeviceBatch", dev_results: "DeviceResults", stream: int = 0):
        """Device-resident buffers; asynchronous on `stream` (0 = the engine's own stream)."""
        cb, cr = dev_batch.as_c(), dev_results.as_c()
        self._check(lib().vlr_call_batch_device(self._ctx, C.byref(cb), C.byref(cr), C.c_void_p(stream or None)))


class DeviceBatch:
    """A LocusBatch resident in device memory (torch tensors are only the allocator)."""

    def __init__(self, batch: LocusBatch, device="cuda:0"):
        import torch
        self.n_loci, self.n_reads, self.n_samples = batch.n_loci, batch.n_reads, batch.n_samples

        def up(a):
            return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.read_offsets = up(batch.read_offsets)
        self.columns = {k: up(v) for k, v in batch.columns.items()}
        self.read_flags = up(batch.read_flags.view(np.int32))
        self.locus_flags = up(batch.locus_flags.view(np.int32))
        self.hart = up(batch.prob_homopolymer_artifact)
        self.hvar = up(batch.prob_homopolymer_variant)
        self.het = up(batch.locus_heterozygosity_phred)
        self.semr = up(batch.locus_semr_phred)

    def as_c(self) -> abi.Batch:
        def p(t, ct):
            return abi.devptr(0 if t is None or t.numel() == 0 else t.data_ptr(), ct)
        b = abi.Batch()
        b.n_loci, b.n_reads = self.n_loci, self.n_reads
        b.read_offsets = p(self.read_offsets, C.c_int64)
        for k in abi.BATCH_F32_COLUMNS:
            setattr(b, k, p(self.columns[k], C.c_float))
        b.read_flags = p(self.read_flags, C.c_uint32)
        b.prob_homopolymer_artifact = p(self.hart, C.c_float)
        b.prob_homopolymer_variant = p(self.hvar, C.c_float)
        b.locus_flags = p(self.locus_flags, C.c_uint32)
        b.locus_heterozygosity_phred = p(self.het, C.c_float)
        b.locus_semr_phred = p(self.semr, C.c_float)
        return b


class DeviceResults:
    def __init__(self, n_loci, n_samples, n_events, afd_capacity=0, device="cuda:0"):
        import torch
        self.n_loci, self.n_samples, self.n_events, self.afd_capacity = n_loci, n_samples, n_events, afd_capacity
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.log_posteriors = torch.empty((n_loci, n_events + 1), **f64)
        self.log_marginal = torch.empty(n_loci, **f64)
        self.map_vaf = torch.empty((n_loci, n_samples), **f64)
        self.map_config = torch.empty(n_loci, **i32)
        self.best_event = torch.empty(n_loci, **i32)
        self.status = torch.empty(n_loci, **i32)
        self.n_base_events = torch.empty(n_loci, **i32)
        if afd_capacity > 0:
            self.afd_count = torch.empty((n_loci, n_samples), **i32)
            self.afd_vaf = torch.empty((n_loci, n_samples, afd_capacity), **f64)
            self.afd_logp = torch.empty((n_loci, n_samples, afd_capacity), **f64)
        else:
            self.afd_count = self.afd_vaf = self.afd_logp = None

    def as_c(self) -> abi.Results:
        def p(t, ct):
            return abi.devptr(0 if t is None else t.data_ptr(), ct)
        r = abi.Results()
        r.log_posteriors = p(self.log_posteriors, C.c_double)
        r.log_marginal = p(self.log_marginal, C.c_double)
        r.map_vaf = p(self.map_vaf, C.c_double)
        r.map_config = p(self.map_config, C.c_int32)
        r.best_event = p(self.best_event, C.c_int32)
        r.status = p(self.status, C.c_uint32)
        r.n_base_events = p(self.n_base_events, C.c_uint32)
        r.afd_capacity = self.afd_capacity
        r.afd_count = p(self.afd_count, C.c_int32)
        r.afd_vaf = p(self.afd_vaf, C.c_double)
        r.afd_logp = p(self.afd_logp, C.c_double)
        return r

    def to_host(self) -> CallResults:
        out = CallResults(self.n_loci, self.n_samples, self.n_events, self.afd_capacity)
        out.log_posteriors[...] = self.log_posteriors.cpu().numpy()
        out.log_marginal[...] = self.log_marginal.cpu().numpy()
        out.map_vaf[...] = self.map_vaf.cpu().numpy()
        out.map_config[...] = self.map_config.cpu().numpy()
        out.best_event[...] = self.best_event.cpu().numpy()
        out.status[...] = self.status.cpu().numpy().view(np.uint32)
        out.n_base_events[...] = self.n_base_events.cpu().numpy().view(np.uint32)
        if self.afd_capacity > 0:
            out.afd_count[...] = self.afd_count.cpu().numpy()
            out.afd_vaf[...] = self.afd_vaf.cpu().numpy()
            out.afd_logp[...] = self.afd_logp.cpu().numpy()
        return out

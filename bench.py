#!/usr/bin/env python
"""bench.py — loci/s of the per-locus posterior engine on BASELINE.json's config 2
(1M synthetic SNV loci, tumor-normal, 100 reads/locus/sample, tumor resolution 0.01 = "101-pt grid").

One "step" = one pass of the hot path over the whole locus batch.
  value     whole-job loci/s with the batch already resident in HBM (CUDA events on the launch stream)
  e2e       the same metric through the C-ABI entry `vlr_call_batch` with pinned HOST buffers
            (chunked H2D, kernels and D2H all inside the timed region)
  roofline  algorithmic HBM bytes of the kernel / its measured duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample of the same workload

`--impl reference` times the reference algorithm's CPU implementation (the oracle port; the Rust reference cannot be
built in this image) with all host threads on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): loci are independent, every rank processes its own shard of the same size
(weak scaling) and the only collective is the final NCCL gather of the fixed-stride result records.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loci_per_sec"
UNIT = "loci/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loci", type=int, default=1_000_000, help="loci per GPU (config 2: 1M)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5], help="BASELINE.json config index")
    ap.add_argument("--cpu-sample", type=int, default=0, help="loci in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(cfg, loci):
    return {2: "cfg2: %d synthetic SNV loci, tumor-normal (purity 0.75), 100 reads/locus/sample, tumor resolution 0.01",
            3: "cfg3: %d synthetic mixed SNV/indel loci, 3-sample pedigree grammar, 100 reads/locus/sample",
            5: "cfg5: %d synthetic SNV loci, tumor-normal, reads/locus/sample log-uniform 10..2000"}[cfg] % loci


def make_batch(cfg, loci, seed, pinned=False):
    """Synthetic batch of `loci` loci, generated in slabs (bounded transient memory). With pinned=True the columns are
    written straight into page-locked arrays (vlr_host_alloc), so a rank holds one copy of its 6.4 GB shard."""
    from varlociraptor_b200 import synth
    from varlociraptor_b200.batch import LocusBatch
    slab = 125_000 if cfg != 5 else 20_000
    parts = []
    scenario = None
    done = 0
    k = 0
    out = None
    row = 0
    while done < loci:
        n = min(slab, loci - done)
        scenario, b = synth.config(cfg, n, seed=seed * 1000 + k)
        if pinned and cfg != 5:
            S = b.n_samples
            if out is None:  # fixed depth: total sizes are known after the first slab
                from varlociraptor_b200 import engine
                total_reads = b.n_reads // n * loci
                out = LocusBatch.__new__(LocusBatch)
                out.n_samples, out.n_loci, out.n_reads = S, loci, total_reads
                out.read_offsets = engine.pinned_empty(loci * S + 1, np.int64)
                out.read_offsets[0] = 0
                out.columns = {c: engine.pinned_empty(total_reads, np.float32) for c in b.columns}
                out.read_flags = engine.pinned_empty(total_reads, np.uint32)
                out.locus_flags = engine.pinned_empty(loci, np.uint32)
                out.prob_homopolymer_artifact = out.prob_homopolymer_variant = None
                out.locus_heterozygosity_phred = out.locus_semr_phred = None
            out.read_offsets[done * S + 1:(done + n) * S + 1] = b.read_offsets[1:] + row
            for c in b.columns:
                out.columns[c][row:row + b.n_reads] = b.columns[c]
            out.read_flags[row:row + b.n_reads] = b.read_flags
            out.locus_flags[done:done + n] = b.locus_flags
            row += b.n_reads
        else:
            parts.append(b)
        done += n
        k += 1
    if out is not None:
        assert row == out.n_reads
        return scenario, out
    batch = parts[0] if len(parts) == 1 else LocusBatch.concat(parts)
    if pinned:
        from varlociraptor_b200 import engine
        batch = engine.pin_batch(batch)
    return scenario, batch


def algorithmic_bytes(batch, n_events):
    """32 B per read (7 f32 probabilities + flag word) + 16 B locus header in,
    8 (E+1) + 9 S bytes out per locus."""
    S = batch.n_samples
    return batch.n_reads * 32 + batch.n_loci * (16 + 8 * (n_events + 1) + 9 * S)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(flat, batch, n_sample, threads):
    from oracle import oracle
    n = min(n_sample, batch.n_loci)
    sub = batch.slice(0, n)
    t0 = time.perf_counter()
    oracle.call_batch(flat, sub, afd_capacity=0, n_threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.cpu_sample or 250 * threads
    scenario, batch = make_batch(args.config, n_sample, seed=20260100 + args.config)
    flat = scenario.flatten()
    from oracle import oracle
    oracle.build()
    vals = []
    for i in range(args.warmup + args.steps):
        v, n, dt = cpu_baseline(flat, batch, n_sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
        if i == 0 and dt * (args.warmup + args.steps) > 240:  # keep the whole run within minutes
            args.warmup, args.steps = 0, max(1, min(args.steps, int(240 / dt)))
    value = float(np.mean([v for v, _ in vals])) if vals else v
    ms = float(np.mean([dt for _, dt in vals]) * 1e3) if vals else dt * 1e3
    sample = "first %d loci of the workload per step, %d threads over disjoint locus ranges" % (batch.n_loci, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.loci), "sample_loci_per_step": int(batch.n_loci)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle = C++ port of the reference algorithm (the Rust reference cannot be built here: no cargo); "
                "upstream `call variants` is single-threaded, the port is run on all host threads",
    }))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device(device))
    from varlociraptor_b200 import engine
    scenario, batch = make_batch(args.config, args.loci, seed=20260100 + args.config + 17 * rank, pinned=True)
    flat = scenario.flatten()
    S, E = flat.n_samples, flat.n_events
    eng = engine.PosteriorEngine(flat, device=local_rank)
    max_reads = int(np.max(batch.read_offsets[S::S] - batch.read_offsets[:-S:S]))
    eng.reserve(max_reads)
    dbatch = engine.DeviceBatch(batch, device)
    dres = engine.DeviceResults(batch.n_loci, S, E, 0, device)
    pinned = batch  # already page-locked
    pres = engine.pinned_results(batch.n_loci, S, E, 0)
    gather_buf = None
    if world > 1:
        rec = torch.empty((batch.n_loci, E + 1 + S + 1), dtype=torch.float64, device=device)
        gather_buf = [torch.empty_like(rec) for _ in range(world)] if rank == 0 else None
    stream = torch.cuda.Stream(device=device)  # a real (non-default) stream: handle 0 would mean "the engine's own"
    torch.cuda.set_stream(stream)

    def step_device():
        eng.call_batch_device(dbatch, dres, stream.cuda_stream)
        if world > 1:  # final gather of fixed-stride result records over NVLink
            rec[:, :E + 1] = dres.log_posteriors
            rec[:, E + 1:E + 1 + S] = dres.map_vaf
            rec[:, E + 1 + S] = dres.status.to(torch.float64)
            dist.gather(rec, gather_buf, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        kev[i][0].record(stream)
        eng.call_batch_device(dbatch, dres, stream.cuda_stream)
        kev[i][1].record(stream)
        if world > 1:
            rec[:, :E + 1] = dres.log_posteriors
            rec[:, E + 1:E + 1 + S] = dres.map_vaf
            rec[:, E + 1 + S] = dres.status.to(torch.float64)
            dist.gather(rec, gather_buf, dst=0)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    launches = args.steps * eng.launches  # kernels of one device-entry call (vlr_last_launch_count) x timed steps

    # end to end through the host-buffer entry of the C-ABI
    for _ in range(1):
        eng.call_batch(pinned, out=pres)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.call_batch(pinned, out=pres)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = eng.launches

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    status = dres.status.cpu().numpy().view(np.uint32)
    n_bad = int(np.count_nonzero(status & 0x83f))  # NaN/overflow/overshoot/workspace bits
    post = dres.log_posteriors.cpu().numpy()
    sum_err = float(np.nanmax(np.abs(np.logaddexp.reduce(post, axis=1))))
    joint_evals = float(dres.n_base_events.cpu().numpy().view(np.uint32).mean())

    if rank == 0:
        total_loci = batch.n_loci * world
        value = total_loci * args.steps / (ms_total * 1e-3)
        e2e = total_loci * args.steps / (e2e_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        abytes = algorithmic_bytes(batch, E)
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        wave = os.environ.get("VLR_WAVE", "1") != "0" and args.config in (2, 5)
        traffic = None  # DRAM bytes per step from the committed ncu capture of this workload, scaled by loci
        kernel_name = "vlr_call_kernel (warp per locus)"
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1_wave.json" if wave else "traffic_r1.json")))
            if args.config == 2:
                traffic = int(tr["dram_bytes_per_locus"] * batch.n_loci)
            if wave:
                kernel_name = tr["kernels"]
        except (OSError, KeyError, ValueError):
            pass
        # the bound that matters: fp64. Algorithmic flops = joint evaluations x reads of the integrated
        # pileup x 5 (2 FMA + 1 MUL per read and abscissa); peak = DFMA microbenchmark on this device.
        reads_leaf = batch.n_reads / max(1, batch.n_loci) / S
        flops = joint_evals * batch.n_loci * reads_leaf * 5.0
        try:
            fp64_peak = engine.measure_fp64_peak(local_rank)
        except Exception:  # noqa: BLE001
            fp64_peak = None
        fp64 = {"achieved": flops / (kernel_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (flops / (kernel_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                "flops_per_step": flops, "peak_source": "vlr_measure_fp64_peak (DFMA microbenchmark, this device)",
                "accounting": "joint evaluations x reads of the integrated pileup x 5 flops"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, args.loci), "loci_per_gpu": int(batch.n_loci),
                       "reads_per_gpu": int(batch.n_reads), "parallelism": "loci sharded, %d rank(s)" % world,
                       "l2": "inputs (%.1f GB per step) are larger than L2" % (abytes / 1e9)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(batch.nbytes()),
                    "d2h_bytes_per_step": int(batch.n_loci * (8 * (E + 1) + 8 * S + 8 + 4 * 4)),
                    "launches_per_step": int(e2e_launches)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kernel_name, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": int(abytes), "peak_source": peak_src, "fp64": fp64,
                         "note": "kernel_ms = all kernels of one vlr_call_batch_device step (CUDA events on the launch "
                                 "stream). The path is compute/latency bound, not HBM bound (DESIGN.md §4): %.0f joint "
                                 "evaluations (each a product over the reads of a pileup) per locus on average, so the "
                                 "HBM fraction is small by construction and `fp64` is the meaningful roofline; traffic "
                                 "= measured DRAM bytes of the committed ncu pass (profiles/), mostly the per-config "
                                 "coefficient arena (written once, re-staged every round), not the %.1f GB of "
                                 "algorithmic bytes" % (joint_evals, abytes / 1e9)},
            "clocks": clocks,
            "checks": {"loci_with_error_status": n_bad, "max_abs_log_sum_of_posteriors": sum_err},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle
            oracle.build()
            threads = os.cpu_count() or 1
            n_sample = args.cpu_sample or 250 * threads
            v, n, dt = cpu_baseline(flat, batch, n_sample, threads)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "first %d loci of the same batch, %.1f s, %d threads over disjoint "
                                             "locus ranges (upstream is single-threaded)" % (n, dt, threads)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

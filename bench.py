#!/usr/bin/env python
"""bench.py -- throughput of the Shannon k-mer front end on B200 (BASELINE.json metric:
reads partitioned/sec & k-mer lookups/sec; HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm
    python bench.py --config 4|5 ...                         # BASELINE.json configs[3] / configs[4]

A "step" is ONE pass of the whole hot path over the workload: K1-mer table build
(load_kmers + lowComplexity), seed ordering, greedy walks, shape + duplicate filters, contig
C-mer graph + components, K1-mer->component map, read packing and read->component partition.
Workload (default --config 3): BASELINE.json configs[2] -- synthetic 10 M 100-bp read pairs from
5 k transcripts, 1 % substitution error, K=24 (generated on the device, RC-doubled like
shannon.py:413-424 and counted with the jellyfish stand-in, all outside the timed region).

N = 1: the single-GPU path.  `value` = read records partitioned per second with inputs resident
in HBM; `e2e` = the same step through the C-ABI with HOST buffers (H2D of the K1-mer list and
the ASCII reads and D2H of the partition inside the timed region).

N > 1: ONE global workload (the same total size, STRONG scaling) on hash-sharded K1-mer tables
(shannon_b200/sharded.py): every rank holds a slice of the lines of the global k1mer.dict_org and
a range of the read records; lines are routed to minimizer-hash owners with NCCL all-to-all, global
K1-mer graph components are labelled across ranks, whole components are re-sharded, walks run per
rank, candidates are merged, reads are partitioned by range.  `dist_parity` = the N-rank result
(contigs, allowed set, contig graph, read partition) equals the 1-GPU result on the same input,
checked on rank 0 whenever the workload fits one GPU.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_partitioned_per_sec"
UNIT = "read records/s"
K = 24
K1 = K + 1
READ_LEN = 100
FRAG_LEN = 300
NVLINK_GBS = 900.0       # NVLink 5 per direction per GPU (B200_PROFILING.md)

# BASELINE.json configs[2..4]
CONFIGS = {
    3: {"pairs": 10_000_000, "transcripts": 5000, "skewed": False,
        "name": "BASELINE.json configs[2]"},
    4: {"pairs": 50_000_000, "transcripts": 20000, "skewed": True,
        "name": "BASELINE.json configs[3]"},
    5: {"pairs": 200_000_000, "transcripts": 60000, "skewed": False,   # BASELINE.json names no skew here
        "name": "BASELINE.json configs[4]"},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=None, help="read pairs of the whole workload")
    ap.add_argument("--transcripts", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--K", type=int, default=24, help="k-mer size (K1 = K+1; 32 -> 128-bit keys)")
    ap.add_argument("--sample-pairs", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-e2e", action="store_true")
    ap.add_argument("--no-dist-parity", action="store_true")
    ap.add_argument("--file-pairs", type=int, default=1_000_000, help="size of the file-level e2e leg")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.pairs is None:
        args.pairs = cfg["pairs"]
    if args.transcripts is None:
        args.transcripts = cfg["transcripts"]
    if args.seed is None:
        args.seed = 1234 + (args.config - 1)
    args.skewed = cfg["skewed"]
    return args


# ------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def device_transcripts(ctx, n_tx, seed, skewed):
    from shannon_b200 import synth
    tx = synth.make_transcripts(n_tx, seed)
    codes, offs = synth.pack_transcripts(tx)
    thr = synth.expression_thresholds(len(tx), [len(t) for t in tx], skewed)
    return len(tx), (ctx.to_device(codes), ctx.to_device(offs), ctx.to_device(thr))


def generate_records(ctx, tx, n_tx, n_pairs, seed, lo, hi):
    """Record pairs [lo, hi) of the RC-doubled read files of the whole workload:
    reads_1 = [R1 ; rc(R2)], reads_2 = [rc(R1) ; R2]   (shannon.py:413-424), as two device ASCII
    arrays.  Pair p of the generator is record p (first half) and record n_pairs + p (second)."""
    from shannon_b200 import synth
    d_tx, d_off, d_thr = tx
    n = hi - lo
    nb = max(n * READ_LEN, 1)
    d_r1, d_r2 = ctx.dev_alloc(nb), ctx.dev_alloc(nb)
    a_lo, a_hi = min(lo, n_pairs), min(hi, n_pairs)          # originals: (R1, rc(R1))
    b_lo, b_hi = max(lo, n_pairs) - n_pairs, max(hi, n_pairs) - n_pairs   # (rc(R2), R2)
    scratch = ctx.dev_alloc(max(max(a_hi - a_lo, b_hi - b_lo) * READ_LEN, 1))
    if a_hi > a_lo:
        m = a_hi - a_lo
        ctx.synth_pairs(d_tx, d_off, d_thr, n_tx, m, a_lo, seed, READ_LEN, FRAG_LEN,
                        synth.ERR_THRESHOLD_24, d_r1, scratch)
        ctx.revcomp_reads(d_r1, d_r2, m, READ_LEN)
    if b_hi > b_lo:
        m = b_hi - b_lo
        off = (a_hi - a_lo) * READ_LEN
        ctx.synth_pairs(d_tx, d_off, d_thr, n_tx, m, b_lo, seed, READ_LEN, FRAG_LEN,
                        synth.ERR_THRESHOLD_24, scratch, d_r2 + off)
        ctx.revcomp_reads(d_r2 + off, d_r1 + off, m, READ_LEN)
    ctx.dev_free(scratch)
    return d_r1, d_r2


class Workload(object):
    """The whole synthetic input on ONE device (N = 1, and the parity check of N > 1)."""

    def __init__(self, ctx, n_pairs, n_tx, seed, skewed=False):
        self.ctx = ctx
        self.n_pairs = n_pairs
        self.n_records = 2 * n_pairs           # record pairs after RC doubling
        n_tx, tx = device_transcripts(ctx, n_tx, seed, skewed)
        self.d_r1, self.d_r2 = generate_records(ctx, tx, n_tx, n_pairs, seed, 0, self.n_records)
        for d in tx:
            ctx.dev_free(d)
        windows = 2 * self.n_records * (READ_LEN - K1 + 1)
        t0 = time.perf_counter()
        self.d_keys, self.d_counts, self.n_kmers = ctx.count_k1mers(
            [self.d_r1, self.d_r2], [self.n_records, self.n_records], READ_LEN, K1,
            max(1 << 20, windows // 6))
        self.count_s = time.perf_counter() - t0
        self.h_offs = ctx.pinned_empty(self.n_records + 1, np.uint64)
        self.h_offs[:] = np.arange(self.n_records + 1, dtype=np.uint64) * np.uint64(READ_LEN)
        self.d_offs = ctx.to_device(self.h_offs)
        self.host = None

    def mates(self, on_device):
        if on_device:
            return [(self.d_r1, self.d_offs, self.n_records, True),
                    (self.d_r2, self.d_offs, self.n_records, True)]
        h = self.host
        return [(h["r1"], self.h_offs, None, False), (h["r2"], self.h_offs, None, False)]

    def stage_host(self):
        """Host copies of every input (pinned when possible) for the e2e leg."""
        ctx = self.ctx
        nb = self.n_records * READ_LEN
        self.host = {
            "r1": ctx.d2h(ctx.pinned_empty(nb, np.uint8), self.d_r1),
            "r2": ctx.d2h(ctx.pinned_empty(nb, np.uint8), self.d_r2),
            "keys": ctx.d2h(ctx.pinned_empty(self.n_kmers * (2 if K1 > 32 else 1), np.uint64), self.d_keys),
            "counts": ctx.d2h(ctx.pinned_empty(self.n_kmers, np.uint32), self.d_counts),
        }
        return sum(a.nbytes for a in self.host.values()) + self.h_offs.nbytes * 2

    def free(self):
        for d in (self.d_r1, self.d_r2, self.d_offs):
            self.ctx.dev_free(d)


def run_step(ctx, wl, on_device):
    from shannon_b200 import pipeline
    if on_device:
        keys, counts = wl.d_keys, wl.d_counts
    else:
        keys, counts = wl.host["keys"], wl.host["counts"]
        if K1 > 32:
            keys = keys.reshape(-1, 2)
    cor, comp_offs, rec_idx, stats = pipeline.frontend_in_memory(
        ctx, keys, counts, K1, wl.mates(on_device), True, 3, 75, 500, on_device, wl.n_kmers)
    # what comes back to the host: the partition, the contigs and the contig graph
    return stats, rec_idx.nbytes + comp_offs.nbytes + cor.sizes["contig_bases"] + \
        8 * (cor.sizes["n_contigs"] + 1) + 16 * cor.sizes["n_edges"] + 4 * (cor.sizes["n_contigs"] + 1), \
        (cor, comp_offs, rec_idx)


def result_digest(cor, comp_offs, rec_idx):
    """sha256 over everything the front end hands on: contigs, contig graph, component packing,
    read partition (used to compare the N-rank result with the 1-GPU result)."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(cor.contigs.bases).tobytes())
    h.update(np.ascontiguousarray(cor.contigs.offs).tobytes())
    for a in cor.adj_csr:
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(np.asarray(cor.comp_of, dtype=np.int64).tobytes())
    h.update(np.asarray(comp_offs, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(np.asarray(rec_idx, dtype=np.uint32)).tobytes())
    return h.hexdigest()


# SURVEY 8(d) algorithmic bytes per unit (the figure roofline.frac is computed from) and, beside
# it, the bytes of THIS design's data layout (64-byte buckets = one DRAM burst per probe).
def algorithmic_bytes(name, s, n_kmers, n_records):
    """(survey bytes, design bytes) of all launches of kernel `name` in one step; None = the kernel
    has no SURVEY 8(d) figure."""
    n_slots = s["n_slots"]
    survey = {
        "table_insert": 76 * n_kmers,                               # 8 B key + 4 B count + 32 B sector RMW
        "walk": 192 * s["n_traversed"],                             # 4 probes x 32 B + 64 B claim RMW
        "l4_assign": 161 * 2 * n_records,                           # per read record (two per record pair)
        "pack_reads": 125 * 2 * n_records,                          # L B in + L/4 B out
        "rmer_entries": 40 * s.get("candidate_bases", 0) or None,   # 32 B probe + 8 B key per base
        "table_lookup": 44 * s.get("n_lookup_queries", 0) or None,
    }
    design = {
        "table_clear": 16 * n_slots,
        "table_insert": 140 * n_kmers,
        "seed_count": 16 * n_slots,
        "seed_emit": 36 * n_slots + 12 * s["n_seeds"],             # slot read + idx parked (4 B) + slot line written back
        "uf_init": 4 * n_slots,
        "uf_edges": 16 * n_slots + (64 + 16) * s["n_loaded"],       # ONE successor probe sequence + parent / aux words
        "uf_flatten": 24 * n_slots,
        "comp_count": 4 * n_slots + 4 * s["n_loaded"],
        "walk": (4 * 64 + 64 + 1) * s["n_traversed"],
        "pack_reads": 2 * (READ_LEN + 32 + 20) * n_records,
        "l4_assign": (2 * (32 + 12)) * n_records + 32 * s["lookups"] + 8 * s["assignments"],
    }
    return survey.get(name), design.get(name)


def kernel_table(prof, steps, ms_per_step, stats, n_kmers, n_records, peak_gbs, traffic):
    kern = []
    for name, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        sv, ds = algorithmic_bytes(name, stats, n_kmers, n_records)
        e = {"kernel": name, "ms_per_step": ms / steps, "launches_per_step": n / steps,
             "share_of_step": (ms / steps) / ms_per_step}
        if sv:
            e["achieved_gbs"] = sv / 1e9 / (ms / steps / 1e3)
            e["frac"] = e["achieved_gbs"] / peak_gbs
        if ds:
            e["frac_design"] = ds / 1e9 / (ms / steps / 1e3) / peak_gbs
        if name in traffic:
            e["dram_bytes_per_launch"] = traffic[name]
        kern.append(e)
    return kern


def load_peaks():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else \
        "fallback 6650 GB/s (B200_PROFILING.md)"
    return peak_gbs, src


def load_traffic(pairs):
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("pairs_per_gpu") == pairs:
            return tj.get("bytes_per_launch", {})
    except Exception:
        pass
    return {}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's Python, 1 core (the path is single-threaded)
# ------------------------------------------------------------------------------------------------
def cpu_sample_case(n_pairs, n_tx, seed, workdir, skewed=False, fast_count=True):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    s1, s2 = helpers.synthetic_seqs(n_tx, n_pairs, seed, skewed=skewed)
    return helpers, helpers.make_case(workdir, K, s1, s2, fast_count=fast_count and K1 <= 32)


def time_oracle(helpers, case, name):
    import contextlib
    from oracle import shannon_oracle as so
    with contextlib.redirect_stdout(sys.stderr):       # gpmetis stand-in / os.system chatter
        t0 = time.perf_counter()
        run = helpers.run_frontend(so.extension_correction, so.kmers_for_component, case, name)
        dt = time.perf_counter() - t0
    return dt, run


def sample_config(args, n_pairs, n_tx):
    return {"workload": "%s SUBSAMPLE actually run by this arm: synthetic %d x 2 x %d bp read pairs from "
                        "%d transcripts (same generator, coverage and error rate as the full workload of "
                        "%d pairs / %d transcripts), K=%d, RC-doubled -> %d read records; full front end "
                        "through the file-level entry points (extension_correction + kmers_for_component)"
                        % (CONFIGS[args.config]["name"], n_pairs, READ_LEN, n_tx, args.pairs,
                           args.transcripts, K, 4 * n_pairs),
            "pairs": n_pairs, "transcripts": n_tx, "K": K, "full_workload_pairs": args.pairs,
            "parallelism": "1 CPU core (the reference path is single-threaded, SURVEY 2.2)"}


def reference_arm(args, rank, world):
    """--impl reference: the reference's own (Python, single-threaded) algorithm for this path,
    timed on the host cores.  The reference is Python 2 source that cannot travel to the GPU box
    (no /root/reference there); oracle/shannon_oracle.py is its pinned restatement ("port").
    The pass is timed ONCE on a bounded subsample (default 250 k pairs = 10^6 read records, about
    25 s) and that time is reported for every step: K + W repetitions of a pass that long would
    not finish within minutes, and a 6 k-pair sample (round 1) is not the same regime."""
    if rank != 0:
        return
    n_pairs = args.sample_pairs or 250_000
    n_tx = max(2, round(args.transcripts * n_pairs / float(args.pairs)))
    work = tempfile.mkdtemp(prefix="shn_ref_")
    t0 = time.perf_counter()
    helpers, case = cpu_sample_case(n_pairs, n_tx, args.seed, work, args.skewed)
    setup_s = time.perf_counter() - t0
    t, _ = time_oracle(helpers, case, "ref")
    ms = 1000.0 * t
    records = 4 * n_pairs
    value = records / t
    sample = ("%d pairs from %d transcripts (same generator and coverage as the workload); ONE timed "
              "pass, reported for all %d steps" % (n_pairs, n_tx, args.steps))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "steps_timed": 1,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "u64" if K1 <= 32 else "u128", "data": "synthetic",
        "config": sample_config(args, n_pairs, n_tx),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "cores_available": os.cpu_count(),
                         "kind": "port", "sample": sample, "seconds": t, "input_setup_seconds": setup_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, n_gpus):
    par = ("1 GPU, one table" if n_gpus == 1 else
           "hash-sharded K1-mer tables over %d ranks: lines routed to minimizer-hash owners (NCCL "
           "all-to-all), cross-rank K1-mer graph components, component re-sharding (all-to-all), "
           "per-rank walks, all-gathered candidates, reads sharded by record range" % n_gpus)
    return {"workload": "%s: synthetic %d x 2 x %d bp read pairs (WHOLE job%s) from %d transcripts (genes "
                        "with shared exons, %s expression), 1%% substitution error, K=%d, RC-doubled "
                        "(shannon.py:413-424) -> %d read records; full front end (table build -> walks -> "
                        "filters -> contig graph -> read partition)"
                        % (CONFIGS[args.config]["name"], args.pairs, READ_LEN,
                           "" if n_gpus == 1 else ", split over %d GPUs" % n_gpus, args.transcripts,
                           "Zipf" if args.skewed else "uniform", K, 4 * args.pairs),
            "pairs": args.pairs, "transcripts": args.transcripts, "K": K,
            "parallelism": par,
            "l2": "inputs (GBs) larger than the 126 MB L2; 256 MB scratch write between steps"}


# ------------------------------------------------------------------------------------------------
# file-level end to end: the reference's own entry points on files
# ------------------------------------------------------------------------------------------------
def file_e2e_leg(ctx, args):
    """extension_correction(argv) + kmers_for_component(..., repartition=True) (shannon.py:457-467)
    on a written k1mer.dict_org and two read FASTA files.  The input files are written by the
    library's native writers from device-generated data (outside the timed region)."""
    import contextlib
    import extension_correction as ec_mod
    import kmers_for_component as kfc_mod
    n_pairs = min(args.file_pairs, args.pairs)
    n_tx = max(2, round(args.transcripts * n_pairs / float(args.pairs)))
    work = tempfile.mkdtemp(prefix="shn_file_", dir=os.environ.get("SHN_BENCH_TMP"))
    t0 = time.perf_counter()
    wl = Workload(ctx, n_pairs, n_tx, args.seed, args.skewed)
    algo = os.path.join(work, "in_algo_input")
    out = os.path.join(work, "out")
    os.makedirs(algo)
    os.makedirs(os.path.join(out, "algo_input"))
    n_rec = wl.n_records
    files = [os.path.join(algo, "reads_1.fasta"), os.path.join(algo, "reads_2.fasta")]
    ident = np.arange(n_rec, dtype=np.uint32)
    for path, d in zip(files, (wl.d_r1, wl.d_r2)):
        h = ctx.d2h(np.empty(n_rec * READ_LEN, np.uint8), d)
        ctx.write_fasta_subset(path, False, h, wl.h_offs, ident, 0, "")
    k1dict_org = os.path.join(algo, "k1mer.dict_org")
    hk = ctx.d2h(np.empty(wl.n_kmers * (2 if K1 > 32 else 1), np.uint64), wl.d_keys)
    hc = ctx.d2h(np.empty(wl.n_kmers, np.uint32), wl.d_counts)
    ctx.write_kmer_file(k1dict_org, hk, hc, K1)
    in_bytes = sum(os.path.getsize(p) for p in files + [k1dict_org])
    wl.free()
    setup_s = time.perf_counter() - t0
    argv = [k1dict_org, os.path.join(out, "algo_input", "k1mer.dict"), "3", "75", out, "500", "1"] + files
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    with contextlib.redirect_stdout(sys.stderr):
        t0 = time.perf_counter()
        allowed, reads = ec_mod.extension_correction(argv, True)
        t_ec = time.perf_counter() - t0
        t0 = time.perf_counter()
        kfc_mod.kmers_for_component(allowed, os.path.join(out, "algo_input"), reads, files, out,
                                    "contigs.txt", True, False, True, True, 500, 2, K, helpers.GPMETIS, 5,
                                    False, False, 1)
        t_kfc = time.perf_counter() - t0
    out_bytes = 0
    for base, _, fns in os.walk(out):
        out_bytes += sum(os.path.getsize(os.path.join(base, f)) for f in fns)
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    tot = t_ec + t_kfc
    return {"value": 4 * n_pairs / tot, "unit": UNIT, "pairs": n_pairs, "transcripts": n_tx,
            "seconds": tot, "extension_correction_seconds": t_ec, "kmers_for_component_seconds": t_kfc,
            "sections_seconds": dict(getattr(ec_mod, "LAST_TIMINGS", {}), **getattr(kfc_mod, "LAST_TIMINGS", {})),
            "input_bytes": in_bytes, "output_bytes": out_bytes, "input_setup_seconds": setup_s,
            "what": "extension_correction(argv, inMem=True) + kmers_for_component(..., repartition=True, "
                    "inMem=False) exactly as shannon.py:457-467 calls them: parses k1mer.dict_org, reads "
                    "both FASTA files, writes contig / component / per-component read and K1-mer files"}


# ------------------------------------------------------------------------------------------------
# N > 1: the sharded path
# ------------------------------------------------------------------------------------------------
class ShardedWorkload(object):
    """This rank's part of ONE global workload: its range of the read records (generated on the
    device) and its slice of the lines of the global, ASCII-sorted k1mer.dict_org (distributed
    jellyfish stand-in: local count, (key, count) routed to key-RANGE owners, merged there, so the
    concatenation of the ranks' lists is the list the 1-GPU counter produces)."""

    def __init__(self, ctx, comm, n_pairs, n_tx, seed, skewed):
        import torch
        from shannon_b200.dist import shard_range
        assert K1 <= 32, "the distributed counter of bench.py carries one-word keys"
        self.ctx, self.comm = ctx, comm
        rank, world = comm.rank, comm.world
        dev = comm.device
        self.n_records_total = 2 * n_pairs
        self.rec_lo, self.rec_hi = shard_range(self.n_records_total, rank, world)
        n = self.n_records = self.rec_hi - self.rec_lo
        n_tx, tx = device_transcripts(ctx, n_tx, seed, skewed)
        self.d_r1, self.d_r2 = generate_records(ctx, tx, n_tx, n_pairs, seed, self.rec_lo, self.rec_hi)
        for d in tx:
            ctx.dev_free(d)
        t0 = time.perf_counter()
        windows = 2 * n * (READ_LEN - K1 + 1)
        d_keys, d_counts, nk = ctx.count_k1mers([self.d_r1, self.d_r2], [n, n], READ_LEN, K1,
                                                max(1 << 20, windows // 6))
        keys = torch.empty(nk, dtype=torch.int64, device=dev)
        cnts = torch.empty(nk, dtype=torch.int32, device=dev)
        ctx.d2d(keys.data_ptr(), d_keys, nk * 8)
        ctx.d2d(cnts.data_ptr(), d_counts, nk * 4)
        ctx.sync()
        ctx.count_release()
        ctx.trim()
        m5 = 0x5555555555555555

        def ascii_order(x):     # A0 G1 C2 T3 pairs <-> A0 C1 G2 T3 pairs (an involution)
            return ((x & m5) << 1) | ((x >> 1) & m5)
        ok = ascii_order(keys)  # ascending already: the local counter emits ASCII order
        del keys
        # key-range owners: equal slices of [0, 4^k1)
        bounds = torch.tensor([(r * (1 << (2 * K1))) // world for r in range(1, world)],
                              dtype=torch.int64, device=dev)
        cut = torch.searchsorted(ok, bounds).tolist()
        send_counts = np.diff([0] + cut + [nk]).tolist()
        rows = torch.stack([ok, cnts.to(torch.int64)], dim=1)
        del ok, cnts
        recv, _ = comm.all_to_all_rows(rows, send_counts)
        del rows
        order = torch.argsort(recv[:, 0])
        sk, sc = recv[order, 0], recv[order, 1]
        del recv, order
        uk, inv = torch.unique_consecutive(sk, return_inverse=True)
        uc = torch.zeros(uk.shape[0], dtype=torch.int64, device=dev).index_add_(0, inv, sc)
        del sk, sc, inv
        self.keys = ascii_order(uk).contiguous()
        self.counts = uc.to(torch.int32).contiguous()
        self.n_lines = int(uk.shape[0])
        del uk, uc
        torch.cuda.empty_cache()
        locs = [v[0] for v in comm.exchange_ints([self.n_lines])]
        self.first_line = sum(locs[:rank])
        self.n_kmers_total = sum(locs)
        self.count_s = time.perf_counter() - t0
        self.h_offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(READ_LEN)
        self.d_offs = ctx.to_device(self.h_offs)
        self.host = None

    def mates(self):
        return [(self.d_r1, self.d_offs, self.n_records, True), (self.d_r2, self.d_offs, self.n_records, True)]

    def stage_host(self):
        import torch
        ctx = self.ctx
        nb = self.n_records * READ_LEN
        self.host = {"r1": ctx.d2h(ctx.pinned_empty(nb, np.uint8), self.d_r1),
                     "r2": ctx.d2h(ctx.pinned_empty(nb, np.uint8), self.d_r2),
                     "keys": self.keys.cpu().pin_memory(), "counts": self.counts.cpu().pin_memory()}
        return 2 * nb + self.n_lines * 12


def sharded_step(comm, ops, ctx, wl, from_host=False):
    from shannon_b200 import sharded
    if from_host:     # e2e: this rank's inputs start in pinned host memory
        import torch
        keys = wl.host["keys"].to(comm.device, non_blocking=True)
        counts = wl.host["counts"].to(comm.device, non_blocking=True)
        mates = [(wl.host["r1"], wl.h_offs, None, False), (wl.host["r2"], wl.h_offs, None, False)]
    else:
        keys, counts, mates = wl.keys, wl.counts, wl.mates()
    return sharded.frontend_sharded(comm, ops, ctx, keys.data_ptr(), counts.data_ptr(), wl.n_lines,
                                    wl.first_line, K1, mates, wl.rec_lo, True, 3, 75, 500)


def main_sharded(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from shannon_b200 import _lib, sharded
    dev = torch.device("cuda", local_rank)
    ctx = _lib.Context(local_rank)
    ops = sharded.GpuOps(ctx, dev)
    comm = sharded.TorchComm(device=dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.perf_counter()
    wl = ShardedWorkload(ctx, comm, args.pairs, args.transcripts, args.seed, args.skewed)
    setup_s = time.perf_counter() - t_setup
    # the first all-to-all of a process sets up NCCL's point-to-point channels: that happened in the
    # distributed counter above
    for _ in range(args.warmup):
        sharded_step(comm, ops, ctx, wl)
        ctx.flush_l2()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.prof_enable(True)
    launches0 = ctx.launch_count()
    comm.bytes_sent = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cor, comp_offs, rec_idx, stats = sharded_step(comm, ops, ctx, wl)
        ctx.flush_l2()
    ev1.record()
    ev1.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    sent = comm.bytes_sent / args.steps
    prof = ctx.prof()
    ctx.prof_enable(False)
    stats["n_slots"] = ctx.table_stats()["n_slots"]
    t = torch.tensor([max(dev_ms, wall_ms) / args.steps, float(sent), float(launches)], device=dev,
                     dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_per_step = float(tmax[0].item())
    records = 2 * wl.n_records_total
    value = records / (ms_per_step / 1000.0)
    peak_gbs, peak_src = load_peaks()
    # this rank's kernels; per-rank units (this rank's shard of lines / records)
    kern = kernel_table(prof, args.steps, ms_per_step, stats, stats.get("n_owned_keys", wl.n_lines),
                        wl.n_records, peak_gbs, {})
    dom = next((k for k in kern if k.get("frac") is not None), None)
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak_gbs,
                    "unit": "GB/s", "frac": dom["frac"], "frac_design": dom.get("frac_design"), "traffic": None,
                    "peak_source": peak_src, "share_of_step": dom["share_of_step"],
                    "note": "rank 0's largest kernel with a SURVEY 8(d) byte figure; units = this rank's share"}
    nvlink = {"bytes_sent_per_step_all_ranks": float(t[1].item()),
              "bytes_sent_per_step_max_rank": float(tmax[1].item()),
              "exchange_ms_per_step": None, "unit": "GB/s", "peak": NVLINK_GBS,
              "note": "payload of the all-to-all / all-gather exchanges of one step (self-sends excluded); "
                      "gbs_step = bytes of the busiest rank / whole step time (the exchanges are a small "
                      "part of the step: it is not link-bound)"}
    nvlink["gbs_step"] = nvlink["bytes_sent_per_step_max_rank"] / 1e9 / (ms_per_step / 1e3)
    nvlink["frac"] = nvlink["gbs_step"] / NVLINK_GBS

    # ---- e2e: host buffers of every rank ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h2d_bytes = wl.stage_host()
        sharded_step(comm, ops, ctx, wl, from_host=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cor, comp_offs, rec_idx, stats_e = sharded_step(comm, ops, ctx, wl, from_host=True)
        torch.cuda.synchronize()
        e_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e_ms, float(h2d_bytes)], device=dev, dtype=torch.float64)
        tem = te.clone()
        dist.all_reduce(tem, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        d2h = 0
        if rank == 0:
            d2h = rec_idx.nbytes + comp_offs.nbytes + cor.sizes["contig_bases"] + 8 * (cor.sizes["n_contigs"] + 1)
        e2e = {"value": records / (float(tem[0].item()) / 1000.0), "unit": UNIT,
               "ms_per_step": float(tem[0].item()), "h2d_bytes_per_step": int(te[1].item()),
               "d2h_bytes_per_step": int(d2h), "stage_wall_ms": stats_e.get("host_timings_ms")}

    # ---- parity of the sharded result against ONE GPU on the same global input --------------------
    dist_parity = None
    parity_note = None
    fits = args.pairs <= 12_000_000 and K1 <= 32
    if not args.no_dist_parity:
        if fits:
            digest_n = result_digest(cor, comp_offs, rec_idx) if rank == 0 else None
            ops.close()                              # back to the context's own stream
            if rank == 0:
                del wl.keys, wl.counts
                torch.cuda.empty_cache()
                ref = Workload(ctx, args.pairs, args.transcripts, args.seed, args.skewed)
                assert ref.n_kmers == wl.n_kmers_total, "distributed counter disagrees with the 1-GPU counter"
                _, _, (cor1, offs1, idx1) = run_step(ctx, ref, True)
                dist_parity = bool(result_digest(cor1, offs1, idx1) == digest_n)
                parity_note = ("sha256 over contigs, contig graph, components and the read partition: "
                               "%d-rank sharded run vs pipeline.frontend_in_memory on one GPU, same input"
                               % world)
            barrier()
        else:
            parity_note = "workload does not fit one GPU: no 1-GPU result to compare with"

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, world),
            "kmer_lookups_per_sec": stats["lookups"] / (ms_per_step / 1000.0),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(t[2].item()),
            "roofline": roofline, "cpu_baseline": None, "dist_parity": dist_parity,
            "dist_parity_note": parity_note, "nvlink": nvlink,
            "kernels": kern[:int(os.environ.get("SHN_BENCH_KERNELS", "14"))],
            "workload_stats": dict((k, int(v)) for k, v in stats.items() if k != "host_timings_ms"),
            "stage_wall_ms": stats.get("host_timings_ms"),
            "setup_seconds": setup_s, "kmer_count_seconds": wl.count_s,
            "n_kmers_total": wl.n_kmers_total,
        }
        print(json.dumps(out), flush=True)
    try:
        ops.close()
    except Exception:
        pass
    ctx.close()
    dist.destroy_process_group()


def main():
    global K, K1
    args = parse_args()
    K, K1 = args.K, args.K + 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from shannon_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        main_sharded(args, rank, world, local_rank)
        return

    ctx = _lib.Context(local_rank)
    t_setup = time.perf_counter()
    wl = Workload(ctx, args.pairs, args.transcripts, args.seed, args.skewed)
    setup_s = time.perf_counter() - t_setup

    # ---- value: inputs resident in HBM --------------------------------------------------------
    for _ in range(args.warmup):
        run_step(ctx, wl, True)
        ctx.flush_l2()
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    sampler.start()
    ctx.prof_enable(True)
    launches0 = ctx.launch_count()
    ctx.sync()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stats_i, d2h_bytes, _ = run_step(ctx, wl, True)
        ctx.flush_l2()
    ctx.sync()
    dev_ms = ctx.timer_stop()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    stats = stats_i
    stats["n_slots"] = ctx.table_stats()["n_slots"]
    launches = ctx.launch_count() - launches0
    prof = ctx.prof()
    ctx.prof_enable(False)
    ms_per_step = max(dev_ms, wall_ms) / args.steps
    records = 2 * wl.n_records                       # read records (both mates)
    value = records / (ms_per_step / 1000.0)
    lookups_per_s = stats["lookups"] / (ms_per_step / 1000.0)

    # ---- roofline of the dominant kernel (CUDA events around every launch, timed region) ------
    peak_gbs, peak_src = load_peaks()
    traffic = load_traffic(args.pairs)
    kern = kernel_table(prof, args.steps, ms_per_step, stats, wl.n_kmers, wl.n_records, peak_gbs, traffic)
    step_kernel_ms = sum(v[0] for v in prof.values()) / args.steps
    dom = kern[0] if kern else None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom.get("achieved_gbs"),
                    "peak": peak_gbs, "unit": "GB/s", "frac": dom.get("frac"),
                    "frac_design": dom.get("frac_design"),
                    "traffic": traffic.get(dom["kernel"]), "peak_source": peak_src,
                    "share_of_step": dom["share_of_step"],
                    "walk_rounds": stats.get("walk_rounds"),
                    "note": "achieved = SURVEY 8(d) algorithmic bytes per unit x units per step (192 B per "
                            "traversed K1-mer, 76 B per inserted line, 161 B per read record) / CUDA-event "
                            "duration inside the timed region; frac_design = the same with this design's "
                            "64-byte-bucket byte model (DESIGN.md section 5; the peak is a COPY bandwidth, so "
                            "a pure read or pure write stream such as seed_count / table_clear can exceed 1).  "
                            "The walk stage is bound by instruction issue and the latency of its dependent "
                            "2-step rounds (walk_rounds on the longest one-warp component), not by bandwidth"}
        if dom["kernel"] == "walk" and stats.get("walk_rounds"):
            roofline["round_time_us"] = 1000.0 * dom["ms_per_step"] / stats["walk_rounds"]

    # ---- e2e: same step through the C-ABI with host buffers ----------------------------------
    e2e = None
    e2e_ok = not args.no_e2e
    if e2e_ok:
        need = 2 * wl.n_records * READ_LEN + 12 * wl.n_kmers * (2 if K1 > 32 else 1)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        e2e_ok = avail is None or need < 0.6 * avail
        if not e2e_ok:
            sys.stderr.write("bench: e2e leg skipped, %d GB of pinned host buffers do not fit\n" % (need >> 30))
    if e2e_ok:
        h2d_bytes = wl.stage_host()
        run_step(ctx, wl, False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stats_e, d2h_bytes, _ = run_step(ctx, wl, False)
        ctx.sync()
        e_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
        e2e = {"value": records / (e_ms / 1000.0), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
               "stage_wall_ms": stats_e.get("host_timings_ms"),
               "what": "pipeline.frontend_in_memory through the C-ABI with pinned HOST buffers (packed K1-mer "
                       "list + ASCII reads in, partition out); the file-level entry points are timed in file_e2e"}

    # ---- CPU baseline on a bounded sample (rank 0) + parity of the GPU path on it --------------
    cpu = None
    if not args.no_cpu_baseline:
        n_pairs = args.sample_pairs or 60_000
        n_tx = max(2, round(args.transcripts * n_pairs / float(args.pairs)))
        work = tempfile.mkdtemp(prefix="shn_cpu_")
        helpers, case = cpu_sample_case(n_pairs, n_tx, args.seed, work, args.skewed)
        t_cpu, run_cpu = time_oracle(helpers, case, "oracle")
        import contextlib
        import extension_correction as ec_mod
        import kmers_for_component as kfc_mod
        with contextlib.redirect_stdout(sys.stderr):   # the modules print the reference's log lines
            run_gpu = helpers.run_frontend(ec_mod.extension_correction,
                                           kfc_mod.kmers_for_component, case, "gpu")
        helpers.assert_same_run(run_cpu, run_gpu, "bench sample: gpu vs oracle")
        cpu = {"value": 4 * n_pairs / t_cpu, "unit": UNIT, "cores": 1,
               "cores_available": os.cpu_count(), "kind": "port", "seconds": t_cpu,
               "sample": "%d pairs from %d transcripts (same generator and coverage as the "
                         "workload); oracle/shannon_oracle.py, the pinned Python restatement of the "
                         "reference (single-threaded like the reference); GPU output on the same "
                         "sample checked bit-identical" % (n_pairs, n_tx)}

    # ---- the reference's own entry points on files ----------------------------------------------
    file_e2e = None
    if not args.no_file_e2e:
        wl.free()
        wl.host = None
        file_e2e = file_e2e_leg(ctx, args)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64" if K1 <= 32 else "u128", "data": "synthetic",
        "config": workload_config(args, world),
        "kmer_lookups_per_sec": lookups_per_s,
        "assign_kernel_lookups_per_sec": (stats["lookups"] / (prof["l4_assign"][0] / args.steps / 1e3)
                                          if "l4_assign" in prof else None),
        "clocks": clocks, "e2e": e2e, "file_e2e": file_e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "kernels": kern[:int(os.environ.get("SHN_BENCH_KERNELS", "14"))],
        "kernel_ms_per_step": step_kernel_ms,
        "host_ms_per_step": ms_per_step - step_kernel_ms,
        "workload_stats": dict((k, int(v)) for k, v in stats.items() if k != "host_timings_ms"),
        "stage_wall_ms": stats_i.get("host_timings_ms"),
        "setup_seconds": setup_s, "kmer_count_seconds": wl.count_s,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    import faulthandler
    faulthandler.enable()
    main()

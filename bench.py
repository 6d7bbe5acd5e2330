#!/usr/bin/env python
"""bench.py -- throughput of the Shannon k-mer front end on B200 (BASELINE.json metric:
reads partitioned/sec & k-mer lookups/sec; HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

A "step" is ONE pass of the whole hot path over the workload: K1-mer table build
(load_kmers + lowComplexity), seed ordering, greedy walks, shape + duplicate filters, contig
C-mer graph + components, K1-mer->component map, read packing and read->component partition.
Workload at N=1: BASELINE.json configs[2] -- synthetic 10 M 100-bp read pairs from 5 k
transcripts, 1 % substitution error, K=24 (generated on the device, RC-doubled like
shannon.py:413-424 and counted with the jellyfish stand-in, all outside the timed region).
`value` = read records partitioned per second with inputs resident in HBM; `e2e` = the same
step through the C-ABI with HOST buffers (H2D of the K1-mer list and the ASCII reads and D2H of
the partition inside the timed region).  For N > 1 every rank runs the path on its own shard
(disjoint transcript sets, weak scaling, no data-path collective yet -- see DESIGN.md).
"""
import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per GPU")
    ap.add_argument("--transcripts", type=int, default=5000, help="transcripts per GPU")
    ap.add_argument("--seed", type=int, default=1234 + 2)
    ap.add_argument("--K", type=int, default=24, help="k-mer size (K1 = K+1; 32 -> 128-bit keys)")
    ap.add_argument("--sample-pairs", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


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
class Workload(object):
    """Synthetic config-3 style input, resident on the device."""

    def __init__(self, ctx, n_pairs, n_tx, seed):
        from shannon_b200 import synth
        self.ctx = ctx
        self.n_pairs = n_pairs
        self.n_records = 2 * n_pairs           # record pairs after RC doubling
        tx = synth.make_transcripts(n_tx, seed)
        codes, offs = synth.pack_transcripts(tx)
        thr = synth.expression_thresholds(len(tx), [len(t) for t in tx], False)
        self.tx_bases = int(offs[-1])
        d_tx, d_off, d_thr = ctx.to_device(codes), ctx.to_device(offs), ctx.to_device(thr)
        nb = self.n_records * READ_LEN
        self.d_r1, self.d_r2 = ctx.dev_alloc(nb), ctx.dev_alloc(nb)
        half = n_pairs * READ_LEN
        # reads_1 = [R1 ; rc(R2)], reads_2 = [rc(R1) ; R2]   (shannon.py:413-424)
        ctx.synth_pairs(d_tx, d_off, d_thr, len(tx), n_pairs, 0, seed, READ_LEN, FRAG_LEN,
                        synth.ERR_THRESHOLD_24, self.d_r1, self.d_r2 + half)
        ctx.revcomp_reads(self.d_r2 + half, self.d_r1 + half, n_pairs, READ_LEN)
        ctx.revcomp_reads(self.d_r1, self.d_r2, n_pairs, READ_LEN)
        for d in (d_tx, d_off, d_thr):
            ctx.dev_free(d)
        windows = 2 * self.n_records * (READ_LEN - K1 + 1)
        t0 = time.perf_counter()
        self.d_keys, self.d_counts, self.n_kmers = ctx.count_k1mers(
            [self.d_r1, self.d_r2], [self.n_records, self.n_records], READ_LEN, K1,
            max(1 << 20, windows // 4))
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
        8 * (cor.sizes["n_contigs"] + 1) + 16 * cor.sizes["n_edges"] + 4 * (cor.sizes["n_contigs"] + 1)


# algorithmic bytes per step of every launch of a kernel (DESIGN.md "Kernels"); s = step stats
def algorithmic_bytes(name, s, wl):
    n_slots = s["n_slots"]
    table = {
        "table_clear": 16 * n_slots,
        "table_insert": 140 * wl.n_kmers,
        "seed_count": 16 * n_slots,
        "seed_emit": 16 * n_slots + 12 * s["n_seeds"],
        "uf_init": 4 * n_slots,
        "uf_edges": 16 * n_slots + (4 * 64 + 8) * s["n_loaded"],
        "uf_flatten": 24 * n_slots,
        "comp_count": 20 * n_slots + 8 * s["n_loaded"],
        "walk": (4 * 64 + 64 + 1) * s["n_traversed"],
        "pack_reads": 2 * (READ_LEN + 32 + 20) * wl.n_records,      # two launches: one per mate file
        "l4_assign": (2 * (32 + 12)) * wl.n_records + 32 * s["lookups"] + 8 * s["assignments"],
    }
    return table.get(name)


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's Python, 1 core (the path is single-threaded)
# ------------------------------------------------------------------------------------------------
def cpu_sample_case(n_pairs, n_tx, seed, workdir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    s1, s2 = helpers.synthetic_seqs(n_tx, n_pairs, seed)
    return helpers, helpers.make_case(workdir, K, s1, s2)


def time_oracle(helpers, case, name):
    import contextlib
    from oracle import shannon_oracle as so
    with contextlib.redirect_stdout(sys.stderr):       # gpmetis stand-in / os.system chatter
        t0 = time.perf_counter()
        run = helpers.run_frontend(so.extension_correction, so.kmers_for_component, case, name)
        dt = time.perf_counter() - t0
    return dt, run


def reference_arm(args, rank, world):
    """--impl reference: the reference's own (Python, single-threaded) algorithm for this path,
    timed on the host cores.  The reference is Python 2 source that cannot travel to the GPU box
    (no /root/reference there); oracle/shannon_oracle.py is its pinned restatement ("port")."""
    if rank != 0:
        return
    n_pairs = args.sample_pairs or 6000
    n_tx = max(2, round(args.transcripts * n_pairs / float(args.pairs)))
    work = tempfile.mkdtemp(prefix="shn_ref_")
    helpers, case = cpu_sample_case(n_pairs, n_tx, args.seed, work)
    times = []
    for i in range(args.warmup + args.steps):
        t, _ = time_oracle(helpers, case, "ref%d" % i)
        if i >= args.warmup:
            times.append(t)
    ms = 1000.0 * sum(times) / len(times)
    records = 4 * n_pairs
    value = records / (ms / 1000.0)
    sample = "%d pairs from %d transcripts (same generator and coverage as the workload)" % (
        n_pairs, n_tx)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, n_gpus):
    return {"workload": "BASELINE.json configs[2]: synthetic %d x 2 x %d bp read pairs per GPU from "
                        "%d transcripts (genes with shared exons), 1%% substitution error, K=%d, "
                        "RC-doubled (shannon.py:413-424) -> %d read records per GPU; full front end "
                        "(table build -> walks -> filters -> contig graph -> read partition)"
                        % (args.pairs, READ_LEN, args.transcripts, K, 4 * args.pairs),
            "pairs_per_gpu": args.pairs, "transcripts_per_gpu": args.transcripts, "K": K,
            "parallelism": "%d independent shard(s), one per GPU" % n_gpus,
            "l2": "inputs (GBs) larger than the 126 MB L2; 256 MB scratch write between steps"}


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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = _lib.Context(local_rank)
    t_setup = time.perf_counter()
    wl = Workload(ctx, args.pairs, args.transcripts, args.seed + 1000 * rank)
    setup_s = time.perf_counter() - t_setup

    # ---- value: inputs resident in HBM --------------------------------------------------------
    for _ in range(args.warmup):
        run_step(ctx, wl, True)
        ctx.flush_l2()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.prof_enable(True)
    launches0 = ctx.launch_count()
    ctx.sync()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stats_i, d2h_bytes = run_step(ctx, wl, True)
        ctx.flush_l2()
    ctx.sync()
    dev_ms = ctx.timer_stop()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    stats = stats_i
    stats["n_slots"] = ctx.table_stats()["n_slots"]
    launches = ctx.launch_count() - launches0
    prof = ctx.prof()
    ctx.prof_enable(False)
    ms_local = max(dev_ms, wall_ms) / args.steps
    t = torch.tensor([ms_local], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    records = 2 * wl.n_records                       # read records (both mates) per GPU
    value = world * records / (ms_per_step / 1000.0)
    lookups_per_s = world * stats["lookups"] / (ms_per_step / 1000.0)

    # ---- roofline of the dominant kernel (CUDA events around every launch, timed region) ------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("pairs_per_gpu") == args.pairs:
            traffic = tj.get("bytes_per_launch", {})
    except Exception:
        pass
    kern = []
    step_kernel_ms = sum(v[0] for v in prof.values()) / args.steps
    for name, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ab = algorithmic_bytes(name, stats, wl)
        e = {"kernel": name, "ms_per_step": ms / args.steps, "launches_per_step": n / args.steps,
             "share_of_step": (ms / args.steps) / ms_per_step}
        if ab is not None:            # ab = algorithmic bytes of all launches of this kernel in a step
            e["achieved_gbs"] = ab / 1e9 / (ms / args.steps / 1e3)
            e["frac"] = e["achieved_gbs"] / peak_gbs
        kern.append(e)
    dom = kern[0] if kern else None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom.get("achieved_gbs"),
                    "peak": peak_gbs, "unit": "GB/s", "frac": dom.get("frac"),
                    "traffic": traffic.get(dom["kernel"]), "peak_source": peak_src,
                    "share_of_step": dom["share_of_step"],
                    "note": "achieved = algorithmic bytes per launch (DESIGN.md) / CUDA-event "
                            "duration of that launch inside the timed region"}

    # ---- e2e: same step through the C-ABI with host buffers ----------------------------------
    e2e = None
    e2e_ok = not args.no_e2e
    if e2e_ok:
        # every rank pins its own copy of the inputs: do not drive the box out of host memory
        need = world * (2 * wl.n_records * READ_LEN + 12 * wl.n_kmers * (2 if K1 > 32 else 1))
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        flag = torch.tensor([1 if (avail is None or need < 0.6 * avail) else 0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        e2e_ok = bool(flag.item())
        if not e2e_ok and rank == 0:
            sys.stderr.write("bench: e2e leg skipped, %d GB of pinned host buffers do not fit\n" % (need >> 30))
    if e2e_ok:
        h2d_bytes = wl.stage_host()
        run_step(ctx, wl, False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stats_e, d2h_bytes = run_step(ctx, wl, False)
        ctx.sync()
        e_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
        t = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
        e2e = {"value": world * records / (e_ms / 1000.0), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
               "stage_wall_ms": stats_e.get("host_timings_ms")}

    # ---- N > 1: the hash-sharded K1-mer table (all-to-all over NVLink) as a lookup service ------
    dist_table = None
    if world > 1 and K1 <= 32:   # the hash-routed table carries one-word keys
        from shannon_b200 import dist as sdist
        ops = sdist.GpuOps(ctx, local_rank)
        tab = sdist.ShardedKmerTable(ops)
        n_k = wl.n_kmers
        tk = torch.empty(n_k, dtype=torch.int64, device="cuda")
        tc = torch.empty(n_k, dtype=torch.int32, device="cuda")
        ctx.sync()
        tk.copy_(torch.frombuffer(ctx.d2h(np.empty(n_k, np.uint64), wl.d_keys), dtype=torch.int64))
        tc.copy_(torch.frombuffer(ctx.d2h(np.empty(n_k, np.uint32), wl.d_counts), dtype=torch.int32))
        counts_all = [None] * world
        dist.all_gather_object(counts_all, n_k)
        first_line = sum(counts_all[:rank])
        # first all-to-all of the process sets up NCCL's point-to-point channels: keep that out of
        # the build time
        wu_in = torch.zeros(world, dtype=torch.int64, device="cuda")
        wu_out = torch.empty(world, dtype=torch.int64, device="cuda")
        dist.all_to_all_single(wu_out, wu_in)
        barrier()
        t0 = time.perf_counter()
        tab.build(tk, tc, first_line, K1)
        barrier()
        t_build = time.perf_counter() - t0
        nq = min(n_k, 50_000_000)
        q = tk[torch.randperm(n_k, device="cuda")[:nq]].contiguous()
        tab.lookup(q[:1000])
        barrier()
        t0 = time.perf_counter()
        w, f = tab.lookup(q)
        barrier()
        t_look = time.perf_counter() - t0
        ok = bool(f.all().item())
        tot_k = torch.tensor([float(n_k), float(nq)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot_k)
        dist_table = {"build_keys_per_s": float(tot_k[0]) / t_build,
                      "lookups_per_s": float(tot_k[1]) / t_look, "all_found": ok,
                      "alltoall_bytes_per_lookup": 8 + 4 + 1,
                      "note": "keys of every rank's shard routed to hash owners with NCCL "
                              "all_to_all_single; lookups = route, probe on the owner, route back"}
        del tk, tc, q, w, f

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) + parity of the GPU path on it ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_pairs = args.sample_pairs or 20000
        n_tx = max(2, round(args.transcripts * n_pairs / float(args.pairs)))
        work = tempfile.mkdtemp(prefix="shn_cpu_")
        helpers, case = cpu_sample_case(n_pairs, n_tx, args.seed, work)
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

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, world),
            "kmer_lookups_per_sec": lookups_per_s,
            "assign_kernel_lookups_per_sec": (stats["lookups"] / (prof["l4_assign"][0] / args.steps / 1e3)
                                              if "l4_assign" in prof else None),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "dist_table": dist_table,
            "kernels": kern[:int(os.environ.get("SHN_BENCH_KERNELS", "14"))],
            "kernel_ms_per_step": step_kernel_ms,
            "host_ms_per_step": ms_per_step - step_kernel_ms,
            "workload_stats": dict((k, int(v)) for k, v in stats.items() if k != "host_timings_ms"),
            "stage_wall_ms": stats_i.get("host_timings_ms"),
            "setup_seconds": setup_s, "kmer_count_seconds": wl.count_s,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.json.gz by running the REAL reference (from /root/reference via
oracle/ref_loader.py) -- run in the build container only:

    python tests/golden/make_golden.py [case names ...]      (default: all cases)

Each fixture holds the inputs (read sequences; k1mer.dict_org is re-derived with the
deterministic stand-in counter and pinned by sha256), the run parameters, and everything the
reference produced: the returned allowed_kmer_dict, every output file, and the normalised
return value of kmers_for_component.  The bundled samples (BASELINE.json configs 1-2) are
stored as sequence lines only under tests/golden/samples/.
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from oracle import ref_loader  # noqa: E402

sys.path.insert(0, HERE)
from cases import CASES, case_inputs, digest, digest_allowed  # noqa: E402


def export_samples():
    d = os.path.join(HERE, "samples")
    os.makedirs(d, exist_ok=True)
    src = os.path.join(ref_loader.REFERENCE_DIR, "Samples")
    pe = {}
    for name in ("SE_read", "PE_read_1", "PE_read_2"):
        seqs = helpers.read_fasta_seqs(os.path.join(src, name + ".fasta"))
        pe[name] = seqs
        with gzip.GzipFile(os.path.join(d, name + ".seqs.gz"), "wb", mtime=0) as f:
            f.write(("\n".join(seqs) + "\n").encode())
    # config 1 names Samples/SE_reads.fastq: its sequence lines are PE_read_1 ++ PE_read_2
    with open(os.path.join(src, "SE_reads.fastq")) as f:
        fq = [l.rstrip("\n") for i, l in enumerate(f) if i % 4 == 1]
    assert fq == pe["PE_read_1"] + pe["PE_read_2"], "SE_reads.fastq is not PE_read_1 ++ PE_read_2"


def main():
    export_samples()
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        work = tempfile.mkdtemp(prefix="golden_")
        seqs1, seqs2 = case_inputs(spec)
        case = helpers.make_case(work, spec["K"], seqs1, seqs2,
                                 double_stranded=spec.get("rc_double", True),
                                 fast_count=spec.get("fast_count", False))
        ec = ref_loader.load("extension_correction")
        kfc = ref_loader.load("kmers_for_component")
        out, allowed, reads, ret = helpers.run_frontend(
            ec.extension_correction, kfc.kmers_for_component, case, "ref", **spec["run"])
        snap = helpers.snapshot(out)
        with open(case.k1mer_org, "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        cb, new_comps, cw, rps = helpers.normalise_ret(ret)
        ret_doc = {"components_broken": dict((str(k), v) for k, v in cb.items()),
                   "new_comps": new_comps, "contig_weights": cw, "rps": rps}
        if spec.get("digest"):   # large case: sha256 of every output instead of the bytes
            files = {}
            for k, v in snap.items():
                if k.endswith("algo_input/k1mer.dict"):
                    v = b"".join(sorted(v.splitlines(True)))
                files[k] = digest(v)
            doc = {"name": name, "digest": True, "k1mer_dict_org_sha256": sha,
                   "n_allowed": len(allowed), "allowed_sha256": digest_allowed(allowed),
                   "files_sha256": files,
                   "ret_sha256": digest(json.dumps(json.loads(json.dumps(ret_doc)), sort_keys=True))}
        else:
            doc = {
                "name": name,
                "k1mer_dict_org_sha256": sha,
                "allowed_kmer_dict": allowed,
                "files": dict((k, v.decode()) for k, v in snap.items()),
                "ret": ret_doc,
            }
        path = os.path.join(HERE, name + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(doc, sort_keys=True).encode())
        print(name, os.path.getsize(path), "bytes;", len(allowed), "allowed k1mers;",
              len(snap), "files")


if __name__ == "__main__":
    main()

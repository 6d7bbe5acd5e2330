"""Row f4: faster_reps.py on the B200 against its CPU restatement
(oracle/reps_oracle.py, pinned to the real faster_reps.py in tests/test_oracle_vs_reference.py)."""
import os

import pytest

import helpers
from oracle import reps_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,ds,n", [(1, False, 60), (2, True, 60), (3, True, 300), (4, False, 300),
                                       (5, True, 3000)])
def test_find_reps_equals_oracle(workdir, seed, ds, n):
    import faster_reps                      # top-level shim, the reference's script name
    src = helpers.transcripts_file(os.path.join(workdir, "t.fasta"), seed, n)
    reps_oracle.find_reps(src, os.path.join(workdir, "ora.fasta"), ds)
    faster_reps.find_reps(src, os.path.join(workdir, "gpu.fasta"), ds)
    a, b = open(os.path.join(workdir, "ora.fasta")).read(), open(os.path.join(workdir, "gpu.fasta")).read()
    assert a == b
    assert 0 < a.count(">") < open(src).read().count(">")


def test_find_reps_on_pipeline_contigs(workdir):
    """contigs of a front-end run: repeats and reverse-complement twins (RC-doubled reads)"""
    import faster_reps
    from oracle import shannon_oracle as so
    s1, s2 = helpers.synthetic_seqs(20, 4000, 9)
    case = helpers.make_case(workdir, 24, s1, s2)
    out, _, _, _ = helpers.run_frontend(so.extension_correction, so.kmers_for_component, case, "ora")
    contigs = open(out + "/algo_input/k1mer.dict_contig").read().split()
    src = os.path.join(workdir, "contigs.fasta")
    with open(src, "w") as f:
        for i, c in enumerate(contigs):
            f.write(">c%d\n%s\n" % (i, c))
    for ds in (False, True):
        reps_oracle.find_reps(src, os.path.join(workdir, "ora.fasta"), ds)
        faster_reps.find_reps(src, os.path.join(workdir, "gpu.fasta"), ds)
        assert open(os.path.join(workdir, "ora.fasta")).read() == open(os.path.join(workdir, "gpu.fasta")).read()


def test_find_reps_edge_cases(workdir):
    import faster_reps
    from shannon_b200 import _lib
    p = os.path.join(workdir, "e.fasta")
    open(p, "w").close()
    faster_reps.find_reps(p, p + ".out", True)
    assert open(p + ".out").read() == ""
    with open(p, "w") as f:
        f.write(">a\nACGT\n>b\n" + "ACGTTGCA" * 5 + "\n")          # shorter than r / a single long one
    faster_reps.find_reps(p, p + ".out", True)
    assert open(p + ".out").read() == open(p).read()
    with open(p, "w") as f:
        f.write(">a\n" + "ACGTNGCA" * 5 + "\n")
    with pytest.raises(_lib.ShnError):
        faster_reps.find_reps(p, p + ".out", False)

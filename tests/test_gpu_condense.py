"""Row f3: de Bruijn load + unitig condensation on the B200 against its CPU restatement
(oracle/mbgraph_oracle.py, pinned to the real multibridging.py / mbgraph.py in
tests/test_oracle_vs_reference.py)."""
import os

import pytest

import helpers
from oracle import mbgraph_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,K,n_seqs,length", [(1, 8, 6, 120), (2, 8, 20, 200), (3, 12, 10, 300), (4, 24, 8, 400),
                                                  (5, 31, 8, 400), (6, 32, 8, 400), (7, 16, 200, 500)])
def test_condense_equals_oracle(workdir, seed, K, n_seqs, length):
    from shannon_b200 import condense
    path = helpers.debruijn_case(os.path.join(workdir, "k1mer.dict"), seed, K=K, n_seqs=n_seqs, length=length)
    exp = mbgraph_oracle.load_and_condense(path, K).snapshot()
    g = condense.load_and_condense(path, K)
    got = g.snapshot()
    assert g.n_cycle_nodes == 0
    assert got[0] == exp[0], "unitig nodes differ"
    assert got[1] == exp[1], "edges differ"
    assert any(c > 1 for _, c, _, _, _ in got[0]) and len(got[1]) > 0


def test_condense_on_pipeline_output(workdir):
    """the consumer path: component k1mer.dict files written by kmers_for_component"""
    from oracle import shannon_oracle as so
    from shannon_b200 import condense
    s1, s2 = helpers.synthetic_seqs(25, 4000, 15)
    case = helpers.make_case(workdir, 24, s1, s2)
    out, _, _, ret = helpers.run_frontend(so.extension_correction, so.kmers_for_component, case, "ora",
                                          partition_size=3)
    n = 0
    for comp in ret[1]:
        path = os.path.join(out, "component" + comp + "k1mers_allowed.dict")
        if os.path.getsize(path) == 0:
            continue
        exp = mbgraph_oracle.load_and_condense(path, 24).snapshot()
        got = condense.load_and_condense(path, 24).snapshot()
        assert got == exp, comp
        n += 1
    assert n >= 2


def test_condense_leaves_pure_cycles_uncondensed(workdir):
    """a ring of unambiguous edges has no first node: the reference's result depends on its visiting
    order; here the ring stays as single K-mer nodes and is reported"""
    from shannon_b200 import condense
    ring = "ACGTTGCAAGGCTTAACCGGTAGC"
    K = 6
    s = ring + ring[:K]
    path = os.path.join(workdir, "ring.dict")
    with open(path, "w") as f:
        for i in range(len(ring)):
            f.write("%s\t%d\n" % (s[i:i + K + 1], i + 1))
        f.write("TTTTTTA\t3\nTTTTTAC\t4\n")                 # plus one ordinary chain
    g = condense.load_and_condense(path, K)
    nodes, edges = g.snapshot()
    assert g.n_cycle_nodes == len(ring)
    assert ("TTTTTTAC", 3.0, 10.0, 3.0, 0.0) in nodes
    assert sum(1 for b, c, _, _, _ in nodes if c == 1.0) == len(ring)
    assert len(edges) == len(ring) and all(cc > 0 for _, _, _, cc in edges)

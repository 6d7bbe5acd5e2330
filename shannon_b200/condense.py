"""The first step of the per-component assembly that consumes the hot path's output (SURVEY.md 8f
row f3): ``multibridging.load_single_jellyfish`` + ``Node.condense_all`` (multibridging.py:145-172,
mbgraph.py:479-507) on the B200 -- de Bruijn nodes and edges from a per-component ``k1mer.dict``
and condensation of every unambiguous edge into unitig nodes (csrc/condense.cu)."""
import numpy as np

from .extension_correction import get_context


class CondensedGraph(object):
    """Unitig graph after condense_all: nodes[i] = (bases, count, prevalence, norm, copy_count) like the
    attributes of mbgraph.Node, edges = (source index, destination index, weight, copy_count)."""

    def __init__(self, K, raw):
        self.K = K
        self.raw = raw
        self.n_cycle_nodes = raw["n_cycle_nodes"]

    def __len__(self):
        return len(self.raw["count"])

    def node_bases(self):
        text = self.raw["bases"].tobytes().decode()
        o = self.raw["offsets"].astype(np.int64).tolist()
        return [text[o[i]:o[i + 1]] for i in range(len(o) - 1)]

    def nodes(self):
        cnt = self.raw["count"].astype(np.float64).tolist()
        prev = self.raw["prevalence"].astype(np.float64).tolist()
        return [(b, c, p, c, 0.0) for b, c, p in zip(self.node_bases(), cnt, prev)]

    def edges(self):
        return list(zip(self.raw["edge_src"].tolist(), self.raw["edge_dst"].tolist(),
                        [self.K - 1] * len(self.raw["edge_src"]),
                        self.raw["edge_copy_count"].astype(np.float64).tolist()))

    def snapshot(self):
        """(sorted node tuples, sorted edge tuples by bases) -- the comparison form of the oracle"""
        names = self.node_bases()
        return sorted(self.nodes()), sorted((names[s], names[d], w, cc) for s, d, w, cc in self.edges())


def kmer_ends(keys, k1):
    """K-mer prefix and suffix (packed, one word each) of packed K1-mers ((n,) or (n, 2) uint64)."""
    keys = np.asarray(keys, dtype=np.uint64)
    K = k1 - 1
    if keys.ndim == 1:
        mask = np.uint64((1 << (2 * K)) - 1) if K < 32 else np.uint64(0xFFFFFFFFFFFFFFFF)
        return keys >> np.uint64(2), keys & mask
    lo, hi = keys[:, 0], keys[:, 1]                      # k1 = 33: 66 bits, low word first
    return (lo >> np.uint64(2)) | (hi << np.uint64(62)), lo.copy()


def load_and_condense(edge_file, K, ctx=None):
    """load_single_jellyfish(edge_file) followed by Node.condense_all() (Read.K = K)."""
    ctx = ctx or get_context()
    keys, counts, k1 = ctx.parse_kmer_file(edge_file)
    assert K == k1 - 1                                   # multibridging.py:161
    pre, suf = kmer_ends(keys, k1)
    return CondensedGraph(K, ctx.condense(pre, suf, counts, K))

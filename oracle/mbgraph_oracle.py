"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the first step of the per-component assembly that
consumes the hot path's output: ``multibridging.load_single_jellyfish``
(multibridging.py:145-172) and ``Node.condense_all`` / ``Edge.condense`` (mbgraph.py:479-507,
184-258): de Bruijn nodes (K-mers) and edges (K1-mers) from a ``k1mer.dict`` file, then condensation
of every unambiguous edge (source has one out-edge, destination one in-edge) into unitig nodes.
Same sequential order as the reference.  Pinned to the real modules in
tests/test_oracle_vs_reference.py."""


class Node(object):
    def __init__(self, graph, bases):
        self.bases = bases
        self.in_edges = []
        self.out_edges = []
        self.norm = 1.0
        self.copy_count = 0.0
        self.prevalence = 0.0
        self.count = 1.0
        self.destroyed = False
        graph.nodes.append(self)           # mbgraph.py:341


class Edge(object):
    def __init__(self, weight, in_node, out_node):      # mbgraph.py:169-176
        self.in_node, self.out_node, self.weight = in_node, out_node, weight
        in_node.out_edges.append(self)
        out_node.in_edges.append(self)
        self.copy_count = 0.0

    def destroy(self):                                  # mbgraph.py:178-182
        self.in_node.out_edges.remove(self)
        self.out_node.in_edges.remove(self)
        self.in_node = self.out_node = None


class Graph(object):
    def __init__(self, K):
        self.K = K
        self.nodes = []

    def load_single_jellyfish(self, edge_file):
        """multibridging.py:145-172: every line `K1MER prevalence` is an edge between its K-mer prefix
        and suffix (created on first sight), weight K-1, copy_count = round(prevalence);
        node.prevalence = sum of the WEIGHTS of its out-edges."""
        by_bases = {}
        with open(edge_file) as f:
            for line in f:
                bases, prevalence = line.split()
                assert self.K == len(bases) - 1
                k1, k2 = bases[:-1], bases[1:]
                if k1 not in by_bases:
                    by_bases[k1] = Node(self, k1)
                if k2 not in by_bases:
                    by_bases[k2] = Node(self, k2)
                e = Edge(self.K - 1, by_bases[k1], by_bases[k2])
                e.copy_count = round(float(prevalence))
        for node in self.nodes:
            node.prevalence = sum(e.weight for e in node.out_edges)

    def condense(self, edge):
        """mbgraph.py:184-258 for source is not destination (the only case condense_all reaches)."""
        source, destination = edge.in_node, edge.out_node
        condensed = Node(self, source.bases + destination.bases[edge.weight:])
        condensed.count = source.count + destination.count
        condensed.prevalence = source.prevalence + destination.prevalence
        condensed.norm = source.norm + destination.norm
        condensed.copy_count = ((source.copy_count * source.norm + destination.copy_count * destination.norm)
                                / condensed.norm)
        for e in list(source.in_edges):                 # new Edge objects: copy_count starts at 0.0
            Edge(e.weight, e.in_node, condensed)
            e.destroy()
        for e in list(destination.out_edges):
            Edge(e.weight, condensed, e.out_node)
            e.destroy()
        edge.destroy()
        source.destroyed = destination.destroyed = True
        return condensed

    def condense_all(self):
        """mbgraph.py:479-498: the loop runs over the GROWING node list, so condensed nodes are
        visited again and chains collapse completely."""
        i = 0
        while i < len(self.nodes):
            n = self.nodes[i]
            i += 1
            if len(n.out_edges) != 1:
                continue
            assert not n.destroyed
            e = n.out_edges[0]
            if len(e.out_node.in_edges) == 1 and n is not e.out_node:
                self.condense(e)
        self.nodes = [n for n in self.nodes if not n.destroyed]

    def snapshot(self):
        """(sorted node tuples, sorted edge tuples) for comparison"""
        nodes = sorted((n.bases, float(n.count), float(n.prevalence), float(n.norm), float(n.copy_count))
                       for n in self.nodes)
        edges = sorted((n.bases, e.out_node.bases, int(e.weight), float(e.copy_count))
                       for n in self.nodes for e in n.out_edges)
        return nodes, edges


def load_and_condense(edge_file, K):
    g = Graph(K)
    g.load_single_jellyfish(edge_file)
    g.condense_all()
    return g

"""Re-weighting of the METIS graph for the second gpmetis pass (the reference's
weight_updated_graph.py:9-44, same signature): every edge cut by the first partition gets its
weight multiplied by ``penalty``.  A text rewrite of one small contig-level graph file, only
reached for components with more than ``partition_size`` contigs; it stays on the host."""


def weight_updated_graph(directory, partition_file, og_graph_file, new_graph_file, contig_file,
                         new_contig_file, penalty=5, randomize=False):
    if randomize:
        raise NotImplementedError("randomize=True is never used by kmers_for_component.py:229-231")
    with open(directory + og_graph_file, 'r') as f:
        graph = f.readlines()
    with open(directory + partition_file, 'r') as f:
        part = [int(x) for x in f.readlines()]
    with open(directory + new_graph_file, 'w') as out:
        out.write(graph[0])
        for node, line in enumerate(graph[1:]):
            tokens = line.split()
            pieces = []
            for j in range(0, len(tokens) - 1, 2):
                weight = tokens[j + 1]
                if part[node] != part[int(tokens[j]) - 1]:
                    weight = str(penalty * int(weight))
                pieces.append(tokens[j] + "\t" + weight + "\t")
            out.write("".join(pieces) + "\n")

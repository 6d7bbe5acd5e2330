"""TEST INFRASTRUCTURE ONLY -- CPU oracle for Shannon's k-mer front end.

A from-scratch Python restatement of the algorithm of the reference's two hot-path modules,
``extension_correction.py`` and ``kmers_for_component.py`` (citations are file:line in
/root/reference).  It exists so the CUDA path can be checked for bit-exact parity on a box
that has no copy of the reference.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
(``shannon_b200/``) never does.

Parity pin: the reference has no tests or golden vectors for this path,
so this oracle is pinned against *outputs of the reference itself*:
``tests/test_oracle_vs_reference.py`` runs the real reference (``oracle/ref_loader.py``, py3
patch in memory) beside this file on the bundled samples and on seeded random inputs and
compares every written file and returned object; ``tests/golden/`` holds fixtures generated
from the real reference by ``tests/golden/make_golden.py`` that travel to the GPU box.

Ordering semantics are those of insertion-ordered dicts (CPython >= 3.7), see SURVEY 8c.

The data model is kept deliberately close to the reference's (str keys in dicts, one Python
loop per k-mer): besides being the checker it is the "port" CPU baseline that bench.py times,
so it has to cost what the reference's own Python costs.
"""
import math
import os
import time

# successor tie-break order, extension_correction.py:10 (NOT alphabetical)
BASES = ("A", "G", "C", "T")
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
R_MER = 15  # extension_correction.py:357


def reverse_complement(s):
    """extension_correction.py:30."""
    return "".join(_COMP[b] for b in reversed(s))


def low_complexity(kmer):
    """extension_correction.py:142-149: within Hamming distance 2 of a homopolymer."""
    n = len(kmer)
    return max(kmer.count("A"), kmer.count("C"), kmer.count("G"), kmer.count("T")) >= n - 2


def load_kmers(path, double_stranded, polyA_del=True):
    """extension_correction.py:202-221.  Returns (insertion-ordered dict kmer->int, K1).

    Weights are integer counts (the reference parses them with float(); integer-valued
    floats < 2**53 add exactly, so ints are the same numbers)."""
    kmers = {}
    with open(path) as f:
        for line in f:
            kmer, weight = line.split()
            kmer = kmer.upper()
            if polyA_del and low_complexity(kmer):
                continue
            w = int(float(weight))
            kmers[kmer] = kmers.get(kmer, 0) + w
            if double_stranded:
                rc = reverse_complement(kmer)
                kmers[rc] = kmers.get(rc, 0) + w
    k1 = len(next(iter(kmers)))
    return kmers, k1


def seed_order(kmers):
    """extension_correction.py:334,343-344: stable ascending sort by weight, consumed from
    the end => heaviest first, ties: later-inserted key first."""
    order = sorted(kmers.items(), key=lambda kv: kv[1])
    order.reverse()
    return order


def _walk_one_side(seed, kmers, traversed, right):
    """extension_correction.py:223-245 (extend / extend_right / extend_left / argmax)."""
    last = seed[1:] if right else seed[:-1]
    bases = []
    tot_w = 0
    while True:
        best_b = None
        best_w = None
        for b in BASES:
            cand = last + b if right else b + last
            w = kmers.get(cand)
            if w is None or cand in traversed:
                continue
            if best_b is None or w > best_w:  # strict '>' keeps the first of equals
                best_b, best_w = b, w
        if best_b is None:
            return bases, tot_w
        node = last + best_b if right else best_b + last
        traversed.add(node)
        bases.append(best_b)
        tot_w += best_w
        last = node[1:] if right else node[:-1]


class Walk(object):
    __slots__ = ("rank", "seed", "contig", "n_left", "n_right", "tot_wt", "tot_kmer",
                 "passes_shape", "duplicate", "accepted", "contig_index")

    def as_tuple(self):
        return (self.seed, self.contig, self.n_left, self.n_right, self.tot_wt, self.tot_kmer,
                self.passes_shape, self.duplicate, self.accepted)


def greedy_walks(kmers, min_weight):
    """The seed loop of run_correction without the accept/index part
    (extension_correction.py:343-354).  Returns the list of started walks in pop order and
    the final traversed set.  ``rank`` is the position in seed_order()."""
    traversed = set()
    walks = []
    for rank, (seed, w) in enumerate(seed_order(kmers)):
        if w < min_weight:
            break
        if seed in traversed:
            continue
        traversed.add(seed)
        right, wr = _walk_one_side(seed, kmers, traversed, True)
        left, wl = _walk_one_side(seed, kmers, traversed, False)
        wk = Walk()
        wk.rank = rank
        wk.seed = seed
        wk.n_left, wk.n_right = len(left), len(right)
        wk.tot_wt = wr + wl + w
        wk.tot_kmer = len(left) + len(right) + 1
        wk.contig = "".join(reversed(left)) + seed + "".join(right)
        walks.append(wk)
    return walks, traversed


def passes_shape(length, tot_wt, tot_kmer, min_weight, min_length):
    """The length and 'hyperbola' terms of extension_correction.py:353,361 with the same
    float expression order."""
    avg_wt = float(tot_wt) / max(1, tot_kmer)
    return (length >= min_length and
            length * math.pow(avg_wt, 1 / 4.0) >= 2 * min_length * math.pow(min_weight, 1 / 4.0))


def duplicate_check(contig, rmer_to_contig, r=R_MER):
    """extension_correction.py:247-270 (f = 0.5)."""
    n = len(contig)
    count = {}
    top = 0
    best = -1
    for i in range(n - r + 1):
        for dup in rmer_to_contig.get(contig[i:i + r], ()):
            c = count.get(dup, 0) + 1
            count[dup] = c
            if c >= top:
                top, best = c, dup
    covered = [False] * n
    for i in range(n - r + 1):
        lst = rmer_to_contig.get(contig[i:i + r])
        if lst is not None and best in lst:
            for j in range(i, i + r):
                covered[j] = True
    return 2 * sum(covered) > n


def accept_walks(walks, k1, min_weight, min_length):
    """The accept / index half of the seed loop (extension_correction.py:356-397) for walks given
    in pop order: duplicate_check against the accepted contigs so far, the length + hyperbola terms,
    allowed K1-mers, contig C-mer graph, r-mer index.  Returns (contigs 1-based, connections,
    allowed as an ordered dict)."""
    rmer_to_contig = {}
    cmer_to_contig = {}
    connections = {}            # contig index -> {neighbour: weight}, insertion ordered
    contigs = [None]            # 1-based, extension_correction.py:341
    allowed = {}                # ordered set
    c_len = k1 - 1
    for wk in walks:
        contig = wk.contig
        wk.duplicate = duplicate_check(contig, rmer_to_contig)          # :358
        wk.passes_shape = passes_shape(len(contig), wk.tot_wt, wk.tot_kmer,
                                       min_weight, min_length)
        wk.accepted = wk.passes_shape and not wk.duplicate               # :361
        wk.contig_index = 0
        if not wk.accepted:
            continue
        idx = len(contigs)
        wk.contig_index = idx
        contigs.append(contig)
        mine = connections.setdefault(idx, {})
        for i in range(len(contig) - k1 + 1):                            # :368-369
            allowed[contig[i:i + k1]] = None
        for i in range(len(contig) - c_len + 1):                         # :375-389
            cmer = contig[i:i + c_len]
            lst = cmer_to_contig.get(cmer)
            if lst is None:
                lst = cmer_to_contig[cmer] = []
            else:
                for other in lst:
                    if other == idx:
                        continue
                    mine[other] = mine.get(other, 0) + 1
                    theirs = connections[other]
                    theirs[idx] = theirs.get(idx, 0) + 1
            lst.append(idx)
        for i in range(len(contig) - R_MER + 1):                         # :393-397
            rmer_to_contig.setdefault(contig[i:i + R_MER], []).append(idx)
    return contigs, connections, allowed


class CorrectionResult(object):
    """Everything run_correction computes, kept for parity checks."""
    pass


def run_correction(infile, outfile, min_weight, min_length, double_stranded,
                   comp_directory_name, comp_size_threshold, polyA_del=True, inMem=False,
                   write_files=True):
    """extension_correction.py:309-524.  Writes the same files and returns a
    CorrectionResult (``.allowed_kmer_dict`` and ``.reads`` are the reference's return)."""
    res = CorrectionResult()
    log = []
    log.append("{:s}: Starting..".format(time.asctime()))
    kmers, k1 = load_kmers(infile, double_stranded, polyA_del)
    res.kmers, res.k1 = kmers, k1
    log.append("{:s}: {:d} K-mers loaded.".format(time.asctime(), len(kmers)))
    log.append("{:s}: Reads loading in background process.".format(time.asctime()))

    walks, traversed = greedy_walks(kmers, min_weight)
    res.walks, res.traversed = walks, traversed

    contigs, connections, allowed = accept_walks(walks, k1, min_weight, min_length)
    res.contigs = contigs
    res.connections = connections
    log.append("{:s}: {:d} K-mers remaining after error correction. ".format(
        time.asctime(), len(allowed)))

    res.allowed_kmer_dict = dict((k, int(kmers[k])) for k in allowed)    # :404-408
    log.append("{:s}: {:d} K-mers written to file. ".format(time.asctime(), len(allowed)))

    # connected components by iterative DFS, extension_correction.py:417-434
    log.append(str(time.asctime()) + ": Before dfs ")
    contig2component = {}
    component2contig = {}
    seen = set()
    for root in connections:
        if root in contig2component:
            continue
        members = component2contig[root] = []
        stack = [root]
        seen.add(root)
        while stack:
            cur = stack.pop()
            contig2component[cur] = root
            members.append(cur)
            for nb in connections[cur]:
                if nb not in seen:
                    stack.append(nb)
                    seen.add(nb)
    res.contig2component, res.component2contig = contig2component, component2contig
    log.append(str(time.asctime()) + ": After dfs ")

    # distinct undirected edges per component, extension_correction.py:439-450
    n_edges = dict((c, 0) for c in component2contig)
    for a in connections:
        for b in connections[a]:
            if a < b:
                n_edges[contig2component[a]] += 1
    log.append(str(time.asctime()) + ": After Edges Loaded ")

    # file packing, extension_correction.py:458-513
    files = {}
    files[outfile + "_contig"] = "".join(c + "\n" for c in contigs[1:])
    if inMem:
        files[outfile] = ""
    else:
        files[outfile] = "".join("{:s}\t{:d}\n".format(k, w)
                                 for k, w in res.allowed_kmer_dict.items())
    singles = []
    remaining = [[]]
    big = []
    cur_size = 0
    for comp, members in component2contig.items():
        if len(members) == 1:
            singles.append(contigs[members[0]])
            continue
        if len(members) > comp_size_threshold:
            code = dict((c, i + 1) for i, c in enumerate(members))
            lines = [str(len(members)) + "\t" + str(n_edges[comp]) + "\t001\n"]
            for c in members:
                lines.append("".join(str(code[nb]) + "\t" + str(wt) + "\t"
                                     for nb, wt in connections[c].items()) + "\n")
            big.append(("".join(lines), "".join(contigs[c] + "\n" for c in members)))
        else:
            remaining[-1].extend(contigs[c] for c in members)
            cur_size += len(members)
            if cur_size > comp_size_threshold:
                remaining.append([])
                cur_size = 0
    d = comp_directory_name
    files[d + "/reconstructed_single_contigs.fasta"] = "".join(
        ">Single_" + str(j) + "\n" + c + "\n" for j, c in enumerate(singles))
    for m, lst in enumerate(remaining):
        files[d + "/remaining_contigs" + str(m + 1) + ".txt"] = "".join(c + "\n" for c in lst)
    for n, (graph, ctg) in enumerate(big):
        files[d + "/component" + str(n + 1) + ".txt"] = graph
        files[d + "/component" + str(n + 1) + "contigs.txt"] = ctg
    res.singles, res.remaining, res.big = singles, remaining, big
    log.append(str(time.asctime()) + ": Metis Input File Created ")
    log.append("{:s}: Read-loader in background process joinig back.".format(time.asctime()))
    log.append("{:s}: {:d} Reads loaded in background process.".format(time.asctime(), 0))
    files[d + "/before_sp_log.txt"] = "".join(l + "\n" for l in log)
    res.files = files
    res.reads = []
    if write_files:
        for path, text in files.items():
            with open(path, "w") as f:
                f.write(text)
    return res


def extension_correction(arguments, inMem=False):
    """extension_correction.py:528-549: same argv contract, same return value."""
    double_stranded = "-d" in arguments
    arguments = [a for a in arguments if len(a) > 0 and a[0] != "-"]
    infile, outfile = arguments[:2]
    min_weight, min_length = int(arguments[2]), int(arguments[3])
    comp_directory_name, comp_size_threshold = arguments[4], int(arguments[5])
    res = run_correction(infile, outfile, min_weight, min_length, double_stranded,
                         comp_directory_name, comp_size_threshold, True, inMem)
    return res.allowed_kmer_dict, res.reads


# ----------------------------------------------------------------------------------------
# L4: kmers_for_component.py
# ----------------------------------------------------------------------------------------

def weight_updated_graph(directory, partition_file, og_graph_file, new_graph_file, penalty=5):
    """weight_updated_graph.py:9-44 (randomize=False branch, the only one the path uses):
    multiply the weight of every edge cut by the first partition by ``penalty``."""
    with open(directory + og_graph_file) as f:
        graph = f.readlines()
    with open(directory + partition_file) as f:
        part = [int(x) for x in f.readlines()]
    out = [graph[0]]
    for node, line in enumerate(graph[1:]):
        tok = line.split()
        new = ""
        for j in range(0, len(tok) - 1, 2):
            nb, wt = tok[j], tok[j + 1]
            if part[node] != part[int(nb) - 1]:
                wt = str(penalty * int(wt))
            new += nb + "\t" + wt + "\t"
        out.append(new + "\n")
    with open(directory + new_graph_file, "w") as f:
        f.writelines(out)


def sample_k1mers(read, k1):
    """kmers_for_component.py:186-192 (get_rmers)."""
    out = []
    i = 0
    while i < len(read) - k1:
        out.append(read[i:i + k1])
        i += k1
    out.append(read[-k1:])
    return out


def read_components(read, k1mers2component, k1):
    """kmers_for_component.py:194-202: UNION of the component sets of all sampled hits."""
    comps = set()
    for km in sample_k1mers(read, k1):
        hit = k1mers2component.get(km)
        if hit is not None:
            comps |= hit[0]
    return comps


def _count_files(pattern):
    n = 0
    while os.path.exists(pattern % (n + 1)):
        n += 1
    return n


def _read_records(handles, NR, counter):
    """The chunk reader of kmers_for_component.py:330-339 / :369-380 for 1 or 2 files.
    Yields (chunk, more) where chunk is a list of tuples of mates."""
    stop = False
    while not stop:
        chunk = []
        while True:
            names = [h.readline()[:-1] for h in handles]
            if not names[0]:
                stop = True
                break
            mates = tuple(h.readline()[:-1] for h in handles)
            if any(m.strip("ACTG") for m in mates):
                continue
            counter[0] += 1
            chunk.append(mates)
            if not mates[0]:
                stop = True
                break
            if counter[0] % NR == 0:
                break
        yield chunk


def build_component_map(directory_name, contig_file_extension, components_broken,
                        n_remaining, k1mer_dictionary, k1, repartition):
    """kmers_for_component.py:239-305.  Returns (new_components, k1mers2component)."""
    new_components = {}
    k1mers2component = {}

    def add(comp, contig):
        new_components.setdefault(comp, []).append(contig)
        for p in range(len(contig) - k1 + 1):
            km = contig[p:p + k1]
            hit = k1mers2component.get(km)
            if hit is None:
                k1mers2component[km] = [set([comp]), k1mer_dictionary.get(km, 0)]
            else:
                hit[0].add(comp)

    for i in components_broken:
        base = directory_name + "/component" + str(i + 1)
        with open(base + contig_file_extension) as f:
            contig_lines = f.readlines()
        passes = [("c", ".txt.part.")]
        if repartition:
            passes.append(("r2_c", "r2.txt.part."))
        for prefix, ext in passes:
            with open(base + ext + str(components_broken[i])) as f:
                for j, line in enumerate(f):
                    comp = prefix + str(i + 1) + "_" + line.split()[0]
                    add(comp, contig_lines[j].split()[0])
    for i in range(n_remaining):
        with open(directory_name + "/remaining_contigs" + str(i + 1) + ".txt") as f:
            for line in f.readlines():
                add("cremaining" + str(i + 1), line.split()[0])
    return new_components, k1mers2component


def kmers_for_component(k1mer_dictionary, kmer_directory, reads, reads_files, directory_name,
                        contig_file_extension, get_partition_k1mers, double_stranded=True,
                        paired_end=False, repartition=False, partition_size=500, overload=1.5,
                        K=24, gpmetis_path="gpmetis", penalty=5, only_reads=False, inMem=False,
                        nJobs=1, NR=10000000):
    """kmers_for_component.py:144-558 restated.  ``double_stranded=True`` (in-process RC
    fan-out, :36-52,117-141,341,381; never passed by shannon.py:427) is restated for nJobs = 1,
    where the reference is deterministic: every chunk of valid reads is followed by its reverse
    complements (SE), resp. the pairs (r1, rc(r2)) by the pairs (r2, rc(r1)) (PE).  With nJobs > 1
    the reference concatenates the workers' results in queue-arrival order."""
    if not get_partition_k1mers:
        return None  # the reference falls off the end of the function (:207)
    k1 = K + 1
    log_path = directory_name + "/before_sp_log.txt"
    f_log = open(log_path, "a" if os.path.exists(log_path) else "w")

    def write_log(s):
        f_log.write(s + "\n")

    n_components = _count_files(directory_name + "/component%dcontigs.txt")
    n_remaining = _count_files(directory_name + "/remaining_contigs%d.txt")

    ufactor = int(1000.0 * overload - 1000.0)
    components_broken = {}
    summary = ""
    for i in range(n_components):                                        # :213-234
        base = directory_name + "/component" + str(i + 1)
        with open(base + contig_file_extension) as f:
            num_contigs = len(f.readlines())
        parts = min(int(math.ceil(float(num_contigs) / float(partition_size))), 100)
        components_broken[i] = parts
        summary += "Component " + str(i) + ": " + str(parts) + " partitions, "
        if num_contigs >= 2:
            os.system(gpmetis_path + " -ufactor=" + str(ufactor) + " " + base + ".txt " + str(parts))
            if repartition:
                write_log(str(time.asctime()) + ": Creating graph for repartition ")
                weight_updated_graph(directory_name,
                                     "/component" + str(i + 1) + ".txt.part." + str(parts),
                                     "/component" + str(i + 1) + ".txt",
                                     "/component" + str(i + 1) + "r2.txt", penalty)
                write_log(str(time.asctime()) + ": Created graph for repartition ")
                os.system(gpmetis_path + " -ufactor=" + str(ufactor) + " " + base + "r2.txt " + str(parts))
    write_log(str(time.asctime()) + ": gpmetis for partitioning is complete \n " + summary)

    new_components, k1mers2component = build_component_map(
        directory_name, contig_file_extension, components_broken, n_remaining,
        k1mer_dictionary, k1, repartition)
    write_log(str(time.asctime()) + ": k1mers2component dictionary created ")

    # read partition, :322-423
    n_files = 2 if paired_end else 1
    part = [dict((c, []) for c in new_components) for _ in range(n_files)]
    offset = dict((c, 0) for c in new_components)
    handles = [open(reads_files[m]) for m in range(n_files)]
    suffix = ["_1", "_2"] if paired_end else [""]
    counter = [0]
    for chunk in _read_records(handles, NR, counter):
        if double_stranded:                                              # :341,381 with nJobs = 1
            if paired_end:
                chunk = [(a, reverse_complement(b.strip())) for a, b in chunk] + \
                        [(b, reverse_complement(a.strip())) for a, b in chunk]
            else:
                chunk = list(chunk) + [(reverse_complement(m[0].strip()),) for m in chunk]
        for mates in chunk:
            assigned = set()
            for m in mates:
                assigned |= read_components(m, k1mers2component, k1)
            for comp in assigned:
                for m in range(n_files):
                    part[m][comp].append(mates[m])
        if not inMem:
            for comp in new_components:
                for m in range(n_files):
                    with open(directory_name + "/reads" + str(comp) + suffix[m] + ".fasta", "a") as f:
                        f.write("".join(">" + str(e + offset[comp]) + suffix[m] + "\n" + r + "\n"
                                        for e, r in enumerate(part[m][comp])))
                offset[comp] += len(part[0][comp])
                for m in range(n_files):
                    part[m][comp][:] = []
    for h in handles:
        h.close()
    write_log(str(time.asctime()) + ": reads partititoned ")

    contig_weights = {}
    if not only_reads:                                                   # :452-477
        write_log(str(time.asctime()) + ": Writing k1mers to file")
        for comp in new_components:
            contig_weights[comp] = []
            with open(directory_name + "/component" + comp + "k1mers_allowed.dict", "w") as f:
                for contig in new_components[comp]:
                    wl = [k1mers2component[contig[p:p + k1]][1]
                          for p in range(len(contig) - k1 + 1)]
                    if inMem:
                        contig_weights[comp].append(wl)
                    else:
                        f.write("".join(contig[p:p + k1] + "\t" + str(w) + "\n"
                                        for p, w in enumerate(wl)))
        write_log(str(time.asctime()) + ": k1mers written to file ")
    write_log(str(time.asctime()) + ": kmers written to file \n")
    f_log.close()

    if inMem:
        new_comps = new_components
        if paired_end:
            rps = dict((c, [[part[0][c]], [part[1][c]]]) for c in new_components)
        else:
            rps = dict((c, [part[0][c]]) for c in new_components)
    else:
        new_comps = list(new_components)
        contig_weights = []
        rps = {}
    return [components_broken, new_comps, contig_weights, rps]

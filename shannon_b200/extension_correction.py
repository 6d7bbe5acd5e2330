"""Drop-in replacement of the reference's ``extension_correction.py`` (same entry point, argv
contract, return value and output files; extension_correction.py:528-549, 309-524) running on
the B200 through libshannon_b200.so.  Host code here only parses arguments, orders contig-level
results (the DFS of :417-434 over the GPU-built contig graph) and writes the text files.
"""
import os
import sys
import time
from collections.abc import Mapping

import numpy as np

from . import _lib
from .pipeline import (contig_adjacency, correct, decode_kmers, dfs_components,  # noqa: F401
                       encode_kmer, keys_array, keys_as_ints, pack_components)

_default_ctx = None
LAST_TIMINGS = {}     # wall-clock seconds of the sections of the last run_correction call


def get_context(device=None):
    """Process-wide default GPU context (device from SHANNON_B200_DEVICE / LOCAL_RANK / 0)."""
    global _default_ctx
    if _default_ctx is None:
        if device is None:
            device = int(os.environ.get("SHANNON_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _default_ctx = _lib.Context(device)
    return _default_ctx


class AllowedKmerDict(Mapping):
    """``allowed_kmer_dict`` (extension_correction.py:404-408) backed by the packed arrays the
    GPU produced.  Behaves like the reference's ``dict[str, int]`` (iteration in contig order,
    ``get``, ``clear``) and lets ``kmers_for_component`` pick up the arrays without a detour
    through millions of Python strings."""

    def __init__(self, keys, weights, k1):
        self.keys_packed = np.asarray(keys, dtype=np.uint64)
        self.weights = np.asarray(weights, dtype=np.uint32)
        self.k1 = k1
        self._index = None

    def _idx(self):
        if self._index is None:
            self._index = dict(zip(keys_as_ints(self.keys_packed), self.weights.tolist()))
        return self._index

    def __len__(self):
        return len(self.keys_packed)

    def __iter__(self):
        if len(self.keys_packed) == 0:
            return iter(())
        mat = decode_kmers(self.keys_packed, self.k1)
        return iter(mat.tobytes().decode()[i:i + self.k1]
                    for i in range(0, mat.size, self.k1))

    def __getitem__(self, kmer):
        if not isinstance(kmer, str) or len(kmer) != self.k1 or kmer.strip("ACGT"):
            raise KeyError(kmer)
        return self._idx()[encode_kmer(kmer)]

    def clear(self):  # shannon.py:469
        self.keys_packed = self.keys_packed[:0]
        self.weights = np.empty(0, dtype=np.uint32)
        self._index = None


def run_correction(infile, outfile, min_weight, min_length, double_stranded,
                   comp_directory_name, comp_size_threshold, polyA_del=True, inMem=False,
                   nJobs=1, reads_files=(), ctx=None):
    """extension_correction.py:309-524.  Returns (allowed_kmer_dict, reads)."""
    if not polyA_del:
        raise NotImplementedError("polyA_del=False is never used by the reference driver")
    ctx = ctx or get_context()
    print('nJobs:' + str(nJobs))
    print('reads_files:' + ' '.join(reads_files))
    f_log = open(comp_directory_name + "/before_sp_log.txt", 'w')
    print("{:s}: Starting Kmer error correction..".format(time.asctime()))
    f_log.write("{:s}: Starting..".format(time.asctime()) + "\n")

    LAST_TIMINGS.clear()
    t0 = time.perf_counter()
    keys, counts, k1 = ctx.parse_kmer_file(infile)
    LAST_TIMINGS["ec_parse_k1mer_file"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cor = correct(ctx, keys, counts, k1, double_stranded, min_weight, min_length)
    LAST_TIMINGS["ec_gpu_and_ordering"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    del keys, counts
    print("{:s}: {:d} K-mers loaded.".format(time.asctime(), cor.n_loaded))
    f_log.write("{:s}: {:d} K-mers loaded.".format(time.asctime(), cor.n_loaded) + "\n")
    f_log.write("{:s}: Reads loading in background process.".format(time.asctime()) + "\n")

    contigs = cor.contigs
    with open(outfile + '_contig', 'w') as f1:
        f1.write("".join(c + "\n" for c in contigs.strings()))

    a_keys, a_w = cor.allowed_keys, cor.allowed_weights
    allowed_kmer_dict = AllowedKmerDict(a_keys, a_w, k1)
    n_allowed = len(allowed_kmer_dict)
    print("{:s}: {:d} K-mers remaining after error correction.".format(time.asctime(), n_allowed))
    f_log.write("{:s}: {:d} K-mers remaining after error correction.".format(
        time.asctime(), n_allowed) + " \n")
    with open(outfile, 'w') as f:
        if not inMem and n_allowed:
            mat = decode_kmers(a_keys, k1)
            lines = np.empty((n_allowed, k1 + 1), dtype=np.uint8)
            lines[:, :k1] = mat
            lines[:, k1] = ord("\t")
            txt = lines.tobytes().decode()
            f.write("".join(txt[i:i + k1 + 1] + str(w) + "\n"
                            for i, w in zip(range(0, len(txt), k1 + 1), a_w.tolist())))
    f_log.write("{:s}: {:d} K-mers written to file.".format(time.asctime(), n_allowed) + " \n")
    f_log.write(str(time.asctime()) + ": " + "Before dfs " + "\n")
    component2contig, n_edges = cor.component2contig, cor.n_edges
    f_log.write(str(time.asctime()) + ": " + "After dfs " + "\n")
    f_log.write(str(time.asctime()) + ": " + "After Edges Loaded " + "\n")

    # file packing of :458-513 (singles / METIS graphs of oversized components / remaining files)
    d = comp_directory_name
    pk = pack_components(cor, comp_size_threshold)
    with open(d + "/reconstructed_single_contigs.fasta", 'w') as f:
        f.write("".join('>Single_' + str(j) + '\n' + contigs[c] + '\n' for j, c in enumerate(pk.singles)))
    for n, (component, members) in enumerate(pk.big):
        code = dict((c, i + 1) for i, c in enumerate(members))
        with open(d + "/component" + str(n + 1) + ".txt", 'w') as f:
            f.write(str(len(members)) + "\t" + str(n_edges[component]) + "\t" + "001" + "\n")
            for c in members:
                f.write("".join(str(code[nb]) + "\t" + str(wt) + "\t"
                                for nb, wt in cor.neighbours(c)) + "\n")
        with open(d + "/component" + str(n + 1) + "contigs" + ".txt", 'w') as f:
            f.write("".join(contigs[c] + "\n" for c in members))
    for m, group in enumerate(pk.remaining):
        with open(d + "/remaining_contigs" + str(m + 1) + ".txt", 'w') as f:
            f.write("".join(contigs[c] + "\n" for c in group))
    LAST_TIMINGS["ec_write_files"] = time.perf_counter() - t0
    f_log.write(str(time.asctime()) + ": " + "Metis Input File Created " + "\n")
    f_log.write("{:s}: Read-loader in background process joinig back.".format(time.asctime()) + "\n")
    reads = []
    f_log.write("{:s}: {:d} Reads loaded in background process.".format(time.asctime(), len(reads)) + "\n")
    f_log.close()
    return allowed_kmer_dict, reads


def extension_correction(arguments, inMem=False):
    """Same argv contract as extension_correction.py:528-549:
    [-d] infile outfile min_weight min_length comp_dir comp_size_threshold [nJobs [reads1 [reads2]]]"""
    double_stranded = '-d' in arguments
    arguments = [a for a in arguments if len(a) > 0 and a[0] != '-']
    infile, outfile = arguments[:2]
    min_weight, min_length = int(arguments[2]), int(arguments[3])
    comp_directory_name, comp_size_threshold = arguments[4], int(arguments[5])
    nJobs = int(arguments[6]) if len(arguments) > 6 else 1
    reads_files = list(arguments[7:9]) if len(arguments) > 7 else []
    return run_correction(infile, outfile, min_weight, min_length, double_stranded,
                          comp_directory_name, comp_size_threshold, True, inMem, nJobs, reads_files)


if __name__ == '__main__':
    if len(sys.argv) == 1:
        argv = ['kmers.dict', 'allowed_kmers.dict', '1', '1', '-d']
    else:
        argv = sys.argv[1:]
    extension_correction(argv)

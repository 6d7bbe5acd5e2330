"""Test-side numpy twin of the four array primitives shannon_b200.dist needs (GpuOps), so the
exchange protocol of the sharded table can run under gloo on CPU.  Test infrastructure only."""
import numpy as np
import torch

from shannon_b200 import synth


def owner_of(keys_u64, world):
    lo = synth.mix64(keys_u64) & np.uint64(0xFFFFFFFF)
    return ((lo * np.uint64(world)) >> np.uint64(32)).astype(np.int64)


class NumpyOps(object):
    def __init__(self):
        self.table = {}

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype)

    def sync(self):
        pass

    def plan(self, keys, world):
        own = owner_of(keys.numpy().view(np.uint64), world)
        perm = np.argsort(own, kind="stable").astype(np.int32)
        counts = np.bincount(own, minlength=world).tolist()
        return torch.from_numpy(perm), counts

    def gather(self, src, perm):
        return src[perm.long()].contiguous()

    def scatter(self, src, perm):
        out = torch.empty_like(src)
        out[perm.long()] = src
        return out

    def build(self, keys, counts, line_idx, k1):
        self.table = {}
        for k, c, l in zip(keys.numpy().view(np.uint64).tolist(), counts.tolist(), line_idx.tolist()):
            e = self.table.setdefault(k, [0, l])
            e[0] += c
            e[1] = min(e[1], l)

    def lookup(self, keys):
        ks = keys.numpy().view(np.uint64).tolist()
        w = torch.tensor([self.table.get(k, [0])[0] for k in ks], dtype=torch.int32)
        f = torch.tensor([int(k in self.table) for k in ks], dtype=torch.uint8)
        return w, f


class OracleL4Ctx(object):
    """Test-side stand-in for the four L4 calls of a shn context, backed by the oracle's
    read-assignment functions: lets partition_reads_sharded run under gloo on CPU."""

    def __init__(self):
        from oracle import shannon_oracle as so
        self.so = so
        self.k2c = {}
        self.mates = {}
        self.lists = []

    def l4_map_add_contigs(self, bases, offsets, comp_of_contig, k1, reset, expected_total):
        if reset:
            self.k2c = {}
        text = np.asarray(bases, dtype=np.uint8).tobytes().decode()
        n = 0
        for c in range(len(comp_of_contig)):
            if int(comp_of_contig[c]) == 0xFFFFFFFF:
                continue
            s = text[int(offsets[c]):int(offsets[c + 1])]
            for p in range(len(s) - k1 + 1):
                self.k2c.setdefault(s[p:p + k1], [set(), 0])[0].add(int(comp_of_contig[c]))
                n += 1
        assert n <= expected_total

    def l4_load_reads(self, mate, bases, offsets, n=None, on_device=False):
        text = np.asarray(bases, dtype=np.uint8).tobytes().decode()
        self.mates[mate] = [text[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)]

    def l4_assign(self, paired, k1):
        files = [self.mates[0]] + ([self.mates[1]] if paired else [])
        per_comp = {}
        n_assign = n_valid = n_look = 0
        for r in range(len(files[0])):
            ms = [f[r] for f in files]
            if any(m.strip("ACTG") for m in ms):
                continue
            n_valid += 1
            comps = set()
            for m in ms:
                comps |= self.so.read_components(m, self.k2c, k1)
                n_look += len(self.so.sample_k1mers(m, k1)) if len(m) >= k1 else 0
            for c in comps:
                per_comp.setdefault(c, []).append(r)
                n_assign += 1
        self.per_comp = per_comp
        return n_assign, n_look, n_valid

    def l4_assignments(self, n_comps, n_assign):
        offs = np.zeros(n_comps + 1, dtype=np.uint64)
        idx = []
        for c in range(n_comps):
            idx += self.per_comp.get(c, [])
            offs[c + 1] = len(idx)
        return offs, np.asarray(idx, dtype=np.uint32)


def partition_case(seed=3, n_tx=6, n_pairs=400):
    """Contigs with component ids + paired read files (host arrays) for the sharded partition tests:
    the contigs are the transcripts themselves, two per component; some reads are dirty."""
    import helpers
    from shannon_b200 import synth
    tx = synth.make_transcripts(n_tx, seed)
    s1, s2 = helpers.synthetic_seqs(n_tx, n_pairs, seed)
    s1[3] = s1[3][:40] + "N" + s1[3][41:]
    s2[7] = s2[7][:20]
    s1[11] = s1[11][:25]

    def arrays(seqs):
        b = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
        o = np.zeros(len(seqs) + 1, dtype=np.uint64)
        o[1:] = np.cumsum([len(s) for s in seqs])
        return b, o
    codes = "AGCT"
    contigs = ["".join(codes[int(c)] for c in t) if not isinstance(t, str) else t for t in tx]
    comp = np.asarray([i // 2 for i in range(len(contigs))], dtype=np.uint32)
    comp[-1] = 0xFFFFFFFF        # a single-contig component is not partitioned
    cb, co = arrays(contigs)
    return (cb, co, comp), int(comp[:-1].max()) + 1, [arrays(s1), arrays(s2)]

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


# ---- CPU twin of sharded.GpuOps (the whole-path protocol under gloo) ------------------------------
_M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = np.asarray(x, dtype=np.uint64) & _M32
    x = (x * np.uint64(0x9E3779B1)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x85EBCA77)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE3D)) & _M32
    x ^= x >> np.uint64(16)
    return x


def minimizer_owner(keys, k1, world, m=11):
    """numpy twin of owner_of() in csrc/shard.cu (one-word keys, k1 > m)."""
    keys = np.asarray(keys, dtype=np.uint64)
    assert k1 > m
    mmask = np.uint64((1 << (2 * m)) - 1)
    best = np.full(len(keys), 0xFFFFFFFF, dtype=np.uint64)
    for p in range(0, k1 - m + 1):
        best = np.minimum(best, _mix32((keys >> np.uint64(2 * p)) & mmask))
    return ((_mix32(best ^ np.uint64(0x5BD1E995)) * np.uint64(world)) >> np.uint64(32)).astype(np.int64)


class OracleShardOps(object):
    """Every per-rank primitive sharded.correct_sharded needs, restated with numpy and the oracle's
    walk / accept functions on CPU tensors (records = int64 rows [key, payload])."""

    def __init__(self):
        from oracle import shannon_oracle as so
        self.so = so
        self.device = torch.device("cpu")

    # -- helpers
    @staticmethod
    def _u(t):
        return t.numpy().view(np.uint64)

    def _kstr(self, key):
        return "".join("AGCT"[(key >> (2 * (self.k1 - 1 - i))) & 3] for i in range(self.k1))

    def _kint(self, s):
        x = 0
        for ch in s:
            x = (x << 2) | "AGCT".index(ch)
        return x

    @staticmethod
    def _group(dest, rows, world):
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=world).tolist()
        return torch.from_numpy(np.ascontiguousarray(rows[order]).view(np.int64)), counts

    # -- routing / build
    def route_lines(self, keys, counts, n, first_line, ds, k1, world):
        self.k1 = k1
        keys = np.asarray(keys, dtype=np.uint64)[:n]
        gl = np.arange(n, dtype=np.uint64) + np.uint64(first_line)
        if ds:
            rc = np.array([self._kint(self.so.reverse_complement(self._kstr(int(k)))) for k in keys], dtype=np.uint64)
            keys = np.stack([keys, rc], axis=1).reshape(-1)
            gl = np.stack([2 * gl, 2 * gl + np.uint64(1)], axis=1).reshape(-1)
            counts = np.repeat(np.asarray(counts)[:n], 2)
        rows = np.stack([keys, (gl << np.uint64(30)) | np.asarray(counts, dtype=np.uint64)[:len(keys)]], axis=1)
        return self._group(minimizer_owner(keys, k1, world), rows, world)

    def build_from_records(self, recs, k1):
        self.k1 = k1
        r = self._u(recs).reshape(-1, 2)
        gl = r[:, 1] >> np.uint64(30)
        order = np.argsort(gl, kind="stable")
        self.table = {}                    # key -> [weight, local first idx], insertion = line order
        for i, j in enumerate(order.tolist()):
            key, w = int(r[j, 0]), int(r[j, 1] & np.uint64(0x3FFFFFFF))
            if self.so.low_complexity(self._kstr(key)):
                continue
            e = self.table.setdefault(key, [0, i])
            e[0] += w
        self.gline = gl[order]

    def n_distinct(self):
        return len(self.table)

    # -- components
    def _succ(self, key):
        mask = (1 << (2 * self.k1)) - 1
        return [((key << 2) & mask) | b for b in range(4)]

    def cc_local(self):
        parent = dict((k, k) for k in self.table)

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for k in self.table:
            for s in self._succ(k):
                if s in self.table:
                    a, b = find(k), find(s)
                    if a != b:
                        parent[max(a, b)] = min(a, b)
        roots = sorted(set(find(k) for k in self.table))
        rid = dict((r, i) for i, r in enumerate(roots))
        self.local_comp = dict((k, rid[find(k)]) for k in self.table)
        return len(roots)

    def cc_cross(self, world, rank, gid_base, k1):
        rows, dest = [], []
        for k in self.table:
            succ = np.array(self._succ(k), dtype=np.uint64)
            own = minimizer_owner(succ, k1, world)
            for s, o in zip(succ.tolist(), own.tolist()):
                if o != rank:
                    rows.append((s, gid_base + self.local_comp[k]))
                    dest.append(o)
        rows = np.array(rows, dtype=np.uint64).reshape(-1, 2)
        return self._group(np.array(dest, dtype=np.int64), rows, world)

    def cc_resolve(self, recs, gid_base):
        out = []
        for key, gid in self._u(recs).reshape(-1, 2).tolist():
            if key in self.table:
                out.append(gid | ((gid_base + self.local_comp[key]) << 32))
        return torch.from_numpy(np.array(out, dtype=np.uint64).view(np.int64))

    def cc_merge(self, edges, n_super):
        parent = list(range(n_super))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for e in self._u(edges).tolist():
            a, b = find(e & 0xFFFFFFFF), find(e >> 32)
            if a != b:
                parent[max(a, b)] = min(a, b)
        roots = sorted(set(find(i) for i in range(n_super)))
        rid = dict((r, i) for i, r in enumerate(roots))
        self.final_of_super = [rid[find(i)] for i in range(n_super)]
        return len(roots)

    def cc_sizes(self, gid_base, n_final):
        sizes = np.zeros(n_final, dtype=np.int64)
        for k, lc in self.local_comp.items():
            sizes[self.final_of_super[gid_base + lc]] += 1
        return torch.from_numpy(sizes)

    def cc_route(self, owner_of_final, gid_base, world, k1):
        own = owner_of_final.numpy()
        gl = self.gline
        rows, dest = [], []
        for k, (w, idx) in self.table.items():
            rows.append((k, (int(gl[idx]) << 30) | w))
            dest.append(int(own[self.final_of_super[gid_base + self.local_comp[k]]]))
        rows = np.array(rows, dtype=np.uint64).reshape(-1, 2)
        return self._group(np.array(dest, dtype=np.int64), rows, world)

    def relieve(self, min_free=0.35):
        pass

    def cc_free(self):
        self.local_comp = self.final_of_super = None

    # -- walks / candidates / filters
    def l3_walks(self, min_weight, min_length):
        self.min_weight, self.min_length = min_weight, min_length
        ordered = sorted(self.table.items(), key=lambda kv: kv[1][1])       # dict insertion order
        kmers = dict((self._kstr(k), w) for k, (w, _) in ordered)
        walks, _ = self.so.greedy_walks(kmers, min_weight)
        self.cands = [wk for wk in walks
                      if self.so.passes_shape(len(wk.contig), wk.tot_wt, wk.tot_kmer, min_weight, min_length)]
        self.kmers = kmers

    def cand_export(self):
        w = [self.kmers[c.seed] for c in self.cands]
        idx = [int(self.gline[self.table[self._kint(c.seed)][1]]) for c in self.cands]
        lens = [len(c.contig) for c in self.cands]
        offs = np.zeros(len(lens) + 1, dtype=np.int64)
        offs[1:] = np.cumsum(lens)
        codes = np.array(["AGCT".index(ch) for c in self.cands for ch in c.contig], dtype=np.uint8)
        return (torch.tensor(w, dtype=torch.int32), torch.tensor(idx, dtype=torch.int64),
                torch.from_numpy(offs), torch.from_numpy(codes))

    def l3_filter(self, codes, offs, n_cand):
        text = "".join("AGCT"[c] for c in codes.tolist())
        o = offs.tolist()
        walks = []
        for j in range(n_cand):
            wk = self.so.Walk()
            wk.contig = text[o[j]:o[j + 1]]
            wk.tot_wt, wk.tot_kmer = 10 ** 12, 1            # shape already decided by the owner rank
            walks.append(wk)
        self.contigs, self.connections, allowed = self.so.accept_walks(walks, self.k1, 1, 1)
        self.allowed = list(allowed)
        return {"n_allowed": len(self.allowed), "n_contigs": len(self.contigs) - 1}

    def allowed_weights(self, n_allowed):
        return torch.tensor([self.kmers.get(k, 0) for k in self.allowed], dtype=torch.int32)

    def set_allowed_weights(self, w):
        self.allowed_w = w.tolist()

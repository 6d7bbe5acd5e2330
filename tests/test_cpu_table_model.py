"""CPU model of the K1-mer table's probe protocol (shannon_b200/csrc/table_dev.cuh) for the one piece
of it that is an algorithmic claim rather than plumbing: with the home bucket a function of the
K-base PREFIX, `table_find_successors` finds the four successors x[1:].b of a K1-mer with ONE walk
over the buckets and stops at the first bucket that has a free slot or no overflow flag.  The model
inserts keys exactly like `table_insert_add` (first bucket of the probe sequence with a free slot;
every full bucket walked past gets the overflow flag) and checks the group lookup against one
independent lookup per key -- including long overflow chains, wrap-around at the end of the table,
and families split over several buckets."""
import random

import pytest

SLOTS = 4


class Table(object):
    def __init__(self, n_buckets, k1, rng):
        self.nb = n_buckets
        self.k1 = k1
        self.b = [[] for _ in range(n_buckets)]
        self.overflow = [False] * n_buckets
        self.salt = rng.getrandbits(61) | 1

    def home(self, key):                      # a function of key >> 2 only (ShnTableView::bucket_of)
        return ((key >> 2) * self.salt >> 7) % self.nb

    def insert(self, key):
        b = self.home(key)
        for _ in range(self.nb):
            if key in self.b[b]:
                return
            if len(self.b[b]) < SLOTS:
                self.b[b].append(key)
                return
            self.overflow[b] = True           # walked past a full bucket
            b = (b + 1) % self.nb
        raise AssertionError("table full")

    def find(self, key):                      # table_find_from
        b = self.home(key)
        while True:
            if key in self.b[b]:
                return (b, self.b[b].index(key))
            if len(self.b[b]) < SLOTS or not self.overflow[b]:
                return None
            b = (b + 1) % self.nb

    def find_successors(self, x):             # table_find_successors
        mask = (1 << (2 * self.k1)) - 1
        pre = (x << 2) & mask
        b = self.home(pre)
        found = {}
        loads = 0
        while True:
            loads += 1
            for j, k in enumerate(self.b[b]):
                if (k & ~3) == pre:
                    found[k & 3] = (b, j)
            if len(self.b[b]) < SLOTS or not self.overflow[b]:
                return found, loads
            b = (b + 1) % self.nb


@pytest.mark.parametrize("seed", range(8))
def test_one_probe_sequence_finds_all_four_successors(seed):
    rng = random.Random(seed)
    k1 = rng.choice([5, 7, 13])
    nb = rng.choice([8, 64, 257])
    load = rng.choice([0.5, 0.8, 0.97])       # high loads: long overflow chains and wrap-around
    t = Table(nb, k1, rng)
    space = 1 << (2 * k1)
    keys = set()
    # successor families: a random prefix with 1..4 extensions
    while len(keys) < int(load * nb * SLOTS):
        pre = rng.randrange(space >> 2) << 2
        for b in rng.sample(range(4), rng.choice([1, 1, 1, 2, 3, 4])):
            keys.add(pre | b)
    keys = list(keys)[:int(load * nb * SLOTS)]
    rng.shuffle(keys)
    for k in keys:
        t.insert(k)
    assert any(t.overflow), "case without overflow"
    present = set(keys)
    queries = [k >> 2 | (rng.randrange(4) << (2 * (k1 - 1))) for k in keys]     # predecessors of stored keys
    queries += [rng.randrange(space) for _ in range(200)]
    total_loads = 0
    for x in queries:
        got, loads = t.find_successors(x)
        total_loads += loads
        pre = (x << 2) & (space - 1)
        for b in range(4):
            assert got.get(b) == t.find(pre | b), "successor %d of %x" % (b, x)
            assert (b in got) == ((pre | b) in present)
    assert total_loads < 4 * len(queries) * (1 if load < 0.9 else 8)

"""A CPU model of the speculative walk protocol of ``walk_spec_kernel`` (shannon_b200/csrc/l3.cu,
DESIGN.md section 4) run under RANDOM schedules, against the sequential seed loop of the oracle
(extension_correction.py:343-350).

The GPU tests exercise the protocol with the schedules the hardware happens to produce; the cases
that decide its correctness -- a later seed walking over an earlier one's territory first, a left
extension meeting its own right extension on a cycle through the seed, claims examined one round
late -- depend on timing.  Here every unit of work (right / left extension of a window position) is
a generator that yields wherever the kernel waits for memory, and a random scheduler interleaves
the units of a window on a random number of "warps".  Whatever the schedule, the committed walks
must equal the oracle's walks, one for one and in pop order.

Modelled as in the kernel: stamps 2(W-p) / 2(W-p)-1 claimed with max(), candidates blocked by
committed-traversed or stamp >= own, two steps per round decided from ONE snapshot of the flags,
optimistic claims whose results are looked at one round later, blocker masks, the intact check,
COMMIT / SKIP / retry in pop order, roll back of stamps that are still one's own, the first position
of a window running both directions on one unit.
"""
import random

import pytest

import helpers  # noqa: F401  (puts the repository root on sys.path)
from oracle import shannon_oracle as so

BASES = "AGCT"  # the reference's successor order (extension_correction.py:10)


class Model(object):
    def __init__(self, kmers, order, window, rng, n_warps):
        self.kmers = kmers          # kmer -> weight
        self.order = order          # [(seed, weight)] in pop order
        self.W = window
        self.rng = rng
        self.n_warps = n_warps
        self.traversed = set()
        self.stamp = {}             # kmer -> claim stamp (absent = 0)
        self.walks = []             # committed (seed, n_left, n_right, tot_wt, contig)
        self.windows = self.retries = self.skips = 0

    # -- what one probe of the kernel sees -----------------------------------------------------
    def look(self, kmer):
        w = self.kmers.get(kmer)
        if w is None:
            return None
        return (w, kmer in self.traversed, self.stamp.get(kmer, 0))

    def claim(self, kmer, s):       # atomicMax; returns the stamp found
        old = self.stamp.get(kmer, 0)
        if s > old:
            self.stamp[kmer] = s
        return old

    def pos_of(self, st):
        return self.W - ((st + 1) >> 1)

    @staticmethod
    def child(cur, b, right):
        return cur[1:] + b if right else b + cur[:-1]

    # -- one unit of work: the directions `dirs` of window position p ------------------------------
    def unit(self, p, seed, dirs, out):
        stamp_r = 2 * (self.W - p)
        go = True
        if 0 in dirs:
            old = self.claim(seed, stamp_r)
            go = old < stamp_r
            out["len_r"] = 1 if go else 0
            out["path_r"] = [seed] if go else []
        else:
            go = self.stamp.get(seed, 0) <= stamp_r
        yield
        if not go:
            return
        for d in dirs:
            right = d == 0
            s = stamp_r - d
            path = out["path_r"] if right else out["path_l"]
            bases = out["bases_r"] if right else out["bases_l"]
            tot = self.kmers[seed] if right else 0
            cur = seed
            pend = []
            poisoned = False
            while True:
                # ONE snapshot for both steps of the round (lanes 0..3 and 4..19 load together)
                lvl1 = [(b, self.child(cur, b, right)) for b in BASES]
                snap1 = {c: self.look(c) for _, c in lvl1}
                snap2 = {}
                for _, c in lvl1:
                    if snap1[c] is not None:
                        for b2 in BASES:
                            g = self.child(c, b2, right)
                            snap2[(c, g)] = self.look(g)
                yield                                   # the loads are in flight
                if any(o >= s for o in pend):           # claims of the previous round
                    poisoned = True
                    break
                pend = []
                best = None
                for b, c in lvl1:
                    v = snap1[c]
                    if v is None:
                        continue
                    w, trav, st = v
                    if not trav and st > stamp_r and self.pos_of(st) != p:
                        out["block"].add(self.pos_of(st))
                    if trav or st >= s:
                        continue
                    if best is None or w > best[0]:
                        best = (w, b, c)
                if best is None:
                    break
                w1, b1, c1 = best
                pend.append(self.claim(c1, s))
                path.append(c1)
                bases.append(b1)
                tot += w1
                if self.rng.random() < 0.5:
                    yield
                best = None
                for b2 in BASES:
                    g = self.child(c1, b2, right)
                    v = snap2[(c1, g)]
                    if v is None:
                        continue
                    w, trav, st = v
                    if not trav and st > stamp_r and self.pos_of(st) != p:
                        out["block"].add(self.pos_of(st))
                    if trav or st >= s or g == c1:
                        continue
                    if best is None or w > best[0]:
                        best = (w, b2, g)
                if best is None:
                    break
                w2, b2, c2 = best
                pend.append(self.claim(c2, s))
                path.append(c2)
                bases.append(b2)
                tot += w2
                cur = c2
            if any(o >= s for o in pend):
                poisoned = True
            if right:
                out["tot_r"] = tot
            else:
                out["tot_l"] = tot
            if poisoned:
                out["poison"] = True
                return

    # -- one window ----------------------------------------------------------------------------------
    def run(self):
        cursor = 0
        n = len(self.order)
        while True:
            win = []
            pos = cursor
            while len(win) < self.W and pos < n:
                if self.order[pos][0] not in self.traversed:
                    win.append(pos)
                pos += 1
            cursor_after = pos
            if not win:
                break
            self.windows += 1
            outs = [dict(len_r=0, path_r=[], path_l=[], bases_r=[], bases_l=[], tot_r=0, tot_l=0,
                         block=set(), poison=False) for _ in win]
            units = [(0, (0, 1))]
            for p in range(1, len(win)):
                units += [(p, (0,)), (p, (1,))]
            todo = list(units)
            running = []
            # `n_warps` workers pull units in order; a random worker advances at every tick
            while todo or running:
                while todo and len(running) < self.n_warps:
                    p, dirs = todo.pop(0)
                    running.append(self.unit(p, self.order[win[p]][0], dirs, outs[p]))
                g = self.rng.choice(running)
                try:
                    next(g)
                except StopIteration:
                    running.remove(g)
            # intact? thief?
            intact, thief = [], []
            for p, o in enumerate(outs):
                stamp_r = 2 * (self.W - p)
                ok = bool(o["path_r"]) and not o["poison"]
                ok = ok and all(self.stamp.get(k, 0) == stamp_r for k in o["path_r"])
                ok = ok and all(self.stamp.get(k, 0) == stamp_r - 1 for k in o["path_l"])
                intact.append(ok)
                sst = self.stamp.get(self.order[win[p]][0], 0)
                thief.append(self.pos_of(sst) if sst > stamp_r else None)
            committed = set()
            P = 0
            status = {}
            for P in range(len(win) + 1):
                if P == len(win):
                    break
                if intact[P] and outs[P]["block"] <= committed:
                    status[P] = 1
                    committed.add(P)
                elif thief[P] is not None and thief[P] in committed:
                    status[P] = 2
                    self.skips += 1
                else:
                    break
            assert P >= 1, "the first position of a window must always commit"
            self.retries += len(win) - P
            for p, o in enumerate(outs):
                stamp_r = 2 * (self.W - p)
                if status.get(p) == 1:
                    seed = self.order[win[p]][0]
                    for k in o["path_r"] + o["path_l"]:
                        assert k not in self.traversed
                        self.traversed.add(k)
                    contig = "".join(reversed(o["bases_l"])) + seed + "".join(o["bases_r"])
                    self.walks.append((seed, len(o["bases_l"]), len(o["bases_r"]), o["tot_r"] + o["tot_l"], contig))
                else:
                    for k in o["path_r"]:
                        if self.stamp.get(k, 0) == stamp_r:
                            self.stamp[k] = 0
                    for k in o["path_l"]:
                        if self.stamp.get(k, 0) == stamp_r - 1:
                            self.stamp[k] = 0
            # after a window every untraversed K1-mer is unclaimed again
            assert all(st == 0 or k in self.traversed for k, st in self.stamp.items())
            cursor = win[P] if P < len(win) else cursor_after
        return self.walks


def _random_kmers(rng, k1, genome_len, n_reads, read_len, err, repeats):
    """K1-mer counts of error-laden reads from a small genome with repeated segments (cycles and
    branches in the K1-mer graph)."""
    genome = "".join(rng.choice("ACGT") for _ in range(genome_len))
    for _ in range(repeats):
        a = rng.randrange(0, genome_len - 12)
        seg = genome[a:a + rng.randrange(k1, 2 * k1)]
        b = rng.randrange(0, len(genome))
        genome = genome[:b] + seg + genome[b:]
    kmers = {}
    for _ in range(n_reads):
        a = rng.randrange(0, len(genome) - read_len)
        read = list(genome[a:a + read_len])
        for i in range(read_len):
            if rng.random() < err:
                read[i] = rng.choice("ACGT")
        read = "".join(read)
        for i in range(read_len - k1 + 1):
            km = read[i:i + k1]
            kmers[km] = kmers.get(km, 0) + 1
    return kmers


@pytest.mark.parametrize("seed", range(12))
def test_speculative_half_walks_equal_the_sequential_loop_under_random_schedules(seed):
    orc = so
    rng = random.Random(1000 + seed)
    k1 = rng.choice([5, 6, 7, 9])
    kmers = _random_kmers(rng, k1, genome_len=rng.choice([60, 150, 400]), n_reads=rng.choice([40, 120]),
                          read_len=rng.choice([20, 30]), err=rng.choice([0.0, 0.02, 0.08]),
                          repeats=rng.choice([0, 3, 8]))
    ref, traversed = orc.greedy_walks(kmers, 1)
    expect = [(w.seed, w.n_left, w.n_right, w.tot_wt, w.contig) for w in ref]
    order = orc.seed_order(kmers)
    for trial in range(6):
        m = Model(kmers, order, window=rng.choice([2, 4, 8, 32]), rng=random.Random(seed * 100 + trial),
                  n_warps=rng.choice([1, 2, 3, 8, 64]))
        got = m.run()
        assert got == expect, "schedule %d of case %d differs from the sequential loop" % (trial, seed)
        assert m.traversed == traversed


def test_model_exercises_skips_retries_and_cycles():
    """The random cases above are only meaningful if the interesting events happen."""
    orc = so
    skips = retries = windows = 0
    for seed in range(12):
        rng = random.Random(1000 + seed)
        k1 = rng.choice([5, 6, 7, 9])
        kmers = _random_kmers(rng, k1, genome_len=rng.choice([60, 150, 400]), n_reads=rng.choice([40, 120]),
                              read_len=rng.choice([20, 30]), err=rng.choice([0.0, 0.02, 0.08]),
                              repeats=rng.choice([0, 3, 8]))
        m = Model(kmers, orc.seed_order(kmers), window=32, rng=random.Random(seed), n_warps=8)
        m.run()
        skips += m.skips
        retries += m.retries
        windows += m.windows
    assert skips > 0 and retries > 0 and windows > 12

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

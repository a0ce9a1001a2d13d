"""Discrete-event model of the synchronisation protocol of conv_tc3_fused_kernel (fastvocoder_b200/csrc/fv_tc.cuh).

TEST INFRASTRUCTURE (no GPU): the persistent kernel's roles — loader warps, weight-ring producer, UMMA issuers,
epiA, epiB — are re-stated as coroutines over a model of mbarriers (count + completed-phase counter), an in-order
tensor pipe (tcgen05.mma / tcgen05.commit) and asynchronous bulk copies.  A random scheduler interleaves them and checks

* no deadlock: every role finishes;
* mbarrier phase discipline: a waiter is never more than one phase ahead of (or behind) the barrier it polls —
  `try_wait.parity` cannot tell completion k from completion k +- 2, so that would be a silent mis-synchronisation;
* data hazards: every UMMA reads the A tile / ring slot content it expects (tile id, conv, tap), every epilogue
  reads the accumulator set of its own tile, and no buffer, ring slot or TMEM set is overwritten before its readers
  are done.

The four plan families tc3_plan produces are covered: resident weights (double-buffered A1 / acc1 / acc2), streamed
weights with single-buffered tall tiles (conv2 issued before conv1 of the next tile), and ping-pong tiles (one
in-place A buffer + one accumulator set per tile in flight) with streamed or resident weights.
"""
from __future__ import annotations

import random


class ProtocolError(AssertionError):
    pass


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.done = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending < 0:
            raise ProtocolError(f"{self.name}: more arrivals than the barrier's count in one phase")
        if self.pending == 0:
            self.done += 1
            self.pending = self.count

    def ready(self, k, who):
        """Completion number k (0-based) awaited with parity k & 1: legal only while done is k or k + 1."""
        if not (k <= self.done <= k + 1):
            raise ProtocolError(f"{who}: waits for completion {k} of {self.name} but {self.done} completed "
                                "(parity wait would alias)")
        return self.done == k + 1


class Sim:
    def __init__(self, *, n_tiles, K, n_issuers, a1_stages, acc1_stages, acc2_stages, resident, w_stages, pp, seed):
        self.n, self.K, self.ni = n_tiles, K, n_issuers
        self.a1, self.acc1, self.acc2 = a1_stages, acc1_stages, acc2_stages
        self.resident, self.wst, self.pp = resident, w_stages, pp
        self.rng = random.Random(seed)
        self.conv2_first = a1_stages == 1
        B = lambda name, c: Bar(name, c)
        # loaders / epilogue warp groups are modelled as one actor each (their arrivals are symmetric) -> count 1
        self.a1_full = [B(f"a1_full{s}", 1) for s in range(2)]
        self.a1_empty = [B(f"a1_empty{s}", n_issuers) for s in range(2)]
        self.acc1_full = [B(f"acc1_full{s}", n_issuers) for s in range(2)]
        self.acc1_empty = [B(f"acc1_empty{s}", 1) for s in range(2)]
        self.a2_full = [B("a2_full0", 1), B("a2_full1", 1)]      # pp: one per buffer; else only [0]
        self.a2_empty = B("a2_empty", n_issuers)
        self.acc2_full = [B(f"acc2_full{s}", n_issuers) for s in range(2)]
        self.acc2_empty = [B(f"acc2_empty{s}", 1) for s in range(2)]
        self.w_full = [B(f"w_full{s}", 1) for s in range(4)]
        self.w_empty = [B(f"w_empty{s}", n_issuers) for s in range(4)]
        # data
        self.A1 = [None, None]            # label of each A1 stage: ("x", tile) or ("h", tile) (pp, in place)
        self.A1_readers = [n_issuers, n_issuers]   # issuers done with the current content (starts "free")
        self.A2 = None
        self.A2_readers = n_issuers
        n_sets = 2 if pp else acc1_stages + acc2_stages
        self.tmem = [None] * n_sets       # ("c1"|"c2", tile, issuers_written)
        self.tmem_drained = [True] * n_sets
        self.slots = [None] * 4           # (kind, tap, g)
        self.slot_readers = [n_issuers] * 4
        self.pipe = []                    # in-order tensor pipe: ("mma", fn) / ("commit", bar)
        self.copies = []                  # in-flight bulk copies: (slot, label)
        self.stored = []                  # tiles written to global by epiB, in order

    # ---- consumption order of the convs (shared by producer and issuers, fv_tc.cuh conv1/conv2 loops)
    def order(self):
        n = self.n
        seq = []
        if self.pp:
            for i in range(0, n, 2):
                seq.append(("c1", i))
                if i + 1 < n: seq.append(("c1", i + 1))
                seq.append(("c2", i))
                if i + 1 < n: seq.append(("c2", i + 1))
            return seq
        if n > 0: seq.append(("c1", 0))
        for i in range(n):
            if self.conv2_first:
                seq.append(("c2", i))
                if i + 1 < n: seq.append(("c1", i + 1))
            else:
                if i + 1 < n: seq.append(("c1", i + 1))
                seq.append(("c2", i))
        return seq

    # ---- roles -----------------------------------------------------------------------------------------------
    def loader(self):
        for it in range(self.n):
            s = it % self.a1
            if it >= self.a1:
                k = it // self.a1 - 1
                while not self.a1_empty[s].ready(k, "loader"): yield
            if self.A1_readers[s] != self.ni:
                raise ProtocolError(f"loader overwrites A1[{s}] ({self.A1[s]}) before its readers are done")
            self.A1[s], self.A1_readers[s] = ("x", it), 0
            self.a1_full[s].arrive()
            yield

    def producer(self):
        g = 0
        for kind, _ in self.order():
            for j in range(self.K):
                slot = g % self.wst
                if g >= self.wst:
                    k = g // self.wst - 1
                    while not self.w_empty[slot].ready(k, "producer"): yield
                if self.slot_readers[slot] != self.ni:
                    raise ProtocolError(f"producer refills ring slot {slot} before its readers are done")
                self.slot_readers[slot] = 0
                self.copies.append((slot, ("w1" if kind == "c1" else "w2", j, g)))
                g += 1
                yield

    def issuer(self, w):
        gw = 0
        for kind, i in self.order():
            who = f"issuer{w} {kind}({i})"
            if kind == "c1":
                s, a = i % self.a1, i % self.acc1
                while not self.a1_full[s].ready(i // self.a1, who): yield
                if self.pp:
                    if i >= 2:
                        while not self.acc2_empty[a].ready(i // 2 - 1, who): yield
                elif i >= self.acc1:
                    while not self.acc1_empty[a].ready(i // self.acc1 - 1, who): yield
                src, want_a, dst = ("A1", s), ("x", i), a
                commits = ([] if self.pp else [self.a1_empty[s]]) + [self.acc1_full[a]]
            else:
                b = i % self.acc2
                if self.pp:
                    while not self.a2_full[b].ready(i // 2, who): yield
                    src, want_a, dst = ("A1", b), ("h", i), b
                    commits = [self.a1_empty[b], self.acc2_full[b]]
                else:
                    while not self.a2_full[0].ready(i, who): yield
                    if i >= self.acc2:
                        while not self.acc2_empty[b].ready(i // self.acc2 - 1, who): yield
                    src, want_a, dst = ("A2", 0), ("h", i), self.acc1 + b
                    commits = [self.a2_empty, self.acc2_full[b]]
            for j in range(self.K):
                slot = None
                if not self.resident:
                    slot = gw % self.wst
                    while not self.w_full[slot].ready(gw // self.wst, who): yield
                self.pipe.append(("mma", self._mma(w, kind, i, j, src, want_a, dst, slot, gw, j == self.K - 1)))
                if slot is not None:
                    self.pipe.append(("release", slot))
                    self.pipe.append(("commit", self.w_empty[slot]))
                gw += 1
                yield
            # bookkeeping for the hazard checks (this issuer's UMMAs of the conv have read their A operand), then the commits
            if kind == "c1":
                self.pipe.append(("readers", ("A1", i % self.a1)) if not self.pp else ("nop", None))
            else:
                self.pipe.append(("readers", ("A1", i % 2) if self.pp else ("A2", 0)))
            for bar in commits:
                self.pipe.append(("commit", bar))
            yield

    def _mma(self, w, kind, i, j, src, want_a, dst, slot, g, last):
        def run():
            have = self.A1[src[1]] if src[0] == "A1" else self.A2
            if have != want_a:
                raise ProtocolError(f"issuer{w} {kind}({i}) tap {j}: A operand holds {have}, expected {want_a}")
            if slot is not None:
                want_w = ("w1" if kind == "c1" else "w2", j, g)
                if self.slots[slot] != want_w:
                    raise ProtocolError(f"issuer{w} {kind}({i}): ring slot {slot} holds {self.slots[slot]}, expected {want_w}")
            cur = self.tmem[dst]
            if j == 0 and (cur is None or cur[:2] != (kind, i)):
                # first issuer to touch the set for this conv: the previous content must have been drained
                if not self.tmem_drained[dst]:
                    raise ProtocolError(f"issuer{w} {kind}({i}) overwrites TMEM set {dst} holding {cur} before it was drained")
                self.tmem[dst], self.tmem_drained[dst] = (kind, i, 0), False
            if last:
                k_, i_, n_ = self.tmem[dst]
                self.tmem[dst] = (k_, i_, n_ + 1)
        return run

    def epiA(self):
        for it in range(self.n):
            a = it % self.acc1
            while not self.acc1_full[a].ready(it // self.acc1, "epiA"): yield
            if not self.pp and it >= 1:
                while not self.a2_empty.ready(it - 1, "epiA"): yield
            if self.tmem[a] != ("c1", it, self.ni):
                raise ProtocolError(f"epiA({it}) reads TMEM set {a} holding {self.tmem[a]}")
            yield
            if self.pp:
                if self.A1[a] != ("x", it):
                    raise ProtocolError(f"epiA({it}) writes h over A1[{a}] holding {self.A1[a]}")
                self.A1[a] = ("h", it)
                self.tmem_drained[a] = True
                self.a2_full[a].arrive()
            else:
                if self.A2_readers != self.ni:
                    raise ProtocolError(f"epiA({it}) overwrites A2 ({self.A2}) before conv2 read it")
                self.A2, self.A2_readers = ("h", it), 0
                self.tmem_drained[a] = True
                self.acc1_empty[a].arrive()
                self.a2_full[0].arrive()
            yield

    def epiB(self):
        for it in range(self.n):
            b = it % self.acc2
            st = b if self.pp else self.acc1 + b
            while not self.acc2_full[b].ready(it // self.acc2, "epiB"): yield
            if self.tmem[st] != ("c2", it, self.ni):
                raise ProtocolError(f"epiB({it}) reads TMEM set {st} holding {self.tmem[st]}")
            yield
            self.stored.append(it)
            self.tmem_drained[st] = True
            self.acc2_empty[b].arrive()
            yield

    # ---- scheduler ---------------------------------------------------------------------------------------------
    def run(self, max_steps=2_000_000):
        roles = {"loader": self.loader(), "epiA": self.epiA(), "epiB": self.epiB()}
        if not self.resident:
            roles["producer"] = self.producer()
        for w in range(self.ni):
            roles[f"issuer{w}"] = self.issuer(w)
        alive = dict(roles)
        stuck = 0
        for _ in range(max_steps):
            choices = list(alive) + (["pipe"] if self.pipe else []) + (["copy"] if self.copies else [])
            if not choices:
                break
            c = self.rng.choice(choices)
            before = self._state()
            if c == "pipe":
                op, arg = self.pipe.pop(0)
                if op == "mma": arg()
                elif op == "commit": arg.arrive()
                elif op == "release": self.slot_readers[arg] += 1
                elif op == "readers":
                    if arg[0] == "A1": self.A1_readers[arg[1]] += 1
                    else: self.A2_readers += 1
            elif c == "copy":
                slot, label = self.copies.pop(self.rng.randrange(len(self.copies)))
                self.slots[slot] = label
                self.w_full[slot].arrive()
            else:
                try:
                    next(alive[c])
                except StopIteration:
                    del alive[c]
            stuck = stuck + 1 if self._state() == before and c not in ("pipe", "copy") else 0
            if stuck > 20000:
                raise ProtocolError(f"deadlock: roles {sorted(alive)} make no progress; stored tiles {self.stored}")
        if alive or self.pipe or self.copies:
            raise ProtocolError(f"did not finish: {sorted(alive)}")
        if self.stored != list(range(self.n)):
            raise ProtocolError(f"tiles stored {self.stored}")
        return True

    def _state(self):
        bars = self.a1_full + self.a1_empty + self.acc1_full + self.acc1_empty + self.a2_full + [self.a2_empty] + \
            self.acc2_full + self.acc2_empty + self.w_full + self.w_empty
        return (tuple((b.done, b.pending) for b in bars), len(self.pipe), len(self.copies), tuple(self.stored),
                tuple(self.A1), self.A2)


PLANS = {
    # name: kwargs mirroring what tc3_plan emits
    "resident double-buffered (C<=32, C=64 k=3)": dict(a1_stages=2, acc1_stages=2, acc2_stages=2, resident=True, w_stages=0, pp=False),
    "resident, single A1 stage": dict(a1_stages=1, acc1_stages=2, acc2_stages=2, resident=True, w_stages=0, pp=False),
    "resident, single acc1": dict(a1_stages=2, acc1_stages=1, acc2_stages=2, resident=True, w_stages=0, pp=False),
    "streamed weights, tall single-buffered tile": dict(a1_stages=1, acc1_stages=1, acc2_stages=1, resident=False, w_stages=4, pp=False),
    "streamed weights, m=1 double-buffered": dict(a1_stages=2, acc1_stages=2, acc2_stages=2, resident=False, w_stages=4, pp=False),
    "streamed weights, 3-slot ring": dict(a1_stages=1, acc1_stages=1, acc2_stages=1, resident=False, w_stages=3, pp=False),
    "ping-pong tiles, streamed weights": dict(a1_stages=2, acc1_stages=2, acc2_stages=2, resident=False, w_stages=4, pp=True),
    "ping-pong tiles, resident weights": dict(a1_stages=2, acc1_stages=2, acc2_stages=2, resident=True, w_stages=0, pp=True),
}

"""A small executable model of the peer-memory protocol of csrc/peer.cuh / maxk_b200/peer.py.

The multi-GPU path has no "done reading" handshake: it relies on two sets of buffers alternating
by step parity plus the ordering that the arrival / publish flags already impose.  This model
makes that argument checkable on a CPU: every rank is a little program of stream-ordered
operations, an adversarial scheduler interleaves the ranks in every possible (small cases) or in
many random (larger cases) ways, buffers carry (writer, epoch) tags, and the checker fails if

  * a rank ever aggregates from a block that does not hold exactly the epoch it waits for, or
  * a block is written while another rank's kernel is in the middle of reading it.

Operations per rank and step e (same order as PeerAggregator.exchange/forward/backward):

  F1  TopK writes the own block of the own gathered buffer g[e & 1]
  F2  push: wait until arrive[rank][p] >= e - 1 for all p (every peer has entered step e - 1 and is done with the
      buffers of step e - 2), write the own block into every peer's g[e & 1], then raise arrive[peer][rank] = e
  F3  wait until arrive[rank][p] >= e for all p; READ every block of the own g[e & 1]   (begin .. end)
      -- a rank WITHOUT tasks on remote columns (directed / asymmetric partitions) never waits here and reads its own
      block only: the wait in F2 is what keeps such a rank from running two steps ahead (`no_wait` ranks below)
  B1  zero + scatter into the own partial buffer part[b & 1]                             (all blocks)
  B2  publish bwd[rank] = b
  B3  wait until bwd[p] >= b for all p; READ block `rank` of every peer's part[b & 1]    (begin .. end)

Forward and backward steps carry their own counters, so the layered training pattern
(F F F B B B) is covered as well as F B F B and forward-only streams.  The model also shows that the
SAME protocol with a single set of buffers is broken (the checker must catch it)."""
import itertools
import random

import pytest


class Violation(Exception):
    pass


class Model:
    def __init__(self, world, programs, buffers=2, no_wait=(), gate_push=True):
        self.w = world
        self.nbuf = buffers
        self.no_wait = set(no_wait)     # ranks whose forward has no task on a remote column
        self.gate_push = gate_push      # F2 waits for the peers' arrivals of the step before
        # g[r][parity][block] = epoch of the data that block holds; part[r][parity] = epoch
        self.g = [[[0] * world for _ in range(buffers)] for _ in range(world)]
        self.part = [[0] * buffers for _ in range(world)]
        self.arrive = [[0] * world for _ in range(world)]   # arrive[dst][src]
        self.bwd = [0] * world
        self.reading_g = [None] * world      # rank -> parity being read by its forward kernel
        self.reading_part = [None] * world   # rank -> parity its reduce-scatter is pulling
        self.prog = programs                  # per rank: list of (op, epoch)
        self.pc = [0] * world

    def enabled(self, r):
        if self.pc[r] >= len(self.prog[r]):
            return False
        op, e = self.prog[r][self.pc[r]]
        if op == "F2" and self.gate_push:
            return all(self.arrive[r][p] >= e - 1 for p in range(self.w) if p != r)
        if op == "F3b":
            return r in self.no_wait or all(self.arrive[r][p] >= e for p in range(self.w) if p != r)
        if op == "B3b":
            return all(self.bwd[p] >= e for p in range(self.w) if p != r)
        return True

    def done(self):
        return all(self.pc[r] >= len(self.prog[r]) for r in range(self.w))

    def step(self, r):
        op, e = self.prog[r][self.pc[r]]
        par = e % self.nbuf
        if op == "F1":
            if self.reading_g[r] == par:
                raise Violation(f"rank {r} overwrites its own block while its kernel reads g[{par}]")
            self.g[r][par][r] = e
        elif op == "F2":
            for p in range(self.w):
                if p == r:
                    continue
                if self.reading_g[p] == par and p not in self.no_wait:   # (a no-wait rank reads its own block only)
                    raise Violation(f"rank {r} pushes epoch {e} into g[{par}] of rank {p} which is being read")
                self.g[p][par][r] = e
                self.arrive[p][r] = e
        elif op == "F3b":
            for blk in ([r] if r in self.no_wait else range(self.w)):
                if self.g[r][par][blk] != e:
                    raise Violation(f"rank {r} step {e}: block {blk} holds epoch {self.g[r][par][blk]}")
            self.reading_g[r] = par
        elif op == "F3e":
            for blk in ([r] if r in self.no_wait else range(self.w)):   # still intact at the end of the kernel
                if self.g[r][par][blk] != e:
                    raise Violation(f"rank {r} step {e}: block {blk} changed under the kernel")
            self.reading_g[r] = None
        elif op == "B1":
            for p in range(self.w):
                if p != r and self.reading_part[p] == par:
                    raise Violation(f"rank {r} rewrites part[{par}] while rank {p} pulls from it")
            self.part[r][par] = e
        elif op == "B2":
            self.bwd[r] = e
        elif op == "B3b":
            for p in range(self.w):
                if self.part[p][par] != e:
                    raise Violation(f"rank {r} bwd {e}: partial of rank {p} holds epoch {self.part[p][par]}")
            self.reading_part[r] = par
        elif op == "B3e":
            for p in range(self.w):
                if self.part[p][par] != e:
                    raise Violation(f"rank {r} bwd {e}: partial of rank {p} changed under the pull")
            self.reading_part[r] = None
        self.pc[r] += 1


def program(pattern):
    """pattern: string of F / B steps, e.g. "FBFB", "FFFBBB", "FFFF"."""
    ops, f, b = [], 0, 0
    for ch in pattern:
        if ch == "F":
            f += 1
            ops += [("F1", f), ("F2", f), ("F3b", f), ("F3e", f)]
        else:
            b += 1
            ops += [("B1", b), ("B2", b), ("B3b", b), ("B3e", b)]
    return ops


def run_random(world, pattern, seed, buffers=2, no_wait=(), gate_push=True):
    rng = random.Random(seed)
    m = Model(world, [program(pattern) for _ in range(world)], buffers, no_wait, gate_push)
    # an adversarial flavour: one rank is favoured so that it runs as far ahead as it can
    fav = rng.randrange(world)
    while not m.done():
        ready = [r for r in range(world) if m.enabled(r)]
        assert ready, "deadlock: nobody can move"
        if fav in ready and rng.random() < 0.7:
            r = fav
        else:
            r = rng.choice(ready)
        m.step(r)


def explore_all(world, pattern, buffers=2, no_wait=(), gate_push=True):
    """Depth-first search over every interleaving (with state memoisation)."""
    progs = [program(pattern) for _ in range(world)]
    seen = set()
    stack = [Model(world, progs, buffers, no_wait, gate_push)]
    count = 0
    import copy
    while stack:
        m = stack.pop()
        key = (tuple(m.pc), tuple(map(tuple, (tuple(x) for x in itertools.chain.from_iterable(m.g)))),
               tuple(map(tuple, m.part)), tuple(map(tuple, m.arrive)), tuple(m.bwd),
               tuple(m.reading_g), tuple(m.reading_part))
        if key in seen:
            continue
        seen.add(key)
        if m.done():
            count += 1
            continue
        ready = [r for r in range(world) if m.enabled(r)]
        assert ready, "deadlock: nobody can move"
        for r in ready:
            n = copy.deepcopy(m)
            n.step(r)
            stack.append(n)
    return len(seen)


@pytest.mark.parametrize("pattern", ["FBFBFB", "FFFBBB", "FFFFF", "FFBBFFBB"])
def test_two_buffers_are_enough_exhaustive_two_ranks(pattern):
    assert explore_all(2, pattern) > 10


def test_two_buffers_are_enough_exhaustive_three_ranks():
    assert explore_all(3, "FBFB") > 100
    assert explore_all(3, "FFF") > 100


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("pattern", ["FBFBFBFB", "FFFBBBFFFBBB", "FFFFFFFF"])
def test_two_buffers_are_enough_random_schedules(world, pattern):
    for seed in range(300):
        run_random(world, pattern, seed)


def test_the_checker_catches_a_single_buffer():
    """Same protocol, one set of buffers: with two forward (or two backward) steps in a row -- the
    layers of one training step -- a fast rank overwrites what a slow one still reads."""
    with pytest.raises(Violation):
        for seed in range(200):
            run_random(2, "FFFF", seed, buffers=1)
    with pytest.raises(Violation):
        explore_all(2, "FFBB", buffers=1)
    with pytest.raises(Violation):
        explore_all(3, "BB", buffers=1)


def test_strict_alternation_would_not_even_need_the_second_buffer():
    """What the model also shows: when forward and backward steps strictly alternate (one layer),
    the flags of each direction order the buffers of the other and one set would do.  The second
    set is what makes F F F B B B (several layers per step) and forward-only streams safe."""
    assert explore_all(2, "FBFBFB", buffers=1) > 10
    assert explore_all(3, "FBFB", buffers=1) > 100


def test_a_rank_without_remote_tasks_cannot_run_ahead():
    """A rank whose own tasks never read a remote column does not wait in its aggregating warps.  With the push
    gated on the peers' arrivals of the step before (peer.cuh: peer_push_block) forward-only streams stay safe in
    every interleaving; WITHOUT that gate -- round 1's protocol -- such a rank runs two steps ahead and overwrites a
    block a slower peer is still reading (the checker must catch it)."""
    for no_wait in ((0,), (1,), (0, 1)):
        assert explore_all(2, "FFFFF", no_wait=no_wait) > 10
    assert explore_all(3, "FFFF", no_wait=(2,)) > 100
    assert explore_all(2, "FFFBBB", no_wait=(0,)) > 10
    for seed in range(200):
        run_random(8, "FFFFFFFF", seed, no_wait=(seed % 8, (seed + 3) % 8))
    with pytest.raises(Violation):
        explore_all(2, "FFFF", no_wait=(0,), gate_push=False)

#!/usr/bin/env python
"""CPU model check of the mbarrier protocol of `umma_gemm_nt_v2_kernel` (csrc/agp_umma.cu).

No GPU involved: the roles of one CTA (TMA producer, MMA issuer, 4 epilogue warps, NCG x 4 converter warps) are
coroutines that execute the kernel's wait / arrive / issue sequence; TMA landings and tensor-core completions are
asynchronous events.  A seeded random scheduler interleaves everything, and every read checks a content tag:

  * a converter warp reading raw-A stage `sa` for k-block g must find k-block g's tile there,
  * the tensor core executing the MMAs of k-block g must find A_hi/A_lo of g in TMEM slot g % TS (written by all
    four warps of the converting group) and B of g in stage g % RB (converted when not pre-split),
  * a TMA landing or a TMEM store must never overwrite data whose consumer has not finished,
  * an epilogue warp must find exactly the unit's k-blocks in the accumulator it drains, and the MMA must never
    write an accumulator that is still being drained,
  * mbarrier parity waits use the hardware rule (`try_wait.parity(p)` succeeds iff the barrier's current phase bit
    differs from p), so a waiter that could lag two phases behind would be caught as a wrong-tag read,
  * the run must end with every role finished (no deadlock).

Usage:  python tools/umma_v2_protocol_sim.py [--trials 300] [--seed 0]
Also imported by tests/test_umma_v2_protocol.py (a few hundred random schedules, CPU only).
"""
from __future__ import annotations

import argparse
import random


class Bar:
    def __init__(self, count):
        self.count0 = count
        self.pending = count
        self.tx = 0
        self.phase = 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count0

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier's count in one phase"
        self.pending -= 1
        self._maybe_flip()

    def expect_tx(self, nbytes):      # mbarrier.arrive.expect_tx
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        assert self.tx >= 0
        self._maybe_flip()

    def test(self, parity):           # try_wait.parity
        return self.phase != parity


class ProtocolError(AssertionError):
    pass


def simulate(units, RA=4, RB=4, TS=4, NCG=2, presplit=True, seed=0, max_steps=4_000_000, bug=None, slow=()):
    """units: list of nkb (k-blocks per work unit of this CTA). Returns the number of scheduler steps.
    bug: drop one wait of the protocol (mutation test of this checker): 'slot' (converter does not wait for the TMEM
    slot), 'a_empty' (producer does not wait for the raw-A stage), 'b_stage' (producer does not wait for the MMA before
    refilling B), 'tmem_empty' (MMA does not wait for the epilogue).
    slow: role indices (0 producer, 1 MMA, 2..5 epilogue warps, 6.. converter warps) scheduled ~50x less often than the rest,
    'tma' / 'tensor' to starve the asynchronous engines instead."""
    rng = random.Random(seed)
    assert RB == TS
    a_full = [Bar(1) for _ in range(RA)]
    a_empty = [Bar(4) for _ in range(RA)]          # 4 warps (128 threads) of the converting group
    b_full = [Bar(1) for _ in range(RB)]
    conv_done = [Bar(4) for _ in range(TS)]
    mma_done = [Bar(1) for _ in range(TS)]
    tmem_full = [Bar(1) for _ in range(2)]
    tmem_empty = [Bar(4) for _ in range(2)]

    TILE = 16384
    a_stage = [None] * RA                            # content tag: k-block index g
    a_readers_left = [0] * RA                        # converter warps that still have to read the stage
    b_stage = [dict(tag=None, parts=0, converted=0, busy=False) for _ in range(RB)]
    tmem_a = [dict(tag=None, quarters=set(), busy=False) for _ in range(TS)]
    acc = [dict(blocks=[], draining=0, unit=None) for _ in range(2)]
    async_events = []                                # callables fired at random later times
    tensor_queue = []                                # in-order tensor-core work

    def tma(bar, nbytes, fn):
        def land():
            fn()
            bar.complete_tx(nbytes)
        async_events.append(land)

    # ---- roles ----
    def producer():
        g = 0
        for nkb in units:
            for _ in range(nkb):
                sa, sb = g % RA, g % RB
                while bug != 'a_empty' and not a_empty[sa].test(((g // RA) & 1) ^ 1):
                    yield
                a_full[sa].expect_tx(TILE)

                def land_a(sa=sa, g=g):
                    if a_readers_left[sa] != 0:
                        raise ProtocolError(f"TMA overwrote raw A stage {sa} before k-block {a_stage[sa]} was read (g={g})")
                    a_stage[sa] = g
                    a_readers_left[sa] = 4
                tma(a_full[sa], TILE, land_a)
                while bug != 'b_stage' and not mma_done[sb].test(((g // RB) & 1) ^ 1):
                    yield
                nparts = 2 if presplit else 1
                b_full[sb].expect_tx(nparts * TILE)
                st = b_stage[sb]
                for _p in range(nparts):
                    def land_b(st=st, g=g, sb=sb):
                        if st["busy"]:
                            raise ProtocolError(f"TMA overwrote B stage {sb} while the tensor core reads it (g={g})")
                        if st["tag"] != g:
                            st["tag"], st["parts"], st["converted"] = g, 0, 0
                        st["parts"] += 1
                    tma(b_full[sb], TILE, land_b)
                g += 1
                yield

    def mma():
        g = 0
        for lt, nkb in enumerate(units):
            ab = lt & 1
            while bug != 'tmem_empty' and not tmem_empty[ab].test(((lt >> 1) & 1) ^ 1):
                yield
            for i in range(nkb):
                s = g % TS
                while not conv_done[s].test((g // TS) & 1):
                    yield
                if presplit:
                    while not b_full[s].test((g // RB) & 1):
                        yield

                def do_mma(g=g, s=s, ab=ab, lt=lt, first=(i == 0)):
                    ta, st = tmem_a[s], b_stage[s]
                    if ta["tag"] != g or ta["quarters"] != {0, 1, 2, 3}:
                        raise ProtocolError(f"MMA of k-block {g} found TMEM slot {s} = {ta}")
                    need_parts = 2 if presplit else 1
                    if st["tag"] != g or st["parts"] != need_parts or (not presplit and st["converted"] != 4):
                        raise ProtocolError(f"MMA of k-block {g} found B stage {s} = {st}")
                    if acc[ab]["draining"]:
                        raise ProtocolError(f"MMA of unit {lt} writes accumulator {ab} while it is being drained")
                    if first:
                        acc[ab]["blocks"] = []
                        acc[ab]["unit"] = lt
                    acc[ab]["blocks"].append(g)
                    ta["busy"] = st["busy"] = False
                tmem_a[s]["busy"] = b_stage[s]["busy"] = True
                tensor_queue.append(do_mma)
                tensor_queue.append(mma_done[s].arrive)            # tcgen05.commit
                if i == nkb - 1:
                    tensor_queue.append(tmem_full[ab].arrive)
                g += 1
                yield

    def epilogue(q):
        g0 = 0
        for lt, nkb in enumerate(units):
            ab = lt & 1
            if nkb > 0:
                while not tmem_full[ab].test((lt >> 1) & 1):
                    yield
                acc[ab]["draining"] += 1
                yield                                              # tcgen05.ld in flight
                want = list(range(g0, g0 + nkb))
                if acc[ab]["unit"] != lt or acc[ab]["blocks"] != want:
                    raise ProtocolError(f"epilogue of unit {lt} found accumulator {ab} = {acc[ab]}, wanted {want}")
                acc[ab]["draining"] -= 1
                tmem_empty[ab].arrive()
                yield                                              # global stores, statistics
            g0 += nkb

    def converter(grp, q):
        g = 0
        for nkb in units:
            for _ in range(nkb):
                if g % NCG == grp:
                    sa, s = g % RA, g % TS
                    while bug != 'slot' and not mma_done[s].test(((g // TS) & 1) ^ 1):
                        yield
                    while not a_full[sa].test((g // RA) & 1):
                        yield
                    if a_stage[sa] != g:
                        raise ProtocolError(f"converter ({grp},{q}) wanted k-block {g} in raw A stage {sa}, found {a_stage[sa]}")
                    yield
                    ta = tmem_a[s]
                    if ta["busy"]:
                        raise ProtocolError(f"converter ({grp},{q}) stores into TMEM slot {s} (k-block {g}) while MMA {ta['tag']} reads it")
                    if ta["tag"] != g:
                        ta["tag"], ta["quarters"] = g, set()
                    ta["quarters"].add(q)
                    a_readers_left[sa] -= 1
                    if not presplit:
                        while not b_full[s].test((g // RB) & 1):
                            yield
                        st = b_stage[s]
                        if st["tag"] != g or st["parts"] != 1:
                            raise ProtocolError(f"converter ({grp},{q}) wanted raw B of k-block {g} in stage {s}, found {st}")
                        if st["busy"]:
                            raise ProtocolError(f"converter ({grp},{q}) rewrites B stage {s} while the tensor core reads it")
                        yield
                        st["converted"] += 1
                    a_empty[sa].arrive()
                    conv_done[s].arrive()
                g += 1
            yield

    roles = [producer(), mma()] + [epilogue(q) for q in range(4)] + [converter(c, q) for c in range(NCG) for q in range(4)]
    alive = list(range(len(roles)))
    steps = 0
    idle = 0
    while alive or async_events or tensor_queue:
        steps += 1
        if steps > max_steps:
            raise ProtocolError("step limit: livelock or deadlock")
        choices = []
        if alive:
            choices += ["role"] * 60
        if async_events:
            choices += ["async"] * (1 if "tma" in slow else 20)
        if tensor_queue:
            choices += ["tensor"] * (1 if "tensor" in slow else 20)
        c = rng.choice(choices)
        if c == "role":
            k = rng.choice(alive)
            if k in slow and rng.random() > 0.02:
                continue
            try:
                next(roles[k])
            except StopIteration:
                alive.remove(k)
        elif c == "async":
            async_events.pop(rng.randrange(len(async_events)))()   # TMA loads complete in any order
        else:
            tensor_queue.pop(0)()                                   # the tensor pipe is in order
        # deadlock detection: nothing asynchronous outstanding and no role can make progress
        if not async_events and not tensor_queue:
            idle += 1
            if idle > 20000:
                raise ProtocolError(f"deadlock: roles {alive} blocked with nothing in flight")
        else:
            idle = 0
    return steps


def random_units(rng, tri=False):
    n = rng.randint(1, 7)
    if tri:
        return [4 * rng.randint(1, 4) for _ in range(n)]
    return [rng.randint(1, 20) for _ in range(n)]


def run_trials(trials=300, seed=0, verbose=False):
    rng = random.Random(seed)
    total = 0
    for t in range(trials):
        units = random_units(rng, tri=bool(t & 1))
        presplit = bool((t >> 1) & 1)
        slow = rng.choice([(), (0,), (1,), (2, 3, 4, 5), (3,), (6, 7, 8, 9), (10,), ("tma",), ("tensor",), (1, "tma")])
        total += simulate(units, presplit=presplit, seed=rng.randrange(1 << 30), slow=slow)
        if verbose and t % 50 == 0:
            print(f"trial {t}: units={units} presplit={presplit} ok")
    return total


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    steps = run_trials(a.trials, a.seed, verbose=True)
    print(f"{a.trials} random schedules, {steps} scheduler steps: protocol consistent (no stale read, no overwrite, no deadlock)")

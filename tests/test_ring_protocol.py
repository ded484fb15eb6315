"""CPU: randomized model check of the smem-ring protocol of ``conv_tc_kernel`` with TWO
MMA-issuing warps (``mma_issuer_alternate`` in ``terran_b200/csrc/conv_tc.cu``).

The model has the actors of the kernel's main loop — the TMA producer, the two issuers that
own alternate global iterations, asynchronous TMA landings and asynchronous ``tcgen05.commit``
arrivals — and mbarriers with the hardware's PARITY semantics (``try_wait.parity P`` succeeds
iff the barrier's current phase parity differs from P).  A random scheduler interleaves them.

It reproduces the hazard found on the GPU (profiles/r01_two_issuers.txt): with an odd number of
ring stages a stage changes owner every ring cycle, an issuer can reach a stage whose previous
phase — the other issuer's — has not completed, and the parity wait passes one phase early.
With the progress words the kernel uses (each issuer publishes how many of its full-barrier
waits have passed; the other waits for that count) no schedule violates the protocol.
"""
import random

import pytest


class MBarrier:
    def __init__(self, count=1):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test_wait(self, parity):
        return (self.phase & 1) != parity


class Violation(Exception):
    pass


def simulate(stages, iterations, seed, progress_words, max_steps=200000):
    """Returns None, or a description of the first protocol violation."""
    rng = random.Random(seed)
    full = [MBarrier() for _ in range(stages)]
    empty = [MBarrier() for _ in range(stages)]
    content = [None] * stages            # which global iteration's operands a stage holds
    prog = [0, 0]
    events = []                          # pending asynchronous completions: (kind, payload)

    def producer():
        for g in range(iterations):
            s, ph = g % stages, (g // stages) & 1
            while not empty[s].test_wait(ph ^ 1):
                yield
            events.append(('land', (s, g)))          # the TMA load is in flight
            yield

    def issuer(w):
        mine = 0
        for g in range(w, iterations, 2):
            s, ph = g % stages, (g // stages) & 1
            if progress_words and (stages & 1) and g >= stages:
                need = (g - stages - (w ^ 1)) // 2 + 1
                while prog[w ^ 1] < need:
                    yield
            while not full[s].test_wait(ph):
                yield
            mine += 1
            prog[w] = mine
            if content[s] != g:
                raise Violation(f'issuer {w} consumed stage {s} holding {content[s]} at iteration {g}')
            yield                                    # MMAs issued
            events.append(('commit', s))             # tcgen05.commit arrives when they retire
            yield

    actors = {'producer': producer(), 'issuer0': issuer(0), 'issuer1': issuer(1)}
    try:
        for _ in range(max_steps):
            choices = list(actors) + ['event'] * min(len(events), 2)
            if not choices:
                return None
            pick = rng.choice(choices)
            if pick == 'event':
                kind, payload = events.pop(rng.randrange(len(events)))    # completions may reorder
                if kind == 'land':
                    s, g = payload
                    content[s] = g
                    full[s].arrive()
                else:
                    content[payload] = None          # the slot is free again
                    empty[payload].arrive()
                continue
            try:
                next(actors[pick])
            except StopIteration:
                del actors[pick]
        return 'no progress (deadlock or live-lock)' if actors else None
    except Violation as v:
        return str(v)


@pytest.mark.parametrize('stages', [2, 3, 4, 5])
def test_two_issuers_with_progress_words_never_violate_the_ring(stages):
    for seed in range(300):
        assert simulate(stages, 40, seed, progress_words=True) is None, (stages, seed)


def test_even_stage_counts_are_safe_without_progress_words():
    """Each issuer then owns fixed stages and consumes every phase of them itself."""
    for stages in (2, 4):
        for seed in range(300):
            assert simulate(stages, 40, seed, progress_words=False) is None, (stages, seed)


def test_odd_stage_count_without_progress_words_reproduces_the_hazard():
    """The fault seen on the GPU: a parity wait passing one phase early."""
    bad = [simulate(3, 40, seed, progress_words=False) for seed in range(300)]
    assert any(b is not None for b in bad)

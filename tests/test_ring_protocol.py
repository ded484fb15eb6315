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


def simulate_segments(stages, segments, seed, max_steps=400000):
    """The whole two-issuer pipeline: ring + double-buffered accumulators handed to the
    epilogue (``tmem_full`` with two arrivals per use, ``tmem_empty`` released by the epilogue;
    an issuer without work in a one-iteration segment arrives plainly after its tmem_empty
    wait).  ``segments`` = iterations per (tile, stream-K range).  Checks that the epilogue
    reads an accumulator only when every MMA of the segment has retired, that no MMA touches an
    accumulator the epilogue has not released, and that each issuer's first MMA of a segment
    overwrites (accumulate = 0) while the others accumulate."""
    rng = random.Random(seed)
    total = sum(segments)
    full = [MBarrier() for _ in range(stages)]
    empty = [MBarrier() for _ in range(stages)]
    tfull = [MBarrier(2), MBarrier(2)]
    tempty = [MBarrier(1), MBarrier(1)]
    content = [None] * stages
    prog = [0, 0]
    events = []
    # accumulator model: per (buffer, issuer) the set of iterations summed so far; owner flag
    acc_sum = [[set(), set()], [set(), set()]]
    acc_busy = [False, False]                # True while the epilogue reads the buffer
    retired = set()                          # iterations whose MMAs have completed
    done_segments = []

    def producer():
        for g in range(total):
            s, ph = g % stages, (g // stages) & 1
            while not empty[s].test_wait(ph ^ 1):
                yield
            events.append(('land', (s, g)))
            yield

    def issuer(w):
        mine, g0 = 0, 0
        for t, n_it in enumerate(segments):
            a, aph = t & 1, (t >> 1) & 1
            first = 0 if (g0 & 1) == w else 1
            while not tempty[a].test_wait(aph ^ 1):
                yield
            if first >= n_it:
                tfull[a].arrive()                     # plain arrive: no work in this segment
                g0 += n_it
                yield
                continue
            last = None
            for i in range(first, n_it, 2):
                g = g0 + i
                s, ph = g % stages, (g // stages) & 1
                if (stages & 1) and g >= stages:
                    need = (g - stages - (w ^ 1)) // 2 + 1
                    while prog[w ^ 1] < need:
                        yield
                while not full[s].test_wait(ph):
                    yield
                mine += 1
                prog[w] = mine
                if content[s] != g:
                    raise Violation(f'issuer {w} consumed stage {s} holding {content[s]} at iteration {g}')
                if acc_busy[a]:
                    raise Violation(f'issuer {w} wrote accumulator {a} while the epilogue reads it')
                if i == first:
                    acc_sum[a][w] = {g}               # accumulate = 0
                else:
                    acc_sum[a][w].add(g)
                yield
                events.append(('commit', (s, g)))
                last = g
            events.append(('tfull', (a, last, w)))    # commit(tmem_full) after my last MMA
            g0 += n_it
            yield

    def epilogue():
        g0 = 0
        for t, n_it in enumerate(segments):
            a, aph = t & 1, (t >> 1) & 1
            while not tfull[a].test_wait(aph):
                yield
            acc_busy[a] = True
            first_w = g0 & 1
            both = n_it >= 2
            got = set(acc_sum[a][first_w]) | (set(acc_sum[a][first_w ^ 1]) if both else set())
            want = set(range(g0, g0 + n_it))
            if got != want:
                raise Violation(f'segment {t}: epilogue read {sorted(got)} instead of {sorted(want)}')
            if not want <= retired:
                raise Violation(f'segment {t}: epilogue started before MMAs {sorted(want - retired)} retired')
            yield
            acc_busy[a] = False
            tempty[a].arrive()
            done_segments.append(t)
            g0 += n_it
            yield

    actors = {'producer': producer(), 'issuer0': issuer(0), 'issuer1': issuer(1), 'epilogue': epilogue()}
    try:
        for _ in range(max_steps):
            choices = list(actors) + ['event'] * min(len(events), 2)
            if not choices:
                return None if len(done_segments) == len(segments) else 'stopped early'
            pick = rng.choice(choices)
            if pick == 'event':
                # tensor-pipe completions of ONE issuing thread retire in its issue order
                idx = rng.randrange(len(events))
                kind, payload = events[idx]
                if kind != 'land':
                    w = payload[1] & 1 if kind == 'commit' else payload[2]
                    idx = next(i for i, (k, pl) in enumerate(events)
                               if (k == 'commit' and (pl[1] & 1) == w) or (k == 'tfull' and pl[2] == w))
                    kind, payload = events[idx]
                events.pop(idx)
                if kind == 'land':
                    s, g = payload
                    content[s] = g
                    full[s].arrive()
                elif kind == 'commit':
                    s, g = payload
                    retired.add(g)
                    content[s] = None
                    empty[s].arrive()
                else:
                    tfull[payload[0]].arrive()
                continue
            try:
                next(actors[pick])
            except StopIteration:
                del actors[pick]
        return 'no progress (deadlock or live-lock)'
    except Violation as v:
        return str(v)


@pytest.mark.parametrize('stages', [2, 3, 4])
def test_two_issuer_pipeline_with_accumulator_handover(stages):
    shapes = [[9] * 6, [5, 5, 5, 5], [1, 9, 1, 2, 3, 1, 1, 7], [4, 9, 9, 5], [2] * 9, [1] * 7]
    for segments in shapes:
        for seed in range(60):
            assert simulate_segments(stages, segments, seed) is None, (stages, segments, seed)
    rng = random.Random(7)
    for seed in range(120):
        segments = [rng.randint(1, 11) for _ in range(rng.randint(3, 9))]
        assert simulate_segments(stages, segments, seed) is None, (stages, segments, seed)


def simulate_halo(stages, tiles, kchunks, iters_kc, seed, max_steps=400000):
    """Halo-path two-issuer protocol (``mma_issuer_alternate_halo``): per chunk-step a resident
    patch in one of two buffers (``pfull`` armed by the TMA, ``pempty`` released by BOTH
    issuers: count 2, plain arrive by an issuer without an iteration in the step), filter
    blocks through the ring, the producer requesting the next step's patch one step ahead."""
    rng = random.Random(seed)
    full = [MBarrier() for _ in range(stages)]
    empty = [MBarrier() for _ in range(stages)]
    pfull = [MBarrier(), MBarrier()]
    pempty = [MBarrier(2), MBarrier(2)]
    tfull = [MBarrier(2), MBarrier(2)]
    tempty = [MBarrier(1), MBarrier(1)]
    content = [None] * stages
    patch = [None, None]                     # chunk-step whose patch a buffer holds
    reading = [0, 0]                         # MMAs in flight that read the buffer
    prog = [0, 0]
    events = []
    n_tile = kchunks * iters_kc
    steps_total = tiles * kchunks
    done = []

    def producer():
        issued = 0

        def issue_patch(q):
            b, ph = q & 1, (q >> 1) & 1
            while not pempty[b].test_wait(ph ^ 1):
                yield
            if reading[b]:
                raise Violation(f'patch buffer {b} overwritten while MMAs read it')
            events.append(('patch', (b, q)))

        g = 0
        for q in range(steps_total):
            if issued == q:
                yield from issue_patch(q)
                issued += 1
            for it in range(iters_kc):
                if it == min(stages, iters_kc - 1) and issued == q + 1 and q + 1 < steps_total:
                    yield from issue_patch(q + 1)
                    issued += 1
                s, ph = g % stages, (g // stages) & 1
                while not empty[s].test_wait(ph ^ 1):
                    yield
                events.append(('land', (s, g)))
                g += 1
                yield

    def issuer(w):
        mine, g, q = 0, 0, 0
        for t in range(tiles):
            a, aph = t & 1, (t >> 1) & 1
            first_tile = 0 if (g & 1) == w else 1
            while not tempty[a].test_wait(aph ^ 1):
                yield
            if first_tile >= n_tile:
                while not pfull[q & 1].test_wait((q >> 1) & 1):
                    yield
                pempty[q & 1].arrive()
                tfull[a].arrive()
                g += n_tile
                q += kchunks
                yield
                continue
            last_mine = max(range(first_tile, n_tile, 2))
            for kc in range(kchunks):
                b = q & 1
                while not pfull[b].test_wait((q >> 1) & 1):
                    yield
                if patch[b] != q:
                    raise Violation(f'issuer {w} step {q}: patch buffer holds {patch[b]}')
                gq = g + kc * iters_kc
                first = 0 if (gq & 1) == w else 1
                its = list(range(first, iters_kc, 2))
                if not its:
                    pempty[b].arrive()
                    q += 1
                    yield
                    continue
                for it in its:
                    gi = gq + it
                    s, ph = gi % stages, (gi // stages) & 1
                    if (stages & 1) and gi >= stages:
                        need = (gi - stages - (w ^ 1)) // 2 + 1
                        while prog[w ^ 1] < need:
                            yield
                    while not full[s].test_wait(ph):
                        yield
                    mine += 1
                    prog[w] = mine
                    if content[s] != gi:
                        raise Violation(f'issuer {w} consumed stage {s} holding {content[s]} at {gi}')
                    reading[b] += 1
                    yield
                    events.append(('commit', (s, gi, b)))
                    if it == its[-1]:
                        events.append(('pempty', (b, gi, w)))
                    if kc * iters_kc + it == last_mine:
                        events.append(('tfull', (a, gi, w)))
                    yield
                q += 1
            g += n_tile

    def epilogue():
        for t in range(tiles):
            a, aph = t & 1, (t >> 1) & 1
            while not tfull[a].test_wait(aph):
                yield
            yield
            tempty[a].arrive()
            done.append(t)
            yield

    actors = {'producer': producer(), 'issuer0': issuer(0), 'issuer1': issuer(1), 'epilogue': epilogue()}

    def issuer_of(kind, payload):
        return payload[1] & 1 if kind == 'commit' else payload[2]

    try:
        for _ in range(max_steps):
            choices = list(actors) + ['event'] * min(len(events), 2)
            if not choices:
                return None if len(done) == tiles else 'stopped early'
            pick = rng.choice(choices)
            if pick == 'event':
                idx = rng.randrange(len(events))
                kind, payload = events[idx]
                if kind in ('commit', 'pempty', 'tfull'):        # one thread's completions are ordered
                    w = issuer_of(kind, payload)
                    idx = next(i for i, (k, pl) in enumerate(events)
                               if k in ('commit', 'pempty', 'tfull') and issuer_of(k, pl) == w)
                    kind, payload = events[idx]
                events.pop(idx)
                if kind == 'land':
                    content[payload[0]] = payload[1]
                    full[payload[0]].arrive()
                elif kind == 'patch':
                    patch[payload[0]] = payload[1]
                    pfull[payload[0]].arrive()
                elif kind == 'commit':
                    s, gi, b = payload
                    content[s] = None
                    reading[b] -= 1
                    empty[s].arrive()
                elif kind == 'pempty':
                    pempty[payload[0]].arrive()
                else:
                    tfull[payload[0]].arrive()
                continue
            try:
                next(actors[pick])
            except StopIteration:
                del actors[pick]
        return 'no progress (deadlock or live-lock)'
    except Violation as v:
        return str(v)


@pytest.mark.parametrize('stages', [2, 3, 4])
def test_halo_two_issuer_protocol(stages):
    for kchunks, iters_kc in [(1, 3), (2, 3), (2, 25), (1, 1), (2, 1), (3, 2), (1, 5)]:
        for seed in range(40):
            r = simulate_halo(stages, 5, kchunks, iters_kc, seed)
            assert r is None, (stages, kchunks, iters_kc, seed, r)

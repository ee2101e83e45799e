"""Protocol model of the two operand-sharing K loops (ss1_pair_merged_kernel, gemm2s_kernel), CPU only.

This is a MODEL of the kernels' synchronisation design, not the kernels: three actors (TMA producer, MMA issuer, epilogue)
written from the schedules in csrc/ss1_kernel.cuh / csrc/gemm2_core.cuh run against mbarriers with the hardware's
phase-parity semantics, with asynchronous TMA / MMA completion fired at random times.  Over many random interleavings and
parameter choices (K blocks, tail / skew, ring depth, tiles per CTA) it checks what the GPU tests can only sample:

  * no deadlock and no parity aliasing (every actor terminates),
  * a shared-memory stage is never overwritten before the MMAs that read it have completed,
  * a TMEM region is never written by an MMA while the epilogue still reads the previous tile out of it, and never read
    before the commit that covers its last MMA has fired,
  * every accumulator receives each of its K blocks exactly once, the first one with accumulate = 0.
"""
import random

import pytest


class MBar:
    """mbarrier with `count` expected arrivals per phase; wait(parity) passes once the phase of that parity completed."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        return self.phase != parity


class World:
    def __init__(self, stages, n_epi, rng):
        self.rng = rng
        self.full = [MBar(1) for _ in range(stages)]
        self.empty = [MBar(1) for _ in range(stages)]
        self.tfull = [MBar(1), MBar(1)]
        self.tempty = [MBar(n_epi), MBar(n_epi)]
        self.tma_q, self.mma_q = [], []            # in-order asynchronous completions
        self.stage_content = [None] * stages       # what the landed TMA put there
        self.stage_busy = [0] * stages             # MMAs issued on the stage and not yet complete
        self.region_readers = [0, 0]               # epilogue warps currently reading the region
        self.region_done = [False, False]          # commit covering the region's last MMA has fired
        self.region_log = [[], []]                 # (tile, product, kb, accumulate) in issue order

    def fire_some(self):
        for q in (self.tma_q, self.mma_q):
            while q and self.rng.random() < 0.6:
                q.pop(0)()


def _run(actors, world, max_steps=2_000_000):
    live = list(actors)
    steps = idle = 0
    while live:
        steps += 1
        assert steps < max_steps
        world.fire_some()
        a = world.rng.choice(live)
        try:
            progressed = next(a)
        except StopIteration:
            live.remove(a)
            continue
        idle = 0 if progressed or world.tma_q or world.mma_q else idle + 1
        assert idle < 10_000, "deadlock: every actor waits and nothing is in flight"
    while world.tma_q or world.mma_q:
        world.fire_some()


def _wait(bar, parity):
    while not bar.passed(parity):
        yield False
    yield True


# ---- schedules: per tile a list of stages, each (A name, [(B name, region role, product name)]) ----------------------
def merged_schedule(K, tail):
    """ss1_pair_merged_kernel: delta.x^T blocks [0, K - tail), merged y^.(delta | y^)^T blocks, delta.x^T blocks [K - tail, K).
    Returns [(part, kb, [(role, product)])]; role 'D' / 'Y'."""
    ks = K - tail if K > tail else 0
    sched = [(0, kb, [("D", "dx")]) for kb in range(ks)]
    sched += [(1, kb, [("D", "yd"), ("Y", "yy")]) for kb in range(K)]
    sched += [(2, kb, [("D", "dx")]) for kb in range(ks, K)]
    return sched


def couple_schedule(K, skew):
    """gemm2s_kernel: half 0 alone on [0, s), both halves on [s, K), half 1 alone on [0, s)."""
    s = min(max(skew, 0), K)
    sched = [(0, kb, [("H0", "h0")]) for kb in range(s)]
    sched += [(1, kb, [("H0", "h0"), ("H1", "h1")]) for kb in range(s, K)]
    sched += [(2, kb, [("H1", "h1")]) for kb in range(s)]
    return sched


def _simulate(kind, K, param, stages, tiles, n_epi, seed, bug=None):
    rng = random.Random(seed)
    w = World(stages, n_epi, rng)
    sched = merged_schedule(K, param) if kind == "merged" else couple_schedule(K, param)

    def regions(seq):
        if kind == "merged":                       # roles swap regions every tile
            d = seq & 1
            return {"D": d, "Y": d ^ 1}
        return {"H0": 0, "H1": 1}

    # which role is complete after which part: merged -> Y after part 1, D after part 2; couples -> H0 after 1, H1 after 2
    done_after = {1: "Y", 2: "D"} if kind == "merged" else {1: "H0", 2: "H1"}
    first_wait = {0: "D", 1: "Y"} if kind == "merged" else {0: "H0", 1: "H1"}     # region that must be free before a part
    tfull_of = {"Y": 1, "D": 0, "H0": 0, "H1": 1}

    def producer():
        stage, phase = 0, 0
        for seq in range(tiles):
            for part, kb, prods in sched:
                if bug != "no_empty_wait":                  # (negative test) refill a stage without waiting for its MMAs
                    yield from _wait(w.empty[stage], phase ^ 1)
                assert w.stage_busy[stage] == 0, "stage overwritten while MMAs still read it"
                content = (seq, part, kb)

                def land(st=stage, c=content):
                    w.stage_content[st] = c
                    w.full[st].arrive()
                w.tma_q.append(land)
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield True

    def mma():
        stage, phase, tph = 0, 0, 0
        for seq in range(tiles):
            reg = regions(seq)
            touched = set()
            last_part = -1
            for part, kb, prods in sched + [(3, 0, [])]:
                # boundaries between parts: commits of completed roles, waits for the regions the next part needs
                for p in range(last_part + 1, part + 1):
                    if p - 1 in done_after and p - 1 >= 0:
                        role = done_after[p - 1]

                        def fire(r=reg[role], t=tfull_of[role]):
                            w.region_done[r] = True
                            w.tfull[t].arrive()
                        w.mma_q.append(fire)
                    if p in first_wait:
                        r = reg[first_wait[p]]
                        yield from _wait(w.tempty[r], tph ^ 1)
                        assert w.region_readers[r] == 0, "MMA writes a TMEM region the epilogue still reads"
                        w.region_done[r] = False
                last_part = part
                if part == 3:
                    break
                yield from _wait(w.full[stage], phase)
                assert w.stage_content[stage] == (seq, part, kb), "MMA consumed a stage holding other data"
                for role, prod in prods:
                    r = reg[role]
                    assert w.region_readers[r] == 0
                    w.region_log[r].append((seq, prod, kb, role in touched))
                    touched.add(role)
                w.stage_busy[stage] += 1

                def done(st=stage):
                    w.stage_busy[st] -= 1
                    w.empty[st].arrive()
                w.mma_q.append(done)
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield True
            tph ^= 1

    def epilogue():
        tph = 0
        order = ["Y", "D"] if kind == "merged" else ["H0", "H1"]
        for seq in range(tiles):
            reg = regions(seq)
            for role in order:
                yield from _wait(w.tfull[tfull_of[role]], tph)
                r = reg[role]
                assert w.region_done[r], "epilogue reads a region before its last MMA completed"
                if bug == "early_release":                  # (negative test) hand the columns back before reading them
                    w.tempty[r].arrive()
                w.region_readers[r] += 1
                for _ in range(rng.randint(0, 6)):          # reading takes a while
                    yield True
                w.region_readers[r] -= 1
                if bug != "early_release":
                    w.tempty[r].arrive()
                yield True
            tph ^= 1

    _run([producer(), mma()] + [epilogue() for _ in range(n_epi)], w)

    # every accumulator received each K block of each of its products exactly once, first write with accumulate = 0
    want = {"merged": {"D": ["dx", "yd"], "Y": ["yy"]}, "couple": {"H0": ["h0"], "H1": ["h1"]}}[kind]
    for seq in range(tiles):
        reg = regions(seq)
        for role, prods in want.items():
            log = [e for e in w.region_log[reg[role]] if e[0] == seq and e[1] in prods]
            assert sorted((p, kb) for _, p, kb, _ in log) == sorted((p, kb) for p in prods for kb in range(K))
            assert [acc for *_, acc in log] == [False] + [True] * (len(log) - 1)


@pytest.mark.parametrize("K,tail", [(35, 4), (35, 0), (35, 40), (3, 4), (1, 1), (9, 8), (35, 34)])
@pytest.mark.parametrize("stages,tiles", [(4, 5), (2, 3), (1, 2), (4, 1)])
def test_merged_stage1_loop_protocol(K, tail, stages, tiles):
    for seed in range(6):
        _simulate("merged", K, tail, stages, tiles, n_epi=2, seed=seed)


@pytest.mark.parametrize("K,skew", [(35, 16), (35, 0), (35, 40), (3, 16), (1, 0), (9, 8), (35, 1)])
@pytest.mark.parametrize("stages,tiles", [(4, 5), (2, 3), (1, 2), (4, 1)])
def test_skewed_couple_loop_protocol(K, skew, stages, tiles):
    for seed in range(6):
        _simulate("couple", K, skew, stages, tiles, n_epi=2, seed=seed)


def _fails_somewhere(kind, param, bug):
    for seed in range(40):
        try:
            _simulate(kind, 12, param, 3, 4, n_epi=2, seed=seed, bug=bug)
        except AssertionError:
            return True
    return False


@pytest.mark.parametrize("kind,param", [("merged", 4), ("couple", 5)])
def test_the_model_catches_broken_protocols(kind, param):
    """The checker is not vacuous: releasing a TMEM region before it is read, or refilling a shared-memory stage without
    waiting for its MMAs, is caught within a few interleavings."""
    assert _fails_somewhere(kind, param, "early_release")
    assert _fails_somewhere(kind, param, "no_empty_wait")

"""CPU model of the hand-over protocol of the warp-specialised WMF row solver (cymf_b200/csrc/als_ws.cu).

The kernel's roles (copy warps, converter warps, the MMA warp, two solver groups of four warps) never tell each other
what comes next: every hand-over is an mbarrier phase whose parity a role derives from its OWN counters, exactly as
restated below (same slot / accumulator / b-slot arithmetic, same arrival counts).  `mbarrier.try_wait.parity P`
only knows the parity of the phase in progress: it succeeds iff that parity differs from P.  A waiter that is a
whole use ahead therefore passes at once -- that was the bug of the first version (every solver group counted ALL
chains, so a group could wait for the second use of an accumulator before the first had completed), found on the
GPU by the kernel's time-out diagnostics.  This model runs the roles as coroutines under random interleavings with
asynchronous completions (cp.async / TMA landing, tcgen05.commit) and checks, for random row lists, that
  * every successful wait really follows the completion of the use it was meant for (no premature pass),
  * the protocol terminates (no deadlock) and every chunk / chain / row is consumed exactly once,
and that the shared-accumulator variant (the first version) is caught by the same checks.
"""
import random

import pytest

NHI, NLO, NACC, NB, NI, NG, CHAIN = 10, 3, 4, 4, 2, 4, 16      # WS_NHI, WS_NLO, WS_NACC, WS_NB, WS_NI, WS_NG, WS_CHAIN


class Mbar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def try_wait(self, parity):                     # what the hardware can tell: parity of the phase in progress
        return (self.phase & 1) != parity


class PrematurePass(AssertionError):
    pass


def wait(bar, use):
    """Coroutine step: wait for completion of use number `use` of `bar` with the parity the kernel computes."""
    while not bar.try_wait(use & 1):
        yield
    if bar.phase < use + 1:
        raise PrematurePass(f"wait for use {use} passed in phase {bar.phase}")


def chunks_of(nnz):
    return (nnz + 31) // 32


def chains_of(nnz):
    return (nnz + 32 * CHAIN - 1) // (32 * CHAIN)


class Model:
    def __init__(self, rows, rng, shared_accumulators=False):
        self.rows, self.rng, self.shared = rows, rng, shared_accumulators
        self.landed = [Mbar(NI) for _ in range(NHI)]            # one (deferred) arrival per copy warp
        self.full = [Mbar(NG) for _ in range(NHI)]
        self.done_hi = [Mbar(1) for _ in range(NHI)]
        self.done_lo = [Mbar(1) for _ in range(NLO)]
        self.acc_full = [Mbar(1) for _ in range(NACC)]
        self.acc_empty = [Mbar(4) for _ in range(NACC)]
        self.b_full = [Mbar(NG) for _ in range(NB)]
        self.b_free = [Mbar(4) for _ in range(NB)]
        self.deferred = []                                      # asynchronous completions: (due step, barrier)
        self.step = 0
        self.folded, self.solved, self.converted, self.issued = [], [], 0, 0

    def later(self, bar, lo=1, hi=40):
        self.deferred.append((self.step + self.rng.randint(lo, hi), bar))

    # ---- roles (one coroutine per warp) -------------------------------------------------------------------------
    def copy_warp(self, w):
        t = 0
        for nnz in self.rows:
            for _ in range(chunks_of(nnz)):
                hs = t % NHI
                if t >= NHI:
                    yield from wait(self.done_hi[hs], t // NHI - 1)
                self.later(self.landed[hs], 1, 120)             # the copies land some time later, in any order
                t += 1
                yield

    def convert_warp(self, w):
        t, nzg = 0, [0, 0]
        for i, nnz in enumerate(self.rows):
            n = chunks_of(nnz)
            for c in range(n):
                hs, ls = t % NHI, t % NLO
                if t >= NLO:
                    yield from wait(self.done_lo[ls], t // NLO - 1)
                yield from wait(self.landed[hs], t // NHI)
                if c == n - 1:
                    og = i & 1
                    nz = nzg[og]
                    bs = 2 * og + (nz & 1)
                    if nz >= 2:
                        yield from wait(self.b_free[bs], (nz >> 1) - 1)
                    self.b_full[bs].arrive()
                    nzg[og] += 1
                self.full[hs].arrive()
                if w == 0:
                    self.converted += 1
                t += 1
                yield

    def mma_warp(self):
        t, chg, chain_all = 0, [0, 0], 0
        for i, nnz in enumerate(self.rows):
            og, n = i & 1, chunks_of(nnz)
            for c in range(n):
                if self.shared:
                    chain, acc, first_use = chain_all, chain_all % NACC, chain_all >= NACC
                    use_prev = chain_all // NACC - 1
                else:
                    chain = chg[og]
                    acc, first_use, use_prev = 2 * og + (chain & 1), chain >= 2, (chain >> 1) - 1
                chain_first = c % CHAIN == 0
                chain_last = c % CHAIN == CHAIN - 1 or c == n - 1
                if chain_first and first_use:
                    yield from wait(self.acc_empty[acc], use_prev)
                yield from wait(self.full[t % NHI], t // NHI)
                self.issued += 1
                due = self.step + self.rng.randint(5, 30)       # tcgen05.commit: arrivals when the MMAs have completed, in order
                if self.deferred_mma and self.deferred_mma[-1][0] >= due:
                    due = self.deferred_mma[-1][0] + 1
                bars = [self.done_hi[t % NHI], self.done_lo[t % NLO]] + ([self.acc_full[acc]] if chain_last else [])
                self.deferred_mma.append((due, bars))
                if chain_last:
                    chg[og] += 1
                    chain_all += 1
                t += 1
                yield

    def solver_warp(self, grp, wq):
        chain, nz, chain_all, nz_all = 0, 0, 0, 0
        for i, nnz in enumerate(self.rows):
            if (i & 1) != grp:
                if nnz > 0:
                    chain_all += chains_of(nnz)
                    nz_all += 1
                continue
            if nnz == 0:
                continue
            for c in range(chains_of(nnz)):
                if self.shared:
                    acc, use = chain_all % NACC, chain_all // NACC
                else:
                    acc, use = 2 * grp + (chain & 1), chain >> 1
                yield from wait(self.acc_full[acc], use)
                for _ in range(self.rng.randint(0, 3)):         # the fold takes a while
                    yield
                self.acc_empty[acc].arrive()
                if wq == 0:
                    self.folded.append((i, c))
                chain += 1
                chain_all += 1
            if self.shared:
                bs, buse = nz_all % NB, nz_all // NB
            else:
                bs, buse = 2 * grp + (nz & 1), nz >> 1
            nz += 1
            nz_all += 1
            yield from wait(self.b_full[bs], buse)
            self.b_free[bs].arrive()
            for _ in range(self.rng.randint(0, 60)):            # CG iterations (the group's named barrier is not modelled)
                yield
            if wq == 0:
                self.solved.append(i)

    # ---- scheduler ----------------------------------------------------------------------------------------------
    def run(self, max_steps=2_000_000):
        self.deferred_mma = []
        warps = [self.copy_warp(w) for w in range(NI)] + [self.convert_warp(w) for w in range(NG)] + [self.mma_warp()]
        warps += [self.solver_warp(g, q) for g in range(2) for q in range(4)]
        live = list(range(len(warps)))
        while live:
            self.step += 1
            if self.step > max_steps:
                raise TimeoutError("deadlock: no role can make progress")
            due = [d for d in self.deferred if d[0] <= self.step]
            self.deferred = [d for d in self.deferred if d[0] > self.step]
            for _, bar in due:
                bar.arrive()
            while self.deferred_mma and self.deferred_mma[0][0] <= self.step:
                for bar in self.deferred_mma.pop(0)[1]:
                    bar.arrive()
            k = self.rng.choice(live)
            try:
                next(warps[k])
            except StopIteration:
                live.remove(k)
        return self


def _row_list(rng, n_rows):
    kinds = [lambda: rng.randint(1, 64), lambda: rng.randint(65, 600), lambda: rng.randint(500, 3000), lambda: 0,
             lambda: 32 * CHAIN, lambda: 32 * CHAIN + 1]
    return [rng.choice(kinds)() for _ in range(n_rows)]


@pytest.mark.parametrize("seed", range(12))
def test_protocol_terminates_without_premature_passes(seed):
    rng = random.Random(seed)
    rows = _row_list(rng, rng.randint(0, 40))
    m = Model(rows, rng).run()
    n_chunks = sum(chunks_of(n) for n in rows)
    assert m.issued == n_chunks and m.converted == n_chunks
    assert sorted(m.folded) == [(i, c) for i, n in enumerate(rows) for c in range(chains_of(n))]
    assert sorted(m.solved) == [i for i, n in enumerate(rows) if n > 0]
    for g in (0, 1):
        mine = [i for i in m.solved if (i & 1) == g]
        assert mine == sorted(mine)


def test_many_single_chain_rows_and_one_very_long_row():
    rng = random.Random(99)
    Model([40] * 60, rng).run()                                 # every slot, accumulator and b slot reused many times
    Model([32 * CHAIN * 9 + 5, 70, 70, 70], rng).run()          # one row cycling its group's two accumulators


def test_shared_accumulators_are_caught():
    """The first version's numbering (chains and b slots shared by both groups, each group skipping the other's):
    a group waits for the second use of an accumulator while the first is still in progress, and passes."""
    caught = 0
    for seed in range(20):
        rng = random.Random(seed)
        rows = [32 * CHAIN * 4] + [rng.randint(60, 120) for _ in range(8)]          # the failing shape: 4 chains, then short rows
        try:
            Model(rows, rng, shared_accumulators=True).run(max_steps=300_000)
        except (PrematurePass, TimeoutError, AssertionError):
            caught += 1
    assert caught >= 15

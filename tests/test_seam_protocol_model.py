"""Discrete-event model of the synchronisation protocol of `bottleneck_next_kernel`
(handmvnet_b200/csrc/bottleneck_next_tc.cu): four roles (operand producer, MMA issuer, slot producer, epilogue),
mbarriers with PTX phase/parity semantics, and asynchronous completions (TMA loads, tcgen05.commit, TMA store reads)
that finish after random delays.  It cannot prove the CUDA code right, but it checks the DESIGN of the hand-shakes for
any (stages, slots, lag) setting without a GPU:

  * every role runs to completion for 1..3 tiles per CTA under many random timings (no deadlock);
  * no operand stage is refilled while MMAs that read it are in flight;
  * no chunk slot is refilled (residual TMA load) or rewritten while a TMA store or an MMA still reads it;
  * no accumulator is overwritten before the epilogue has drained it, and none is read before its MMAs retired.

The constants below mirror the kernel; `test_shipped_configuration` pins the combination that is compiled in.
"""
import heapq
import random
import re
import os

import pytest

NCH = 8            # conv3 chunks per tile (kBnNch)
KB3 = 4            # K blocks per conv3 chunk (kBnKb3)
E1_SLOTS = 4       # conv1 output slots per tile (kBnP / kBnSlotCols)
SLOTS_PER_TILE = 2 * NCH + E1_SLOTS


class MBar:
    """mbarrier: `count` arrivals (+ outstanding transaction bytes) complete a phase."""

    def __init__(self, count, name):
        self.count, self.pending, self.tx, self.phase, self.name = count, count, 0, 0, name

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, f"{self.name}: more arrivals than the barrier count in one phase"
        self.pending -= 1
        self._maybe_flip()

    def arrive_expect_tx(self, n):
        self.tx += n
        self.arrive()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_flip()

    def ready(self, parity):               # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, stages, slots, lag, tiles, seed, fault=None):
        self.S, self.NS, self.lag, self.tiles, self.fault = stages, slots, lag, tiles, fault
        self.rng = random.Random(seed)
        self.now = 0.0
        self.events = []                   # (time, seq, fn)
        self.seq = 0
        self.full = [MBar(1, f"full{i}") for i in range(stages)]
        self.empty = [MBar(1, f"empty{i}") for i in range(stages)]
        self.t3full = [MBar(1, f"t3full{i}") for i in range(2)]
        self.t3empty = [MBar(8, f"t3empty{i}") for i in range(2)]
        self.t1full, self.t1empty = MBar(1, "t1full"), MBar(8, "t1empty")
        self.sres = [MBar(1, f"sres{i}") for i in range(slots)]
        self.aready = [MBar(8, f"aready{i}") for i in range(slots)]
        self.sfree = [MBar(5, f"sfree{i}") for i in range(slots)]
        # hazard tracking
        self.stage_readers = [0] * stages          # MMAs in flight that read the stage
        self.slot_readers = [0] * slots            # TMA stores / MMAs in flight that read the slot
        self.acc3_busy = [False, False]            # written by MMAs, not yet drained
        self.acc3_ready = [False, False]
        self.acc1_busy = False
        self.mma_queue_time = 0.0                  # MMAs retire in issue order

    # ---- async machinery ----
    def later(self, lo, hi, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + self.rng.uniform(lo, hi), self.seq, fn))

    def mma_retire(self, fns):
        """tcgen05.commit: `fns` run once every MMA issued so far has retired (in order)."""
        self.mma_queue_time = max(self.mma_queue_time, self.now) + self.rng.uniform(0.1, 1.0)
        self.seq += 1
        heapq.heappush(self.events, (self.mma_queue_time, self.seq, lambda: [f() for f in fns]))

    # ---- schedule shared by the operand producer and the MMA issuer (bn_tile_schedule) ----
    def schedule(self):
        ops = []
        for c in range(NCH):
            ops.append(("t3", c))
            if c >= self.lag:
                ops.append(("t1", c - self.lag))
        for c in range(NCH - self.lag, NCH):
            ops.append(("t1", c))
        return ops

    # ---- roles (generators yielding a predicate to wait on) ----
    def producer(self):
        stage, phase = 0, 0
        for _ in range(self.tiles):
            for kind, _c in self.schedule():
                for _ in range(KB3 if kind == "t3" else 2):
                    yield lambda s=stage, p=phase: self.empty[s].ready(p ^ 1)
                    assert self.stage_readers[stage] == 0, "stage refilled while MMAs still read it"
                    self.full[stage].arrive_expect_tx(1)
                    self.later(0.5, 3.0, lambda s=stage: self.full[s].complete_tx(1))
                    stage += 1
                    if stage == self.S:
                        stage, phase = 0, phase ^ 1

    def mma(self):
        stage, phase, q3 = 0, 0, 0
        for i in range(self.tiles):
            gbase = i * SLOTS_PER_TILE
            for kind, c in self.schedule():
                if kind == "t3":
                    s, use = q3 & 1, q3 >> 1
                    yield lambda s=s, use=use: self.t3empty[s].ready((use & 1) ^ 1)
                    assert not self.acc3_busy[s], "conv3 accumulator overwritten before it was drained"
                    self.acc3_busy[s] = True
                    for _ in range(KB3):
                        yield lambda st=stage, p=phase: self.full[st].ready(p)
                        self.stage_readers[stage] += 1
                        self.mma_retire([lambda st=stage: self._stage_done(st)])
                        stage += 1
                        if stage == self.S:
                            stage, phase = 0, phase ^ 1
                    self.mma_retire([lambda s=s: self._acc3_full(s)])
                    q3 += 1
                else:
                    for j in range(2):
                        g = gbase + 2 * c + j
                        slot, use = g % self.NS, g // self.NS
                        if c == 0 and j == 0:
                            yield lambda i=i: self.t1empty.ready((i & 1) ^ 1)
                            assert not self.acc1_busy, "conv1 accumulator overwritten before it was drained"
                            self.acc1_busy = True
                        yield lambda slot=slot, use=use: self.aready[slot].ready(use & 1)
                        yield lambda st=stage, p=phase: self.full[st].ready(p)
                        self.stage_readers[stage] += 1
                        self.slot_readers[slot] += 1
                        self.mma_retire([lambda st=stage: self._stage_done(st), lambda sl=slot: self._slot_mma_done(sl)])
                        stage += 1
                        if stage == self.S:
                            stage, phase = 0, phase ^ 1
                    if c == NCH - 1:
                        self.mma_retire([self.t1full.arrive])

    def _stage_done(self, st):
        self.stage_readers[st] -= 1
        self.empty[st].arrive()

    def _slot_mma_done(self, sl):
        self.slot_readers[sl] -= 1
        self.sfree[sl].arrive()

    def _acc3_full(self, s):
        self.acc3_ready[s] = True
        self.t3full[s].arrive()

    def slot_producer(self):
        g = 0
        for _ in range(self.tiles):
            for c in range(SLOTS_PER_TILE):
                slot, use = g % self.NS, g // self.NS
                yield lambda slot=slot, use=use: self.sfree[slot].ready((use & 1) ^ 1)
                assert self.slot_readers[slot] == 0, "slot refilled while a store / MMA still reads it"
                if c < 2 * NCH:
                    self.sres[slot].arrive_expect_tx(1)
                    self.later(0.5, 4.0, lambda sl=slot: self.sres[sl].complete_tx(1))
                else:
                    self.sres[slot].arrive()
                    if self.fault != "missing_conv1_slot_arrival":
                        self.sfree[slot].arrive()      # stands in for the MMA commit
                g += 1

    def epilogue(self):
        """The 8 epilogue warps move in lock step through named barriers: one role, 8 arrivals where each warp arrives."""
        g, q3, pending = 0, 0, None
        store_fifo_time = [0.0]

        def do_slot(slot, use, release, is_conv3):
            nonlocal g, pending
            yield lambda: self.sres[slot].ready(use & 1)
            assert self.slot_readers[slot] == 0, "slot rewritten while a store / MMA still reads it"
            if release is not None:
                release()
            if is_conv3 or self.fault != "aready_only_for_conv3_slots":
                for _ in range(8):
                    self.aready[slot].arrive()         # every slot use: keeps the barrier's phase equal to the use count
            # TMA store of the slot (4 issuers): reads complete in order, some time later
            self.slot_readers[slot] += 1
            store_fifo_time[0] = max(store_fifo_time[0], self.now) + self.rng.uniform(0.2, 2.0)
            done_at = store_fifo_time[0]
            state = {"done": False}
            self.seq += 1
            heapq.heappush(self.events, (done_at, self.seq, lambda sl=slot, st=state: self._store_read_done(sl, st)))
            if pending is not None:                    # bulk_wait_read<1>: the PREVIOUS store has been read
                prev_slot, prev_state = pending
                yield lambda st=prev_state: st["done"]
                for _ in range(4):
                    self.sfree[prev_slot].arrive()
            pending = (slot, state)
            g += 1

        for i in range(self.tiles):
            for c in range(NCH):
                s = q3 & 1
                yield lambda s=s, q=q3: self.t3full[s].ready((q >> 1) & 1)
                assert self.acc3_ready[s], "conv3 accumulator read before its MMAs retired"
                for cc in range(2):
                    rel = None
                    if cc == 1:
                        def rel(s=s):
                            self.acc3_busy[s] = False
                            self.acc3_ready[s] = False
                            for _ in range(8):
                                self.t3empty[s].arrive()
                    yield from do_slot(g % self.NS, g // self.NS, rel, True)
                q3 += 1
            yield lambda i=i: self.t1full.ready(i & 1)
            for cc in range(E1_SLOTS):
                rel = None
                if cc == E1_SLOTS - 1:
                    def rel():
                        self.acc1_busy = False
                        for _ in range(8):
                            self.t1empty.arrive()
                yield from do_slot(g % self.NS, g // self.NS, rel, False)

    def _store_read_done(self, sl, st):
        st["done"] = True
        self.slot_readers[sl] -= 1

    # ---- driver ----
    def run(self):
        roles = {"producer": self.producer(), "mma": self.mma(), "slots": self.slot_producer(), "epilogue": self.epilogue()}
        waiting = {}
        for name, gen in list(roles.items()):
            try:
                waiting[name] = next(gen)
            except StopIteration:
                del roles[name]
        steps = 0
        while roles:
            steps += 1
            assert steps < 2_000_000, "model did not terminate"
            progressed = False
            for name in list(roles):
                while name in roles and waiting[name]():
                    progressed = True
                    try:
                        waiting[name] = next(roles[name])
                    except StopIteration:
                        del roles[name]
                        del waiting[name]
            if not roles:
                break
            if not progressed:
                if not self.events:
                    raise AssertionError(f"deadlock: {sorted(roles)} blocked with no event pending "
                                         f"(stages={self.S}, slots={self.NS}, lag={self.lag})")
                t, _, fn = heapq.heappop(self.events)
                self.now = max(self.now, t)
                fn()
        return True


CONFIGS = [(5, 4, 1), (4, 6, 1), (3, 8, 2)]      # (stages, slots, lag): the three settings measured on the GPU (all ran correctly)


@pytest.mark.parametrize("stages,slots,lag", CONFIGS)
@pytest.mark.parametrize("tiles", [1, 2, 3])
def test_protocol_has_no_deadlock_and_no_hazard(stages, slots, lag, tiles):
    for seed in range(25):
        assert Sim(stages, slots, lag, tiles, seed).run()


@pytest.mark.parametrize("stages,slots,lag", [(2, 2, 1), (1, 2, 1), (8, 3, 1), (6, 4, 2)])
def test_other_settings_are_deadlock_free_too(stages, slots, lag):
    """The hand-shakes do not depend on how many stages / slots there are (fewer only serialise more) ..."""
    for seed in range(10):
        assert Sim(stages, slots, lag, 2, seed).run()


def test_a_single_slot_deadlocks():
    """... down to two slots: a slot is handed back when the NEXT slot's store has been committed (`bulk_wait_read<1>`),
    so with one slot the next store can never be issued."""
    with pytest.raises(AssertionError, match="deadlock"):
        Sim(4, 1, 1, 1, 0).run()


@pytest.mark.parametrize("fault", ["missing_conv1_slot_arrival", "aready_only_for_conv3_slots"])
def test_model_detects_protocol_faults(fault):
    """Two mistakes that were considered while writing the kernel must show up in the model: a conv1 slot that never gets its
    5th `sfree` arrival (nothing replaces the MMA commit) dead-locks as soon as the slot is reused; an `aready` barrier that
    is only arrived on for conv3 slots falls out of step with the use count, so the MMA issuer reads a slot too early
    (hazard) or waits for ever."""
    with pytest.raises(AssertionError):
        for seed in range(25):
            Sim(5, 4, 1, 3, seed, fault=fault).run()


def test_shipped_configuration():
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "handmvnet_b200", "csrc",
                            "bottleneck_next_tc.cu")).read()
    got = tuple(int(re.search(rf"constexpr int {k} = (\d+);", src).group(1)) for k in ("kBnStages", "kBnSlots", "kBnLag"))
    assert got in CONFIGS, f"kernel constants {got} are not covered by the protocol model: add them to CONFIGS"
    assert int(re.search(r"constexpr int kBnNch = kBnN3 / kBnChunk;", src) is not None)

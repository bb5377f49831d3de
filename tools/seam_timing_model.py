"""Timing model of `bottleneck_next_kernel` (layer3 seam) for exploring (stages, slots, lag) and schedule variants
without a GPU.  Same protocol as tests/test_seam_protocol_model.py, but with deterministic latencies in SM cycles:

    TMA operand load (L2 hit)      L_OP   cycles from issue to `full`
    TMA residual load (HBM)        L_RES
    one K block of MMAs            C_MMA  (4 tcgen05.mma, M=128, K=16: ~128 cycles each whatever N is)
    epilogue work per 64-col slot  C_EPI  (tcgen05.ld + bias/residual/ReLU + st.shared + fence + barrier)
    TMA store smem read            L_ST

Calibration (B=64, 2560 tiles on 148 SMs = 17.3 tiles per CTA, 1.6 GHz): measured 0.50 / 0.457 / 0.40 ms for
(3, 8, 2) / (4, 6, 1) / (5, 4, 1) = 46k / 42k / 37k cycles per tile.

    python tools/seam_timing_model.py

Outcome (kept as a record): the fit needs a per-SM TMA fill port of ~48 B/clk and then predicts -20 % from multicasting the
weight tiles over a 2-CTA cluster; the variant was built (HMV_SEAM_CLUSTER=1) and measured -1 %, i.e. the fill-bandwidth
hypothesis is wrong and the kernel is bound by a latency chain the model does not capture.
"""
import heapq
import itertools

NCH, KB3, E1_SLOTS = 8, 4, 4
SLOTS_PER_TILE = 2 * NCH + E1_SLOTS


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0
        self.waiters = []

    def _flip(self, sim):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count
            w, self.waiters = self.waiters, []
            for role in w:
                sim.wake(role)

    def arrive(self, sim):
        self.pending -= 1
        self._flip(sim)

    def expect(self, sim):
        self.tx += 1
        self.arrive(sim)

    def complete(self, sim):
        self.tx -= 1
        self._flip(sim)

    def ready(self, parity):
        return (self.phase & 1) != parity


class Model:
    def __init__(self, stages, slots, lag, tiles=4, L_OP=1400, L_RES=2200, C_MMA=512, C_EPI=1350, L_ST=400, order=None,
                 BW=None, op_bytes=(32768, 32768), res_bytes=16384):
        # BW: bytes per cycle the SM can take in through TMA (None = unlimited); op_bytes = (T3 stage, T1 stage)
        self.BW, self.op_bytes, self.res_bytes = BW, op_bytes, res_bytes
        self.port_free = 0
        self.S, self.NS, self.lag, self.tiles = stages, slots, lag, tiles
        self.L_OP, self.L_RES, self.C_MMA, self.C_EPI, self.L_ST = L_OP, L_RES, C_MMA, C_EPI, L_ST
        self.order = order
        self.now = 0
        self.ev = []
        self.seq = itertools.count()
        self.full = [Bar(1) for _ in range(stages)]
        self.empty = [Bar(1) for _ in range(stages)]
        self.t3full = [Bar(1), Bar(1)]
        self.t3empty = [Bar(8), Bar(8)]
        self.t1full, self.t1empty = Bar(1), Bar(8)
        self.sres = [Bar(1) for _ in range(slots)]
        self.aready = [Bar(8) for _ in range(slots)]
        self.sfree = [Bar(5) for _ in range(slots)]
        self.pipe_free = 0              # tensor pipe: MMAs execute in order
        self.tile_done = []
        self.runnable = []
        self.mma_busy = 0
        self.epi_busy = 0

    def load(self, nbytes, latency, fn):
        """A TMA load: queued on the SM's fill port (FIFO, BW bytes / cycle), complete `latency` cycles after its last byte."""
        if self.BW is None:
            self.at(self.now + latency, fn)
            return
        start = max(self.port_free, self.now)
        self.port_free = start + nbytes / self.BW
        self.at(self.port_free + latency, fn)

    def at(self, t, fn):
        heapq.heappush(self.ev, (t, next(self.seq), fn))

    def wake(self, role):
        self.at(self.now, lambda: self.step(role))

    def schedule(self):
        if self.order is not None:
            return self.order
        ops = []
        for c in range(NCH):
            ops.append(("t3", c))
            if c >= self.lag:
                ops.append(("t1", c - self.lag))
        for c in range(NCH - self.lag, NCH):
            ops.append(("t1", c))
        return ops

    # roles yield ("wait", bar, parity) or ("busy", cycles)
    def producer(self):
        st, ph = 0, 0
        for _ in range(self.tiles):
            for kind, _c in self.schedule():
                for _ in range(KB3 if kind == "t3" else 2):
                    yield ("wait", self.empty[st], ph ^ 1)
                    self.full[st].expect(self)
                    self.load(self.op_bytes[0] if kind == "t3" else self.op_bytes[1], self.L_OP, lambda s=st: self.full[s].complete(self))
                    yield ("busy", 40)
                    st += 1
                    if st == self.S:
                        st, ph = 0, ph ^ 1

    def commit(self, fns):
        t = max(self.pipe_free, self.now)
        self.at(t, lambda: [f() for f in fns])

    def issue_kblock(self):
        start = max(self.pipe_free, self.now)
        self.pipe_free = start + self.C_MMA
        self.mma_busy += self.C_MMA

    def mma(self):
        st, ph, q3 = 0, 0, 0
        for i in range(self.tiles):
            gbase = i * SLOTS_PER_TILE
            for kind, c in self.schedule():
                if kind == "t3":
                    s, use = q3 & 1, q3 >> 1
                    yield ("wait", self.t3empty[s], (use & 1) ^ 1)
                    for _ in range(KB3):
                        yield ("wait", self.full[st], ph)
                        self.issue_kblock()
                        self.commit([lambda x=st: self.empty[x].arrive(self)])
                        yield ("busy", 60)
                        st += 1
                        if st == self.S:
                            st, ph = 0, ph ^ 1
                    self.commit([lambda x=s: self.t3full[x].arrive(self)])
                    q3 += 1
                else:
                    for j in range(2):
                        g = gbase + 2 * c + j
                        slot, use = g % self.NS, g // self.NS
                        if c == 0 and j == 0:
                            yield ("wait", self.t1empty, (i & 1) ^ 1)
                        yield ("wait", self.aready[slot], use & 1)
                        yield ("wait", self.full[st], ph)
                        self.issue_kblock()
                        self.commit([lambda x=st: self.empty[x].arrive(self), lambda x=slot: self.sfree[x].arrive(self)])
                        yield ("busy", 60)
                        st += 1
                        if st == self.S:
                            st, ph = 0, ph ^ 1
                    if c == NCH - 1:
                        self.commit([lambda: self.t1full.arrive(self)])

    def slots(self):
        g = 0
        for _ in range(self.tiles):
            for c in range(SLOTS_PER_TILE):
                slot, use = g % self.NS, g // self.NS
                yield ("wait", self.sfree[slot], (use & 1) ^ 1)
                if c < 2 * NCH:
                    self.sres[slot].expect(self)
                    self.load(self.res_bytes, self.L_RES, lambda x=slot: self.sres[x].complete(self))
                else:
                    self.sres[slot].arrive(self)
                    self.sfree[slot].arrive(self)
                yield ("busy", 40)
                g += 1

    def epilogue(self):
        g, q3 = 0, 0
        pending = None
        for i in range(self.tiles):
            for c in range(NCH):
                s = q3 & 1
                yield ("wait", self.t3full[s], (q3 >> 1) & 1)
                for cc in range(2):
                    slot, use = g % self.NS, g // self.NS
                    yield ("wait", self.sres[slot], use & 1)
                    yield ("busy", self.C_EPI)
                    self.epi_busy += self.C_EPI
                    if cc == 1:
                        for _ in range(8):
                            self.t3empty[s].arrive(self)
                    for _ in range(8):
                        self.aready[slot].arrive(self)
                    done = Bar(1)
                    self.at(self.now + self.L_ST, lambda b=done: b.arrive(self))
                    if pending is not None:
                        pslot, pdone = pending
                        yield ("wait", pdone, 0)
                        for _ in range(4):
                            self.sfree[pslot].arrive(self)
                    pending = (slot, done)
                    g += 1
                q3 += 1
            yield ("wait", self.t1full, i & 1)
            for cc in range(E1_SLOTS):
                slot, use = g % self.NS, g // self.NS
                yield ("wait", self.sres[slot], use & 1)
                yield ("busy", self.C_EPI * 0.8)
                self.epi_busy += self.C_EPI * 0.8
                if cc == E1_SLOTS - 1:
                    for _ in range(8):
                        self.t1empty.arrive(self)
                for _ in range(8):
                    self.aready[slot].arrive(self)
                done = Bar(1)
                self.at(self.now + self.L_ST, lambda b=done: b.arrive(self))
                if pending is not None:
                    pslot, pdone = pending
                    yield ("wait", pdone, 0)
                    for _ in range(4):
                        self.sfree[pslot].arrive(self)
                pending = (slot, done)
                g += 1
            self.tile_done.append(self.now)

    def step(self, role):
        gen = role["gen"]
        while True:
            req = role.get("req")
            if req is None:
                try:
                    req = next(gen)
                except StopIteration:
                    role["done"] = True
                    return
            if req[0] == "wait":
                _, bar, parity = req
                if bar.ready(parity):
                    role["req"] = None
                    continue
                role["req"] = req
                bar.waiters.append(role)
                return
            role["req"] = None
            self.at(self.now + req[1], lambda r=role: self.step(r))
            return

    def run(self):
        roles = [{"gen": g()} for g in (self.producer, self.mma, self.slots, self.epilogue)]
        for r in roles:
            self.wake(r)
        while self.ev:
            t, _, fn = heapq.heappop(self.ev)
            self.now = max(self.now, t)
            fn()
        assert all(r.get("done") for r in roles), "model deadlocked"
        per_tile = (self.tile_done[-1] - self.tile_done[0]) / (len(self.tile_done) - 1)
        return per_tile, self.mma_busy / self.tiles, self.epi_busy / self.tiles


def calibrate():
    """Which (fill bandwidth, operand latency) reproduces the three measured settings?"""
    meas = {(3, 8, 2): 46000, (4, 6, 1): 42000, (5, 4, 1): 37000}
    best = None
    for bw in (32, 40, 48, 56, 64, 80, 96, 128, None):
        for lop in (800, 1400, 2000, 3000, 4000):
            err = 0.0
            row = []
            for cfg, m in meas.items():
                t, _, _ = Model(*cfg, BW=bw, L_OP=lop, L_RES=lop + 800).run()
                row.append(t)
                err += (t / m - 1) ** 2
            if best is None or err < best[0]:
                best = (err, bw, lop, row)
    print("best fit: fill bandwidth", best[1], "B/clk, operand latency", best[2], "cycles ->",
          [f"{v:.0f}" for v in best[3]], "vs measured [46000, 42000, 37000]; rms rel err %.3f" % (best[0] / 3) ** 0.5)
    return best[1], best[2]


def main():
    bw, lop = calibrate()
    print("\nvariants at the fitted parameters (cycles per tile; 17.3 tiles per CTA at B=64):")
    base = dict(BW=bw, L_OP=lop, L_RES=lop + 800)
    for name, cfg, kw in [
        ("shipped (5 stages, 4 slots)", (5, 4, 1), {}),
        ("+ W3 / W1 multicast over a 2-CTA cluster (weights cross L2->SM once per pair)", (5, 4, 1), dict(op_bytes=(16384 + 8192, 16384))),
        ("+ conv2 output tile resident (no A reload per chunk)", (5, 4, 1), dict(op_bytes=(16384, 32768))),
        ("+ both", (5, 4, 1), dict(op_bytes=(8192, 16384))),
    ]:
        t, mb, eb = Model(*cfg, **base, **kw).run()
        print(f"  {name:82s} {t:8.0f}   {t * 17.3 / 1.6e9 * 1e3:.3f} ms")
    print()
    print("measured: (3,8,2) 46k  (4,6,1) 42k  (5,4,1) 37k cycles per tile\n")
    print(f"{'stages':>6} {'slots':>5} {'lag':>3} {'cycles/tile':>12} {'mma busy':>9} {'epi busy':>9}   ms per B=64 launch @1.6 GHz")
    for cfg in [(3, 8, 2), (4, 6, 1), (5, 4, 1), (6, 2, 1), (5, 4, 2), (7, 4, 1), (8, 6, 1), (10, 8, 2)]:
        t, mb, eb = Model(*cfg, **base).run()
        print(f"{cfg[0]:6d} {cfg[1]:5d} {cfg[2]:3d} {t:12.0f} {mb:9.0f} {eb:9.0f}   {t * 17.3 / 1.6e9 * 1e3:.3f}"
              + ("   (needs more shared memory than one CTA has)" if 32 * cfg[0] + 16 * cfg[1] > 224 else ""))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Executable model of the mbarrier protocol of conv_umma_kernel (3d-wsis_b200/csrc/conv_umma.cu).

The kernel is a set of cooperating loops (record producer, weight producers, gather warps, builder groups, MMA
issuers, epilogue warps) that hand shared-memory / tensor-memory buffers to each other through mbarriers whose waits
only test a PARITY bit.  A parity wait is sound only while the waiter can never be two phases away from the phase it
means; one version of the kernel violated that (more builder warps than operand stages) and hung on the GPU.  This
model transcribes the control flow of every role -- the same ring indices, the same parities, the same arrival
counts -- and runs the roles under a random scheduler while checking, at every wait that passes, that the LOGICAL
generation the waiter needs has really completed, and at every buffer use that the buffer holds what the consumer
expects.  A violation or a deadlock raises.  tests/test_cpu.py runs it over every launch plan wsis_conv_umma_plan
can produce.

    python tools/protocol_model.py            # quick self-check
"""
import random


class Barrier(object):
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0   # phase = number of completed phases

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def test(self, parity):                                       # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class ProtocolError(AssertionError):
    pass


def wait(bar, parity, needed_phases, what):
    """Generator step: blocks until the parity test passes, then checks the logical condition."""
    while not bar.test(parity):
        yield
    if bar.phase < needed_phases:
        raise ProtocolError("parity wait passed early: %s needs %d completed phases, barrier has %d"
                            % (what, needed_phases, bar.phase))


class Cta(object):
    """One persistent CTA working through `tiles` = list of (number of active offsets, KB).  A pipeline STAGE holds up
    to `us` consecutive active offsets of one (tile, channel block) pass."""

    def __init__(self, tiles, na, nrc, nrec, nbg, nmma, nbuf=2, resident=False, nwp=2, us=4, gather_warps=4, epi_warps=4):
        assert na & (na - 1) == 0
        self.tiles, self.na, self.nrc, self.nrec, self.nbg, self.nmma = tiles, na, nrc, nrec, nbg, nmma
        self.nbuf, self.resident, self.nwp, self.us = nbuf, resident, nwp, us
        self.G, self.E = gather_warps, epi_warps
        self.lna = na.bit_length() - 1
        self.afull = [Barrier(4 + (0 if resident else 1)) for _ in range(na)]  # 4 builder warps (+ weight expect_tx)
        self.aempty = [Barrier(1) for _ in range(na)]                      # the owning issuer's tcgen05.commit
        self.rcf = [Barrier(gather_warps) for _ in range(nrc)]
        self.rce = [Barrier(4 * nbg) for _ in range(nrc)]
        self.recf = [Barrier(1) for _ in range(nrec)]
        self.rece = [Barrier(4 * nbg + nmma) for _ in range(nrec)]
        self.nul = 4                                                       # ring of unique-row lists (gatherers)
        self.ulf = [Barrier(1) for _ in range(self.nul)]
        self.ule = [Barrier(gather_warps) for _ in range(self.nul)]
        self.ul = [None] * self.nul
        self.accf = [Barrier(nmma) for _ in range(2)]
        self.acce = [Barrier(epi_warps) for _ in range(2)]
        # buffer contents (what the consumer must find)
        self.stage_rows = [[None] * 4 for _ in range(na)]                  # stage id written by each builder warp
        self.stage_w = [None] * na                                         # stage id of the weight blocks
        self.rc = [None] * nrc                                             # pass id
        self.rec = [None] * nrec                                           # tile iteration
        self.acc_tile = [None, None]                                       # tile whose MMAs went into the buffer
        self.done_stages, self.done_tiles = [], []

    def _nq(self, nact):
        return -(-nact // self.us)

    # ---- roles (each a generator; `yield` = be descheduled) ----
    def record_producer(self):
        for it in range(len(self.tiles)):
            rb = it % self.nrec
            yield from wait(self.rece[rb], ((it // self.nrec) & 1) ^ 1, it // self.nrec, "record buffer free")
            self.rec[rb] = it
            yield
            self.recf[rb].arrive()                                         # expect_tx arrive + bytes landed

    def list_producer(self):
        for it in range(len(self.tiles)):
            ub = it % self.nul
            yield from wait(self.ule[ub], ((it // self.nul) & 1) ^ 1, it // self.nul, "list buffer free")
            self.ul[ub] = it
            yield
            self.ulf[ub].arrive()

    def weight_producer(self, wi):
        if self.resident:
            return
        Q = 0
        for nact, KB in self.tiles:
            for _ in range(KB * self._nq(nact)):
                if Q % self.nwp == wi:
                    s = Q & (self.na - 1)
                    yield from wait(self.aempty[s], ((Q >> self.lna) & 1) ^ 1, Q >> self.lna, "weight slot free")
                    self.stage_w[s] = Q
                    yield
                    self.afull[s].arrive()
                Q += 1

    def gatherer(self, g):
        q = 0
        for it, (nact, KB) in enumerate(self.tiles):
            ub = it % self.nul
            yield from wait(self.ulf[ub], (it // self.nul) & 1, it // self.nul + 1, "list ready (gatherer)")
            if self.ul[ub] != it:
                raise ProtocolError("gatherer reads list of tile %s, wants %d" % (self.ul[ub], it))
            for _ in range(KB):
                slot = q % self.nrc
                yield                                                       # loads in flight
                yield from wait(self.rce[slot], ((q // self.nrc) & 1) ^ 1, q // self.nrc, "row cache free")
                if g == 0:
                    self.rc[slot] = q
                yield
                self.rcf[slot].arrive()
                q += 1
            self.ule[ub].arrive()

    def builder(self, bw):
        g, w4 = bw // 4, bw % 4
        q = Q = 0
        for it, (nact, KB) in enumerate(self.tiles):
            rb = it % self.nrec
            yield from wait(self.recf[rb], (it // self.nrec) & 1, it // self.nrec + 1, "record ready (builder)")
            if self.rec[rb] != it:
                raise ProtocolError("builder reads record of tile %s, wants %d" % (self.rec[rb], it))
            for kb in range(KB):
                slot = q % self.nrc
                yield from wait(self.rcf[slot], (q // self.nrc) & 1, q // self.nrc + 1, "row cache full")
                for _ in range(self._nq(nact)):
                    if Q % self.nbg == g:
                        stage, phase = Q & (self.na - 1), (Q >> self.lna) & 1
                        if self.rc[slot] != q:
                            raise ProtocolError("builder reads row cache of pass %s, wants %d" % (self.rc[slot], q))
                        yield                                               # rows -> registers
                        yield from wait(self.aempty[stage], phase ^ 1, Q >> self.lna, "operand slots free")
                        self.stage_rows[stage][w4] = Q
                        yield
                        self.afull[stage].arrive()
                    Q += 1
                self.rce[slot].arrive()
                q += 1
            self.rece[rb].arrive()

    def issuer(self, mi):
        Q = 0
        acc, aph = 0, 0
        for it, (nact, KB) in enumerate(self.tiles):
            rb = it % self.nrec
            yield from wait(self.recf[rb], (it // self.nrec) & 1, it // self.nrec + 1, "record ready (issuer)")
            if self.rec[rb] != it:
                raise ProtocolError("issuer reads record of tile %s, wants %d" % (self.rec[rb], it))
            yield from wait(self.acce[acc], aph ^ 1, it // self.nbuf, "accumulator free")
            for _ in range(self._nq(nact) * KB):
                if Q & (self.nmma - 1) == mi:                              # stage Q belongs to issuer Q mod nmma
                    sa = Q & (self.na - 1)
                    yield from wait(self.afull[sa], (Q >> self.lna) & 1, (Q >> self.lna) + 1, "stage full")
                    if (not self.resident and self.stage_w[sa] != Q) or any(r != Q for r in self.stage_rows[sa]):
                        raise ProtocolError("issuer %d, stage %d: holds rows %s weights %s"
                                            % (mi, Q, self.stage_rows[sa], self.stage_w[sa]))
                    if self.acc_tile[acc] not in (None, it):
                        raise ProtocolError("accumulator %d still holds tile %s" % (acc, self.acc_tile[acc]))
                    self.acc_tile[acc] = it
                    yield                                                   # MMAs execute
                    self.done_stages.append(Q)
                    self.aempty[sa].arrive()                                # tcgen05.commit
                Q += 1
            yield
            self.accf[acc].arrive()
            self.rece[rb].arrive()
            acc += 1
            if acc == self.nbuf:
                acc, aph = 0, aph ^ 1

    def epilogue(self, w):
        acc, aph = 0, 0
        for it in range(len(self.tiles)):
            yield from wait(self.accf[acc], aph, it // self.nbuf + 1, "accumulator full")
            if self.acc_tile[acc] not in (it, None):                        # None: no issuer had a stage in this tile yet
                raise ProtocolError("epilogue reads tile %s, wants %d" % (self.acc_tile[acc], it))
            yield
            if w == 0:
                self.done_tiles.append(it)
            self.acce[acc].arrive()
            if self.acce[acc].pending == self.acce[acc].count:              # last warp released the buffer
                self.acc_tile[acc] = None
            acc += 1
            if acc == self.nbuf:
                acc, aph = 0, aph ^ 1

    def run(self, seed=0, max_steps=10 ** 7):
        rng = random.Random(seed)
        roles = [self.record_producer(), self.list_producer()] + [self.weight_producer(w) for w in range(self.nwp)]
        roles += [self.gatherer(g) for g in range(self.G)]
        roles += [self.builder(b) for b in range(4 * self.nbg)]
        roles += [self.issuer(m) for m in range(self.nmma)]
        roles += [self.epilogue(w) for w in range(self.E)]
        alive = list(range(len(roles)))
        idle = 0
        for _ in range(max_steps):
            if not alive:
                break
            i = rng.choice(alive)
            before = self._state()
            try:
                next(roles[i])
            except StopIteration:
                alive.remove(i)
                idle = 0
                continue
            idle = idle + 1 if self._state() == before else 0
            if idle > 200 * len(roles):
                raise ProtocolError("deadlock: no role makes progress")
        else:
            raise ProtocolError("step budget exhausted")
        total = sum(self._nq(n) * k for n, k in self.tiles)
        if sorted(self.done_stages) != list(range(total)) or self.done_tiles != list(range(len(self.tiles))):
            raise ProtocolError("work lost: %d of %d stages, tiles %s" % (len(self.done_stages), total, self.done_tiles))
        return True

    def _state(self):
        bars = self.afull + self.aempty + self.rcf + self.rce + self.recf + self.rece + self.accf + self.acce + self.ulf + self.ule
        return tuple((b.phase, b.pending) for b in bars) + (len(self.done_stages),)


class CtaBarrier(object):
    """__syncthreads() / a named barrier among `count` warps: arrive, then wait for the generation to complete."""

    def __init__(self, count):
        self.count, self.pending, self.gen = count, count, 0

    def arrive(self):
        g = self.gen
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.gen = self.count, self.gen + 1
        return g

    def done(self, g):
        return self.gen > g


class SmallKernelModel(object):
    """Random-schedule runner shared by the two small tcgen05 kernels below: `roles` are generators, `engines` model the
    asynchronous units (TMA bulk copies, the tensor core) that arrive on an mbarrier some time after they were started."""

    def _run(self, roles, state, seed, max_steps=10 ** 6):
        rng = random.Random(seed)
        alive = list(range(len(roles)))
        idle = 0
        for _ in range(max_steps):
            if not alive:
                return True
            i = rng.choice(alive)
            before = state()
            try:
                next(roles[i])
            except StopIteration:
                alive.remove(i)
                idle = 0
                continue
            idle = idle + 1 if state() == before else 0
            if idle > 400 * len(roles):
                raise ProtocolError("deadlock: no role makes progress")
        raise ProtocolError("step budget exhausted")


class EccMessagesCta(SmallKernelModel):
    """ecc_messages_kernel (csrc/ecc_umma.cu): 4 warps; per (tile, quarter) every warp waits for the operand load (barL),
    warp 0 issues the MMAs, every warp waits for them (barM), thread 0 starts the NEXT load, all warps run the epilogue and
    meet at __syncthreads().  `fixed=False` is the protocol as first written: the next load is started as soon as warp 0
    has seen barM, although a slower warp may not have tested barL for the CURRENT quarter yet -- the load needs nobody's
    participation, so barL can run two phases ahead of that warp, whose parity test then never passes.  `fixed=True`:
    warps 1-3 arrive on a named barrier after their barL wait and warp 0 syncs on it before it starts the next load."""

    def __init__(self, quarters, fixed):
        self.Q, self.fixed = quarters, fixed
        self.barL, self.barM = Barrier(1), Barrier(1)
        self.cta, self.named = CtaBarrier(4), CtaBarrier(4)
        self.loads, self.mmas = [], []          # started, not yet completed
        self.seen = [0, 0, 0, 0]

    def engine(self, queue, bar):
        while True:
            if queue:
                yield                               # the unit takes a while
                queue.pop(0)
                bar.arrive()
            if self.finished:
                return
            yield

    def warp(self, w):
        phL = phM = 0
        for k in range(self.Q):
            yield from wait(self.barL, phL, k + 1, "warp %d operands of quarter %d" % (w, k))
            phL ^= 1
            self.seen[w] = k + 1
            g_named = self.named.arrive() if self.fixed and w != 0 else None
            if w == 0:
                self.mmas.append(k)
            yield
            yield from wait(self.barM, phM, k + 1, "warp %d accumulator of quarter %d" % (w, k))
            phM ^= 1
            if w == 0:
                if self.fixed:
                    g_named = self.named.arrive()
                    while not self.named.done(g_named):
                        yield
                if k + 1 < self.Q:
                    self.loads.append(k + 1)       # needs nobody's participation
            yield                                   # epilogue
            g = self.cta.arrive()
            while not self.cta.done(g):
                yield
        self.done_warps += 1
        if self.done_warps == 4:
            self.finished = True

    def run(self, seed=0):
        self.finished, self.done_warps = False, 0
        self.loads.append(0)
        roles = [self.warp(w) for w in range(4)] + [self.engine(self.loads, self.barL), self.engine(self.mmas, self.barM)]
        return self._run(roles, lambda: (self.barL.phase, self.barM.phase, self.cta.gen, self.named.gen, tuple(self.seen),
                                         len(self.loads), len(self.mmas), self.done_warps), seed)


class WgradCta(SmallKernelModel):
    """wgrad_umma_kernel (csrc/wgrad_umma.cu): 16 warps, two operand buffers; before group `it` is gathered into buffer
    it & 1 every warp waits for the commit of group it - 2 (parity ((it >> 1) - 1) & 1), then gathers, __syncthreads(),
    warp 0 issues the group's MMAs and commits to the buffer's barrier."""

    def __init__(self, groups, warps=4):
        self.G, self.W = groups, warps
        self.bar = [Barrier(1), Barrier(1)]
        self.cta = CtaBarrier(warps)
        self.mmas = [[], []]
        self.buf = [None, None]

    def engine(self, b):
        while True:
            if self.mmas[b]:
                yield
                it = self.mmas[b].pop(0)
                if self.buf[b] != it:
                    raise ProtocolError("MMAs of group %d read buffer %d holding group %s" % (it, b, self.buf[b]))
                self.bar[b].arrive()
            if self.finished:
                return
            yield

    def warp(self, w):
        for it in range(self.G):
            b = it & 1
            if it >= 2:
                yield from wait(self.bar[b], ((it >> 1) - 1) & 1, it >> 1, "warp %d buffer %d before group %d" % (w, b, it))
            self.buf[b] = it                        # gather: overwrite the operand buffer
            yield
            g = self.cta.arrive()
            while not self.cta.done(g):
                yield
            if w == 0:
                self.mmas[b].append(it)
        self.done_warps += 1
        if self.done_warps == self.W:
            self.finished = True

    def run(self, seed=0):
        self.finished, self.done_warps = False, 0
        roles = [self.warp(w) for w in range(self.W)] + [self.engine(0), self.engine(1)]
        return self._run(roles, lambda: (self.bar[0].phase, self.bar[1].phase, self.cta.gen, len(self.mmas[0]),
                                         len(self.mmas[1]), self.done_warps), seed)


def random_tiles(rng, n_tiles, max_units=27, max_kb=3):
    return [(rng.randint(1, max_units), rng.randint(1, max_kb)) for _ in range(n_tiles)]


if __name__ == "__main__":
    r = random.Random(1)
    for na, nrc, nbg, nmma, nbuf, res, us in ((8, 3, 2, 4, 2, True, 1), (8, 3, 2, 2, 2, False, 1), (4, 2, 2, 1, 2, False, 1),
                                              (8, 2, 2, 1, 1, False, 1), (4, 1, 2, 4, 2, False, 2), (2, 2, 2, 2, 2, False, 4)):
        for seed in range(20):
            Cta(random_tiles(r, 6), na, nrc, 2, nbg, nmma, nbuf=nbuf, resident=res, us=us).run(seed)
    for seed in range(100):
        EccMessagesCta(12, fixed=True).run(seed)
        WgradCta(14).run(seed)
    caught = 0
    for seed in range(100):
        try:
            EccMessagesCta(12, fixed=False).run(seed)
        except ProtocolError:
            caught += 1
    print("protocol ok; the first ecc_messages protocol deadlocks under %d of 100 random schedules" % caught)
    try:  # four issuers on a two-stage ring: an issuer's next stage is two generations ahead on the same buffer
        for seed in range(50):
            Cta(random_tiles(r, 6), 2, 2, 2, 2, 4).run(seed)
        print("(the issuers > stages configuration was not caught)")
    except ProtocolError as e:
        print("issuers > stages is caught:", e)

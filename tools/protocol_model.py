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


def random_tiles(rng, n_tiles, max_units=27, max_kb=3):
    return [(rng.randint(1, max_units), rng.randint(1, max_kb)) for _ in range(n_tiles)]


if __name__ == "__main__":
    r = random.Random(1)
    for na, nrc, nbg, nmma, nbuf, res, us in ((8, 3, 2, 4, 2, True, 1), (8, 3, 2, 2, 2, False, 1), (4, 2, 2, 1, 2, False, 1),
                                              (8, 2, 2, 1, 1, False, 1), (4, 1, 2, 4, 2, False, 2), (2, 2, 2, 2, 2, False, 4)):
        for seed in range(20):
            Cta(random_tiles(r, 6), na, nrc, 2, nbg, nmma, nbuf=nbuf, resident=res, us=us).run(seed)
    print("protocol ok")
    try:  # four issuers on a two-stage ring: an issuer's next stage is two generations ahead on the same buffer
        for seed in range(50):
            Cta(random_tiles(r, 6), 2, 2, 2, 2, 4).run(seed)
        print("(the issuers > stages configuration was not caught)")
    except ProtocolError as e:
        print("issuers > stages is caught:", e)

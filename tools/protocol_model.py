#!/usr/bin/env python
"""Executable model of the mbarrier protocol of conv_umma_kernel (3d-wsis_b200/csrc/conv_umma.cu).

The kernel is nine cooperating loops (record producer, weight producer, gather warps, builder warps, MMA issuers,
epilogue warps) that hand shared-memory buffers to each other through mbarriers whose waits only test a PARITY bit.
A parity wait is sound only while the waiter can never be two phases away from the phase it means; one version of
the kernel violated that (more builder warps than operand stages) and hung on the GPU.  This model transcribes the
control flow of every role -- the same ring indices, the same parities, the same arrival counts -- and runs the
roles under a random scheduler while checking, at every wait that passes, that the LOGICAL generation the waiter
needs has really completed, and at every buffer use that the buffer holds what the consumer expects.  A violation
or a deadlock raises.  tests/test_cpu.py runs it over every launch plan wsis_conv_umma_plan can produce.

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
    """One persistent CTA working through `tiles` = list of (number of active offsets, KB)."""

    def __init__(self, tiles, na, nrc, nrec, nb, nmma, builder_halves=2, gather_warps=4, epi_warps=4):
        assert na & (na - 1) == 0
        self.tiles, self.na, self.nrc, self.nrec, self.nb, self.nmma = tiles, na, nrc, nrec, nb, nmma
        self.halves, self.G, self.E = builder_halves, gather_warps, epi_warps
        self.lna = na.bit_length() - 1
        self.afull = [Barrier(builder_halves + 1) for _ in range(na)]      # builders + weight producer (expect_tx)
        self.aempty = [Barrier(1) for _ in range(na)]                      # tcgen05.commit
        self.rcf = [Barrier(gather_warps) for _ in range(nrc)]
        self.rce = [Barrier(builder_halves * nb) for _ in range(nrc)]
        self.recf = [Barrier(1) for _ in range(nrec)]
        self.rece = [Barrier(gather_warps + builder_halves * nb) for _ in range(nrec)]
        self.accf = [Barrier(nmma) for _ in range(2)]
        self.acce = [Barrier(epi_warps) for _ in range(2)]
        # buffer contents (what the consumer must find)
        self.stage_rows = [[None] * builder_halves for _ in range(na)]     # unit id written by each builder half
        self.stage_w = [None] * na                                         # unit id of the weight block
        self.rc = [None] * nrc                                             # pass id
        self.rec = [None] * nrec                                           # tile iteration
        self.acc_tile = [None, None]                                       # tile whose MMAs went into the buffer
        self.done_units, self.done_tiles = [], []

    # ---- roles (each a generator; `yield` = be descheduled) ----
    def record_producer(self):
        for it in range(len(self.tiles)):
            rb = it % self.nrec
            yield from wait(self.rece[rb], ((it // self.nrec) & 1) ^ 1, it // self.nrec, "record buffer free")
            self.rec[rb] = it
            yield
            self.recf[rb].arrive()                                         # expect_tx arrive + bytes landed

    def weight_producer(self):
        j = 0
        for nact, KB in self.tiles:
            for _ in range(KB * nact):
                s = j & (self.na - 1)
                yield from wait(self.aempty[s], ((j >> self.lna) & 1) ^ 1, j >> self.lna, "weight slot free")
                self.stage_w[s] = j
                yield
                self.afull[s].arrive()
                j += 1

    def gatherer(self, g):
        q = 0
        for it, (nact, KB) in enumerate(self.tiles):
            rb = it % self.nrec
            yield from wait(self.recf[rb], (it // self.nrec) & 1, it // self.nrec + 1, "record ready (gatherer)")
            if self.rec[rb] != it:
                raise ProtocolError("gatherer reads record of tile %s, wants %d" % (self.rec[rb], it))
            for _ in range(KB):
                slot = q % self.nrc
                yield                                                       # loads in flight
                yield from wait(self.rce[slot], ((q // self.nrc) & 1) ^ 1, q // self.nrc, "row cache free")
                if g == 0:
                    self.rc[slot] = q
                yield
                self.rcf[slot].arrive()
                q += 1
            self.rece[rb].arrive()

    def builder(self, bw):
        b, half = bw % self.nb, bw // self.nb
        if half >= self.halves:
            return
        q = j0 = 0
        for it, (nact, KB) in enumerate(self.tiles):
            rb = it % self.nrec
            yield from wait(self.recf[rb], (it // self.nrec) & 1, it // self.nrec + 1, "record ready (builder)")
            if self.rec[rb] != it:
                raise ProtocolError("builder reads record of tile %s, wants %d" % (self.rec[rb], it))
            for kb in range(KB):
                slot = q % self.nrc
                yield from wait(self.rcf[slot], (q // self.nrc) & 1, q // self.nrc + 1, "row cache full")
                jb = j0 + kb * nact
                ak = (b + self.nb - jb % self.nb) % self.nb
                while ak < nact:
                    j = jb + ak
                    stage, phase = j & (self.na - 1), (j >> self.lna) & 1
                    if self.rc[slot] != q:
                        raise ProtocolError("builder reads row cache of pass %s, wants %d" % (self.rc[slot], q))
                    yield                                                   # prefetch rows into registers
                    yield from wait(self.aempty[stage], phase ^ 1, j >> self.lna, "operand stage free")
                    self.stage_rows[stage][half] = j
                    yield
                    self.afull[stage].arrive()
                    ak += self.nb
                self.rce[slot].arrive()
                q += 1
            j0 += nact * KB
            self.rece[rb].arrive()

    def issuer(self, mi):
        j0 = 0
        for it, (nact, KB) in enumerate(self.tiles):
            n, acc = nact * KB, it & 1
            yield from wait(self.acce[acc], ((it >> 1) & 1) ^ 1, it >> 1, "accumulator free")
            u = (mi - j0) & (self.nmma - 1)
            while u < n:
                j = j0 + u
                sa = j & (self.na - 1)
                yield from wait(self.afull[sa], (j >> self.lna) & 1, (j >> self.lna) + 1, "stage full")
                if self.stage_w[sa] != j or any(r != j for r in self.stage_rows[sa]):
                    raise ProtocolError("issuer %d, unit %d: stage holds rows %s weights %s"
                                        % (mi, j, self.stage_rows[sa], self.stage_w[sa]))
                if self.acc_tile[acc] not in (None, it):
                    raise ProtocolError("accumulator %d still holds tile %s" % (acc, self.acc_tile[acc]))
                self.acc_tile[acc] = it
                yield                                                       # MMAs execute
                self.done_units.append(j)
                self.aempty[sa].arrive()                                    # tcgen05.commit
                u += self.nmma
            j0 += n
            yield
            self.accf[acc].arrive()

    def epilogue(self, w):
        for it in range(len(self.tiles)):
            acc = it & 1
            yield from wait(self.accf[acc], (it >> 1) & 1, (it >> 1) + 1, "accumulator full")
            if self.acc_tile[acc] != it:
                raise ProtocolError("epilogue reads tile %s, wants %d" % (self.acc_tile[acc], it))
            yield
            if w == 0:
                self.done_tiles.append(it)
            self.acce[acc].arrive()
            if self.acce[acc].pending == self.acce[acc].count:              # last warp released the buffer
                self.acc_tile[acc] = None

    def run(self, seed=0, max_steps=10 ** 7):
        rng = random.Random(seed)
        roles = [self.record_producer(), self.weight_producer()]
        roles += [self.gatherer(g) for g in range(self.G)]
        roles += [self.builder(b) for b in range(self.halves * self.nb)]
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
        total = sum(n * k for n, k in self.tiles)
        if sorted(self.done_units) != list(range(total)) or self.done_tiles != list(range(len(self.tiles))):
            raise ProtocolError("work lost: %d of %d units, tiles %s" % (len(self.done_units), total, self.done_tiles))
        return True

    def _state(self):
        bars = self.afull + self.aempty + self.rcf + self.rce + self.recf + self.rece + self.accf + self.acce
        return tuple((b.phase, b.pending) for b in bars) + (len(self.done_units),)


def random_tiles(rng, n_tiles, max_units=27, max_kb=3):
    return [(rng.randint(1, max_units), rng.randint(1, max_kb)) for _ in range(n_tiles)]


if __name__ == "__main__":
    r = random.Random(1)
    for na, nrc, nb, nmma in ((4, 3, 4, 2), (8, 3, 4, 2), (2, 2, 2, 2), (1, 1, 1, 1), (4, 2, 4, 1)):
        for seed in range(20):
            Cta(random_tiles(r, 6), na, nrc, 2, nb, nmma).run(seed)
    print("protocol ok")
    try:  # the configuration that hung on the GPU: 6 single-warp builders on 4 stages
        for seed in range(50):
            Cta(random_tiles(r, 6), 4, 3, 2, 6, 2, builder_halves=1).run(seed)
        print("(the nb > na configuration was not caught)")
    except ProtocolError as e:
        print("nb > na is caught:", e)

"""Randomised discrete-event model of the barrier protocol of the channel-streamed persistent kernel
(segmantic_b200/csrc/conv_cs.cu).

Roles as coroutines -- the A producer and the W producer (one TMA ring each), `ni` MMA issuer threads that own disjoint
row tiles of a unit (an issuer without tiles still waits and commits), `ne` epilogue warps, the asynchronous engines (TMA
loads complete at random times; the tensor pipe executes every issuer's MMAs and commits in that issuer's order, issuers
interleaved at random) -- and mbarriers with the hardware's parity semantics (`try_wait.parity P` succeeds iff the
current phase's parity differs from P, so a waiter two phases behind, or one early, aliases).  Random scheduling with
random per-role speeds.  Detects deadlocks and these hazards:

* an MMA EXECUTES while its A stage or W stage does not hold the fill it was issued for (not landed yet, or already
  overwritten by a later fill: the producer passed an "empty" barrier before the stage's readers were done);
* an MMA of unit u writes a TMEM half that the epilogue of the unit that used it before has not drained;
* an epilogue warp reads a TMEM half before every MMA of its unit has executed.

`python tests/sim_cs_protocol.py [seeds]`; tests/test_cs_protocol_sim.py runs a few seeds on the CPU."""
import random
import sys


class MBar:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"over-arrival on {self.name}"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def try_wait(self, parity):
        return (self.phase & 1) != parity   # the phase with this parity has completed


def run(units, nkc, ngw, astages, wstages, nbuf, ntiles, seed, ni=4, ne=8, bug=None):
    """`bug`: None, or a deliberately broken protocol the model must catch ("no_aempty": the A producer does not wait
    for its ring stage to be released; "no_tempty": the issuers do not wait for the epilogue to drain a TMEM half)."""
    rnd = random.Random(seed)
    AFULL = [MBar(1, f"AFULL{s}") for s in range(astages)]
    AEMPTY = [MBar(ni, f"AEMPTY{s}") for s in range(astages)]
    WFULL = [MBar(1, f"WFULL{s}") for s in range(wstages)]
    WEMPTY = [MBar(ni, f"WEMPTY{s}") for s in range(wstages)]
    TFULL = [MBar(ni, f"TFULL{b}") for b in range(2)]
    TEMPTY = [MBar(ne, f"TEMPTY{b}") for b in range(2)]
    a_content = [None] * astages     # fill id the stage holds (None: nothing / being overwritten)
    w_content = [None] * wstages
    tma_queue = []                   # pending loads: (kind, stage, fill id, barrier)
    mma_queues = [[] for _ in range(ni)]
    executed = set()                 # (unit, kc, gi, issuer) groups of MMAs that have run
    tmem_owner = [None, None]        # unit whose accumulators the half holds (None after the epilogue drained it)
    drained = set()
    errors = []

    def a_producer():
        s, ph, wrapped, fill = 0, 0, False, 0
        for u in range(units):
            for kc in range(nkc):
                if wrapped and bug != "no_aempty":
                    while not AEMPTY[s].try_wait(ph ^ 1):
                        yield ("wait", f"A producer AEMPTY{s} fill {fill}")
                a_content[s] = None                       # the TMA engine may start overwriting at once
                tma_queue.append(("A", s, fill, AFULL[s]))
                fill += 1
                s += 1
                if s == astages:
                    s, ph, wrapped = 0, ph ^ 1, True
                yield None

    def w_producer():
        s, ph, wrapped, fill = 0, 0, False, 0
        for u in range(units):
            for kc in range(nkc):
                for gi in range(ngw):
                    if wrapped:
                        while not WEMPTY[s].try_wait(ph ^ 1):
                            yield ("wait", f"W producer WEMPTY{s} fill {fill}")
                    w_content[s] = None
                    tma_queue.append(("W", s, fill, WFULL[s]))
                    fill += 1
                    s += 1
                    if s == wstages:
                        s, ph, wrapped = 0, ph ^ 1, True
                    yield None

    def issuer(iw):
        q = mma_queues[iw]
        sa = sw = 0
        pa = pw = 0
        afill = wfill = 0
        has_tiles = any(tt % ni == iw for tt in range(ntiles))
        for u in range(units):
            buf = (u & 1) if nbuf == 2 else 0
            use = (u >> 1) if nbuf == 2 else u
            if use > 0 and bug != "no_tempty":
                while not TEMPTY[buf].try_wait((use - 1) & 1):
                    yield ("wait", f"issuer{iw} TEMPTY{buf} unit {u}")
            for kc in range(nkc):
                while not AFULL[sa].try_wait(pa):
                    yield ("wait", f"issuer{iw} AFULL{sa} unit {u} kc {kc}")
                for gi in range(ngw):
                    while not WFULL[sw].try_wait(pw):
                        yield ("wait", f"issuer{iw} WFULL{sw} unit {u} kc {kc} gi {gi}")
                    if has_tiles:
                        q.append(("mma", (u, kc, gi, iw, sa, afill, sw, wfill, buf)))
                    q.append(("commit", WEMPTY[sw]))
                    wfill += 1
                    sw += 1
                    if sw == wstages:
                        sw, pw = 0, pw ^ 1
                    yield None
                q.append(("commit", AEMPTY[sa]))
                afill += 1
                sa += 1
                if sa == astages:
                    sa, pa = 0, pa ^ 1
            q.append(("commit", TFULL[buf]))
            yield None

    def epilogue(e):
        for u in range(units):
            buf = (u & 1) if nbuf == 2 else 0
            use = (u >> 1) if nbuf == 2 else u
            while not TFULL[buf].try_wait(use & 1):
                yield ("wait", f"epilogue{e} TFULL{buf} unit {u}")
            for kc in range(nkc):
                for gi in range(ngw):
                    for iw in range(ni):
                        if any(tt % ni == iw for tt in range(ntiles)) and (u, kc, gi, iw) not in executed:
                            errors.append(f"epilogue{e} read unit {u} before MMA group ({kc},{gi}) of issuer {iw} ran")
            if tmem_owner[buf] != u and ntiles > 0:
                errors.append(f"epilogue{e} found unit {tmem_owner[buf]} in TMEM half {buf}, expected {u}")
            for _ in range(rnd.randrange(0, 4 * nkc * ngw)):   # a slow epilogue (stores, residual loads): many scheduler turns
                yield None
            drained.add((u, e))
            TEMPTY[buf].arrive()
            yield None

    def engines():
        while True:
            did = False
            qs = [q for q in mma_queues if q]
            if qs and rnd.random() < 0.7:
                kind, x = rnd.choice(qs).pop(0)
                if kind == "mma":
                    u, kc, gi, iw, sa, afill, sw, wfill, buf = x
                    if a_content[sa] != afill:
                        errors.append(f"MMA of unit {u} kc {kc} ran with A stage {sa} holding {a_content[sa]}, expected fill {afill}")
                    if w_content[sw] != wfill:
                        errors.append(f"MMA of unit {u} kc {kc} gi {gi} ran with W stage {sw} holding {w_content[sw]}, expected fill {wfill}")
                    prev = tmem_owner[buf]
                    if prev is not None and prev != u and any((prev, e) not in drained for e in range(ne)):
                        errors.append(f"MMA of unit {u} wrote TMEM half {buf} before unit {prev} was drained")
                    tmem_owner[buf] = u
                    executed.add((u, kc, gi, iw))
                else:
                    x.arrive()
                did = True
            if tma_queue and rnd.random() < 0.5:
                kind, s, fill, bar = tma_queue.pop(rnd.randrange(min(3, len(tma_queue))))
                if kind == "A":
                    a_content[s] = fill
                else:
                    w_content[s] = fill
                bar.arrive()
                did = True
            yield None if did else ("idle", "engines")

    procs = {"aprod": a_producer(), "wprod": w_producer(), "engines": engines()}
    for iw in range(ni):
        procs[f"iss{iw}"] = issuer(iw)
    for e in range(ne):
        procs[f"epi{e}"] = epilogue(e)
    weights = {k: rnd.choice([0.2, 1.0, 5.0]) for k in procs}
    alive = set(procs) - {"engines"}
    status = {}
    steps = 0
    while alive:
        steps += 1
        names = list(procs)
        k = rnd.choices(names, [weights[n] for n in names])[0]
        try:
            r = next(procs[k])
        except StopIteration:
            alive.discard(k)
            del procs[k]
            continue
        status[k] = r
        if errors:
            return "HAZARD", errors[:5], None
        if steps % 2000 == 0 and not any(mma_queues) and not tma_queue and \
                all(isinstance(status.get(n), tuple) for n in procs):
            stuck = True
            for n in list(procs):
                try:
                    r = next(procs[n])
                except StopIteration:
                    alive.discard(n)
                    del procs[n]
                    stuck = False
                    continue
                status[n] = r
                if not isinstance(r, tuple):
                    stuck = False
            if stuck and not any(mma_queues) and not tma_queue:
                return "DEADLOCK", {n: status[n][1] for n in procs if n != "engines"}, errors
    return "OK", steps, errors


CONFIGS = (  # (units per CTA, K chunks, weight groups per chunk, A stages, W stages, TMEM halves, row tiles per unit)
    (4, 8, 3, 4, 4, 2, 3),     # bottom.unit0: 128 channels, 27 taps in 3 groups, 3 tiles
    (3, 16, 7, 4, 4, 1, 3),    # N = 128: 7 groups of 4 taps, the unit needs the whole TMEM (no double buffering)
    (5, 1, 3, 2, 4, 2, 2),     # strided 16-channel block: a single K chunk per unit
    (6, 2, 2, 2, 4, 2, 1),     # one row tile: three issuers without tiles
    (3, 4, 2, 3, 4, 2, 8),     # two windows per unit: 8 tiles, two per issuer
    (1, 24, 2, 4, 4, 1, 3),    # transposed conv: 24 chunks of 8 shifts
)

if __name__ == "__main__":
    bad = 0
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
        for cfg in CONFIGS:
            r = run(*cfg, seed)
            if r[0] != "OK":
                bad += 1
                print(seed, cfg, r[0], r[1])
    print("bad", bad)

"""Randomised discrete-event model of the barrier protocol of the row-sweep head kernel (segmantic_b200/csrc/conv_rs.cu).

Roles as coroutines (TMA producer, MMA issuer, 4 x nq epilogue warps, the in-order tensor pipe / TMA engine), mbarriers
with the hardware's parity semantics (`try_wait.parity P` succeeds iff the current phase's parity differs from P -- so a
waiter two phases behind, or one phase early, aliases), random scheduling with random per-role speeds.  Detects
deadlocks and three hazards: the issuer passing a "row cleared" barrier before every lane quarter cleared the row, an
epilogue warp passing a "row multiplied" barrier before that MMA completed, anyone passing "plane landed" early.
`python tests/sim_rs_protocol.py [seeds]`; tests/test_rs_protocol_sim.py runs a few seeds on the CPU."""
import random, sys

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
        return (self.phase & 1) != parity   # phase with this parity has completed

def run(D0, t1, R, nq, units, seed, res=True, verbose=False, RC=4, ni=1):
    rnd = random.Random(seed)
    H1 = t1 + 2   # RC: input rows per tcgen05.commit (one 'chunk multiplied' barrier per chunk and plane)
    PFULL = [MBar(1, f"PFULL{s}") for s in range(R)]
    PEMPTY = [MBar(ni + (4 * nq if res else 0), f"PEMPTY{s}") for s in range(R)]   # every issuer warp commits
    FULL = [MBar(ni, f"FULL{c}") for c in range((H1 + 3) // 4)]   # per chunk of four input rows, both issuers commit
    CLR = [MBar(nq * max(min(4, t1 - 4 * c), 1), f"CLR{c}") for c in range((t1 + 3) // 4)]   # per chunk of four output rows
    mma_queues = [[], []]   # per issuer thread, in order: ('mma', tag) / ('commit', bar); the pipe interleaves them
    tma_queue = []   # pending loads: bar
    # ground truth tracking for hazards
    state = dict(mma_done=set(), cleared={}, true_phase={})
    errors = []
    seen = {}

    def producer():
        G = 0
        for kl in range(units):
            for x0 in range(D0):
                slot = G % R
                if G >= R:
                    par = (G // R - 1) & 1
                    while not PEMPTY[slot].try_wait(par): yield ("wait", f"prod PEMPTY{slot} G={G}")
                tma_queue.append((PFULL[slot], G))
                G += 1
                yield None

    def issuer(iw):
        mma_queue = mma_queues[iw]
        G = 0
        for kl in range(units):
            for xi in range(D0 + 2):
                V0 = kl * (D0 + 2) + xi
                Vc = V0 - 1
                real = xi < D0
                if real:
                    slot = G % R
                    while not PFULL[slot].try_wait((G // R) & 1): yield ("wait", f"iss{iw} PFULL{slot} G={G}")
                    if ("loaded", G) not in state["mma_done"]: errors.append(f"issuer passed PFULL before plane {G} loaded")
                next_commit = 0
                for i in range(iw, H1, ni):
                    c = i >> 2
                    if Vc >= 0 and (i & 3) == iw and 4 * c < t1:
                        while not CLR[c].try_wait(Vc & 1): yield ("wait", f"iss{iw} CLR{c} Vc={Vc}")
                        for o in range(4 * c, min(4 * c + 4, t1)):
                            if state["cleared"].get((Vc, o), 0) != nq: errors.append(f"issuer{iw} passed CLR[{c}] Vc={Vc}: row {o} has {state['cleared'].get((Vc,o),0)} clears")
                        seen.setdefault((iw, Vc), set()).update(range(4 * c, min(4 * c + 4, t1)))
                    if Vc >= 0:
                        for o in range(max(i - 2, 0), min(i, t1 - 1) + 1):
                            if o not in seen.get((iw, Vc), set()): errors.append(f"issuer{iw} row {i} touches unobserved row {o}")
                    if real:
                        mma_queue.append(("mma", (G, i)))
                    if (i & 3) == 4 - ni + iw or i + ni >= H1:
                        mma_queue.append(("commit", FULL[c]))
                        next_commit = c + 1
                    yield None
                for c in range(next_commit, ((H1 - 1) >> 2) + 1):
                    mma_queue.append(("commit", FULL[c]))
                if real:
                    mma_queue.append(("commit", PEMPTY[slot]))
                    G += 1
                yield None

    def epi(g, q):
        for kl in range(units):
            for v in range(-1, D0 + 1):
                V = kl * (D0 + 2) + v + 1
                xin = min(max(v + 1, 0), D0 - 1)
                Gin = kl * D0 + xin
                Gv = kl * (D0 + 2) + v + 1
                real = 0 <= v < D0
                if real and res:
                    Gres = kl * D0 + v
                    rslot = Gres % R
                    while not PFULL[rslot].try_wait((Gres // R) & 1): yield ("wait", f"epi{g}{q} PFULL{rslot} Gres={Gres}")
                    if ("loaded", Gres) not in state["mma_done"]: errors.append(f"epi passed PFULL before plane {Gres} loaded")
                for oa in range(2 * g, t1, 8):       # two adjacent output rows per step
                    rows = [oa] + ([oa + 1] if oa + 1 < t1 else [])
                    ilast = (rows[-1] + 2) >> 2
                    while not FULL[ilast].try_wait(Gv & 1): yield ("wait", f"epi{g}{q} FULL{ilast} Gv={Gv} v={v}")
                    for o in rows:
                        for i in range(o, o + 3):
                            if (Gin, i) not in state["mma_done"]: errors.append(f"epi{g}{q} passed FULL[{ilast}] before MMA({Gin},{i}) done (v={v}, o={o})")
                    yield None
                    for o in rows:
                        state["cleared"][(V, o)] = state["cleared"].get((V, o), 0) + 1
                        CLR[o >> 2].arrive()
                    yield None
                if real and res:
                    PEMPTY[rslot].arrive()
                yield None

    def pipe():   # tensor pipe + TMA engine: complete queued work in order, at random times
        while True:
            did = False
            qs = [q for q in mma_queues if q]
            if qs and rnd.random() < 0.7:
                kind, x = rnd.choice(qs).pop(0)
                if kind == "mma": state["mma_done"].add(x)
                else: x.arrive()
                did = True
            if tma_queue and rnd.random() < 0.5:
                bar, G = tma_queue.pop(rnd.randrange(min(2, len(tma_queue))))
                state["mma_done"].add(("loaded", G))
                bar.arrive()
                did = True
            yield None if did else ("idle", "pipe")

    procs = {"prod": producer(), "pipe": pipe()}
    for iw in range(ni):
        procs[f"iss{iw}"] = issuer(iw)
    for g in range(4):
        for q in range(nq):
            procs[f"epi{g}{q}"] = epi(g, q)
    weights = {k: rnd.choice([0.2, 1.0, 5.0]) for k in procs}
    alive = set(procs) - {"pipe"}
    status = {}
    steps = 0
    while alive:
        steps += 1
        names = list(procs)
        k = rnd.choices(names, [weights[n] for n in names])[0]
        try:
            r = next(procs[k])
        except StopIteration:
            alive.discard(k); del procs[k]; continue
        status[k] = r
        if steps % 2000 == 0:
            # deadlock check: everyone alive is waiting and pipe idle and queues empty
            if not any(mma_queues) and not tma_queue and all(isinstance(status.get(n), tuple) for n in procs):
                # run a sweep: poll everyone once more to confirm
                stuck = True
                for n in list(procs):
                    try: r = next(procs[n])
                    except StopIteration: alive.discard(n); del procs[n]; stuck = False; continue
                    status[n] = r
                    if not isinstance(r, tuple): stuck = False
                if stuck and not any(mma_queues) and not tma_queue:
                    return "DEADLOCK", {n: status[n][1] for n in procs if n != "pipe"}, errors
        if errors: return "HAZARD", errors[:5], None
    return "OK", steps, errors

if __name__ == "__main__":
    bad = 0
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
        for (D0, t1, R, nq, units) in ((6, 3, 4, 3, 1), (6, 3, 4, 3, 2), (8, 9, 4, 3, 3), (5, 6, 3, 4, 2), (7, 12, 4, 3, 2)):
            r = run(D0, t1, R, nq, units, seed, ni=1 + seed % 2)
            if r[0] != "OK":
                bad += 1
                print(seed, (D0, t1, R, nq, units), r[0], r[1])
    print("bad", bad)

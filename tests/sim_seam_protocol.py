"""Randomised model of the NVLink seam exchange of the multi-GPU driver (segmantic_b200/seg/p2p.py, csrc/p2p.cu).

Every rank has two in-order streams -- compute (tail windows, rest windows, wait for DATA, blend, ACK) and copy (wait for
the tail event, wait for the peer's ACK of the previous volume, push, raise DATA, record "push done") -- whose operations
are queued by a host that never blocks and executed at random times; the counters live in the peer's memory.  Buffers
carry the volume number they hold.  Checked over many volumes in a row, for any number of ranks:

* no deadlock;
* a blend of volume k reads a receive buffer that holds volume k (pushed completely, not yet overwritten by k + 1);
* a push of volume k copies a send region that holds volume k (the next volume's tail windows have not started
  overwriting it) and lands only after the receiver has blended volume k - 1.

`bug`: a deliberately broken variant the model must catch ("no_ack": pushes do not wait for the receiver's ACK;
"no_begin": the next volume's tail windows do not wait for the previous push to leave the send region)."""
import random


def run(world, volumes, seed, bug=None):
    rnd = random.Random(seed)
    data = [0] * world          # DATA counter in rank r's memory (raised by r - 1)
    ack = [0] * world           # ACK counter in rank r's memory (raised by r + 1)
    recv = [0] * world          # volume number held by rank r's receive buffer (-1 while a push is landing)
    send = [0] * world          # volume number held by rank r's send region (-1 while being recomputed)
    blended = [0] * world       # last volume rank r has blended
    tail_done = [0] * world     # events (monotonic volume numbers)
    put_done = [0] * world
    errors = []
    # queues of (stream op, volume); "compute" ops: begin, tail_start, tail_end, rest, wait_data, blend, ack
    comp = [[] for _ in range(world)]
    copy = [[] for _ in range(world)]
    for r in range(world):
        for k in range(1, volumes + 1):
            comp[r] += [("begin", k), ("tail_start", k), ("tail_end", k), ("rest", k)]
            if r > 0:
                comp[r] += [("wait_data", k)]
            comp[r] += [("blend", k)]
            if r > 0:
                comp[r] += [("ack", k)]
            if r + 1 < world:
                copy[r] += [("wait_tail", k), ("wait_ack", k), ("put_start", k), ("put_end", k), ("signal", k)]

    def ready(r, stream, op, k):
        if op == "begin":
            return bug == "no_begin" or put_done[r] >= k - 1 or r + 1 >= world
        if op == "wait_data":
            return data[r] >= k
        if op == "wait_tail":
            return tail_done[r] >= k
        if op == "wait_ack":
            return bug == "no_ack" or ack[r] >= k - 1
        return True

    def execute(r, op, k):
        if op == "tail_start":
            if r + 1 < world and put_done[r] < k - 1:
                errors.append(f"rank {r}: tail windows of volume {k} overwrite the send region while push {k - 1} is in flight")
            send[r] = -1
        elif op == "tail_end":
            send[r] = k
            tail_done[r] = k
        elif op == "blend":
            if r > 0 and recv[r] != k:
                errors.append(f"rank {r}: blend of volume {k} reads a receive buffer holding {recv[r]}")
            blended[r] = k
        elif op == "ack":
            ack[r - 1] = k
        elif op == "put_start":
            if send[r] != k:
                errors.append(f"rank {r}: push of volume {k} copies a send region holding {send[r]}")
            if blended[r + 1] < k - 1:
                errors.append(f"rank {r}: push of volume {k} lands before rank {r + 1} blended volume {k - 1}")
            recv[r + 1] = -1
        elif op == "put_end":
            if send[r] != k:
                errors.append(f"rank {r}: the send region changed to {send[r]} during the push of volume {k}")
            recv[r + 1] = k
        elif op == "signal":
            data[r + 1] = k
            put_done[r] = k

    idle = 0
    while any(comp[r] or copy[r] for r in range(world)):
        r = rnd.randrange(world)
        q = comp[r] if rnd.random() < 0.5 else copy[r]
        if not q or not ready(r, q, *q[0]):
            idle += 1
            if idle > 20000:
                heads = {(rr, "comp" if qq is comp[rr] else "copy"): qq[0] for rr in range(world)
                         for qq in (comp[rr], copy[rr]) if qq}
                return "DEADLOCK", heads
            continue
        idle = 0
        op, k = q.pop(0)
        execute(r, op, k)
        if errors:
            return "HAZARD", errors[:3]
    return "OK", None

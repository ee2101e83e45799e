"""CPU model of the peer-window protocol of the row-sharded evaluation (csrc/api.cu: self_sim_sharded_sym, cov_sharded_*,
peer_allreduce; csrc/kernels.cuh: peer_allreduce_kernel).  No GPU: ranks are replayed as programs over shared "windows" under
random interleavings, and every access is checked against the version it must see.

Per evaluation e a rank executes, in stream order (the order of the launches in eval_impl):
    scatter   write its partial covariance tiles into the owners' windows            (EpiGramScatter, peer stores)
    copy      write its sign blocks into the windows of the ranks that own their rows  (cudaMemcpy2DAsync; awaited before A)
    A         one-shot allreduce #1: push payload into its slot of every window (parity of the epoch), publish flag = epoch,
              wait until every sender's flag in its own window has reached the epoch, reduce the slots
    owner     read the partial tiles it owns (all senders), write sign tiles into EVERY rank's Sg
    stage2c   read the sign blocks it received
    B         one-shot allreduce #2 (the other slot parity)
    covbwd    read its Sg
There is no other synchronisation.  Claims checked for every interleaving:
  * read-after-write: a rank reads tiles / blocks / slots / Sg only after every writer of THIS evaluation has written them;
  * write-after-read: nobody overwrites a location with evaluation e + 1 data before its reader has consumed evaluation e;
  * the two slot parities suffice (a sender is never more than one collective ahead of the slowest rank).
Two broken variants (the first collective removed; a single slot parity) must be caught."""
import random

import pytest


class Violation(Exception):
    pass


def simulate(world, evals, seed, skip_barrier_a=False, single_parity=False):
    rng = random.Random(seed)
    # version of the data each location holds: -1 = nothing yet
    gram = [[-1] * world for _ in range(world)]        # gram[owner][sender]
    pblk = [[-1] * world for _ in range(world)]        # pblk[receiver][sender]
    sg = [[-1] * world for _ in range(world)]          # sg[reader][owner]
    slots = [[[-1] * world for _ in range(world)] for _ in range(2)]   # slots[parity][receiver][sender] = epoch
    flags = [[0] * world for _ in range(world)]        # flags[receiver][sender] = last epoch published
    consumed = {"gram": [-1] * world, "pblk": [-1] * world, "sg": [-1] * world, "slot": [[0] * world for _ in range(2)]}
    ops = ["scatter", "copy", "A_push", "A_wait", "owner", "stage2c", "B_push", "B_wait", "covbwd"]
    pc = [0] * world                                    # index into the unrolled program of each rank
    total = evals * len(ops)

    def write(table, row, col, version, last_consumed, what):
        # the previous version must have been consumed by its reader before it is overwritten
        if table[row][col] >= 0 and last_consumed < table[row][col]:
            raise Violation(f"{what}: version {table[row][col]} at [{row}][{col}] overwritten by {version} before it was read")
        table[row][col] = version

    while any(p < total for p in pc):
        ready = []
        for r in range(world):
            if pc[r] >= total:
                continue
            e, op = divmod(pc[r], len(ops))
            name = ops[op]
            if name in ("A_wait", "B_wait"):
                epoch = 2 * e + (1 if name == "A_wait" else 2)
                if skip_barrier_a and name == "A_wait":
                    ready.append(r)
                elif all(flags[r][s] >= epoch for s in range(world)):
                    ready.append(r)
            else:
                ready.append(r)
        if not ready:
            raise Violation("deadlock")
        r = rng.choice(ready)
        e, op = divmod(pc[r], len(ops))
        name = ops[op]
        if name == "scatter":
            for owner in range(world):
                write(gram, owner, r, e, consumed["gram"][owner], "partial tiles")
        elif name == "copy":
            for recv in range(world):
                if recv != r:
                    write(pblk, recv, r, e, consumed["pblk"][recv], "sign blocks")
        elif name in ("A_push", "B_push"):
            epoch = 2 * e + (1 if name == "A_push" else 2)
            par = 0 if single_parity else epoch & 1
            for recv in range(world):
                write(slots[par], recv, r, epoch, consumed["slot"][par][recv], "allreduce slot")
            for recv in range(world):
                flags[recv][r] = epoch
        elif name in ("A_wait", "B_wait"):
            epoch = 2 * e + (1 if name == "A_wait" else 2)
            par = 0 if single_parity else epoch & 1
            if not (skip_barrier_a and name == "A_wait"):
                for s in range(world):
                    if slots[par][r][s] != epoch:
                        raise Violation(f"allreduce {epoch}: slot of sender {s} at rank {r} holds epoch {slots[par][r][s]}")
                consumed["slot"][par][r] = epoch
        elif name == "owner":
            for s in range(world):
                if gram[r][s] != e:
                    raise Violation(f"owner step of eval {e} at rank {r}: tiles of sender {s} are version {gram[r][s]}")
            consumed["gram"][r] = e
            for reader in range(world):
                write(sg, reader, r, e, consumed["sg"][reader], "sign matrix")
        elif name == "stage2c":
            for s in range(world):
                if s != r and pblk[r][s] != e:
                    raise Violation(f"stage 2c of eval {e} at rank {r}: block of sender {s} is version {pblk[r][s]}")
            consumed["pblk"][r] = e
        elif name == "covbwd":
            for owner in range(world):
                if sg[r][owner] != e:
                    raise Violation(f"covariance backward of eval {e} at rank {r}: Sg tiles of owner {owner} are version {sg[r][owner]}")
            consumed["sg"][r] = e
        pc[r] += 1


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_protocol_holds_under_random_interleavings(world):
    for seed in range(200):
        simulate(world, evals=4, seed=seed)


def test_missing_first_collective_is_caught():
    caught = 0
    for seed in range(50):
        try:
            simulate(4, evals=3, seed=seed, skip_barrier_a=True)
        except Violation:
            caught += 1
    assert caught >= 45          # without the barrier the owner step reads tiles that have not arrived


def test_single_slot_parity_is_caught():
    caught = 0
    for seed in range(200):
        try:
            simulate(4, evals=3, seed=seed, single_parity=True)
        except Violation:
            caught += 1
    assert caught >= 1           # a fast rank pushes collective e + 1 into a slot the slow rank has not reduced yet

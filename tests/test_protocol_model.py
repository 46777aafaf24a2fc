"""Model check (CPU, randomized schedules) of the two-buffer halo protocol of csrc/halo.cu + csrc/ops.cu.

Every rank repeats:  push(x): store my values into each receiver's buffer x & 1, then raise my flag = x at EVERY peer;
wait(x): until all peers' flags >= x;  read(x): gather from my buffer x & 1.  Ranks progress independently (no other
synchronisation, as in back-to-back cmb_op_apply calls).  The claim in DESIGN.md §7: a peer can only start exchange
x+2 after it consumed my flag of x+1, which I raised after my read of x — so two buffers suffice.  The model also shows
why the flags go to every peer and not only to the receivers of data: with a one-directional pattern a rank that
receives nothing would otherwise never wait and could overwrite a buffer that is still being read."""
import random

import pytest


def simulate(P, sends, nexch, flag_everyone, seed, steps_per_read=3):
    """sends[r] = set of ranks that r stores data into.  Returns None or a description of the first violation."""
    rng = random.Random(seed)
    buf = [[{s: 0 for s in range(P) if r in sends[s]} for _ in range(2)] for r in range(P)]  # buf[r][parity][sender]
    flag = [[0] * P for _ in range(P)]  # flag[r][s]: raised at r by s
    # per-rank program counter: (exchange, phase, reads_left); phases: 0 push, 1 wait, 2 read
    pc = [[1, 0, 0] for _ in range(P)]
    receivers_of = sends
    waits_on = [[s for s in range(P) if s != r and (flag_everyone or r in sends[s])] for r in range(P)]
    done = 0
    while done < P:
        r = rng.randrange(P)
        x, ph, left = pc[r]
        if x > nexch:
            continue
        if ph == 0:
            for q in receivers_of[r]:
                buf[q][x & 1][r] = x
            for q in range(P):
                if q != r and (flag_everyone or q in receivers_of[r]):
                    flag[q][r] = x
            pc[r] = [x, 1, 0]
        elif ph == 1:
            if all(flag[r][s] >= x for s in waits_on[r]):
                pc[r] = [x, 2, steps_per_read]
        else:
            for s, tag in buf[r][x & 1].items():  # the read takes several scheduler steps: re-checked every time
                if tag != x:
                    return "rank %d reading exchange %d found data of exchange %d from rank %d" % (r, x, tag, s)
            if left > 1:
                pc[r] = [x, 2, left - 1]
            else:
                pc[r] = [x + 1, 0, 0]
                if x + 1 > nexch:
                    done += 1
    return None


@pytest.mark.parametrize("P", [2, 3, 4, 8])
def test_two_buffers_suffice_when_everyone_is_flagged(P):
    patterns = {
        "ring": [{(r + 1) % P, (r - 1) % P} - {r} for r in range(P)],
        "one_directional": [{r - 1} if r > 0 else set() for r in range(P)],   # rank q reads from q+1 only
        "all_to_all": [set(range(P)) - {r} for r in range(P)],
        "star": [set(range(1, P)) if r == 0 else set() for r in range(P)],      # only rank 0 sends
    }
    for name, sends in patterns.items():
        for seed in range(40):
            assert simulate(P, sends, nexch=12, flag_everyone=True, seed=seed) is None, (name, seed)


def test_flagging_only_the_receivers_is_not_enough():
    # rank 1 stores into rank 0 and receives nothing: without flags from rank 0 it never waits, runs two exchanges
    # ahead and overwrites the buffer rank 0 is still reading
    sends = [set(), {0}]
    found = [simulate(2, sends, nexch=12, flag_everyone=False, seed=s) for s in range(200)]
    assert any(f is not None for f in found)
    assert all(simulate(2, sends, nexch=12, flag_everyone=True, seed=s) is None for s in range(200))


def simulate_mailbox(P, nred, ring, seed):
    """Gram-Schmidt coefficient mailboxes (csrc/device_utils.cuh mail_*): reduction k = every rank stores its partial
    into slot k % ring of every mailbox and raises flag k there; a rank sums the P partials of slot k % ring once all
    flags are >= k, and only then goes on to push reduction k+1 (stream order of the kernel chain)."""
    rng = random.Random(seed)
    slot = [[[0] * P for _ in range(ring)] for _ in range(P)]  # slot[r][k % ring][sender] = tag
    flag = [[0] * P for _ in range(P)]
    pc = [[1, 0] for _ in range(P)]  # (reduction, phase): 0 push, 1 pull
    done = 0
    while done < P:
        r = rng.randrange(P)
        k, ph = pc[r]
        if k > nred:
            continue
        if ph == 0:
            for q in range(P):
                slot[q][k % ring][r] = k
                flag[q][r] = k
            pc[r] = [k, 1]
        elif all(flag[r][s] >= k for s in range(P)):
            for s in range(P):
                if slot[r][k % ring][s] != k:
                    return "rank %d summing reduction %d found the partial of reduction %d from rank %d" % (
                        r, k, slot[r][k % ring][s], s)
            pc[r] = [k + 1, 0]
            if k + 1 > nred:
                done += 1
    return None


@pytest.mark.parametrize("P", [2, 4, 8])
def test_mailbox_ring_of_four_slots_is_race_free(P):
    for seed in range(60):
        assert simulate_mailbox(P, nred=30, ring=4, seed=seed) is None
        assert simulate_mailbox(P, nred=30, ring=2, seed=seed) is None  # two would already do; four is margin
    # a single slot is not enough: a fast rank overwrites a partial that a slow one has not summed yet
    assert any(simulate_mailbox(P, nred=30, ring=1, seed=seed) is not None for seed in range(60))

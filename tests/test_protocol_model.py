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


def simulate_slab_push(P, partners, nsteps, nbuf, reductions, seed, steps_per_read=3):
    """Slab exchange of the matrix-free Heisenberg operator, fused into the pass that produces w (csrc/cgs.cu SlabPush,
    csrc/heisenberg.cu).  One Lanczos step of rank r:
        apply x      : wait until every partner's flag >= x, then read receive buffer x % nbuf (takes several scheduler steps)
        reductions   : `reductions` all-rank reductions (push a partial to everybody, then wait for everybody's) — the
                       mailbox exchanges of DOT / UPDATE_DOT; 0 models back-to-back applies
        UPDATE_NORM  : store the slabs of exchange x+1 into the partners' buffers (x+1) % nbuf, raise flag x+1 there
    partners[r] is symmetric.  Returns None or the first violation."""
    rng = random.Random(seed)
    buf = [[{s: 0 for s in partners[r]} for _ in range(nbuf)] for r in range(P)]
    flag = [[0] * P for _ in range(P)]
    red = [[0] * P for _ in range(P)]  # red[r][s]: number of the last reduction partial rank s delivered to rank r
    # exchange 1 is pushed by a stand-alone kernel before the first apply
    pc = [[1, "push", 0, 0] for _ in range(P)]  # exchange, phase, reads left / reductions done, -
    done = 0
    while done < P:
        r = rng.randrange(P)
        x, ph, cnt, _ = pc[r]
        if x > nsteps:
            continue
        if ph == "push":  # stores of exchange x, then the flags
            for q in partners[r]:
                buf[q][x % nbuf][r] = x
                flag[q][r] = x
            pc[r] = [x, "wait", 0, 0]
        elif ph == "wait":
            if all(flag[r][s] >= x for s in partners[r]):
                pc[r] = [x, "read", steps_per_read, 0]
        elif ph == "read":
            for s, tag in buf[r][x % nbuf].items():
                if tag != x:
                    return "rank %d reading exchange %d found the slab of exchange %d from rank %d" % (r, x, tag, s)
            pc[r] = [x, "read", cnt - 1, 0] if cnt > 1 else [x, "reduce_push", 0, 0]
        elif ph == "reduce_push":
            if cnt == reductions:
                pc[r] = [x + 1, "push", 0, 0]  # UPDATE_NORM of step x carries exchange x + 1
                if x + 1 > nsteps:
                    done += 1
            else:
                k = (x - 1) * reductions + cnt + 1
                for q in range(P):
                    red[q][r] = k
                pc[r] = [x, "reduce_pull", cnt, 0]
        else:  # reduce_pull
            k = (x - 1) * reductions + cnt + 1
            if all(red[r][s] >= k for s in range(P)):
                pc[r] = [x, "reduce_push", cnt + 1, 0]
    return None


@pytest.mark.parametrize("P", [2, 4, 8])
def test_fused_slab_push_needs_two_receive_buffers_and_no_more(P):
    import itertools

    def heisenberg_partners(P):
        p = P.bit_length() - 1
        out = []
        for r in range(P):
            s = {r ^ 1, r ^ (1 << (p - 1))}  # straddle bond, periodic bond
            for b in range(p - 1):
                if ((r >> b) ^ (r >> (b + 1))) & 1:
                    s.add(r ^ (3 << b))  # anti-aligned rank-rank bond
            out.append(s - {r})
        return out

    patterns = {"heisenberg": heisenberg_partners(P), "all_to_all": [set(range(P)) - {r} for r in range(P)]}
    for (name, partners), reductions in itertools.product(patterns.items(), (0, 2)):
        assert all(r in partners[q] for r in range(P) for q in partners[r]), "partners must be symmetric"
        for seed in range(40):
            assert simulate_slab_push(P, partners, 12, nbuf=2, reductions=reductions, seed=seed) is None, (name, reductions, seed)
    # one buffer is not enough: a partner that finished its apply stores the next exchange while I am still reading
    assert any(simulate_slab_push(P, patterns["heisenberg"], 12, nbuf=1, reductions=0, seed=s) is not None for s in range(100))

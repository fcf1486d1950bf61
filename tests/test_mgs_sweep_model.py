"""Packet protocol of mgs_sweep_kernel (pymor_b200/csrc/mgs.cu), checked without a GPU.

Every CTA of every GPU runs the kernel's communication steps as a coroutine; a random scheduler interleaves
them (a post is visible to the others as soon as it is made, a wait yields until its flag shows up).  Checked:
no deadlock; every CTA of every rank obtains exactly sum_ranks(sum_ctas(partial)) for every step of every
sweep; a packet is never overwritten while somebody still has to read it (that would surface as a deadlock on
the exact-flag wait or as a wrong sum); two buffers per slot are enough, one is not.  The GPU parity of the
kernel is tests/test_gpu_parity.py::test_mgs_sweep_parity (+ tests/sharded_worker.py for several GPUs).
"""
import pytest


def partial(rank, cta, step):
    return float((rank + 1) * 1000 + (cta + 1) * 10) + step * 0.5


def simulate(ranks, ctas, sweeps, rnd, buffers=2):
    cta_box = [[[(0.0, 0)] * ctas for _ in range(buffers)] for _ in range(ranks)]      # [rank][parity][cta]
    inbox = [[[(0.0, 0)] * buffers for _ in range(ranks)] for _ in range(ranks)]       # [owner][sender][parity]
    results = {}

    def program(rank, cta):
        seq = 0
        for nq in sweeps:
            for j in range(nq + 1):
                seq += 1
                par = seq % buffers                   # parity of the GLOBAL step number (not of j: two
                                                      # sweeps in a row would share a buffer)
                cta_box[rank][par][cta] = (partial(rank, cta, seq), seq)                 # ll_post
                g = 0.0
                for i in range(ctas):                                                    # ll_wait on every CTA packet
                    while cta_box[rank][par][i][1] != seq:
                        yield
                    g += cta_box[rank][par][i][0]
                if ranks > 1:
                    if cta == 0:
                        for r in range(ranks):
                            inbox[r][rank][par] = (g, seq)
                            yield                                                        # posts are separate stores
                    tot = 0.0
                    for r in range(ranks):
                        while inbox[rank][r][par][1] != seq:
                            yield
                        tot += inbox[rank][r][par][0]
                    g = tot
                results[(rank, cta, seq)] = g
                yield

    progs = {(r, c): program(r, c) for r in range(ranks) for c in range(ctas)}
    idle = 0
    while progs:
        keys = list(progs)
        key = keys[rnd.integers(len(keys))]
        before = len(results)
        try:
            next(progs[key])
        except StopIteration:
            del progs[key]
        idle = 0 if len(results) != before else idle + 1
        if idle > 200000:
            return None                                                                   # deadlock
    steps = sum(nq + 1 for nq in sweeps)
    for seq in range(1, steps + 1):
        want = sum(partial(r, c, seq) for r in range(ranks) for c in range(ctas))
        for r in range(ranks):
            for c in range(ctas):
                if results[(r, c, seq)] != want:
                    return False
    return True


@pytest.mark.parametrize('ranks,ctas', [(1, 1), (1, 5), (2, 3), (3, 4), (8, 2)])
def test_sweep_packet_protocol(rng, ranks, ctas):
    for _ in range(4):
        assert simulate(ranks, ctas, [0, 1, 5, 2, 7], rng) is True


def test_one_buffer_is_not_enough(rng):
    """With a single buffer per slot a fast CTA overwrites a packet a slow one still waits for: the model must
    notice (deadlock on the exact-flag wait), i.e. it is able to see this class of bug."""
    outcomes = [simulate(2, 3, [6, 6], rng, buffers=1) for _ in range(20)]
    assert any(o is not True for o in outcomes)

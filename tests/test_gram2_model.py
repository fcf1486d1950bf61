"""Index maps of the second-generation Gram tile kernel (pymor_b200/csrc/gram2.cu), checked without a GPU.

tests/native/gram2_model.cpp replays the kernel lane by lane on the host with the SAME header of index
functions the kernel includes (pymor_b200/csrc/gram2_plan.h): swizzled TMA tiles, weight scaling slices,
DMMA fragment ownership, warp -> block plans (diagonal tiles: triangular blocks + K-split blocks; ragged
tiles: 8-row sub-block counts), epilogue scatter and the reduction incl. the mirrored K-halves.  The GPU
parity of the kernel itself is tests/test_gpu_parity.py::test_gram2_variant_parity.
"""
import ctypes
import subprocess

import numpy as np
import pytest

from conftest import ROOT

G2_ENABLE, G2_NO_TRI, G2_NO_KSPLIT, G2_NO_RAGGED = 16, 32, 64, 128


@pytest.fixture(scope='module')
def model(tmp_path_factory):
    so = tmp_path_factory.mktemp('g2') / 'libg2model.so'
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', str(so),
                    str(ROOT / 'tests' / 'native' / 'gram2_model.cpp')], check=True)
    lib = ctypes.CDLL(str(so))
    P, I64, I = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.g2_model_gram.argtypes = [P, I64, I, P, I64, I, I64, P, P, I64, I, I, I64, P]
    lib.g2_model_gram.restype = I
    return lib


def run_model(lib, A, B, w, symmetric, opts, rows_per_split, counts=None):
    n, kA = A.shape
    kB = B.shape[1]
    A = np.asfortranarray(A)
    B = A if symmetric else np.asfortranarray(B)
    out = np.full((kA, kB), np.nan, order='F')
    rc = lib.g2_model_gram(A.ctypes.data, n, kA, B.ctypes.data, n, kB, n, None if w is None else w.ctypes.data,
                           out.ctypes.data, kA, int(symmetric), opts, rows_per_split,
                           None if counts is None else counts.ctypes.data)
    assert rc == 0, f'model rc={rc} (-2: two warps wrote the same partial entry)'
    return out


OPTS = [G2_ENABLE, G2_ENABLE | G2_NO_TRI, G2_ENABLE | G2_NO_KSPLIT, G2_ENABLE | G2_NO_RAGGED,
        G2_ENABLE | G2_NO_TRI | G2_NO_KSPLIT | G2_NO_RAGGED]


@pytest.mark.parametrize('opts', OPTS)
@pytest.mark.parametrize('weighted', [False, True])
@pytest.mark.parametrize('k', [16, 97, 128, 129, 200, 232, 256, 377, 500])
def test_syrk_model(model, rng, k, weighted, opts):
    n = 70                                            # 3 splits of 32 rows, the last one ragged (6 rows)
    A = rng.standard_normal((n, k))
    w = rng.uniform(0.5, 1.5, n) if weighted else None
    G = run_model(model, A, A, w, True, opts, 32)
    ref = A.T @ (A if w is None else w[:, None] * A)
    assert not np.isnan(G).any()
    np.testing.assert_array_equal(G, G.T)
    np.testing.assert_allclose(G, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize('opts', [G2_ENABLE, G2_ENABLE | G2_NO_RAGGED, G2_ENABLE | G2_NO_KSPLIT])
@pytest.mark.parametrize('weighted', [False, True])
@pytest.mark.parametrize('ka,kb', [(17, 130), (130, 17), (128, 128), (129, 257), (200, 64), (100, 100), (250, 250),
                                   (224, 32), (32, 224), (128, 8), (300, 20), (40, 40), (64, 128), (16, 16)])
def test_general_inner_model(model, rng, ka, kb, weighted, opts):
    n = 53
    A, B = rng.standard_normal((n, ka)), rng.standard_normal((n, kb))
    w = rng.uniform(0.5, 1.5, n) if weighted else None
    G = run_model(model, A, B, w, False, opts, 48)
    ref = A.T @ (B if w is None else w[:, None] * B)
    assert not np.isnan(G).any()
    np.testing.assert_allclose(G, ref, rtol=0, atol=1e-12)


def test_syrk_1000_and_smsp_balance(model, rng):
    """k = 1000 (the benchmark width): 36 tiles, the last tile row is ragged (104 valid rows).  DMMA
    instructions per SMSP and stage: 256 for a full tile; a diagonal tile carries 136 on every SMSP
    (triangular block 40 + full block 64 + K-half 32); a ragged off-diagonal tile 208 on every SMSP."""
    n, k = 40, 1000
    A = rng.standard_normal((n, k))
    w = rng.uniform(0.5, 1.5, n)
    counts = np.zeros((36, 4), dtype=np.int64)
    G = run_model(model, A, A, w, True, G2_ENABLE, 48, counts)
    np.testing.assert_allclose(G, A.T @ (w[:, None] * A), rtol=0, atol=1e-12)
    # tile numbering: column by column of the lower triangle -> diagonal tiles at 0, 8, 15, 21, ...
    diag_tiles, t = [], 0
    for tj in range(8):
        diag_tiles.append(t)
        t += 8 - tj
    for tile in range(36):
        c = counts[tile]
        if tile in diag_tiles[:-1]:
            assert (c == 136).all(), (tile, c)
        elif tile == diag_tiles[-1]:                  # ragged diagonal tile: no K-split blocks
            assert c.max() <= 136
        elif (tile + 1) in diag_tiles[1:] or tile == 35 - 1:   # last tile of a column = tile row 7 (ragged)
            assert (c == 64 * 3 + 16).all(), (tile, c)
        else:
            assert (c == 256).all(), (tile, c)
    total, old_total = counts.max(axis=1).sum(), 28 * 256 + 8 * 192
    assert total < 0.92 * old_total                   # critical-path DMMA count vs gram_dmma_kernel's plan


def test_fragment_loads_are_bank_conflict_free_and_scaling_covers_the_tile(model):
    """Every fragment LDS.64 of a half-warp touches 16 distinct 8-byte bank pairs (shared memory has 32
    4-byte banks = 16 pairs), i.e. two wavefronts per warp load, the minimum; the address formula agrees
    with the swizzled tile layout; the 16 warps' scaling slices tile the stage exactly once and each
    128-bit access of a quarter-warp is one contiguous 128-byte row."""
    for wblk in range(4):
        for sub in range(4):
            for ks in range(4):
                addrs = [model.g2_model_frag_addr(wblk, sub, ks, lane) for lane in range(32)]
                assert addrs == [model.g2_model_frag_addr_by_layout(wblk, sub, ks, lane) for lane in range(32)]
                for half in (addrs[:16], addrs[16:]):
                    assert len({a % 16 for a in half}) == 16
    seen = np.zeros(2048, dtype=int)
    e, wo = ctypes.c_int(), ctypes.c_int()
    for warp in range(16):
        for part in range(2):
            elems = []
            for lane in range(32):
                model.g2_model_scale_slice(warp, lane, part, ctypes.byref(e), ctypes.byref(wo))
                assert e.value % 2 == 0 and wo.value % 2 == 0 and 0 <= wo.value < 16
                seen[e.value:e.value + 2] += 1
                elems.append(e.value)
            for q in range(4):                           # quarter-warp = 8 lanes x 16 B = one 128-byte tile row
                assert elems[8 * q:8 * q + 8] == list(range(elems[8 * q], elems[8 * q] + 16, 2))
                assert elems[8 * q] % 16 == 0
    assert (seen == 1).all()


# ------------------------------------------------------------------------------------------------------
# mbarrier protocol of gram2_dmma_kernel<WEIGHTED> as a randomly scheduled discrete-event simulation
# ------------------------------------------------------------------------------------------------------
class _MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):            # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


def _simulate_protocol(nk, weighted, diag, rnd, stages=6, look=3, warps=16):
    full = [_MBar(1) for _ in range(stages)]
    empty = [_MBar(warps) for _ in range(stages)]
    ready = [_MBar(warps) for _ in range(stages)]
    slotA, slotB = [None] * stages, [None] * stages     # stage index the slot holds
    scaled = [set() for _ in range(nk)]                  # warps that scaled stage kt
    computed = [set() for _ in range(nk)]
    in_flight = []                                       # TMA loads issued, not landed

    def issue(kt):
        in_flight.append(kt)

    def land(kt):
        s = kt % stages
        if kt >= stages:                                 # the slot's previous stage must be dead
            assert len(computed[kt - stages]) == warps
            assert not weighted or len(scaled[kt - stages]) == warps
        slotA[s] = kt
        if not diag:
            slotB[s] = ('raw', kt)
        full[s].arrive()

    def scale(w, kt):
        s = kt % stages
        yield lambda: full[s].done((kt // stages) & 1)
        assert slotA[s] == kt
        if diag:                                         # scaled copy goes to the B slot: its old readers are done
            assert kt < stages or len(computed[kt - stages]) == warps
        else:
            assert slotB[s] in (('raw', kt), ('scaled', kt))
        slotB[s] = ('scaled', kt)
        assert w not in scaled[kt]
        scaled[kt].add(w)
        ready[s].arrive()

    def warp_program(w):
        if w == 0:
            for kt in range(min(nk, stages)):
                issue(kt)
        if weighted:
            for kt in range(min(nk, look)):
                yield from scale(w, kt)
        for kt in range(nk):
            s = kt % stages
            if w == 0 and kt >= 1 and kt - 1 + stages < nk:
                sp = (kt - 1) % stages
                yield lambda: empty[sp].done(((kt - 1) // stages) & 1)
                issue(kt - 1 + stages)
            bar = ready[s] if weighted else full[s]
            yield lambda: bar.done((kt // stages) & 1)
            assert slotA[s] == kt
            if weighted:
                assert len(scaled[kt]) == warps and slotB[s] == ('scaled', kt)
            elif not diag:
                assert slotB[s] == ('raw', kt)
            computed[kt].add(w)
            yield lambda: True                            # the MMA takes a while: let others run
            if weighted and kt + look < nk:
                yield from scale(w, kt + look)
            empty[s].arrive()

    progs = {w: warp_program(w) for w in range(warps)}
    waiting = {}
    for w, g in list(progs.items()):
        try:
            waiting[w] = next(g)
        except StopIteration:
            del progs[w]
    steps = 0
    while progs or in_flight:
        steps += 1
        assert steps < 10_000_000
        choices = [('warp', w) for w in progs if waiting[w]()] + [('tma', i) for i in range(len(in_flight))]
        assert choices, f'deadlock: warps {sorted(progs)} blocked, nothing in flight'
        kind, x = choices[rnd.integers(len(choices))]
        if kind == 'tma':
            land(in_flight.pop(x))
        else:
            try:
                waiting[x] = next(progs[x])
            except StopIteration:
                del progs[x], waiting[x]
    assert all(len(c) == warps for c in computed)


@pytest.mark.parametrize('weighted', [False, True])
@pytest.mark.parametrize('diag', [False, True])
def test_barrier_protocol_simulation(rng, weighted, diag):
    """Random interleavings of the 16 warps and the TMA completions: no deadlock, no phase aliasing (a
    parity wait never passes on a stale phase), every stage is scaled exactly once by every warp before
    anybody multiplies it, and no slot is overwritten (by TMA or by the scaling) while it still has readers."""
    for nk in (1, 2, 3, 4, 6, 7, 13, 40):
        for _ in range(6):
            _simulate_protocol(nk, weighted, diag, rng)


def test_narrow_tiles_use_all_warps(model, rng):
    """A 128 x 32 tile (a panel of block Gram-Schmidt against 128 previous vectors) has 4 blocks of data: their K
    range is split four ways, so all 16 warps work and every SMSP issues the same 64 DMMA per stage; a 128 x 64
    tile splits two ways; with the split disabled the same work sits on 4 (8) warps."""
    n = 40
    A, B = rng.standard_normal((n, 128)), rng.standard_normal((n, 32))
    for kb, opts in ((32, G2_ENABLE), (64, G2_ENABLE), (32, G2_ENABLE | G2_NO_KSPLIT)):
        B = rng.standard_normal((n, kb))
        counts = np.zeros((1, 4), dtype=np.int64)
        G = run_model(model, A, B, None, False, opts, 48, counts)
        np.testing.assert_allclose(G, A.T @ B, rtol=0, atol=1e-12)
        assert (counts[0] == 16 * kb // 8).all(), counts          # 4 (8) blocks x 64 DMMA over the four SMSPs

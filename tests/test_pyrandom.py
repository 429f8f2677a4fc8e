"""cpg_pyrandom_draw_shuffles against CPython's own `random`: the permutation, k and the blinders
of every proof, and the generator state afterwards, must be exactly what the reference's Python calls
(random.shuffle, random_scalar = randint(1, r - 1); cp/util.py:21-24, cp/whisk_interface.py:114-116) produce."""
import ctypes
import random

import pytest

from curdleproofs_pie_b200 import runtime as rt


def python_draw(rng, ell, n_rand, B):
    perms, ks, rands = [], b"", b""
    for _ in range(B):
        perm = list(range(ell))
        rng.shuffle(perm)
        perms += perm
        ks += rng.randint(1, rt.R_ORDER - 1).to_bytes(32, "little")
        rands += b"".join(rng.randint(1, rt.R_ORDER - 1).to_bytes(32, "little") for _ in range(n_rand))
    return perms, ks, rands


@pytest.mark.parametrize("seed,ell,n_rand,B,burn", [(1234, 4, 37, 3, 0), (7, 124, 397, 2, 5), (2**70 + 3, 60, 205, 4, 623), (0, 1, 1, 5, 1)])
def test_draws_equal_cpythons(seam_lib, seed, ell, n_rand, B, burn):
    a, b = random.Random(seed), random.Random(seed)
    for _ in range(burn):                                   # start at an arbitrary position inside the 624-word block
        a.getrandbits(32); b.getrandbits(32)
    want = python_draw(a, ell, n_rand, B)
    version, words, gauss = b.getstate()
    st = (ctypes.c_uint32 * 625)(*words)
    perms = (ctypes.c_uint32 * (B * ell))()
    ks = ctypes.create_string_buffer(B * 32)
    rand = ctypes.create_string_buffer(B * n_rand * 32)
    seam_lib.check(seam_lib.c.cpg_pyrandom_draw_shuffles(st, ell, n_rand, B, perms, ks, rand))
    assert list(perms) == want[0]
    assert ks.raw == want[1]
    assert rand.raw == want[2]
    b.setstate((version, tuple(st), gauss))
    assert [a.random() for _ in range(5)] == [b.random() for _ in range(5)]      # the stream continues identically


def test_batch_prover_draws_through_c_and_leaves_python_state(seam_lib):
    import prove_cases as pc
    import shuffle_cases as sc
    from curdleproofs_pie_b200 import whisk

    case = sc.load_case("shuffle_N8_seed1234.json")
    ell = case["N"] - 4
    prover = whisk.BatchProver(bytes.fromhex(case["crs"]), ell, fixed_window=4, lib=seam_lib)
    # replay the fixture's stream up to the shuffle (oracle/gen_golden.py::one_case): CRS draws
    random.seed(case["seed"])
    for _ in range(ell + 4 + 3):
        random.randint(1, rt.R_ORDER - 1)
    # the reference draws vec_R / vec_S between k and the blinders, so one proof cannot be replayed through
    # draw_batch; what must hold is the order perm -> k -> blinders and Python's state afterwards
    probe = random.Random(99)
    twin = random.Random(99)
    perms, ks, rand = prover.draw_batch(probe, 2)
    want = python_draw(twin, ell, prover.n_rand, 2)
    assert (list(perms), ks, rand) == want
    assert probe.random() == twin.random()
    # a generator that is not a CPython Mersenne Twister is driven call by call, same order
    class Counting(random.Random):
        pass
    c1, c2 = Counting(5), random.Random(5)
    assert tuple(map(lambda x: list(x) if not isinstance(x, bytes) else x, prover.draw_batch(c1, 1))) == python_draw(c2, ell, prover.n_rand, 1)
    prover.close()
    assert pc is not None

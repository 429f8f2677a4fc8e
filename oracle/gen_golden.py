"""ORACLE tooling: generate tests/golden/*.json by running the UNMODIFIED reference
(/root/reference/curdleproofs + merlin_transcripts, imported in place, nothing copied) on top of
the oracle's py_arkworks_bls12381 stand-in (oracle/standin/, C backend) under fixed Python
``random`` seeds.  The reference cannot travel to the GPU box, so its outputs are committed here
as fixtures together with this script.

    python oracle/gen_golden.py            # needs /root/reference

Per case: seed, N, CRS bytes (cp/crs.py:93-102), compressed vec_R/vec_S/vec_T/vec_U/M, the proof's
wire bytes (cp/curdleproofs.py:275-285), and the verifier's verdict on it and on corrupted
variants (cp/test_curdleproofs.py:643-670 style).  Input construction order is the reference
test's (cp/test_curdleproofs.py:576-593): CRS -> shuffle(perm) -> k -> vec_R -> vec_S.
"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "standin"), os.path.join(REF, "curdleproofs"), os.path.join(REF, "merlin_transcripts"), ROOT]

from oracle import ark_surface, cref_binding, merlin_py  # noqa: E402

ark_surface.set_backend("c")

# the reference's pure-Python Keccak costs 0.7 ms/permutation; swap in the C permutation (same
# function, checked in tests/test_oracle_kat.py) underneath the reference's own Strobe class
import merlin_transcripts.keccak as ref_keccak  # noqa: E402
import merlin_transcripts.strobe as ref_strobe  # noqa: E402

_c = cref_binding.load()


def _fast_f1600(state):
    return _c.keccak_f1600(state)


from curdleproofs.crs import CurdleproofsCrs  # noqa: E402
from curdleproofs.curdleproofs import CurdleProofsProof, N_BLINDERS, shuffle_permute_and_commit_input  # noqa: E402
from curdleproofs.util import BufReader, get_random_point, point_projective_to_bytes, random_scalar  # noqa: E402


def one_case(seed, N):
    random.seed(seed)
    ell = N - N_BLINDERS
    crs = CurdleproofsCrs.new(ell, N_BLINDERS)
    perm = list(range(ell))
    random.shuffle(perm)
    k = random_scalar()
    vec_R = [get_random_point() for _ in range(ell)]
    vec_S = [get_random_point() for _ in range(ell)]
    vec_T, vec_U, M, bl = shuffle_permute_and_commit_input(crs, vec_R, vec_S, perm, k)
    proof = CurdleProofsProof.new(crs=crs, vec_R=vec_R, vec_S=vec_S, vec_T=vec_T, vec_U=vec_U, M=M,
                                  permutation=perm, k=k, vec_m_blinders=bl)
    wire = proof.to_bytes()
    proof.verify(crs, vec_R, vec_S, vec_T, vec_U, M)            # the reference accepts its own proof

    def verdict(R_, S_, T_, U_, M_):
        try:
            CurdleProofsProof.from_bytes(BufReader(wire), N).verify(crs, R_, S_, T_, U_, M_)
            return True
        except AssertionError:
            return False

    enc = lambda pts: [point_projective_to_bytes(p).hex() for p in pts]  # noqa: E731
    case = {
        "seed": seed, "N": N,
        "crs": crs.to_bytes().hex(),
        "vec_R": enc(vec_R), "vec_S": enc(vec_S), "vec_T": enc(vec_T), "vec_U": enc(vec_U),
        "M": point_projective_to_bytes(M).hex(),
        "perm": perm, "k": int(k),
        "proof": wire.hex(), "proof_sha256": hashlib.sha256(wire).hexdigest(),
        "verdicts": {
            "honest": verdict(vec_R, vec_S, vec_T, vec_U, M),
            "swap_R_S": verdict(vec_S, vec_R, vec_T, vec_U, M),
            "swap_T_U": verdict(vec_R, vec_S, vec_U, vec_T, M),
            "wrong_M": verdict(vec_R, vec_S, vec_T, vec_U, M + M),
            "rotated_T": verdict(vec_R, vec_S, vec_T[1:] + vec_T[:1], vec_U, M),
        },
    }
    return case


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    # N=8 runs with the reference's own pure-Python Keccak; the larger ones use the C permutation
    for seed, N, fast in ((1234, 8, False), (77, 16, True), (2024, 64, True), (4096, 128, True)):
        ref_strobe.KeccakF1600 = _fast_f1600 if fast else ref_keccak.KeccakF1600
        case = one_case(seed, N)
        case["keccak"] = "c" if fast else "reference-python"
        path = os.path.join(out_dir, "shuffle_N%d_seed%d.json" % (N, seed))
        with open(path, "w") as f:
            json.dump(case, f, indent=0)
        print(path, len(case["proof"]) // 2, "bytes", case["proof_sha256"][:16], case["verdicts"])


if __name__ == "__main__":
    main()

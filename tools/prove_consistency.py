"""Diagnostic: the batched prover must emit the same bytes for the same (inputs, randomness) whatever the batch split
(lanes), batch size or transcript placement.  Prints, per configuration, how many proofs differ from the single-lane
host-transcript run and the first differing byte offset (which names the first wrong output)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import shuffle_cases as sc
    from curdleproofs_pie_b200 import runtime as rt, whisk

    lib = rt.get_lib()
    name = sys.argv[1] if len(sys.argv) > 1 else "shuffle_N128_seed4096.json"
    case = sc.load_case(name)
    ell = case["N"] - 4
    crs = bytes.fromhex(case["crs"])
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    B = int(os.environ.get("PC_B", "96"))

    def run(lanes, mode, nb=B):
        prover = whisk.BatchProver(crs, ell, lib=lib)
        if lanes:
            prover.set_lanes(*lanes)
        prover.set_transcript(mode)
        rng = random.Random(2718)
        perms, ks, rands = [], [], []
        for _ in range(nb):
            p = list(range(ell)); rng.shuffle(p)
            perms.append(p); ks.append(rng.randint(1, rt.R_ORDER - 1)); rands.append(prover.draw_randomness(rng))
        res = prover.prove([pre] * nb, perms, ks, rands)
        prover.close()
        return res

    ref = run(None, "host")
    ver = whisk.BatchVerifier(crs, ell, lib=lib)
    print("reference run valid:", sum(ver.verify([pre + tu for tu, _ in ref], [pr for _, pr in ref])), "of", B)
    for label, lanes, mode, nb in [("1 lane device", None, "device", B), ("4 lanes host", (4, 2), "host", B), ("4 lanes device", (4, 2), "device", B),
                                   ("2 lanes host", (2, 2), "host", B), ("1 lane host B=24", None, "host", 24), ("1 lane device B=24", None, "device", 24),
                                   ("1 lane host again", None, "host", B)]:
        got = run(lanes, mode, nb)
        bad = [(i, next((j for j in range(len(g[1])) if g[1][j] != r[1][j]), -1), g[0] != r[0]) for i, (g, r) in enumerate(zip(got, ref)) if g != r]
        print("%-22s differing proofs: %d  first: %s" % (label, len(bad), bad[:6]))


if __name__ == "__main__":
    main()

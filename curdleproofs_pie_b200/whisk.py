"""Batched Whisk-facing API: the reference's bytes-in / bool-out entry points, B proofs per call.

Mirrors /root/reference/curdleproofs/curdleproofs/whisk_interface.py:
    IsValidWhiskShuffleProof(crs, pre_trackers, post_trackers, proof_bytes) -> bool        (:74-87)
becomes
    IsValidWhiskShuffleProofBatch(crs, [pre_trackers], [post_trackers], [proof_bytes]) -> [bool]
with the same acceptance rule per proof (any decoding problem or failed check -> False).
All group arithmetic runs on the GPU through cpg_verify_batch (include/cpg.h); there is no CPU path.
"""
import ctypes

from . import runtime as _rt


class BatchVerifier:
    """Holds the device-resident CRS (fixed-base tables) for one (ell, n_blinders)."""

    def __init__(self, crs_bytes, ell, n_blinders=4, fixed_window=0, host_threads=0, lib=None, group=0, sharded=False):
        """group: proofs per aggregated check (cpg_verifier_set_group): 0 = adaptive (default: the library re-picks
        the group size after every batch from the observed rate of failing proofs), 1 = one MSM per proof.
        sharded: every proof is checked by ALL ranks of the library's communicator together (comm.init; BASELINE
        config 5) - every rank makes the same calls with the same arguments."""
        self.lib = lib or _rt.get_lib()
        self.ell = int(ell)
        self.n_blinders = int(n_blinders)
        crs_bytes = bytes(crs_bytes)
        if len(crs_bytes) != 48 * (self.ell + self.n_blinders + 5):
            raise ValueError("crs_bytes must be CurdleproofsCrs.to_bytes() for (ell, n_blinders)")
        create = self.lib.c.cpg_verifier_create_sharded if sharded else self.lib.c.cpg_verifier_create
        self.handle = create(crs_bytes, self.ell, self.n_blinders, fixed_window, host_threads)
        if not self.handle:
            raise _rt.CpgError("cpg_verifier_create failed: " + self.lib.last_error())
        self.proof_len = int(self.lib.c.cpg_verifier_proof_bytes(self.handle))
        self.input_len = int(self.lib.c.cpg_verifier_input_bytes(self.handle))
        self.set_group(group)

    def set_window(self, c):
        self.lib.check(self.lib.c.cpg_verifier_set_window(self.handle, int(c)), "cpg_verifier_set_window")

    def set_transcript(self, mode):
        """True / "device": transcript + coefficients per proof on the GPU, one thread per proof; "warp": one warp per
        proof (shorter, for tens to a few thousand proofs); False / "host": on host threads; "auto" (the default): by
        batch size."""
        mode = {"host": 0, "device": 1, "auto": 2, "warp": 3, True: 1, False: 0}.get(mode, mode)
        self.lib.check(self.lib.c.cpg_verifier_set_transcript(self.handle, int(mode)), "cpg_verifier_set_transcript")

    def set_streams(self, n):
        self.lib.check(self.lib.c.cpg_verifier_set_streams(self.handle, int(n)), "cpg_verifier_set_streams")

    def set_group(self, group, window=0):
        """Cross-proof aggregation: `group` consecutive proofs share one MSM; failing groups are re-checked
        proof by proof (verdicts stay exact).  1 = off, 0 = adaptive (re-picked after every batch)."""
        self.lib.check(self.lib.c.cpg_verifier_set_group(self.handle, int(group), int(window)), "cpg_verifier_set_group")

    def group(self):
        return int(self.lib.c.cpg_verifier_group(self.handle))

    def rechecked(self):
        return int(self.lib.c.cpg_verifier_rechecked(self.handle))

    def set_cache(self, log2_slots):
        """Device-resident cache of decompressed tracker points keyed by their 48-byte encodings (0 = off): pre-shuffle
        trackers of one Whisk shuffle are post-shuffle trackers of an earlier one, and duplicates inside a batch are
        decompressed once.  2^log2_slots entries of 160 B."""
        self.lib.check(self.lib.c.cpg_verifier_set_cache(self.handle, int(log2_slots)), "cpg_verifier_set_cache")

    def cache_reset(self):
        self.lib.check(self.lib.c.cpg_verifier_cache_reset(self.handle), "cpg_verifier_cache_reset")

    def cache_stats(self):
        """{lookups, served, claimed} since creation / cache_reset"""
        out = (ctypes.c_uint64 * 3)()
        self.lib.check(self.lib.c.cpg_verifier_cache_stats(self.handle, out), "cpg_verifier_cache_stats")
        return {"lookups": int(out[0]), "served": int(out[1]), "claimed": int(out[2])}

    def verify_raw(self, inputs, proofs, B):
        """inputs: B*input_len bytes, proofs: B*proof_len bytes -> bytes of B verdicts."""
        out = ctypes.create_string_buffer(max(1, B))
        self.lib.check(self.lib.c.cpg_verify_batch(self.handle, inputs, proofs, B, out), "cpg_verify_batch")
        return out.raw[:B]

    def verify(self, inputs, proofs):
        """inputs[i] = vec_R|vec_S|vec_T|vec_U (48-byte points), proofs[i] = M|proof wire bytes.
        Wrong-length entries are rejected (the reference's BufReader would raise -> False)."""
        B = len(inputs)
        if len(proofs) != B:
            raise ValueError("inputs and proofs must have the same length")
        ok_len = [len(inputs[i]) == self.input_len and len(proofs[i]) >= self.proof_len for i in range(B)]
        idx = [i for i in range(B) if ok_len[i]]
        verdicts = [False] * B
        if idx:
            raw = self.verify_raw(b"".join(bytes(inputs[i]) for i in idx), b"".join(bytes(proofs[i])[:self.proof_len] for i in idx), len(idx))
            for j, i in enumerate(idx):
                verdicts[i] = bool(raw[j])
        return verdicts

    def close(self):
        if self.handle:
            self.lib.c.cpg_verifier_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def trackers_to_input(pre_trackers, post_trackers):
    """[(r_G, k_r_G)...] pre/post -> vec_R|vec_S|vec_T|vec_U bytes (whisk_interface.py:96-100).
    Trackers may be WhiskTracker-like objects (.r_G/.k_r_G) or (r_G, k_r_G) tuples."""
    def halves(ts):
        a, b = [], []
        for t in ts:
            r, k = (t.r_G, t.k_r_G) if hasattr(t, "r_G") else t
            a.append(bytes(r)); b.append(bytes(k))
        return b"".join(a), b"".join(b)

    R, S = halves(pre_trackers)
    T, U = halves(post_trackers)
    return R + S + T + U


_CACHE = {}


def IsValidWhiskShuffleProofBatch(crs, pre_shuffle_trackers, post_shuffle_trackers, whisk_shuffle_proofs, cache_log2=None):
    """crs: object with to_bytes()/vec_G/vec_H (the reference's CurdleproofsCrs) or (crs_bytes, ell).
    cache_log2: None leaves the (per-CRS, process-wide) verifier's tracker cache as it is (off unless set before);
    k > 0 gives it 2^k slots (BatchVerifier.set_cache: trackers met in earlier calls are not decompressed again); 0 = off."""
    if isinstance(crs, tuple):
        crs_bytes, ell = crs
        nbl = 4
    else:
        crs_bytes, ell, nbl = crs.to_bytes(), len(crs.vec_G), len(crs.vec_H)
    key = (bytes(crs_bytes), ell, nbl)
    ver = _CACHE.get(key)
    if ver is None:
        ver = _CACHE[key] = BatchVerifier(crs_bytes, ell, nbl)
        ver._cache_log2 = 0
    if cache_log2 is not None and int(cache_log2) != ver._cache_log2:
        ver.set_cache(int(cache_log2))
        ver._cache_log2 = int(cache_log2)
    inputs = []
    for pre, post in zip(pre_shuffle_trackers, post_shuffle_trackers):
        try:
            inputs.append(trackers_to_input(pre, post))
        except Exception:
            inputs.append(b"")
    return ver.verify(inputs, list(whisk_shuffle_proofs))


class BatchProver:
    """Device-resident CRS tables + lock-step proof generation (cpg_prove_batch)."""

    def __init__(self, crs_bytes, ell, n_blinders=4, fixed_window=0, lib=None, sharded=False):
        """sharded: ONE proof at a time over ALL ranks of the library's communicator (comm.init; BASELINE config 5) -
        every rank makes the same calls with the same arguments and gets the same bytes."""
        self.lib = lib or _rt.get_lib()
        self.ell = int(ell)
        self.n = self.ell + int(n_blinders)
        crs_bytes = bytes(crs_bytes)
        if len(crs_bytes) != 48 * (self.n + 5):
            raise ValueError("crs_bytes must be CurdleproofsCrs.to_bytes() for (ell, n_blinders)")
        create = self.lib.c.cpg_prover_create_sharded if sharded else self.lib.c.cpg_prover_create
        self.handle = create(crs_bytes, self.ell, int(n_blinders), fixed_window)
        if not self.handle:
            raise _rt.CpgError("cpg_prover_create failed: " + self.lib.last_error())
        self.proof_len = int(self.lib.c.cpg_prover_proof_bytes(self.handle))
        self.n_rand = int(self.lib.c.cpg_prover_rand_scalars(self.handle))

    def set_window(self, c):
        self.lib.check(self.lib.c.cpg_prover_set_window(self.handle, int(c)), "cpg_prover_set_window")

    def set_table_window(self, c):
        self.lib.check(self.lib.c.cpg_prover_set_table_window(self.handle, int(c)), "cpg_prover_set_table_window")

    def set_transcript(self, mode):
        """0 / "host": Fiat-Shamir on host threads; 1 / "device": one GPU thread per proof; 2 / "auto": by batch size"""
        mode = {"host": 0, "device": 1, "auto": 2}.get(mode, mode)
        self.lib.check(self.lib.c.cpg_prover_set_transcript(self.handle, int(mode)), "cpg_prover_set_transcript")

    def set_lanes(self, k, min_proofs_per_lane=0):
        self.lib.check(self.lib.c.cpg_prover_set_lanes(self.handle, int(k), int(min_proofs_per_lane)), "cpg_prover_set_lanes")

    def draw_randomness(self, rng):
        """Blinders for ONE proof in the reference's draw order, from a `random`-like
        object; call it right after drawing the permutation and k as the reference does."""
        order = _rt.R_ORDER
        return b"".join(rng.randint(1, order - 1).to_bytes(32, "little") for _ in range(self.n_rand))

    def draw_batch(self, rng, B):
        """(perms, ks, rand) for B proofs - flat u32 array, B*32 bytes, B*n_rand*32 bytes - drawn from `rng` in the
        reference's order (shuffle, k, blinders per proof).  A CPython Mersenne Twister (the `random` module or a
        random.Random) is continued in C (cpg_pyrandom_draw_shuffles) and left where Python would have left it;
        any other generator is driven call by call."""
        import array
        import random as _random

        is_mt = rng is _random or type(rng) is _random.Random
        if is_mt:
            version, words, gauss = rng.getstate()
            if version == 3 and len(words) == 625:
                st = (ctypes.c_uint32 * 625)(*words)
                perms = (ctypes.c_uint32 * (B * self.ell))()
                ks = ctypes.create_string_buffer(max(1, B * 32))
                rand = ctypes.create_string_buffer(max(1, B * self.n_rand * 32))
                self.lib.check(self.lib.c.cpg_pyrandom_draw_shuffles(st, self.ell, self.n_rand, B, perms, ks, rand), "cpg_pyrandom_draw_shuffles")
                rng.setstate((version, tuple(st), gauss))
                return perms, ks.raw[:B * 32], rand.raw[:B * self.n_rand * 32]
        perms, ks, rands = array.array("I"), bytearray(), bytearray()
        for _ in range(B):
            perm = list(range(self.ell))
            rng.shuffle(perm)
            perms.extend(perm)
            ks += rng.randint(1, _rt.R_ORDER - 1).to_bytes(32, "little")
            rands += self.draw_randomness(rng)
        return (ctypes.c_uint32 * len(perms)).from_buffer(perms), bytes(ks), bytes(rands)

    def prove_drawn(self, pre_inputs, rng):
        """Proves len(pre_inputs) shuffles with randomness drawn from `rng` as the reference draws it.
        Returns [(vec_T|vec_U bytes, M|proof bytes)]."""
        B = len(pre_inputs)
        pre_inputs = self._checked_inputs(pre_inputs)
        perms, ks, rand = self.draw_batch(rng, B)
        out_tu = ctypes.create_string_buffer(max(1, B * 2 * self.ell * 48))
        out_pr = ctypes.create_string_buffer(max(1, B * self.proof_len))
        status = ctypes.create_string_buffer(max(1, B))
        self.lib.check(self.lib.c.cpg_prove_batch(self.handle, b"".join(pre_inputs), perms, ks, rand, B, out_tu, out_pr, status), "cpg_prove_batch")
        if any(status.raw[:B]):
            raise ValueError("serialised data seems to be invalid (lanes %s)" % [i for i, s in enumerate(status.raw[:B]) if s])
        w = 2 * self.ell * 48
        tu, pr = out_tu.raw, out_pr.raw
        return [(tu[i * w:(i + 1) * w], pr[i * self.proof_len:(i + 1) * self.proof_len]) for i in range(B)]

    def _checked_inputs(self, pre_inputs):
        """every entry must be exactly vec_R|vec_S = 2*ell*48 bytes: libcpg reads B full rows from the joined buffer"""
        want = 2 * self.ell * 48
        out = [bytes(p) for p in pre_inputs]
        for i, p in enumerate(out):
            if len(p) != want:
                raise ValueError("pre_inputs[%d] holds %d bytes, expected 2*ell*48 = %d" % (i, len(p), want))
        return out

    def prove_raw(self, inputs, perms, ks, rand, B):
        """Flat buffers for B proofs; every length is checked here because libcpg reads B full rows of each.
        Malformed VALUES (bad encodings, k or a blinder >= r, a perms row that is not a permutation) come back in the
        per-lane status bytes."""
        import array

        perm_arr = array.array("I", perms)
        if perm_arr.itemsize != 4 or len(perm_arr) != B * self.ell:
            raise ValueError("perms must hold B*ell = %d u32 entries, got %d" % (B * self.ell, len(perm_arr)))
        for name, buf, want in (("inputs", inputs, B * 2 * self.ell * 48), ("ks", ks, B * 32), ("rand", rand, B * self.n_rand * 32)):
            if len(buf) != want:
                raise ValueError("%s holds %d bytes, expected %d" % (name, len(buf), want))
        out_tu = ctypes.create_string_buffer(B * 2 * self.ell * 48)
        out_pr = ctypes.create_string_buffer(B * self.proof_len)
        status = ctypes.create_string_buffer(max(1, B))
        pbuf = (ctypes.c_uint32 * len(perm_arr)).from_buffer(perm_arr)
        self.lib.check(self.lib.c.cpg_prove_batch(self.handle, inputs, pbuf, ks, rand, B, out_tu, out_pr, status), "cpg_prove_batch")
        return out_tu.raw, out_pr.raw, status.raw[:B]

    def prove(self, pre_inputs, perms, ks, rands):
        """pre_inputs[i] = vec_R|vec_S bytes, perms[i] = list of ell ints, ks[i] = int, rands[i] = bytes
        from draw_randomness.  Returns [(vec_T|vec_U bytes, M|proof bytes)]; raises on a malformed input."""
        B = len(pre_inputs)
        if not (len(perms) == len(ks) == len(rands) == B):
            raise ValueError("pre_inputs, perms, ks and rands must have the same length")
        pre_inputs = self._checked_inputs(pre_inputs)
        for i in range(B):
            if len(perms[i]) != self.ell:
                raise ValueError("perms[%d] has %d entries, expected ell = %d" % (i, len(perms[i]), self.ell))
            if any(not 0 <= int(x) < self.ell for x in perms[i]):
                raise IndexError("perms[%d] holds an index outside [0, ell)" % i)          # the reference's list index raises the same
            if len(rands[i]) != self.n_rand * 32:
                raise ValueError("rands[%d] holds %d bytes, expected n_rand*32 = %d" % (i, len(rands[i]), self.n_rand * 32))
            if not 0 <= int(ks[i]) < 1 << 256:
                raise ValueError("ks[%d] does not fit 32 bytes" % i)
        flat_perm = [int(x) for p in perms for x in p]
        tu, pr, st = self.prove_raw(b"".join(pre_inputs), flat_perm, b"".join(int(k).to_bytes(32, "little") for k in ks), b"".join(bytes(r) for r in rands), B)
        if any(st):
            raise ValueError("serialised data seems to be invalid (lanes %s)" % [i for i, s in enumerate(st) if s])
        w = 2 * self.ell * 48
        return [(tu[i * w:(i + 1) * w], pr[i * self.proof_len:(i + 1) * self.proof_len]) for i in range(B)]

    def close(self):
        if self.handle:
            self.lib.c.cpg_prover_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WhiskTracker:
    """(r_G, k_r_G) as 48-byte encodings - the shape of the reference's WhiskTracker (whisk_interface.py:30-33);
    also unpacks like the (r_G, k_r_G) tuple earlier versions returned."""

    __slots__ = ("r_G", "k_r_G")

    def __init__(self, r_G, k_r_G):
        self.r_G = bytes(r_G)
        self.k_r_G = bytes(k_r_G)

    def __iter__(self):
        return iter((self.r_G, self.k_r_G))

    def __eq__(self, other):
        try:
            a, b = other
        except (TypeError, ValueError):
            return NotImplemented
        return (self.r_G, self.k_r_G) == (bytes(a), bytes(b))

    def __repr__(self):
        return "WhiskTracker(r_G=%s..., k_r_G=%s...)" % (self.r_G[:4].hex(), self.k_r_G[:4].hex())


_PROVER_CACHE = {}


def GenerateWhiskShuffleProofBatch(crs, pre_shuffle_trackers_per_proof, rng=None):
    """Batched GenerateWhiskShuffleProof (whisk_interface.py:111-140): for each proof draws the
    permutation, k and the blinders from `rng` (default: the `random` module) exactly as the
    reference does, then proves all of them on the GPU.  Returns [(post_trackers, proof_bytes)] with
    post_trackers a list of WhiskTracker.
    The default generator is Python's Mersenne Twister, as in the reference (cp/util.py:21-24) - reproducible, NOT a
    CSPRNG; pass random.SystemRandom() for blinders that are actually unpredictable.
    The prover (CRS tables, device and pinned buffers) is kept per (crs, ell, n_blinders) like the verifier's."""
    import random as _random

    rng = rng or _random
    if isinstance(crs, tuple):
        crs_bytes, ell = crs
        nbl = 4
    else:
        crs_bytes, ell, nbl = crs.to_bytes(), len(crs.vec_G), len(crs.vec_H)
    inputs = []
    for i, trackers in enumerate(pre_shuffle_trackers_per_proof):
        trackers = list(trackers)
        if len(trackers) != ell:
            raise ValueError("proof %d: %d pre-shuffle trackers, expected ell = %d" % (i, len(trackers), ell))
        row = trackers_to_input(trackers, [])
        if len(row) != 2 * ell * 48:
            raise ValueError("proof %d: every tracker half must be a 48-byte compressed point" % i)
        inputs.append(row)
    key = (bytes(crs_bytes), ell, nbl)
    prover = _PROVER_CACHE.get(key)
    if prover is None:
        prover = _PROVER_CACHE[key] = BatchProver(crs_bytes, ell, nbl)
    out = []
    for tu, proof in prover.prove_drawn(inputs, rng):
        T, U = tu[:48 * ell], tu[48 * ell:]
        post = [WhiskTracker(T[48 * i:48 * i + 48], U[48 * i:48 * i + 48]) for i in range(ell)]
        out.append((post, proof))
    return out

// CPython-exact randomness for the batched prover.
// The reference draws the permutation, k and every blinder from Python's `random` module
// (cp/util.py:21-24 random_scalar = Scalar(random.randint(1, CURVE_ORDER - 1)); whisk_interface.py:114-116
// random.shuffle + random_scalar), ~3n + 14 draws per proof: in pure Python that is 0.4 ms per proof, a third
// of the GPU's time per proof.  This restates CPython 3's generator on the caller's own state
// (random.getstate()[1]: 624 words + index), so the stream continues exactly where Python left it and
// random.setstate() of the returned words resumes it: under random.seed(s) the batched prover still emits the
// reference's proof bytes.  Restated from the published algorithms: MT19937 (Matsumoto & Nishimura, genrand_int32),
// and CPython's Random.getrandbits / _randbelow_with_getrandbits / randint / shuffle (Lib/random.py, _randommodule.c).
namespace {

struct PyMT {
    uint32_t* mt;                 // 624 state words
    uint32_t idx;
    uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; k++) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    uint32_t bits(uint32_t k) { return next() >> (32 - k); }                 // getrandbits(k), 1 <= k <= 32
    uint32_t below(uint32_t n) {                                              // _randbelow(n), 1 <= n < 2^32
        uint32_t k = 0;
        while ((n >> k) != 0) k++;
        uint32_t r = bits(k);
        while (r >= n) r = bits(k);
        return r;
    }
    // randint(1, r - 1) = 1 + _randbelow(r - 1): getrandbits(255) fills 32-bit words from the least significant
    // one, the last word keeps its top 31 bits; rejected while >= r - 1
    void scalar(uint8_t* out32) {
        static const uint32_t RM1[8] = {0x00000000u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        uint32_t w[8];
        for (;;) {
            for (int i = 0; i < 7; i++) w[i] = next();
            w[7] = next() >> 1;
            bool lt = false;
            for (int i = 7; i >= 0; i--) if (w[i] != RM1[i]) { lt = w[i] < RM1[i]; break; }
            if (lt) break;
        }
        uint64_t carry = 1;                                                   // + 1
        for (int i = 0; i < 8; i++) { uint64_t v = (uint64_t)w[i] + carry; w[i] = (uint32_t)v; carry = v >> 32; }
        for (int i = 0; i < 8; i++) { out32[4 * i] = (uint8_t)w[i]; out32[4 * i + 1] = (uint8_t)(w[i] >> 8); out32[4 * i + 2] = (uint8_t)(w[i] >> 16); out32[4 * i + 3] = (uint8_t)(w[i] >> 24); }
    }
};

}  // namespace

extern "C" {

/* For B proofs in the reference's order (cp/whisk_interface.py:114-116, then CurdleProofsProof.new's draws,
 * SURVEY A.4):  perm = list(range(ell)); random.shuffle(perm); k = random_scalar(); n_rand x random_scalar().
 * state: the 625 words of random.getstate()[1] (version 3), advanced in place.  Pure host code: needs no device. */
int cpg_pyrandom_draw_shuffles(uint32_t* state625, size_t ell, size_t n_rand, size_t B, uint32_t* perms, uint8_t* ks, uint8_t* rand) {
    if (!state625 || !perms || !ks || !rand) return fail("cpg_pyrandom_draw_shuffles: null argument");
    if (state625[624] > 624 || ell == 0 || ell >= 0x7fffffffULL) return fail("cpg_pyrandom_draw_shuffles: bad state or size");
    PyMT g{state625, state625[624]};
    for (size_t b = 0; b < B; b++) {
        uint32_t* p = perms + b * ell;
        for (size_t i = 0; i < ell; i++) p[i] = (uint32_t)i;
        for (size_t i = ell - 1; i >= 1; i--) {                               // Random.shuffle
            uint32_t j = g.below((uint32_t)i + 1);
            uint32_t t = p[i]; p[i] = p[j]; p[j] = t;
        }
        g.scalar(ks + 32 * b);
        for (size_t r = 0; r < n_rand; r++) g.scalar(rand + (b * n_rand + r) * 32);
    }
    state625[624] = g.idx;
    return 0;
}

}  // extern "C"

// Test seam: compiles the device arithmetic headers for the HOST (carry flag emulated in
// bigint.cuh) so the exact algorithm code the kernels run can be checked on a CPU-only box.
// Built by tests/conftest.py into tests/_build/libhostseam.so; never loaded by the product.
#include "../../curdleproofs_pie_b200/csrc/field.cuh"
#include "../../curdleproofs_pie_b200/csrc/g1.cuh"
#include <string.h>
using namespace cpg;

extern "C" {
void hs_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.l, a, 48); memcpy(y.l, b, 48); Fq z = mul(x, y); memcpy(r, z.l, 48); }
void hs_fq_sqr(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = sqr(x); memcpy(r, z.l, 48); }
void hs_fr_sqr(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.l, a, 32); Fr z = sqr(x); memcpy(r, z.l, 32); }
void hs_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.l, a, 48); memcpy(y.l, b, 48); Fq z = add(x, y); memcpy(r, z.l, 48); }
void hs_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.l, a, 48); memcpy(y.l, b, 48); Fq z = sub(x, y); memcpy(r, z.l, 48); }
void hs_fq_inv(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = fq_inv(x); memcpy(r, z.l, 48); }
void hs_fq_inv_fermat(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = fq_inv_fermat(x); memcpy(r, z.l, 48); }
void hs_fq_sqrt(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = fq_sqrt_candidate(x); memcpy(r, z.l, 48); }
// lazy chain forms (operands and results in [0, 2p), no final subtraction) and the single reduction that ends a chain
void hs_fq_mul_lazy(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.l, a, 48); memcpy(y.l, b, 48); Fq z = mul_lz<true>(x, y); memcpy(r, z.l, 48); }
void hs_fq_sqr_lazy(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = sqr_n_lz<true>(x, 1); memcpy(r, z.l, 48); }
void hs_fq_reduce_once(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = reduce_once(x); memcpy(r, z.l, 48); }
void hs_fq_to_mont(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = to_mont(x); memcpy(r, z.l, 48); }
void hs_fq_from_mont(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); Fq z = from_mont(x); memcpy(r, z.l, 48); }
int hs_fq_lex_largest(const uint32_t* a) { Fq x; memcpy(x.l, a, 48); return fq_is_lex_largest(x); }
void hs_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = mul(x, y); memcpy(r, z.l, 32); }
void hs_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = add(x, y); memcpy(r, z.l, 32); }
void hs_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = sub(x, y); memcpy(r, z.l, 32); }
void hs_fr_inv(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.l, a, 32); Fr z = fr_inv(x); memcpy(r, z.l, 32); }
}

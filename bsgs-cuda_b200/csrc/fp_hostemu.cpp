// fp_hostemu.cpp -- exports fp.cuh compiled for the HOST (PTX primitives emulated) so that the limb
// arithmetic can be unit-tested on a machine without a GPU.  Test-only; not linked into the product.
#include "fp.cuh"
extern "C" {
// number of carries that left a chain the code declares carry-free (fresh top slot); must stay 0
uint64_t he_lost_carries(void) { return fp::g_lost_carries; }
uint64_t he_rare_calls(void) { return fp::g_rare_calls; }
void he_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { fp::fe_mul(r, a, b); }
void he_sqr(uint32_t* r, const uint32_t* a) { fp::fe_sqr(r, a); }
void he_inv(uint32_t* r, const uint32_t* a) { fp::fe_inv(r, a); }
void he_sub_canon(uint32_t* r, const uint32_t* a, const uint32_t* b) { fp::fe_sub_canon(r, a, b); }
void he_add_lazy(uint32_t* r, const uint32_t* a, const uint32_t* b) { fp::fe_add_lazy(r, a, b); }
void he_canon(uint32_t* r) { fp::fe_canon(r); }
void he_reduce512(uint32_t* r, const uint32_t* v16) { uint32_t v[16]; for (int i = 0; i < 16; i++) v[i] = v16[i]; fp::reduce512(r, v); }
uint64_t he_add_low64_canon(const uint32_t* a, const uint32_t* u) { return fp::fe_add_low64_canon(a, u); }
// returns the key; *exact_needed = 1 when the caller must take the exact path
uint64_t he_key_low64(const uint32_t* L, const uint32_t* A, const uint32_t* Q, int* exact_needed) {
    const uint64_t a_lo = ((uint64_t)A[1] << 32) | A[0], q_lo = ((uint64_t)Q[1] << 32) | Q[0];
    bool ex = false;
    uint64_t k = fp::fe_key_low64(L, a_lo - q_lo, (int64_t)A[7] - (int64_t)Q[7], ex);
    *exact_needed = ex ? 1 : 0;
    return k;
}
// key from an unreduced 512-bit v: returns the key; *exact_needed as for he_key_low64
uint64_t he_key_from_wide(const uint32_t* v16, const uint32_t* A, const uint32_t* Q, uint32_t thr, int* exact_needed) {
    const uint64_t a_lo = ((uint64_t)A[1] << 32) | A[0], q_lo = ((uint64_t)Q[1] << 32) | Q[0];
    bool ex = false;
    uint64_t k = fp::fe_key_from_wide(v16, a_lo - q_lo, (int64_t)A[7] - (int64_t)Q[7], thr, ex);
    *exact_needed = ex ? 1 : 0;
    return k;
}
void he_sqr_wide(uint32_t* v16, const uint32_t* a) { fp::fe_sqr_wide(v16, a); }
}

/* TEST INFRASTRUCTURE — the checker, never the product path.  See tfhe_oracle.h.
 *
 * Exact-integer restatement of TFHEpp's gate bootstrap (128-bit parameters).
 * TFHEPP/ = /root/reference/thirdparty/cuFHE/thirdparties/TFHEpp/
 * All torus arithmetic wraps modulo the word size exactly as the reference's
 * unsigned C++ types do; the only deliberate difference from the reference is
 * that polynomial products are exact negacyclic convolutions mod 2^32 instead
 * of a double-precision FFT (TFHEPP/include/mulfft.hpp:69-134), i.e. this is the
 * value the reference approximates to within its FFT rounding error.
 */
#include "tfhe_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define N0 ORC_N0
#define N1 ORC_N1
#define L ORC_L
#define BGBIT ORC_BGBIT
#define BG (1u << BGBIT)
#define ROWS ORC_ROWS

/* ------------------------------------------------------------------ */
/* deterministic integer-only randomness (NOT the reference's Randen) */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r)
{ /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline rng_t rng_fork(uint64_t seed, uint64_t stream)
{
    rng_t r = { seed ^ (0xD1B54A32D192ED03ull * (stream + 1)) };
    rng_next(&r);
    rng_t out = { rng_next(&r) };
    return out;
}
/* Irwin-Hall(12) approximation of a centred Gaussian with stdev sigma_q16/65536
 * torus units; integer arithmetic only so keys are reproducible on any host. */
static inline int32_t rng_gauss(rng_t *r, int64_t sigma_q16)
{
    int64_t s = 0;
    for (int i = 0; i < 6; i++) {
        uint64_t v = rng_next(r);
        s += (int64_t)(uint32_t)v + (int64_t)(uint32_t)(v >> 32);
    }
    s -= 6ll << 32;                                   /* stdev(s) = 2^32 */
    return (int32_t)(((__int128)s * sigma_q16 + ((__int128)1 << 47)) >> 48);
}
/* lvl0param::alpha * 2^w0 and lvl1param::alpha * 2^32 in Q16 (params/128bit.hpp:16,38; params/CGGI16.hpp:15,35) */
#ifdef ORC_80BIT
#define SIGMA0_Q16 ((int64_t)(2.44e-5 * 4294967296.0 * 65536.0 + 0.5))
#define SIGMA1_Q16 ((int64_t)(3.73e-9 * 4294967296.0 * 65536.0 + 0.5))
#else
#define SIGMA0_Q16 ((int64_t)(0.0000925119974676756 * 65536.0 * 65536.0 + 0.5))
#define SIGMA1_Q16 ((int64_t)(128.0 * 65536.0))
#endif

/* ------------------------------------------------------------------ */
/* primitives                                                         */
/* ------------------------------------------------------------------ */

/* TFHEPP/include/utils.hpp:113-128  PolynomialMulByXai, a in [0, 2N] */
void orc_mul_xai(const uint32_t *poly, uint32_t a, uint32_t *out)
{
    if (a == 0) {
        memcpy(out, poly, N1 * 4);
    } else if (a < N1) {
        for (uint32_t i = 0; i < a; i++) out[i] = -poly[i - a + N1];
        for (uint32_t i = a; i < N1; i++) out[i] = poly[i - a];
    } else {
        uint32_t aa = a - N1;
        for (uint32_t i = 0; i < aa; i++) out[i] = poly[i - aa + N1];
        for (uint32_t i = aa; i < N1; i++) out[i] = -poly[i - aa];
    }
}

/* TFHEPP/include/utils.hpp:130-144  PolynomialMulByXaiMinusOne, a in [0, 2N] */
void orc_mul_xai_minus_one(const uint32_t *poly, uint32_t a, uint32_t *out)
{
    if (a < N1) {
        for (uint32_t i = 0; i < a; i++) out[i] = -poly[i - a + N1] - poly[i];
        for (uint32_t i = a; i < N1; i++) out[i] = poly[i - a] - poly[i];
    } else {
        uint32_t aa = a - N1;
        for (uint32_t i = 0; i < aa; i++) out[i] = poly[i - aa + N1] - poly[i];
        for (uint32_t i = aa; i < N1; i++) out[i] = -poly[i - aa] - poly[i];
    }
}

/* TFHEPP/include/trgsw.hpp:12-21 offsetgen, :62-78 Decomposition (non-optimal branch) */
void orc_decompose(const uint32_t *poly, int32_t *digits)
{
    uint32_t offset = 0;
    for (int i = 1; i <= L; i++) offset += (BG / 2) * (1u << (32 - i * BGBIT));
    const uint32_t roundoffset = 1u << (32 - L * BGBIT - 1);
    const uint32_t mask = BG - 1, half = BG / 2;
    for (int i = 0; i < N1; i++) {
        uint32_t v = poly[i] + offset + roundoffset;
        for (int l = 0; l < L; l++)
            digits[l * N1 + i] = (int32_t)((v >> (32 - (l + 1) * BGBIT)) & mask) - (int32_t)half;
    }
}

/* acc += d (*) b in Z_{2^32}[X]/(X^N+1), exact (uint32 wrap-around is the ring).
 * ext[m] = -b[m] (m<N), b[m-N] (m>=N)  =>  acc[k] += sum_i d[i]*ext[N+k-i]. */
void orc_negacyclic_mul(const int32_t *d, const uint32_t *b, uint32_t *acc)
{
    uint32_t ext[2 * N1];
    for (int j = 0; j < N1; j++) { ext[j] = -b[j]; ext[N1 + j] = b[j]; }
    for (int i = 0; i < N1; i++) {
        const uint32_t di = (uint32_t)d[i];
        if (di == 0) continue;
        const uint32_t *e = ext + N1 - i;
        for (int k = 0; k < N1; k++) acc[k] += di * e[k];
    }
}

/* TFHEPP/include/trgsw.hpp:102-131 trgswfftExternalProduct, with exact products.
 * Row order: digits of polynomial A (rows 0..l-1) then of B (rows l..2l-1). */
void orc_external_product(const uint32_t *trlwe, const uint32_t *trgsw, uint32_t *out)
{
    int32_t dig[L * N1];
    uint32_t res[2 * N1];
    memset(res, 0, sizeof(res));
    for (int p = 0; p < 2; p++) {
        orc_decompose(trlwe + p * N1, dig);
        for (int l = 0; l < L; l++) {
            const uint32_t *row = trgsw + (size_t)(p * L + l) * 2 * N1;
            orc_negacyclic_mul(dig + l * N1, row, res);
            orc_negacyclic_mul(dig + l * N1, row + N1, res + N1);
        }
    }
    memcpy(out, res, sizeof(res));
}

/* TFHEPP/include/detwfa.hpp:36-49 CMUXFFTwithPolynomialMulByXaiMinusOne (key_value_diff==1) */
void orc_cmux_step(uint32_t *acc, const uint32_t *trgsw, uint32_t abar)
{
    uint32_t tmp[2 * N1], prod[2 * N1];
    orc_mul_xai_minus_one(acc, abar, tmp);
    orc_mul_xai_minus_one(acc + N1, abar, tmp + N1);
    orc_external_product(tmp, trgsw, prod);
    for (int i = 0; i < 2 * N1; i++) acc[i] += prod[i];
}

/* TFHEPP/include/gatebootstrapping.hpp:26-30 (b, no rounding) and :58-65 (a, rounded).
 * uint16 operands are promoted to int in the reference, so abar may be 2N (== 0 mod 2N); with the 32-bit lvl0 torus
 * of the 80-bit set the sum is formed in uint32 and wraps, so abar stays below 2N. */
void orc_mod_switch(const orc_t0 *c, uint32_t *abar, uint32_t *bbar)
{
    const int shift = ORC_T0_BITS - 1 - ORC_NBIT;          /* digits - 1 - nbit = 5 (21) */
    *bbar = 2 * N1 - ((uint32_t)c[N0] >> shift);
    for (int i = 0; i < N0; i++) abar[i] = (uint32_t)((uint32_t)c[i] + (1u << (shift - 1))) >> shift;
}

/* TFHEPP/include/gatebootstrapping.hpp:19-71 BlindRotate with testvector mu_polygen (:233-239) */
void orc_blind_rotate(const orc_t0 *c, const uint32_t *bk, uint32_t *acc)
{
    uint32_t abar[N0], bbar, tv[N1];
    orc_mod_switch(c, abar, &bbar);
    for (int i = 0; i < N1; i++) tv[i] = ORC_MU1;
    memset(acc, 0, N1 * 4);
    orc_mul_xai(tv, bbar, acc + N1);
    for (int i = 0; i < N0; i++) {
        if (abar[i] == 0) continue;                /* :66 */
        orc_cmux_step(acc, bk + (size_t)i * ROWS * 2 * N1, abar[i]);
    }
}

/* TFHEPP/include/trlwe.hpp:213-223 SampleExtractIndex(index = 0) */
void orc_sample_extract0(const uint32_t *acc, uint32_t *tlwe1)
{
    tlwe1[0] = acc[0];
    for (int i = 1; i < N1; i++) tlwe1[i] = -acc[N1 - i];
    tlwe1[N1] = acc[N1];
}

/* TFHEPP/include/keyswitch.hpp:11-52 IdentityKeySwitch<lvl10param> */
void orc_keyswitch(const uint32_t *tlwe1, const orc_t0 *ksk, orc_t0 *out)
{
    const uint32_t prec_offset = 1u << (32 - (1 + ORC_BASEBIT * ORC_T));
    const uint32_t mask = (1u << ORC_BASEBIT) - 1;
    orc_t0 res[ORC_TLWE0];
    memset(res, 0, sizeof(res));
#if ORC_T0_BITS == 32
    res[N0] = tlwe1[N1];                                             /* same width: plain copy, :27-29 */
#else
    res[N0] = (orc_t0)((tlwe1[N1] + (1u << 15)) >> 16);              /* 32 -> 16 bit rounding, :30-34 */
#endif
    for (int i = 0; i < N1; i++) {
        const uint32_t aibar = tlwe1[i] + prec_offset;
        for (int j = 0; j < ORC_T; j++) {
            const uint32_t aij = (aibar >> (32 - (j + 1) * ORC_BASEBIT)) & mask;
            if (aij == 0) continue;
            const orc_t0 *row = ksk + (((size_t)i * ORC_T + j) * 3 + (aij - 1)) * ORC_TLWE0;
            for (int k = 0; k <= N0; k++) res[k] -= row[k];
        }
    }
    memcpy(out, res, sizeof(res));
}

void orc_bootstrap_to_lvl1(const orc_t0 *c, const uint32_t *bk, uint32_t *tlwe1)
{
    uint32_t acc[2 * N1];
    orc_blind_rotate(c, bk, acc);
    orc_sample_extract0(acc, tlwe1);
}

/* ------------------------------------------------------------------ */
/* gates: TFHEPP/include/gate.hpp                                     */
/* ------------------------------------------------------------------ */
typedef struct { int sa, sb, off; } gate_coef;   /* off in units of mu0 */
static int gate_table(uint8_t op, gate_coef *g)
{
    switch (op) {
    case ORC_NAND:   *g = (gate_coef){-1, -1, +1}; return 1;  /* gate.hpp:65  */
    case ORC_NOR:    *g = (gate_coef){-1, -1, -1}; return 1;  /* :82  */
    case ORC_XNOR:   *g = (gate_coef){-2, -2, -2}; return 1;  /* :99  */
    case ORC_AND:    *g = (gate_coef){+1, +1, -1}; return 1;  /* :116 */
    case ORC_OR:     *g = (gate_coef){+1, +1, +1}; return 1;  /* :133 */
    case ORC_XOR:    *g = (gate_coef){+2, +2, +2}; return 1;  /* :150 */
    case ORC_ANDNY:  *g = (gate_coef){-1, +1, -1}; return 1;  /* :167 */
    case ORC_ANDNOT: *g = (gate_coef){+1, -1, -1}; return 1;  /* :184 HomANDYN (iyokan_tfhepp.hpp:133) */
    case ORC_ORNY:   *g = (gate_coef){-1, +1, +1}; return 1;  /* :201 */
    case ORC_ORNOT:  *g = (gate_coef){+1, -1, +1}; return 1;  /* :218 HomORYN (iyokan_tfhepp.hpp:136) */
    default: return 0;
    }
}

int orc_num_bootstraps(uint8_t op)
{
    gate_coef g;
    if (op == ORC_MUX) return 2;
    return gate_table(op, &g) ? 1 : 0;
}

void orc_gate(uint8_t op, const orc_t0 *in0, const orc_t0 *in1, const orc_t0 *in2,
              orc_t0 *out, const uint32_t *bk, const orc_t0 *ksk)
{
    gate_coef g;
    orc_t0 c[ORC_TLWE0];
    uint32_t u[ORC_TLWE1];
    if (gate_table(op, &g)) {
        /* HomGate, gate.hpp:8-18 (uint16 wrap) */
        for (int i = 0; i <= N0; i++) c[i] = (orc_t0)((uint32_t)g.sa * in0[i] + (uint32_t)g.sb * in1[i]);
        c[N0] = (orc_t0)(c[N0] + (uint32_t)g.off * ORC_MU0);
        orc_bootstrap_to_lvl1(c, bk, u);
        orc_keyswitch(u, ksk, out);
    } else if (op == ORC_MUX) {
        /* HomMUX<lvl0param>, gate.hpp:231-262 with cs=in2, c1=in1, c0=in0 */
        orc_t0 c0[ORC_TLWE0];
        uint32_t u0[ORC_TLWE1];
        for (int i = 0; i <= N0; i++) c[i] = (orc_t0)(in2[i] + in1[i]);
        for (int i = 0; i <= N0; i++) c0[i] = (orc_t0)(0u - in2[i] + in0[i]);
        c[N0] -= ORC_MU0;
        c0[N0] -= ORC_MU0;
        orc_bootstrap_to_lvl1(c, bk, u);
        orc_bootstrap_to_lvl1(c0, bk, u0);
        for (int i = 0; i <= N1; i++) u0[i] += u[i];
        orc_keyswitch(u0, ksk, out);
        out[N0] += ORC_MU0;
    } else if (op == ORC_NOT) {                                   /* gate.hpp:47-51 */
        for (int i = 0; i <= N0; i++) out[i] = (orc_t0)(0u - in0[i]);
    } else if (op == ORC_COPY) {                                  /* gate.hpp:53-57 */
        memmove(out, in0, ORC_TLWE0 * sizeof(orc_t0));
    } else if (op == ORC_CONST1 || op == ORC_CONST0) {            /* gate.hpp:32-44 */
        memset(out, 0, ORC_TLWE0 * sizeof(orc_t0));
        out[N0] = (op == ORC_CONST1) ? (orc_t0)ORC_MU0 : (orc_t0)(0u - ORC_MU0);
    } else {
        abort();
    }
}

void orc_gate_batch(const uint8_t *ops, const orc_t0 *in0, const orc_t0 *in1, const orc_t0 *in2,
                    orc_t0 *out, size_t count, const uint32_t *bk, const orc_t0 *ksk, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (long g = 0; g < (long)count; g++) {
        const size_t o = (size_t)g * ORC_TLWE0;
        orc_gate(ops[g], in0 ? in0 + o : NULL, in1 ? in1 + o : NULL, in2 ? in2 + o : NULL,
                 out + o, bk, ksk);
    }
}

/* ------------------------------------------------------------------ */
/* keys / encryption: same distributions and layouts as the reference */
/* ------------------------------------------------------------------ */

/* tlweSymEncrypt<lvl0param>, TFHEPP/include/tlwe.hpp:12-25 */
static void tlwe0_encrypt(rng_t *r, orc_t0 m, const orc_t0 *sk0, orc_t0 *out)
{
    orc_t0 b = (orc_t0)(m + (orc_t0)rng_gauss(r, SIGMA0_Q16));
    const int per = 64 / ORC_T0_BITS;              /* uniform a_i: 4 (2) torus words per 64-bit draw */
    for (int i = 0; i < N0; i += per) {
        uint64_t v = rng_next(r);
        for (int j = 0; j < per && i + j < N0; j++) {
            orc_t0 a = (orc_t0)(v >> (ORC_T0_BITS * j));
            out[i + j] = a;
            b = (orc_t0)(b + a * sk0[i + j]);
        }
    }
    out[N0] = b;
}

void orc_encrypt_bits(uint64_t seed, const orc_t0 *sk0, const uint8_t *bits, size_t count,
                      orc_t0 *out)
{
#pragma omp parallel for
    for (long g = 0; g < (long)count; g++) {
        rng_t r = rng_fork(seed, (uint64_t)g);
        /* bootsSymEncrypt, tlwe.hpp:101-110: message = bit ? mu : -mu */
        tlwe0_encrypt(&r, bits[g] ? (orc_t0)ORC_MU0 : (orc_t0)(0u - ORC_MU0), sk0,
                      out + (size_t)g * ORC_TLWE0);
    }
}

/* tlweSymDecrypt, tlwe.hpp:77-87: bit = (signed phase > 0) */
void orc_phase(const orc_t0 *sk0, const orc_t0 *in, size_t count, orc_s0 *phase)
{
    for (size_t g = 0; g < count; g++) {
        const orc_t0 *c = in + g * ORC_TLWE0;
        orc_t0 ph = c[N0];
        for (int i = 0; i < N0; i++) ph = (orc_t0)(ph - c[i] * sk0[i]);
        phase[g] = (orc_s0)ph;
    }
}

void orc_decrypt_bits(const orc_t0 *sk0, const orc_t0 *in, size_t count, uint8_t *bits)
{
    for (size_t g = 0; g < count; g++) {
        orc_s0 ph;
        orc_phase(sk0, in + g * ORC_TLWE0, 1, &ph);
        bits[g] = ph > 0;
    }
}

void orc_phase1(const int32_t *sk1, const uint32_t *tlwe1, size_t count, int32_t *phase)
{
    for (size_t g = 0; g < count; g++) {
        const uint32_t *c = tlwe1 + g * ORC_TLWE1;
        uint32_t ph = c[N1];
        for (int i = 0; i < N1; i++) ph -= c[i] * (uint32_t)sk1[i];
        phase[g] = (int32_t)ph;
    }
}

/* keys: lweKey ctor (TFHEPP/src/key.cpp:5-16; lvl0 binary, lvl1 ternary),
 * bkgen<lvl01param> (cloudkey.hpp:18-52) -> trgswSymEncrypt (trgsw.hpp:301-318, hgen :231-244)
 *   -> trlweSymEncryptZero (trlwe.hpp:7-24), ikskgen<lvl10param> (cloudkey.hpp:188-203). */
void orc_keygen(uint64_t seed, orc_t0 *sk0, int32_t *sk1, uint32_t *bk, orc_t0 *ksk)
{
    rng_t r = rng_fork(seed, 0xA11CE);
    for (int i = 0; i < N0; i++) sk0[i] = (orc_t0)(rng_next(&r) & 1);
    for (int i = 0; i < N1; i++) sk1[i] = (int32_t)(rng_next(&r) % 3) - 1;

#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < N0; i++) {
        rng_t ri = rng_fork(seed, 0x10000u + (uint64_t)i);
        for (int row = 0; row < ROWS; row++) {
            uint32_t *a = bk + ((size_t)i * ROWS + row) * 2 * N1;
            uint32_t *b = a + N1;
            for (int j = 0; j < N1; j += 2) {
                uint64_t v = rng_next(&ri);
                a[j] = (uint32_t)v;
                a[j + 1] = (uint32_t)(v >> 32);
            }
            for (int j = 0; j < N1; j++) b[j] = (uint32_t)rng_gauss(&ri, SIGMA1_Q16);
            /* b += a (*) s1, negacyclic, ternary key */
            for (int k = 0; k < N1; k++) {
                int32_t s = sk1[k];
                if (s == 0) continue;
                uint32_t us = (uint32_t)s;
                for (int j = 0; j < N1 - k; j++) b[j + k] += us * a[j];
                for (int j = N1 - k; j < N1; j++) b[j + k - N1] -= us * a[j];
            }
            /* gadget: trgsw[l + p*L][p][0] += [s0_i == 1] * 2^(32-(l+1)*Bgbit) */
            const int p = row / L, l = row % L;
            if (sk0[i] == 1) (p == 0 ? a : b)[0] += 1u << (32 - (l + 1) * BGBIT);
        }
    }

#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < N1; i++) {
        rng_t ri = rng_fork(seed, 0x20000u + (uint64_t)i);
        for (int j = 0; j < ORC_T; j++)
            for (uint32_t k = 0; k < 3; k++) {
                /* domainkey[i]*(k+1)*2^(w0-(j+1)*basebit), truncated to the lvl0 torus */
                orc_t0 m = (orc_t0)((uint32_t)sk1[i] * (k + 1) * (1u << (ORC_T0_BITS - (j + 1) * ORC_BASEBIT)));
                tlwe0_encrypt(&ri, m, sk0, ksk + (((size_t)i * ORC_T + j) * 3 + k) * ORC_TLWE0);
            }
    }
}

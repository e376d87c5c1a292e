/* TEST INFRASTRUCTURE — the checker, never the product path.
 *
 * Exact-integer CPU restatement of the reference's TFHE gate-bootstrap path
 * (TFHEpp, vendored at thirdparty/cuFHE/thirdparties/TFHEpp; 128-bit and 80-bit parameter sets).
 * Every polynomial product is an exact negacyclic convolution modulo 2^32, so the
 * result is the mathematically canonical ciphertext the reference's floating-point
 * FFT approximates.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks it against vectors the
 * unmodified reference produced in this container (tests/golden/, generator script
 * tests/golden/make_golden.py): IdentityKeySwitch bit-exact, Decomposition /
 * PolynomialMulByXaiMinusOne bit-exact, BlindRotate within the reference's own FFT
 * rounding error, every gate's decrypted bit identical.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Two compile-time flavours, like the reference (USE_80BIT_SECURITY, TFHEpp include/params.hpp:14-28):
 * libtfhe_oracle.so (default, 128-bit) and libtfhe_oracle80.so (-DORC_80BIT). */
#ifdef ORC_80BIT
/* 80-bit parameter set: TFHEpp include/params/CGGI16.hpp:6-68 */
typedef uint32_t orc_t0;     /* lvl0param::T */
typedef int32_t orc_s0;
#define ORC_T0_BITS 32
#define ORC_N0 500      /* lvl0 dimension n */
#define ORC_L 2         /* gadget length l */
#define ORC_BGBIT 10
#define ORC_T 8         /* key-switch digits t */
#define ORC_MU0 (1u << 29)   /* lvl0param::mu, uint32 torus */
#else
/* 128-bit parameter set: TFHEpp include/params/128bit.hpp:7-76 */
typedef uint16_t orc_t0;
typedef int16_t orc_s0;
#define ORC_T0_BITS 16
#define ORC_N0 636      /* lvl0 dimension n */
#define ORC_L 3         /* gadget length l */
#define ORC_BGBIT 6
#define ORC_T 7         /* key-switch digits t */
#define ORC_MU0 (1u << 13)   /* lvl0param::mu, uint16 torus */
#endif
#define ORC_N1 1024     /* lvl1 ring degree N */
#define ORC_NBIT 10
#define ORC_BASEBIT 2
#define ORC_MU1 (1u << 29)   /* lvl1param::mu, uint32 torus */
#define ORC_TLWE0 (ORC_N0 + 1)
#define ORC_TLWE1 (ORC_N1 + 1)
#define ORC_ROWS (2 * ORC_L) /* (k+1)*l TRGSW rows */

/* opcode numbering shared with include/b200fhe.h */
enum {
    ORC_AND = 0, ORC_NAND, ORC_ANDNOT, ORC_OR, ORC_NOR, ORC_ORNOT, ORC_XOR, ORC_XNOR,
    ORC_MUX, ORC_NOT, ORC_COPY, ORC_CONST0, ORC_CONST1, ORC_ANDNY, ORC_ORNY, ORC_NUM_OPS
};

/* ---- keys, encryption (deterministic, integer-only sampler; not the reference RNG) ---- */
void orc_keygen(uint64_t seed, orc_t0 *sk0 /*[636]*/, int32_t *sk1 /*[1024]*/,
                uint32_t *bk /*[636][6][2][1024]*/, orc_t0 *ksk /*[1024][7][3][637]*/);
void orc_encrypt_bits(uint64_t seed, const orc_t0 *sk0, const uint8_t *bits, size_t count,
                      orc_t0 *out /*[count][637]*/);
void orc_decrypt_bits(const orc_t0 *sk0, const orc_t0 *in, size_t count, uint8_t *bits);
void orc_phase(const orc_t0 *sk0, const orc_t0 *in, size_t count, orc_s0 *phase);
void orc_phase1(const int32_t *sk1, const uint32_t *tlwe1, size_t count, int32_t *phase);

/* ---- primitives (each cites the reference lines it restates in the .c file) ---- */
void orc_decompose(const uint32_t *poly /*[1024]*/, int32_t *digits /*[3][1024]*/);
void orc_mul_xai(const uint32_t *poly, uint32_t a, uint32_t *out);
void orc_mul_xai_minus_one(const uint32_t *poly, uint32_t a, uint32_t *out);
void orc_negacyclic_mul(const int32_t *d, const uint32_t *b, uint32_t *acc /* += d*b */);
void orc_external_product(const uint32_t *trlwe /*[2][1024]*/, const uint32_t *trgsw /*[6][2][1024]*/,
                          uint32_t *out /*[2][1024]*/);
void orc_cmux_step(uint32_t *acc /*[2][1024] in/out*/, const uint32_t *trgsw, uint32_t abar);
void orc_mod_switch(const orc_t0 *c /*[637]*/, uint32_t *abar /*[636]*/, uint32_t *bbar);
void orc_blind_rotate(const orc_t0 *c /*[637]*/, const uint32_t *bk, uint32_t *acc /*[2][1024]*/);
void orc_sample_extract0(const uint32_t *acc, uint32_t *tlwe1 /*[1025]*/);
void orc_keyswitch(const uint32_t *tlwe1 /*[1025]*/, const orc_t0 *ksk, orc_t0 *out /*[637]*/);

/* ---- gates ---- */
/* in0/in1/in2 follow Iyokan's input(0..2) order (MUX: in2 ? in1 : in0). Unused may be NULL. */
void orc_gate(uint8_t op, const orc_t0 *in0, const orc_t0 *in1, const orc_t0 *in2,
              orc_t0 *out, const uint32_t *bk, const orc_t0 *ksk);
/* dense arrays [count][637]; OpenMP over gates; nthreads<=0 -> all cores */
void orc_gate_batch(const uint8_t *ops, const orc_t0 *in0, const orc_t0 *in1, const orc_t0 *in2,
                    orc_t0 *out, size_t count, const uint32_t *bk, const orc_t0 *ksk, int nthreads);
/* blind rotation + sample extract only (lvl1 TLWE out) for the linear combination
 * sa*in0 + sb*in1 + off (used to pin against GateBootstrappingTLWE2TLWEFFT) */
void orc_bootstrap_to_lvl1(const orc_t0 *c /*[637]*/, const uint32_t *bk, uint32_t *tlwe1);
int orc_num_bootstraps(uint8_t op);

#ifdef __cplusplus
}
#endif
#endif

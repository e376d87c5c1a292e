// 32-bit modular arithmetic for the exact negacyclic NTT.
//
// Why one 29-bit prime is enough for an EXACT product modulo 2^32:
// an external product sums (k+1)*l = 6 negacyclic products of a signed digit polynomial
// (|d| <= Bg/2 = 32) with a bootstrapping-key polynomial.  The key coefficients are split
// into three centred limbs bk = x0 + 2^11*x1 + 2^22*x2 (|x0|,|x1| <= 1024, |x2| <= 512), so
// every per-limb integer result is bounded by 6*1024*32*1024 = 201,326,592 < p/2 and is
// recovered exactly from its residue mod p; the three results are recombined mod 2^32.
// (SURVEY.md Appendix A derives the 2^48.6 bound that forces either this split or a >49-bit modulus.)
// 80-bit flavour (l = 2, Bg = 2^10): 4 rows x 1024 x 512 x 64 = 2^27 < p/2 with FIVE limbs of 7,7,6,6,6 bits.
//
// p = 2^29 - 14335 = 536856577, p = 1 (mod 2048).  Values are kept lazily in [0, 8p) (8p < 2^32):
//   shoup_mul : any 32-bit y  -> y*w mod p in [0, 2p)      (Harvey / Shoup, 3 integer multiplies)
//   fix29     : any 32-bit x  -> x mod p   in [0, p + 8c)  (c = 2^29 - p), one shift + one multiply-add
#pragma once
#include "fhe_params.h"
#include "hd.h"

namespace b200 {

constexpr uint32_t P = 536856577u;
constexpr uint32_t P2 = 2u * P;
constexpr uint32_t P4 = 4u * P;
constexpr uint32_t PC = (1u << 29) - P;      // 14335
constexpr uint32_t PINVNEG = 331335679u;     // -p^{-1} mod 2^32
constexpr uint32_t PSI = 127625803u;         // primitive 2048-th root of unity mod p (5^((p-1)/2048))
constexpr uint32_t CONV_BOUND = (uint32_t)ROWS * 1024u * (1u << (BGBIT - 1)) * (1u << (limb_width(0) - 1));  // widest limb first
static_assert((uint64_t)P * PINVNEG % (1ull << 32) == (1ull << 32) - 1, "PINVNEG");
static_assert(CONV_BOUND < P / 2, "exactness bound");
static_assert((1u << 28) + 8u * PC + CONV_BOUND < P, "centred lift is unambiguous");

B200_HD uint32_t shoup_mul(uint32_t y, tw_t t)
{
    const uint32_t q = mulhi32(y, t.ws);
    return y * t.w - q * P;  // in [0, 2p)
}

B200_HD uint32_t fix29(uint32_t x) { return x - (x >> 29) * P; }  // < p + 8c

// ALU-only conditional subtractions (unsigned-min trick: x - k wraps above x when x < k).
// The integer-multiply pipe is the kernel's bottleneck, so range fixes run on the ALU pipe.
B200_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
B200_HD uint32_t fix_lt8p_to_lt4p(uint32_t x) { return umin32(x, x - P4); }
B200_HD uint32_t fix_lt4p_to_lt2p(uint32_t x) { return umin32(x, x - P2); }
B200_HD uint32_t fix_lt8p_to_lt2p(uint32_t x) { return fix_lt4p_to_lt2p(fix_lt8p_to_lt4p(x)); }

// full reduction of any 32-bit value to [0, p)
B200_HD uint32_t reduce_full(uint32_t x)
{
    x = fix29(x);
    return x >= P ? x - P : x;
}
// exact signed integer v with v = x (mod p), valid when |v| <= CONV_BOUND and x < 4p; ALU only
B200_HD int32_t centered_lift_alu(uint32_t x)
{
    x = fix_lt4p_to_lt2p(x);
    x = umin32(x, x - P);                      // [0, p)
    return (int32_t)(x > P / 2 ? x - P : x);  // centred
}

// Montgomery reduction of a 64-bit accumulator: returns acc * 2^-32 mod p, lazily, < acc/2^32 + p.
B200_HD uint32_t redc64(uint64_t acc)
{
    const uint32_t m = (uint32_t)acc * PINVNEG;
    return (uint32_t)((acc + (uint64_t)m * P) >> 32);
}

// exact signed integer v with v = x (mod p), valid when |v| <= CONV_BOUND and x < 8p
B200_HD int32_t centered_lift(uint32_t x)
{
    const uint32_t q = (x + (1u << 28)) >> 29;
    return (int32_t)(x - q * P);
}

// ---- host-side helpers (table generation) ----
inline uint32_t mod_mul(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b % P); }
inline uint32_t mod_pow(uint32_t a, uint64_t e)
{
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = mod_mul(r, a);
        a = mod_mul(a, a);
        e >>= 1;
    }
    return r;
}
inline uint32_t mod_inv(uint32_t a) { return mod_pow(a, P - 2); }
inline tw_t make_tw(uint32_t w) { return tw_t{w, (uint32_t)(((uint64_t)w << 32) / P)}; }

}  // namespace b200

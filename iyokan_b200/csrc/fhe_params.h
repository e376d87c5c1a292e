// TFHE parameter set and device data-layout constants.  Two compile-time flavours, selected the way the reference
// selects them (-DIYOKAN_80BIT_SECURITY=On -> USE_80BIT_SECURITY, CMakeLists.txt:28-30, TFHEpp include/params.hpp:14-28):
//   default        128-bit: TFHEpp include/params/128bit.hpp:7-76
//                  lvl0 n=636 / uint16 torus, lvl1 N=1024 / uint32, l=3, Bgbit=6; key switch t=7, basebit=2
//   B200FHE_80BIT   80-bit: TFHEpp include/params/CGGI16.hpp:6-68
//                  lvl0 n=500 / uint32 torus, lvl1 N=1024 / uint32, l=2, Bgbit=10; key switch t=8, basebit=2
// The 80-bit library (libb200fhe80.so) is built from the same sources; it carries the generic blind-rotation
// kernel (brg_phases.h) only, the 128-bit one adds the shapes specialised for l = 3 (br_phases.h, br7_phases.h, ...).
#pragma once
#include "hd.h"

namespace b200 {

constexpr int N1 = 1024;           // lvl1 ring degree N
constexpr int NBIT = 10;
constexpr int KS_BASEBIT = 2;
constexpr uint32_t MU1 = 1u << 29;  // lvl1param::mu on the 32-bit torus
constexpr int TLWE1_LEN = N1 + 1;

#if defined(B200FHE_80BIT)
using torus0_t = uint32_t;          // lvl0param::T
constexpr int T0_BITS = 32;
constexpr int N0 = 500;             // lvl0 dimension n
constexpr int GL = 2;               // gadget length l
constexpr int BGBIT = 10;
constexpr int KS_T = 8;
constexpr uint32_t MU0 = 1u << 29;  // lvl0param::mu
// Exactness with the 29-bit NTT prime: 4 rows x 1024 x |digit| <= 512 x |limb| <= 64 = 2^27 < p/2, so the 32-bit key
// coefficient is split into five centred limbs of 7, 7, 6, 6 and 6 bits (modarith.h).
constexpr int LIMBS = 5;
B200_HD constexpr int limb_width(int l) { return l < 2 ? 7 : 6; }
B200_HD constexpr int limb_shift(int l) { return l == 0 ? 0 : l == 1 ? 7 : l == 2 ? 14 : l == 3 ? 20 : 26; }
constexpr int SLOT_STRIDE = 512;    // uint32 per TLWE slot in the arena (2048 B; 501 used)
#else
using torus0_t = uint16_t;
constexpr int T0_BITS = 16;
constexpr int N0 = 636;
constexpr int GL = 3;
constexpr int BGBIT = 6;
constexpr int KS_T = 7;
constexpr uint32_t MU0 = 1u << 13;  // lvl0param::mu on the 16-bit torus
// bootstrapping key split into 11 + 11 + 10-bit centred limbs: 6 x 1024 x 32 x 1024 < p/2
constexpr int LIMBS = 3;
B200_HD constexpr int limb_width(int l) { return l < 2 ? 11 : 10; }
B200_HD constexpr int limb_shift(int l) { return 11 * l; }
constexpr int SLOT_STRIDE = 640;    // uint16 per TLWE slot in the arena (1280 B; 637 used)
#endif
constexpr int LIMB_BITS = 11;       // 128-bit shapes: limb l sits at bit 11*l

constexpr uint32_t T0_MASK = T0_BITS == 32 ? 0xFFFFFFFFu : ((1u << (T0_BITS & 31)) - 1);
constexpr int ROWS = 2 * GL;        // (k+1)*l TRGSW rows
constexpr int TLWE0_LEN = N0 + 1;   // 637 uint16 = 1274 B (501 uint32 = 2004 B) on the wire
constexpr int SLOT_BYTES = SLOT_STRIDE * (T0_BITS / 8);
constexpr int SLOT_WORDS = SLOT_BYTES / 4;   // 32-bit words per slot (320 / 512)
constexpr int KSK_ROW = SLOT_STRIDE;         // torus0 elements per key-switching-key row
constexpr int U_STRIDE = 1028;      // uint32 per lvl1 TLWE in the rotation scratch buffer
constexpr int BK_COLS = 2 * LIMBS;  // NTT-domain output columns c = poly*LIMBS + limb
// mod switch lvl0 -> 2N (gatebootstrapping.hpp:26-30, 58-65): shift = digits - 1 - nbit, a is rounded, b is not
constexpr int MODSW_SHIFT = T0_BITS - 1 - NBIT;
constexpr uint32_t MODSW_ROUND = 1u << (MODSW_SHIFT - 1);

// Decomposition constants, TFHEpp include/trgsw.hpp:12-21,62-78
constexpr uint32_t dec_offset_()
{
    uint32_t o = 0;
    for (int i = 1; i <= GL; i++) o += (1u << (BGBIT - 1)) * (1u << (32 - i * BGBIT));
    return o;
}
constexpr uint32_t DEC_OFFSET = dec_offset_();
constexpr uint32_t DEC_ROUND = 1u << (32 - GL * BGBIT - 1);

// opcodes of the C ABI (include/b200fhe.h)
enum : uint8_t {
    OP_AND = 0, OP_NAND, OP_ANDNOT, OP_OR, OP_NOR, OP_ORNOT, OP_XOR, OP_XNOR,
    OP_MUX, OP_NOT, OP_COPY, OP_CONST0, OP_CONST1, OP_ANDNY, OP_ORNY, OP_NUM
};

}  // namespace b200

// 128-bit TFHE parameter set and device data-layout constants.
// Values follow the reference: TFHEpp include/params/128bit.hpp:7-76
// (lvl0 n=636/uint16, lvl1 N=1024/uint32, l=3, Bgbit=6; key switch t=7, basebit=2).
#pragma once
#include "hd.h"

namespace b200 {

constexpr int N0 = 636;            // lvl0 dimension n
constexpr int N1 = 1024;           // lvl1 ring degree N
constexpr int NBIT = 10;
constexpr int GL = 3;              // gadget length l
constexpr int BGBIT = 6;
constexpr int ROWS = 2 * GL;       // (k+1)*l TRGSW rows
constexpr int KS_T = 7;
constexpr int KS_BASEBIT = 2;
constexpr uint32_t MU0 = 1u << 13;  // lvl0param::mu on the 16-bit torus
constexpr uint32_t MU1 = 1u << 29;  // lvl1param::mu on the 32-bit torus
constexpr int TLWE0_LEN = N0 + 1;   // 637 uint16 = 1274 B on the wire
constexpr int TLWE1_LEN = N1 + 1;

// device layouts (padded for 16-byte vector access / TMA bulk copies)
constexpr int SLOT_STRIDE = 640;    // uint16 per TLWE slot in the arena (1280 B)
constexpr int KSK_ROW = 640;        // uint16 per key-switching-key row (1280 B)
constexpr int U_STRIDE = 1028;      // uint32 per lvl1 TLWE in the rotation scratch buffer
constexpr int LIMBS = 3;            // bootstrapping key split into 11+11+10-bit centred limbs
constexpr int LIMB_BITS = 11;
constexpr int BK_COLS = 2 * LIMBS;  // NTT-domain output columns c = poly*3 + limb

// Decomposition constants, TFHEpp include/trgsw.hpp:12-21,62-78
constexpr uint32_t DEC_OFFSET =
    (1u << (BGBIT - 1)) * ((1u << (32 - BGBIT)) + (1u << (32 - 2 * BGBIT)) + (1u << (32 - 3 * BGBIT)));
constexpr uint32_t DEC_ROUND = 1u << (32 - GL * BGBIT - 1);

// opcodes of the C ABI (include/b200fhe.h)
enum : uint8_t {
    OP_AND = 0, OP_NAND, OP_ANDNOT, OP_OR, OP_NOR, OP_ORNOT, OP_XOR, OP_XNOR,
    OP_MUX, OP_NOT, OP_COPY, OP_CONST0, OP_CONST1, OP_ANDNY, OP_ORNY, OP_NUM
};

}  // namespace b200

// Host-side translation of one gate frontier into kernel jobs (pure C++, shared by the CUDA
// library and the CPU simulator so the opcode table is tested without a GPU).
//
// Reference semantics: DEFINE_TASK_GATE in src/iyokan_tfhepp.hpp:109-144 maps each Task to a
// TFHEpp call; TFHEpp's HomGate table (include/gate.hpp:59-230) gives (casign, cbsign, offset);
// HomMUX<lvl0param> (gate.hpp:231-262) is two rotations, one key switch, + mu afterwards.
#pragma once
#include "br_phases.h"
#include "ks_phases.h"

namespace b200 {

struct BatchCounts {
    size_t nbr = 0, nks = 0, nun = 0;
};

// (sa, sb, off/mu0) of HomGate<.., casign, cbsign, offset>
inline bool gate_coef(uint8_t op, int& sa, int& sb, int& off)
{
    switch (op) {
    case OP_NAND:   sa = -1; sb = -1; off = +1; return true;  // gate.hpp:65
    case OP_NOR:    sa = -1; sb = -1; off = -1; return true;  // :82
    case OP_XNOR:   sa = -2; sb = -2; off = -2; return true;  // :99
    case OP_AND:    sa = +1; sb = +1; off = -1; return true;  // :116
    case OP_OR:     sa = +1; sb = +1; off = +1; return true;  // :133
    case OP_XOR:    sa = +2; sb = +2; off = +2; return true;  // :150
    case OP_ANDNY:  sa = -1; sb = +1; off = -1; return true;  // :167
    case OP_ANDNOT: sa = +1; sb = -1; off = -1; return true;  // :184 HomANDYN
    case OP_ORNY:   sa = -1; sb = +1; off = +1; return true;  // :201
    case OP_ORNOT:  sa = +1; sb = -1; off = +1; return true;  // :218 HomORYN
    default: return false;
    }
}

// br needs room for 2n jobs, ks and un for n.  Returns nullptr on success or an error string.
inline const char* build_gate_jobs(const uint8_t* opcode, const uint32_t* in0, const uint32_t* in1,
                                   const uint32_t* in2, const uint32_t* out, size_t n, size_t n_slots, BrJob* br,
                                   KsJob* ks, UnaryJob* un, BatchCounts& cnt)
{
    cnt = BatchCounts{};
    auto slot = [&](const uint32_t* arr, size_t i, uint32_t& dst) -> bool {
        if (!arr || arr[i] >= n_slots) return false;
        dst = arr[i];
        return true;
    };
    for (size_t i = 0; i < n; i++) {
        const uint8_t op = opcode[i];
        uint32_t o = 0, a = 0, b = 0, s = 0;
        if (!slot(out, i, o)) return "output slot missing or out of range";
        int sa, sb, off;
        if (gate_coef(op, sa, sb, off)) {
            if (!slot(in0, i, a) || !slot(in1, i, b)) return "input slot missing or out of range";
            BrJob& j = br[cnt.nbr];
            j.in[0] = a; j.in[1] = b; j.in[2] = 0;
            j.sgn[0] = (int8_t)sa; j.sgn[1] = (int8_t)sb; j.sgn[2] = 0; j.pad = 0;
            j.off = ((uint32_t)off * MU0) & T0_MASK;
            ks[cnt.nks++] = KsJob{(uint32_t)cnt.nbr, KS_NONE, o, 0u};
            cnt.nbr++;
        } else if (op == OP_MUX) {
            // cs = in2, c1 = in1, c0 = in0:  (cs + c1 - mu) and (-cs + c0 - mu), gate.hpp:236-240
            if (!slot(in0, i, a) || !slot(in1, i, b) || !slot(in2, i, s)) return "input slot missing or out of range";
            BrJob& j1 = br[cnt.nbr];
            j1.in[0] = a; j1.in[1] = b; j1.in[2] = s;
            j1.sgn[0] = 0; j1.sgn[1] = 1; j1.sgn[2] = 1; j1.pad = 0;
            j1.off = (0u - MU0) & T0_MASK;
            BrJob& j0 = br[cnt.nbr + 1];
            j0.in[0] = a; j0.in[1] = b; j0.in[2] = s;
            j0.sgn[0] = 1; j0.sgn[1] = 0; j0.sgn[2] = -1; j0.pad = 0;
            j0.off = (0u - MU0) & T0_MASK;
            ks[cnt.nks++] = KsJob{(uint32_t)cnt.nbr, (uint32_t)cnt.nbr + 1, o, MU0};  // + mu after the switch, :260
            cnt.nbr += 2;
        } else if (op == OP_NOT || op == OP_COPY) {
            if (!slot(in0, i, a)) return "input slot missing or out of range";
            un[cnt.nun++] = UnaryJob{a, o, op};
        } else if (op == OP_CONST0 || op == OP_CONST1) {
            un[cnt.nun++] = UnaryJob{0u, o, op};
        } else {
            return "unknown opcode";
        }
    }
    return nullptr;
}

}  // namespace b200

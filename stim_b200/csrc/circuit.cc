// circuit.cc — parser + gate table for the stand-alone host side. See circuit.h.
#include "circuit.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "program.h"

namespace gstim {

// ---------------------------------------------------------------------------------------------
// Frame action of the two-qubit Cliffords, written as the same XOR sequences the reference applies
// (/root/reference/src/stim/simulators/frame_simulator.inl:387-630) and folded into a 4x4 GF(2)
// matrix: output o in (x1,z1,x2,z2) gets nibble o; bit i of the nibble = input i contributes.
// ---------------------------------------------------------------------------------------------
namespace {
struct Sym {
    uint8_t v[4] = {1, 2, 4, 8};  // x1 z1 x2 z2
    constexpr uint16_t mat() const {
        return (uint16_t)(v[0] | (v[1] << 4) | (v[2] << 8) | (v[3] << 12));
    }
};
enum { X1 = 0, Z1 = 1, X2 = 2, Z2 = 3 };
constexpr void sx(Sym &s, int a, int b) {
    uint8_t t = s.v[a];
    s.v[a] = s.v[b];
    s.v[b] = t;
}
constexpr uint16_t m_cx() { Sym s; s.v[Z1] ^= s.v[Z2]; s.v[X2] ^= s.v[X1]; return s.mat(); }
constexpr uint16_t m_cy() { Sym s; s.v[Z1] ^= s.v[X2] ^ s.v[Z2]; s.v[Z2] ^= s.v[X1]; s.v[X2] ^= s.v[X1]; return s.mat(); }
constexpr uint16_t m_cz() { Sym s; s.v[Z1] ^= s.v[X2]; s.v[Z2] ^= s.v[X1]; return s.mat(); }
constexpr uint16_t m_swap() { Sym s; sx(s, Z1, Z2); sx(s, X1, X2); return s.mat(); }
constexpr uint16_t m_iswap() {
    Sym s;
    uint8_t dx = s.v[X1] ^ s.v[X2];
    uint8_t t1 = s.v[Z1] ^ dx, t2 = s.v[Z2] ^ dx;
    s.v[Z1] = t2; s.v[Z2] = t1;
    sx(s, X1, X2);
    return s.mat();
}
constexpr uint16_t m_cxswap() { Sym s; s.v[Z2] ^= s.v[Z1]; s.v[Z1] ^= s.v[Z2]; s.v[X1] ^= s.v[X2]; s.v[X2] ^= s.v[X1]; return s.mat(); }
constexpr uint16_t m_czswap() { Sym s; sx(s, Z1, Z2); sx(s, X1, X2); s.v[Z1] ^= s.v[X2]; s.v[Z2] ^= s.v[X1]; return s.mat(); }
constexpr uint16_t m_swapcx() { Sym s; s.v[Z1] ^= s.v[Z2]; s.v[Z2] ^= s.v[Z1]; s.v[X2] ^= s.v[X1]; s.v[X1] ^= s.v[X2]; return s.mat(); }
constexpr uint16_t m_sqrt_xx() { Sym s; uint8_t d = s.v[Z1] ^ s.v[Z2]; s.v[X1] ^= d; s.v[X2] ^= d; return s.mat(); }
constexpr uint16_t m_sqrt_yy() { Sym s; uint8_t d = s.v[X1] ^ s.v[Z1] ^ s.v[X2] ^ s.v[Z2]; s.v[X1] ^= d; s.v[Z1] ^= d; s.v[X2] ^= d; s.v[Z2] ^= d; return s.mat(); }
constexpr uint16_t m_sqrt_zz() { Sym s; uint8_t d = s.v[X1] ^ s.v[X2]; s.v[Z1] ^= d; s.v[Z2] ^= d; return s.mat(); }
constexpr uint16_t m_xcx() { Sym s; s.v[X1] ^= s.v[Z2]; s.v[X2] ^= s.v[Z1]; return s.mat(); }
constexpr uint16_t m_xcy() { Sym s; s.v[X1] ^= s.v[X2] ^ s.v[Z2]; s.v[X2] ^= s.v[Z1]; s.v[Z2] ^= s.v[Z1]; return s.mat(); }
constexpr uint16_t m_ycx() { Sym s; s.v[X2] ^= s.v[X1] ^ s.v[Z1]; s.v[X1] ^= s.v[Z2]; s.v[Z1] ^= s.v[Z2]; return s.mat(); }
constexpr uint16_t m_ycy() {
    Sym s;
    uint8_t y1 = s.v[X1] ^ s.v[Z1], y2 = s.v[X2] ^ s.v[Z2];
    s.v[X1] ^= y2; s.v[Z1] ^= y2; s.v[X2] ^= y1; s.v[Z2] ^= y1;
    return s.mat();
}
static_assert(m_cx() == GSTIM_MAT_CX, "CX matrix constant out of sync with program.h");
// XCZ a b == CX b a ; YCZ a b == CY b a  (frame_simulator.inl:592-598, 624-630): swap roles.
constexpr uint16_t swap_roles(uint16_t m) {
    // relabel inputs and outputs (x1,z1)<->(x2,z2)
    auto perm_nib = [](uint8_t n) -> uint8_t { return (uint8_t)(((n & 3) << 2) | ((n >> 2) & 3)); };
    uint8_t o0 = perm_nib(m & 15), o1 = perm_nib((m >> 4) & 15), o2 = perm_nib((m >> 8) & 15), o3 = perm_nib((m >> 12) & 15);
    return (uint16_t)(o2 | (o3 << 4) | (o0 << 8) | (o1 << 12));
}

// single qubit matrices: bit0 x'<-x, bit1 x'<-z, bit2 z'<-x, bit3 z'<-z
constexpr uint16_t C1_H = 0x6;      // swap(x,z)                      frame_simulator.inl:345-350
constexpr uint16_t C1_HXY = 0xD;    // z ^= x                         :353-358
constexpr uint16_t C1_HYZ = 0xB;    // x ^= z                         :361-366
constexpr uint16_t C1_CXYZ = 0x7;   // x ^= z ; z ^= x                :369-375
constexpr uint16_t C1_CZYX = 0xE;   // z ^= x ; x ^= z                :378-384

constexpr uint16_t meas(uint32_t basis, uint32_t kind) { return (uint16_t)(basis | (kind << 2)); }

const GateInfo GATES[] = {
    // annotations / no-ops on the frame (frame_simulator.inl:1097-1108)
    {"TICK", GateCat::NOOP, 0, 0, TR_NONE, false},
    {"QUBIT_COORDS", GateCat::NOOP, 0, ARGS_ANY, TR_QUBITS, false},
    {"SHIFT_COORDS", GateCat::NOOP, 0, ARGS_ANY, TR_NONE, false},
    {"I", GateCat::NOOP, 0, 0, TR_QUBITS, false},
    {"X", GateCat::NOOP, 0, 0, TR_QUBITS, false},
    {"Y", GateCat::NOOP, 0, 0, TR_QUBITS, false},
    {"Z", GateCat::NOOP, 0, 0, TR_QUBITS, false},
    {"II", GateCat::NOOP, 0, 0, TR_PAIRS, false},
    {"I_ERROR", GateCat::NOOP, 0, ARGS_ANY, TR_QUBITS, true},
    {"II_ERROR", GateCat::NOOP, 0, ARGS_ANY, TR_PAIRS, true},
    {"DETECTOR", GateCat::DETECTOR, 0, ARGS_ANY, TR_REC, false},
    {"OBSERVABLE_INCLUDE", GateCat::OBSERVABLE_INCLUDE, 0, 1, TR_REC_OR_PAULI, false},
    {"REPEAT", GateCat::REPEAT, 0, 0, TR_BLOCK, false},
    // single-qubit Cliffords (gate -> handler aliasing frame_simulator.inl:1025-1095)
    {"H", GateCat::CLIFF1, C1_H, 0, TR_QUBITS, false},
    {"H_NXZ", GateCat::CLIFF1, C1_H, 0, TR_QUBITS, false},
    {"SQRT_Y", GateCat::CLIFF1, C1_H, 0, TR_QUBITS, false},
    {"SQRT_Y_DAG", GateCat::CLIFF1, C1_H, 0, TR_QUBITS, false},
    {"S", GateCat::CLIFF1, C1_HXY, 0, TR_QUBITS, false},
    {"S_DAG", GateCat::CLIFF1, C1_HXY, 0, TR_QUBITS, false},
    {"H_XY", GateCat::CLIFF1, C1_HXY, 0, TR_QUBITS, false},
    {"H_NXY", GateCat::CLIFF1, C1_HXY, 0, TR_QUBITS, false},
    {"SQRT_X", GateCat::CLIFF1, C1_HYZ, 0, TR_QUBITS, false},
    {"SQRT_X_DAG", GateCat::CLIFF1, C1_HYZ, 0, TR_QUBITS, false},
    {"H_YZ", GateCat::CLIFF1, C1_HYZ, 0, TR_QUBITS, false},
    {"H_NYZ", GateCat::CLIFF1, C1_HYZ, 0, TR_QUBITS, false},
    {"C_XYZ", GateCat::CLIFF1, C1_CXYZ, 0, TR_QUBITS, false},
    {"C_NXYZ", GateCat::CLIFF1, C1_CXYZ, 0, TR_QUBITS, false},
    {"C_XNYZ", GateCat::CLIFF1, C1_CXYZ, 0, TR_QUBITS, false},
    {"C_XYNZ", GateCat::CLIFF1, C1_CXYZ, 0, TR_QUBITS, false},
    {"C_ZYX", GateCat::CLIFF1, C1_CZYX, 0, TR_QUBITS, false},
    {"C_NZYX", GateCat::CLIFF1, C1_CZYX, 0, TR_QUBITS, false},
    {"C_ZNYX", GateCat::CLIFF1, C1_CZYX, 0, TR_QUBITS, false},
    {"C_ZYNX", GateCat::CLIFF1, C1_CZYX, 0, TR_QUBITS, false},
    // two-qubit Cliffords
    {"CX", GateCat::CLIFF2, m_cx(), 0, TR_PAIRS_BITS, false},
    {"CY", GateCat::CLIFF2, m_cy(), 0, TR_PAIRS_BITS, false},
    {"CZ", GateCat::CLIFF2, m_cz(), 0, TR_PAIRS_BITS, false},
    {"XCZ", GateCat::CLIFF2, swap_roles(m_cx()), 0, TR_PAIRS_BITS, false},
    {"YCZ", GateCat::CLIFF2, swap_roles(m_cy()), 0, TR_PAIRS_BITS, false},
    {"XCX", GateCat::CLIFF2, m_xcx(), 0, TR_PAIRS, false},
    {"XCY", GateCat::CLIFF2, m_xcy(), 0, TR_PAIRS, false},
    {"YCX", GateCat::CLIFF2, m_ycx(), 0, TR_PAIRS, false},
    {"YCY", GateCat::CLIFF2, m_ycy(), 0, TR_PAIRS, false},
    {"SWAP", GateCat::CLIFF2, m_swap(), 0, TR_PAIRS, false},
    {"ISWAP", GateCat::CLIFF2, m_iswap(), 0, TR_PAIRS, false},
    {"ISWAP_DAG", GateCat::CLIFF2, m_iswap(), 0, TR_PAIRS, false},
    {"CXSWAP", GateCat::CLIFF2, m_cxswap(), 0, TR_PAIRS, false},
    {"SWAPCX", GateCat::CLIFF2, m_swapcx(), 0, TR_PAIRS, false},
    {"CZSWAP", GateCat::CLIFF2, m_czswap(), 0, TR_PAIRS, false},
    {"SQRT_XX", GateCat::CLIFF2, m_sqrt_xx(), 0, TR_PAIRS, false},
    {"SQRT_XX_DAG", GateCat::CLIFF2, m_sqrt_xx(), 0, TR_PAIRS, false},
    {"SQRT_YY", GateCat::CLIFF2, m_sqrt_yy(), 0, TR_PAIRS, false},
    {"SQRT_YY_DAG", GateCat::CLIFF2, m_sqrt_yy(), 0, TR_PAIRS, false},
    {"SQRT_ZZ", GateCat::CLIFF2, m_sqrt_zz(), 0, TR_PAIRS, false},
    {"SQRT_ZZ_DAG", GateCat::CLIFF2, m_sqrt_zz(), 0, TR_PAIRS, false},
    // collapsing gates
    {"M", GateCat::MEASURE, meas(GB_Z, GK_M), ARGS_ZERO_OR_ONE, TR_QUBITS_INV, true},
    {"MX", GateCat::MEASURE, meas(GB_X, GK_M), ARGS_ZERO_OR_ONE, TR_QUBITS_INV, true},
    {"MY", GateCat::MEASURE, meas(GB_Y, GK_M), ARGS_ZERO_OR_ONE, TR_QUBITS_INV, true},
    {"MR", GateCat::MEASURE, meas(GB_Z, GK_MR), ARGS_ZERO_OR_ONE, TR_QUBITS_INV, true},
    {"MRX", GateCat::MEASURE, meas(GB_X, GK_MR), ARGS_ZERO_OR_ONE, TR_QUBITS_INV, true},
    {"MRY", GateCat::MEASURE, meas(GB_Y, GK_MR), ARGS_ZERO_OR_ONE, TR_QUBITS_INV, true},
    {"R", GateCat::MEASURE, meas(GB_Z, GK_R), 0, TR_QUBITS, false},
    {"RX", GateCat::MEASURE, meas(GB_X, GK_R), 0, TR_QUBITS, false},
    {"RY", GateCat::MEASURE, meas(GB_Y, GK_R), 0, TR_QUBITS, false},
    {"MPAD", GateCat::MPAD, 0, ARGS_ZERO_OR_ONE, TR_MPAD, true},
    {"MPP", GateCat::MPP, 0, ARGS_ZERO_OR_ONE, TR_PRODUCTS, true},
    {"SPP", GateCat::SPP, 0, 0, TR_PRODUCTS, false},
    {"SPP_DAG", GateCat::SPP, 0, 0, TR_PRODUCTS, false},
    {"MXX", GateCat::MPAIR, GB_X, ARGS_ZERO_OR_ONE, TR_PAIRS_INV, true},
    {"MYY", GateCat::MPAIR, GB_Y, ARGS_ZERO_OR_ONE, TR_PAIRS_INV, true},
    {"MZZ", GateCat::MPAIR, GB_Z, ARGS_ZERO_OR_ONE, TR_PAIRS_INV, true},
    // noise
    {"X_ERROR", GateCat::NOISE1, 1, 1, TR_QUBITS, true},
    {"Z_ERROR", GateCat::NOISE1, 2, 1, TR_QUBITS, true},
    {"Y_ERROR", GateCat::NOISE1, 3, 1, TR_QUBITS, true},
    {"DEPOLARIZE1", GateCat::NOISE1, 4, 1, TR_QUBITS, true},
    {"DEPOLARIZE2", GateCat::DEPOLARIZE2, 0, 1, TR_PAIRS, true},
    {"PAULI_CHANNEL_1", GateCat::PAULI_CHANNEL_1, 0, 3, TR_QUBITS, true},
    {"PAULI_CHANNEL_2", GateCat::PAULI_CHANNEL_2, 0, 15, TR_PAIRS, true},
    {"E", GateCat::CORR, 1, 1, TR_PAULIS, true},
    {"ELSE_CORRELATED_ERROR", GateCat::CORR, 0, 1, TR_PAULIS, true},
    {"HERALDED_ERASE", GateCat::HERALDED_ERASE, 0, 1, TR_QUBITS, true},
    {"HERALDED_PAULI_CHANNEL_1", GateCat::HERALDED_PAULI_CHANNEL_1, 0, 4, TR_QUBITS, true},
};

// aliases: /root/reference/src/stim/gates/gate_data_*.cc add_gate_alias calls
const std::pair<const char *, const char *> ALIASES[] = {
    {"MZ", "M"},       {"MRZ", "MR"},  {"RZ", "R"},        {"ZCX", "CX"},
    {"CNOT", "CX"},    {"ZCY", "CY"},  {"ZCZ", "CZ"},      {"H_XZ", "H"},
    {"CORRELATED_ERROR", "E"},         {"SQRT_Z", "S"},    {"SQRT_Z_DAG", "S_DAG"},
    {"SWAPCZ", "CZSWAP"},
};

const std::map<std::string, const GateInfo *> &gate_map() {
    static const std::map<std::string, const GateInfo *> m = [] {
        std::map<std::string, const GateInfo *> r;
        for (const auto &g : GATES) {
            r[g.name] = &g;
        }
        for (const auto &a : ALIASES) {
            r[a.first] = r.at(a.second);
        }
        return r;
    }();
    return m;
}

[[noreturn]] void fail(const std::string &msg) {
    throw std::invalid_argument(msg);
}

struct Reader {
    std::string_view s;
    size_t p = 0;
    size_t line = 1;

    bool eof() const { return p >= s.size(); }
    char peek() const { return p < s.size() ? s[p] : '\0'; }
    void skip_inline_space() {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\r')) {
            p++;
        }
    }
    // skips whitespace, newlines and comments between instructions
    void skip_dead_space() {
        while (p < s.size()) {
            char c = s[p];
            if (c == '\n') {
                line++;
                p++;
            } else if (c == ' ' || c == '\t' || c == '\r') {
                p++;
            } else if (c == '#') {
                while (p < s.size() && s[p] != '\n') {
                    p++;
                }
            } else {
                break;
            }
        }
    }
    [[noreturn]] void err(const std::string &msg) const {
        fail("Circuit parse error at line " + std::to_string(line) + ": " + msg);
    }
    uint64_t read_uint(uint64_t limit, const char *what) {
        if (!isdigit((unsigned char)peek())) {
            err(std::string("Expected a digit while reading ") + what + ".");
        }
        uint64_t v = 0;
        while (isdigit((unsigned char)peek())) {
            v = v * 10 + (uint64_t)(s[p] - '0');
            if (v > limit) {
                err(std::string("Number too large while reading ") + what + ".");
            }
            p++;
        }
        return v;
    }
    void expect(char c) {
        if (peek() != c) {
            err(std::string("Expected '") + c + "'.");
        }
        p++;
    }
    void expect_word(const char *w) {
        for (const char *q = w; *q; q++) {
            expect(*q);
        }
    }
};

uint32_t read_target(Reader &r) {
    char c = r.peek();
    if (isdigit((unsigned char)c)) {
        return (uint32_t)r.read_uint(T_VALUE_MASK, "a qubit target");
    }
    if (c == '!') {
        r.p++;
        char d = r.peek();
        uint32_t t;
        if (d == 'X' || d == 'Y' || d == 'Z' || d == 'x' || d == 'y' || d == 'z') {
            t = read_target(r);
        } else {
            t = (uint32_t)r.read_uint(T_VALUE_MASK, "an inverted qubit target");
        }
        return t | T_INVERTED;
    }
    if (c == 'X' || c == 'Y' || c == 'Z' || c == 'x' || c == 'y' || c == 'z') {
        uint32_t m = 0;
        char u = (char)toupper((unsigned char)c);
        if (u == 'X') {
            m = T_PAULI_X;
        } else if (u == 'Y') {
            m = T_PAULI_X | T_PAULI_Z;
        } else {
            m = T_PAULI_Z;
        }
        r.p++;
        if (r.peek() == ' ') {
            r.err("Pauli target followed by a space instead of a qubit index.");
        }
        return m | (uint32_t)r.read_uint(T_VALUE_MASK, "a Pauli target");
    }
    if (c == 'r') {
        r.expect_word("rec[-");
        uint64_t k = r.read_uint(T_VALUE_MASK, "a record lookback");
        r.expect(']');
        if (k == 0) {
            r.err("rec[-0] is not a valid measurement record target.");
        }
        return T_REC | (uint32_t)k;
    }
    if (c == 's') {
        r.expect_word("sweep[");
        uint64_t k = r.read_uint(T_VALUE_MASK, "a sweep index");
        r.expect(']');
        return T_SWEEP | (uint32_t)k;
    }
    if (c == '*') {
        r.p++;
        return T_COMBINER;
    }
    r.err(std::string("Unrecognized target prefix '") + c + "'.");
}

void validate(const Instruction &inst, const Reader &r) {
    const GateInfo &g = *inst.gate;
    const auto &ts = inst.targets;
    const std::string name = g.name;

    // argument count
    if (g.arg_count == ARGS_ZERO_OR_ONE) {
        if (inst.args.size() > 1) {
            r.err("Gate " + name + " takes 0 or 1 parens arguments.");
        }
    } else if (g.arg_count != ARGS_ANY && (int)inst.args.size() != g.arg_count) {
        r.err("Gate " + name + " was given " + std::to_string(inst.args.size()) + " parens arguments but takes " +
              std::to_string(g.arg_count) + ".");
    }
    if (g.args_are_probs) {
        double total = 0;
        for (double p : inst.args) {
            if (!(p >= 0 && p <= 1)) {
                r.err("Gate " + name + " only takes probability arguments, but one of its arguments wasn't a probability.");
            }
            total += p;
        }
        if (total > 1.0000001) {
            r.err("The disjoint probability arguments given to gate " + name + " sum to more than 1.");
        }
    }
    if (g.cat == GateCat::OBSERVABLE_INCLUDE) {
        double a = inst.args[0];
        if (a < 0 || a != std::round(a)) {
            r.err("Gate OBSERVABLE_INCLUDE only takes non-negative integer arguments.");
        }
    }

    auto is_plain = [](uint32_t t) { return (t & ~T_VALUE_MASK) == 0; };
    auto is_bit = [](uint32_t t) { return (t & (T_REC | T_SWEEP)) != 0; };
    auto is_pauli = [](uint32_t t) { return (t & (T_PAULI_X | T_PAULI_Z)) != 0 && !(t & (T_REC | T_SWEEP | T_COMBINER)); };
    auto check_pairs = [&]() {
        if (ts.size() & 1) {
            r.err("Two qubit gate " + name + " requires an even number of targets.");
        }
        for (size_t k = 0; k < ts.size(); k += 2) {
            if (ts[k] == ts[k + 1]) {
                r.err("The two qubit gate " + name + " was applied to a target pair with the same target twice.");
            }
        }
    };
    auto check_combiners = [&]() {
        bool allowed = false, just_saw = false, bad = false;
        for (uint32_t t : ts) {
            if (t == T_COMBINER) {
                bad |= !allowed;
                allowed = false;
                just_saw = true;
            } else {
                allowed = true;
                just_saw = false;
            }
        }
        if (bad || just_saw) {
            r.err("Gate " + name + " given combiners ('*') that aren't between other targets.");
        }
    };

    switch (g.targets) {
        case TR_NONE:
            if (!ts.empty()) {
                r.err("Gate " + name + " takes no targets.");
            }
            break;
        case TR_QUBITS:
            for (uint32_t t : ts) {
                if (!is_plain(t)) {
                    r.err("Gate " + name + " only takes qubit targets.");
                }
            }
            break;
        case TR_QUBITS_INV:
            for (uint32_t t : ts) {
                if (!is_plain(t & ~T_INVERTED)) {
                    r.err("Gate " + name + " only takes (optionally inverted) qubit targets.");
                }
            }
            break;
        case TR_PAIRS:
            check_pairs();
            for (uint32_t t : ts) {
                if (!is_plain(t)) {
                    r.err("Gate " + name + " only takes qubit targets.");
                }
            }
            break;
        case TR_PAIRS_INV:
            check_pairs();
            for (uint32_t t : ts) {
                if (!is_plain(t & ~T_INVERTED)) {
                    r.err("Gate " + name + " only takes (optionally inverted) qubit targets.");
                }
            }
            break;
        case TR_PAIRS_BITS:
            check_pairs();
            for (uint32_t t : ts) {
                if (!is_plain(t) && !(is_bit(t) && !(t & (T_INVERTED | T_PAULI_X | T_PAULI_Z | T_COMBINER)))) {
                    r.err("Gate " + name + " only takes qubit, rec[-k] or sweep[k] targets.");
                }
            }
            break;
        case TR_REC:
            for (uint32_t t : ts) {
                if ((t & ~T_VALUE_MASK) != T_REC) {
                    r.err("Gate " + name + " only takes measurement record targets (rec[-k]).");
                }
            }
            break;
        case TR_REC_OR_PAULI:
            for (uint32_t t : ts) {
                if ((t & ~T_VALUE_MASK) != T_REC && !is_pauli(t & ~T_INVERTED)) {
                    r.err("Gate " + name + " only takes measurement record targets and Pauli targets (rec[-k], Xk, Yk, Zk).");
                }
            }
            break;
        case TR_PAULIS:
            for (uint32_t t : ts) {
                if (!is_pauli(t)) {
                    r.err("Gate " + name + " only takes Pauli targets ('X2', 'Y3', 'Z5', etc).");
                }
            }
            break;
        case TR_PRODUCTS:
            check_combiners();
            for (uint32_t t : ts) {
                if (t != T_COMBINER && !is_pauli(t & ~T_INVERTED)) {
                    r.err("Gate " + name + " only takes Pauli targets ('X2', 'Y3', 'Z5', etc).");
                }
            }
            break;
        case TR_PRODUCTS_BITS:
            check_combiners();
            for (uint32_t t : ts) {
                if (t != T_COMBINER && !is_pauli(t & ~T_INVERTED) && !is_bit(t)) {
                    r.err("Gate " + name + " only takes Pauli targets or bit targets.");
                }
            }
            break;
        case TR_MPAD:
            for (uint32_t t : ts) {
                if (t != 0 && t != 1) {
                    r.err("Gate MPAD only takes 0 or 1 as targets.");
                }
            }
            break;
        case TR_BLOCK:
            break;
    }
}

void read_ops(Reader &r, Circuit &out, bool in_block) {
    while (true) {
        r.skip_dead_space();
        if (r.eof()) {
            if (in_block) {
                r.err("Unterminated block. Got a '{' without an eventual '}'.");
            }
            return;
        }
        if (r.peek() == '}') {
            if (!in_block) {
                r.err("Uninitiated block. Got a '}' without a '{'.");
            }
            r.p++;
            return;
        }
        // gate name
        size_t start = r.p;
        while (isalnum((unsigned char)r.peek()) || r.peek() == '_') {
            r.p++;
        }
        if (r.p == start) {
            r.err(std::string("Unexpected character '") + r.peek() + "'.");
        }
        std::string name(r.s.substr(start, r.p - start));
        const GateInfo *g = find_gate(name);
        if (g == nullptr) {
            r.err("Gate not found: '" + name + "'.");
        }
        Instruction inst;
        inst.gate = g;
        // optional tag
        if (r.peek() == '[') {
            while (!r.eof() && r.peek() != ']' && r.peek() != '\n') {
                r.p++;
            }
            r.expect(']');
        }
        // optional parens args
        if (r.peek() == '(') {
            r.p++;
            while (true) {
                r.skip_inline_space();
                if (r.peek() == ')') {
                    r.p++;
                    break;
                }
                const char *b = r.s.data() + r.p;
                // strtod needs a terminated buffer; copy the token.
                size_t e = r.p;
                while (e < r.s.size() && r.s[e] != ',' && r.s[e] != ')' && r.s[e] != '\n') {
                    e++;
                }
                std::string tok(b, e - r.p);
                char *endp = nullptr;
                double v = strtod(tok.c_str(), &endp);
                while (endp && (*endp == ' ' || *endp == '\t')) {
                    endp++;
                }
                if (tok.empty() || endp == tok.c_str() || *endp != '\0') {
                    r.err("Not a real number: '" + tok + "'.");
                }
                inst.args.push_back(v);
                r.p = e;
                r.skip_inline_space();
                if (r.peek() == ',') {
                    r.p++;
                } else if (r.peek() != ')') {
                    r.err("Parens arguments must be separated by commas and end with ')'.");
                }
            }
        }
        if (g->cat == GateCat::REPEAT) {
            r.skip_inline_space();
            inst.repeat_count = r.read_uint((uint64_t)1 << 62, "a repetition count");
            r.skip_inline_space();
            if (r.peek() != '{') {
                r.err("Missing '{' at start of REPEAT block.");
            }
            r.p++;
            if (inst.repeat_count == 0) {
                r.err("Repeating 0 times is not supported.");
            }
            inst.block_index = (uint32_t)out.blocks.size();
            out.blocks.emplace_back();
            // note: out.blocks may reallocate during recursion only for nested blocks of the child.
            Circuit body;
            read_ops(r, body, true);
            out.blocks[inst.block_index] = std::move(body);
            out.ops.push_back(std::move(inst));
            continue;
        }
        // targets until end of line
        while (true) {
            size_t before = r.p;
            r.skip_inline_space();
            char c = r.peek();
            if (c == '\n' || c == '\0' || c == '#' || c == '}') {
                break;
            }
            if (c == '{') {
                r.err("Unexpected '{'.");
            }
            bool had_space = r.p > before;
            bool prev_comb = !inst.targets.empty() && inst.targets.back() == T_COMBINER;
            if (!had_space && c != '*' && !prev_comb) {
                r.err("Targets must be separated by spacing.");
            }
            inst.targets.push_back(read_target(r));
        }
        validate(inst, r);
        out.ops.push_back(std::move(inst));
    }
}

}  // namespace

const GateInfo *find_gate(std::string_view name) {
    std::string up(name);
    for (auto &c : up) {
        c = (char)toupper((unsigned char)c);
    }
    const auto &m = gate_map();
    auto it = m.find(up);
    return it == m.end() ? nullptr : it->second;
}

Circuit Circuit::from_text(std::string_view text) {
    Reader r{text};
    Circuit c;
    read_ops(r, c, false);
    return c;
}

// ---------------------------------------------------------------------------------------------
// Stats (mirrors CircuitInstruction::add_stats_to, circuit_instruction.cc:44-99).
// ---------------------------------------------------------------------------------------------
namespace {
uint64_t count_results(const Instruction &op) {
    switch (op.gate->cat) {
        case GateCat::MEASURE:
            return ((op.gate->param >> 2) == GK_R) ? 0 : op.targets.size();
        case GateCat::MPAD:
        case GateCat::HERALDED_ERASE:
        case GateCat::HERALDED_PAULI_CHANNEL_1:
            return op.targets.size();
        case GateCat::MPAIR:
            return op.targets.size() / 2;
        case GateCat::MPP: {
            uint64_t n = op.targets.size();
            for (uint32_t t : op.targets) {
                if (t == T_COMBINER) {
                    n -= 2;
                }
            }
            return n;
        }
        default:
            return 0;
    }
}

void add_stats(const Circuit &c, CircuitStats &s) {
    for (const auto &op : c.ops) {
        if (op.gate->cat == GateCat::REPEAT) {
            CircuitStats body;
            add_stats(c.blocks[op.block_index], body);
            s.num_qubits = std::max(s.num_qubits, body.num_qubits);
            s.max_lookback = std::max(s.max_lookback, body.max_lookback);
            s.num_sweep_bits = std::max(s.num_sweep_bits, body.num_sweep_bits);
            s.num_observables = std::max(s.num_observables, body.num_observables);
            s.num_measurements += body.num_measurements * op.repeat_count;
            s.num_detectors += body.num_detectors * op.repeat_count;
            continue;
        }
        for (uint32_t t : op.targets) {
            if (t == T_COMBINER) {
                continue;
            }
            uint32_t v = t & T_VALUE_MASK;
            if (t & T_REC) {
                s.max_lookback = std::max<uint64_t>(s.max_lookback, v);
            } else if (t & T_SWEEP) {
                s.num_sweep_bits = std::max<uint64_t>(s.num_sweep_bits, (uint64_t)v + 1);
            } else if (op.gate->cat != GateCat::MPAD) {
                s.num_qubits = std::max<uint64_t>(s.num_qubits, (uint64_t)v + 1);
            }
        }
        s.num_measurements += count_results(op);
        if (op.gate->cat == GateCat::DETECTOR) {
            s.num_detectors++;
        } else if (op.gate->cat == GateCat::OBSERVABLE_INCLUDE) {
            s.num_observables = std::max<uint64_t>(s.num_observables, (uint64_t)op.args[0] + 1);
        }
    }
}
}  // namespace

CircuitStats compute_stats(const Circuit &c) {
    CircuitStats s;
    add_stats(c, s);
    return s;
}

}  // namespace gstim

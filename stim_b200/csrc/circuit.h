// circuit.h — minimal host-side circuit IR + parser for the Stim circuit language.
//
// In a real integration Stim's own parser stays (north_star: "the C++ host stays the same");
// this stand-alone reader exists so the library can be driven without linking the reference.
// It accepts the file format documented in /root/reference/doc/file_format_stim_circuit.md
// (instruction := NAME[tag](args) targets ; REPEAT n { ... } ; '#' comments) and applies the
// same per-gate validation rules as CircuitInstruction::validate
// (/root/reference/src/stim/circuit/circuit_instruction.cc:103-260): pair gates need an even
// target count, probabilities must lie in [0,1] and sum to <=1, record/sweep/Pauli targets only
// where the gate allows them. Violations throw std::invalid_argument like the reference.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace gstim {

// Target encoding (32-bit word). Value in the low 24 bits.
constexpr uint32_t T_VALUE_MASK = (1u << 24) - 1;
constexpr uint32_t T_INVERTED = 1u << 31;
constexpr uint32_t T_PAULI_X = 1u << 30;
constexpr uint32_t T_PAULI_Z = 1u << 29;
constexpr uint32_t T_REC = 1u << 28;
constexpr uint32_t T_COMBINER = 1u << 27;
constexpr uint32_t T_SWEEP = 1u << 26;

enum class GateCat : uint8_t {
    NOOP,         // TICK, QUBIT_COORDS, SHIFT_COORDS, I, X, Y, Z, II, I_ERROR, II_ERROR
    CLIFF1,       // single-qubit Clifford: param = 2x2 GF(2) matrix
    CLIFF2,       // two-qubit Clifford: param = 4x4 GF(2) matrix
    MEASURE,      // M MX MY MR MRX MRY R RX RY : param = basis | kind<<2
    MPAD,
    MPP,
    SPP,          // SPP and SPP_DAG (identical frame action)
    MPAIR,        // MXX MYY MZZ : param = basis
    NOISE1,       // X_ERROR Y_ERROR Z_ERROR DEPOLARIZE1 : param = kind
    DEPOLARIZE2,
    PAULI_CHANNEL_1,
    PAULI_CHANNEL_2,
    CORR,         // E / ELSE_CORRELATED_ERROR : param = 1 for E (resets flag)
    HERALDED_ERASE,
    HERALDED_PAULI_CHANNEL_1,
    DETECTOR,
    OBSERVABLE_INCLUDE,
    REPEAT,
};

// How many parens arguments a gate takes.
constexpr int ARGS_ANY = -1;
constexpr int ARGS_ZERO_OR_ONE = -2;

enum TargetRule : uint8_t {
    TR_NONE,         // no targets allowed (TICK, SHIFT_COORDS)
    TR_QUBITS,       // plain qubits
    TR_QUBITS_INV,   // qubits, optionally inverted with '!' (measurements)
    TR_PAIRS,        // plain qubit pairs
    TR_PAIRS_INV,    // qubit pairs, optionally inverted (MXX..)
    TR_PAIRS_BITS,   // qubit pairs where rec[-k]/sweep[k] may appear (CX CY CZ XCZ YCZ)
    TR_REC,          // rec[-k] only (DETECTOR)
    TR_REC_OR_PAULI, // rec[-k] or Pauli targets (OBSERVABLE_INCLUDE)
    TR_PAULIS,       // Pauli targets, no combiners (E, ELSE_CORRELATED_ERROR)
    TR_PRODUCTS,     // Pauli products with '*' combiners, optional '!' (MPP)
    TR_PRODUCTS_BITS,// Pauli products that may include rec/sweep bits (SPP, SPP_DAG)
    TR_MPAD,         // literal 0 / 1
    TR_BLOCK,        // REPEAT
};

struct GateInfo {
    const char *name;
    GateCat cat;
    uint16_t param;
    int arg_count;
    TargetRule targets;
    bool args_are_probs;
};

const GateInfo *find_gate(std::string_view name);  // case-insensitive, resolves aliases; nullptr if unknown

struct Instruction {
    const GateInfo *gate = nullptr;
    std::vector<double> args;
    std::vector<uint32_t> targets;
    // REPEAT only:
    uint64_t repeat_count = 0;
    uint32_t block_index = 0;
};

struct Circuit {
    std::vector<Instruction> ops;
    std::vector<Circuit> blocks;

    static Circuit from_text(std::string_view text);

    // Visits every non-REPEAT instruction in execution order with REPEAT blocks unrolled
    // (same traversal as stim::Circuit::for_each_operation, /root/reference/src/stim/circuit/circuit.h:181-193).
    template <typename F>
    void for_each_operation(F &&f) const {
        for (const auto &op : ops) {
            if (op.gate->cat == GateCat::REPEAT) {
                const Circuit &body = blocks[op.block_index];
                for (uint64_t r = 0; r < op.repeat_count; r++) {
                    body.for_each_operation(f);
                }
            } else {
                f(op);
            }
        }
    }
    // Visits every non-REPEAT instruction of the text once (REPEAT bodies are not unrolled).
    template <typename F>
    void for_each_instruction_once(F &&f) const {
        for (const auto &op : ops) {
            if (op.gate->cat == GateCat::REPEAT) {
                blocks[op.block_index].for_each_instruction_once(f);
            } else {
                f(op);
            }
        }
    }
};

// Aggregate sizes, mirroring stim::CircuitStats (/root/reference/src/stim/circuit/circuit_instruction.h:30-50).
struct CircuitStats {
    uint64_t num_qubits = 0;
    uint64_t num_measurements = 0;
    uint64_t num_detectors = 0;
    uint64_t num_observables = 0;
    uint64_t max_lookback = 0;
    uint64_t num_sweep_bits = 0;
};
CircuitStats compute_stats(const Circuit &c);

}  // namespace gstim

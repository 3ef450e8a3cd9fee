// program.h — the lowered instruction stream shared by the host lowering (lowering.cc)
// and the sm_100a interpreter kernel (kernels.cu).
//
// There is no counterpart of this format in the reference: Stim's FrameSimulator walks the
// Circuit object directly (/root/reference/src/stim/simulators/frame_simulator.inl:166-170, 915-1113).
// Here the circuit is flattened once on the host (REPEAT unrolled, MPP/MXX/SPP decomposed,
// PAULI_CHANNEL folded to a single site, qubit indices compacted, rec[-k] made absolute) into
// a flat array of uint32 words that a thread block streams through shared memory.
//
// A program is a sequence of BATCHES. A batch = 12 header words + payload. Every item of a
// batch touches resources (qubits, record rows, output rows) disjoint from every other item
// of the same batch, so items are executed concurrently, item i by thread-group (i % slots).
// The program is cut into CHUNKS of chunk_words words; no batch straddles a chunk boundary
// (the tail of a chunk is an OP_NEXT_CHUNK header). Chunks are what the kernel's bulk-async
// (TMA 1D) copies move into the shared-memory ring.
#pragma once
#include <stdint.h>

#define GSTIM_HDR_WORDS 12u

// Shots are simulated in "columns" of 128 shots (one uint4 per qubit per column).
#define GSTIM_COL_SHOTS 128u

enum GstimOp : uint32_t {
    GOP_END = 0,         // end of program
    GOP_NEXT_CHUNK = 1,  // rest of this chunk is padding
    GOP_CLIFF1 = 2,      // aux = 2x2 GF(2) matrix: bit0 x'<-x, bit1 x'<-z, bit2 z'<-x, bit3 z'<-z. item: qubit
    GOP_CLIFF2 = 3,      // aux = 4x4 GF(2) matrix over (x1,z1,x2,z2), 4 bits per output. item: q1 | q2<<16
    GOP_NOISE1 = 4,      // single-target Pauli noise site per item. item: qubit (clock + frame target)
    GOP_NOISE2 = 5,      // two-target Pauli noise site per item. item: q1 | q2<<16 (clock = q1)
    GOP_MEASURE = 6,     // aux = basis | kind<<2. item: qubit (physical row) | logical qubit index << 16
    GOP_RECZERO = 7,     // zero record rows rec0 .. rec0+n-1 (no payload)
    GOP_XORROWS = 8,     // out row (^)= XOR of record rows. payload: dst[n], off[n+1], idx[...]
    GOP_OBS_PAULI = 9,   // out row ^= frame component. payload per item: dst row, qubit | x<<30 | z<<31
    GOP_FEEDBACK = 10,   // frame ^= record row. payload per item: rec index, qubit | x<<30 | z<<31
    GOP_CORR = 11,       // E / ELSE_CORRELATED_ERROR: one site, payload = Pauli targets qubit | x<<30 | z<<31
    GOP_QMAP = 12,       // not executed: payload[i] = logical index of physical frame row GH_EXTRA + i
    GOP_SWEEP = 13,      // never serialised (m2d lowering only): frame ^= sweep bit. payload per item: sweep index, qubit | x<<30 | z<<31
};

// header word indices
enum GstimHdr : uint32_t {
    GH_OP = 0,       // op | flags<<8 | aux<<16
    GH_N = 1,        // number of items
    GH_WORDS = 2,    // total words of this batch (header + payload)
    GH_EXTRA = 3,    // op specific (NOISE: clock override qubit+1 or 0; CORR: clock qubit)
    GH_LAMBDA_LO = 4,  // lambda = -log1p(-p) per shot as u64 fixed point, unit 2^-56 nat (lo word)
    GH_LAMBDA_HI = 5,
    GH_SITE0 = 6,    // noise group of the batch (Philox counter word 0 of its event draws)
    GH_CSITE0 = 7,   // measure group of the batch (Philox counter word 0 of its collapse draws)
    GH_REC0 = 8,     // absolute measurement index of item 0
    GH_T1 = 9,       // NOISE1: category thresholds on a uniform u32
    GH_T2 = 10,
    GH_T3 = 11,
};

// header flags (bits 8..15 of word 0)
#define GF_BARRIER 0x01u     // __syncthreads() before executing this batch
#define GF_REC 0x02u         // NOISE1: every event also flips bit in record row rec0+i (heralds, M(p) noise)
#define GF_ACCUM 0x04u       // XORROWS: dst ^= value (instead of dst = value)
#define GF_RESET_FLAG 0x08u  // CORR: clear the "correlated error occurred" row first (E vs ELSE)
#define GF_TABLE 0x10u       // NOISE2: 15 cumulative u32 thresholds follow the header (PAULI_CHANNEL_2)
#define GF_NOFRAME 0x20u     // NOISE1: item is not a frame qubit (MPAD noise); clock = GH_EXTRA-1
#define GF_NOENTRY 0x80u     // NOISE1/NOISE2: the previous batch was a noise batch too: its exit barrier serves as this entry barrier
#define GF_DET 0x40u         // MEASURE: 3 payload words per item: qubit word, detector row to write (or 0xFFFFFFFF), record slot
                             //          to XOR the fresh result with (detector fused into the measurement, lowering.cc)

// CLIFF2 matrix of CX (the interpreter has a dedicated path for it; circuit.cc static_asserts the value)
#define GSTIM_MAT_CX 0x85A1u

// MEASURE aux encoding
#define GB_X 0u
#define GB_Y 1u
#define GB_Z 2u
#define GK_M 0u   // measure, keep
#define GK_MR 1u  // measure, reset
#define GK_R 2u   // reset only (no record)

// NOISE1 aux: four 2-bit Pauli categories c0..c3 (bit0 = flip x, bit1 = flip z), c_j at bits 2j..2j+1.
//   v = uniform u32;  v < T1 -> c0;  v < T2 -> c1;  v < T3 -> c2;  else c3.

// Philox counters (key = seed):
//   collapse of a qubit:    (measure group, logical qubit, global column lo, 'COLL' ^ global column hi)
//   call c of a noise slice: (noise group, 0x80000000 | slice index in the group, col0 lo, col0 hi | c << 15)
//     -> draws 2c (words 0, 1) and 2c + 1 (words 2, 3); a draw = (gap word, Pauli word of the event the gap leads to)
// col0 = global column of the shot block's first column (< 2^47).
#define GTAG_COLLAPSE 0x434F4C4Cu
#define GSTIM_SLICE_FLAG 0x80000000u
#define GSTIM_DRAW_SHIFT 15u

// Noise schedule: one info record per noise batch (NOISE1 / NOISE2 / CORR, numbered in program order; the
// ordinal is stored in the batch header's GH_CSITE0 word). The sites of a noise group (in target order) are
// cut into SLICES of GSTIM_NOISE_SLICE consecutive sites; a slice x a shot block is one Bernoulli sequence
// (site-major, then shot) walked with geometric gaps drawn from the slice's own Philox stream, like the
// reference's RareErrorIterator over targets x shots. Lowering never cuts a group into batches inside a slice.
//   slice (8 words): noise group, slice index in the group, noise batch ordinal | rate index << 16,
//                    first item of the slice in its batch | number of sites << 11,
//                    then what an event needs from its batch header: op | flags << 8 | aux << 16, T1, T2, T3
//   rate  (2 u64):   lam, floor((2^64 - 1) / lam)
// The kernel's event pre-pass walks the slices and leaves compact event records for the interpreter:
//   record = shot (bits 0-11) | item (12-22) | flips x1,z1,x2,z2 (23-26) | record flip (27) | conflict (28).
#define GSTIM_NOISE_SLICE 32u
#define GSTIM_SLICE_WORDS 8u
#define GSTIM_RATE_SMEM_MAX 64u    // the first 64 rates are mirrored in shared memory
#define GSTIM_NOISE_INFO_WORDS 12u
enum GstimNoiseInfo : uint32_t {
    GNI_H0 = 0,         // op | flags<<8 | aux<<16 of the batch
    GNI_N = 1,          // number of sites
    GNI_LAM_LO = 2,
    GNI_LAM_HI = 3,
    GNI_GROUP = 4,      // noise group (Philox counter word 0)
    GNI_T1 = 5,
    GNI_T2 = 6,
    GNI_T3 = 7,
    GNI_TABLE_OFF = 8,  // word offset of the 15 PAULI_CHANNEL_2 thresholds in the program (0 = none)
};
#define GSTIM_EV_SHOT_BITS 12u
#define GSTIM_EV_ITEM_SHIFT 12u
#define GSTIM_EV_ITEM_MASK 0x7FFu
#define GSTIM_EV_FLIP_SHIFT 23u
#define GSTIM_EV_CONFLICT 0x10000000u  // another record of the batch flips the same 32-bit frame words: apply atomically
#define GSTIM_MAX_BATCH_ITEMS 2047u
#define GSTIM_EV_STAGE 512u      // event records per noise batch prefetched into shared memory
#define GSTIM_EV_SMEM_MAX 2048u  // event counters / segment offsets live in shared memory up to this many noise batches

// Plan: everything the kernel needs besides the program words.
struct GstimPlan {
    uint32_t num_qubits;     // compacted qubit count Q (frame rows)
    uint32_t q_pitch;        // Q rounded up to odd (shared-memory row pitch in uint4)
    uint32_t num_meas;       // M
    uint32_t num_det;        // D
    uint32_t num_obs;        // L
    uint32_t rec_ring;       // power of two >= max lookback distance (detector mode), else 0
    uint32_t n_words;        // program length in words (multiple of chunk_words)
    uint32_t chunk_words;    // chunk size
    uint32_t n_chunks;
    uint32_t slots;          // thread groups the hazard analysis assumed (threads / lanes_per_item)
    uint32_t mode;           // 0 = detectors(+observables), 1 = measurements
    uint32_t max_items;      // largest batch
    uint32_t n_batches;
    uint32_t n_barriers;
};

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
//
// Noise is not a batch of its own when it acts on the items of a neighbouring gate batch: a Pauli-noise
// instruction whose target list equals the item list of the gate / measurement batch directly before it
// (DEPOLARIZE2 after CX, DEPOLARIZE1 after H, X_ERROR after MR ...) is ATTACHED to that batch as its
// post-noise, and one directly in front of a measurement batch as its pre-noise: the batch header names
// the RNG slices (below) whose pre-sampled events the batch applies, and the WARP that executes items
// 32 s .. 32 s + 31 applies the events of slice s right before / after its own items. No block barrier,
// no second item list. Noise that matches no neighbour is a stand-alone GOP_NOISE1 / GOP_NOISE2 batch
// with its own item list, applied the same warp-local way.
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
    GOP_NOISE1 = 4,      // stand-alone single-target Pauli noise. item: qubit
    GOP_NOISE2 = 5,      // stand-alone two-target Pauli noise. item: q1 | q2<<16
    GOP_MEASURE = 6,     // aux = basis | kind<<2. item: qubit (physical row) | logical qubit index << 16
    GOP_RECZERO = 7,     // zero record rows rec0 .. rec0+n-1 (no payload)
    GOP_XORROWS = 8,     // out row (^)= XOR of record rows. payload: dst[n], off[n+1], idx[...]
    GOP_OBS_PAULI = 9,   // out row ^= frame component. payload per item: dst row, qubit | x<<30 | z<<31
    GOP_FEEDBACK = 10,   // frame ^= record row. payload per item: rec index, qubit | x<<30 | z<<31
    GOP_CORR = 11,       // E / ELSE_CORRELATED_ERROR: one site, payload = Pauli targets qubit | x<<30 | z<<31
    GOP_QMAP = 12,       // not executed: payload[i] = logical index of physical frame row GH_EXTRA + i
};

// header word indices
enum GstimHdr : uint32_t {
    GH_OP = 0,        // op | flags<<8 | aux<<16
    GH_N = 1,         // number of items
    GH_WORDS = 2,     // total words of this batch (header + payload)
    GH_EXTRA = 3,     // op specific (QMAP: first physical row)
    GH_CSITE0 = 4,    // measure group of the batch (Philox counter word 0 of its collapse draws)
    GH_REC0 = 5,      // absolute measurement index of item 0 (MEASURE / RECZERO / record-flipping noise)
    GH_PRE = 6,       // pre-noise application: first RNG slice | parity << 31, or GSTIM_NO_NOISE (MEASURE only)
    GH_PRE_NEXT = 7,  // the noise application that follows it in program order: first slice | log2(slices per 32 items) << 28,
                      // or GSTIM_NO_NOISE
    GH_POST = 8,      // post-noise application (for GOP_NOISE1 / GOP_NOISE2 / GOP_CORR: the batch's own noise)
    GH_POST_NEXT = 9,
    GH_PERM = 10,     // word offset (from the header) of the byte table site -> item position inside its 32-item group,
                      // 0 = identity (gate batches whose items were reordered for bank spreading)
    GH_WIDTHS = 11,   // log2(sites per slice) of the pre-noise (bits 0-3) and of the post-noise (bits 4-7)
};
#define GSTIM_NO_NOISE 0xFFFFFFFFu

// header flags (bits 8..15 of word 0)
#define GF_BARRIER 0x01u     // block barrier (of the interpreter warps) before executing this batch
#define GF_REC 0x02u         // NOISE1: every event also flips bit in record row rec0+i (heralds, M(p) noise)
#define GF_ACCUM 0x04u       // XORROWS: dst ^= value (instead of dst = value)
#define GF_RESET_FLAG 0x08u  // CORR: clear the "correlated error occurred" row first (E vs ELSE)
#define GF_TABLE 0x10u       // NOISE2: PAULI_CHANNEL_2 (15 cumulative u32 thresholds; they live in the noise schedule)
#define GF_NOFRAME 0x20u     // NOISE1: item is not a frame qubit (MPAD noise)
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

// Gap arithmetic (spec v6; all integer, restated bit for bit by oracle/philox.py):
//   E   = exp_draw_q26(gap word): -ln(v / 2^32) with v = word | 1, in units of 2^-26 nat, through a 256-entry log2 table
//         (Q26 base + forward difference) with 13-bit linear interpolation and one multiply-high by ln 2 (Q32)
//   gap = (E * INV) >> SH  as a 64-bit product: INV = floor(2^32 * m), SH = 58 - e for 1 / lambda = m * 2^e (m in [0.5, 1)),
//         lambda = -log1p(-p); floor(Exp(1) / lambda) is Geometric(p) (RareErrorIterator, probability_util.cc:33-43).
//         p >= 1: INV = 0 (an event at every shot). p below 2^-58 is treated as 0.
#define GSTIM_LN2_Q32 2977044472u

// Noise schedule. The sites of a noise group (one noise instruction, or one run of it without a repeated qubit), in
// target order, are cut into SLICES of 2^w consecutive sites; a slice x a shot block is one Bernoulli sequence
// (site-major, then shot) walked with geometric gaps drawn from the slice's own Philox stream, like the reference's
// RareErrorIterator over targets x shots. The slice width follows the probability, so that a slice holds a bounded
// number of events however dense the noise: w = 5 (GSTIM_NOISE_SLICE = 32 sites) for p < 2^-6, one less for every
// doubling of p, w = 0 (one site per slice) from p >= 1/4 (gstim_slice_width_log2). Slices are numbered in program
// order; a noise application (the pre- / post-noise of a batch, or a stand-alone noise batch) over n items owns
// ceil(n / 2^w) consecutive slices, 32 >> w of them per group of 32 items.
//   slice descriptor (8 words): noise group, slice index in the group, rate index | number of sites << 16,
//                    op | flags << 8 | aux << 16 of the noise, T1, T2, T3 (NOISE1 thresholds; NOISE2 + GF_TABLE: T1 = word
//                    offset of the 15 PAULI_CHANNEL_2 thresholds in the schedule's table area), spare
//   rate (2 words):  INV, SH
// The kernel's producer warps walk the slices of shot block r + 1 while block r is interpreted and leave, per slice, one
// 128-byte LINE in this CTA's event buffer: word 0 = number of events, words 1..31 = the first 31 event records; further
// records go to the slice's overflow segment behind the lines.
//   record = shot (bits 0-11) | site in the slice (12-16) | flips x1,z1,x2,z2 (17-20) | record flip (21) | conflict (22).
#define GSTIM_NOISE_SLICE 32u
#define GSTIM_SLICE_WORDS 8u
#ifdef __cplusplus
// log2 of the sites per slice for event probability p (already narrowed to float): 2^-6 > p -> 5, ..., p >= 2^-2 -> 0.
static inline uint32_t gstim_slice_width_log2(double p) {
    uint32_t w = 5;
    for (double t = 1.0 / 64; w > 0 && p >= t; t *= 2) {
        w--;
    }
    return w;
}
#endif
enum GstimSliceWord : uint32_t {
    GSL_GROUP = 0,
    GSL_INDEX = 1,
    GSL_RATE_SITES = 2,
    GSL_H0 = 3,
    GSL_T1 = 4,
    GSL_T2 = 5,
    GSL_T3 = 6,
    GSL_SPARE = 7,
};
#define GSTIM_RATE_SMEM_MAX 64u    // the first 64 rates are mirrored in shared memory
#define GSTIM_EV_SHOT_BITS 12u
#define GSTIM_EV_SITE_SHIFT 12u
#define GSTIM_EV_FLIP_SHIFT 17u
#define GSTIM_EV_CONFLICT 0x00400000u  // the neighbouring record of the slice flips the same 32-bit frame words: apply atomically
#define GSTIM_EV_LINE_WORDS 32u
#define GSTIM_MAX_BATCH_ITEMS 2047u

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
    uint32_t lanes_log2;     // log2(lanes per item) the hazard analysis assumed (a warp executes 32 >> lanes_log2 items)
    uint32_t n_slices;       // RNG slices (noise schedule)
    uint32_t mode;           // 0 = detectors(+observables), 1 = measurements
    uint32_t max_items;      // largest batch
    uint32_t n_batches;
    uint32_t n_barriers;
};

// kernels.cuh — launch parameter blocks shared between kernels.cu and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace gstim {

struct InterpParams {
    const uint32_t *prog;      // lowered program in global memory (16-byte aligned, n_chunks*chunk_words words)
    uint32_t n_chunks;
    uint32_t chunk_words;
    uint32_t Q;                // compact qubit count
    uint32_t q_pitch;          // shared-memory row pitch (uint4 units), odd
    uint32_t K;                // 128-shot columns per thread block
    uint32_t G_log2;           // log2(lanes per item)
    uint32_t slots;            // threads_interp >> G_log2
    uint32_t threads_interp;   // interpreter threads = the first warps of the block; the remaining warps produce noise events
    uint32_t n_blocks;         // shot blocks in this launch
    uint64_t col0_base;        // global column index (shot/128) of block 0
    uint32_t seed_lo, seed_hi; // Philox key
    uint4 *rec;                // measurement record rows
    uint64_t rec_block_stride; // uint4 units between consecutive shot blocks (measurement mode: rows live in the table)
    uint64_t rec_cta_stride;   // uint4 units between CTAs (detector mode: per-CTA L2-resident ring)
    uint64_t rec_k_stride;     // uint4 units between the 128-shot columns of a record row (rows of one column are contiguous)
    uint32_t rec_mask;         // ring mask (detector mode) or 0xFFFFFFFF
    uint4 *out;                // detector/observable table, column-major: out[column * out_k_stride + row]; block g owns columns [g*K,(g+1)*K)
    uint64_t out_k_stride;     // uint4 units between columns (= number of rows)
    uint32_t dbg_flags;        // GSTIM_DEBUG_FLAGS: timing experiments only (results become wrong): bit0 no noise events, bits 8+op skip opcode
    unsigned long long *dbg_cycles;  // optional (GSTIM_DEBUG_CYCLES=1): [op] cycles and [16+op] batch counts seen by block 0
    // noise schedule (program.h) and the per-CTA event scratch the pre-pass fills
    uint32_t n_noise;                 // noise batches
    uint32_t n_rates;                 // distinct rates
    const uint32_t *noise_info;       // n_noise * GSTIM_NOISE_INFO_WORDS
    const ulonglong2 *rates;          // per rate: lam, floor((2^64 - 1) / lam)
    const uint4 *slices;              // RNG slices, 2 uint4 each (program.h "Noise schedule")
    uint32_t n_slices;
    const uint32_t *ev_segoff;        // n_noise + 1 : event segment offsets for this launch's block size
    uint32_t *ev_counts;              // gridDim.x * 2 * n_noise (two event buffers per CTA)
    uint32_t *ev_buf;                 // gridDim.x * 2 * ev_segoff[n_noise]
    uint32_t *ev_overflow;            // set to 1 if any segment overflowed (host retries with more room)
};

// Shared memory the interpreter needs for (Q, K, chunk_words).
size_t interp_smem_bytes(uint32_t q_pitch, uint32_t Q, uint32_t K, uint32_t chunk_words, uint32_t n_noise);
cudaError_t launch_interp(const InterpParams &p, uint32_t grid, uint32_t threads, size_t smem, cudaStream_t stream);
cudaError_t interp_set_max_smem(size_t smem);
int interp_max_blocks_per_sm(uint32_t threads, size_t smem);

struct TransposeParams {
    const uint32_t *table;     // column-major bit table: uint4 table[column * n_rows + row], column = 128 shots
    uint64_t n_rows;           // rows per column
    const uint32_t *row_map;   // n_bits entries: source row | invert<<31
    uint32_t n_bits;           // bits per shot in the output
    uint64_t n_shots;          // shots to emit (may be less than row_words*32)
    uint8_t *out;              // dense shot-major output
    uint64_t out_pitch;        // bytes per shot in `out` (>= ceil(n_bits/8))
};
cudaError_t launch_transpose_b8(const TransposeParams &p, cudaStream_t stream);

// flip counts of the output bits (rows through row_map, bit 31 = invert; null = identity) over the first n_shots shots:
// single[n_bits] and (optional) pair[n_bits - 1] = counts of bit j AND bit j + 1; both uint64, accumulated
cudaError_t launch_bit_counts(const uint32_t *table, uint64_t n_rows, uint64_t n_shots, const uint32_t *row_map, uint32_t n_bits,
                              unsigned long long *single, unsigned long long *pair, cudaStream_t stream);

// LOP3 microbenchmark on the current device: 32-bit three-input logic lane-operations per clock per SM, per second
// (whole chip) and the SM clock the probe ran at (MHz, from clock64 span / event time).
cudaError_t measure_lop3_peak(int num_sms, double *lane_ops_per_clk_per_sm, double *lane_ops_per_sec, double *sm_mhz);

}  // namespace gstim

// sparse.cuh — host interface of the event-driven engine (sparse.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "response.h"

namespace gstim {

// Owns the device copy of a response table and launches the sampling kernel. Not thread-safe (like the sampler handle).
class SparseEngine {
  public:
    // slice_events: expected events per slice and tile (part of the random stream's definition; 0 = automatic)
    // tile_buffers: tile images a block should be able to hold (decides the tile height; 0 = default 3)
    // compress_mb: tables larger than this many MB are stored periodically (per class: head + one period + tail) when the
    // circuit has a round structure
    SparseEngine(ResponseTable &&rt, uint32_t mode, uint32_t D, uint32_t L, uint32_t M, int device, uint32_t slice_events,
                 uint32_t tile_buffers, uint32_t compress_mb);
    uint64_t device_table_entries() const;  // entries stored on the device (after folding)
    ~SparseEngine();
    SparseEngine(const SparseEngine &) = delete;
    SparseEngine &operator=(const SparseEngine &) = delete;

    const ResponseTable &table() const;
    uint32_t tile_shots() const;     // S: shots per tile (power of two <= 128); shot offsets must be multiples of it
    uint32_t blocks_per_sm() const;
    const std::vector<uint32_t> &slices() const;  // 4 words per slice: class, trials, first table entry, 0

    // Output layout: flags = GSTIM_PREPEND_OBS | GSTIM_APPEND_OBS | GSTIM_SEPARATE_OBS (detector mode). Re-encodes the
    // device table when the layout changes (synchronises `stream`).
    void set_layout(uint32_t flags, cudaStream_t stream);
    uint32_t main_bits() const;
    uint32_t obs_bits() const;
    // Measurement mode: packed reference sample every row starts from (null / all zero: rows start at zero).
    void set_reference_row(const uint8_t *packed, size_t n_bytes);

    // Samples shots [first_shot, first_shot + n_shots) (global indices; first_shot a multiple of tile_shots()) into dense
    // device rows. Pitches of 0 mean dense rows. Asynchronous on `stream`.
    void launch(uint64_t first_shot, uint64_t n_shots, uint8_t *main_out, uint64_t main_pitch, uint8_t *obs_out, uint64_t obs_pitch,
                uint64_t seed, cudaStream_t stream);

  private:
    struct Impl;
    Impl *impl;
};

// single[b] += number of shots with bit b set, pair[b] += bit b AND bit b + 1 (pair may be null), over dense b8 rows
cudaError_t launch_count_b8(const uint8_t *rows, uint64_t pitch, uint64_t n_shots, uint32_t n_bits, unsigned long long *single,
                            unsigned long long *pair, cudaStream_t stream);

// Dense b8 rows -> per-shot records of their non-zero bytes (u16 count, u16 offsets, u8 values; sparse.cu "Sparse host
// delivery"): index[shot] = byte offset of the record in stream_buf, *cursor = bytes used (start it at 0), *overflow = 1 when a
// record did not fit in `capacity`. row_bytes must be below 65536.
cudaError_t launch_compress_rows(const uint8_t *rows, uint64_t pitch, uint32_t row_bytes, uint64_t n_shots, uint8_t *stream_buf,
                                 uint64_t capacity, unsigned long long *cursor, unsigned long long *index, uint32_t *overflow,
                                 cudaStream_t stream);

}  // namespace gstim

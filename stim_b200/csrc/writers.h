// writers.h — host-side encoders for Stim's result formats, fed from the dense shot-major b8
// rows the device transposer produces (or, for ptb64, from the device's bit-major rows).
// Byte-for-byte behaviour of the reference's MeasureRecordWriterFormat{01,B8,Hits,R8,Dets}
// (/root/reference/src/stim/io/measure_record_writer.cc:61-211) and of the ptb64 branch of
// write_table_data (/root/reference/src/stim/io/measure_record_writer.h:122-135);
// formats are specified in /root/reference/doc/result_formats.md.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>

namespace gstim {

enum class Format { F01, B8, R8, HITS, DETS, PTB64 };

// Throws std::invalid_argument for an unknown name.
Format parse_format(const char *name);

// rows: n_shots rows of `pitch` bytes, bit k of a shot at rows[shot*pitch + k/8] >> (k%8). Encoded by up to 16 host threads
// (slices of rows into per-thread buffers, written in order); the sparse formats scan for set bits eight bytes at a time.
// prefix1 is used for bits [0, transition), prefix2 for [transition, n_bits) (DETS format only).
void write_shots(
    FILE *f, const uint8_t *rows, size_t pitch, size_t n_shots, size_t n_bits, Format fmt, char prefix1, char prefix2, size_t transition);

// ptb64 from packed shot-major rows (n_shots a multiple of 64): per group of 64 shots one little-endian u64 per bit.
void write_ptb64_from_rows(FILE *f, const uint8_t *rows, size_t pitch, size_t n_shots, size_t n_bits);

// ptb64 from bit-major 32-bit rows: for each group of 64 shots, for each output bit, one u64.
// row_map[bit] = source row | invert<<31. n_shots must be a multiple of 64.
void write_ptb64(FILE *f, const uint32_t *table, size_t n_rows, const uint32_t *row_map, size_t n_bits, size_t n_shots);

}  // namespace gstim

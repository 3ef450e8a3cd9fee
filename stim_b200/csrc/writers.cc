// writers.cc — see writers.h.
#include "writers.h"

#include <cstring>
#include <stdexcept>
#include <vector>

namespace gstim {

Format parse_format(const char *name) {
    std::string s = name ? name : "";
    if (s == "01") return Format::F01;
    if (s == "b8") return Format::B8;
    if (s == "r8") return Format::R8;
    if (s == "hits") return Format::HITS;
    if (s == "dets") return Format::DETS;
    if (s == "ptb64") return Format::PTB64;
    throw std::invalid_argument("Unrecognized result format '" + s + "'. Expected one of 01,b8,r8,hits,dets,ptb64.");
}

namespace {
struct Buf {
    FILE *f;
    std::vector<char> b;
    explicit Buf(FILE *f) : f(f) {
        b.reserve(1 << 20);
    }
    void put(char c) {
        b.push_back(c);
    }
    void put_uint(unsigned long long v) {
        char tmp[24];
        int n = 0;
        do {
            tmp[n++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (n) {
            b.push_back(tmp[--n]);
        }
    }
    void maybe_flush() {
        if (b.size() >= (1 << 20) - 4096) {
            flush();
        }
    }
    void flush() {
        if (!b.empty()) {
            if (fwrite(b.data(), 1, b.size(), f) != b.size()) {
                throw std::runtime_error("Failed to write result data.");
            }
            b.clear();
        }
    }
};
inline bool bit_of(const uint8_t *row, size_t k) {
    return (row[k >> 3] >> (k & 7)) & 1;
}
}  // namespace

void write_shots(
    FILE *f, const uint8_t *rows, size_t pitch, size_t n_shots, size_t n_bits, Format fmt, char prefix1, char prefix2, size_t transition) {
    Buf out(f);
    if (transition > n_bits) {
        transition = n_bits;
    }
    for (size_t s = 0; s < n_shots; s++) {
        const uint8_t *row = rows + s * pitch;
        switch (fmt) {
            case Format::F01:
                for (size_t k = 0; k < n_bits; k++) {
                    out.put(bit_of(row, k) ? '1' : '0');
                }
                out.put('\n');
                break;
            case Format::B8: {
                size_t full = n_bits >> 3;
                out.b.insert(out.b.end(), (const char *)row, (const char *)row + full);
                if (n_bits & 7) {
                    out.put((char)(row[full] & ((1u << (n_bits & 7)) - 1)));
                }
                break;
            }
            case Format::R8: {
                // each byte = number of 0s before the next 1 (0xFF = 255 zeros, no 1); a final byte
                // closes the shot as if a 1 followed the end (measure_record_writer.cc:133-169).
                unsigned run = 0;
                for (size_t k = 0; k < n_bits; k++) {
                    if (bit_of(row, k)) {
                        out.put((char)run);
                        run = 0;
                    } else if (++run == 255) {
                        out.put((char)255);
                        run = 0;
                    }
                }
                out.put((char)run);
                break;
            }
            case Format::HITS: {
                bool first = true;
                for (size_t k = 0; k < n_bits; k++) {
                    if (bit_of(row, k)) {
                        if (!first) {
                            out.put(',');
                        }
                        first = false;
                        out.put_uint(k);
                    }
                }
                out.put('\n');
                break;
            }
            case Format::DETS: {
                out.put('s');
                out.put('h');
                out.put('o');
                out.put('t');
                for (size_t k = 0; k < n_bits; k++) {
                    if (bit_of(row, k)) {
                        out.put(' ');
                        if (k < transition) {
                            out.put(prefix1);
                            out.put_uint(k);
                        } else {
                            out.put(prefix2);
                            out.put_uint(k - transition);
                        }
                    }
                }
                out.put('\n');
                break;
            }
            case Format::PTB64:
                throw std::logic_error("ptb64 is written from bit-major rows (write_ptb64).");
        }
        out.maybe_flush();
    }
    out.flush();
}

void write_ptb64(FILE *f, const uint32_t *table, size_t n_rows, const uint32_t *row_map, size_t n_bits, size_t n_shots) {
    if (n_shots % 64 != 0) {
        throw std::invalid_argument("shots must be a multiple of 64 to use ptb64 format.");
    }
    std::vector<uint64_t> line(n_bits);
    for (size_t g = 0; g < n_shots / 64; g++) {
        for (size_t m = 0; m < n_bits; m++) {
            uint32_t rm = row_map[m];
            // column-major table: uint4 table[column * n_rows + row], 64-shot group g = half of column g / 2
            const uint32_t *r = table + ((g >> 1) * n_rows + (size_t)(rm & 0x7FFFFFFFu)) * 4 + 2 * (g & 1);
            uint64_t v = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
            if (rm >> 31) {
                v = ~v;
            }
            line[m] = v;
        }
        if (fwrite(line.data(), 8, n_bits, f) != n_bits) {
            throw std::runtime_error("Failed to write result data.");
        }
    }
}

}  // namespace gstim

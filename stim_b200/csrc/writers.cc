// writers.cc — see writers.h.
#include "writers.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <thread>
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

// Output of one encoder thread: a byte array with a write cursor; room() guarantees headroom for a whole row up front, so the
// per-character appends are plain stores.
struct Buf {
    std::vector<char> b;
    size_t n = 0;
    void room(size_t extra) {
        if (b.size() < n + extra) {
            b.resize(std::max(b.size() * 2, n + extra));
        }
    }
    void put(char c) {
        b[n++] = c;
    }
    void put8(uint64_t packed, unsigned len) {  // up to 8 characters packed little-endian
        memcpy(b.data() + n, &packed, 8);
        n += len;
    }
};

// decimal strings of 0 .. count - 1, packed little-endian into 7 bytes + length in the top byte (indices below 10^7)
std::vector<uint64_t> decimal_table(size_t count) {
    std::vector<uint64_t> t(count);
    for (size_t k = 0; k < count; k++) {
        char tmp[8];
        int len = 0;
        size_t v = k;
        do {
            tmp[len++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        uint64_t w = 0;
        for (int i = 0; i < len; i++) {
            w |= (uint64_t)(uint8_t)tmp[len - 1 - i] << (8 * i);
        }
        t[k] = w | ((uint64_t)len << 56);
    }
    return t;
}

// f(k) for every set bit k < n_bits of a packed row, in increasing order; eight bytes per probe (the rows of a QEC experiment
// are ~98 % zeros, so the sparse formats cost a few hundred steps per shot instead of one per bit)
template <typename F>
inline void for_each_set_bit(const uint8_t *row, size_t n_bits, F &&f) {
    const size_t n_bytes = (n_bits + 7) / 8, full = n_bytes / 8;
    for (size_t i = 0; i < full; i++) {
        uint64_t w;
        memcpy(&w, row + 8 * i, 8);  // (fixed size: one load)
        if (i * 64 + 64 > n_bits) {
            w &= (1ull << (n_bits - i * 64)) - 1;
        }
        while (w) {
            f(i * 64 + (size_t)__builtin_ctzll(w));
            w &= w - 1;
        }
    }
    if (full * 8 < n_bytes) {
        uint64_t w = 0;
        for (size_t b = full * 8; b < n_bytes; b++) {
            w |= (uint64_t)row[b] << (8 * (b - full * 8));
        }
        const size_t valid = n_bits - full * 64;  // 1..63
        w &= (1ull << valid) - 1;
        while (w) {
            f(full * 64 + (size_t)__builtin_ctzll(w));
            w &= w - 1;
        }
    }
}

const uint64_t *lut01() {  // byte -> eight '0' / '1' characters, bit 0 first
    static const struct Lut {
        uint64_t v[256];
        Lut() {
            for (int x = 0; x < 256; x++) {
                uint64_t w = 0;
                for (int k = 0; k < 8; k++) {
                    w |= (uint64_t)('0' + ((x >> k) & 1)) << (8 * k);
                }
                v[x] = w;
            }
        }
    } lut;
    return lut.v;
}

void encode_rows(Buf &out, const uint8_t *rows, size_t pitch, size_t n_shots, size_t n_bits, Format fmt, char prefix1, char prefix2,
                 size_t transition, const uint64_t *dec) {
    const uint64_t *lut = lut01();
    // the most one row can take: 01: one character per bit; r8: one byte per bit + 1; hits / dets: separator, prefix, <= 7 digits
    const size_t worst = fmt == Format::HITS || fmt == Format::DETS ? n_bits * 9 + 16 : n_bits + 16;
    auto put_index = [&](size_t k) {
        const uint64_t w = dec[k];
        out.put8(w & 0x00FFFFFFFFFFFFFFull, (unsigned)(w >> 56));
    };
    for (size_t s = 0; s < n_shots; s++) {
        const uint8_t *row = rows + s * pitch;
        out.room(worst);
        switch (fmt) {
            case Format::F01: {
                const size_t full = n_bits >> 3;
                char *dst = out.b.data() + out.n;
                out.n += n_bits + 1;
                for (size_t i = 0; i < full; i++) {
                    memcpy(dst + 8 * i, &lut[row[i]], 8);
                }
                for (size_t k = full * 8; k < n_bits; k++) {
                    dst[k] = (char)('0' + ((row[k >> 3] >> (k & 7)) & 1));
                }
                dst[n_bits] = '\n';
                break;
            }
            case Format::B8: {
                size_t full = n_bits >> 3;
                memcpy(out.b.data() + out.n, row, full);
                out.n += full;
                if (n_bits & 7) {
                    out.put((char)(row[full] & ((1u << (n_bits & 7)) - 1)));
                }
                break;
            }
            case Format::R8: {
                // each byte = number of 0s before the next 1 (0xFF = 255 zeros, no 1); a final byte
                // closes the shot as if a 1 followed the end (measure_record_writer.cc:133-169).
                size_t next = 0;  // first position not yet accounted for
                auto gap = [&](size_t zeros) {
                    while (zeros >= 255) {
                        out.put((char)255);
                        zeros -= 255;
                    }
                    out.put((char)zeros);
                };
                for_each_set_bit(row, n_bits, [&](size_t k) {
                    gap(k - next);
                    next = k + 1;
                });
                gap(n_bits - next);
                break;
            }
            case Format::HITS: {
                bool first = true;
                for_each_set_bit(row, n_bits, [&](size_t k) {
                    if (!first) {
                        out.put(',');
                    }
                    first = false;
                    put_index(k);
                });
                out.put('\n');
                break;
            }
            case Format::DETS: {
                out.put('s');
                out.put('h');
                out.put('o');
                out.put('t');
                for_each_set_bit(row, n_bits, [&](size_t k) {
                    out.put(' ');
                    if (k < transition) {
                        out.put(prefix1);
                        put_index(k);
                    } else {
                        out.put(prefix2);
                        put_index(k - transition);
                    }
                });
                out.put('\n');
                break;
            }
            case Format::PTB64:
                throw std::logic_error("ptb64 is written from whole groups of 64 rows (write_ptb64_from_rows).");
        }
    }
}

void transpose64(uint64_t *a) {  // bit c of a[r] <-> bit r of a[c]
    uint64_t m = 0x00000000FFFFFFFFull;
    for (unsigned j = 32; j != 0; j >>= 1, m ^= m << j) {
        for (unsigned k = 0; k < 64; k = (k + j + 1) & ~j) {
            const uint64_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= t << j;
            a[k + j] ^= t;
        }
    }
}

// Encodes [0, n_units) with up to 16 threads, `batch` units at a time, and writes the pieces in order.
template <typename ENC>
void encode_parallel(FILE *f, size_t n_units, size_t batch, ENC &&enc) {
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    static thread_local std::vector<Buf> bufs;  // kept between calls: a stream of chunks reuses the grown arrays
    if (bufs.size() < hw) {
        bufs.resize(hw);
    }
    for (size_t u0 = 0; u0 < n_units; u0 += batch) {
        const size_t cnt = std::min(batch, n_units - u0);
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(hw, cnt / 16));
        const size_t per = (cnt + nt - 1) / nt;
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < nt; t++) {
            const size_t a = std::min(cnt, t * per), b = std::min(cnt, a + per);
            bufs[t].n = 0;
            if (a >= b) {
                continue;
            }
            if (nt == 1) {
                enc(bufs[t], u0 + a, u0 + b);
            } else {
                Buf *bt = &bufs[t];
                ts.emplace_back([&enc, bt, u0, a, b] { enc(*bt, u0 + a, u0 + b); });
            }
        }
        for (auto &t : ts) {
            t.join();
        }
        for (unsigned t = 0; t < nt; t++) {
            if (bufs[t].n != 0 && fwrite(bufs[t].b.data(), 1, bufs[t].n, f) != bufs[t].n) {
                throw std::runtime_error("Failed to write result data.");
            }
        }
    }
}
}  // namespace

void write_shots(
    FILE *f, const uint8_t *rows, size_t pitch, size_t n_shots, size_t n_bits, Format fmt, char prefix1, char prefix2, size_t transition) {
    if (transition > n_bits) {
        transition = n_bits;
    }
    if (fmt == Format::PTB64) {
        write_ptb64_from_rows(f, rows, pitch, n_shots, n_bits);
        return;
    }
    // batches bound the encoders' buffers (01: n_bits + 1 characters per shot)
    const size_t batch = std::max<size_t>(64, (size_t)(256u << 20) / std::max<size_t>(n_bits + 1, 1));
    static thread_local std::vector<uint64_t> dec;
    if ((fmt == Format::HITS || fmt == Format::DETS) && dec.size() < n_bits) {
        if (n_bits > 9999999) {
            throw std::invalid_argument("more than 10^7 bits per shot are not supported by the hits / dets writers");
        }
        dec = decimal_table(n_bits);
    }
    const uint64_t *dec_p = dec.data();
    encode_parallel(f, n_shots, batch, [&](Buf &out, size_t a, size_t b) {
        encode_rows(out, rows + a * pitch, pitch, b - a, n_bits, fmt, prefix1, prefix2, transition, dec_p);
    });
}

void write_ptb64_from_rows(FILE *f, const uint8_t *rows, size_t pitch, size_t n_shots, size_t n_bits) {
    if (n_shots % 64 != 0) {
        throw std::invalid_argument("shots must be a multiple of 64 to use ptb64 format.");
    }
    // a group of 64 shots = n_bits little-endian 64-bit words, word m holding bit m of the 64 shots: 64 x 64 bit blocks of the
    // rows, transposed
    const size_t n_bytes = (n_bits + 7) / 8;
    encode_parallel(f, n_shots / 64, 4096, [&](Buf &out, size_t g0, size_t g1) {
        uint64_t blk[64];
        for (size_t g = g0; g < g1; g++) {
            out.room(n_bits * 8);
            char *dst = out.b.data() + out.n;
            out.n += n_bits * 8;
            for (size_t w = 0; w * 64 < n_bits; w++) {
                const size_t take = n_bytes - w * 8 < 8 ? n_bytes - w * 8 : 8;
                for (size_t r = 0; r < 64; r++) {
                    blk[r] = 0;
                    memcpy(&blk[r], rows + (g * 64 + r) * pitch + w * 8, take);
                }
                transpose64(blk);
                const size_t cols = n_bits - w * 64 < 64 ? n_bits - w * 64 : 64;
                memcpy(dst + w * 64 * 8, blk, cols * 8);
            }
        }
    });
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

// hostpipe.h — moving packed rows between caller (host) memory and device staging without stalling on pageable copies.
// A cudaMemcpy from / to pageable memory is a synchronous, single-threaded bounce through the driver's own staging (~6-10 GB/s);
// here rows go through two page-locked buffers in sub-chunks: the DMA of sub-chunk k overlaps the threaded memcpy of sub-chunk
// k - 1 (up to 16 host threads). Page-locked caller memory is used directly. Used by the measurement converter (m2d.cu) and the
// detector-error-model sampler (dem.cu); the bulk samplers have their own pipeline in api.cu.
#pragma once
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace gstim {

inline void hp_ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error(std::string("CUDA: ") + what + ": " + cudaGetErrorString(e));
    }
}

inline bool hp_is_pinned(const void *ptr) {
    if (ptr == nullptr) {
        return false;
    }
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// A large result array that host threads are about to touch for the first time: ask for transparent huge pages, so that the
// kernel zeroes and maps 2 MiB at a time instead of 4 KiB (a hint; ignored where THP is off or the range is already mapped).
inline void hp_hugepage_hint(void *ptr, uint64_t bytes) {
    if (ptr == nullptr || bytes < (64ull << 20)) {
        return;
    }
    const uintptr_t a = (reinterpret_cast<uintptr_t>(ptr) + 4095) & ~(uintptr_t)4095;
    const uintptr_t b = (reinterpret_cast<uintptr_t>(ptr) + bytes) & ~(uintptr_t)4095;
    if (b > a) {
        madvise(reinterpret_cast<void *>(a), b - a, MADV_HUGEPAGE);
    }
}

// f(i0, i1) over [0, n) on up to 16 threads (at least `grain` rows per thread)
template <typename F>
void hp_parallel(uint64_t n, uint64_t grain, F &&f) {
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(hw, n / std::max<uint64_t>(grain, 1)));
    if (nt <= 1) {
        f((uint64_t)0, n);
        return;
    }
    std::vector<std::thread> ts;
    const uint64_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const uint64_t a = std::min(n, t * per), b = std::min(n, a + per);
        if (a < b) {
            ts.emplace_back([&f, a, b] { f(a, b); });
        }
    }
    for (auto &t : ts) {
        t.join();
    }
}

struct HostStager {
    uint8_t *buf[2] = {nullptr, nullptr};
    size_t cap = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    HostStager() = default;
    HostStager(const HostStager &) = delete;
    HostStager &operator=(const HostStager &) = delete;
    ~HostStager() {
        for (int i = 0; i < 2; i++) {
            if (buf[i]) {
                cudaFreeHost(buf[i]);
            }
            if (ev[i]) {
                cudaEventDestroy(ev[i]);
            }
        }
    }
    void ensure(size_t bytes) {
        for (int i = 0; i < 2; i++) {
            if (!ev[i]) {
                hp_ck(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming), "cudaEventCreate");
            }
        }
        if (bytes <= cap) {
            return;
        }
        for (int i = 0; i < 2; i++) {
            if (buf[i]) {
                cudaFreeHost(buf[i]);
                buf[i] = nullptr;
            }
            hp_ck(cudaHostAlloc((void **)&buf[i], bytes, cudaHostAllocDefault), "cudaHostAlloc (staging)");
        }
        cap = bytes;
    }
    void wait(int i) {
        if (busy[i]) {
            hp_ck(cudaEventSynchronize(ev[i]), "cudaEventSynchronize");
            busy[i] = false;
        }
    }
};

constexpr uint64_t HP_SUB_BYTES = 32ull << 20;  // staging sub-chunk

// If a consumer throws while copies into / out of the staging buffers are still in flight, wait for them before the exception
// leaves (the buffers may be freed or reused by whoever handles it).
struct HpDrainOnUnwind {
    cudaStream_t st;
    HostStager &S;
    int live;
    HpDrainOnUnwind(cudaStream_t st, HostStager &S) : st(st), S(S), live(std::uncaught_exceptions()) {}
    ~HpDrainOnUnwind() {
        if (std::uncaught_exceptions() > live) {
            cudaStreamSynchronize(st);
            S.busy[0] = S.busy[1] = false;
        }
    }
};

// Device rows (dense, `pitch` bytes apart) -> rowfn(staged_row, i) for every row i in [0, n), the DMA of one sub-chunk
// overlapping the host work on the previous one. The device rows must stay valid until the stream has passed the copies.
template <typename ROWFN>
void hp_staged_d2h(HostStager &S, cudaStream_t st, const uint8_t *d, uint64_t pitch, uint64_t n, ROWFN &&rowfn) {
    if (n == 0 || pitch == 0) {
        return;
    }
    const uint64_t sub = std::max<uint64_t>(1, HP_SUB_BYTES / pitch);
    S.ensure(sub * pitch);
    HpDrainOnUnwind guard(st, S);
    const uint64_t n_sub = (n + sub - 1) / sub;
    auto process = [&](uint64_t k) {
        const uint64_t r0 = k * sub, cnt = std::min(sub, n - r0);
        S.wait((int)(k & 1));
        const uint8_t *src = S.buf[k & 1];
        hp_parallel(cnt, 256, [&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; i++) {
                rowfn(src + i * pitch, r0 + i);
            }
        });
    };
    for (uint64_t k = 0; k < n_sub; k++) {
        const uint64_t r0 = k * sub, cnt = std::min(sub, n - r0);
        S.wait((int)(k & 1));
        hp_ck(cudaMemcpyAsync(S.buf[k & 1], d + r0 * pitch, cnt * pitch, cudaMemcpyDeviceToHost, st), "D2H");
        hp_ck(cudaEventRecord(S.ev[k & 1], st), "cudaEventRecord");
        S.busy[k & 1] = true;
        if (k > 0) {
            process(k - 1);
        }
    }
    process(n_sub - 1);
}

// The same, handing whole staged sub-chunks to blockfn(rows, first_row, count) in order (for consumers that parallelise over
// the rows themselves, e.g. the result-format encoders).
template <typename BLOCKFN>
void hp_staged_d2h_blocks(HostStager &S, cudaStream_t st, const uint8_t *d, uint64_t pitch, uint64_t n, uint64_t align_rows, BLOCKFN &&blockfn) {
    if (n == 0 || pitch == 0) {
        return;
    }
    uint64_t sub = std::max<uint64_t>(1, HP_SUB_BYTES / pitch);
    sub = std::max<uint64_t>(align_rows, sub / align_rows * align_rows);
    S.ensure(sub * pitch);
    HpDrainOnUnwind guard(st, S);
    const uint64_t n_sub = (n + sub - 1) / sub;
    auto process = [&](uint64_t k) {
        const uint64_t r0 = k * sub, cnt = std::min(sub, n - r0);
        S.wait((int)(k & 1));
        blockfn((const uint8_t *)S.buf[k & 1], r0, cnt);
    };
    for (uint64_t k = 0; k < n_sub; k++) {
        const uint64_t r0 = k * sub, cnt = std::min(sub, n - r0);
        S.wait((int)(k & 1));
        hp_ck(cudaMemcpyAsync(S.buf[k & 1], d + r0 * pitch, cnt * pitch, cudaMemcpyDeviceToHost, st), "D2H");
        hp_ck(cudaEventRecord(S.ev[k & 1], st), "cudaEventRecord");
        S.busy[k & 1] = true;
        if (k > 0) {
            process(k - 1);
        }
    }
    process(n_sub - 1);
}

// Caller rows (row_bytes each, src_pitch apart) -> dense device rows (row_bytes apart), through the staging buffers unless the
// caller's memory is page-locked.
inline void hp_h2d_rows(HostStager &S, cudaStream_t st, const uint8_t *h, uint64_t src_pitch, uint8_t *d, uint64_t row_bytes, uint64_t n) {
    if (n == 0 || row_bytes == 0) {
        return;
    }
    if (hp_is_pinned(h)) {
        hp_ck(cudaMemcpy2DAsync(d, row_bytes, h, src_pitch, row_bytes, n, cudaMemcpyHostToDevice, st), "H2D");
        return;
    }
    const uint64_t sub = std::max<uint64_t>(1, HP_SUB_BYTES / row_bytes);
    S.ensure(sub * row_bytes);
    HpDrainOnUnwind guard(st, S);
    for (uint64_t k = 0, r0 = 0; r0 < n; k++, r0 += sub) {
        const uint64_t cnt = std::min(sub, n - r0);
        S.wait((int)(k & 1));
        uint8_t *dst = S.buf[k & 1];
        hp_parallel(cnt, 256, [&](uint64_t a, uint64_t b) {
            if (src_pitch == row_bytes) {
                memcpy(dst + a * row_bytes, h + (r0 + a) * src_pitch, (b - a) * row_bytes);
            } else {
                for (uint64_t i = a; i < b; i++) {
                    memcpy(dst + i * row_bytes, h + (r0 + i) * src_pitch, row_bytes);
                }
            }
        });
        hp_ck(cudaMemcpyAsync(d + r0 * row_bytes, dst, cnt * row_bytes, cudaMemcpyHostToDevice, st), "H2D");
        hp_ck(cudaEventRecord(S.ev[k & 1], st), "cudaEventRecord");
        S.busy[k & 1] = true;
    }
}

// bits [bit0, bit0 + n_bits) of one packed row -> a packed (or one byte per bit) row of the caller
inline void hp_slice_row(const uint8_t *r, uint32_t bit0, uint32_t n_bits, bool packed, uint8_t *dst) {
    const uint64_t out_bytes = (n_bits + 7) / 8;
    if (packed && (bit0 & 7) == 0) {
        memcpy(dst, r + (bit0 >> 3), out_bytes);
        if (n_bits & 7) {
            dst[out_bytes - 1] &= (uint8_t)((1u << (n_bits & 7)) - 1);
        }
    } else if (packed) {
        memset(dst, 0, out_bytes);
        for (uint32_t b = 0; b < n_bits; b++) {
            dst[b >> 3] |= (uint8_t)(((r[(bit0 + b) >> 3] >> ((bit0 + b) & 7)) & 1) << (b & 7));
        }
    } else if ((bit0 & 7) == 0) {
        static const struct Lut {
            uint64_t v[256];
            Lut() {
                for (int x = 0; x < 256; x++) {
                    uint64_t w = 0;
                    for (int k = 0; k < 8; k++) {
                        w |= (uint64_t)((x >> k) & 1) << (8 * k);
                    }
                    v[x] = w;
                }
            }
        } lut;
        const uint8_t *src = r + (bit0 >> 3);
        const uint32_t full = n_bits / 8;
        for (uint32_t i = 0; i < full; i++) {
            memcpy(dst + 8 * i, &lut.v[src[i]], 8);
        }
        for (uint32_t b = full * 8; b < n_bits; b++) {
            dst[b] = (src[b >> 3] >> (b & 7)) & 1;
        }
    } else {
        for (uint32_t b = 0; b < n_bits; b++) {
            dst[b] = (r[(bit0 + b) >> 3] >> ((bit0 + b) & 7)) & 1;
        }
    }
}

}  // namespace gstim

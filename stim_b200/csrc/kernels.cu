// kernels.cu — sm_100a device code of the Pauli-frame sampler.
//
//   (the interpreter kernel lives in interp.cu)
//   gstim_transpose_kernel  bit-major rows -> dense shot-major b8 bytes.
//                         Replaces simd_bit_table::transposed + write_table_data
//                         (/root/reference/src/stim/io/measure_record_writer.h:101-166).
//   gstim_bitcount_kernel per-bit flip counts and adjacent-pair counts (feed the optional multi-GPU allreduce).
#include "kernels.cuh"

#include <algorithm>
#include <vector>

namespace gstim {

// ------------------------------------------------------------------------------------------------
// Output transposer. Tile = TP_SHOTS shots (TP_COLS columns) x 1024 bits; small tiles so that several thread
// blocks per SM overlap their gather / transpose / store phases.
//   1. every thread gathers TP_RPT rows x TP_COLS columns of table entries (a warp reads 512 contiguous bytes of a
//      column) through the row map and parks them in shared memory, one uint4 of padding per 32 rows;
//   2. thread (col, w, rg) picks up the 32 words (rows rg*32.., shot word w of column col), transposes the
//      32x32 bit block in registers and parks it in the (aliased) output staging area [shot][40 words];
//   3. warps stream whole 128-byte shot segments to the dense (arbitrarily aligned) output rows.
// ------------------------------------------------------------------------------------------------
constexpr int TP_COLS = 2;
constexpr int TP_SHOTS = TP_COLS * 128;
constexpr int TP_THREADS = TP_SHOTS;         // one thread per (column, shot word, 32-row group)
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int TP_RPT = 1024 / TP_THREADS;     // rows gathered per thread and column
constexpr int TP_BITS = 1024;
constexpr int TP_PITCH = 40;                       // words per shot row in the output staging (8 mod 32: conflict-free)
constexpr int TP_IN_COL = TP_BITS + TP_BITS / 32;  // uint4 per column in the input staging (1 pad per 32 rows)

__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
    // After this, a[j] bit i == (input a[i]) bit j. The 16- and 8-bit stages are byte permutes (one PRMT per
    // output word), the 4-, 2- and 1-bit stages masked swaps.
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t x = a[k], y = a[k + 16];
        a[k] = __byte_perm(x, y, 0x5410);       // (x.lo16, y.lo16)
        a[k + 16] = __byte_perm(x, y, 0x7632);  // (x.hi16, y.hi16)
    }
#pragma unroll
    for (int k = 0; k < 32; k++) {
        if ((k & 8) == 0) {
            const uint32_t x = a[k], y = a[k + 8];
            a[k] = __byte_perm(x, y, 0x6240);      // bytes x0 y0 x2 y2
            a[k + 8] = __byte_perm(x, y, 0x7351);  // bytes x1 y1 x3 y3
        }
    }
#pragma unroll
    for (int st = 2; st < 5; st++) {
        const int j = 16 >> st;
        const uint32_t m = j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k | j]) & m;
                a[k] ^= t << j;
                a[k | j] ^= t;
            }
        }
    }
}

__global__ void __launch_bounds__(TP_THREADS, 4) gstim_transpose_kernel(const TransposeParams p) {
    extern __shared__ __align__(16) uint32_t tile[];
    uint4 *in = reinterpret_cast<uint4 *>(tile);
    const uint32_t tid = threadIdx.x;
    // consecutive blocks take consecutive bit tiles of the same shots, so the 128-byte pieces of an output row are
    // written at about the same time and reach DRAM as whole rows
    const uint32_t n_bit_tiles = (p.n_bits + TP_BITS - 1) / TP_BITS;
    const uint32_t tile_y = blockIdx.x % n_bit_tiles;
    const uint64_t tile_x = blockIdx.x / n_bit_tiles;
    const uint64_t col0 = tile_x * TP_COLS;
    const uint64_t n_cols = (p.n_shots + 127) / 128;
    const uint32_t bit0 = tile_y * TP_BITS;
    const uint4 *table = reinterpret_cast<const uint4 *>(p.table);

    // ---- 1. gather
    {
        uint4 v[TP_RPT][TP_COLS];
        uint32_t inv[TP_RPT];
#pragma unroll
        for (int j = 0; j < TP_RPT; j++) {
            const uint32_t bit = bit0 + tid + TP_THREADS * j;
            const bool ok = bit < p.n_bits;
            const uint32_t rm = ok ? p.row_map[bit] : 0u;
            inv[j] = (uint32_t)0 - (rm >> 31);
#pragma unroll
            for (int c = 0; c < TP_COLS; c++) {
                v[j][c] = (ok && col0 + c < n_cols) ? table[(col0 + c) * p.n_rows + (rm & 0x7FFFFFFFu)] : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int j = 0; j < TP_RPT; j++) {
            const uint32_t r = tid + TP_THREADS * j;
#pragma unroll
            for (int c = 0; c < TP_COLS; c++) {
                uint4 x = v[j][c];
                x.x ^= inv[j];
                x.y ^= inv[j];
                x.z ^= inv[j];
                x.w ^= inv[j];
                in[c * TP_IN_COL + r + (r >> 5)] = x;
            }
        }
    }
    __syncthreads();

    // ---- 2. 32x32 bit transposes
    const uint32_t w = tid & 3, rg = (tid >> 2) & 31, col = tid >> 7;
    uint32_t a[32];
#pragma unroll
    for (int i = 0; i < 32; i++) {
        a[i] = tile[(col * TP_IN_COL + rg * 33 + i) * 4 + w];
    }
    transpose32(a);
    __syncthreads();  // the output staging aliases the input staging
    // shot col*128 + w*32 + s is parked at row s*(4 TP_COLS) + sw (sw = col*4 + w): bank = (8*(row) + rg) % 32
    const uint32_t sw = col * 4 + w;
#pragma unroll
    for (int s = 0; s < 32; s++) {
        tile[(s * (4 * TP_COLS) + sw) * TP_PITCH + rg] = a[s];
    }
    __syncthreads();

    // ---- 3. dense rows out: warp `warp` streams the 128-byte segments of shots warp, warp + 16, ...
    const uint32_t nbytes = (p.n_bits + 7) / 8;
    const uint32_t seg0 = tile_y * (TP_BITS / 8);
    const uint32_t seg_len = min((uint32_t)(TP_BITS / 8), nbytes - seg0);
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint64_t shot0 = tile_x * TP_SHOTS;
    const uint32_t n_valid = (uint32_t)min((uint64_t)TP_SHOTS, p.n_shots - shot0);
    uint8_t *dst = p.out + (shot0 + warp) * p.out_pitch + seg0;
    const uint64_t dst_step = TP_WARPS * p.out_pitch;
    if (seg_len == TP_BITS / 8) {
        // full segment: aligned word `lane` of the destination holds segment bytes [4 lane - mis, 4 lane - mis + 4).
        // Branch-free per row: one (predicated) word store per lane, and the 4 - mis head bytes / mis tail bytes as
        // single predicated byte stores from lanes 0..2 (a per-byte loop run by one lane still costs the whole warp).
        const uint32_t mis_step = (uint32_t)(dst_step & 3);
        uint32_t mis = (uint32_t)((uintptr_t)dst & 3);
        const uint32_t *S = &tile[(warp * (4 * TP_COLS)) * TP_PITCH];  // row of shot sl = warp + TP_WARPS * t
#pragma unroll 4
        for (uint32_t sl = warp; sl < n_valid; sl += TP_WARPS, dst += dst_step, mis = (mis + mis_step) & 3) {
            const uint32_t *R = S + (((sl - warp) & 31) * (4 * TP_COLS) + (sl >> 5)) * TP_PITCH;
            const uint32_t sh = 32 - 8 * mis;
            const uint32_t hi = R[lane], lo = R[(lane + 31) & 31], first = R[0], last = R[31];
            if (lane != 0 || mis == 0) {
                reinterpret_cast<uint32_t *>(dst - mis)[lane] = __funnelshift_rc(lo, hi, sh);  // (sh = 32 -> hi)
            }
            if (mis != 0 && lane < 4 - mis) {
                dst[lane] = (uint8_t)(first >> (8 * lane));  // segment bytes 0 .. 3 - mis
            }
            if (lane < mis) {
                dst[128 - mis + lane] = (uint8_t)(last >> (sh + 8 * lane));  // segment bytes 128 - mis .. 127
            }
        }
        return;
    }
    for (uint32_t sl = warp; sl < n_valid; sl += TP_WARPS, dst += dst_step) {
        const uint32_t *S = &tile[((sl & 31) * (4 * TP_COLS) + (sl >> 5)) * TP_PITCH];
        const uint32_t mis = (uint32_t)((uintptr_t)dst & 3);
        uint8_t *base = dst - mis;  // 4-byte aligned
        // aligned destination word w holds segment bytes [4w - mis, 4w - mis + 4)
        for (uint32_t wd = lane; wd * 4 < mis + seg_len; wd += 32) {
            const uint32_t lo = wd >= 1 ? S[wd - 1] : 0u;
            const uint32_t hi = wd < 32 ? S[wd] : 0u;
            const uint32_t val = mis == 0 ? hi : __funnelshift_r(lo, hi, 8 * (4 - mis));
            const int first = (int)(4 * wd) - (int)mis;  // segment byte index of this word's byte 0
            if (first >= 0 && first + 4 <= (int)seg_len) {
                *(uint32_t *)(base + 4 * wd) = val;
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int sb = first + b;
                    if (sb >= 0 && sb < (int)seg_len) {
                        base[4 * wd + b] = (uint8_t)(val >> (8 * b));
                    }
                }
            }
        }
    }
}

cudaError_t launch_transpose_b8(const TransposeParams &p, cudaStream_t stream) {
    if (p.n_bits == 0 || p.n_shots == 0) {
        return cudaSuccess;
    }
    const uint64_t n_tiles = ((p.n_shots + TP_SHOTS - 1) / TP_SHOTS) * ((p.n_bits + TP_BITS - 1) / TP_BITS);
    if (n_tiles >= (1ull << 31)) {
        return cudaErrorInvalidValue;
    }
    dim3 grid((unsigned)n_tiles);
    static bool attr_set = false;
    const size_t smem = std::max((size_t)TP_SHOTS * TP_PITCH * 4, (size_t)TP_COLS * TP_IN_COL * 16);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gstim_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            return e;
        }
        attr_set = true;
    }
    gstim_transpose_kernel<<<grid, TP_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Flip counts of the output bits of a column-major table: single[j] = popcount of output bit j over the first n_shots
// shots, pair[j] = popcount of (bit j AND bit j + 1) (the adjacent-pair correlations of the statistical parity tests).
// Output bit j = table row (row_map[j] & 0x7FFFFFFF), inverted when bit 31 of the map entry is set (reference sample).
// thread = output bit (a warp reads 512 contiguous bytes of a column when the map is the identity), blockIdx.y = slice
// of columns. `pair` may be null.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 masked_row(const uint4 *table, uint64_t c, uint64_t n_rows, uint32_t rm, uint64_t n_shots) {
    uint4 v = table[c * n_rows + (rm & 0x7FFFFFFFu)];
    const uint32_t inv = (uint32_t)0 - (rm >> 31);
    uint32_t wds[4] = {v.x ^ inv, v.y ^ inv, v.z ^ inv, v.w ^ inv};
    const uint64_t left = n_shots - c * 128;  // shots of this column that count
    if (left < 128) {
        for (uint32_t i = 0; i < 4; i++) {
            const uint64_t lo = 32ull * i;
            wds[i] = left <= lo ? 0u : (left - lo >= 32 ? wds[i] : wds[i] & ((1u << (left - lo)) - 1));
        }
    }
    return make_uint4(wds[0], wds[1], wds[2], wds[3]);
}

__global__ void gstim_bitcount_kernel(const uint4 *table, uint64_t n_rows, uint64_t n_shots, const uint32_t *row_map, uint32_t n_bits,
                                      unsigned long long *single, unsigned long long *pair) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_bits) {
        return;
    }
    const uint32_t rm0 = row_map ? row_map[j] : (uint32_t)j;
    const bool has_next = pair != nullptr && j + 1 < n_bits;
    const uint32_t rm1 = has_next ? (row_map ? row_map[j + 1] : (uint32_t)(j + 1)) : rm0;
    const uint64_t n_cols = (n_shots + 127) / 128;
    unsigned long long a1 = 0, a2 = 0;
    for (uint64_t c = blockIdx.y; c < n_cols; c += gridDim.y) {
        const uint4 v = masked_row(table, c, n_rows, rm0, n_shots);
        a1 += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        if (has_next) {
            const uint4 w = masked_row(table, c, n_rows, rm1, n_shots);
            a2 += __popc(v.x & w.x) + __popc(v.y & w.y) + __popc(v.z & w.z) + __popc(v.w & w.w);
        }
    }
    if (a1) {
        atomicAdd(&single[j], a1);
    }
    if (a2) {
        atomicAdd(&pair[j], a2);
    }
}

cudaError_t launch_bit_counts(const uint32_t *table, uint64_t n_rows, uint64_t n_shots, const uint32_t *row_map, uint32_t n_bits,
                              unsigned long long *single, unsigned long long *pair, cudaStream_t stream) {
    if (n_bits == 0 || n_shots == 0) {
        return cudaSuccess;
    }
    const uint64_t n_cols = (n_shots + 127) / 128;
    dim3 grid((unsigned)((n_bits + 255) / 256), (unsigned)std::min<uint64_t>(n_cols, 1024));
    gstim_bitcount_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(table), n_rows, n_shots, row_map, n_bits, single, pair);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Integer-ALU roofline probe ("lanes/clk/SM for 32-bit logic must be measured with a LOP3
// microbenchmark on the box"). Every thread runs 8 independent chains of three-input LOP3 (xor3), fully unrolled;
// blocks report their own clock64 span, the host divides the lane-operations of an SM by the longest span.
// ------------------------------------------------------------------------------------------------
constexpr int LOP3_CHAINS = 8, LOP3_UNROLL = 32;
__global__ void __launch_bounds__(1024, 2) gstim_lop3_probe_kernel(uint32_t iters, uint32_t seed, uint32_t *sink, long long *spans) {
    uint32_t v[LOP3_CHAINS];
#pragma unroll
    for (int c = 0; c < LOP3_CHAINS; c++) {
        v[c] = seed * (threadIdx.x + 1u) + c;
    }
    const uint32_t a = seed ^ 0x9E3779B9u, b = seed + blockIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < LOP3_UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < LOP3_CHAINS; c++) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(a), "r"(b + u));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < LOP3_CHAINS; c++) {
        acc ^= v[c];
    }
    if (acc == 0x12345u) {
        sink[0] = acc;  // keeps the chains alive
    }
    if (threadIdx.x == 0) {
        spans[2 * blockIdx.x] = t0;
        spans[2 * blockIdx.x + 1] = t1;
    }
}

cudaError_t measure_lop3_peak(int num_sms, double *lane_ops_per_clk_per_sm, double *lane_ops_per_sec, double *sm_mhz) {
    const uint32_t iters = 4096, grid = (uint32_t)num_sms * 2;
    uint32_t *sink = nullptr;
    long long *spans = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&sink, 16)) != cudaSuccess) {
        return e;
    }
    if ((e = cudaMalloc(&spans, (size_t)grid * 16)) != cudaSuccess) {
        cudaFree(sink);
        return e;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best_per_clk = 0, best_per_sec = 0, mhz = 0;
    std::vector<long long> h((size_t)grid * 2);
    for (int rep = 0; rep < 5 && e == cudaSuccess; rep++) {
        cudaEventRecord(e0);
        gstim_lop3_probe_kernel<<<grid, 1024>>>(iters, 12345u + rep, sink, spans);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) {
            break;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h.data(), spans, (size_t)grid * 16, cudaMemcpyDeviceToHost);
        long long longest = 1;
        for (uint32_t b = 0; b < grid; b++) {
            longest = std::max(longest, h[2 * b + 1] - h[2 * b]);
        }
        // two 1024-thread blocks share an SM for (almost) the same span
        const double ops_per_sm = 2.0 * 1024.0 * iters * LOP3_UNROLL * LOP3_CHAINS;
        const double per_clk = ops_per_sm / (double)longest;
        const double per_sec = ops_per_sm * num_sms / (ms * 1e-3);
        if (rep > 0 && per_sec > best_per_sec) {  // (rep 0 = warm-up)
            best_per_sec = per_sec;
            best_per_clk = per_clk;
            mhz = (double)longest / (ms * 1e-3) * 1e-6;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(spans);
    if (e == cudaSuccess) {
        e = cudaGetLastError();
    }
    *lane_ops_per_clk_per_sm = best_per_clk;
    *lane_ops_per_sec = best_per_sec;
    *sm_mhz = mhz;
    return e;
}

}  // namespace gstim

// kernels.cu — sm_100a device code of the Pauli-frame sampler.
//
//   (the interpreter kernel lives in interp.cu)
//   gstim_transpose_kernel  bit-major rows -> dense shot-major b8 bytes.
//                         Replaces simd_bit_table::transposed + write_table_data
//                         (/root/reference/src/stim/io/measure_record_writer.h:101-166).
//   gstim_popcount_kernel per-row flip counts (feeds the optional multi-GPU allreduce).
#include "kernels.cuh"

#include <algorithm>

namespace gstim {

// ------------------------------------------------------------------------------------------------
// Output transposer. Tile = 512 shots x 1024 bits. Thread (sw, rg): sw = shot word 0..15 of the tile,
// rg = group of 32 output bits. Each thread gathers 32 source rows' words (coalesced across sw),
// transposes the 32x32 bit block in registers, parks it in shared memory, then warps stream whole
// 128-byte shot segments to the dense (arbitrarily aligned) output rows.
// ------------------------------------------------------------------------------------------------
constexpr int TP_SHOTS = 512;
constexpr int TP_BITS = 1024;
constexpr int TP_PITCH = 33;  // words per shot row in shared memory (+1 to spread banks)

__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
    // After this, a[j] bit i == (input a[i]) bit j.
#pragma unroll
    for (int st = 0; st < 5; st++) {
        const int j = 16 >> st;
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
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

__global__ void __launch_bounds__(512) gstim_transpose_kernel(const TransposeParams p) {
    extern __shared__ __align__(16) uint32_t tile[];
    const uint32_t tid = threadIdx.x;
    const uint32_t sw = tid & 15, rg = tid >> 4;
    const uint64_t shot_word0 = (uint64_t)blockIdx.x * (TP_SHOTS / 32);
    const uint32_t bit0 = blockIdx.y * TP_BITS;

    uint32_t a[32];
    const uint64_t sword = shot_word0 + sw;
    const bool in_range = sword < p.row_words;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const uint32_t bit = bit0 + rg * 32 + i;
        uint32_t v = 0;
        if (bit < p.n_bits && in_range) {
            const uint32_t rm = p.row_map[bit];
            v = p.table[(uint64_t)(rm & 0x7FFFFFFFu) * p.row_words + sword];
            v ^= (uint32_t)0 - (rm >> 31);
        }
        a[i] = v;
    }
    transpose32(a);
    // shot (sw*32+s) is parked at tile row (s*16+sw): bank = (16 s + sw + rg) % 32
#pragma unroll
    for (int s = 0; s < 32; s++) {
        tile[(s * 16 + sw) * TP_PITCH + rg] = a[s];
    }
    __syncthreads();

    const uint32_t nbytes = (p.n_bits + 7) / 8;
    const uint32_t seg0 = blockIdx.y * (TP_BITS / 8);
    const uint32_t seg_len = min((uint32_t)(TP_BITS / 8), nbytes - seg0);
    const uint32_t lane = tid & 31, warp = tid >> 5;
    for (uint32_t sl = warp; sl < TP_SHOTS; sl += 16) {
        const uint64_t shot = (uint64_t)blockIdx.x * TP_SHOTS + sl;
        if (shot >= p.n_shots) {
            break;
        }
        const uint32_t *S = &tile[((sl & 31) * 16 + (sl >> 5)) * TP_PITCH];
        uint8_t *dst = p.out + shot * p.out_pitch + seg0;
        const uint32_t mis = (uint32_t)((uintptr_t)dst & 3);
        uint8_t *base = dst - mis;  // 4-byte aligned
        // aligned destination word w holds segment bytes [4w - mis, 4w - mis + 4)
        for (uint32_t w = lane; w * 4 < mis + seg_len; w += 32) {
            const uint32_t lo = w >= 1 ? S[w - 1] : 0u;
            const uint32_t hi = w < 32 ? S[w] : 0u;
            const uint32_t val = mis == 0 ? hi : __funnelshift_r(lo, hi, 8 * (4 - mis));
            const int first = (int)(4 * w) - (int)mis;  // segment byte index of this word's byte 0
            if (first >= 0 && first + 4 <= (int)seg_len) {
                *(uint32_t *)(base + 4 * w) = val;
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int sb = first + b;
                    if (sb >= 0 && sb < (int)seg_len) {
                        base[4 * w + b] = (uint8_t)(val >> (8 * b));
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
    dim3 grid((unsigned)((p.n_shots + TP_SHOTS - 1) / TP_SHOTS), (p.n_bits + TP_BITS - 1) / TP_BITS);
    static bool attr_set = false;
    const size_t smem = (size_t)TP_SHOTS * TP_PITCH * 4;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gstim_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            return e;
        }
        attr_set = true;
    }
    gstim_transpose_kernel<<<grid, 512, smem, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row popcounts: one warp per (row, slice of shots).
// ------------------------------------------------------------------------------------------------
__global__ void gstim_popcount_kernel(
    const uint32_t *table, uint64_t row_words, uint32_t n_rows, uint64_t n_shots, unsigned long long *counts) {
    const uint32_t row = blockIdx.x;
    if (row >= n_rows) {
        return;
    }
    const uint64_t full_words = n_shots / 32;
    const uint32_t tail_bits = (uint32_t)(n_shots & 31);
    const uint32_t *r = table + (uint64_t)row * row_words;
    unsigned long long acc = 0;
    for (uint64_t w = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; w < full_words; w += (uint64_t)gridDim.y * blockDim.x) {
        acc += __popc(r[w]);
    }
    if (tail_bits && blockIdx.y == 0 && threadIdx.x == 0) {
        acc += __popc(r[full_words] & ((1u << tail_bits) - 1));
    }
    for (int o = 16; o; o >>= 1) {
        acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    }
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicAdd(&counts[row], acc);
    }
}

cudaError_t launch_row_popcount(
    const uint32_t *table, uint64_t row_words, uint32_t n_rows, uint64_t n_shots, unsigned long long *counts, cudaStream_t stream) {
    if (n_rows == 0 || n_shots == 0) {
        return cudaSuccess;
    }
    uint32_t slices = (uint32_t)std::min<uint64_t>(64, (n_shots / 32 + 255) / 256);
    if (slices == 0) {
        slices = 1;
    }
    dim3 grid(n_rows, slices);
    gstim_popcount_kernel<<<grid, 256, 0, stream>>>(table, row_words, n_rows, n_shots, counts);
    return cudaGetLastError();
}

}  // namespace gstim

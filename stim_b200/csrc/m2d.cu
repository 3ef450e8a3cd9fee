// m2d.cu — measurements -> detection events.
//
// Replaces measurements_to_detection_events_helper<W> (/root/reference/src/stim/simulators/measurements_to_detection_events.inl:30-131),
// the batch loop of stream_measurements_to_detection_events (:147-330) behind `stim m2d`, and
// stim.CompiledMeasurementsToDetectionEventsConverter.convert (src/stim/simulators/measurements_to_detection_events.pybind.cc:78-137).
//
// The reference walks the noiseless circuit once per batch of shots: a detector row is the XOR of the recorded
// measurement rows it names, inverted when the reference sample says the noiseless parity is 1, XORed with the
// detector-flip row of a FrameSimulator that only sees the sweep bits (`CX sweep[k] q` etc. flip the frame, which
// flips later measurements, which flips detectors; frame randomisation off). Everything in that walk is GF(2)-linear,
// so the converter is a fixed sparse matrix, built once per circuit:
//     output j  =  const_j  ^  XOR_{m in recs(j)} measurement[m]  ^  XOR_{k in sweeps(j)} sweep[k]
// recs(j) come from the DETECTOR / OBSERVABLE_INCLUDE targets (lowering.cc, with_sweep), const_j is the parity of the
// reference sample over recs(j), and sweeps(j) is the transpose of the per-sweep-bit responses that the backward
// sensitivity pass of response.cc computes (the same pass that builds the event engine's table; it covers classical
// feedback between the sweep-controlled Pauli and the detector).
//
// Device side: a block stages the packed measurement (+ sweep) rows of a tile of shots in shared memory with coalesced
// loads, then every thread computes output bytes: 8 outputs x their sources, each a bit read from the staged rows, and
// stores them (consecutive threads -> consecutive bytes of a row). Input and output cross PCIe once; the kernel is far
// from being the bound (c3: 2029 B in + 1951 B out per shot).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/gstim.h"
#include "circuit.h"
#include "hostpipe.h"
#include "lowering.h"
#include "response.h"
#include "tableau_ref.h"

namespace gstim {

struct M2dParams {
    const uint32_t *src_off;   // n_out + 1
    const uint32_t *src;       // source bit index in the staged row: measurement m, or meas_bytes * 8 + sweep bit k
    const uint8_t *const_bits; // packed constants per output (layout order)
    const uint8_t *meas;       // device, packed rows
    uint64_t meas_pitch;
    const uint8_t *sweep;      // device, packed rows or null
    uint64_t sweep_pitch;
    uint32_t meas_bytes, sweep_bytes;
    uint32_t n_out;            // output bits of the main rows
    uint32_t out_bytes;
    uint8_t *out;
    uint64_t out_pitch;
    uint32_t obs0, n_obs;      // separate observables: outputs [obs0, obs0 + n_obs) of the source tables -> obs_out
    uint32_t obs_bytes;
    uint8_t *obs_out;
    uint64_t obs_pitch;
    uint64_t n_shots;
    uint32_t tile_shots;
};

__global__ void __launch_bounds__(512) gstim_m2d_kernel(const M2dParams p) {
    extern __shared__ uint4 smem4[];
    uint8_t *const rows = reinterpret_cast<uint8_t *>(smem4);
    const uint32_t in_bytes = p.meas_bytes + p.sweep_bytes;  // staged row: measurement bytes, then sweep bytes
    for (uint64_t t0 = (uint64_t)blockIdx.x * p.tile_shots; t0 < p.n_shots; t0 += (uint64_t)gridDim.x * p.tile_shots) {
        const uint32_t n = (uint32_t)min((uint64_t)p.tile_shots, p.n_shots - t0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n * in_bytes; i += blockDim.x) {
            const uint32_t s = i / in_bytes, b = i - s * in_bytes;
            rows[i] = b < p.meas_bytes ? p.meas[(t0 + s) * p.meas_pitch + b] : (p.sweep ? p.sweep[(t0 + s) * p.sweep_pitch + (b - p.meas_bytes)] : (uint8_t)0);
        }
        __syncthreads();
        auto emit = [&](uint32_t first_out, uint32_t n_bits, uint32_t n_bytes, uint8_t *dst, uint64_t pitch) {
            for (uint32_t i = threadIdx.x; i < n * n_bytes; i += blockDim.x) {
                const uint32_t s = i / n_bytes, b = i - s * n_bytes;
                const uint8_t *row = rows + (size_t)s * in_bytes;
                uint32_t v = 0;
                for (uint32_t k = 0; k < 8 && b * 8 + k < n_bits; k++) {
                    const uint32_t j = first_out + b * 8 + k;
                    uint32_t bit = (p.const_bits[j >> 3] >> (j & 7)) & 1u;
                    for (uint32_t e = p.src_off[j]; e < p.src_off[j + 1]; e++) {
                        const uint32_t idx = p.src[e];
                        bit ^= (row[idx >> 3] >> (idx & 7)) & 1u;
                    }
                    v |= bit << k;
                }
                dst[(t0 + s) * pitch + b] = (uint8_t)v;
            }
        };
        if (p.out != nullptr && p.out_bytes) {
            emit(0, p.n_out, p.out_bytes, p.out, p.out_pitch);
        }
        if (p.obs_out != nullptr && p.obs_bytes) {
            emit(p.obs0, p.n_obs, p.obs_bytes, p.obs_out, p.obs_pitch);
        }
    }
}

// 32 x 32 bit transpose across a warp: lane l gives row l, gets column l (bit s of the result = bit l of lane s's word).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
    for (uint32_t j = 16, m = 0x0000FFFFu; j != 0; j >>= 1, m ^= m << j) {
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
        x = (lane & j) ? (((y >> j) & m) | (x & ~m)) : ((x & m) | ((y & m) << j));
    }
    return x;
}

// Bit-sliced converter: a block takes 32 shots at a time. Their packed rows are staged in shared memory (coalesced), turned
// into one 32-bit word per input bit (bit s = shot s; warp transposes of 32 x 32 bit blocks), every output is the XOR of its
// source WORDS (one shared-memory load per source and 32 shots, where the row-wise kernel above does one per source and shot),
// and the output words are transposed back into packed rows. Used when the three arrays fit in shared memory.
//   smem: stage[32][pitch_w] words (pitch_w odd), inT[in_words * 32], outT[out_words * 32]
__global__ void __launch_bounds__(512) gstim_m2d_sliced_kernel(const M2dParams p, uint32_t pitch_w, uint32_t in_words, uint32_t out_words) {
    extern __shared__ uint4 smem4[];
    uint32_t *const stage = reinterpret_cast<uint32_t *>(smem4);
    uint32_t *const inT = stage + 32u * pitch_w;
    uint32_t *const outT = inT + in_words * 32u;
    uint8_t *const stage8 = reinterpret_cast<uint8_t *>(stage);
    const uint32_t in_bytes = p.meas_bytes + p.sweep_bytes, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t pitch_b = pitch_w * 4u;
    for (uint64_t t0 = (uint64_t)blockIdx.x * 32u; t0 < p.n_shots; t0 += (uint64_t)gridDim.x * 32u) {
        const uint32_t n = (uint32_t)min((uint64_t)32u, p.n_shots - t0);
        __syncthreads();
        // stage the rows (zero padding behind each row and for missing shots)
        for (uint32_t i = threadIdx.x; i < 32u * pitch_w; i += blockDim.x) {
            stage[i] = 0;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n * in_bytes; i += blockDim.x) {
            const uint32_t s = i / in_bytes, b = i - s * in_bytes;
            stage8[s * pitch_b + b] =
                b < p.meas_bytes ? p.meas[(t0 + s) * p.meas_pitch + b] : (p.sweep ? p.sweep[(t0 + s) * p.sweep_pitch + (b - p.meas_bytes)] : (uint8_t)0);
        }
        __syncthreads();
        for (uint32_t w = warp; w < in_words; w += n_warps) {
            inT[w * 32u + lane] = warp_transpose32(stage[lane * pitch_w + w], lane);
        }
        __syncthreads();
        auto emit = [&](uint32_t first_out, uint32_t n_bits, uint32_t n_bytes, uint8_t *dst, uint64_t pitch) {
            const uint32_t words = (n_bits + 31u) / 32u;
            for (uint32_t j = threadIdx.x; j < words * 32u; j += blockDim.x) {
                uint32_t v = 0;
                if (j < n_bits) {
                    const uint32_t o = first_out + j;
                    v = ((p.const_bits[o >> 3] >> (o & 7)) & 1u) ? 0xFFFFFFFFu : 0u;
                    for (uint32_t e = p.src_off[o]; e < p.src_off[o + 1]; e++) {
                        v ^= inT[p.src[e]];
                    }
                }
                outT[j] = v;
            }
            __syncthreads();
            for (uint32_t w = warp; w < words; w += n_warps) {
                stage[lane * pitch_w + w] = warp_transpose32(outT[w * 32u + lane], lane);
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < n * n_bytes; i += blockDim.x) {
                const uint32_t s = i / n_bytes, b = i - s * n_bytes;
                dst[(t0 + s) * pitch + b] = stage8[s * pitch_b + b];
            }
            __syncthreads();
        };
        if (p.out != nullptr && p.out_bytes) {
            emit(0, p.n_out, p.out_bytes, p.out, p.out_pitch);
        }
        if (p.obs_out != nullptr && p.obs_bytes) {
            emit(p.obs0, p.n_obs, p.obs_bytes, p.obs_out, p.obs_pitch);
        }
    }
}

}  // namespace gstim

using namespace gstim;

void gstim_set_last_error(const char *msg);

struct gstim_m2d {
    int device = 0;
    int num_sms = 0;
    size_t smem_optin = 0;
    uint64_t M = 0, D = 0, L = 0, n_sweep = 0;
    // per output id (detector d, observable D + l): sources and constant
    std::vector<std::vector<uint32_t>> recs, sweeps;
    std::vector<uint8_t> konst;
    // device copy of the tables for the layout last used
    uint32_t layout = 0xFFFFFFFFu;
    void *d_off = nullptr, *d_src = nullptr, *d_const = nullptr, *d_in = nullptr, *d_sweep = nullptr, *d_out = nullptr, *d_obs = nullptr;
    size_t cap_in = 0, cap_sweep = 0, cap_out = 0, cap_obs = 0;
    cudaStream_t stream = nullptr;
    gstim::HostStager stage_in, stage_out;  // page-locked staging pairs (hostpipe.h)
    ~gstim_m2d() {
        for (void *p : {d_off, d_src, d_const, d_in, d_sweep, d_out, d_obs}) {
            if (p) {
                cudaFree(p);
            }
        }
        if (stream) {
            cudaStreamDestroy(stream);
        }
    }
};

namespace {

void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
        }
        throw std::runtime_error(std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what);
    }
}

template <typename F>
int m2d_guarded(F &&f) {
    try {
        f();
        return GSTIM_OK;
    } catch (const std::invalid_argument &e) {
        gstim_set_last_error(e.what());
        return GSTIM_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range &e) {
        gstim_set_last_error(e.what());
        return GSTIM_ERR_OUT_OF_RANGE;
    } catch (const std::exception &e) {
        gstim_set_last_error(e.what());
        return std::string(e.what()).rfind("CUDA", 0) == 0 ? GSTIM_ERR_CUDA : GSTIM_ERR_INTERNAL;
    }
}

void ensure(void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) {
        return;
    }
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
        *cap = 0;
    }
    ck(cudaMalloc(p, bytes), "cudaMalloc");
    *cap = bytes;
}

// Uploads the source tables in the order of the requested output layout: main rows = detectors (+ observables when
// appended), then the observables again for a separate observable output.
void upload_layout(gstim_m2d *h, bool append_obs) {
    const uint32_t want = append_obs ? 1u : 0u;
    if (h->layout == want) {
        return;
    }
    std::vector<uint32_t> off{0}, src;
    std::vector<uint8_t> konst((h->D + 2 * h->L + 7) / 8 + 1, 0);
    const uint32_t sweep_base = (uint32_t)((h->M + 7) / 8) * 8;
    auto put = [&](uint64_t id) {
        const size_t j = off.size() - 1;
        // (a measurement named twice by one detector cancels)
        std::vector<uint32_t> r = h->recs[id];
        std::sort(r.begin(), r.end());
        for (size_t i = 0; i < r.size(); i++) {
            if (i + 1 < r.size() && r[i] == r[i + 1]) {
                i++;
                continue;
            }
            src.push_back(r[i]);
        }
        for (uint32_t k : h->sweeps[id]) {
            src.push_back(sweep_base + k);
        }
        off.push_back((uint32_t)src.size());
        if (h->konst[id]) {
            konst[j >> 3] |= (uint8_t)(1u << (j & 7));
        }
    };
    for (uint64_t d = 0; d < h->D; d++) {
        put(d);
    }
    for (uint64_t l = 0; l < h->L; l++) {  // outputs [D, D + L): appended or separate observables
        put(h->D + l);
    }
    for (void **p : {&h->d_off, &h->d_src, &h->d_const}) {
        if (*p) {
            cudaFree(*p);
            *p = nullptr;
        }
    }
    ck(cudaMalloc(&h->d_off, off.size() * 4), "cudaMalloc");
    ck(cudaMalloc(&h->d_src, std::max<size_t>(src.size(), 1) * 4), "cudaMalloc");
    ck(cudaMalloc(&h->d_const, konst.size()), "cudaMalloc");
    ck(cudaMemcpy(h->d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice), "upload");
    if (!src.empty()) {
        ck(cudaMemcpy(h->d_src, src.data(), src.size() * 4, cudaMemcpyHostToDevice), "upload");
    }
    ck(cudaMemcpy(h->d_const, konst.data(), konst.size(), cudaMemcpyHostToDevice), "upload");
    h->layout = want;
}

}  // namespace

// A converter from explicit source lists: output j = XOR of the input bits recs[j] (the DEM sampler's error replay, dem.cu).
gstim_m2d *gstim_m2d_from_lists(int device, uint64_t n_inputs, uint64_t D, uint64_t L, std::vector<std::vector<uint32_t>> recs) {
    auto h = std::make_unique<gstim_m2d>();
    h->M = n_inputs;
    h->D = D;
    h->L = L;
    h->n_sweep = 0;
    h->recs = std::move(recs);
    h->recs.resize(D + L);
    h->sweeps.resize(D + L);
    h->konst.assign(D + L, 0);
    h->device = device;
    ck(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    ck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    ck(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    ck(cudaFuncSetAttribute(gstim_m2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin), "smem attribute");
    ck(cudaFuncSetAttribute(gstim_m2d_sliced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin), "smem attribute");
    return h.release();
}

extern "C" {

int gstim_m2d_create_from_text(const char *circuit_text, size_t text_len, int skip_reference_sample, int device, gstim_m2d **out) {
    return m2d_guarded([&] {
        if (circuit_text == nullptr || out == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        *out = nullptr;
        auto h = std::make_unique<gstim_m2d>();
        Circuit c = Circuit::from_text(std::string_view(circuit_text, text_len));
        LoweredCircuit lc = lower_circuit(c, 0, (1u << 20), true);
        ResponseTable rt = build_response_table(lc, true);
        h->M = lc.stats.num_measurements;
        h->D = lc.stats.num_detectors;
        h->L = lc.stats.num_observables;
        h->n_sweep = lc.stats.num_sweep_bits;
        const uint64_t n = h->D + h->L;
        h->recs.resize(n);
        h->sweeps.resize(n);
        h->konst.assign(n, 0);
        std::vector<uint8_t> ref;
        if (!skip_reference_sample) {
            ref = reference_sample(c);
        }
        for (uint64_t j = 0; j < n; j++) {
            for (uint64_t m : lc.out_recs[j]) {
                h->recs[j].push_back((uint32_t)m);
                if (!ref.empty() && ref[m]) {
                    h->konst[j] ^= 1;
                }
            }
        }
        // (a table that was not built — chains longer than 16 — has no sweep responses; such circuits have to do without sweep bits)
        if (!rt.eligible && h->n_sweep > 0) {
            throw std::invalid_argument("measurement conversion with sweep bits is not supported for this circuit: " + rt.why_not);
        }
        for (size_t k = 0; k < rt.sweep_responses.size(); k++) {
            for (uint32_t id : rt.sweep_responses[k]) {
                h->sweeps[id].push_back((uint32_t)k);
            }
        }
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            throw std::runtime_error("CUDA: no usable device: this library has no CPU fallback.");
        }
        if (device < 0 || device >= ndev) {
            throw std::invalid_argument("CUDA device ordinal out of range.");
        }
        h->device = device;
        ck(cudaSetDevice(device), "cudaSetDevice");
        cudaDeviceProp prop;
        ck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
        h->num_sms = prop.multiProcessorCount;
        h->smem_optin = prop.sharedMemPerBlockOptin;
        ck(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ck(cudaFuncSetAttribute(gstim_m2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin), "smem attribute");
    ck(cudaFuncSetAttribute(gstim_m2d_sliced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin), "smem attribute");
        *out = h.release();
    });
}

void gstim_m2d_destroy(gstim_m2d *h) {
    if (h) {
        cudaSetDevice(h->device);
        delete h;
    }
}

int gstim_m2d_get_sizes(const gstim_m2d *h, uint64_t *num_measurements, uint64_t *num_detectors, uint64_t *num_observables,
                        uint64_t *num_sweep_bits) {
    return m2d_guarded([&] {
        if (h == nullptr) {
            throw std::invalid_argument("NULL converter.");
        }
        if (num_measurements) {
            *num_measurements = h->M;
        }
        if (num_detectors) {
            *num_detectors = h->D;
        }
        if (num_observables) {
            *num_observables = h->L;
        }
        if (num_sweep_bits) {
            *num_sweep_bits = h->n_sweep;
        }
    });
}

int gstim_m2d_convert(gstim_m2d *h, uint64_t shots, uint32_t flags, const void *measurements, int64_t meas_stride, const void *sweep_bits,
                      int64_t sweep_stride, void *dets_out, int64_t dets_stride, void *obs_out, int64_t obs_stride) {
    return m2d_guarded([&] {
        if (h == nullptr) {
            throw std::invalid_argument("NULL converter.");
        }
        const bool append = (flags & GSTIM_APPEND_OBS) != 0, separate = (flags & GSTIM_SEPARATE_OBS) != 0;
        if (flags & GSTIM_PREPEND_OBS) {
            throw std::invalid_argument("prepend_observables is not an option of the measurement converter.");
        }
        if (meas_stride < 0 || sweep_stride < 0 || dets_stride < 0 || obs_stride < 0) {
            throw std::invalid_argument("negative strides are not supported.");
        }
        if (shots == 0) {
            return;
        }
        if (measurements == nullptr && h->M > 0) {
            throw std::invalid_argument("measurements must not be NULL.");
        }
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        upload_layout(h, append);
        const uint32_t meas_bytes = (uint32_t)((h->M + 7) / 8), sweep_bytes = (uint32_t)((h->n_sweep + 7) / 8);
        const uint32_t n_out = (uint32_t)(h->D + (append ? h->L : 0)), out_bytes = (n_out + 7) / 8;
        const uint32_t obs_bytes = separate ? (uint32_t)((h->L + 7) / 8) : 0;
        const uint64_t mp = meas_stride ? (uint64_t)meas_stride : meas_bytes, sp = sweep_stride ? (uint64_t)sweep_stride : sweep_bytes;
        const uint64_t dp = dets_stride ? (uint64_t)dets_stride : out_bytes, op = obs_stride ? (uint64_t)obs_stride : obs_bytes;
        const uint32_t in_bytes = meas_bytes + sweep_bytes;
        const uint32_t tile = (uint32_t)std::max<size_t>(1, std::min<size_t>(128, (h->smem_optin - 1024) / std::max<uint32_t>(in_bytes, 1)));
        if ((size_t)in_bytes > h->smem_optin - 1024) {
            throw std::invalid_argument("a packed measurement row does not fit in shared memory");
        }
        // chunks of shots through device staging (dense rows on the device)
        const uint64_t chunk = std::max<uint64_t>(1, (256ull << 20) / std::max<uint32_t>(in_bytes + out_bytes + obs_bytes, 1));
        if (dets_out && !hp_is_pinned(dets_out)) {
            hp_hugepage_hint(dets_out, shots * dp);
        }
        const bool timing = getenv("GSTIM_M2D_TIMING") != nullptr;
        double t_in = 0, t_kernel = 0, t_out = 0;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        for (uint64_t s0 = 0; s0 < shots; s0 += chunk) {
            const uint64_t n = std::min(chunk, shots - s0);
            double t0 = timing ? now() : 0;
            ensure(&h->d_in, &h->cap_in, std::max<uint64_t>(n * meas_bytes, 16));
            ensure(&h->d_out, &h->cap_out, std::max<uint64_t>(n * out_bytes, 16));
            // caller rows -> device: threaded packing into page-locked staging + async DMA (direct DMA from page-locked memory)
            if (meas_bytes) {
                hp_h2d_rows(h->stage_in, h->stream, (const uint8_t *)measurements + s0 * mp, mp, (uint8_t *)h->d_in, meas_bytes, n);
            }
            const bool have_sweep = sweep_bits != nullptr && sweep_bytes > 0;
            if (have_sweep) {
                ensure(&h->d_sweep, &h->cap_sweep, n * sweep_bytes);
                hp_h2d_rows(h->stage_in, h->stream, (const uint8_t *)sweep_bits + s0 * sp, sp, (uint8_t *)h->d_sweep, sweep_bytes, n);
            }
            if (obs_bytes) {
                ensure(&h->d_obs, &h->cap_obs, n * obs_bytes);
            }
            M2dParams p{};
            p.src_off = (const uint32_t *)h->d_off;
            p.src = (const uint32_t *)h->d_src;
            p.const_bits = (const uint8_t *)h->d_const;
            p.meas = (const uint8_t *)h->d_in;
            p.meas_pitch = meas_bytes;
            p.sweep = have_sweep ? (const uint8_t *)h->d_sweep : nullptr;
            p.sweep_pitch = sweep_bytes;
            p.meas_bytes = meas_bytes;
            p.sweep_bytes = sweep_bytes;
            p.n_out = n_out;
            p.out_bytes = out_bytes;
            p.out = dets_out ? (uint8_t *)h->d_out : nullptr;
            p.out_pitch = out_bytes;
            p.obs0 = (uint32_t)h->D;
            p.n_obs = (uint32_t)h->L;
            p.obs_bytes = obs_out ? obs_bytes : 0;
            p.obs_out = obs_out && obs_bytes ? (uint8_t *)h->d_obs : nullptr;
            p.obs_pitch = obs_bytes;
            p.n_shots = n;
            p.tile_shots = tile;
            if (timing) {
                cudaStreamSynchronize(h->stream);
                t_in += now() - t0;
                t0 = now();
            }
            // bit-sliced kernel when its three shared-memory arrays fit, else the row-wise one
            const uint32_t widest = std::max(std::max(in_bytes, out_bytes), obs_bytes);
            const uint32_t pitch_w = ((widest + 3) / 4 + 1) | 1u, in_words = (in_bytes * 8 + 31) / 32;
            const uint32_t out_words = (std::max<uint32_t>(n_out, separate ? (uint32_t)h->L : 0) + 31) / 32;
            const size_t sliced_smem = ((size_t)32 * pitch_w + (size_t)in_words * 32 + (size_t)out_words * 32) * 4;
            if (sliced_smem <= h->smem_optin && !getenv("GSTIM_M2D_ROWWISE")) {
                const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 31) / 32, (uint64_t)h->num_sms);
                gstim_m2d_sliced_kernel<<<grid, 512, sliced_smem, h->stream>>>(p, pitch_w, in_words, out_words);
            } else {
                const uint32_t grid = (uint32_t)std::min<uint64_t>((n + tile - 1) / tile, (uint64_t)h->num_sms * 2);
                gstim_m2d_kernel<<<grid, 512, (size_t)tile * in_bytes + 16, h->stream>>>(p);
            }
            ck(cudaGetLastError(), "gstim_m2d_kernel launch");
            if (timing) {
                cudaStreamSynchronize(h->stream);
                t_kernel += now() - t0;
                t0 = now();
            }
            // device rows -> caller: direct DMA into page-locked memory, else staged sub-chunks copied out by host threads
            auto rows_out = [&](const void *d, uint64_t row_bytes, uint8_t *dst, uint64_t pitch) {
                if (hp_is_pinned(dst)) {
                    ck(cudaMemcpy2DAsync(dst, pitch, d, row_bytes, row_bytes, n, cudaMemcpyDeviceToHost, h->stream), "D2H");
                } else {
                    hp_staged_d2h(h->stage_out, h->stream, (const uint8_t *)d, row_bytes, n,
                                  [&](const uint8_t *row, uint64_t i) { memcpy(dst + i * pitch, row, row_bytes); });
                }
            };
            if (dets_out && out_bytes) {
                rows_out(h->d_out, out_bytes, (uint8_t *)dets_out + s0 * dp, dp);
            }
            if (obs_out && obs_bytes) {
                rows_out(h->d_obs, obs_bytes, (uint8_t *)obs_out + s0 * op, op);
            }
            ck(cudaStreamSynchronize(h->stream), "cudaStreamSynchronize");
            if (timing) {
                t_out += now() - t0;
            }
        }
        if (timing) {
            fprintf(stderr, "m2d: %llu shots, tile %u, in %.1f ms, kernel %.1f ms, out %.1f ms\n", (unsigned long long)shots, tile, t_in * 1e3,
                    t_kernel * 1e3, t_out * 1e3);
        }
    });
}

}  // extern "C"

/* gstim.h — C ABI of the B200-native Pauli-frame sampler (libgstim.so).
 *
 * This is the drop-in boundary for the ONE hot path of quantumlib/Stim that this library replaces:
 * the FrameSimulator bulk sampler behind `stim detect`, `stim sample`,
 * `stim.Circuit.compile_detector_sampler()` and `stim.Circuit.compile_sampler()`.
 * The reference has no FFI for this path (FrameSimulator<W> is a header template instantiated in
 * its callers); each entry point below names the reference call it replaces. All paths are
 * relative to /root/reference/.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints. No C++/torch types.
 *   - every function returns 0 on success or a GSTIM_ERR_* code; the message is available from
 *     gstim_last_error(). The C++/Python shim re-raises INVALID_ARGUMENT as
 *     std::invalid_argument/ValueError and OUT_OF_RANGE as std::out_of_range/IndexError, matching
 *     the exception types the reference throws (src/stim/main_namespaced.cc:113-122 and pybind).
 *   - a sampler handle is NOT thread-safe (like the reference object, which owns a mutable RNG);
 *     distinct handles are independent. Calls are synchronous: on return the results are in the
 *     caller's buffer.
 *   - the caller owns every input and output buffer; the library owns device memory and streams.
 *   - there is NO CPU fallback: creating a sampler without a usable CUDA device fails with
 *     GSTIM_ERR_CUDA.
 *   - successive sample calls on one handle continue the random stream (they never repeat shots),
 *     as with the reference samplers (src/stim/py/compiled_detector_sampler.pybind.cc:180-196).
 */
#ifndef GSTIM_H
#define GSTIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSTIM_OK 0
#define GSTIM_ERR_INVALID_ARGUMENT 1 /* std::invalid_argument in the reference */
#define GSTIM_ERR_OUT_OF_RANGE 2     /* std::out_of_range in the reference */
#define GSTIM_ERR_CUDA 3
#define GSTIM_ERR_OOM 4
#define GSTIM_ERR_IO 5
#define GSTIM_ERR_INTERNAL 6

/* which record a sampler produces */
#define GSTIM_MODE_DETECTORS 0    /* detection events + observables (stim detect / compile_detector_sampler) */
#define GSTIM_MODE_MEASUREMENTS 1 /* measurement record (stim sample / compile_sampler) */

/* output flags */
#define GSTIM_BIT_PACKED 0x01u   /* uint8[shots, ceil(n/8)] little-endian bits (b8); else one byte per bit */
#define GSTIM_PREPEND_OBS 0x02u  /* observables before detectors in each shot (--prepend_observables) */
#define GSTIM_APPEND_OBS 0x04u   /* observables after detectors in each shot (--append_observables) */
#define GSTIM_SEPARATE_OBS 0x08u /* observables into obs_out (separate_observables=True / --obs_out) */

typedef struct gstim_sampler gstim_sampler;

/* mirrors stim::CircuitStats (src/stim/circuit/circuit_instruction.h:30-50) + the chosen device plan */
typedef struct gstim_stats {
    uint64_t num_qubits;       /* highest qubit index + 1 */
    uint64_t num_measurements;
    uint64_t num_detectors;
    uint64_t num_observables;
    uint64_t max_lookback;
    uint64_t active_qubits;    /* qubits kept in the frame after compaction */
    uint64_t program_words;    /* lowered instruction stream length (uint32 words) */
    uint64_t num_batches;      /* concurrent batches in the stream */
    uint64_t num_barriers;     /* batches that need a block-wide barrier */
    uint64_t num_noise_sites;  /* Philox site counter at the end of the circuit */
    uint64_t num_collapse_sites;
    uint32_t threads;          /* threads per block */
    uint32_t lanes_per_item;   /* G */
    uint32_t slots;            /* concurrent items per batch pass */
    uint32_t max_columns;      /* largest K (128-shot columns per block) that fits in shared memory */
    uint32_t chunk_words;
    uint32_t smem_bytes_max;   /* dynamic shared memory at max_columns */
} gstim_stats;

int gstim_version(void);

/* Message of the most recent failure on this thread (never NULL). */
const char *gstim_last_error(void);

/* Number of usable CUDA devices (0 when none / no driver). */
int gstim_device_count(void);

/* Parse + validate a circuit on the host only (no GPU needed): fills the CircuitStats part of *out.
 * Replaces: Circuit::compute_stats()  src/stim/circuit/circuit.cc:719-725 */
int gstim_circuit_stats(const char *circuit_text, size_t text_len, gstim_stats *out);

/* Host-only (no GPU): the noiseless reference sample of the circuit, little-endian packed into bits_out
 * (n_bits = number of measurements). Random measurement outcomes are fixed to 0.
 * Replaces: TableauSimulator::reference_sample_circuit  src/stim/simulators/tableau_simulator.inl:1435-1438 */
int gstim_reference_sample(const char *circuit_text, size_t text_len, uint8_t *bits_out, size_t n_bits);

/* Host-only lowering (no GPU needed): circuit text -> the uint32 instruction stream the interpreter
 * kernel executes (format: stim_b200/csrc/program.h), with barrier flags computed for `slots`
 * concurrent thread groups and cut into chunks of `chunk_words` words (0 = library default).
 * plan_out receives the 14 uint32 fields of GstimPlan (program.h). Call with words == NULL to get
 * the required length in *n_words. No reference analogue (new subsystem: the lowering). */
int gstim_lower_text(
    const char *circuit_text,
    size_t text_len,
    int mode,
    uint32_t slots,
    uint32_t chunk_words,
    uint32_t *words,
    size_t *n_words,
    uint32_t plan_out[16]);

/* Parse + lower + upload a circuit given in Stim's circuit file format.
 * Replaces: CompiledDetectorSampler(circuit, rng)  src/stim/py/compiled_detector_sampler.pybind.cc:28-32
 *           CompiledMeasurementSampler(ref, circuit, skip_ref, rng)  src/stim/py/compiled_measurement_sampler.pybind.cc:26-29
 *           Circuit::from_file + FrameSimulator construction in src/stim/cmd/command_detect.cc:65-77,
 *           src/stim/cmd/command_sample.cc:58-69
 * mode: GSTIM_MODE_*.  device: CUDA ordinal.  seed: Philox key. */
int gstim_create_from_text(
    const char *circuit_text, size_t text_len, int mode, uint64_t seed, int device, gstim_sampler **out);

/* Same, on several CUDA devices of this process ("devices[], n_devices"). The host-output calls
 * (gstim_sample_detectors, gstim_sample_measurements) cut their shots into one contiguous range per device, sample the
 * ranges concurrently and fill the caller's arrays; there is no inter-GPU traffic. With the event engine the result equals
 * a single-device call bit for bit. Every other call acts on devices[0]. */
int gstim_create_from_text_multi(
    const char *circuit_text, size_t text_len, int mode, uint64_t seed, const int *devices, int n_devices, gstim_sampler **out);

void gstim_destroy(gstim_sampler *s);

int gstim_get_stats(const gstim_sampler *s, gstim_stats *out);

/* Copy of the lowered instruction stream (for inspection / the program-level emulator in oracle/).
 * Call with words == NULL to get the length in *n_words. */
int gstim_get_program(const gstim_sampler *s, uint32_t *words, size_t *n_words);

/* Binds the noiseless reference sample that is XORed into measurement results
 * (MODE_MEASUREMENTS only). bits: num_measurements bits, little-endian packed. NULL = all zero
 * (== skip_reference_sample=True).
 * Replaces: the `reference_sample` argument of sample_batch_measurements
 *           src/stim/simulators/frame_simulator_util.h:103-109, :123-130 */
int gstim_set_reference_sample(gstim_sampler *s, const uint8_t *bits, size_t n_bits);

/* Shot offset (in shots) the next call will start at; lets a caller reproduce / resume a range. */
int gstim_get_shot_offset(const gstim_sampler *s, uint64_t *offset);
int gstim_set_shot_offset(gstim_sampler *s, uint64_t offset);

/* Detection-event sampling into HOST memory, shot-major.
 * Replaces: CompiledDetectorSampler::sample_to_numpy  src/stim/py/compiled_detector_sampler.pybind.cc:34-86
 *           sample_batch_detection_events<W>         src/stim/simulators/frame_simulator_util.h:48-50
 * dets_out: [shots][n] where n = D (+L with PREPEND/APPEND); with GSTIM_BIT_PACKED each shot is
 *   ceil(n/8) bytes, else n bytes of 0/1. dets_shot_stride = bytes between consecutive shots
 *   (0 = dense). obs_out/obs_shot_stride likewise for the L observables when GSTIM_SEPARATE_OBS. */
int gstim_sample_detectors(
    gstim_sampler *s,
    uint64_t shots,
    uint32_t flags,
    void *dets_out,
    int64_t dets_shot_stride,
    void *obs_out,
    int64_t obs_shot_stride);

/* Measurement sampling into HOST memory, shot-major.
 * Replaces: CompiledMeasurementSampler::sample_to_numpy  src/stim/py/compiled_measurement_sampler.pybind.cc:31-35
 *           sample_batch_measurements<W>                 src/stim/simulators/frame_simulator_util.inl:291-315 */
int gstim_sample_measurements(gstim_sampler *s, uint64_t shots, uint32_t flags, void *out, int64_t shot_stride);

/* Same as the two calls above but the output buffers are DEVICE pointers on the sampler's device
 * and only the bit-packed (b8) layout is produced. Results stay in HBM (no PCIe traffic). */
int gstim_sample_detectors_device(
    gstim_sampler *s,
    uint64_t shots,
    uint32_t flags,
    void *dets_out_dev,
    int64_t dets_shot_stride,
    void *obs_out_dev,
    int64_t obs_shot_stride);
int gstim_sample_measurements_device(gstim_sampler *s, uint64_t shots, void *out_dev, int64_t shot_stride);

/* Streaming to files in Stim's result formats "01", "b8", "r8", "hits", "dets", "ptb64".
 * Replaces: sample_batch_detection_events_writing_results_to_disk  src/stim/simulators/frame_simulator_util.h:67-77
 *           sample_batch_measurements_writing_results_to_disk      src/stim/simulators/frame_simulator_util.h:123-130
 * fd / obs_fd are open file descriptors owned by the caller (obs_fd < 0 = none).
 * Error behaviour follows the reference: ptb64 needs shots % 64 == 0 (invalid_argument,
 * src/stim/io/measure_record_writer.h:123-125); combining prepend/append/obs_out is out_of_range
 * (src/stim/simulators/frame_simulator_util.inl:127-129). */
int gstim_sample_detectors_to_fd(
    gstim_sampler *s, uint64_t shots, uint32_t flags, int fd, const char *format, int obs_fd, const char *obs_format);
int gstim_sample_measurements_to_fd(gstim_sampler *s, uint64_t shots, int fd, const char *format);

/* Host-only (no GPU): encode shot-major bit-packed rows (bit k of a shot at rows[shot*row_pitch + k/8] >> k%8)
 * into any of Stim's result formats. In the "dets" format bits [0, prefix_transition) are printed with prefix1
 * and the rest with prefix2 ('D'/'L'/'M').
 * Replaces: write_table_data + MeasureRecordWriterFormat*  src/stim/io/measure_record_writer.h:111-166,
 *           src/stim/io/measure_record_writer.cc:61-211 */
int gstim_write_shots_to_fd(
    const uint8_t *rows,
    size_t row_pitch,
    uint64_t shots,
    uint64_t n_bits,
    int fd,
    const char *format,
    char prefix1,
    char prefix2,
    uint64_t prefix_transition);

/* Per-detector and per-observable flip counts over `shots` fresh shots: counts[D+L] (uint64),
 * written to HOST memory; counts_dev (optional, may be NULL) receives the same on the device so a
 * multi-GPU caller can allreduce it (NCCL sum over uint64[D+L]) without a host round trip. */
int gstim_detector_flip_counts(gstim_sampler *s, uint64_t shots, uint64_t *counts_host, void *counts_dev);

/* Flip counts of every output bit over `shots` fresh shots, computed on the device from the bit-major table (no
 * transposition, no result transfer): single[n] and pair[n - 1] (count of shots where bits j and j + 1 are both set),
 * n = D + L in detector mode (detectors then observables) or M in measurement mode (reference sample applied).
 * Any of the four outputs may be NULL; the *_dev pointers are device memory (uint64), e.g. for an NCCL allreduce.
 * This is the statistic of the reference's own sampling tests (src/stim/cmd/command_sample.test.cc:39-71), and what
 * sinter's collection loop reduces shots to (glue/sample/src/sinter/_decoding/_stim_then_decode_sampler.py:162-185). */
int gstim_bit_counts(gstim_sampler *s, uint64_t shots, uint64_t *single_host, uint64_t *pair_host, void *single_dev, void *pair_dev);

/* ---- detector error model sampling --------------------------------------------------------
 * Replaces: DemSampler<W>::resample / sample_write          src/stim/simulators/dem_sampler.inl:52-130
 *           stim sample_dem                                  src/stim/cmd/command_sample_dem.cc:25-93
 *           stim.CompiledDemSampler.sample / sample_write    src/stim/simulators/dem_sampler.pybind.cc
 * dem_text is a detector error model in Stim's .dem format (error / detector / logical_observable / shift_detectors /
 * repeat). Every `error(p)` is an independent Bernoulli(p) row XORed into its targets' rows; parse errors are
 * GSTIM_ERR_INVALID_ARGUMENT. Successive calls continue the random stream. */
typedef struct gstim_dem_sampler gstim_dem_sampler;
int gstim_dem_counts(const char *dem_text, size_t text_len, uint64_t *num_detectors, uint64_t *num_observables, uint64_t *num_errors);
int gstim_dem_create_from_text(const char *dem_text, size_t text_len, uint64_t seed, int device, gstim_dem_sampler **out);
void gstim_dem_destroy(gstim_dem_sampler *s);
int gstim_dem_set_shot_offset(gstim_dem_sampler *s, uint64_t offset);
/* Host outputs, shot-major: dets [shots, D], obs [shots, L], errs [shots, E] (any may be NULL), one byte per bit or
 * bit-packed little-endian with GSTIM_BIT_PACKED; strides in bytes (0 = dense). errs = which error mechanisms fired
 * (the reference's return_errors / --err_out). Replaying recorded errors (--replay_err_in) is not supported. */
int gstim_dem_sample(gstim_dem_sampler *s, uint64_t shots, uint32_t flags, void *dets_out, int64_t dets_stride, void *obs_out,
                     int64_t obs_stride, void *errs_out, int64_t errs_stride);
/* A detector error model is a response table (one site per error mechanism, classes of equal probability): unless the
 * fired errors are asked for (errs_out / err_fd), the model is sampled by the event engine (GSTIM_ENGINE=interp switches
 * that off). Arrays of the table as in gstim_get_response_table (site_group = index of the mechanism in the flattened
 * model); what = 8: one word, the tile height. */
/* Replays recorded errors instead of sampling (DemSampler::resample with replay_errors, src/stim/simulators/dem_sampler.inl:52-130;
 * sample_dem --replay_err_in): errors [shots][ceil(E/8)] bit-packed host rows in, detector / observable rows out
 * (bit-packed, host; NULL to skip; strides in bytes, 0 = dense). No randomness is consumed. */
int gstim_dem_replay(gstim_dem_sampler *s, uint64_t shots, const void *errors, int64_t errors_stride, void *dets_out, int64_t dets_stride,
                     void *obs_out, int64_t obs_stride);
int gstim_dem_get_response_table(gstim_dem_sampler *s, int what, uint32_t *words, size_t *n_words);
/* Flip counts of the D + L output bits (detectors, then observables) and of adjacent pairs over `shots` fresh shots,
 * reduced on the device (the statistic of the parity tests; see gstim_bit_counts). pair_host may be NULL. */
int gstim_dem_bit_counts(gstim_dem_sampler *s, uint64_t shots, uint64_t *single_host, uint64_t *pair_host);
/* Streams to files in any result format; a negative fd skips that output. The outputs of one chunk are written in the
 * reference's order (errors, observables, detectors). */
int gstim_dem_sample_to_fd(gstim_dem_sampler *s, uint64_t shots, int det_fd, const char *det_format, int obs_fd, const char *obs_format,
                           int err_fd, const char *err_format);

/* ---- sampling engines ------------------------------------------------------------------------------------------
 * The library has two implementations of the path, chosen per sampler:
 *   GSTIM_ENGINE_INTERPRETER  the frame interpreter (interp.cu): x/z frame words pushed through every lowered
 *                             instruction, like FrameSimulator::do_circuit (src/stim/simulators/frame_simulator.inl:166-170).
 *   GSTIM_ENGINE_EVENTS       the event-driven engine (sparse.cu): the frame simulation is GF(2)-linear in its random
 *                             bits, so every (noise site, Pauli) has a fixed response - the set of output bits it
 *                             flips - computed once per circuit by propagating sensitivities backwards through the
 *                             lowered program (response.cc). A shot is the XOR of the responses of the events that fired,
 *                             sampled with geometric skipping over (site x shot) like RareErrorIterator
 *                             (src/stim/util_bot/probability_util.cc:23-43). Same distribution as the reference for
 *                             every eligible circuit; bit-exact on deterministic circuits (p in {0, 1}).
 * A circuit is eligible for the event engine unless it contains an ELSE_CORRELATED_ERROR chain or its table would be
 * too large; collapse randomisation that reaches an output (non-deterministic detectors, measurement sampling) becomes
 * p = 1/2 sites. GSTIM_ENGINE_AUTO (default; env GSTIM_ENGINE=auto|interp|events overrides) picks the event engine when
 * the circuit is eligible and its cost model (expected events per shot vs lowered items) favours it.
 * The two engines draw different random streams: with GSTIM_ENGINE_EVENTS the stream is a function of
 * (seed, shot offset) only, and shot offsets must be multiples of the tile height (gstim_engine_info). */
#define GSTIM_ENGINE_AUTO 0
#define GSTIM_ENGINE_INTERPRETER 1
#define GSTIM_ENGINE_EVENTS 2
typedef struct gstim_engine_info {
    int32_t eligible;          /* 1 when the event engine can sample this circuit */
    int32_t favoured;          /* 1 when GSTIM_ENGINE_AUTO picks it */
    int32_t last_engine;       /* engine of the most recent sampling call */
    uint32_t tile_shots;       /* shots per thread-block tile (power of two <= 128) */
    uint32_t blocks_per_sm;
    uint32_t num_classes;      /* site classes (probability + outcome chooser) */
    uint32_t num_slices;       /* RNG slices per tile */
    uint32_t max_response;     /* most output bits flipped by one (site, outcome) */
    uint64_t num_sites;        /* noise + live collapse sites */
    uint64_t num_entries;      /* (site, outcome) table entries of 16 bytes */
    uint64_t device_entries;   /* entries stored on the device (tables beyond GSTIM_TABLE_COMPRESS_MB are folded to one round) */
    uint64_t overflow_words;
    double events_per_shot;    /* expected events per shot */
    double flips_per_shot;     /* expected bit flips per shot */
    char why_not[96];          /* reason when not eligible */
} gstim_engine_info;
int gstim_set_engine(gstim_sampler *s, int engine);
int gstim_get_engine_info(const gstim_sampler *s, gstim_engine_info *out);
/* Copies one array of the response table (tests / the oracle re-derive the responses by forward injection and restate
 * the sampling from it). what: 0 classes (24 words each: lam lo, lam hi, inv, sh, kind, n_out, thr[15], n_sites, entry0, dense threshold),
 * 1 entries (4 words each: output ids - detector d, observable D + l, measurement m - 0xFFFFFFFF = empty, a 4th word with
 * bit 31 = offset into the overflow array), 2 overflow (count, ids...), 3 site noise group (bit 31: collapse site of that
 * measure group), 4 site index in its group (collapse sites: logical qubit), 5 representative chooser word per class
 * outcome, 6 slices (4 words each: class, trials = sites * tile_shots, first entry, 0), 7 round structure per class (4 words: a, p, n,
 * delta: entries of site s + p = entries of site s with delta added to the detector ids, for s in [a, a + n - p); p = 0 none).
 * Call with words == NULL to get the length in *n_words. */
int gstim_get_response_table(const gstim_sampler *s, int what, uint32_t *words, size_t *n_words);
/* Host-only (no GPU needed): lowers the circuit and builds its response table; `what` selectors 0-5 and 7 as above.
 * gstim_response_table_info fills the table statistics of a gstim_engine_info struct: eligible, why_not, sites, entries, .... */
typedef struct gstim_response_table gstim_response_table;
int gstim_response_table_create(const char *circuit_text, size_t text_len, int mode, gstim_response_table **out);
void gstim_response_table_destroy(gstim_response_table *t);
int gstim_response_table_info(const gstim_response_table *t, gstim_engine_info *out);
int gstim_response_table_get(const gstim_response_table *t, int what, uint32_t *words, size_t *n_words);

/* ---- measurements -> detection events ----------------------------------------------------
 * Replaces: measurements_to_detection_events_helper<W>          src/stim/simulators/measurements_to_detection_events.inl:30-131
 *           stim m2d (stream_measurements_to_detection_events)   src/stim/simulators/measurements_to_detection_events.inl:147-330
 *           stim.CompiledMeasurementsToDetectionEventsConverter.convert
 *                                                                src/stim/simulators/measurements_to_detection_events.pybind.cc:78-137
 * The circuit's noise is ignored (the reference converts with circuit.aliased_noiseless_circuit()). With
 * skip_reference_sample = 0 the noiseless reference sample (host stabilizer simulation) decides which detectors are
 * inverted. All rows are shot-major, bit-packed little-endian, in HOST memory; strides in bytes (0 = dense):
 * measurements [shots][ceil(M/8)], sweep_bits [shots][ceil(S/8)] or NULL (all zero), dets_out [shots][ceil(n/8)] with
 * n = D (+ L with GSTIM_APPEND_OBS), obs_out [shots][ceil(L/8)] or NULL (GSTIM_SEPARATE_OBS). */
typedef struct gstim_m2d gstim_m2d;
int gstim_m2d_create_from_text(const char *circuit_text, size_t text_len, int skip_reference_sample, int device, gstim_m2d **out);
void gstim_m2d_destroy(gstim_m2d *h);
int gstim_m2d_get_sizes(const gstim_m2d *h, uint64_t *num_measurements, uint64_t *num_detectors, uint64_t *num_observables,
                        uint64_t *num_sweep_bits);
int gstim_m2d_convert(gstim_m2d *h, uint64_t shots, uint32_t flags, const void *measurements, int64_t meas_stride, const void *sweep_bits,
                      int64_t sweep_stride, void *dets_out, int64_t dets_stride, void *obs_out, int64_t obs_stride);

/* ---- interactive flip simulator -----------------------------------------------------------
 * Replaces: stim.FlipSimulator over FrameSimulator<W>   src/stim/simulators/frame_simulator.pybind.cc:507-1561,
 *           src/stim/simulators/frame_simulator.inl:153-1113 (safe_do_circuit / do_gate one fragment at a time).
 * A batch of shots whose frame (x, z rows per qubit), measurement flip record, detector flips and observable flips live
 * in device memory between calls. Tables are bit-major: a row = row_words uint32 words over the instances (instance i =
 * bit i % 32 of word i / 32; bits beyond batch_size are unspecified). Table selectors: 0 x, 1 z, 2 measurement flips,
 * 3 detector flips, 4 observable flips. */
typedef struct gstim_flipsim gstim_flipsim;
int gstim_flipsim_create(uint64_t batch_size, int disable_stabilizer_randomization, uint64_t num_qubits, uint64_t seed, int device,
                         gstim_flipsim **out);
void gstim_flipsim_destroy(gstim_flipsim *h);
/* FlipSimulator.copy (frame_simulator.pybind.cc:1476-1488): an independent simulator with the same frame, records and sizes.
 * copy_rng != 0: the copy continues the same random stream as `src` (seed is ignored); else it starts the stream of `seed`. */
int gstim_flipsim_copy(const gstim_flipsim *src, int copy_rng, uint64_t seed, gstim_flipsim **out);
int gstim_flipsim_sizes(const gstim_flipsim *h, uint64_t *batch_size, uint64_t *num_qubits, uint64_t *num_measurements,
                        uint64_t *num_detectors, uint64_t *num_observables, uint64_t *row_words);
/* FlipSimulator.do: applies a circuit fragment (any instructions, REPEAT blocks, noise, detectors; rec[-k] looks back
 * over everything measured so far). */
int gstim_flipsim_do_text(gstim_flipsim *h, const char *circuit_text, size_t text_len);
/* Rows [first_row, first_row + n_rows) of a table to / from host memory (n_rows * row_words words). set_rows with
 * xor_in != 0 XORs instead of overwriting; writing measurement rows at first_row == num_measurements appends them
 * (append_measurement_flips); writing x / z rows beyond num_qubits grows the simulator. */
int gstim_flipsim_get_rows(gstim_flipsim *h, int what, uint64_t first_row, uint64_t n_rows, uint32_t *words_out);
int gstim_flipsim_set_rows(gstim_flipsim *h, int what, uint64_t first_row, uint64_t n_rows, const uint32_t *words_in, int xor_in);
/* broadcast_pauli_errors: pauli (0 I, 1 X, 2 Y, 3 Z) applied where mask (n_rows qubit rows, bit-major) is set, each with
 * probability p. generate_bernoulli_samples: n_words words of Bernoulli(p) bits. clear: back to the start state. */
int gstim_flipsim_broadcast(gstim_flipsim *h, int pauli, const uint32_t *mask_words, uint64_t n_rows, double p);
int gstim_flipsim_bernoulli(gstim_flipsim *h, uint64_t n_words, double p, uint32_t *words_out);
int gstim_flipsim_clear(gstim_flipsim *h);

/* Pins the number of 128-shot columns per thread block (0 = choose per call from the shot count, the default). The
 * random stream is a function of (seed, shot offset, columns per block): callers that split one global shot range over
 * several handles / GPUs and need the union to equal a single-handle run pin the same value everywhere and keep every
 * shard a multiple of columns * 128 shots. `columns` must be a multiple of lanes_per_item and <= max_columns
 * (gstim_get_stats). */
int gstim_set_block_columns(gstim_sampler *s, uint32_t columns);

/* Integer-ALU roofline probe: runs a LOP3 microbenchmark on `device` and reports 32-bit logic lane-operations per clock
 * per SM, per second for the whole chip, and the SM clock during the probe (MHz). Measurement aid for bench.py
 * (SURVEY.md 8d); not on the sampling path. */
int gstim_measure_lop3_peak(int device, double *lane_ops_per_clk_per_sm, double *lane_ops_per_sec, double *sm_mhz);

/* Kernel launches issued by the most recent sampling call (for bench accounting). */
int gstim_last_launch_count(const gstim_sampler *s, uint64_t *launches);

/* 128-shot columns per thread block (K) used by the most recent sampling call. Together with the seed
 * and the shot offset this pins the random stream (DESIGN.md "RNG addressing"). */
int gstim_last_block_columns(const gstim_sampler *s, uint32_t *columns);

/* Device time (CUDA events on the library's stream, ms) from the first kernel start to the last kernel
 * end of the most recent sampling call. */
int gstim_last_call_ms(const gstim_sampler *s, float *ms);

/* Device-time (CUDA events, ms) of the interpreter and transposer kernels in the most recent call. */
int gstim_last_kernel_ms(const gstim_sampler *s, float *interp_ms, float *transpose_ms);

#ifdef __cplusplus
}
#endif
#endif /* GSTIM_H */

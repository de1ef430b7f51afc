/*
 * fastsk_b200 -- C ABI of the B200-native gapped k-mer kernel-matrix build.
 *
 * This is the drop-in boundary for the reference's pybind11 module `fastsk._fastsk`
 * (/root/reference/src/fastsk/_fastsk/bindings.cpp:12-51).  Each entry point names the
 * reference interface it replaces.  Plain pointers and sizes only; every buffer is owned by
 * the caller; a handle is used from one host thread at a time.  All functions return
 * FSK_OK (0) or an error code; fsk_last_error() gives the message.
 *
 * Sequences are passed flat: `codes` = all sequences back to back (train rows first, then
 * test rows), `offsets[i] .. offsets[i+1]` = sequence i.  Any non-negative int32 values are
 * accepted (only equality of characters matters; the library re-codes them densely, which
 * the reference requires of its caller: fastsk.cpp:70-85, shared.cpp:171-172).
 *
 * There is no CPU fallback: every compute entry point fails with FSK_ECUDA when no sm_100
 * device is usable.
 */
#ifndef FASTSK_B200_H
#define FASTSK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsk_handle fsk_handle;

enum {
    FSK_OK = 0,
    FSK_EINVAL = 1,   /* bad argument (Python: ValueError) -- replaces the reference's printf + exit(1), shared.cpp:380-412 */
    FSK_ECUDA = 2,    /* CUDA runtime / device error (Python: RuntimeError) */
    FSK_ENOMEM = 3,   /* device or host memory exhausted */
    FSK_ESTATE = 4    /* call out of order (e.g. getter before compute) */
};

/* dtype tags for fsk_partial_buffer */
enum { FSK_DT_I64 = 0, FSK_DT_F64 = 1 };

/* ---- lifetime ------------------------------------------------------------------------- */

/* FastSK::FastSK(g, m, t, approx, delta, max_iters, skip_variance)  -- fastsk.cpp:19-28,
 * defaults bindings.cpp:14-22 (t=-1 => 20 streams, fastsk_kernel.cpp:54-60).
 * `t` is the number of VIRTUAL streams of approx mode; it never limits GPU use. */
int fsk_create(fsk_handle** out, int g, int m, int t, int approx, double delta, int max_iters, int skip_variance);
void fsk_destroy(fsk_handle* h);
/* message of the last failure on `h` (or of the last failed fsk_create when h == NULL) */
const char* fsk_last_error(const fsk_handle* h);
const char* fsk_version(void);

/* ---- configuration (before fsk_compute / fsk_upload) ------------------------------------ */

int fsk_set_device(fsk_handle* h, int device);              /* CUDA ordinal; default 0; may be called again later (moves the handle) */
/* In-process multi-GPU: one fsk_compute call fans out over `n` GPUs, one host thread per GPU -- the reference fans one
 * compute_kernel call out over T std::threads (fastsk_kernel.cpp:54-94).  The combinations (integer modes) or virtual
 * streams (variance mode) are dealt round-robin to the devices; the partial kernels are merged (fastsk_kernel.cpp:285-315)
 * inside the normalisation kernel, which reads every peer's partial over NVLink; every device normalises its share of the
 * output rows and the getters copy all shares to the host at the same time, each over its own PCIe link.
 * devices == NULL with n == -1 selects every visible GPU.  No NCCL, no torch.  Alternative to fsk_set_shard. */
int fsk_set_devices(fsk_handle* h, const int* devices, int n);
/* The reference shuffles the C(g,m) combinations with std::default_random_engine seeded by
 * time(0) (fastsk_kernel.cpp:31-47).  Same libstdc++ calls here, with an injectable seed ...  */
int fsk_set_seed(fsk_handle* h, uint64_t seed);
/* ... or an explicit processing order (combination numbers in the lexicographic order of
 * getCombinations, shared.cpp:347-360).  n may be smaller than C(g,m).  n == 0 clears it. */
int fsk_set_combo_sequence(fsk_handle* h, const int32_t* combos, int64_t n);
/* Multi-GPU: this handle builds shard `rank` of `world` (combinations in the integer modes,
 * virtual streams in the variance mode).  Default 0 of 1.  Between fsk_build_partial and fsk_finalize
 * the caller either hands every rank the others' partial buffers (fsk_set_peer_partials: the merge then
 * happens inside the normalisation, see below) or sums the buffers itself (one NCCL all-reduce). */
int fsk_set_shard(fsk_handle* h, int rank, int world);
/* One process per GPU (torchrun): after fsk_build_partial every rank exports its partial kernel (a CUDA IPC handle of
 * FSK_IPC_HANDLE_BYTES bytes), the caller gathers the handles of all ranks (any transport: torch.distributed, MPI, a file),
 * fsk_set_peer_partials maps the other ranks' buffers over NVLink, and -- after a barrier: every rank's build must be
 * complete -- fsk_finalize then normalises only this rank's share of the output rows (fsk_output_rows), summing the partial
 * kernels of all ranks as it reads them: the merge of fastsk_kernel.cpp:285-315 fused into the normalisation, no reduced
 * copy of K, no collective library.  fsk_release_peers forgets the peers (also done by the next upload and by fsk_destroy); a
 * rank must not upload again or be destroyed while another rank still reads its partial (barrier first).  The mappings
 * themselves are cached per process and re-used when the same buffers come back (they do: the device blocks are cached
 * too); fsk_trim_cache closes them -- call it on every rank, after a barrier, before a rank's memory should really be freed. */
#define FSK_IPC_HANDLE_BYTES 64
int fsk_ipc_export_partial(fsk_handle* h, void* handle_out);
int fsk_set_peer_partials(fsk_handle* h, const void* handles /* world x FSK_IPC_HANDLE_BYTES, rank order */, int world);
/* the same for handles that live in ONE process (the caller's own threads, or two shards on one GPU): plain device pointers
 * from fsk_partial_buffer, peer access already enabled by the caller where the devices differ */
int fsk_set_peer_pointers(fsk_handle* h, void* const* parts /* world pointers, rank order */, int world);
int fsk_release_peers(fsk_handle* h);
/* Share of the output rows per rank in a sharded finalisation (default: equal).  On a box whose GPUs do not reach host memory
 * equally fast (measured here: GPUs 4-7 together 173 GB/s, GPUs 0-3 together 71 GB/s, all eight 91 GB/s) the results are handed
 * back through the fast ones: weight 0 = this rank normalises and copies no rows (its partial kernel is still read by the
 * others over NVLink).  fsk_probe_d2h times one device -> host copy of `bytes` from `device` into pinned memory; run it on
 * all ranks at once (and on subsets) to find the weights -- fastsk_b200/fastsk.py::output_weights does. */
int fsk_set_output_weights(fsk_handle* h, const double* weights /* world values, or NULL for equal shares */, int world);
int fsk_probe_d2h(int device, size_t bytes, double* seconds);
/* rows of the train / test kernels this handle holds after fsk_finalize (all of them unless the finalisation was sharded;
 * a team of in-process devices reports all rows: its getters collect every member's share) */
int fsk_output_rows(fsk_handle* h, int64_t* train_r0, int64_t* train_nr, int64_t* test_r0, int64_t* test_nr);
/* tuning / diagnostics: "batch" (combinations per launch group, 0 = auto), "profile" (1 = time
 * every kernel class with CUDA events and count entries / runs / pair updates), "acc_path" (0 = auto,
 * 1 = global RED on the packed triangle, 2 = row-stationary shared-memory accumulate, 3 = dense regime:
 * per-sequence k-mer counts contracted as K += C C^T by tcgen05 tensor-core MMAs, no sort; needs at most
 * 12 key bits and 2048 windows per sequence, chosen automatically when its cost model wins); unknown
 * keys give FSK_EINVAL.  "heavy_tau": -1 off, 0 auto, > 0 forced run-length threshold above which a run's update goes to the
 * tensor cores instead of the row path; "heavy_cap": columns of that contraction's list (0 auto); "gemm_shape": 0 auto, 1 or 2
 * output tiles per CTA of the contraction; "seg_fused": 2 = fused last sort pass + segmentation for two-digit keys (opt-in);
 * "seg_dir": 0 auto, 1 off, 2 on: directory form of the segmentation (one entry per k-mer and block of rows
 * instead of one filed task per record; at most 16 key bits), "dir_blocks": target number of row blocks (32); "seg_lean": 0 auto, 1 off, 2 on: register-blocked segmentation kernel
 * (records that carry the sequence id); "count_updates": 0 = "profile"
 * times the kernels but does not count entries / runs / pair updates;
 * "spec_depth": iterations of a virtual stream per launch group in variance mode (0 auto); "wf_regs": 1 (default) the tensor-core
 * Welford contraction keeps the running means in registers over a round's slots, 0 streams them through L2 every slot; "dense_u8":
 * 1 (default) byte operands / int32 accumulators in the dense regime when no sequence has more than 255 windows, 0 always fp16 / fp32;
 * "acc_prefetch", "acc_unroll", "wave", "rows_threads", "pad", "overlap", "safe_rank": tuning of the row path (defaults are the
 * measured optimum); test hooks: "acc_cols" (forced column-window width of the row path), "ids32" (32-bit id stream) */
int fsk_set_option(fsk_handle* h, const char* key, int64_t value);

/* ---- compute ---------------------------------------------------------------------------- */

/* FastSK::compute_kernel(Xtrain, Xtest) -- fastsk.cpp:30-118 (n_test == 0: compute_train,
 * fastsk.cpp:120-188).  Host buffers in; equals fsk_upload + fsk_build_partial + fsk_finalize.
 * FSK_EINVAL if g is longer than the shortest sequence (reference: exit(1), fastsk.cpp:53-58). */
int fsk_compute(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test);

/* The same with the two arguments of compute_kernel(Xtrain, Xtest) as they come: train sequences in (codes, offsets[0..n_train]),
 * test sequences in (codes_test, offsets_test[0..n_test]) -- no concatenation on the caller's side (codes_test == NULL: all N
 * sequences are in codes, as above). */
int fsk_compute_split(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, const int32_t* codes_test,
                      const int64_t* offsets_test, int64_t n_test);
int fsk_upload_split(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, const int32_t* codes_test,
                     const int64_t* offsets_test, int64_t n_test);

/* staged form of the same call (multi-GPU, benchmarking with resident inputs) */
int fsk_upload(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test);
int fsk_build_partial(fsk_handle* h);   /* KernelFunction::kernel_build_parallel for this shard, fastsk_kernel.cpp:145-322 */
/* device pointer / element count / dtype of this rank's partial (packed lower triangle) */
int fsk_partial_buffer(fsk_handle* h, void** dev_ptr, int64_t* n_elems, int* dtype);
int fsk_finalize(fsk_handle* h);        /* normalisation, fastsk_kernel.cpp:96-103 */

/* Benchmark hook: accumulate the integer partial kernels of the given combinations into the
 * resident partial buffer (inputs already uploaded; no normalisation; asynchronous on the
 * handle's stream unless `sync`). */
int fsk_accumulate_combos(fsk_handle* h, const int32_t* combos, int64_t n, int sync);
int fsk_reset_partial(fsk_handle* h);
int fsk_stream(fsk_handle* h, void** cuda_stream);          /* cudaStream_t of the handle */
int fsk_synchronize(fsk_handle* h);

/* ---- results ------------------------------------------------------------------------------ */

int fsk_shape(fsk_handle* h, int64_t* n_train, int64_t* n_test, int64_t* nfeat, int64_t* n_combos);
/* FastSK::get_train_kernel -- fastsk.cpp:190-200: n_train x n_train row-major fp64.  `out` is always the full matrix; a
 * handle whose finalisation was sharded writes only its rows (fsk_output_rows) -- ranks that share one host buffer
 * (shared memory) fill it together, each over its own PCIe link. */
int fsk_get_train_kernel(fsk_handle* h, double* out);
/* FastSK::get_test_kernel -- fastsk.cpp:202-217: n_test x n_train row-major fp64 */
int fsk_get_test_kernel(fsk_handle* h, double* out);
/* the reference's `double* K`: normalised packed lower triangle, N(N+1)/2 */
int fsk_get_kernel_packed(fsk_handle* h, double* out);
/* unnormalised kernel before fastsk_kernel.cpp:96-103; the i64 form exists in the integer
 * modes (exact, approx+skip_variance) only */
int fsk_get_unnormalised_i64(fsk_handle* h, int64_t* out);
int fsk_get_unnormalised_f64(fsk_handle* h, double* out);
/* device-resident results for DLPack hand-off (valid until the next compute / destroy) */
int fsk_train_kernel_device(fsk_handle* h, void** dev_ptr);
int fsk_test_kernel_device(fsk_handle* h, void** dev_ptr);
/* FastSK::get_stdevs -- fastsk.cpp:219-221 (stream 0's sd sequence, fastsk_kernel.cpp:243-250) */
int fsk_get_stdevs(fsk_handle* h, double* out, int64_t cap, int64_t* n);
/* FastSK::save_kernel -- fastsk.cpp:223-237: N lines of "%d:%e " (1-based column ids) */
int fsk_save_kernel(fsk_handle* h, const char* path);
/* the combination order in use (after shuffle / fsk_set_combo_sequence) */
int fsk_get_queue(fsk_handle* h, int32_t* out, int64_t cap, int64_t* n);

/* What this shard (fsk_set_shard) will process: combination numbers in the integer modes, virtual
 * stream ids in the variance mode.  Needs no device. */
int fsk_get_shard_work(fsk_handle* h, int32_t* out, int64_t cap, int64_t* n);

/* ---- host memory, device memory ----------------------------------------------------------- */

/* pinned host buffers for inputs / outputs (DMA at PCIe speed), registration of an existing range (e.g. shared memory),
 * and release of the library's process-wide cache of device blocks (buffers are re-used across computes and handles) */
int fsk_host_alloc(void** out, size_t bytes);
int fsk_host_free(void* p);
int fsk_host_register(void* p, size_t bytes);
int fsk_host_unregister(void* p);
int fsk_trim_cache(void);
/* self-test: the variance mode divides by the iteration number with a reciprocal + two fused multiply-adds instead of the
 * general division routine; this compares the two on `n` pseudo-random operand pairs on the device (must report 0) */
int fsk_selftest_division(int device, uint64_t seed, uint64_t n, uint64_t* mismatches);

/* ---- statistics --------------------------------------------------------------------------- */

typedef struct fsk_stats {
    int64_t n_seq, nfeat, n_pairs, n_combos_total;
    int64_t combos_done;        /* combinations processed by this handle since upload */
    int64_t pair_updates;       /* sum over runs of d(d+1)/2 (only counted when "profile" is on) */
    int64_t entries;            /* distinct (k-mer, sequence) cells seen ("profile") */
    int64_t runs;               /* distinct k-mers seen ("profile") */
    int64_t kernel_launches;    /* CUDA kernels launched by this handle since upload */
    int32_t key_bits, id_bits, record_bytes, sort_passes, alphabet, bits_per_char, batch, acc_bytes;
    /* per kernel class, milliseconds on the handle's stream ("profile" only) */
    double ms_pack, ms_sort, ms_segment, ms_accumulate, ms_welford, ms_normalise, ms_total;
    /* accumulate path in use: 1 global RED, 2 shared-memory rows, 3 dense tensor-core contraction (no sort) */
    int32_t acc_path;
    /* runs longer than heavy_tau records (0 = feature off) leave the row path: their update is one column of a tensor-core
     * contraction; heavy_runs counts them since upload (pair_updates then counts the row path's share only) */
    int32_t heavy_tau;
    int64_t heavy_runs;
    int32_t n_devices;          /* GPUs driven by this handle (fsk_set_devices); counts above add up over them, times are the slowest's */
    int32_t seg_mode;           /* segmentation: 0 one task filed per record, 1 run directory (key spaces <= 2^16), 2 fused bucket form; + 4 when
                                   the register-blocked kernel (fsk_segment.cuh) does it */
    int32_t dense_mode;         /* dense regime (acc_path 3): 1 fp16 operands / fp32 accumulators, 2 byte operands / int32 accumulators
                                   (no sequence has more than 255 windows); + 4 when the variance mode's Welford contraction keeps the
                                   running means in registers over a round; 0 outside the dense regime */
} fsk_stats;
int fsk_get_stats(fsk_handle* h, fsk_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* FASTSK_B200_H */

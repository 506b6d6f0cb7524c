/*
 * gsa_cuda.h -- C ABI of libgsa_cuda.so: the B200 (sm_100a) implementation of the Sobol path of
 * SciML/GlobalSensitivity.jl,   gsa(f, Sobol(order, nboot, conf_level), A, B; batch=true).
 *
 * The reference has no FFI for this path (it is pure Julia); the seam this library replaces is the
 * Julia method   gsa(f, method::Sobol, A::AbstractMatrix, B::AbstractMatrix; batch, Ei_estimator)
 * (reference src/sobol_sensitivity.jl:110-168) and its internal analysis function
 * gsa_sobol_all_y_analysis (:169-405).  INTEGRATION.md shows the `ccall` shim a maintainer adds.
 *
 * Conventions
 *   - every entry point returns an int status (GSA_OK == 0); gsa_last_error() gives the message of
 *     the last failure on the calling thread.  No C++ exception crosses this boundary.
 *   - all matrices use Julia's column-major layout verbatim:
 *         A[i + d*c]            parameter i of sample (column) c, d values of a sample contiguous
 *         Y[o + nout*c]         output o of column c
 *         S1/ST[o + nout*k]     (a length-d vector when nout == 1)
 *         S2[k + d*j + d*d*o]   only k<j is non-zero (reference :194-200, :256-264)
 *   - counts that scale with the number of samples are int64_t.
 *   - the caller owns every buffer it passes; the library keeps no caller pointer after a call returns.
 *   - a handle is bound to one CUDA device and serialises its own calls (one in flight per handle).
 */
#ifndef GSA_CUDA_H
#define GSA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSA_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define GSA_API __attribute__((visibility("default")))
#else
#define GSA_API
#endif

/* status codes */
enum {
    GSA_OK = 0,
    GSA_ERR_INVALID = 1,     /* bad argument (message says which) */
    GSA_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
    GSA_ERR_UNSUPPORTED = 3, /* valid request outside what this build implements */
    GSA_ERR_NOMEM = 4,
    GSA_ERR_NCCL = 5         /* reserved (no collective library is used any more) */
};

/* Ei_estimator of the reference (src/sobol_sensitivity.jl:112, :202-236) */
enum { GSA_EI_HOMMA1996 = 0, GSA_EI_SOBOL2007 = 1, GSA_EI_JANSEN1999 = 2, GSA_EI_JANON2014 = 3 };

/* where a data pointer lives */
enum { GSA_LOC_HOST = 0, GSA_LOC_DEVICE = 1 };

/* built-in device models (gsa_model_lookup names) */
enum {
    GSA_MODEL_ISHIGAMI = 0,    /* params {a, b};  y = sin x1 + a sin^2 x2 + b x3^4 sin x1          */
    GSA_MODEL_SOBOL_G = 1,     /* params a[d];    y = prod_i (|4 x_i - 2| + a_i) / (1 + a_i)       */
    GSA_MODEL_LINEAR = 2,      /* params W[nout x d] column-major; y = W x  (multi-output)         */
    GSA_MODEL_ISHI_LINEAR = 3, /* the 2-output model of reference test/sobol_method.jl:146-158     */
    GSA_MODEL_USER_BASE = 1000 /* ids returned by gsa_model_register_fatbin                        */
};

typedef struct gsa_handle gsa_handle; /* opaque */
typedef struct gsa_multi gsa_multi;   /* opaque: one handle per device of a single process */

/* mirrors `struct Sobol` (reference :77-83) + the call's keyword arguments (:112) */
typedef struct gsa_sobol_opts {
    int second_order;   /* 2 in method.order                                   */
    int nboot;          /* replicates; n = N / nboot columns each (:116-127)   */
    double conf_level;  /* used when nboot > 1 (:359)                          */
    int ei_estimator;   /* GSA_EI_*; the reference default is Jansen1999       */
    int input_location; /* GSA_LOC_* of A, B (or Y)                            */
} gsa_sobol_opts;

/* mirrors `SobolResult` (reference :85-92); caller-allocated, a NULL member is skipped.
 * The *_conf_int members are written only when nboot > 1, S2* only when second_order. */
typedef struct gsa_sobol_result {
    double* S1;          /* nout*d   */
    double* S1_conf_int; /* nout*d   */
    double* S2;          /* d*d*nout */
    double* S2_conf_int; /* d*d*nout */
    double* ST;          /* nout*d   */
    double* ST_conf_int; /* nout*d   */
} gsa_sobol_result;

/* ---- library / handle ------------------------------------------------------------------------ */
GSA_API int gsa_version(void);
/* copies the calling thread's last error message (NUL-terminated, truncated to n) */
GSA_API int gsa_last_error(char* buf, size_t n);
/* device < 0: use the calling thread's current CUDA device */
GSA_API int gsa_create(int device, gsa_handle** out);
GSA_API int gsa_destroy(gsa_handle* h);
/* external != 0: all work of the handle is issued on `cuda_stream` (a cudaStream_t; NULL is the legacy default
 * stream), e.g. the caller framework's current stream.  external == 0: back to the handle-owned stream. */
GSA_API int gsa_set_stream(gsa_handle* h, void* cuda_stream, int external);
/* block until the handle's stream is idle */
GSA_API int gsa_synchronize(gsa_handle* h);

/* Pin a host array in place (cudaHostRegister): a Julia `Matrix{Float64}` is pageable, and pageable designs are staged through
 * the library's pinned buffers at ~40 GB/s instead of streamed at the ~55 GB/s PCIe Gen5 delivers.  Worth it when the same arrays
 * are analysed more than once (pinning itself costs about as much as one staged pass).  Unregister before the array is freed. */
GSA_API int gsa_host_register(void* ptr, size_t bytes);
GSA_API int gsa_host_unregister(void* ptr);

/* ---- device-model registry (replaces the Julia closure `f`) ------------------------------------ */
GSA_API int gsa_model_lookup(const char* name, int* model_id);
/* number of outputs of a model for given (nparams, d); replaces `all_y isa AbstractMatrix` (:144) */
GSA_API int gsa_model_nout(gsa_handle* h, int model_id, int nparams, int d, int* nout);
/* Link a user model into the fused kernel.  `image` is a fatbin/cubin/PTX/LTO-IR image built with
 * `nvcc -rdc=true` for sm_100a that defines
 *     extern "C" __device__ void <symbol>(const double* x, int d, const double* p, int np, double* y, int nout);
 * The linked module is cached in the handle.  Returns an id >= GSA_MODEL_USER_BASE. */
GSA_API int gsa_model_register_fatbin(gsa_handle* h, const void* image, size_t len, const char* symbol, int nout,
                              int* model_id);

/* Structured user models (SURVEY 7 "separable-model registry hook").  The reference's `f` is arbitrary (:143), so the generic
 * kernel evaluates it d+2 (2d+2) times per sample at O(d) each.  A model that is a PRODUCT or a SUM of per-coordinate functions
 * can say so and gets the O(d)-per-sample kernel that makes BASELINE config 5 HBM-bound:
 *     GSA_USER_PRODUCT    f(x) = prod_k g(k, x_k)     extern "C" __device__ double <symbol>(int k, double x, const double* p, int np);
 *     GSA_USER_ADDITIVE   f(x) = sum_k  g(k, x_k)     same signature
 *     GSA_USER_GENERIC    the full-vector form of gsa_model_register_fatbin
 * (default symbols: gsa_user_factor / gsa_user_term / gsa_user_model; other names go through a PTX trampoline).  One output.
 * First/total order with Jansen1999, Sobol2007 or Homma1996 run on the separable kernel; second order and Janon2014 run on the
 * generic kernel, which evaluates the whole model through the same per-coordinate function.  Linked cubins are cached per
 * (kind, symbol, image): registering the same image on several handles / devices links once.
 * An image that also carries LTO-IR (nvcc ... -gencode arch=compute_100a,code=lto_100a -gencode arch=compute_100a,code=sm_100a)
 * and uses the default symbol name gets its function INLINED: the kernel a launch needs is LTO-linked on first use (~1.5 s once
 * per kernel shape, cached), and runs at the speed of a built-in model instead of paying a call per coordinate.
 * Register budget: the linked kernels run 3-4 CTAs per SM, so a user function may use at most 128 (generic) / 168 (separable)
 * registers -- nvJitLink refuses the link otherwise and the message says so (build the model with -maxrregcount=128). */
enum { GSA_USER_GENERIC = 0, GSA_USER_PRODUCT = 1, GSA_USER_ADDITIVE = 2 };
GSA_API int gsa_model_register_separable(gsa_handle* h, const void* image, size_t len, const char* symbol, int kind, int* model_id);

/* ---- K1: unrandomised Sobol A/B design --------------------------------------------------------- */
/* Replaces QuasiMonteCarlo.generate_design_matrices(n, lb, ub, SobolSample(), num_mats) (reference
 * call sites src/sobol_sensitivity.jl:408-413, README.md:55-56, test/sobol_method.jl:29-30): one
 * (num_mats*d)-dimensional Gray-code Sobol sequence (Joe-Kuo 21201 table), first emitted 0-based index
 * 2^floor(log2 n), split by rows into num_mats matrices of d x n, scaled (ub-lb)*u + lb.
 * Writes columns [col0, col0+ncols) of every matrix: out[m] has d*ncols doubles (m < num_mats);
 * `out` is an array of num_mats pointers, all in `out_location`.  Bit-exact.  n < 2^31: the sequence index of the last point,
 * 2^floor(log2 n) + n - 1, must fit the 32-bit Gray code the reference's generator (Sobol.jl) uses. */
GSA_API int gsa_sobol_design(gsa_handle* h, int64_t n, int d, int num_mats, const double* lb, const double* ub,
                     int64_t col0, int64_t ncols, double* const* out, int out_location);

/* ---- K2..K4: the path ---------------------------------------------------------------------------- */
/* Doubles per (replicate, output) sufficient-statistics record for these options. */
GSA_API int gsa_sobol_nstat(const gsa_sobol_opts* opts, int d, int* nstat);

/* gsa(model, Sobol(...), A, B; batch=true)  (:110-149 + :169-405) on one GPU.  A, B are d x N.
 * `params` (here and below) is always HOST memory. */
GSA_API int gsa_sobol_run(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                  const double* A, const double* B, int d, int64_t N, gsa_sobol_result* out);

/* Sharded form (one process per GPU).  This rank holds columns [col0, col0+ncols) of the d x N design
 * (A, B point at its first local column).  Fills `partials` -- nboot*nout*nstat doubles in DEVICE memory,
 * zero where this shard has no samples -- with per-replicate sums that are ADDITIVE across shards:
 * all-reduce (sum) them over ranks, then call gsa_sobol_finalize.  Asynchronous on the handle's stream. */
GSA_API int gsa_sobol_partials(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params,
                       int nparams, const double* A, const double* B, int d, int64_t N, int64_t col0,
                       int64_t ncols, double* partials_dev);

/* The sharded form in ONE call per rank: partial sums of the rank's shard, all-reduce over the peer group the handle is connected
 * to (gsa_peer_export / gsa_peer_connect below; no group or a group of one: the shard is the whole design) and the epilogue, so
 * every rank returns the same SobolResult.  Collective: every rank of the group calls it with the same options.  Where the
 * kernel has a fused tail (product-form models) the whole step is ONE kernel launch: its last CTA sums the tile records, pushes
 * them to the peers over NVLink, adds the peers' in rank order and writes the result to pinned host memory. */
GSA_API int gsa_sobol_run_sharded(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                          const double* A, const double* B, int d, int64_t N, int64_t col0, int64_t ncols, gsa_sobol_result* out);
GSA_API int gsa_sobol_run_sharded_generated(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                                    const double* lb, const double* ub, int d, int64_t samples, int64_t col0, int64_t ncols,
                                    gsa_sobol_result* out);

/* gsa(model, Sobol(...), p_range; samples) (:407-417) WITHOUT materialising A and B: the fused kernel generates the Sobol points
 * of the samples it evaluates (K1 fused into K2+K3), so the path has no HBM reads and no design memory at all.  The design is
 * the reference's: one 2*nboot*d-dimensional sequence, replicate r = dimensions [r*d, (r+1)*d) (A) and [(nboot+r)*d, ...) (B)
 * at sample indices 0..samples-1, i.e. N = nboot*samples columns; the points are bit-identical to gsa_sobol_design's.
 * Available where a generating kernel exists (product-form models, S1/ST, Jansen1999); otherwise GSA_ERR_UNSUPPORTED and the
 * caller uses gsa_sobol_design + gsa_sobol_run.  The sharded form takes the global column range [col0, col0+ncols). */
GSA_API int gsa_sobol_run_generated(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                            const double* lb, const double* ub, int d, int64_t samples, gsa_sobol_result* out);
GSA_API int gsa_sobol_partials_generated(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params,
                                 int nparams, const double* lb, const double* ub, int d, int64_t samples, int64_t col0,
                                 int64_t ncols, double* partials_dev);

/* Same sums from a precomputed model output `Y` = f(all_points), nout x (nboot*n*step), block order
 * [A | B | AB_1..AB_d (| BA_1..BA_d)] per replicate (:105-107, :181-190): the estimator-only kernel. */
GSA_API int gsa_sobol_partials_y(gsa_handle* h, const gsa_sobol_opts* opts, const double* Y, int nout, int d,
                         int64_t n, double* partials_dev);

/* Sum of `nparts` partial-sum vectors (HOST memory, parts[g*count + i], count = nboot*nout*nstat) in part order; the four
 * double-double (hi, lo) pairs at the head of every record are added error-free, so the result does not depend on how the
 * design was sharded beyond the last ulp.  What a caller that gathers the shards' sums itself (MPI, torch.distributed
 * all_gather, Distributed.jl) calls before gsa_sobol_finalize. */
GSA_API int gsa_sobol_combine(const double* parts_host, int nparts, size_t count, int nstat, double* out_host);

/* (:318-404) partial sums (HOST memory, nboot*nout*nstat) -> SobolResult.  n = N / nboot. Pure host code. */
GSA_API int gsa_sobol_finalize(const gsa_sobol_opts* opts, const double* partials_host, int nout, int d, int64_t n,
                       gsa_sobol_result* out);

/* gsa_sobol_all_y_analysis (:169-405) on a host or device `Y`: partials_y + finalize. */
GSA_API int gsa_sobol_analyze_y(gsa_handle* h, const gsa_sobol_opts* opts, const double* Y, int nout, int d, int64_t n,
                        gsa_sobol_result* out);

/* fuse_designs (:94-108) + replicate concatenation (:116-134) written to DEVICE/HOST memory, for callers
 * that still evaluate their own closure: P is d x (nboot*n*step). */
GSA_API int gsa_sobol_all_points(gsa_handle* h, const gsa_sobol_opts* opts, const double* A, const double* B, int d,
                         int64_t N, double* P, int p_location);

/* ---- eFAST (reference src/eFAST_sensitivity.jl:75-213) ------------------------------------------------------ */
/* mirrors `eFASTResult` (:70-73): S1, ST are nout x d, column-major (a 1 x d row for a scalar model); caller-allocated */
typedef struct gsa_efast_result {
    double* S1;
    double* ST;
} gsa_efast_result;
/* The search-curve design (:98-135): P is d x (samples*d); block i (columns [i*samples, (i+1)*samples)) is the curve in which
 * parameter i carries the frequency omega[0] = (samples-1) / (2*num_harmonics).  p_range entries are [lb_j, ub_j] (uniform; a
 * degenerate range lb_j == ub_j gives a constant, :113).  phi[d] replaces the reference's `2 rand(rng)` per curve (:111); NULL = 0. */
GSA_API int gsa_efast_design(gsa_handle* h, const double* lb, const double* ub, int d, int64_t samples, int num_harmonics,
                     const double* phi, double* P, int p_location);
/* gsa(model, eFAST(num_harmonics), p_range; samples, batch=true) (:75-166): the design is generated in registers, the device model
 * is evaluated on it, and the spectrum is analysed (:167-213) -- nothing but lb/ub/phi goes in, S1/ST come out. */
GSA_API int gsa_efast_run(gsa_handle* h, int num_harmonics, int model_id, const double* params, int nparams, const double* lb,
                  const double* ub, const double* phi, int d, int64_t samples, gsa_efast_result* out);
/* gsa_efast_all_y_analysis (:167-213) on a host or device all_y = f(design): nout x (samples*d) */
GSA_API int gsa_efast_analyze_y(gsa_handle* h, int num_harmonics, const double* Y, int y_location, int nout, int d, int64_t samples,
                        gsa_efast_result* out);

/* Device times (ms, CUDA events on the handle's stream) of the last partials / design call: the whole call,
 * and its dominant kernel alone (the last fused / estimator launch; 0 if none); plus the number of kernels
 * the call launched.  Used by bench.py's roofline line. */
GSA_API int gsa_last_timing(gsa_handle* h, float* total_ms, float* kernel_ms, int* launches);

/* ---- K5 over NVLink peer memory (one process per GPU) ------------------------------------------------------ */
/* All-reduce of the partial sums without a collective library: every rank exports a mailbox (cudaIpcMemHandle), maps its
 * peers' mailboxes, and ONE kernel per rank pushes its vector into all peers, waits for theirs and adds them in rank order
 * (double-double entries error-free), so all ranks hold bit-identical sums.  Collective calls: every rank of the group must
 * make the same sequence of gsa_peer_allreduce calls with the same count.
 *   gsa_peer_export     allocates this rank's mailbox for vectors of up to max_doubles and writes its 64-byte IPC handle
 *   gsa_peer_connect    handles = world * 64 bytes in rank order (exchanged by the caller, e.g. an all-gather)
 *   gsa_peer_allreduce  in place on `data_dev` (count doubles = records of nstat doubles, see gsa_sobol_nstat), asynchronous on
 *                       the handle's stream; host_out != NULL (pinned host memory) also receives the sums, written by the
 *                       kernel itself (no device-to-host copy to wait for)
 *   gsa_peer_status     0, or 1 after a wait for a peer timed out (GSA_PEER_TIMEOUT_MS, default 10000)
 *   teardown            every rank: gsa_peer_unmap (closes its mappings of the peers' mailboxes), a barrier of the caller's, then
 *                       gsa_peer_disconnect (frees its own mailbox, which no peer maps any more)                         */
#define GSA_PEER_HANDLE_BYTES 64
GSA_API int gsa_peer_export(gsa_handle* h, size_t max_doubles, void* handle_out);
GSA_API int gsa_peer_connect(gsa_handle* h, int rank, int world, const void* handles);
GSA_API int gsa_peer_allreduce(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out);
GSA_API int gsa_peer_status(gsa_handle* h, int* status);
GSA_API int gsa_peer_unmap(gsa_handle* h);
GSA_API int gsa_peer_disconnect(gsa_handle* h);

/* ---- single-process, multi-device form (a plain Julia `ccall` from one process) ------------------ */
/* devices == NULL: devices 0..ndev-1; ndev == 0 (with devices == NULL): every visible device.  Creates one gsa_handle per device and enables peer access from the first device to
 * the others (up to 16 devices). */
GSA_API int gsa_multi_create(const int* devices, int ndev, gsa_multi** out);
GSA_API int gsa_multi_destroy(gsa_multi* m);
GSA_API int gsa_multi_device_count(gsa_multi* m, int* ndev);
/* the i-th per-device handle (owned by `m`), e.g. for gsa_sobol_design on that device */
GSA_API int gsa_multi_handle(gsa_multi* m, int i, gsa_handle** out);
/* gsa_model_register_fatbin on every device of `m`: the same image is linked once per device and gets the same id on all */
GSA_API int gsa_multi_model_register_fatbin(gsa_multi* m, const void* image, size_t len, const char* symbol, int nout,
                                    int* model_id);
GSA_API int gsa_multi_model_register_separable(gsa_multi* m, const void* image, size_t len, const char* symbol, int kind, int* model_id);
/* gsa(model, Sobol(...), A, B; batch=true) (:110-149 + :169-405) with the HOST design sharded by contiguous column ranges
 * over the devices: every device streams and reduces its shard, ONE kernel on the first device gathers and adds the
 * nboot*nout*nstat partial sums of all devices over NVLink peer memory (error-free for the double-double entries) and
 * writes them to pinned host memory; then the host epilogue. */
GSA_API int gsa_multi_sobol_run(gsa_multi* m, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                        const double* A, const double* B, int d, int64_t N, gsa_sobol_result* out);

/* the same over the devices of `m` with the design generated inside the fused kernel (see gsa_sobol_run_generated) */
GSA_API int gsa_multi_sobol_run_generated(gsa_multi* m, const gsa_sobol_opts* opts, int model_id, const double* params,
                                  int nparams, const double* lb, const double* ub, int d, int64_t samples, gsa_sobol_result* out);

#ifdef __cplusplus
}
#endif
#endif /* GSA_CUDA_H */

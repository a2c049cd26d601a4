/* =============================================================================
 * exa_b200.h — C ABI of the B200-native evaluator for the ExaModels.jl per-pattern
 * NLP callback path (obj, cons!, grad!, jac_coord!, jac_structure!, hess_coord!,
 * hess_structure!).
 *
 * This is the drop-in boundary: what a `B200Backend` method of the reference's
 * `build_extension(c::ExaCore; prod)` (/root/reference/src/nlp.jl:898, KA method at
 * ext/ExaModelsKernelAbstractions.jl:33-191) would `ccall` once at model build, and
 * what the callback methods on `AbstractExaModel{T,VT,E}` (ext:212-547; CPU forms at
 * src/nlp.jl:1798-1978) would `ccall` on every solver iteration.  Names and argument
 * order follow the only C-ABI precedent in the reference tree, the "cnlp ABI v0.1"
 * exported by ExaModelsCompiler (`P_obj(id,x*,out*)`, `P_hess(id,x*,y*,obj_weight,
 * vals*)`, ... /root/reference/ExaModelsCompiler/src/ExaModelsCompiler.jl:1564-1732),
 * with a handle instead of an id, DEVICE pointers instead of host pointers, and a
 * CUDA stream.
 *
 * Conventions (same as cnlp ABI v0.1, ExaModelsCompiler.jl:181-186):
 *   - every function returns an int status, 0 on success; nothing throws across
 *     the boundary; `exb_last_error()` gives the message of the last failure on the
 *     calling thread;
 *   - indices are 1-based; the Hessian is the lower triangle in COO form with
 *     duplicates ("partially compressed", src/simdfunction.jl:78-100);
 *   - outputs are fully overwritten (no caller pre-zero needed; the reference zero
 *     fills, ext:317,343,521);
 *   - work is enqueued on the caller's stream and NOT synchronised, except where a
 *     scalar is returned to the host (`exb_obj`), mirroring the KA callbacks, which
 *     never call `synchronize`;
 *   - a handle is thread-compatible (one caller at a time, on ONE stream at a time: its scratch -- conbuffer, gradient buffer,
 *     objective partials, host staging -- is per handle, so `exb_cons` on one stream and `exb_jprod` on another would race);
 *     handles are independent;
 *   - the first call of a value callback ranks the launch-shape variants of its kernel and synchronises the stream ONCE
 *     (exb_tune / EXB_FLAG_TUNE_AT_CREATE move that to model build);
 *   - parity with the reference is defined on finite points: multiplications by the structural zeros of the linear operators are
 *     folded at build time, so a non-finite adjoint yields 0 where the reference's literal `adj * zero(x)` yields NaN.
 *
 * There is no CPU fallback: every evaluation entry point fails with EXB_ERR_CUDA if
 * no sm_100 device / kernel module is available.
 *
 * --- §IR: the pattern IR (int64 little-endian word stream) ------------------
 * The language-neutral image of `SIMDFunction` (src/simdfunction.jl:21-30) +
 * `Objective / Constraint / ConstraintAugmentation` (src/nlp.jl:107-177), patterns in
 * ADD ORDER.  Emitters: examodels.jl_b200/nlp.py (Python host) and the Julia shim of
 * INTEGRATION.md.
 *
 *   header : MAGIC(0x0031425845 "EXB1") VERSION(1) nvar npar npatterns ndatabufs
 *   pattern: kind(0 obj|1 con|2 aug) nitr itr_kind(0 range|1 AoS) range_start
 *            databuf(-1 if range) stride_bytes
 *            nfields { byte_offset type(0 i64|1 f64|2 i32|3 f32) }*nfields
 *            o0 o1 o2            (-1: assign by the counter rules of nlp.jl:1474-1482,
 *                                 1597-1611,1730-1738)
 *            base                (aug: index of the base Constraint pattern, else -1)
 *            nidx { idx_root }*nidx { dim }*nidx      (aug row = o0 + idxx(coord, dims),
 *                                                      nlp.jl:1986-2001,2012-2015)
 *            nnodes { tag a b payload }*nnodes  root
 *            ncomp1 {comp1}*  ncomp2 {comp2}*   (optional; recomputed and cross-checked)
 *   node tags: 0 CONST_I(payload) 1 CONST_F(payload=f64 bits) 2 DATA_SELF
 *              3 DATA_FIELD(a=field) 4 VAR(a=index node) 5 PAR(a=index node)
 *              6 NULL(payload=f64 bits) 7 OP1(a=child, payload=op) 8 OP2(a,b,payload=op)
 *              9 VAL(payload)   -- Val{p} exponent (specialization.jl:199-202)
 *   OP1 codes follow the order of src/functionlist.jl:6-60 (0 '+', 1 '-', 2 inv, ...),
 *   OP2 codes the order of src/functionlist.jl:71-81 (+ - * / ^ atan hypot max min).
 *   The SpecialFunctions extension (ext/functionlist.jl:6-126) continues both tables in its
 *   registration order: OP1 52.. erf erfc erfi erfcx digamma trigamma invdigamma gamma airyai
 *   airybi airyaiprime airybiprime besselj0 bessely0 besselj1 bessely1 dawson erfinv erfcinv;
 *   OP2 9 beta, 10 logbeta.
 *   Children precede parents.
 * ============================================================================= */
#ifndef EXA_B200_H
#define EXA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXB_ABI_VERSION 2 /* 2: exb_options.tune_x0 */

enum {
  EXB_OK = 0,
  EXB_ERR_HANDLE = 1,   /* invalid handle            (cnlp ABI: id invalid -> 1) */
  EXB_ERR_INTERNAL = 2, /* caught exception          (cnlp ABI: exception  -> 2) */
  EXB_ERR_IR = 3,       /* malformed / inconsistent IR */
  EXB_ERR_COMPILE = 4,  /* nvcc failed or module could not be loaded */
  EXB_ERR_CUDA = 5,     /* CUDA runtime error / no device */
  EXB_ERR_ARG = 6
};

typedef struct exb_plan exb_plan;   /* host-only analysis of an IR (no GPU needed) */
typedef struct exb_model exb_model; /* a model resident on one GPU */

typedef struct exb_options {
  int32_t device;      /* CUDA device ordinal; -1 = current device */
  int32_t rank;        /* shard of every pattern's iterator this handle evaluates ... */
  int32_t world;       /* ... out of `world` contiguous shards (1 = whole model) */
  int32_t flags;       /* EXB_FLAG_* */
  int64_t fuse_below;  /* reserved (every pattern of a callback shares one launch); pass 0 */
  const double* tune_x0; /* EXB_FLAG_TUNE_AT_CREATE: HOST pointer to nvar doubles to tune at (the model's x0); NULL = all ones */
} exb_options;

#define EXB_FLAG_NO_COMPILE 1 /* fail instead of invoking nvcc when the module is not cached */
#define EXB_FLAG_SORTED_PRODUCTS 2 /* exb_jprod / exb_jtprod / exb_hprod: the reference's scheme (COO values into scratch, then
                                     SpMV over a pre-sorted structure, ext:353-511; bitwise reproducible) instead of the default
                                     kernels fused into the derivative sweep (jtprod / hprod use FP64 atomic adds) */

#define EXB_FLAG_TUNE_AT_CREATE 4 /* rank the launch-shape variants of every kernel inside exb_create (on scratch outputs, at
                                     exb_options.tune_x0), so that no callback ever synchronises; default: each kernel is ranked at
                                     its first call, which synchronises the stream once.  Verdicts are remembered per (kernel,
                                     device, grid-size class) next to the cached modules either way. */

/* out[0..7] = nvar ncon nnzj nnzh nobj nnzg nconaug npar
 * (NLPModelMeta fields of src/nlp.jl:765-798 plus the scratch sizes of ext:21-31) */
#define EXB_NDIMS 8

/* ---- analysis (replaces src/simdfunction.jl:66-100 + the counters of nlp.jl) ---- */
int exb_plan_create(const void* ir, size_t ir_bytes, const exb_options* opt, exb_plan** out);
/* same with the iterator data at hand (as exb_create has it): integer fields that hold v0, v0 + 1, v0 + 2, ... (the `i` of an array
 * of NamedTuples built from 1:n) are recognised, which turns a data-indexed pattern into a shift-indexed one */
int exb_plan_create_data(const void* ir, size_t ir_bytes, const void* const* host_data, int n_data, const exb_options* opt, exb_plan** out);
int exb_plan_destroy(exb_plan* p);
int exb_plan_dims(const exb_plan* p, int64_t* out8);
int exb_plan_npatterns(const exb_plan* p);
/* out[0..8] = kind nitr o0 o1 o2 o1step o2step len(comp1) len(comp2) */
int exb_plan_pattern(const exb_plan* p, int k, int64_t* out9);
/* which = 1: comp1 (Compressor of the gradient/Jacobian pass), 2: comp2 (Hessian) */
int exb_plan_comp(const exb_plan* p, int k, int which, int64_t* out);
/* out[0] = 1 if the duplicate-free Hessian is emitted by the fused column-tile kernel (shift-indexed model), out[1] = its
 * number of unique entries (closed form), out[2] = distinct row - column distances, out[3] = halo points per tile */
int exb_plan_tile(const exb_plan* p, int64_t* out4);
/* generated CUDA C++ of the model's kernel module; *len excludes the NUL */
int exb_plan_source(const exb_plan* p, const char** src, size_t* len);
/* path of the compiled module for this plan (hash of source + flags) */
int exb_plan_module_path(const exb_plan* p, char* buf, size_t buflen);
/* run nvcc for this plan if the module is not cached (no GPU needed: cross-compiles) */
int exb_plan_compile(exb_plan* p);

/* ---- model lifetime (replaces build_extension, ext:33-191) ---------------- */
/* `host_data[k]` = host pointer to AoS data buffer k (the iterator arrays, element
 * stride as given in the IR); they are re-laid-out into device SoA columns and need
 * not outlive the call. */
int exb_create(const void* ir, size_t ir_bytes, const void* const* host_data, int n_data,
               const exb_options* opt, exb_model** out);
int exb_destroy(exb_model* m);
int exb_dims(const exb_model* m, int64_t* out8);
/* rank the launch-shape variants of every kernel now, at the DEVICE vectors x[nvar] / y[ncon] (NULL: all ones), on scratch
 * outputs; synchronises.  Collective on sharded handles with a communicator.  See EXB_FLAG_TUNE_AT_CREATE. */
int exb_tune(exb_model* m, const double* x, const double* y, void* stream);
/* out[0..4] = seconds spent in: planning + code generation, nvcc (0: module came from the cache), module load + data upload +
 * build-time sorts, tuning so far, exb_create in total */
int exb_build_info(const exb_model* m, double* out5);
/* theta: pointer to npar doubles, DEVICE or HOST (cudaMemcpyDefault; copied; set_value!, src/nlp.jl:1217-1287) */
int exb_set_params(exb_model* m, const double* theta, void* stream);

/* ---- callbacks: all pointers are DEVICE pointers, stream is a cudaStream_t ---- */
/* obj (src/nlp.jl:1827-1839 | ext:253-271): *out_host written after a stream sync */
int exb_obj(exb_model* m, const double* x, double* out_host, void* stream);
/* same, result left on the device (no sync): *out_dev */
int exb_obj_async(exb_model* m, const double* x, double* out_dev, void* stream);
/* grad! (src/nlp.jl:1858-1868 | ext:310-336): g[nvar] */
int exb_grad(exb_model* m, const double* x, double* g, void* stream);
/* cons_nln! (src/nlp.jl:1841-1854 | ext:273-308): c[ncon] */
int exb_cons(exb_model* m, const double* x, double* c, void* stream);
/* jac_structure! (src/nlp.jl:1798-1807 | ext:212-226) */
int exb_jac_structure64(exb_model* m, int64_t* rows, int64_t* cols, void* stream);
int exb_jac_structure32(exb_model* m, int32_t* rows, int32_t* cols, void* stream);
/* jac_coord! (src/nlp.jl:1870-1880 | ext:338-351): vals[nnzj] */
int exb_jac(exb_model* m, const double* x, double* vals, void* stream);
/* hess_structure! (src/nlp.jl:1809-1825 | ext:229-250): lower triangle */
int exb_hess_structure64(exb_model* m, int64_t* rows, int64_t* cols, void* stream);
int exb_hess_structure32(exb_model* m, int32_t* rows, int32_t* cols, void* stream);
/* hess_coord! (src/nlp.jl:1906-1940 | ext:515-547): vals[nnzh];
 * y == NULL is the objective-only form (src/nlp.jl:1906-1915) */
int exb_hess(exb_model* m, const double* x, const double* y, double obj_weight, double* vals,
             void* stream);

/* ---- several callbacks from ONE sweep (the composition of src/nlp.jl:1827-1940: a solver asks for obj, grad!, cons!, jac_coord!
 * and hess_coord! at the same x, and the reference walks every pattern's tree once per callback).  mask = EXB_EVAL_ALL: every
 * data point is evaluated once by one generated launch (exb_eval_g0) that writes c, jac, hess, the gradient slots and the
 * objective's partial sums; the small finishing steps of the separate callbacks follow (fixed-order sum, owner-computed
 * gradient, segmented sums, collectives of a sharded model).  On an unsharded handle whose model has ONE objective pattern with
 * gradient slots, shift-indexed, the sweep writes g itself (no gradient launch; same summation order as exb_grad).
 * mask = EXB_EVAL_FIRST / EXB_EVAL_VALUES: the same with a first-order /
 * value-only sweep.  Any other mask: the requested callbacks one by one.  Outputs not
 * requested may be NULL; *obj_dev is a DEVICE double (no synchronisation); y == NULL is the objective-only Hessian form. */
#define EXB_EVAL_OBJ 1u
#define EXB_EVAL_GRAD 2u
#define EXB_EVAL_CONS 4u
#define EXB_EVAL_JAC 8u
#define EXB_EVAL_HESS 16u
#define EXB_EVAL_ALL 31u
#define EXB_EVAL_FIRST 15u  /* obj | grad | cons | jac: ONE first-order sweep (exb_eval1_g0) */
#define EXB_EVAL_VALUES 5u  /* obj | cons: ONE value sweep (exb_eval0_g0), e.g. a line-search probe */
int exb_eval(exb_model* m, unsigned mask, const double* x, const double* y, double obj_weight, double* obj_dev, double* g, double* c,
             double* jac, double* hess, void* stream);

/* ---- matrix-free products: jprod_nln! / jtprod_nln! / hprod! (src/nlp.jl:1882-1978; device form
 * ext/ExaModelsKernelAbstractions.jl:353-511, `ExaModel(c; prod = true)`).  Default: fused into the derivative sweep --
 * every point multiplies its first- / second-order slots (still in registers) with v; jprod assigns / segment-sums rows
 * (deterministic), jtprod and hprod add into the pre-zeroed output with FP64 atomics; nothing of size nnzj / nnzh is
 * written or read and sharded handles return their partial sums.  With EXB_FLAG_SORTED_PRODUCTS: the reference's scheme,
 * COO values into a buffer owned by the handle, multiplied through row- / column-sorted copies of the structure (built on
 * first use, ext:56-175); bitwise reproducible; a sharded handle sorts the slots of its own points and returns the shard's partial
 * product (completed by the all-reduce when a communicator is attached). */
int exb_jprod(exb_model* m, const double* x, const double* v, double* Jv, void* stream);    /* Jv[ncon]  */
int exb_jtprod(exb_model* m, const double* x, const double* v, double* Jtv, void* stream);  /* Jtv[nvar] */
int exb_hprod(exb_model* m, const double* x, const double* y, const double* v, double obj_weight, double* Hv,
              void* stream);                                                                 /* Hv[nvar]  */

/* ---- duplicate-free COO: the CompressedNLPModel role (src/utils.jl:425-579 | ext:1290-1319) ----
 * Unique (row, col) coordinates in the reference's order (sorted by (col, row)); values of duplicates are summed in
 * ascending slot order (the order `_compress!`, utils.jl:564-571, meets them in the stably sorted list).
 *   Hessian of a SHIFT-INDEXED model (every variable of every pattern with second-order slots is `range value + const`: LV,
 *   any banded / stencil model): FUSED -- one generated launch (exb_hessc_g0, csrc/exb_device.cuh exb_tile_body) evaluates the
 *   points around a tile of columns, sums the slots that share a coordinate in shared memory / registers and writes each
 *   unique entry once (LV: 2N - 1 doubles instead of 9N - 15, and no second pass).  Structure and count are closed forms of
 *   the pattern shifts: no sort, no nnzh-sized buffer.  Works on sharded handles: a rank owns a contiguous range of columns
 *   (exb_owned), hence a contiguous range of the unique entries (exb_compressed_shard); duplicates that straddle two shards
 *   need no exchange because the owner evaluates both contributing points (x, y are replicated).
 *   Anything else (Jacobian; models indexed through iterator data or with fixed-index variables): the reference's scheme --
 *   raw COO values into a buffer owned by the handle, then a segmented sum through the pre-sorted list, built on first use;
 *   not available on sharded handles. */
int exb_compressed_dims(exb_model* m, int64_t* nnzj_unique, int64_t* nnzh_unique);   /* either pointer may be NULL */
/* out[0..1] = 0-based half-open range of the duplicate-free Hessian values this handle writes, out[2] = 1 if fused */
int exb_compressed_shard(exb_model* m, int64_t* out3);
int exb_jac_structure_compressed64(exb_model* m, int64_t* rows, int64_t* cols, void* stream);
int exb_jac_structure_compressed32(exb_model* m, int32_t* rows, int32_t* cols, void* stream);
int exb_hess_structure_compressed64(exb_model* m, int64_t* rows, int64_t* cols, void* stream);
int exb_hess_structure_compressed32(exb_model* m, int32_t* rows, int32_t* cols, void* stream);
int exb_jac_compressed(exb_model* m, const double* x, double* vals, void* stream);
int exb_hess_compressed(exb_model* m, const double* x, const double* y, double obj_weight, double* vals,
                        void* stream);

/* ---- host-buffer shims: the WrapperNLPModel role (src/utils.jl:16-267) ------
 * Same callbacks with HOST pointers; H2D / D2H copies go through pinned staging owned
 * by the handle and are part of the call.  They synchronise before returning. */
int exb_host_obj(exb_model* m, const double* x, double* out);
int exb_host_grad(exb_model* m, const double* x, double* g);
int exb_host_cons(exb_model* m, const double* x, double* c);
int exb_host_jac(exb_model* m, const double* x, double* vals);
int exb_host_hess(exb_model* m, const double* x, const double* y, double obj_weight, double* vals);
/* duplicate-free forms: the D2H copy carries the unique entries only */
int exb_host_jac_compressed(exb_model* m, const double* x, double* vals);
int exb_host_hess_compressed(exb_model* m, const double* x, const double* y, double obj_weight, double* vals);
int exb_host_jprod(exb_model* m, const double* x, const double* v, double* Jv);
int exb_host_jtprod(exb_model* m, const double* x, const double* v, double* Jtv);
int exb_host_hprod(exb_model* m, const double* x, const double* y, const double* v, double obj_weight, double* Hv);
int exb_host_jac_structure32(exb_model* m, int32_t* rows, int32_t* cols);   /* the AOT path passes Cint (ExaModelsCompiler.jl:1650-1668) */
int exb_host_hess_structure32(exb_model* m, int32_t* rows, int32_t* cols);
int exb_host_jac_structure64(exb_model* m, int64_t* rows, int64_t* cols);
int exb_host_hess_structure64(exb_model* m, int64_t* rows, int64_t* cols);

/* ---- multi-GPU: one handle per GPU, each evaluating shard `rank` of `world` (exb_options) ---------------------------
 * The reference has a single device (ext/ExaModelsKernelAbstractions.jl:212-547); this is the path BASELINE.json's north
 * star adds: "NCCL over NVLink only to allreduce the scalar objective/gradient overlaps and to sum duplicated COO entries".
 * A communicator over the `world` handles is either created here from an ncclUniqueId that the host distributes by any
 * means it has (MPI, a socket, torch.distributed): rank 0 calls exb_comm_unique_id, every rank calls exb_comm_init with the
 * same 128 bytes (collective, like ncclCommInitRank) -- or borrowed from the host (exb_comm_attach, an ncclComm_t whose rank /
 * size equal the handle's).  NCCL is dlopen'ed (libnccl.so.2; EXB_NCCL_LIB overrides), there is no link-time dependency.
 *
 * With a communicator the reducing callbacks complete themselves on the caller's stream, with no host synchronisation:
 *   exb_obj / exb_obj_async   partial sums -> ncclAllReduce of ONE double (8 bytes)
 *   exb_grad                  variables are owned in contiguous ranges [nvar r / W, nvar (r + 1) / W) (exb_owned).  Objective
 *                             patterns with shift-indexed variables are owner-computed PER VARIABLE (x is replicated), so a
 *                             rank's g is exact on the range it owns and NOTHING is exchanged -- not even the halo; REPLICATE
 *                             mode then all-gathers the ranges.  Models with other objective patterns (slots scattered by
 *                             iterator data) fall back to ncclAllReduce over nvar.
 *   exb_cons / exb_jprod      base rows are owned with their points (exb_shard); REPLICATE mode all-gathers them (one NCCL
 *                             group of broadcasts).  Augmentation terms land in arbitrary rows -> ncclAllReduce over ncon.
 *   exb_jtprod / exb_hprod    partial products -> ncclAllReduce over nvar
 *   exb_jac / exb_hess        every rank writes contiguous, disjoint slices per pattern (exb_shard): no collective;
 *                             exb_comm_gather_coo replicates them when the consumer is not sharded.
 * Summation order across ranks differs from the single-GPU order: compare at 1e-10, not bitwise.  Without a communicator a
 * sharded handle returns its partial results (zero outside what it computes) and the host reduces them itself. */
#define EXB_COMM_ID_BYTES 128
#define EXB_COMM_REPLICATE 0 /* default: obj, g, c, Jv, Jtv, Hv are complete on every rank after the call */
#define EXB_COMM_OWNER 1     /* sharded consumer: g is valid on the owned variables, c / Jv on the rows of the own points
                                (at least; results that had to be all-reduced are complete everywhere) */
int exb_comm_unique_id(void* id128);
int exb_comm_init(exb_model* m, const void* id128);
int exb_comm_attach(exb_model* m, void* nccl_comm);
int exb_comm_destroy(exb_model* m);
int exb_comm_set_mode(exb_model* m, int mode);
/* which = 1: jac, 2: hess -- in-place replication of the sharded COO values */
int exb_comm_gather_coo(exb_model* m, int which, double* vals, void* stream);
/* out[0..1] = 0-based half-open range of the variables this handle owns */
int exb_owned(const exb_model* m, int64_t* out2);
/* out[0] = collectives issued since creation, out[1] = by the last callback, out[2] = 1 if a communicator is attached, out[3] = mode */
int exb_comm_stats(const exb_model* m, int64_t* out4);

/* ---- sharding / introspection ------------------------------------------- */
/* For pattern k: out[0..5] = lo hi (0-based local point range of this rank)
 * jac_lo jac_hi hess_lo hess_hi (0-based half-open slices of vals this rank writes) */
int exb_shard(const exb_model* m, int k, int64_t* out6);
/* out[0] = kernels launched since creation, out[1] = last callback launches,
 * out[2] = device bytes owned by the handle, out[3] = 1 if module came from cache */
int exb_stats(const exb_model* m, int64_t* out4);
/* Which generated kernel a value callback launches (callback: 0 obj, 1 grad, 2 cons, 3 jac, 4 hess), as decided by the
 * first-call tuner: out[0] = min-blocks-per-SM of the launch-shape variant (-1: not tuned yet), out[1] = 1 if the
 * persistent shared-memory-window form is in use (hess), out[2] = grid size, out[3] = 1 if grad uses the owner-computes kernel */
int exb_kernel_choice(const exb_model* m, int callback, int64_t* out4);
/* out[0] / out[1] = bytes copied host -> device / device -> host by the last exb_host_* call on this handle (a sharded
 * handle uploads only the part of x its points can read and its own rows of y, and downloads only the slices it wrote) */
int exb_host_bytes(const exb_model* m, int64_t* out2);

/* Per-callback device timing: the TimedNLPModel role (src/utils.jl:271-408).  When on, every value callback is
 * bracketed by CUDA events on the caller's stream (no synchronisation); exb_timings synchronises on them and returns
 * the accumulated milliseconds and call counts for obj grad cons jac hess jprod jtprod hprod (8 entries each). */
int exb_set_timing(exb_model* m, int on);
int exb_timings(exb_model* m, double* ms8, int64_t* calls8, int reset);

const char* exb_last_error(void);
int exb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EXA_B200_H */

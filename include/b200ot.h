/*
 * b200ot.h -- C ABI of the B200-native entropic optimal-transport engine.
 *
 * This is the drop-in boundary for the OT hot path of the reference
 * (SURVEY.md section 8b).  The reference has no FFI of its own: the path is a
 * chain of Python calls into POT / ott-jax / perturbot.  Each entry point below
 * names the reference call it replaces (file:line relative to /root/reference).
 * The Python host (b200ot/) binds these symbols with ctypes and re-exposes them
 * under the reference's own function names.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer, including the workspace; the library never
 *     allocates or frees device memory;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 *     except b200ot_sinkhorn_solve, which polls a pinned flag between chunks;
 *   - return value: 0 on success, <0 = B200OT_E_* (never a C++ exception);
 *   - matrices are row-major fp32 with an explicit leading dimension (elements).
 *     The single-sweep Sinkhorn path needs ldc % 4 == 0 and 16-byte aligned
 *     bases; other shapes are served by the generic kernels.
 *   - potentials cross the ABI in natural units (f, g with P = exp((f+g-C)/eps)).
 */
#ifndef B200OT_H_
#define B200OT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200OT_VERSION 100

/* status codes */
#define B200OT_OK 0
#define B200OT_E_INVALID (-1)   /* bad argument (null pointer, non-positive size, misalignment) */
#define B200OT_E_WORKSPACE (-2) /* workspace too small */
#define B200OT_E_LAUNCH (-3)    /* CUDA launch / runtime failure; see b200ot_last_cuda_error */
#define B200OT_E_UNSUPPORTED (-4)
#define B200OT_E_NUMERIC (-5)   /* non-finite value met by the fast path (caller may retry robust) */

/* error norm of the column-marginal violation */
#define B200OT_NORM_L2 0   /* POT 0.9.6 sinkhorn_knopp (MRI_PET_OT_nojax.py:143)             */
#define B200OT_NORM_L2SQ 1 /* in-tree mirror perturbot/perturbot/match/utils.py:88-89        */
#define B200OT_NORM_L1 2   /* ott-jax 0.6.0 Sinkhorn (perturbot/perturbot/match/fot.py:129)   */

/* which kernels run one Sinkhorn iteration */
#define B200OT_PATH_AUTO 0
#define B200OT_PATH_FUSED 1  /* one sweep of C per iteration (cluster kernel, TMA bulk ring)   */
#define B200OT_PATH_ROBUST 2 /* two sweeps, running-max logsumexp in both directions          */

/* cost kinds */
#define B200OT_COST_SQEUCLIDEAN 0
#define B200OT_COST_COSINE 1

typedef struct b200ot_params {
  float eps;           /* entropic regularisation, absolute on the cost passed in            */
  int max_iter;        /* numItermax / max_iterations                                        */
  float tol;           /* stopThr / threshold                                                */
  int check_every;     /* 10 in POT, the mirror and ott                                      */
  int check_phase;     /* error checked after 1-based iteration it when                      */
                       /*   it % check_every == check_phase % check_every                    */
                       /*   (1 = POT / mirror `cpt % 10 == 0`, 0 = ott inner_iterations)     */
  int err_norm;        /* B200OT_NORM_*                                                      */
  int stop_inclusive;  /* 1: stop when err <= tol (mirror); 0: err < tol (POT, ott)          */
  int path;            /* B200OT_PATH_*                                                      */
  int floor_patience;  /* 0 = the reference rule only.  k > 0: also stop once the error has set no   */
                       /* new minimum for k consecutive checks while already below 1e-4 of |b| --    */
                       /* fp32 cannot resolve POT's default stopThr = 1e-9 on small problems, and    */
                       /* without this the solve would spin to numItermax (result status 1).         */
} b200ot_params;

/* result block written by b200ot_sinkhorn_finish (device memory, 32 bytes) */
typedef struct b200ot_result {
  int n_iter;     /* completed (g,f) updates                                                 */
  int converged;  /* 1 if the stopping rule fired                                            */
  int status;     /* 0 ok, 1 stopped at the fp32 resolution floor (floor_patience),          */
                  /* B200OT_E_NUMERIC if the fast path met a non-finite / vanished sum       */
  int n_err;      /* number of recorded error checks                                         */
  float err;      /* last evaluated marginal error                                           */
  float reserved[3];
} b200ot_result;

int b200ot_version(void);
const char* b200ot_strerror(int code);
/* last CUDA runtime error string seen by this thread's failing call (host pointer) */
const char* b200ot_last_cuda_error(void);

/* ---- cost construction ---------------------------------------------------
 * C_ij = |x_i|^2 + |y_j|^2 - 2 x_i.y_j   (kind = SQEUCLIDEAN) or 1 - cos(x_i,y_j).
 * Replaces ot.dist(X, Y) (MRI_PET_OT_nojax.py:70-71) and is the T = I case of the
 * feature cost M = t1 (+) t2 - 2 X^T Ts Y (MRI_PET_OT_nojax.py:121-136,
 * perturbot/perturbot/match/utils.py:125-184).  X is n x d (ldx), Y is m x d (ldy).
 * b200ot_cost runs the contraction on tcgen05 (bf16 tensor cores, fp32 accumulators in TMEM)
 * with an fp32-accurate split over bf16 parts x = x1 + x2 + x3: terms = 6 evaluates
 * x1.y3 + x3.y1 + x2.y2 + x1.y2 + x2.y1 + x1.y1 (fp32-grade, the default of the Python host),
 * terms = 3 drops the first three (error ~2^-17 per product), terms = 1 is the plain bf16 product.
 * terms = B200OT_TERMS_F16_3 (the Python host's default) scales every row by a power of two (largest entry into
 * [2^9, 2^10)), splits it into TWO fp16 parts (11 + 11 bits) and evaluates x1.y2 + x2.y1 + x1.y1: representation
 * error <= 2^-23 of the row maximum per entry -- fp32-grade relative to |x||y| -- at half the tensor work of
 * terms = 6; B200OT_TERMS_F16_4 adds x2.y2.
 * `ws` (1024-byte aligned) needs b200ot_cost_workspace_bytes(n, m, d) bytes.
 * cost_simt is the fp32 FMA version; `norms` is scratch for n + m floats.          */
#define B200OT_TERMS_F16_3 19
#define B200OT_TERMS_F16_4 20
size_t b200ot_cost_workspace_bytes(int n, int m, int d);
int b200ot_cost(const float* X, int ldx, const float* Y, int ldy, int n, int m, int d, int kind,
                float* C, int ldc, void* ws, size_t ws_bytes, int terms, void* stream);
/* The two halves of b200ot_cost for callers that keep the bf16 parts resident (the online solver re-builds
 * row panels of C every iteration but splits X and Y once).  side = 0: rows of X (128-row tiles), 1: rows of Y
 * (256-row tiles).  parts needs b200ot_cost_parts_bytes bytes (1024-byte aligned), norms rows (padded to the
 * tile) floats -- TWICE that with the fp16 terms, which store {norm, 2^-s} per row.  cost_gemm builds n rows of C starting at 128-row tile `row_tile0` of the X parts.        */
size_t b200ot_cost_parts_bytes(int rows, int d, int side);
int b200ot_cost_split(const float* X, int ldx, int rows, int d, int kind, int terms, int side, void* parts,
                      float* norms, void* stream);
int b200ot_cost_gemm(const void* partsA, const float* normsA, int row_tile0, int n, const void* partsB,
                     const float* normsB, int m, int d, int kind, int terms, float* C, int ldc, void* stream);
int b200ot_cost_simt(const float* X, int ldx, const float* Y, int ldy, int n, int m, int d,
                     int kind, float* C, int ldc, float* norms, void* stream);

/* FOT feature cost M = (A.^2)^T w1 (+) (B.^2)^T w2 - 2 A^T Ts B with A n x d, B n2 x d2,
 * Ts n x n2 (MRI_PET_OT_nojax.py:121-136; perturbot/perturbot/match/fot.py:118-128).
 * tmp must hold n*d2 + d + d2 floats.  w1 (n), w2 (n2) are the marginals the caller
 * chose (the two reference variants sum Ts over different axes).                  */
int b200ot_fot_cost(const float* A, int lda, const float* B, int ldb, const float* Ts, int ldt,
                    const float* w1, const float* w2, int n, int n2, int d, int d2, float* M,
                    int ldm, float* tmp, void* stream);

/* The same feature cost with both contractions on tcgen05 (the split / GEMM pair of b200ot_cost, 6-term bf16 split,
 * fp32-grade): G = -2 B^T Ts^T, then M = t1 (+) t2 - 2 A^T (-G/2)^T.  SURVEY 8(f-1): the device-resident caller's
 * X^T.Ts.Y chain.  `ws` (1024-byte aligned) needs b200ot_fot_cost_tc_workspace_bytes bytes.            */
size_t b200ot_fot_cost_tc_workspace_bytes(int n, int n2, int d, int d2);
int b200ot_fot_cost_tc(const float* A, int lda, const float* B, int ldb, const float* Ts, int ldt,
                       const float* w1, const float* w2, int n, int n2, int d, int d2, float* M,
                       int ldm, void* ws, size_t ws_bytes, void* stream);

/* max of a matrix (ott Geometry(scale_cost="max_cost"), fot.py:129-133) and in-place scale */
int b200ot_matrix_max(const float* C, int ldc, int n, int m, float* out_max, void* stream);
int b200ot_matrix_scale_by_inv(float* C, int ldc, int n, int m, const float* denom, void* stream);

/* ---- Sinkhorn (log domain, fp32) -----------------------------------------
 * Replaces ot.sinkhorn(a, b, M, reg, numItermax, stopThr) (MRI_PET_OT_nojax.py:143),
 * sinkhorn_scaling (perturbot/perturbot/match/utils.py:6-115) and
 * ott linear.solve(Geometry(cost_matrix=M, epsilon, scale_cost)).matrix
 * (perturbot/perturbot/match/fot.py:129-134).  One iteration = g update then f update.
 *
 * Life cycle:  init -> enqueue (any number of times) -> finish.
 *   init     stores a, b, the scaled start potentials and the stopping rule in ws and
 *            runs the first g update (f0/g0 may be NULL = zero potentials; POT's
 *            u=1/n start is f0 = eps*log(1/n)).  setup is init without the first g
 *            update (row-sharded callers run it through shard_prologue/finalize).
 *   enqueue  snapshots the potentials, then queues `iters` iterations on `path`
 *            (B200OT_PATH_*); kernels no-op once the stopping rule has fired.
 *   rewind   restores the snapshot of the last enqueue (used to replay a chunk whose
 *            fast path lost a sum on the ROBUST path).
 *   peek     copies 8 ints {it, done, converged, cur, bad, n_err, -, -} to flags8
 *            (device or pinned host memory).
 *   finish   writes f, g (natural units) and the result block (device memory).     */
size_t b200ot_sinkhorn_workspace_bytes(int n, int m);
int b200ot_sinkhorn_setup(int n, int m, const float* a, const float* b, const float* f0,
                          const float* g0, const b200ot_params* prm, void* ws, size_t ws_bytes,
                          void* stream);
int b200ot_sinkhorn_init(const float* C, int ldc, int n, int m, const float* a, const float* b,
                         const float* f0, const float* g0, const b200ot_params* prm, void* ws,
                         size_t ws_bytes, void* stream);
int b200ot_sinkhorn_enqueue(const float* C, int ldc, int n, int m, int iters, int path, void* ws,
                            void* stream);
int b200ot_sinkhorn_snapshot(int n, int m, void* ws, void* stream);
int b200ot_sinkhorn_rewind(int n, int m, void* ws, void* stream);
int b200ot_sinkhorn_peek(void* ws, int* flags8, void* stream);
/* host-side counters of this process: which = 0 iterations launched in the fused form (sweep + fold + exchange +
 * finalize + state machine in ONE cooperative cluster launch), 1 = fused launches that were refused and fell back
 * to separate launches (the fused form is then disabled for the process; b200ot_last_cuda_error says why).    */
long long b200ot_sinkhorn_counter(int which);
/* human-readable description of the kernel configuration chosen for an n x m problem (host buffer) */
int b200ot_sinkhorn_describe(int n, int m, char* buf_host, int buf_len);
int b200ot_sinkhorn_finish(int n, int m, void* ws, float* f, float* g, b200ot_result* result,
                           float* err_hist, int err_hist_cap, void* stream);
/* blocking convenience driver: init, chunks ending on check iterations with the flags of
 * chunk c-1 read while chunk c runs, automatic ROBUST replay of a chunk that hit
 * E_NUMERIC, finish.  result_host is a HOST pointer (may be NULL).                  */
int b200ot_sinkhorn_solve(const float* C, int ldc, int n, int m, const float* a, const float* b,
                          const float* f0, const float* g0, const b200ot_params* prm, void* ws,
                          size_t ws_bytes, float* f, float* g, b200ot_result* result_host,
                          float* err_hist, int err_hist_cap, void* stream);

/* Row-sharded form: this rank owns n_local rows of C (and of a, f); b and g
 * are replicated.  After setup, shard_prologue leaves this rank's column sums for the first
 * g update in s_local (m floats); shard_sweep does one f update on the local rows and
 * leaves the local column sums of the new plan.  The caller all-reduces s_local (NCCL sum)
 * and hands the total to shard_finalize, which every rank runs identically (marginal
 * error, stopping rule, next g).                                                    */
int b200ot_sinkhorn_shard_prologue(const float* C, int ldc, int n_local, int m, void* ws,
                                   float* s_local, void* stream);
int b200ot_sinkhorn_shard_sweep(const float* C, int ldc, int n_local, int m, int path, void* ws,
                                float* s_local, void* stream);
int b200ot_sinkhorn_shard_finalize(int n_local, int m, void* ws, const float* s_total,
                                   int is_prologue, void* stream);

/* Row panels of ONE problem (online / C-free solver): ws describes the whole n x m problem, Cpanel holds rows
 * [row0, row0 + rows) of the cost (rebuilt on the fly, L2-resident); the sweep updates f for those rows and adds
 * the panel's column sums into s_accum (accumulate = 0 on the first panel of an iteration).  After the last
 * panel: b200ot_sinkhorn_shard_finalize(n, m, ws, s_accum, is_prologue).                                */
int b200ot_sinkhorn_panel_prologue(const float* Cpanel, int ldc, int n, int m, int row0, int rows, void* ws,
                                   float* s_accum, int accumulate, void* stream);
int b200ot_sinkhorn_panel_sweep(const float* Cpanel, int ldc, int n, int m, int row0, int rows, int path,
                                void* ws, float* s_accum, int accumulate, void* stream);

/* Same loop with the all-reduce issued from C on the compute stream (NCCL is bound at run time from the
 * libnccl.so.2 already loaded in the process; no link-time dependency).  unique_id: rank 0 creates the
 * 128-byte id, the host broadcasts it, every rank calls nccl_init (collective).  shard_start = first g
 * update (column sums -> all-reduce -> finalize); shard_run = `iters` x (single-sweep -> fold partials ->
 * ncclAllReduce of m floats -> finalize), all on `stream`, CUDA-graph capturable.  comm == NULL runs the
 * loop without the collective.  s_buf: m floats of device scratch.                              */
int b200ot_nccl_unique_id(unsigned char* id128_host);
int b200ot_nccl_init(const unsigned char* id128_host, int world, int rank, void** comm_out);
int b200ot_nccl_destroy(void* comm);
int b200ot_sinkhorn_shard_start(const float* C, int ldc, int n_local, int m, void* ws, float* s_buf,
                                void* comm, void* stream);
int b200ot_sinkhorn_shard_run(const float* C, int ldc, int n_local, int m, int iters, int path, void* ws,
                              float* s_buf, void* comm, void* stream);

/* The same row-sharded loop WITHOUT a collective library call in it (the all-reduce of m column
 * partials fused into the kernels that produce and consume them).  Every rank owns an exchange buffer of
 * b200ot_peer_exchange_bytes(world, m) bytes that its peers map with CUDA IPC (peer_alloc on the owner -> 64-byte
 * handle -> peer_open on the others; one box, NVLink / NVSwitch).  shard_push runs the local sweep (or, with
 * is_prologue, the column pass of the first g update), folds the cluster partials and STORES the m column sums as
 * tagged 64-bit words {value, tag} straight into slab [parity][rank] of every peer's buffer; shard_finalize_peer
 * polls the world slabs of this rank's own buffer, adds them in rank order (identical result on every rank, no
 * atomics) and applies the usual finalize (marginal error, stopping rule, next g).  `epoch` distinguishes solves
 * that reuse the buffers (same value on every rank, change it per solve); the iteration part of the tag comes from
 * the device-side state, so shard_run_peer can queue any number of iterations ahead and stays graph-capturable.
 * peer_bufs: HOST array of `world` device pointers, entry r = rank r's buffer as mapped in this process.        */
int b200ot_peer_alloc(size_t bytes, void** dptr, unsigned char* handle64_host);
int b200ot_peer_open(const unsigned char* handle64_host, void** dptr);
int b200ot_peer_close(void* dptr);
int b200ot_peer_free(void* dptr);
size_t b200ot_peer_exchange_bytes(int world, int m);
int b200ot_sinkhorn_shard_push(const float* C, int ldc, int n_local, int m, int path, void* ws,
                               void* const* peer_bufs, int world, int rank, unsigned epoch, int is_prologue,
                               void* stream);
int b200ot_sinkhorn_shard_finalize_peer(int n_local, int m, void* ws, const void* my_buf, int world, unsigned epoch,
                                        int is_prologue, void* stream);
int b200ot_sinkhorn_shard_run_peer(const float* C, int ldc, int n_local, int m, int iters, int path, void* ws,
                                   void* const* peer_bufs, int world, int rank, unsigned epoch, void* stream);

/* ---- batched small problems (one problem per CTA, float64, kernel domain) --
 * BASELINE config 2: per-training-step minibatch OT (MRI_PET_OT_nojax.py:679-715).
 * Runs POT's sinkhorn_knopp arithmetic itself (u = 1/n start, v then u update, error every
 * check_every iterations, previous-iterate restore on 0/NaN/Inf) in float64 with K resident
 * in shared memory, so iteration counts equal the reference's.  C is batch x n x m
 * contiguous fp32 (or NULL with X, Y given: batch x n x d and batch x m x d embeddings,
 * cost built in-kernel in float64).  Outputs: P (batch x n x m fp32), u, v (float64),
 * n_iter/err per problem.  n, m <= 128.                                            */
int b200ot_sinkhorn_batched(const float* C, const float* X, const float* Y, int batch, int n,
                            int m, int d, const float* a, const float* b,
                            const b200ot_params* prm, float* P, double* u, double* v, int* n_iter,
                            float* err, void* stream);

/* ---- epilogues -------------------------------------------------------------
 * plan       P_ij = exp((f_i + g_j - C_ij)/eps)   (POT's diag(u) K diag(v), utils.py:111-115;
 *            ott's .matrix).
 * ot_cost    <P, C> without materialising P (fot.py:137); out is one double.
 * apply      Z = P V (normalise=0) or diag(1/rowsum(P)) P V (normalise=1), V m x dv,
 *            without materialising P: the barycentric projection of
 *            perturbot/perturbot/eval/match.py:202-206 and, with V = pet^T, the
 *            reference's `pet @ T.t()` (MRI_PET_OT_OT_per_epoch_attn.py:728).
 * apply_t    Z = P^T U (U n x du, Z m x du), same options.
 * These four are the generic SIMT kernels (any shape and alignment); large problems use the tensor-core forms
 * b200ot_apply_plan_tc / b200ot_envelope_bwd declared below.                                     */
int b200ot_plan(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                float* P, int ldp, void* stream);
/* ot_cost: `out` must hold B200OT_OT_COST_DOUBLES doubles; out[0] is <P, C>, the rest is scratch for the
 * fixed-order two-stage fold (the value is bit-reproducible: no floating-point atomics).        */
#define B200OT_OT_COST_DOUBLES (1 + 148 * 8 + 1)
int b200ot_ot_cost(const float* C, int ldc, int n, int m, const float* f, const float* g,
                   float eps, double* out, void* stream);
/* Per-step plan guard of the reference (MRI_PET_OT_nojax.py:704-715): NaN -> 1e-8, then every row divided by its
 * sum, zero sums replaced by 1e-8.  T == NULL: the plan of (C, f, g, eps) is evaluated on the fly and written
 * normalised; T != NULL (n x m, ldt): an already materialised plan is sanitised (C, f, g ignored).   */
int b200ot_plan_guard_rownorm(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                              const float* T, int ldt, float* P, int ldp, void* stream);
int b200ot_apply_plan(const float* C, int ldc, int n, int m, const float* f, const float* g,
                      float eps, const float* V, int ldv, int dv, int normalise, float* Z,
                      int ldz, void* stream);
int b200ot_apply_plan_t(const float* C, int ldc, int n, int m, const float* f, const float* g,
                        float eps, const float* U, int ldu, int du, int normalise, float* Z,
                        int ldz, void* stream);
/* Tensor-core form of apply / apply_t (csrc/apply_tc.cu): ONE pass over C whatever dv <= 512 is.  The plan entries
 * are evaluated once, split into two bf16 parts and fed to tcgen05.mma as the A operand against the pre-split
 * right-hand side (3-term product p1.v2 + p2.v1 + p1.v1, fp32 accumulators in TMEM, error ~2^-16 per product);
 * row sums of P come out of the same registers in fp32.  transpose = 0: Z (n x dv) = P V, V is m x dv;
 * transpose = 1: Z (m x dv) = P^T V, V is n x dv.  normalise = 1 divides every row of Z by the row sum of P
 * (0 -> 1e-30, perturbot/perturbot/eval/match.py:203-204).  rowsum_out (may be NULL) receives those row sums
 * (P 1, or P^T 1 in the transposed form).  ws: 1024-byte aligned, b200ot_apply_plan_tc_workspace_bytes bytes
 * (the pre-tiled bf16 parts of V and, for short outputs whose K range is split over several CTAs, the partial
 * accumulators).                                                                                          */
size_t b200ot_apply_plan_tc_workspace_bytes(int n, int m, int dv, int transpose);
int b200ot_apply_plan_tc(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                         const float* V, int ldv, int dv, int transpose, int normalise, float* Z, int ldz,
                         float* rowsum_out, void* ws, size_t ws_bytes, void* stream);
/* Envelope gradient of <P, C(X, Y)> for the squared-Euclidean cost at fixed P (new capability; the reference never
 * differentiates through OT, MRI_PET_OT_nojax.py:683-684), both halves in ONE launch of the same kernel:
 *   dX = scale (diag(P 1) X - P Y)   (n x d),   dY = scale (diag(P^T 1) Y - P^T X)   (m x d),  scale = 2 for <P, C>.
 * rowsum_out (n) / colsum_out (m) may be NULL.                                                              */
size_t b200ot_envelope_bwd_workspace_bytes(int n, int m, int d);
int b200ot_envelope_bwd(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                        const float* X, int ldx, const float* Y, int ldy, int d, float scale, float* dX, int lddx,
                        float* dY, int lddy, float* rowsum_out, float* colsum_out, void* ws, size_t ws_bytes,
                        void* stream);
/* Weighted forms of the two calls above: the contracted entries are W_ij = P_ij (w0 + w1 C_ij + wrow_i + wcol_j)
 * (wrow: n values over the rows of C, wcol: m values over its columns, either may be NULL) instead of P_ij, and
 * the row sums are those of W.  (w0, w1) = (0, 1) gives the C-weighted products (P o C) V; with
 * w0 = 1, w1 = -1/eps, wrow = lambda/eps, wcol = mu/eps, where H (lambda, mu) = ((P o C) 1, (P o C)^T 1) and
 * H = [[diag(P1), P], [P^T, diag(P^T 1)]], W is d<P, C>/dC THROUGH the Sinkhorn fixed point, so
 * envelope_bwd_weighted returns the implicit-function-theorem gradient of the transport cost with respect to the
 * embeddings (b200ot.torch_ops.ot_loss(value="primal", grad="implicit")).  Workspaces as for the unweighted calls. */
int b200ot_apply_plan_tc_weighted(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                                  float w0, float w1, const float* wrow, const float* wcol, const float* V, int ldv,
                                  int dv, int transpose, float* Z, int ldz, float* rowsum_out, void* ws,
                                  size_t ws_bytes, void* stream);
int b200ot_envelope_bwd_weighted(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                                 float w0, float w1, const float* wrow, const float* wcol, const float* X, int ldx,
                                 const float* Y, int ldy, int d, float scale, float* dX, int lddx, float* dY, int lddy,
                                 void* ws, size_t ws_bytes, void* stream);
/* fused fusion loss of the reference forward: 1 - mean_i cos(A_i, B_i)
 * (MRI_PET_OT_nojax.py:552-560); out is one float.                                  */
int b200ot_cosine_loss(const float* A, int lda, const float* B, int ldb, int rows, int d,
                       float* out, void* stream);

/* FOSCTTM plan-quality metric (perturbot/perturbot/eval/utils.py:18-45): D is the n x n matrix of squared
 * distances between predictions (rows) and true targets (columns), e.g. from b200ot_cost; out[i] = (rank of
 * D[i][i] within row i, ties at their mean position) / (n - 1).                                     */
int b200ot_foscttm(const float* D, int ldd, int n, float* out, void* stream);

/* Entropic Gromov-Wasserstein sample couplings, one label per CTA, float64, everything in shared memory
 * (get_coupling_egw_ott_fixed, MRI_PET_OT_OT_per_epoch_attn.py:129-186).  X: all labels'
 * rows concatenated ((sum n_l) x dx fp32), xoff: nprob + 1 row offsets (device int32); same for Y; n_l, m_l <= 64.
 * Geometries: squared Euclidean / max (PointCloud(scale_cost="max_cost"), :155-156); square loss; uniform
 * marginals; outer loop = ott GromovWasserstein(epsilon, max_iterations = gw_max_iter, min_iterations,
 * threshold), inner = log-domain Sinkhorn(max_iterations = sk_max_iter, threshold, inner_iterations =
 * sk_check_every), warm-started (:168-173).  Outputs: T (float64, label l at element offset toff[l], n_l x m_l),
 * info4[l] = {outer iterations, outer converged, last inner converged, inner iterations in total}, cost[l].   */
int b200ot_egw_batched(const float* X, const float* Y, const int* xoff, const int* yoff, const long long* toff,
                       int nprob, int max_n, int max_m, int dx, int dy, float eps, int gw_max_iter, int gw_min_iter,
                       float gw_threshold, int sk_max_iter, int sk_check_every, float sk_threshold, double* T,
                       int* info4, double* cost, void* stream);

/* ---- attention fusion core ----------------------------------------------------
 * softmax(Q K^T / sqrt(dh)) V for the S <= 4 fusion tokens of the reference's SelfAttentionBlock
 * (MRI_PET_OT_OT_per_epoch_attn.py:523-549, tokens built at :731-738; one token in
 * MRI_PET_OT_nojax.py:664-669).  qkv is (S, B, 3E) as produced by nn.MultiheadAttention's
 * in_proj (batch_first=False), out is (S, B, E), probs (B, H, S, S) keeps the softmax for the
 * backward pass.  keep_mask (B, H, S, S; may be NULL) with keep_scale = 1/(1-p) is the attention
 * dropout of training mode.  One warp per (sample, head), forward and backward.          */
int b200ot_token_attention_fwd(const float* qkv, int S, int B, int E, int H, const float* keep_mask,
                               float keep_scale, float* out, float* probs, void* stream);
int b200ot_token_attention_bwd(const float* qkv, const float* probs, const float* keep_mask,
                               float keep_scale, const float* dout, int S, int B, int E, int H,
                               float* dqkv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200OT_H_ */

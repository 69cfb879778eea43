/*
 * roman_b200.h — C ABI of libroman_b200.so: the B200-native (sm_100a) replacement for the
 * CLIPPER registration hot path that roman.align drives through the `clipperpy` pybind11 module.
 *
 * The reference has no C ABI for this path (pybind11 is a C++ ABI); each entry point below
 * names the reference interface it replaces (paths relative to the reference repo,
 * clipper/ = dependencies/clipper/).  Plain pointers and sizes only — no torch, no C++ types.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative rb_status on error; the message is
 *     available from rb_last_error().  Per-problem soft failures (too few inliers) are
 *     reported in `status[b]`, never as a return code (mirrors the Python exceptions caught
 *     per pair in roman/align/submap_align.py:179-184).
 *   - "Device" entry points take DEVICE pointers and a cudaStream_t (passed as void*); they
 *     only enqueue work on that stream and never synchronise (two documented exceptions: rb_set_dense
 *     reads back a validity flag, and rb_solve with rounding == DSD runs its host epilogue).  "Host" entry points (_host
 *     suffix) take HOST pointers, do the H2D/D2H copies themselves and return after the
 *     results are in the host buffers.
 *   - A batch is B independent submap-pair problems.  Problem b has n1[b], n2[b] objects and
 *     m[b] putative associations (m = n1*n2 when all-to-all).  Per-object feature rows are
 *     object-major: D1 is the concatenation over b of n1[b] rows of d doubles — exactly the
 *     memory of the (d x n1) column-major Eigen matrix the reference receives
 *     (clipper/bindings/python/py_clipper.cpp:316-318, roman/align/roman_registration.py:88-96).
 *   - Per-association arrays (A, u0, u, nodes, sel) are the concatenation over b of m[b]
 *     entries, problem b starting at rb_assoc_offset(h, b).
 */
#ifndef ROMAN_B200_H
#define ROMAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_MAX_RATIO_FEATURES 8

typedef struct rb_handle rb_handle;

typedef enum {
  RB_OK = 0,
  RB_ERR_ARG = -1,       /* bad argument / unsupported configuration */
  RB_ERR_CUDA = -2,      /* CUDA runtime error */
  RB_ERR_NOPLAN = -3,    /* stage called before rb_plan / out of order */
  RB_ERR_TOO_LARGE = -4, /* m exceeds what the solver's shared-memory layout supports */
  RB_ERR_ALLOC = -5
} rb_status;

/* per-problem status written by rb_arun / rb_align_batch */
#define RB_PAIR_OK 0
#define RB_PAIR_INSUFFICIENT 1 /* fewer than point_dim inliers: T filled with NaN
                                  (InsufficientAssociationsException, object_registration.py:107-108) */

/* Which invariant scores the pairs. */
#define RB_INVARIANT_ROMAN 0     /* clipper::invariants::ROMAN      (pairwise + single) */
#define RB_INVARIANT_EUCLIDEAN 1 /* clipper::invariants::EuclideanDistance (pairwise)   */
#define RB_INVARIANT_ROMAN_PAIRWISE 2 /* the ROMAN invariant scored through CLIPPER::scorePairwiseConsistency
                                         (clipper/src/clipper.cpp:29-73): ROMAN::operator() is pairwise_similarity alone —
                                         no single-association scores, no single-score gate, no fusion (abstract.h:99-102) */

/* POD mirror of clipper::invariants::ROMAN::Params (clipper/include/clipper/invariants/roman.h:27-53)
 * and EuclideanDistance::Params (euclidean_distance.h:24-29; uses sigma, epsilon, mindist only). */
typedef struct {
  int32_t kind;              /* RB_INVARIANT_* */
  int32_t point_dim;         /* 2 or 3 */
  int32_t ratio_feature_dim; /* 0..RB_MAX_RATIO_FEATURES */
  int32_t cos_feature_dim;   /* 0 or D */
  int32_t fusion_method;     /* 0 GEOMETRIC_MEAN, 1 ARITHMETIC_MEAN, 2 PRODUCT (roman.h:22-26) */
  int32_t gravity_guided;
  int32_t drift_aware;
  int32_t drift_scale_sigma;
  double sigma, epsilon, mindist;
  double distance_weight, ratio_weight, cosine_weight;
  double ratio_epsilon[RB_MAX_RATIO_FEATURES];
  double cosine_min, cosine_max;
  double gravity_unc_ang_rad;
  double drift_scale;
} rb_invariant_params;

/* POD mirror of clipper::Params (clipper/include/clipper/clipper.h:27-60). */
typedef struct {
  double tol_u, tol_F, tol_Fop;
  int32_t maxiniters, maxoliters;
  double beta;
  int32_t maxlsiters;
  double eps;
  double affinityeps;
  int32_t rescale_u0;
  int32_t rounding; /* 0 NONZERO, 1 DSD (exact densest subgraph; host epilogue, synchronises), 2 DSD_HEU */
} rb_solver_params;

/* Fill with the reference defaults (roman.h:27-53, clipper.h:27-60). */
void rb_default_invariant_params(rb_invariant_params* p);
void rb_default_solver_params(rb_solver_params* p);

/* ---- lifetime ------------------------------------------------------------------------- */

/* One handle per (device, stream user).  Owns only scratch: packed M storage, association
 * records, work vectors; sized lazily by rb_plan and reused.  Not thread-safe.
 * Replaces: construction of clipperpy.CLIPPER / CLIPPERPairwiseAndSingle
 * (py_clipper.cpp:258-332; roman/align/roman_registration.py:82-86). */
int rb_create(rb_handle** out, int device);
int rb_destroy(rb_handle* h);
const char* rb_last_error(const rb_handle* h);

/* Replaces clipperpy.invariants.ROMANParams / ROMAN(...) / EuclideanDistanceParams
 * (py_clipper.cpp:51-122) and clipperpy.Params (py_clipper.cpp:226-242). */
int rb_set_invariant_params(rb_handle* h, const rb_invariant_params* p);
int rb_set_solver_params(rb_handle* h, const rb_solver_params* p);

/* ---- planning ------------------------------------------------------------------------- */

/* Describe the batch shapes (HOST arrays of length B).  m == NULL means all-to-all
 * (m[b] = n1[b]*n2[b], A implicit: utils::createAllToAll, clipper/include/clipper/utils.h:61-71).
 * d = doubles per object row.  Computes offsets, (re)allocates scratch, uploads descriptors.
 * Does not touch the compute stream; call it outside timed regions when shapes repeat. */
int rb_plan(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2,
            const int32_t* m);

/* Same, for batches whose problems share objects (the all-pairs loop-closure search of
 * roman/align/submap_align.py:93-94 registers every submap of run A against every submap of run B):
 * D1 / D2 are shared tables of total_obj1 / total_obj2 object rows, uploaded once, and problem b uses
 * rows [obj1_offset[b], obj1_offset[b] + n1[b]) of D1 and [obj2_offset[b], ... + n2[b]) of D2. */
int rb_plan_indexed(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m,
                    const int64_t* obj1_offset, const int64_t* obj2_offset, int64_t total_obj1, int64_t total_obj2);
/* Same as rb_plan_indexed, for batches in which several problems are the SAME submap pair solved from different start
 * vectors (multi-start: CLIPPER is a local method and the reference draws one random start per solve,
 * clipper/src/utils.cpp:22-29; roman/align/object_registration.py:57-86 re-solves one pair several times).
 * share[b] == b: problem b owns its affinity matrix; share[b] == a < b (with share[a] == a, identical n1, n2, m and
 * object ranges): problem b uses a's matrix, association records and edge list — K1 builds and stores the matrix
 * once, K2 / K3 run for every problem with its own u0 and outputs.  The solver clusters of a group stream the same
 * tiles at the same time, so each pass reads the matrix from HBM once and from L2 for the other starts; per-start
 * results are bit-identical to solving the problems separately. */
int rb_plan_shared(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m,
                   const int64_t* obj1_offset, const int64_t* obj2_offset, int64_t total_obj1, int64_t total_obj2,
                   const int32_t* share);
int64_t rb_assoc_offset(const rb_handle* h, int32_t b); /* b == B gives the total */
int64_t rb_obj1_offset(const rb_handle* h, int32_t b);  /* in objects (rows of d doubles) */
int64_t rb_obj2_offset(const rb_handle* h, int32_t b);
int64_t rb_packed_bytes(const rb_handle* h);            /* bytes of packed M for the batch */

/* ---- device-pointer stages ---------------------------------------------------------------- */

/* K1a + K1.  Replaces CLIPPER::scorePairwiseConsistency (clipper/src/clipper.cpp:29-73) and
 * CLIPPERPairwiseAndSingle::scorePairwiseAndSingleConsistency (:77-136) with the ROMAN /
 * EuclideanDistance invariants (clipper/src/invariants/roman.cpp, euclidean_distance.cpp).
 * A: device int32 (sum m) x 2 row-major, or NULL when the plan is all-to-all.
 * Result stays in the handle (packed strict-upper M in 32x32 tiles; C = pattern(M)). */
int rb_affinity(rb_handle* h, const double* D1, const double* D2, const int32_t* A, void* stream);

/* K2.  Replaces CLIPPER::solve / findDenseClique (clipper/src/clipper.cpp:140-149, 243-394),
 * utils::findIndicesOfkLargest (clipper/src/utils.cpp:33-55) and
 * utils::selectInlierAssociations (:101-108).
 * u0: device (sum m) doubles — REQUIRED (the reference draws it from std::random_device,
 *     clipper/src/utils.cpp:22-29; callers wanting that behaviour fill u0 with U[0,1)).
 * Outputs (device; any may be NULL): nodes (sum m int32, first n_nodes[b] valid per problem,
 * ordered as the reference: descending u), sel (sum m x 2 int32: rows of A at nodes),
 * n_nodes (B), u (sum m), score (B), ifinal (B), n_passes (B: fused (M x, C x) evaluations = 2 + line-search trials;
 * see rb_solver_streamed_passes for how many of them read M from HBM). */
int rb_solve(rb_handle* h, const double* u0, int32_t* nodes, int32_t* sel, int32_t* n_nodes,
             double* u, double* score, int32_t* ifinal, int32_t* n_passes, void* stream);

/* K3.  Replaces ObjectRegistration.T_align (roman/align/object_registration.py:88-129):
 * Arun/Umeyama with unit weights on the matched object centres (first point_dim entries of
 * the rows).  sel / n_sel as produced by rb_solve (or any correspondences laid out the same).
 * T: device B x 16 doubles, row-major 4x4 mapping frame-2 points into frame 1 (for
 * point_dim == 2 the 2x2 rotation and translation sit in rows/cols {0,1} and column 3).
 * status: device B int32 (RB_PAIR_*). */
int rb_arun(rb_handle* h, const double* D1, const double* D2, const int32_t* sel,
            const int32_t* n_sel, double* T, int32_t* status, void* stream);

/* rb_affinity + rb_solve + rb_arun on one stream: one ObjectRegistration.register + T_align
 * per problem (object_registration.py:22-29, 88-129).  Optional outputs may be NULL. */
int rb_align_batch(rb_handle* h, const double* D1, const double* D2, const int32_t* A,
                   const double* u0, int32_t* sel, int32_t* n_sel, double* T, int32_t* status,
                   double* score, int32_t* n_passes, void* stream);

/* Host-only building block of Rounding::DSD.  Replaces clipper::dsd::solve(A, S) (clipper/src/dsd.cpp:265-314 with
 * densest_subgraph :157-262): Goldberg's exact maximum-density subgraph of the weighted graph induced by the k nodes
 * S (global ids, ascending) of a problem with n nodes, W = k x k row-major symmetric weights (upper triangle
 * authoritative, diagonal ignored).  Uses the reference's search constants, hence returns the same set: the source
 * side of the last non-trivial minimum cut, ascending ids, written to out (capacity k); *n_out = its size.
 * k == 1 returns S; an edgeless S returns nothing (the reference reads uninitialised memory in both cases).
 * rb_solve / rb_align_batch* call it per problem when rb_solver_params.rounding == 1: the kernel leaves
 * S = {i : u_i > 0} (clipper.cpp:364-367), the S x S block of M is copied to the host, and the result overwrites
 * nodes / sel / n_nodes — this mode synchronises the stream.  No GPU is needed to call this function. */
int rb_dsd_host(int64_t n, int32_t k, const int32_t* S, const double* W, int32_t* out, int32_t* n_out);

/* Submap-descriptor prefilter of the pair loop.  Replaces Submap.similarity (roman/map/map.py:145-162)
 * as called per pair at roman/align/submap_align.py:131-134: sim[i * S1 + j] = max over (frame a of
 * submap i in run 0, frame b of submap j in run 1) of cos(F0[a], F1[b]), a term being 0 when the
 * product of the two norms is within 1e-9 of zero; NaN terms propagate like numpy's max.
 * off0 (S0 + 1) / off1 (S1 + 1): device int64 row offsets of each submap's stacked frame
 * descriptors in F0 / F1 (device, row-major, D doubles per row; a mean descriptor is one row).
 * Every submap must have at least one row (numpy raises on an empty max).  sim: device S0 x S1. */
int rb_submap_similarity(rb_handle* h, int32_t S0, int32_t S1, int32_t D, const int64_t* off0,
                         const int64_t* off1, const double* F0, const double* F1, double* sim,
                         void* stream);

/* Replaces CLIPPER::getAffinityMatrix / getConstraintMatrix (clipper/src/clipper.cpp:202-216):
 * dense m x m row-major M + I and C + I of problem b into device buffers (either may be NULL). */
int rb_get_dense(rb_handle* h, int32_t b, double* M_out, double* C_out, void* stream);

/* Replaces CLIPPER::setMatrixData (clipper/src/clipper.cpp:220-229): take the strict upper
 * triangles of dense row-major device matrices as problem b's M and C.  Entries with C != 0
 * and M == 0 are kept (stored as -0.0); entries with M != 0 and C == 0 are unsupported. */
int rb_set_dense(rb_handle* h, int32_t b, const double* M, const double* C, void* stream);

/* Tuning knob of K2: how many CTAs (SMs) of one thread-block cluster share one problem
 * (0 = auto, else 1, 2, 4 or 8).  Takes effect at the next rb_plan.  Results do not depend on
 * it beyond the last bits of u (different summation tree). */
int rb_set_solver_cluster(rb_handle* h, int32_t ctas_per_problem);
int rb_get_solver_cluster(const rb_handle* h);
int rb_get_solver_max_clusters(const rb_handle* h); /* co-resident clusters of that size (after rb_plan) */

/* Storage format of the packed affinity matrix.
 * off (default) — DENSE: every tile of the packed upper triangle is written by K1 and streamed by K2 each pass.
 * on — TILE SKIP, the sparse fast path: K1 records per 32x32 tile whether it holds any non-zero word, writes only
 *   such tiles and lists the edges; K2 reads only what is stored and solves small problems from a shared-memory CSR.
 * The stored matrices are bitwise identical in both formats, the solves agree to rounding (summation order).
 * Changing the format invalidates the plan: call rb_plan again (the solver's cluster size is chosen per format). */
int rb_set_tile_skip(rb_handle* h, int32_t on);

/* After rb_solve: device int32[B] — 0: the problem's stored tiles were streamed every pass; 1: resident from the start
 * (tile-skip format), the shared-memory CSR was built from one pass over the stored tiles; 2: resident from the start,
 * the CSR was built from the edge list K1 wrote (no tile read at all); 3: dense format, promoted after
 * rb_set_solver_promote streamed passes by one more (extraction) pass over the tiles; 4: dense format, promoted from
 * K1's edge list (no extra pass). */
int rb_solver_resident(rb_handle* h, int32_t* out, void* stream);

/* Dense format only — late residency.  A problem that is still iterating after `passes` streamed passes of K2 builds a
 * shared-memory CSR of its non-zero words — from the edge list K1 wrote beside the tiles (12 bytes per edge), or, for
 * matrices given by rb_set_dense, by reading its tiles once more — and runs every later pass from there (if
 * they fit: about 6-7 thousand edges per SM of the cluster at m = 2000; otherwise it keeps streaming).  Slow-converging problems — tens of passes where a batch averages ~5 —
 * would otherwise pin a cluster to tens of full streams of a matrix holding a few hundred edges.  Default 8;
 * 0 = never.  rb_solver_resident reports 3 / 4 for promoted problems.  Results agree with the unpromoted solve to
 * rounding (summation order), like the two storage formats do. */
int rb_set_solver_promote(rb_handle* h, int32_t passes);

/* After rb_solve: device int32[B] — how many passes of each problem read the packed matrix from HBM (every pass of a
 * streamed problem; the passes before promotion plus the extraction pass of a promoted one; 1 or 0 for problems
 * resident from the start).  n_passes of rb_solve counts ALL (M x, C x) evaluations regardless of where M came from. */
int rb_solver_streamed_passes(rb_handle* h, int32_t* out, void* stream);

/* After rb_affinity (both formats): edges (non-zero words of the strict upper triangle) K1 found per problem,
 * device int32[B]; a count above the list capacity (8 * 32 * ceil(m/32)) means the list overflowed and is not used. */
int rb_edge_counts(rb_handle* h, int32_t* out, void* stream);

/* After rb_affinity: associations per problem that can carry an edge at all (not padding, single score not below
 * affinityeps — clipper/src/clipper.cpp:104-108), device int32[B].  L live associations give L(L-1)/2 evaluated pairs:
 * the flop count of K1 (bench.py's fp64 accounting). */
int rb_live_counts(rb_handle* h, int32_t* out, void* stream);

/* Tile-skip format only: number of stored (flagged) 32x32 tiles of each problem, device int32[B]. */
int rb_stored_tiles(rb_handle* h, int32_t* out, void* stream);

/* Device milliseconds of the last rb_align_batch / stage calls, measured with CUDA events on
 * the launch stream: ms[0]=prep(K1a) ms[1]=affinity(K1) ms[2]=solve(K2) ms[3]=arun(K3).
 * Synchronises on the recorded events.  Enabled by rb_set_profiling(h, 1). */
int rb_set_profiling(rb_handle* h, int32_t on);
int rb_get_stage_ms(rb_handle* h, float ms[4]);

/* ---- host-buffer entry point (what a cgo / ctypes / JNI binding of the reference would call) ---- */

/* Same as rb_align_batch with HOST pointers: copies D1, D2, A, u0 to the device, runs the
 * three stages, copies sel, n_sel, T, status, score, n_passes back, synchronises.  Pinned host
 * buffers make the copies asynchronous; pageable ones work too.  Requires a prior rb_plan.
 * The uploads are pipelined with the kernels over problem ranges fixed by rb_plan (cut by work, so ragged batches
 * balance).  `sel` keeps the layout of rb_solve ((sum m) x 2, problem b at rb_assoc_offset) but only the first
 * n_sel[b] rows of each problem are written: the device packs them compactly (at most min(n1, n2) rows per problem
 * can form a set of distinct associations) and only that block crosses PCIe. */
int rb_align_batch_host(rb_handle* h, const double* D1, const double* D2, const int32_t* A,
                        const double* u0, int32_t* sel, int32_t* n_sel, double* T,
                        int32_t* status, double* score, int32_t* n_passes);

/* Bytes the last rb_align_batch_host call moved: direction 0 = host to device, 1 = device to host. */
int64_t rb_host_copy_bytes(const rb_handle* h, int32_t direction);

/* Diagnostics (used by tools/): first call enables per-CTA cycle counters in the solver kernel, later calls copy
 * them out: 8 x int64 per CTA — total, mat-vec, cluster combine, queue wait, rounding, list/CSR build. */
int rb_debug_solver_cycles(rb_handle* h, long long* out, int n_ctas);

/* Library/build identification. */
const char* rb_version(void);
int rb_device_sm_count(const rb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ROMAN_B200_H */

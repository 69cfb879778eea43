// rb_kernels.cuh — launch wrappers shared between the kernel translation units and rb_api.cu.
#pragma once
#include "rb_common.cuh"

// Invariant parameters as the kernels see them (roman.h:55-60 caches cos/sin of the angle).
struct RbDevParams {
  rb_invariant_params inv;
  double cos_unc, sin_unc;
  double affinityeps;
};

// Everything the persistent solver kernel needs.
struct RbSolveArgs {
  const RbProblem* probs;
  int32_t B;
  int32_t b0;              // batch index of probs[0] (problem ranges): RbProblem::base is a batch index
  int32_t* counter;        // dynamic problem queue (zeroed before launch)
  const int32_t* order;    // queue position -> problem index (longest problem first), or null = identity
  const double* Mp;        // packed tiles
  const uint8_t* flags;    // per stored tile: 1 = holds a non-zero word; null = every tile is streamed (dense mode)
  int32_t nt_max;          // max stored tiles of one problem (sizes the shared-memory copy of the flags)
  const unsigned* edge_ij; // K1's edge lists (RB_EDGE_FACTOR * nb * 32 slots per problem, i << 16 | j) ...
  const double* edge_v;    // ... their values ...
  const int* edge_cnt;     // ... and counts (may exceed the capacity: overflow, list unusable); null = no lists
  const int32_t* i1;       // association object indices (padded arrays), for `sel`
  const int32_t* i2;
  const double* u0;        // user layout (off_assoc)
  double* u_work;          // padded scratch (off_pad): current iterate u
  double* g_work;          // padded scratch: gradient at u
  int64_t pad_stride;      // u_work/g_work hold `cluster` private copies, this many doubles apart
  int32_t cluster;         // CTAs per problem (1, 2, 4, 8)
  int32_t slots, depth;    // shared-memory configuration chosen by rb_solver_config
  long long* debug;        // optional per-CTA cycle counters (8 per CTA), or null
  // outputs, user layout (off_assoc); rb_api substitutes handle scratch for NULL user pointers,
  // so only u_out may be null here
  int32_t* nodes;
  int32_t* sel;
  int32_t* n_nodes;
  double* u_out;
  double* score;
  int32_t* ifinal;
  int32_t* n_passes;
  int32_t* resident;       // per problem: 0 tiles streamed every pass; 1 / 2 resident from the start (tile-skip format);
                           // 3 promoted: streamed `promote_after` passes, then extracted into the shared-memory CSR
  int32_t* streamed;       // per problem: passes (incl. an extraction pass) that read the packed matrix from HBM
  int promote_after;       // dense format: after this many streamed passes try to make the problem resident (0 = never)
  rb_solver_params sp;
  int32_t mp_max;          // max padded m over the batch (sizes the shared-memory vectors)
  int32_t smem_total;      // dynamic shared memory of the launch (the resident CSR uses everything past the vectors)
};

cudaError_t rb_launch_prep(const RbProblem* probs, int B, int max_mp, int max_n1, int max_n2, bool all_to_all,
                           const RbAssocSoA& rec, const double* D1, const double* D2, const int32_t* A,
                           const RbDevParams& P, cudaStream_t st);
cudaError_t rb_launch_affinity(const RbProblem* probs, int B, int64_t tile_base, int64_t total_tiles, int max_nb,
                               const RbAssocSoA& rec, uint32_t* masks, double* Mp, uint8_t* flags,
                               unsigned* edge_ij, double* edge_v, int* edge_cnt, const RbDevParams& P, int sm_count,
                               cudaStream_t st);
cudaError_t rb_launch_count_flags(const RbProblem* probs, int B, const uint8_t* flags, int32_t* out, cudaStream_t st);
cudaError_t rb_launch_unpack(const RbProblem& pr, const double* Mp, const uint8_t* flags, double* Mout, double* Cout,
                             cudaStream_t st);
cudaError_t rb_launch_pack(const RbProblem& pr, const double* Min, const double* Cin, double* Mp, int* bad,
                           cudaStream_t st);

// Returns cudaErrorInvalidValue if mp_max does not fit the shared-memory layout.
size_t rb_solver_smem_bytes(int mp_max, int slots, int depth);
bool rb_solver_config(int mp_max, size_t optin, int* slots, int* depth);
cudaError_t rb_solver_prepare(int device);  // opt in to the large dynamic shared-memory carve-out
cudaError_t rb_solver_max_clusters(int mp_max, int cluster, int* out);
cudaError_t rb_launch_solver(const RbSolveArgs& a, int max_clusters, cudaStream_t st);

cudaError_t rb_launch_gather_sub(const RbProblem& pr, const double* Mp, const uint8_t* flags, const int32_t* nodes, int k,
                                 double* W, cudaStream_t st);
cudaError_t rb_launch_select(const int32_t* nodes, int k, const int32_t* i1, const int32_t* i2, int32_t* sel,
                             cudaStream_t st);
cudaError_t rb_launch_similarity(int S0, int S1, int D, const int64_t* off0, const int64_t* off1, const double* F0,
                                 const double* F1, double* sim, cudaStream_t st);
cudaError_t rb_launch_arun(const RbProblem* probs, int B, int point_dim, const double* D1, const double* D2,
                           const int32_t* sel, const int32_t* n_sel, double* T, int32_t* status,
                           int32_t* selpad, int kcap, cudaStream_t st);
cudaError_t rb_launch_live_counts(const RbProblem* probs, int B, const uint32_t* masks, int32_t* out, cudaStream_t st);

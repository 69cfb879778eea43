// rb_api.cu — handle, batch planning and the extern "C" entry points of include/roman_b200.h.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "rb_kernels.cuh"

#define RB_VERSION_STR "roman_b200 0.1.0 (sm_100a)"

struct rb_handle {
  int device = 0;
  int sm_count = 0;
  std::string err;
  rb_invariant_params inv;
  rb_solver_params sol;

  // plan (host copies)
  bool planned = false;
  int32_t B = 0, d = 0;
  bool all_to_all = true;
  bool indexed = false;      // rb_plan_indexed: problems reference ranges of shared object tables
  bool shared = false;       // rb_plan_shared: some problems use another problem's matrix (one range only)
  std::vector<RbProblem> probs;
  std::vector<int64_t> tile_prefix;  // B+1
  int64_t total_assoc = 0, total_pad = 0, total_tiles = 0, total_obj1 = 0, total_obj2 = 0;
  int32_t max_mp = 0, max_n1 = 0, max_n2 = 0;
  int32_t cluster_req = 0;   // 0 = auto
  int32_t slots = 2, depth = 8;  // solver shared-memory configuration (rb_solver_config)
  int32_t cluster = 1;       // resolved at plan time
  int32_t max_clusters = 1;  // co-resident clusters of that size

  // device scratch
  RbProblem* d_probs = nullptr; size_t cap_probs = 0;
  int64_t* d_tile_prefix = nullptr; size_t cap_prefix = 0;
  double* d_M = nullptr; size_t cap_M = 0;            // packed tiles
  uint8_t* d_flags = nullptr; size_t cap_flags = 0;   // per stored tile: holds a non-zero word (tile-skip mode)
  bool tile_skip = false;                             // rb_set_tile_skip (default: dense format)
  double* d_pad_f64 = nullptr; size_t cap_pad = 0;    // padded double arrays (7 record + u, g per cluster CTA)
  int32_t* d_pad_i32 = nullptr;                       // 2 padded int arrays
  uint32_t* d_masks = nullptr;                        // one live-mask word per 32-association block
  unsigned* d_edge_ij = nullptr; double* d_edge_v = nullptr;  // K1's edge lists (padded like the records)
  int* d_edge_cnt = nullptr;                          // edges per problem (B)
  int32_t* d_counter = nullptr;
  int* d_bad = nullptr;
  long long* d_debug = nullptr;  // dev builds: per-CTA cycle counters of the solver
  double* d_W = nullptr; size_t cap_W = 0;  // Rounding::DSD: |S| x |S| block of one problem's affinity matrix
  // internal outputs used when the caller passes NULL
  int32_t* d_nodes = nullptr; int32_t* d_sel = nullptr; size_t cap_assoc = 0;
  int32_t* d_nsel = nullptr; double* d_score = nullptr; int32_t* d_ifinal = nullptr; int32_t* d_npass = nullptr;
  int32_t* d_resident = nullptr;
  int32_t* d_streamed = nullptr;
  // solver queue order (longest problem first): d_order for whole-batch launches, d_order_chunks for the problem
  // ranges of the host entry point (entries are indices relative to the start of their range)
  int32_t* d_order = nullptr; int32_t* d_order_chunks = nullptr; size_t cap_order = 0, cap_order_chunks = 0;
  int nch = 1;               // problem ranges of rb_align_batch_host (fixed at plan time)
  int bounds[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  // compact read-back of the inlier associations: the first kcap rows of every problem's `sel`
  int32_t kcap = 0;
  int32_t* d_selpad = nullptr; size_t cap_selpad = 0;
  int32_t* h_selpad = nullptr; int32_t* h_nsel = nullptr; size_t cap_hsel = 0, cap_hnsel = 0;  // pinned host staging
  int32_t promote_after = 8;  // dense format: streamed passes before a still-iterating problem is made resident
  int32_t* d_status = nullptr; double* d_T = nullptr; size_t cap_B = 0;
  // staging for the host entry point
  double* s_D1 = nullptr; double* s_D2 = nullptr; size_t cap_sD1 = 0, cap_sD2 = 0;
  int32_t* s_A = nullptr; double* s_u0 = nullptr; size_t cap_sA = 0, cap_su0 = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t own_stream2 = nullptr;  // second compute stream: consecutive ranges alternate, so the next range's
                                       // kernels fill the SMs the previous range's solver tail leaves idle
  cudaEvent_t join_ev = nullptr;
  cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t last_stream = nullptr;  // stream of the most recent device-pointer stage call
  bool last_valid = false;
  cudaEvent_t order_ev = nullptr;

  bool affinity_done = false;
  bool masks_valid = false;   // rb_affinity wrote the block live masks (not after rb_set_dense)
  int64_t last_h2d = 0, last_d2h = 0;  // bytes the last rb_align_batch_host call copied
  bool edges_valid = false;   // K1 wrote the edge lists of the current packed matrix (not after rb_set_dense)
  bool profiling = false;
  // e0 prep start, e1 prep end = affinity start, e2 affinity end, e3/e4 solve, e5/e6 arun
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid[7] = {false, false, false, false, false, false, false};
};

namespace {

int fail(rb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
int cuda_fail(rb_handle* h, cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  return fail(h, RB_ERR_CUDA, buf);
}
#define RB_CUDA(h, call)                                   \
  do {                                                     \
    cudaError_t _e = (call);                               \
    if (_e != cudaSuccess) return cuda_fail(h, _e, #call); \
  } while (0)

template <typename T>
int ensure(rb_handle* h, T** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t n = need ? need : 1;
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc");
  *cap = n;
  return 0;
}

RbAssocSoA soa_of(rb_handle* h) {
  RbAssocSoA r;
  const size_t n = (size_t)h->total_pad;
  double* f = h->d_pad_f64;
  r.ax = f; r.ay = f + n; r.az = f + 2 * n; r.bx = f + 3 * n; r.by = f + 4 * n; r.bz = f + 5 * n;
  r.single = f + 6 * n;
  r.i1 = h->d_pad_i32;
  r.i2 = h->d_pad_i32 + n;
  return r;
}
constexpr int kMaxCluster = 8;
constexpr int kPadArrays = 7 + 2 * kMaxCluster;  // 7 record arrays + per-CTA copies of u and g
double* u_work(rb_handle* h) { return h->d_pad_f64 + 7 * (size_t)h->total_pad; }
double* g_work(rb_handle* h) { return h->d_pad_f64 + (7 + kMaxCluster) * (size_t)h->total_pad; }

RbDevParams dev_params(const rb_handle* h) {
  RbDevParams P;
  P.inv = h->inv;
  P.cos_unc = cos(h->inv.gravity_unc_ang_rad);  // roman.h:57-58
  P.sin_unc = sin(h->inv.gravity_unc_ang_rad);
  P.affinityeps = h->sol.affinityeps;
  return P;
}

// A handle is tied to one device; stage calls may arrive while another device is current.
#define RB_ENTER(h) RB_CUDA(h, cudaSetDevice((h)->device))

void record(rb_handle* h, int i, cudaStream_t st) {
  if (!h->profiling) return;
  if (!h->ev[i]) cudaEventCreate(&h->ev[i]);
  cudaEventRecord(h->ev[i], st);
  h->ev_valid[i] = true;
}

}  // namespace

extern "C" {

void rb_default_invariant_params(rb_invariant_params* p) {
  memset(p, 0, sizeof *p);
  p->kind = RB_INVARIANT_ROMAN;
  p->point_dim = 3;
  p->fusion_method = 0;
  p->sigma = 0.01;
  p->epsilon = 0.06;
  p->mindist = 0.0;
  p->distance_weight = p->ratio_weight = p->cosine_weight = 1.0;
  p->cosine_min = 0.85;
  p->cosine_max = 0.95;
  p->drift_scale = 0.1;
}

void rb_default_solver_params(rb_solver_params* p) {
  p->tol_u = 1e-8; p->tol_F = 1e-9; p->tol_Fop = 1e-10;
  p->maxiniters = 200; p->maxoliters = 1000;
  p->beta = 0.25; p->maxlsiters = 99;
  p->eps = 1e-9; p->affinityeps = 1e-4;
  p->rescale_u0 = 1; p->rounding = 2;
}

const char* rb_version(void) { return RB_VERSION_STR; }

int rb_create(rb_handle** out, int device) {
  if (!out) return RB_ERR_ARG;
  *out = nullptr;
  rb_handle* h = new rb_handle();
  h->device = device;
  rb_default_invariant_params(&h->inv);
  rb_default_solver_params(&h->sol);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete h; return RB_ERR_CUDA; }
  int sm = 0;
  e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) { delete h; return RB_ERR_CUDA; }
  h->sm_count = sm;
  e = rb_solver_prepare(device);
  if (e != cudaSuccess) { delete h; return RB_ERR_CUDA; }
  if (cudaMalloc((void**)&h->d_counter, 8 * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc((void**)&h->d_bad, sizeof(int)) != cudaSuccess) {
    rb_destroy(h);
    return RB_ERR_ALLOC;
  }
  *out = h;
  return RB_OK;
}

int rb_destroy(rb_handle* h) {
  if (!h) return RB_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_probs); cudaFree(h->d_tile_prefix); cudaFree(h->d_M); cudaFree(h->d_flags); cudaFree(h->d_pad_f64);
  cudaFree(h->d_pad_i32); cudaFree(h->d_masks); cudaFree(h->d_edge_ij); cudaFree(h->d_edge_v); cudaFree(h->d_edge_cnt);
  cudaFree(h->d_counter); cudaFree(h->d_bad); cudaFree(h->d_nodes); cudaFree(h->d_sel);
  cudaFree(h->d_nsel); cudaFree(h->d_score); cudaFree(h->d_ifinal); cudaFree(h->d_npass);
  cudaFree(h->d_status); cudaFree(h->d_T); cudaFree(h->d_resident); cudaFree(h->d_streamed);
  cudaFree(h->s_D1); cudaFree(h->s_D2); cudaFree(h->s_A); cudaFree(h->s_u0); cudaFree(h->d_debug); cudaFree(h->d_W);
  cudaFree(h->d_order); cudaFree(h->d_order_chunks); cudaFree(h->d_selpad);
  if (h->h_selpad) cudaFreeHost(h->h_selpad);
  if (h->h_nsel) cudaFreeHost(h->h_nsel);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream2) cudaStreamDestroy(h->own_stream2);
  if (h->join_ev) cudaEventDestroy(h->join_ev);
  for (auto& e : h->chunk_ev) if (e) cudaEventDestroy(e);
  delete h;
  return RB_OK;
}

const char* rb_last_error(const rb_handle* h) { return h ? h->err.c_str() : "null handle"; }
int rb_device_sm_count(const rb_handle* h) { return h ? h->sm_count : 0; }

int rb_set_invariant_params(rb_handle* h, const rb_invariant_params* p) {
  if (!h || !p) return RB_ERR_ARG;
  if (p->kind != RB_INVARIANT_ROMAN && p->kind != RB_INVARIANT_EUCLIDEAN && p->kind != RB_INVARIANT_ROMAN_PAIRWISE)
    return fail(h, RB_ERR_ARG, "unknown invariant kind");
  if (p->kind != RB_INVARIANT_EUCLIDEAN) {
    if (p->point_dim != 2 && p->point_dim != 3) return fail(h, RB_ERR_ARG, "point_dim must be 2 or 3");
    if (p->gravity_guided && p->point_dim != 3) return fail(h, RB_ERR_ARG, "gravity_guided requires point_dim == 3");
    if (p->ratio_feature_dim < 0 || p->ratio_feature_dim > RB_MAX_RATIO_FEATURES)
      return fail(h, RB_ERR_ARG, "ratio_feature_dim out of range");
    if (p->cos_feature_dim < 0) return fail(h, RB_ERR_ARG, "cos_feature_dim < 0");
    if (p->fusion_method < 0 || p->fusion_method > 2) return fail(h, RB_ERR_ARG, "unknown fusion_method");
  }
  h->inv = *p;
  return RB_OK;
}

int rb_set_solver_params(rb_handle* h, const rb_solver_params* p) {
  if (!h || !p) return RB_ERR_ARG;
  if (p->rounding < 0 || p->rounding > 2) return fail(h, RB_ERR_ARG, "unknown rounding");
  h->sol = *p;
  return RB_OK;
}

int rb_set_solver_cluster(rb_handle* h, int32_t ctas_per_problem) {
  if (!h) return RB_ERR_ARG;
  if (ctas_per_problem != 0 && ctas_per_problem != 1 && ctas_per_problem != 2 && ctas_per_problem != 4 &&
      ctas_per_problem != 8)
    return fail(h, RB_ERR_ARG, "rb_set_solver_cluster: must be 0 (auto), 1, 2, 4 or 8");
  h->cluster_req = ctas_per_problem;
  h->planned = false;  // takes effect at the next rb_plan
  return RB_OK;
}

int rb_get_solver_cluster(const rb_handle* h) { return h ? h->cluster : 0; }
int rb_get_solver_max_clusters(const rb_handle* h) { return h ? h->max_clusters : 0; }

/* dev: enable per-CTA cycle counters of the solver and read them back (8 x int64 per CTA, up to 1024 CTAs) */
int rb_debug_solver_cycles(rb_handle* h, long long* out, int n_ctas) {
  if (!h) return RB_ERR_ARG;
  if (!h->d_debug) {
    RB_CUDA(h, cudaMalloc((void**)&h->d_debug, 1024 * 8 * sizeof(long long)));
    RB_CUDA(h, cudaMemset(h->d_debug, 0, 1024 * 8 * sizeof(long long)));
    return RB_OK;
  }
  if (out && n_ctas > 0) RB_CUDA(h, cudaMemcpy(out, h->d_debug, (size_t)(n_ctas > 1024 ? 1024 : n_ctas) * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  return RB_OK;
}

int rb_set_tile_skip(rb_handle* h, int32_t on) {
  if (!h) return RB_ERR_ARG;
  if (h->tile_skip != (on != 0)) {
    h->tile_skip = on != 0;
    h->affinity_done = false;  // the packed buffer of a previous rb_affinity is in the other format
    h->planned = false;        // the solver's cluster size is chosen per format: takes effect at the next rb_plan
  }
  return RB_OK;
}

int rb_edge_counts(rb_handle* h, int32_t* out, void* stream) {
  if (!h || !out) return RB_ERR_ARG;
  if (!h->planned || !h->affinity_done) return fail(h, RB_ERR_NOPLAN, "rb_edge_counts before rb_affinity");
  if (!h->edges_valid) return fail(h, RB_ERR_ARG, "rb_edge_counts: no edge lists (matrix given by rb_set_dense)");
  if (h->B > 0)
    RB_CUDA(h, cudaMemcpyAsync(out, h->d_edge_cnt, (size_t)h->B * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return RB_OK;
}

int rb_solver_resident(rb_handle* h, int32_t* out, void* stream) {
  if (!h || !out) return RB_ERR_ARG;
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_solver_resident before rb_plan");
  if (h->B > 0)
    RB_CUDA(h, cudaMemcpyAsync(out, h->d_resident, (size_t)h->B * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return RB_OK;
}

int rb_solver_streamed_passes(rb_handle* h, int32_t* out, void* stream) {
  if (!h || !out) return RB_ERR_ARG;
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_solver_streamed_passes before rb_plan");
  if (h->B > 0)
    RB_CUDA(h, cudaMemcpyAsync(out, h->d_streamed, (size_t)h->B * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return RB_OK;
}

int rb_set_solver_promote(rb_handle* h, int32_t passes) {
  if (!h) return RB_ERR_ARG;
  if (passes < 0) return fail(h, RB_ERR_ARG, "rb_set_solver_promote: passes must be >= 0 (0 = never)");
  h->promote_after = passes;
  return RB_OK;
}

int rb_stored_tiles(rb_handle* h, int32_t* out, void* stream) {
  if (!h || !out) return RB_ERR_ARG;
  if (!h->planned || !h->affinity_done) return fail(h, RB_ERR_NOPLAN, "rb_stored_tiles before rb_affinity / rb_set_dense");
  if (!h->tile_skip) return fail(h, RB_ERR_ARG, "rb_stored_tiles: the dense format stores every tile (rb_set_tile_skip is off)");
  RB_CUDA(h, rb_launch_count_flags(h->d_probs, h->B, h->d_flags, out, (cudaStream_t)stream));
  return RB_OK;
}

int rb_live_counts(rb_handle* h, int32_t* out, void* stream) {
  if (!h || !out) return RB_ERR_ARG;
  if (!h->planned || !h->affinity_done || !h->masks_valid) return fail(h, RB_ERR_NOPLAN, "rb_live_counts before rb_affinity");
  RB_ENTER(h);
  RB_CUDA(h, rb_launch_live_counts(h->d_probs, h->B, h->d_masks, out, (cudaStream_t)stream));
  return RB_OK;
}

int64_t rb_host_copy_bytes(const rb_handle* h, int32_t direction) {
  if (!h) return -1;
  return direction == 0 ? h->last_h2d : h->last_d2h;
}

int rb_set_profiling(rb_handle* h, int32_t on) {
  if (!h) return RB_ERR_ARG;
  h->profiling = on != 0;
  return RB_OK;
}

int rb_get_stage_ms(rb_handle* h, float ms[4]) {
  if (!h || !ms) return RB_ERR_ARG;
  static const int beg[4] = {0, 1, 3, 5}, end[4] = {1, 2, 4, 6};
  for (int i = 0; i < 4; ++i) {
    ms[i] = 0.f;
    if (h->ev_valid[beg[i]] && h->ev_valid[end[i]]) {
      RB_CUDA(h, cudaEventSynchronize(h->ev[end[i]]));
      RB_CUDA(h, cudaEventElapsedTime(&ms[i], h->ev[beg[i]], h->ev[end[i]]));
    }
  }
  return RB_OK;
}

static int plan_impl(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m,
                     const int64_t* off1, const int64_t* off2, int64_t tot1, int64_t tot2, const int32_t* share = nullptr);

int rb_plan(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m) {
  return plan_impl(h, B, d, n1, n2, m, nullptr, nullptr, 0, 0);
}

int rb_plan_shared(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m,
                   const int64_t* obj1_offset, const int64_t* obj2_offset, int64_t total_obj1, int64_t total_obj2,
                   const int32_t* share) {
  if (B > 0 && (!obj1_offset || !obj2_offset || !share)) return h ? fail(h, RB_ERR_ARG, "rb_plan_shared: null argument") : RB_ERR_ARG;
  return plan_impl(h, B, d, n1, n2, m, obj1_offset, obj2_offset, total_obj1, total_obj2, share);
}

int rb_plan_indexed(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m,
                    const int64_t* obj1_offset, const int64_t* obj2_offset, int64_t total_obj1, int64_t total_obj2) {
  if (B > 0 && (!obj1_offset || !obj2_offset)) return h ? fail(h, RB_ERR_ARG, "rb_plan_indexed: null offsets") : RB_ERR_ARG;
  return plan_impl(h, B, d, n1, n2, m, obj1_offset, obj2_offset, total_obj1, total_obj2);
}

static int plan_impl(rb_handle* h, int32_t B, int32_t d, const int32_t* n1, const int32_t* n2, const int32_t* m,
                     const int64_t* off1, const int64_t* off2, int64_t tot1, int64_t tot2, const int32_t* share) {
  if (!h) return RB_ERR_ARG;
  if (B < 0 || d < 1 || (B > 0 && (!n1 || !n2))) return fail(h, RB_ERR_ARG, "rb_plan: bad arguments");
  RB_CUDA(h, cudaSetDevice(h->device));
  h->planned = false;
  h->affinity_done = false;
  h->probs.assign((size_t)B, RbProblem());
  h->tile_prefix.assign((size_t)B + 1, 0);
  int64_t oa = 0, op = 0, ot = 0, o1 = 0, o2 = 0;
  int32_t max_mp = 0, mx1 = 0, mx2 = 0;
  for (int b = 0; b < B; ++b) {
    if (n1[b] < 0 || n2[b] < 0) return fail(h, RB_ERR_ARG, "rb_plan: negative object count");
    const int64_t mm = m ? (int64_t)m[b] : (int64_t)n1[b] * n2[b];
    if (mm < 0 || mm > (int64_t)1 << 24) return fail(h, RB_ERR_ARG, "rb_plan: association count out of range");
    RbProblem& p = h->probs[(size_t)b];
    p.n1 = n1[b]; p.n2 = n2[b]; p.m = (int32_t)mm; p.nb = (int32_t)((mm + RB_TILE - 1) / RB_TILE);
    p.d = d; p.all_to_all = m ? 0 : 1;
    p.off_obj1 = off1 ? off1[b] : o1; p.off_obj2 = off2 ? off2[b] : o2;
    if (off1 && (off1[b] < 0 || off1[b] + n1[b] > tot1 || off2[b] < 0 || off2[b] + n2[b] > tot2))
      return fail(h, RB_ERR_ARG, "rb_plan_indexed: object range outside the shared tables");
    p.off_assoc = oa; p.off_pad = op; p.off_tile = ot;
    p.n_tiles = (int64_t)p.nb * (p.nb + 1) / 2;
    p.base = b; p.is_copy = 0; p.off_rec = op;
    h->tile_prefix[(size_t)b] = ot;
    if (share && share[b] != b) {  // uses the matrix, records and edge list of an earlier problem of the same shape
      const int sb = share[b];
      if (sb < 0 || sb > b || share[sb] != sb) return fail(h, RB_ERR_ARG, "rb_plan_shared: share[b] must name an earlier problem that owns its matrix");
      const RbProblem& q = h->probs[(size_t)sb];
      if (q.n1 != p.n1 || q.n2 != p.n2 || q.m != p.m || q.off_obj1 != p.off_obj1 || q.off_obj2 != p.off_obj2)
        return fail(h, RB_ERR_ARG, "rb_plan_shared: problems that share a matrix must have identical shapes and object ranges");
      p.base = sb; p.is_copy = 1; p.off_rec = q.off_pad; p.off_tile = q.off_tile;
      o1 += n1[b]; o2 += n2[b]; oa += mm; op += (int64_t)p.nb * RB_TILE;   // private vectors and outputs, no tiles
      if (p.nb * RB_TILE > max_mp) max_mp = p.nb * RB_TILE;
      if (n1[b] > mx1) mx1 = n1[b];
      if (n2[b] > mx2) mx2 = n2[b];
      continue;
    }
    o1 += n1[b]; o2 += n2[b]; oa += mm; op += (int64_t)p.nb * RB_TILE; ot += p.n_tiles;
    if (p.nb * RB_TILE > max_mp) max_mp = p.nb * RB_TILE;
    if (n1[b] > mx1) mx1 = n1[b];
    if (n2[b] > mx2) mx2 = n2[b];
  }
  h->tile_prefix[(size_t)B] = ot;
  int optin = 0;
  RB_CUDA(h, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  int slots = 1, depth = 2;
  if (!rb_solver_config(max_mp, (size_t)optin, &slots, &depth)) {
    char buf[200];
    snprintf(buf, sizeof buf, "rb_plan: m (padded %d) does not fit the solver's shared-memory layout (%zu B needed at minimum, device offers %d)",
             max_mp, rb_solver_smem_bytes(max_mp, 1, 2), optin);
    return fail(h, RB_ERR_TOO_LARGE, buf);
  }
  h->slots = slots;
  h->depth = depth;
  h->B = B; h->d = d; h->all_to_all = (m == nullptr);
  h->total_assoc = oa; h->total_pad = op; h->total_tiles = ot;
  h->total_obj1 = off1 ? tot1 : o1; h->total_obj2 = off2 ? tot2 : o2;
  h->indexed = (off1 != nullptr);
  h->shared = (share != nullptr);
  h->max_mp = max_mp;
  h->max_n1 = mx1;
  h->max_n2 = mx2;
  {
    // CTAs per problem.  Auto: spread a problem over enough SMs that a slow-converging problem
    // cannot leave a long single-SM tail, but never more clusters' worth than there are problems.
    int c = h->cluster_req;
    if (c == 0) {
      // Measured on B200 (profiles/r1_cluster_sweep.txt): 2 CTAs per problem keeps a slow-converging
      // problem from leaving a long single-SM tail and still uses all 148 SMs (74 clusters);
      // larger clusters only pay when there are too few problems to fill the chip otherwise.
      c = 2;
      if ((int64_t)B * 4 <= h->sm_count) c = 4;
      if ((int64_t)B * 8 <= h->sm_count) c = 8;
      if (max_mp <= 256) c = 1;
      // With single-association scores (cosine / shape ratios) most associations cannot carry an edge, the
      // tile-skip format then usually fits the solver's resident (shared-memory CSR) mode, where one CTA per
      // problem is fastest: there is nothing left to split.
      if (h->tile_skip && h->inv.kind == RB_INVARIANT_ROMAN && (h->inv.cos_feature_dim > 0 || h->inv.ratio_feature_dim > 0)) c = 1;
    }
    int mc = 0;
    for (;;) {
      cudaError_t e = rb_solver_max_clusters(max_mp, c, &mc);
      if (e == cudaSuccess && mc >= 1) break;
      cudaGetLastError();
      if (c == 1) return fail(h, RB_ERR_CUDA, "rb_plan: solver kernel cannot be resident on this device");
      c >>= 1;
    }
    h->cluster = c;
    h->max_clusters = mc;
  }

  int rc;
  if ((rc = ensure(h, &h->d_probs, &h->cap_probs, (size_t)B))) return rc;
  if ((rc = ensure(h, &h->d_tile_prefix, &h->cap_prefix, (size_t)B + 1))) return rc;
  if ((rc = ensure(h, &h->d_M, &h->cap_M, (size_t)ot * RB_TILE_ELEMS))) return rc;
  if ((rc = ensure(h, &h->d_flags, &h->cap_flags, (size_t)ot))) return rc;
  {
    const size_t need = (size_t)op;
    if (need > h->cap_pad || !h->d_pad_f64) {
      if (h->d_pad_f64) cudaFree(h->d_pad_f64);
      if (h->d_pad_i32) cudaFree(h->d_pad_i32);
      h->d_pad_f64 = nullptr; h->d_pad_i32 = nullptr; h->cap_pad = 0;
      const size_t n = need ? need : 1;
      RB_CUDA(h, cudaMalloc((void**)&h->d_pad_f64, (size_t)kPadArrays * n * sizeof(double)));
      RB_CUDA(h, cudaMalloc((void**)&h->d_pad_i32, 2 * n * sizeof(int32_t)));
      if (h->d_masks) cudaFree(h->d_masks);
      h->d_masks = nullptr;
      RB_CUDA(h, cudaMalloc((void**)&h->d_masks, (n / RB_TILE + 1) * sizeof(uint32_t)));
      if (h->d_edge_ij) cudaFree(h->d_edge_ij);
      if (h->d_edge_v) cudaFree(h->d_edge_v);
      h->d_edge_ij = nullptr; h->d_edge_v = nullptr;
      RB_CUDA(h, cudaMalloc((void**)&h->d_edge_ij, (size_t)RB_EDGE_FACTOR * n * sizeof(unsigned)));
      RB_CUDA(h, cudaMalloc((void**)&h->d_edge_v, (size_t)RB_EDGE_FACTOR * n * sizeof(double)));
      h->cap_pad = n;
    }
  }
  {
    const size_t need = (size_t)oa;
    if (need > h->cap_assoc || !h->d_nodes) {
      if (h->d_nodes) cudaFree(h->d_nodes);
      if (h->d_sel) cudaFree(h->d_sel);
      h->d_nodes = nullptr; h->d_sel = nullptr; h->cap_assoc = 0;
      const size_t n = need ? need : 1;
      RB_CUDA(h, cudaMalloc((void**)&h->d_nodes, n * sizeof(int32_t)));
      RB_CUDA(h, cudaMalloc((void**)&h->d_sel, 2 * n * sizeof(int32_t)));
      h->cap_assoc = n;
    }
  }
  if ((size_t)B > h->cap_B || !h->d_nsel) {
    cudaFree(h->d_nsel); cudaFree(h->d_score); cudaFree(h->d_ifinal); cudaFree(h->d_npass); cudaFree(h->d_status); cudaFree(h->d_T);
    cudaFree(h->d_resident); h->d_resident = nullptr;
    cudaFree(h->d_streamed); h->d_streamed = nullptr;
    h->d_nsel = h->d_ifinal = h->d_npass = h->d_status = nullptr; h->d_score = h->d_T = nullptr; h->cap_B = 0;
    const size_t n = B ? (size_t)B : 1;
    RB_CUDA(h, cudaMalloc((void**)&h->d_nsel, n * sizeof(int32_t)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_score, n * sizeof(double)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_ifinal, n * sizeof(int32_t)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_npass, n * sizeof(int32_t)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_resident, n * sizeof(int32_t)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_streamed, n * sizeof(int32_t)));
    cudaFree(h->d_edge_cnt); h->d_edge_cnt = nullptr;
    RB_CUDA(h, cudaMalloc((void**)&h->d_edge_cnt, n * sizeof(int)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_status, n * sizeof(int32_t)));
    RB_CUDA(h, cudaMalloc((void**)&h->d_T, n * 16 * sizeof(double)));
    h->cap_B = n;
  }
  {
    // Problem ranges of the host entry point (rb_align_batch_host): the copy stream uploads range c+1 while the compute
    // streams run K1a/K1/K2/K3 of range c.  Big batches use 5 ranges growing from a small first one — the kernels cannot
    // start before the first range has arrived, and a range must not finish computing before the next one has been
    // copied (measured on B200, 592 pairs of m = 2000: 52.1k alignments/s against 49.2k with 4 equal ranges).  Smaller
    // batches: equal ranges of at least two solver waves.  RB_HOST_SPLIT="w0,w1,..." (dev) overrides the relative sizes.
    constexpr int RB_HOST_CHUNKS = 8;
    const int mc = h->max_clusters > 0 ? h->max_clusters : 1;
    int nch = B / (2 * mc);
    if (nch < 1) nch = 1;
    if (nch > 4) nch = 4;
    double wts[RB_HOST_CHUNKS] = {0.6, 1.0, 1.4, 1.4, 1.4, 0.0, 0.0, 0.0};
    int nw = (B >= 5 * mc) ? 5 : 0;
    if (const char* env = getenv("RB_HOST_SPLIT")) {
      nw = 0;
      const char* q = env;
      while (*q && nw < RB_HOST_CHUNKS) {
        char* endp = nullptr;
        const double v = strtod(q, &endp);
        if (endp == q) break;
        if (v > 0.0) wts[nw++] = v;
        q = (*endp == ',') ? endp + 1 : endp;
      }
    }
    if (nw >= 1 && B >= 2 * nw) nch = nw;
    else nw = 0;
    // ranges exist to overlap the upload with the kernels: with little to upload (shared object tables, small maps)
    // one range — one launch sequence, one solver tail — is faster
    const double h2d_bytes = 8.0 * ((double)(off1 ? tot1 : o1) + (double)(off2 ? tot2 : o2)) * d + 8.0 * (double)oa + (m ? 8.0 * (double)oa : 0.0);
    if ((!getenv("RB_HOST_SPLIT") && h2d_bytes < 24.0e6) || share) { nch = 1; nw = 0; }  // sharing groups must not straddle ranges
    // ranges are cut by WORK (tiles), not by problem count: ragged batches arrive longest-first
    double tot = 0.0, acc = 0.0;
    for (int c = 0; c < nch; ++c) tot += nw ? wts[c] : 1.0;
    h->bounds[0] = 0;
    int b = 0;
    for (int c = 0; c < nch; ++c) {
      acc += nw ? wts[c] : 1.0;
      const double target = (double)ot * acc / tot;
      while (b < B && (double)h->tile_prefix[(size_t)b + 1] <= target + 0.5) ++b;
      if (c + 1 == nch) b = B;
      h->bounds[c + 1] = b;
    }
    h->nch = nch;
    // queue orders: descending m (stable), for the whole batch and inside every range
    if ((rc = ensure(h, &h->d_order, &h->cap_order, (size_t)B))) return rc;
    if ((rc = ensure(h, &h->d_order_chunks, &h->cap_order_chunks, (size_t)B))) return rc;
    std::vector<int32_t> ord((size_t)B), ordc((size_t)B);
    auto fill = [&](int b0, int b1, int32_t* out) {
      for (int i = b0; i < b1; ++i) out[i] = i - b0;
      std::stable_sort(out + b0, out + b1, [&](int32_t x, int32_t y) { return h->probs[(size_t)(b0 + x)].m > h->probs[(size_t)(b0 + y)].m; });
    };
    fill(0, B, ord.data());
    for (int c = 0; c < nch; ++c) fill(h->bounds[c], h->bounds[c + 1], ordc.data());
    if (B > 0) {
      RB_CUDA(h, cudaMemcpy(h->d_order, ord.data(), (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice));
      RB_CUDA(h, cudaMemcpy(h->d_order_chunks, ordc.data(), (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    // compact read-back: an inlier set respects the distinctness constraint, so it has at most min(n1, n2, m) members
    // (problems that exceed the capacity anyway are read back in full, see rb_align_batch_host)
    int32_t kc = 1;
    for (int i = 0; i < B; ++i) {
      const RbProblem& p = h->probs[(size_t)i];
      int32_t k = p.n1 < p.n2 ? p.n1 : p.n2;
      if (k > p.m) k = p.m;
      if (k > kc) kc = k;
    }
    h->kcap = kc;
    if ((rc = ensure(h, &h->d_selpad, &h->cap_selpad, (size_t)(B ? B : 1) * kc * 2))) return rc;
  }
  // Descriptor upload is synchronous on the legacy stream: plan is not part of any timed region.
  if (B > 0) {
    RB_CUDA(h, cudaMemcpy(h->d_probs, h->probs.data(), (size_t)B * sizeof(RbProblem), cudaMemcpyHostToDevice));
  }
  RB_CUDA(h, cudaMemcpy(h->d_tile_prefix, h->tile_prefix.data(), ((size_t)B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
  RB_CUDA(h, cudaDeviceSynchronize());
  h->planned = true;
  return RB_OK;
}

int64_t rb_assoc_offset(const rb_handle* h, int32_t b) {
  if (!h || !h->planned || b < 0 || b > h->B) return -1;
  return b == h->B ? h->total_assoc : h->probs[(size_t)b].off_assoc;
}
int64_t rb_obj1_offset(const rb_handle* h, int32_t b) {
  if (!h || !h->planned || b < 0 || b > h->B) return -1;
  return b == h->B ? h->total_obj1 : h->probs[(size_t)b].off_obj1;
}
int64_t rb_obj2_offset(const rb_handle* h, int32_t b) {
  if (!h || !h->planned || b < 0 || b > h->B) return -1;
  return b == h->B ? h->total_obj2 : h->probs[(size_t)b].off_obj2;
}
int64_t rb_packed_bytes(const rb_handle* h) {
  if (!h || !h->planned) return -1;
  return h->total_tiles * (int64_t)RB_TILE_BYTES;
}

// ---- stage launches on a contiguous range [b0, b1) of the planned problems --------------------
// (the public stage calls use the whole batch; rb_align_batch_host pipelines sub-ranges)
namespace {

int check_affinity_args(rb_handle* h, const double* D1, const double* D2, const int32_t* A) {
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_affinity before rb_plan");
  if (h->B == 0) return RB_OK;
  if (!D1 || !D2) return fail(h, RB_ERR_ARG, "rb_affinity: null feature table");
  if (!h->all_to_all && !A) return fail(h, RB_ERR_ARG, "rb_affinity: plan has explicit associations but A is null");
  const rb_invariant_params& p = h->inv;
  if (p.kind != RB_INVARIANT_EUCLIDEAN) {
    if (h->d < p.point_dim + p.ratio_feature_dim + p.cos_feature_dim)
      return fail(h, RB_ERR_ARG, "rb_affinity: row length d smaller than point_dim + ratio_feature_dim + cos_feature_dim");
  } else if (h->d > 3) {
    return fail(h, RB_ERR_ARG, "rb_affinity: EuclideanDistance invariant supports d <= 3");
  }
  return RB_OK;
}

int affinity_range(rb_handle* h, int b0, int b1, const double* D1, const double* D2, const int32_t* A, cudaStream_t st) {
  if (b1 <= b0) return RB_OK;
  const RbDevParams P = dev_params(h);
  const RbAssocSoA rec = soa_of(h);
  record(h, 0, st);
  RB_CUDA(h, rb_launch_prep(h->d_probs + b0, b1 - b0, h->max_mp, h->max_n1, h->max_n2, h->all_to_all, rec, D1, D2,
                            h->all_to_all ? nullptr : A, P, st));
  record(h, 1, st);
  const int64_t base = h->tile_prefix[(size_t)b0];
  RB_CUDA(h, rb_launch_affinity(h->d_probs + b0, b1 - b0, base, h->tile_prefix[(size_t)b1] - base, h->max_mp / RB_TILE,
                                rec, h->d_masks, h->d_M, h->tile_skip ? h->d_flags : nullptr, h->d_edge_ij, h->d_edge_v,
                                h->d_edge_cnt + b0, P, h->sm_count, st));
  record(h, 2, st);
  h->edges_valid = true;  // K1 lists the edges in both formats
  h->masks_valid = true;
  return RB_OK;
}

// Rounding::DSD (clipper.cpp:364-371): the kernel left S = support of u in `nodes` (NONZERO rounding); per problem, pull
// the S x S block of M, run Goldberg's exact densest subgraph on the host (rb_dsd.cu) and write the result back.
// Synchronises `st`.
int dsd_epilogue(rb_handle* h, int b0, int b1, int32_t* nodes, int32_t* sel, int32_t* n_nodes, cudaStream_t st) {
  const size_t nb = (size_t)(b1 - b0);
  const int64_t a0 = h->probs[(size_t)b0].off_assoc;
  const int64_t a1 = b1 < h->B ? h->probs[(size_t)b1].off_assoc : h->total_assoc;
  std::vector<int32_t> hn(nb), hnodes((size_t)(a1 - a0)), out;
  std::vector<double> W;
  RB_CUDA(h, cudaMemcpyAsync(hn.data(), n_nodes, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (a1 > a0)
    RB_CUDA(h, cudaMemcpyAsync(hnodes.data(), nodes + a0, (size_t)(a1 - a0) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RB_CUDA(h, cudaStreamSynchronize(st));
  for (size_t q = 0; q < nb; ++q) {
    const RbProblem& pr = h->probs[(size_t)b0 + q];
    const int k = hn[q];
    if (k <= 0) continue;
    int rc;
    if ((rc = ensure(h, &h->d_W, &h->cap_W, (size_t)k * k))) return rc;
    RB_CUDA(h, rb_launch_gather_sub(pr, h->d_M, h->tile_skip ? h->d_flags : nullptr, nodes + pr.off_assoc, k, h->d_W, st));
    W.resize((size_t)k * k);
    RB_CUDA(h, cudaMemcpyAsync(W.data(), h->d_W, W.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(h, cudaStreamSynchronize(st));
    out.assign((size_t)k, 0);
    int32_t n_out = 0;
    if (rb_dsd_host(pr.m, k, hnodes.data() + (pr.off_assoc - a0), W.data(), out.data(), &n_out) != RB_OK)
      return fail(h, RB_ERR_ARG, "rb_solve: densest-subgraph rounding failed");
    hn[q] = n_out;
    if (n_out > 0) {
      RB_CUDA(h, cudaMemcpyAsync(nodes + pr.off_assoc, out.data(), (size_t)n_out * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      RB_CUDA(h, rb_launch_select(nodes + pr.off_assoc, n_out, h->d_pad_i32 + pr.off_rec,
                                  h->d_pad_i32 + (size_t)h->total_pad + pr.off_rec, sel + 2 * pr.off_assoc, st));
      RB_CUDA(h, cudaStreamSynchronize(st));  // `out` is reused by the next problem
    }
  }
  RB_CUDA(h, cudaMemcpyAsync(n_nodes, hn.data(), nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  RB_CUDA(h, cudaStreamSynchronize(st));
  return RB_OK;
}

int solve_range(rb_handle* h, int b0, int b1, int queue, const int32_t* order, const double* u0, int32_t* nodes, int32_t* sel,
                int32_t* n_nodes, double* u, double* score, int32_t* ifinal, int32_t* n_passes, cudaStream_t st) {
  if (b1 <= b0) return RB_OK;
  RbSolveArgs a;
  a.probs = h->d_probs + b0; a.B = b1 - b0; a.b0 = b0; a.counter = h->d_counter + queue; a.Mp = h->d_M;
  a.order = order;
  a.flags = h->tile_skip ? h->d_flags : nullptr;
  a.edge_ij = h->edges_valid ? h->d_edge_ij : nullptr;
  a.edge_v = h->d_edge_v;
  a.edge_cnt = h->edges_valid ? h->d_edge_cnt + b0 : nullptr;
  a.nt_max = (h->max_mp / RB_TILE) * (h->max_mp / RB_TILE + 1) / 2;
  a.i1 = h->d_pad_i32; a.i2 = h->d_pad_i32 + (size_t)h->total_pad;
  a.u0 = u0; a.u_work = u_work(h); a.g_work = g_work(h);
  a.nodes = nodes ? nodes : h->d_nodes;            // per-association arrays are addressed by absolute offsets
  a.sel = sel ? sel : h->d_sel;
  a.n_nodes = (n_nodes ? n_nodes : h->d_nsel) + b0;  // per-problem arrays are indexed relative to the range
  a.u_out = u;
  a.score = (score ? score : h->d_score) + b0;
  a.ifinal = (ifinal ? ifinal : h->d_ifinal) + b0;
  a.n_passes = (n_passes ? n_passes : h->d_npass) + b0;
  a.resident = h->d_resident + b0;
  a.streamed = h->d_streamed + b0;
  a.promote_after = h->tile_skip ? 0 : h->promote_after;
  a.sp = h->sol;
  const bool dsd = (h->sol.rounding == 1);
  if (dsd) a.sp.rounding = 0;  // the kernel leaves the support of u; the exact rounding follows on the host
  a.mp_max = h->max_mp;
  a.pad_stride = h->total_pad;
  a.cluster = h->cluster;
  a.debug = h->d_debug;
  a.slots = h->slots;
  a.depth = h->depth;
  RB_CUDA(h, cudaMemsetAsync(h->d_counter + queue, 0, sizeof(int32_t), st));
  record(h, 3, st);
  RB_CUDA(h, rb_launch_solver(a, h->max_clusters, st));
  record(h, 4, st);
  if (dsd) return dsd_epilogue(h, b0, b1, a.nodes, a.sel, a.n_nodes, st);
  return RB_OK;
}

int arun_range(rb_handle* h, int b0, int b1, const double* D1, const double* D2, const int32_t* sel,
               const int32_t* n_sel, double* T, int32_t* status, bool pack_sel, cudaStream_t st) {
  if (b1 <= b0) return RB_OK;
  const int pd = h->inv.kind != RB_INVARIANT_EUCLIDEAN ? h->inv.point_dim : (h->d < 3 ? h->d : 3);
  if (pd < 2) return fail(h, RB_ERR_ARG, "rb_arun: point dimension must be 2 or 3");
  record(h, 5, st);
  RB_CUDA(h, rb_launch_arun(h->d_probs + b0, b1 - b0, pd, D1, D2, sel, n_sel + b0, (T ? T : h->d_T) + (size_t)b0 * 16,
                            (status ? status : h->d_status) + b0,
                            pack_sel ? h->d_selpad + (size_t)b0 * h->kcap * 2 : nullptr, h->kcap, st));
  record(h, 6, st);
  return RB_OK;
}

}  // namespace

int rb_affinity(rb_handle* h, const double* D1, const double* D2, const int32_t* A, void* stream) {
  if (!h) return RB_ERR_ARG;
  int rc = check_affinity_args(h, D1, D2, A);
  if (rc) return rc;
  RB_ENTER(h);
  cudaStream_t st = (cudaStream_t)stream;
  h->last_stream = st; h->last_valid = true;
  rc = affinity_range(h, 0, h->B, D1, D2, A, st);
  if (rc) return rc;
  h->affinity_done = true;
  return RB_OK;
}

int rb_solve(rb_handle* h, const double* u0, int32_t* nodes, int32_t* sel, int32_t* n_nodes, double* u,
             double* score, int32_t* ifinal, int32_t* n_passes, void* stream) {
  if (!h) return RB_ERR_ARG;
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_solve before rb_plan");
  if (!h->affinity_done) return fail(h, RB_ERR_NOPLAN, "rb_solve before rb_affinity / rb_set_dense");
  if (h->B == 0) return RB_OK;
  if (!u0 && h->total_assoc > 0) return fail(h, RB_ERR_ARG, "rb_solve: u0 is required");
  RB_ENTER(h);
  cudaStream_t st = (cudaStream_t)stream;
  h->last_stream = st; h->last_valid = true;
  return solve_range(h, 0, h->B, 0, h->d_order, u0, nodes, sel, n_nodes, u, score, ifinal, n_passes, st);
}

int rb_arun(rb_handle* h, const double* D1, const double* D2, const int32_t* sel, const int32_t* n_sel, double* T,
            int32_t* status, void* stream) {
  if (!h) return RB_ERR_ARG;
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_arun before rb_plan");
  if (h->B == 0) return RB_OK;
  if (!D1 || !D2 || !sel || !n_sel) return fail(h, RB_ERR_ARG, "rb_arun: null input");
  RB_ENTER(h);
  cudaStream_t st = (cudaStream_t)stream;
  h->last_stream = st; h->last_valid = true;
  return arun_range(h, 0, h->B, D1, D2, sel, n_sel, T, status, false, st);
}

int rb_submap_similarity(rb_handle* h, int32_t S0, int32_t S1, int32_t D, const int64_t* off0, const int64_t* off1,
                         const double* F0, const double* F1, double* sim, void* stream) {
  if (!h) return RB_ERR_ARG;
  if (S0 < 0 || S1 < 0 || D < 1) return fail(h, RB_ERR_ARG, "rb_submap_similarity: bad sizes");
  if (S0 > 65535) return fail(h, RB_ERR_TOO_LARGE, "rb_submap_similarity: more than 65535 submaps in run 0");
  if (S0 == 0 || S1 == 0) return RB_OK;
  if (!off0 || !off1 || !F0 || !F1 || !sim) return fail(h, RB_ERR_ARG, "rb_submap_similarity: null pointer");
  RB_ENTER(h);
  cudaError_t e = rb_launch_similarity(S0, S1, D, off0, off1, F0, F1, sim, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(h, RB_ERR_CUDA, std::string("rb_submap_similarity: ") + cudaGetErrorString(e));
  return RB_OK;
}

int rb_align_batch(rb_handle* h, const double* D1, const double* D2, const int32_t* A, const double* u0,
                   int32_t* sel, int32_t* n_sel, double* T, int32_t* status, double* score, int32_t* n_passes,
                   void* stream) {
  if (!h) return RB_ERR_ARG;
  int rc = rb_affinity(h, D1, D2, A, stream);
  if (rc) return rc;
  int32_t* sel_p = sel ? sel : h->d_sel;
  int32_t* nsel_p = n_sel ? n_sel : h->d_nsel;
  rc = rb_solve(h, u0, nullptr, sel_p, nsel_p, nullptr, score, nullptr, n_passes, stream);
  if (rc) return rc;
  return rb_arun(h, D1, D2, sel_p, nsel_p, T, status, stream);
}

int rb_get_dense(rb_handle* h, int32_t b, double* M_out, double* C_out, void* stream) {
  if (!h) return RB_ERR_ARG;
  if (!h->planned || !h->affinity_done) return fail(h, RB_ERR_NOPLAN, "rb_get_dense before rb_affinity / rb_set_dense");
  if (b < 0 || b >= h->B) return fail(h, RB_ERR_ARG, "rb_get_dense: problem index out of range");
  RB_ENTER(h);
  RB_CUDA(h, rb_launch_unpack(h->probs[(size_t)b], h->d_M, h->tile_skip ? h->d_flags : nullptr, M_out, C_out,
                              (cudaStream_t)stream));
  return RB_OK;
}

int rb_set_dense(rb_handle* h, int32_t b, const double* M, const double* C, void* stream) {
  if (!h) return RB_ERR_ARG;
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_set_dense before rb_plan");
  if (b < 0 || b >= h->B || !M) return fail(h, RB_ERR_ARG, "rb_set_dense: bad arguments");
  RB_ENTER(h);
  cudaStream_t st = (cudaStream_t)stream;
  const RbProblem& pr = h->probs[(size_t)b];
  if (pr.is_copy) return fail(h, RB_ERR_ARG, "rb_set_dense: problem shares another problem's matrix (rb_plan_shared); set the owner");
  RB_CUDA(h, cudaMemsetAsync(h->d_bad, 0, sizeof(int), st));
  // set_matrix_data carries no association table: nodes map to themselves (sel = (node, 0))
  const size_t n = (size_t)pr.nb * RB_TILE;
  RB_CUDA(h, cudaMemsetAsync(h->d_pad_i32 + pr.off_pad, 0, n * sizeof(int32_t), st));
  RB_CUDA(h, cudaMemsetAsync(h->d_pad_i32 + (size_t)h->total_pad + pr.off_pad, 0, n * sizeof(int32_t), st));
  RB_CUDA(h, rb_launch_pack(pr, M, C, h->d_M, h->d_bad, st));
  RB_CUDA(h, cudaMemsetAsync(h->d_flags + pr.off_tile, 1, (size_t)pr.n_tiles, st));  // every tile of b is stored
  h->edges_valid = false;  // no edge list for matrices given by the caller: the solver reads the tiles
  h->masks_valid = false;
  int bad = 0;
  RB_CUDA(h, cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  RB_CUDA(h, cudaStreamSynchronize(st));
  if (bad) return fail(h, RB_ERR_ARG, "rb_set_dense: M has a non-zero entry where C is zero (unsupported)");
  h->affinity_done = true;
  return RB_OK;
}

int rb_align_batch_host(rb_handle* h, const double* D1, const double* D2, const int32_t* A, const double* u0,
                        int32_t* sel, int32_t* n_sel, double* T, int32_t* status, double* score,
                        int32_t* n_passes) {
  if (!h) return RB_ERR_ARG;
  if (!h->planned) return fail(h, RB_ERR_NOPLAN, "rb_align_batch_host before rb_plan");
  if (h->B == 0) return RB_OK;
  RB_CUDA(h, cudaSetDevice(h->device));
  if (!D1 || !D2 || !u0) return fail(h, RB_ERR_ARG, "rb_align_batch_host: null input");
  if (!h->all_to_all && !A) return fail(h, RB_ERR_ARG, "rb_align_batch_host: A required");
  int rc = check_affinity_args(h, D1, D2, A);
  if (rc) return rc;
  if (!h->own_stream) RB_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  if (!h->copy_stream) RB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  if (!h->own_stream2) RB_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream2, cudaStreamNonBlocking));
  if (!h->join_ev) RB_CUDA(h, cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
  cudaStream_t st = h->own_stream, cs = h->copy_stream;
  cudaStream_t cst[2] = {h->own_stream, h->own_stream2};
  if (h->last_valid && h->last_stream != st) {  // scratch may still be in use by work queued on the caller's stream
    if (!h->order_ev) RB_CUDA(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
    RB_CUDA(h, cudaEventRecord(h->order_ev, h->last_stream));
    RB_CUDA(h, cudaStreamWaitEvent(st, h->order_ev, 0));
    RB_CUDA(h, cudaStreamWaitEvent(cst[1], h->order_ev, 0));
    RB_CUDA(h, cudaStreamWaitEvent(cs, h->order_ev, 0));
  }
  const size_t nD1 = (size_t)h->total_obj1 * h->d, nD2 = (size_t)h->total_obj2 * h->d, nA = (size_t)h->total_assoc;
  if ((rc = ensure(h, &h->s_D1, &h->cap_sD1, nD1))) return rc;
  if ((rc = ensure(h, &h->s_D2, &h->cap_sD2, nD2))) return rc;
  if ((rc = ensure(h, &h->s_u0, &h->cap_su0, nA))) return rc;
  if (!h->all_to_all && (rc = ensure(h, &h->s_A, &h->cap_sA, 2 * nA))) return rc;

  // Pipeline: the batch is cut into contiguous problem ranges (fixed by rb_plan); the copy stream uploads range c+1
  // while the compute streams run K1a/K1/K2/K3 of range c (inputs are consumed in place, no reordering).
  const int nch = h->nch;
  const int* bounds = h->bounds;
  for (int c = 0; c < nch; ++c)
    if (!h->chunk_ev[c]) RB_CUDA(h, cudaEventCreateWithFlags(&h->chunk_ev[c], cudaEventDisableTiming));
  if (h->last_valid) {  // the copy stream must not overwrite staging buffers the previous call's kernels still read
    if (!h->order_ev) RB_CUDA(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
    RB_CUDA(h, cudaEventRecord(h->order_ev, st));
    RB_CUDA(h, cudaStreamWaitEvent(cs, h->order_ev, 0));
  }
  const size_t d = (size_t)h->d;
  h->last_h2d = (int64_t)((nD1 + nD2 + nA) * sizeof(double) + (h->all_to_all ? 0 : 2 * nA * sizeof(int32_t)));
  if (h->indexed) {  // shared object tables: uploaded once, every range may reference any part of them
    RB_CUDA(h, cudaMemcpyAsync(h->s_D1, D1, nD1 * sizeof(double), cudaMemcpyHostToDevice, cs));
    RB_CUDA(h, cudaMemcpyAsync(h->s_D2, D2, nD2 * sizeof(double), cudaMemcpyHostToDevice, cs));
  }
  for (int c = 0; c < nch; ++c) {
    const int b0 = bounds[c], b1 = bounds[c + 1];
    const RbProblem& p0 = h->probs[(size_t)b0];
    const size_t o1 = (size_t)p0.off_obj1 * d, o2 = (size_t)p0.off_obj2 * d, oa = (size_t)p0.off_assoc;
    const size_t e1 = (b1 < h->B ? (size_t)h->probs[(size_t)b1].off_obj1 : (size_t)h->total_obj1) * d;
    const size_t e2 = (b1 < h->B ? (size_t)h->probs[(size_t)b1].off_obj2 : (size_t)h->total_obj2) * d;
    const size_t ea = b1 < h->B ? (size_t)h->probs[(size_t)b1].off_assoc : (size_t)h->total_assoc;
    if (!h->indexed) {
      RB_CUDA(h, cudaMemcpyAsync(h->s_D1 + o1, D1 + o1, (e1 - o1) * sizeof(double), cudaMemcpyHostToDevice, cs));
      RB_CUDA(h, cudaMemcpyAsync(h->s_D2 + o2, D2 + o2, (e2 - o2) * sizeof(double), cudaMemcpyHostToDevice, cs));
    }
    RB_CUDA(h, cudaMemcpyAsync(h->s_u0 + oa, u0 + oa, (ea - oa) * sizeof(double), cudaMemcpyHostToDevice, cs));
    if (!h->all_to_all)
      RB_CUDA(h, cudaMemcpyAsync(h->s_A + 2 * oa, A + 2 * oa, 2 * (ea - oa) * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    RB_CUDA(h, cudaEventRecord(h->chunk_ev[c], cs));
    cudaStream_t ks = cst[c & 1];
    RB_CUDA(h, cudaStreamWaitEvent(ks, h->chunk_ev[c], 0));
    if ((rc = affinity_range(h, b0, b1, h->s_D1, h->s_D2, h->all_to_all ? nullptr : h->s_A, ks))) return rc;
    if ((rc = solve_range(h, b0, b1, c, h->d_order_chunks + b0, h->s_u0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ks))) return rc;
    if ((rc = arun_range(h, b0, b1, h->s_D1, h->s_D2, h->d_sel, h->d_nsel, nullptr, nullptr, sel != nullptr, ks))) return rc;
  }
  if (nch > 1) {  // join the second compute stream before reading results back
    RB_CUDA(h, cudaEventRecord(h->join_ev, cst[1]));
    RB_CUDA(h, cudaStreamWaitEvent(st, h->join_ev, 0));
  }
  h->affinity_done = true;
  h->last_stream = st; h->last_valid = true;
  const size_t B = (size_t)h->B;
  // `sel` comes back compactly: K3 packed the first kcap rows of every problem (kcap = max min(n1, n2), an upper bound
  // on any inlier set that respects the distinctness constraint) into d_selpad; the host scatters n_sel[b] rows into
  // the caller's layout.  (Copying the whole padded (sum m) x 2 table would be 9.5 MB for 592 pairs of m = 2000.)
  const size_t kc = (size_t)h->kcap;
  if (sel) {
    if (B * kc * 2 > h->cap_hsel) {
      if (h->h_selpad) cudaFreeHost(h->h_selpad);
      h->h_selpad = nullptr; h->cap_hsel = 0;
      RB_CUDA(h, cudaMallocHost((void**)&h->h_selpad, B * kc * 2 * sizeof(int32_t)));
      h->cap_hsel = B * kc * 2;
    }
    if (B > h->cap_hnsel) {
      if (h->h_nsel) cudaFreeHost(h->h_nsel);
      h->h_nsel = nullptr; h->cap_hnsel = 0;
      RB_CUDA(h, cudaMallocHost((void**)&h->h_nsel, B * sizeof(int32_t)));
      h->cap_hnsel = B;
    }
    RB_CUDA(h, cudaMemcpyAsync(h->h_selpad, h->d_selpad, B * kc * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RB_CUDA(h, cudaMemcpyAsync(h->h_nsel, h->d_nsel, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  if (n_sel) RB_CUDA(h, cudaMemcpyAsync(n_sel, h->d_nsel, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (T) RB_CUDA(h, cudaMemcpyAsync(T, h->d_T, B * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status) RB_CUDA(h, cudaMemcpyAsync(status, h->d_status, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (score) RB_CUDA(h, cudaMemcpyAsync(score, h->d_score, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (n_passes) RB_CUDA(h, cudaMemcpyAsync(n_passes, h->d_npass, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RB_CUDA(h, cudaStreamSynchronize(st));
  h->last_d2h = (int64_t)((sel ? B * kc * 2 * sizeof(int32_t) + B * sizeof(int32_t) : 0) + (n_sel ? B * 4 : 0) + (T ? B * 128 : 0) +
                          (status ? B * 4 : 0) + (score ? B * 8 : 0) + (n_passes ? B * 4 : 0));
  if (sel) {
    bool big = false;
    for (size_t b = 0; b < B; ++b) {
      const size_t k = (size_t)h->h_nsel[b];
      const size_t o = (size_t)h->probs[b].off_assoc;
      if (k > kc) {  // larger than any distinct inlier set: read this problem's rows straight from the full table
        RB_CUDA(h, cudaMemcpyAsync(sel + 2 * o, h->d_sel + 2 * o, 2 * k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        h->last_d2h += (int64_t)(2 * k * sizeof(int32_t));
        big = true;
      } else if (k > 0) {
        memcpy(sel + 2 * o, h->h_selpad + b * kc * 2, 2 * k * sizeof(int32_t));
      }
    }
    if (big) RB_CUDA(h, cudaStreamSynchronize(st));
  }
  return RB_OK;
}

}  // extern "C"

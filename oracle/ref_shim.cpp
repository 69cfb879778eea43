// ref_shim.cpp — C entry points around the REFERENCE's own CLIPPER classes, for oracle/_ref/libclipper_ref.so.
//
// TEST INFRASTRUCTURE.  This file is ours; everything it calls is the reference's unmodified C++
// (dependencies/clipper/src/{clipper,utils,dsd}.cpp, src/invariants/{roman,euclidean_distance}.cpp), compiled where it
// lies under /root/reference against the Eigen stand-in of oracle/eigen_standin (recipe: oracle/Makefile, `make ref`).
// The two solvers of clipper.h that need un-vendored third parties (PMC max-clique, SCS SDP) are stubbed: nothing on
// the roman.align path calls them.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "clipper/clipper.h"
#include "clipper/utils.h"

namespace clipper {
namespace maxclique {
std::vector<int> solve(const Eigen::MatrixXd&, const Params&) { throw std::runtime_error("maxclique (PMC) is not vendored"); }
}  // namespace maxclique
namespace sdp {
Solution solve(const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Params&) { throw std::runtime_error("sdp (SCS) is not vendored"); }
}  // namespace sdp
}  // namespace clipper

extern "C" {

// same POD layouts as oracle/clipper_oracle.c (orc_invariant_params / orc_solver_params)
#define REF_MAX_RATIO 8
typedef struct {
  int32_t kind, point_dim, ratio_feature_dim, cos_feature_dim, fusion_method, gravity_guided, drift_aware, drift_scale_sigma;
  double sigma, epsilon, mindist, distance_weight, ratio_weight, cosine_weight;
  double ratio_epsilon[REF_MAX_RATIO];
  double cosine_min, cosine_max, gravity_unc_ang_rad, drift_scale;
} ref_invariant_params;

typedef struct {
  double tol_u, tol_F, tol_Fop;
  int32_t maxiniters, maxoliters;
  double beta;
  int32_t maxlsiters;
  double eps, affinityeps;
  int32_t rescale_u0, rounding;
} ref_solver_params;

int orc_arun(int k, const double* pts1, const double* pts2, double* T);  // oracle/clipper_oracle.c (numpy T_align restated)
}

namespace {

clipper::invariants::ROMAN::Params roman_params(const ref_invariant_params* p) {
  clipper::invariants::ROMAN::Params q;
  q.point_dim = (uint32_t)p->point_dim;
  q.ratio_feature_dim = (uint32_t)p->ratio_feature_dim;
  q.cos_feature_dim = (uint32_t)p->cos_feature_dim;
  q.fusion_method = (clipper::invariants::ROMAN::SimilarityFusionMethod)p->fusion_method;
  q.sigma = p->sigma; q.epsilon = p->epsilon; q.mindist = p->mindist;
  q.distance_weight = p->distance_weight; q.ratio_weight = p->ratio_weight; q.cosine_weight = p->cosine_weight;
  q.ratio_epsilon = Eigen::VectorXd::Zero(p->ratio_feature_dim);
  for (int i = 0; i < p->ratio_feature_dim; ++i) q.ratio_epsilon(i) = p->ratio_epsilon[i];
  q.cosine_min = p->cosine_min; q.cosine_max = p->cosine_max;
  q.gravity_guided = p->gravity_guided != 0;
  q.gravity_unc_ang_rad = p->gravity_unc_ang_rad;
  q.drift_aware = p->drift_aware != 0; q.drift_scale_sigma = p->drift_scale_sigma != 0; q.drift_scale = p->drift_scale;
  return q;
}

clipper::Params solver_params(const ref_solver_params* s) {
  clipper::Params q;
  q.tol_u = s->tol_u; q.tol_F = s->tol_F; q.tol_Fop = s->tol_Fop;
  q.maxiniters = s->maxiniters; q.maxoliters = s->maxoliters;
  q.beta = s->beta; q.maxlsiters = s->maxlsiters; q.eps = s->eps; q.affinityeps = s->affinityeps;
  q.rescale_u0 = s->rescale_u0 != 0;
  q.rounding = (clipper::Params::Rounding)s->rounding;
  return q;
}

// object-major rows of d doubles == the memory of the (d x n) column-major Eigen matrix (py_clipper.cpp:316-318)
clipper::invariants::Data data_of(const double* D, int d, int n) {
  clipper::invariants::Data M(d, n);
  if (d > 0 && n > 0) std::memcpy(M.data(), D, sizeof(double) * (size_t)d * (size_t)n);
  return M;
}

clipper::Association assoc_of(const int32_t* A, int64_t m) {
  if (!A) return clipper::Association();  // empty => all-to-all (clipper.cpp:32, :81)
  clipper::Association a(m, 2);
  for (int64_t i = 0; i < m; ++i) { a(i, 0) = A[2 * i]; a(i, 1) = A[2 * i + 1]; }
  return a;
}

// kind 0 = ROMAN through CLIPPERPairwiseAndSingle::scorePairwiseAndSingleConsistency, 1 = EuclideanDistance through
// CLIPPER::scorePairwiseConsistency, 2 = ROMAN through CLIPPER::scorePairwiseConsistency (pairwise part only)
std::unique_ptr<clipper::CLIPPER> scored(const ref_invariant_params* ip, const ref_solver_params* sp, int d, int n1,
                                         const double* D1, int n2, const double* D2, int64_t m, const int32_t* A,
                                         int parallelize) {
  const clipper::Params P = solver_params(sp);
  const clipper::invariants::Data d1 = data_of(D1, d, n1), d2 = data_of(D2, d, n2);
  const clipper::Association a = assoc_of(A, m);
  if (ip->kind == 0) {
    auto inv = std::make_shared<clipper::invariants::ROMAN>(roman_params(ip));
    auto* c = new clipper::CLIPPERPairwiseAndSingle(inv, P);
    c->setParallelize(parallelize != 0);
    c->scorePairwiseAndSingleConsistency(d1, d2, a);
    return std::unique_ptr<clipper::CLIPPER>(c);
  }
  clipper::invariants::PairwiseInvariantPtr inv;
  if (ip->kind == 1) {
    clipper::invariants::EuclideanDistance::Params e;
    e.sigma = ip->sigma; e.epsilon = ip->epsilon; e.mindist = ip->mindist;
    inv = std::make_shared<clipper::invariants::EuclideanDistance>(e);
  } else {
    inv = std::make_shared<clipper::invariants::ROMAN>(roman_params(ip));
  }
  std::unique_ptr<clipper::CLIPPER> c(new clipper::CLIPPER(inv, P));
  c->setParallelize(parallelize != 0);
  c->scorePairwiseConsistency(d1, d2, a);
  return c;
}

void copy_solution(clipper::CLIPPER& c, int32_t* sel, int32_t* n_sel, double* score, int32_t* ifinal, double* u_out,
                   int32_t* nodes_out) {
  const clipper::Solution& s = c.getSolution();
  // (after setMatrixData there is no association table: getSelectedAssociations() would index an empty A_)
  const clipper::Association Ain = sel ? c.getSelectedAssociations() : clipper::Association();
  *n_sel = (int32_t)s.nodes.size();
  for (size_t i = 0; i < s.nodes.size(); ++i) {
    if (sel) { sel[2 * i] = Ain((Eigen::Index)i, 0); sel[2 * i + 1] = Ain((Eigen::Index)i, 1); }
    if (nodes_out) nodes_out[i] = s.nodes[i];
  }
  if (score) *score = s.score;
  if (ifinal) *ifinal = s.ifinal;
  if (u_out) for (Eigen::Index i = 0; i < s.u.size(); ++i) u_out[i] = s.u(i);
}

}  // namespace

extern "C" {

// scorePairwise[AndSingle]Consistency -> getAffinityMatrix / getConstraintMatrix (dense m x m, ROW-major out, M + I)
int ref_score(const ref_invariant_params* ip, const ref_solver_params* sp, int d, int n1, const double* D1, int n2,
              const double* D2, int64_t m, const int32_t* A, int parallelize, double* M_out, double* C_out) {
  try {
    auto c = scored(ip, sp, d, n1, D1, n2, D2, m, A, parallelize);
    const clipper::Affinity M = c->getAffinityMatrix();
    const clipper::Constraint C = c->getConstraintMatrix();
    const Eigen::Index mm = M.rows();
    for (Eigen::Index i = 0; i < mm; ++i)
      for (Eigen::Index j = 0; j < mm; ++j) {
        if (M_out) M_out[i * mm + j] = M(i, j);
        if (C_out) C_out[i * mm + j] = C(i, j);
      }
    return 0;
  } catch (...) { return -1; }
}

// one ObjectRegistration.register + T_align (object_registration.py:22-29, 88-129): score -> solve(u0) -> selected
// associations -> Arun.  Returns 0, 1 (fewer than 3 inliers, T = NaN) or -1.
int ref_align_pair(const ref_invariant_params* ip, const ref_solver_params* sp, int d, int n1, const double* D1, int n2,
                   const double* D2, int64_t m, const int32_t* A, const double* u0, int parallelize, int32_t* sel,
                   int32_t* n_sel, double* T, double* score, int32_t* ifinal, double* u_out, int32_t* nodes_out) {
  try {
    auto c = scored(ip, sp, d, n1, D1, n2, D2, m, A, parallelize);
    const int64_t mm = A ? m : (int64_t)n1 * n2;
    Eigen::VectorXd u(mm);
    for (int64_t i = 0; i < mm; ++i) u(i) = u0[i];
    c->solve(u);
    copy_solution(*c, sel, n_sel, score, ifinal, u_out, nodes_out);
    if (!T) return 0;
    const int k = *n_sel;
    if (k < 3) { for (int i = 0; i < 16; ++i) T[i] = NAN; return 1; }
    std::vector<double> p1((size_t)3 * k), p2((size_t)3 * k);
    for (int i = 0; i < k; ++i)
      for (int a = 0; a < 3; ++a) { p1[3 * i + a] = D1[(int64_t)sel[2 * i] * d + a]; p2[3 * i + a] = D2[(int64_t)sel[2 * i + 1] * d + a]; }
    orc_arun(k, p1.data(), p2.data(), T);
    return 0;
  } catch (...) { return -1; }
}

// setMatrixData(M, C) + solve(u0) on dense ROW-major symmetric inputs (C may be null: C = pattern of M + I)
int ref_solve_dense(const ref_solver_params* sp, int64_t m, const double* M, const double* C, const double* u0,
                    int32_t* nodes_out, int32_t* n_nodes, double* score, int32_t* ifinal, double* u_out) {
  try {
    clipper::CLIPPER c(nullptr, solver_params(sp));
    clipper::Affinity Mm(m, m);
    clipper::Constraint Cm(m, m);
    for (int64_t i = 0; i < m; ++i)
      for (int64_t j = 0; j < m; ++j) {
        Mm(i, j) = M[i * m + j];
        Cm(i, j) = C ? C[i * m + j] : ((M[i * m + j] != 0.0 || i == j) ? 1.0 : 0.0);
      }
    c.setMatrixData(Mm, Cm);
    Eigen::VectorXd u(m);
    for (int64_t i = 0; i < m; ++i) u(i) = u0[i];
    c.solve(u);
    copy_solution(c, nullptr, n_nodes, score, ifinal, u_out, nodes_out);
    return 0;
  } catch (...) { return -1; }
}

double ref_single_similarity(const ref_invariant_params* ip, int d, const double* a, const double* b) {
  clipper::invariants::ROMAN inv(roman_params(ip));
  clipper::invariants::Datum x(d), y(d);
  for (int i = 0; i < d; ++i) { x(i) = a[i]; y(i) = b[i]; }
  return inv.single_similarity(x, y);
}

double ref_pairwise_similarity(const ref_invariant_params* ip, int d, const double* ai, const double* aj, const double* bi,
                               const double* bj) {
  clipper::invariants::Datum x(d), y(d), z(d), w(d);
  for (int i = 0; i < d; ++i) { x(i) = ai[i]; y(i) = aj[i]; z(i) = bi[i]; w(i) = bj[i]; }
  if (ip->kind == 1) {
    clipper::invariants::EuclideanDistance::Params e;
    e.sigma = ip->sigma; e.epsilon = ip->epsilon; e.mindist = ip->mindist;
    clipper::invariants::EuclideanDistance inv(e);
    return inv(x, y, z, w);
  }
  clipper::invariants::ROMAN inv(roman_params(ip));
  return inv.pairwise_similarity(x, y, z, w);
}

double ref_fusion(const ref_invariant_params* ip, double pair_ij, double si, double sj) {
  clipper::invariants::ROMAN inv(roman_params(ip));
  return inv.pairwise_single_fusion(pair_ij, si, sj);
}

int ref_k_largest(const double* x, int64_t n, int k, int32_t* out) {
  Eigen::VectorXd v(n);
  for (int64_t i = 0; i < n; ++i) v(i) = x[i];
  const std::vector<int> r = clipper::utils::findIndicesOfkLargest(v, k);
  for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
  return (int)r.size();
}

void ref_k2ij(uint64_t k, uint64_t n, uint64_t* i, uint64_t* j) {
  size_t a, b;
  std::tie(a, b) = clipper::utils::k2ij((size_t)k, (size_t)n);
  *i = a; *j = b;
}

// clipper::dsd::solve(M, S) on a dense ROW-major symmetric weight matrix (upper triangle authoritative)
int ref_dsd(int64_t n, const double* W, int32_t k, const int32_t* S, int32_t* out) {
  try {
    Eigen::MatrixXd M = Eigen::MatrixXd::Zero(n, n);
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = i + 1; j < n; ++j) M(i, j) = W[i * n + j];
    std::vector<int> s(S, S + k);
    const std::vector<int> r = clipper::dsd::solve(M, s);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
  } catch (...) { return -1; }
}

}  // extern "C"

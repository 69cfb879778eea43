// rb_affinity.cu — K1a (association records + single scores) and K1 (pairwise affinity tiles).
//
// Replaces, for a batch of problems at once:
//   CLIPPERPairwiseAndSingle::scorePairwiseAndSingleConsistency  clipper/src/clipper.cpp:77-136
//   CLIPPER::scorePairwiseConsistency                            clipper/src/clipper.cpp:29-73
//   ROMAN::{single_similarity,pairwise_similarity,pairwise_single_fusion}
//                                                                clipper/src/invariants/roman.cpp
//   EuclideanDistance::operator()                                clipper/src/invariants/euclidean_distance.cpp:13-31
//   CLIPPER::getAffinityMatrix/getConstraintMatrix/setMatrixData clipper/src/clipper.cpp:202-229
//
// Arithmetic notes: the translation unit is compiled with -fmad=false so that a*b+c keeps two
// roundings, like the reference's plain C++ expressions; sqrt and division are IEEE; exp/pow
// are CUDA's (<= 2 ulp from libm).  All gates (distinctness, single-score, mindist, epsilon,
// affinityeps) are evaluated exactly as written in the reference, in the same order.
#include "rb_kernels.cuh"

namespace {

// roman.cpp:11-12 — truncated literals, verbatim
constexpr double kSqrtTwoThirds = 0.81649658092;
constexpr double kSqrtOneThird = 0.57735026919;

struct Rec {
  double ax, ay, az, bx, by, bz, s;
  int32_t i1, i2;
};

__device__ __forceinline__ double norm3(double dx, double dy, double dz) {
  return sqrt((dx * dx + dy * dy) + dz * dz);
}

// roman.cpp:17-85
__device__ __forceinline__ double pairwise_roman(const RbDevParams& P, const Rec& i, const Rec& j) {
  const rb_invariant_params& p = P.inv;
  const double dax = i.ax - j.ax, day = i.ay - j.ay, daz = i.az - j.az;
  const double dbx = i.bx - j.bx, dby = i.by - j.by, dbz = i.bz - j.bz;
  const double l1 = norm3(dax, day, daz);
  const double l2 = norm3(dbx, dby, dbz);
  if (p.mindist > 0 && (l1 < p.mindist || l2 < p.mindist)) return 0.0;

  double epsilon = p.epsilon, sigma = p.sigma;
  if (p.drift_aware) {
    const double e2 = p.epsilon * p.drift_scale * 0.5 * (l1 + l2);
    epsilon = (p.epsilon < e2) ? e2 : p.epsilon;  // std::max
    if (p.drift_scale_sigma) {
      const double s2 = p.sigma * p.drift_scale * 0.5 * (l1 + l2);
      sigma = (p.sigma < s2) ? s2 : p.sigma;
    }
  }

  if (p.gravity_guided) {
    const double xy1 = sqrt(dax * dax + day * day);
    const double xy2 = sqrt(dbx * dbx + dby * dby);
    const double z1 = daz, z2 = dbz;
    const double c_xy = fabs(xy1 - xy2);
    const double c_z = fabs(z1 - z2);
    double sigma_xy = sigma, sigma_z = sigma, eps_xy = epsilon, eps_z = epsilon;
    if (p.gravity_unc_ang_rad > 0.0) {
      const double xym = 0.5 * (xy1 + xy2);
      const double zm = 0.5 * (fabs(z1) + fabs(z2));
      const double dxy = fabs(xym * P.cos_unc - xym);
      const double dz = fabs(zm * P.sin_unc);
      sigma_xy += dxy;
      sigma_z += dz;
      eps_xy += dxy;
      eps_z += dz;
    }
    if (c_xy > kSqrtTwoThirds * eps_xy || c_z > kSqrtOneThird * eps_z) return 0.0;
    return exp(-0.5 * (c_xy * c_xy / (2.0 / 3.0 * sigma_xy * sigma_xy) +
                       c_z * c_z / (sigma_z * sigma_z / 3.0)));
  } else {
    const double c = fabs(l1 - l2);
    if (c > epsilon) return 0.0;
    return exp(-0.5 * c * c / (sigma * sigma));
  }
}

// euclidean_distance.cpp:13-31 (strict c < epsilon)
__device__ __forceinline__ double pairwise_euclid(const RbDevParams& P, const Rec& i, const Rec& j) {
  const rb_invariant_params& p = P.inv;
  const double l1 = norm3(i.ax - j.ax, i.ay - j.ay, i.az - j.az);
  const double l2 = norm3(i.bx - j.bx, i.by - j.by, i.bz - j.bz);
  if (p.mindist > 0 && (l1 < p.mindist || l2 < p.mindist)) return 0.0;
  const double c = fabs(l1 - l2);
  return (c < p.epsilon) ? exp(-0.5 * c * c / (p.sigma * p.sigma)) : 0.0;
}

// roman.cpp:87-114
__device__ __forceinline__ double fusion(const RbDevParams& P, double pair_ij, double si, double sj) {
  const rb_invariant_params& p = P.inv;
  if (p.ratio_feature_dim > 0 || p.cos_feature_dim > 0) {
    switch (p.fusion_method) {
      case 0: {
        const double dsp = pow(pair_ij, p.distance_weight);
        return pow(dsp * si * sj, 1.0 / (p.distance_weight + 2.0));
      }
      case 1:
        return (p.distance_weight * pair_ij + si + sj) / (p.distance_weight + 2.0);
      case 2:
        return pair_ij * si * sj;
      default:
        return 0.0;
    }
  }
  return pair_ij;
}

// One entry M(i,j) with i < j in association order (clipper.cpp:96-127 / :40-64).
__device__ __forceinline__ double pair_score(const RbDevParams& P, const Rec& i, const Rec& j) {
  if (i.i1 < 0 || j.i1 < 0) return 0.0;                  // padding
  if (i.i1 == j.i1 || i.i2 == j.i2) return 0.0;          // distinctness constraint
  double scr;
  if (P.inv.kind == RB_INVARIANT_ROMAN) {
    if (i.s < P.affinityeps || j.s < P.affinityeps) return 0.0;
    const double pw = pairwise_roman(P, i, j);
    scr = fusion(P, pw, i.s, j.s);
  } else if (P.inv.kind == RB_INVARIANT_ROMAN_PAIRWISE) {  // scorePairwiseConsistency: ROMAN::operator() (abstract.h:99-102)
    scr = pairwise_roman(P, i, j);
  } else {
    scr = pairwise_euclid(P, i, j);
  }
  return (scr > P.affinityeps) ? scr : 0.0;
}

// ---------------------------------------------------------------------------------------------
// The z gate ahead of everything else.  K1 screens all (live row, live column) pairs of a tile with the two tests of
// pair_score() that need no square root — distinctness, and (gravity-guided, not drift-aware ROMAN) the reference's own
// z test on the same fp64 expressions (roman.cpp:54-77: c_z > sqrt(1/3) * eps_z returns 0) — and evaluates pair_score(),
// the reference's expressions untouched, on the survivors only, compacted onto full warps.  At realistic densities nine
// pairs in ten fail the z test, and with the gates inside the scoring code nearly every 32-pair sweep still walked the
// whole square-root / exp / pow path for its one or two survivors.  `false` <=> pair_score() returns exactly 0.0 through
// one of these two exits (no other effect); the operands' order does not matter (differences are negated exactly).
// (fp32 screens of the length / xy gates were tried as well: conversions and square roots on every z survivor cost more
// than the few extra pairs they drop.)
struct ZGate {
  bool on;
  double eps, sin_unc;
  bool unc;
};
__device__ __forceinline__ ZGate make_zgate(const RbDevParams& P) {
  ZGate g;
  const rb_invariant_params& p = P.inv;
  g.on = (p.kind == RB_INVARIANT_ROMAN || p.kind == RB_INVARIANT_ROMAN_PAIRWISE) && p.gravity_guided && !p.drift_aware;
  g.eps = p.epsilon;
  g.sin_unc = P.sin_unc;
  g.unc = p.gravity_unc_ang_rad > 0.0;
  return g;
}
__device__ __forceinline__ bool zgate_pass(const ZGate& g, double az_i, double bz_i, double az_j, double bz_j) {
  const double daz = az_i - az_j, dbz = bz_i - bz_j;
  double eps_z = g.eps;
  if (g.unc) eps_z += fabs(0.5 * (fabs(daz) + fabs(dbz)) * g.sin_unc);
  return !(fabs(daz - dbz) > kSqrtOneThird * eps_z);
}

// ---------------------------------------------------------------------------------------------
// K1a: one warp per association — gathers the two object centres and evaluates
// ROMAN::single_similarity (roman.cpp:116-164); descriptor dot products are warp-strided
// fp64 sums reduced with a fixed xor tree.
__global__ void __launch_bounds__(256) rb_prep_kernel(const RbProblem* __restrict__ probs, RbAssocSoA rec,
                                                      const double* __restrict__ D1, const double* __restrict__ D2,
                                                      const int32_t* __restrict__ A, RbDevParams P) {
  const RbProblem pr = probs[blockIdx.y];
  if (pr.is_copy) return;  // shares another problem's matrix and records (rb_plan_shared)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mp = pr.nb * RB_TILE;
  const rb_invariant_params& p = P.inv;
  for (int k = blockIdx.x * 8 + warp; k < mp; k += gridDim.x * 8) {
    const int64_t o = pr.off_pad + k;
    if (k >= pr.m) {
      if (lane == 0) {
        rec.ax[o] = rec.ay[o] = rec.az[o] = rec.bx[o] = rec.by[o] = rec.bz[o] = 0.0;
        rec.single[o] = 0.0;
        rec.i1[o] = -1;
        rec.i2[o] = -1;
      }
      continue;
    }
    int i1, i2;
    if (pr.all_to_all) {  // utils::createAllToAll, utils.h:61-71
      i1 = k / pr.n2;
      i2 = k - i1 * pr.n2;
    } else {
      i1 = A[2 * (pr.off_assoc + k)];
      i2 = A[2 * (pr.off_assoc + k) + 1];
    }
    const double* a = D1 + (pr.off_obj1 + i1) * (int64_t)pr.d;
    const double* b = D2 + (pr.off_obj2 + i2) * (int64_t)pr.d;

    double single = 1.0;
    if (p.kind == RB_INVARIANT_ROMAN && (p.cos_feature_dim > 0 || p.ratio_feature_dim > 0)) {
      bool zero = false;
      double cos_scaled = 0.0, ratio_score = 0.0;
      if (p.cos_feature_dim > 0) {
        const double* fa = a + p.point_dim + p.ratio_feature_dim;
        const double* fb = b + p.point_dim + p.ratio_feature_dim;
        double dot = 0.0, na = 0.0, nb = 0.0;
        for (int t = lane; t < p.cos_feature_dim; t += 32) {
          const double x = fa[t], y = fb[t];
          dot += x * y;
          na += x * x;
          nb += y * y;
        }
        dot = rb_warp_sum(dot);
        na = rb_warp_sum(na);
        nb = rb_warp_sum(nb);
        const double c = dot / (sqrt(na) * sqrt(nb));
        if (c >= p.cosine_max) cos_scaled = 1.0;
        else if (c <= p.cosine_min) zero = true;
        else cos_scaled = (c - p.cosine_min) / (p.cosine_max - p.cosine_min);
      }
      if (!zero && p.ratio_feature_dim > 0) {
        double prod = 1.0;
        bool below = false;
        for (int t = 0; t < p.ratio_feature_dim; ++t) {
          const double x = a[p.point_dim + t], y = b[p.point_dim + t];
          const double r = x < y ? x / y : y / x;
          if (r < p.ratio_epsilon[t]) below = true;
          prod *= r;
        }
        if (below) zero = true;
        else ratio_score = pow(prod, 1.0 / p.ratio_feature_dim);
      }
      if (zero) single = 0.0;
      else if (p.ratio_feature_dim == 0) single = cos_scaled;
      else if (p.cos_feature_dim == 0) single = ratio_score;
      else
        single = pow(pow(ratio_score, p.ratio_weight) * pow(cos_scaled, p.cosine_weight),
                     1.0 / (p.ratio_weight + p.cosine_weight));
    }
    if (lane == 0) {
      const int pd = (p.kind != RB_INVARIANT_EUCLIDEAN) ? p.point_dim : (pr.d < 3 ? pr.d : 3);
      rec.ax[o] = a[0];
      rec.ay[o] = pd > 1 ? a[1] : 0.0;
      rec.az[o] = pd > 2 ? a[2] : 0.0;
      rec.bx[o] = b[0];
      rec.by[o] = pd > 1 ? b[1] : 0.0;
      rec.bz[o] = pd > 2 ? b[2] : 0.0;
      rec.single[o] = single;
      rec.i1[o] = i1;
      rec.i2[o] = i2;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// K1a for all-to-all plans: the n1 x n2 cosine table is the one dense contraction of this path (descriptors^T *
// descriptors, n1 x D x n2) and is evaluated as one — on the fp64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA) —
// instead of re-reading two D-vectors per association.
//
// One CTA per (<= 64) x (<= 64) block of (object of map 1, object of map 2).  Compute warp w owns the 8 map-1 objects
// 8w .. 8w+7 and all NCB 8-object column blocks of map 2: per 4 descriptor columns it loads one A fragment and NCB B
// fragments (one double per lane each) and issues NCB DMMAs, accumulating two doubles per lane per column block.  One
// more warp accumulates the squared norms of all rows of the block.  40 x 50 objects: 5 compute warps x 7 column
// blocks; two such CTAs share an SM.
//
// Why the tensor pipe although its fp64 rate equals the FMA pipe's: shared-memory bandwidth.  An FMA register tile
// needs (MR + NR) operand words per MR * NR multiply-adds per lane and every 8-byte shared load costs two wavefronts
// (measured: a 5 x 4 tile stalls at 54 % of the fp64 peak, tools/micro/fp64_rate.cu); the MMA fragments of a warp
// hold 32 operands for 256 multiply-adds — a third of the wavefronts.
//
// Staging: descriptor columns [k0, k0 + 32) of all rows of the block form one stage ([row][k], pitch 36 doubles: a
// fragment load touches rows g = 0..7 x columns t = 0..3, and 36 = 4 mod 16 puts the 16 words of a half-warp into 16
// different bank pairs); a ring of 3 stages is filled with asynchronous global->shared copies (cp.async, 8 bytes each
// — SASS LDGSTS) issued two stages ahead; ONE block barrier per 32 columns.  Why not the bulk / TMA form the solver
// uses for its 8 KB tiles: object rows are d = 3 + r + D doubles long (775 for the `roman` method, odd), so every
// other row starts 8 bytes off a 16-byte boundary; cp.async.bulk and tensor maps need 16-byte aligned sources.
//
// Arithmetic: every dot product is a chain of fp64 fused multiply-adds over the descriptor index in ascending order
// (the m8n8k4 DMMA accumulates its four products in k order, each with a single rounding), the squared norms likewise
// in the norm warp: the same values as a scalar loop with fma(), so the cosine gate `<= cosine_min` sees what it saw
// in round 1.  Nothing is split across threads or tree-reduced; results do not depend on the launch shape.
// measured on B200 (592 problems of 40 x 50 objects, D = 768; tools/micro/variants): columns per stage / stages / CTAs
// per SM -> 16/3/4: 0.182 ms, 32/2/3: 0.188, 32/3/2: 0.175, 64/2/2: 0.186, 16/4/3: 0.233 (1.33 waves), 128/2/1: 0.240
#ifndef RB_PREP_KC
#define RB_PREP_KC 32
#define RB_PREP_STAGES 3
#endif
#ifndef RB_PREP_MINB
#define RB_PREP_MINB 2
#endif
constexpr int PREP_KC = RB_PREP_KC;   // descriptor columns per stage: 512 contiguous bytes per row and request (HBM pages are 1 KB)
constexpr int PREP_LD = PREP_KC + 4;  // shared-memory row pitch (doubles): 4 mod 16, see above
constexpr int PREP_STAGES = RB_PREP_STAGES;  // copies run one stage ahead of the arithmetic
constexpr int PREP_HEAD = 8;          // leading (non-descriptor) columns of a row kept for the epilogue: centre + ratio features
constexpr int PREP_MAXCB = 8;         // column blocks per warp (n2 <= 64 per block of the grid)

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// D(8x8) += A(8x4, row) * B(4x8, col), fp64: lane l holds A[l>>2][l&3], B[l&3][l>>2], D[l>>2][2*(l&3) + {0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NCB, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) rb_prep_table_kernel(const RbProblem* __restrict__ probs, RbAssocSoA rec,
                                                               const double* __restrict__ D1,
                                                               const double* __restrict__ D2, RbDevParams P,
                                                               int tiles2_max, int NRB) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  const RbProblem pr = probs[blockIdx.y];
  if (pr.is_copy) return;  // shares another problem's matrix and records (rb_plan_shared)
  const rb_invariant_params& p = P.inv;
  const int PT1 = 8 * NRB, PT2 = 8 * NCB, R = PT1 + PT2;
  const int NTH = blockDim.x;                  // 32 * NRB compute threads + 32 (norm warp)
  double* stage0 = reinterpret_cast<double*>(prep_smem);                // [PREP_STAGES][R][PREP_LD]
  double* s_norm = stage0 + (size_t)PREP_STAGES * R * PREP_LD;          // [R] squared norms
  double* s_head = s_norm + R;                                          // [R][PREP_HEAD] centres and ratio features
  const double** s_rowp = reinterpret_cast<const double**>(s_head + (size_t)R * PREP_HEAD);   // [R] descriptor row starts (null: no such object)
  const int t1 = blockIdx.x / tiles2_max, t2 = blockIdx.x % tiles2_max;
  const int i1_0 = t1 * PT1, i2_0 = t2 * PT2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 0) {  // padding records of this problem
    for (int k = pr.m + tid; k < pr.nb * RB_TILE; k += NTH) {
      const int64_t o = pr.off_pad + k;
      rec.ax[o] = rec.ay[o] = rec.az[o] = rec.bx[o] = rec.by[o] = rec.bz[o] = 0.0;
      rec.single[o] = 0.0;
      rec.i1[o] = -1;
      rec.i2[o] = -1;
    }
  }
  if (i1_0 >= pr.n1 || i2_0 >= pr.n2) return;
  const bool norm_warp = warp == NRB;
  const double* A0 = D1 + pr.off_obj1 * (int64_t)pr.d;
  const double* B0 = D2 + pr.off_obj2 * (int64_t)pr.d;
  const int foff = p.point_dim + p.ratio_feature_dim;
  const int Dc = (p.kind == RB_INVARIANT_ROMAN) ? p.cos_feature_dim : 0;
  const int nchunks = (Dc + PREP_KC - 1) / PREP_KC;
  const int nhead = foff < PREP_HEAD ? foff : PREP_HEAD;

  // stage row q < PT1 is object i1_0 + q of map 1, q >= PT1 object i2_0 + q - PT1 of map 2 (rows / columns that do
  // not exist are zeros)
  auto row_ptr = [&](int q) -> const double* {
    const bool is1 = q < PT1;
    const int o = is1 ? i1_0 + q : i2_0 + q - PT1;
    if (is1 ? o >= pr.n1 : o >= pr.n2) return nullptr;
    return (is1 ? A0 : B0) + (int64_t)o * pr.d;
  };
  for (int q = tid; q < R; q += NTH) {
    const double* src = row_ptr(q);
    s_rowp[q] = src ? src + foff : nullptr;
  }
  for (int e = tid; e < R * PREP_HEAD; e += NTH) {  // centres and ratio features for the epilogue
    const int q = e / PREP_HEAD, k = e % PREP_HEAD;
    const double* src = row_ptr(q);
    s_head[e] = (src != nullptr && k < nhead) ? src[k] : 0.0;
  }
  __syncthreads();
  auto issue = [&](int c) {  // asynchronous copies of chunk c into stage c % PREP_STAGES
    double* st = stage0 + (size_t)(c % PREP_STAGES) * R * PREP_LD;
    const int k0 = c * PREP_KC;
    for (int e = tid; e < R * PREP_KC; e += NTH) {  // consecutive threads -> consecutive columns of one row: coalesced
      const int q = e / PREP_KC, k = e % PREP_KC;   // PREP_KC is a power of two
      const double* src = s_rowp[q];
      double* dst = st + q * PREP_LD + k;
      if (src != nullptr && (k0 + k) < Dc) cp_async_8(dst, src + k0 + k);
      else *dst = 0.0;
    }
  };

  double acc[NCB][2];
#pragma unroll
  for (int c = 0; c < NCB; ++c) acc[c][0] = acc[c][1] = 0.0;
  constexpr int NJ = 4;              // norm warp: rows lane, lane + 32, ... of the block (R <= 128)
  double nrm[NJ] = {0.0, 0.0, 0.0, 0.0};
  const int g = lane >> 2, t = lane & 3;

  for (int c = 0; c < PREP_STAGES - 1; ++c) {
    if (c < nchunks) issue(c);
    cp_async_commit();
  }
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<PREP_STAGES - 2>();  // this thread's copies of chunk c have landed ...
    __syncthreads();                   // ... and everybody's; also: everybody is done with the stage refilled next
    if (c + PREP_STAGES - 1 < nchunks) issue(c + PREP_STAGES - 1);
    cp_async_commit();
    const double* st = stage0 + (size_t)(c % PREP_STAGES) * R * PREP_LD;
    if (!norm_warp) {
      const double* pa = st + (warp * 8 + g) * PREP_LD + t;
      const double* pb = st + (PT1 + g) * PREP_LD + t;
#pragma unroll
      for (int k = 0; k < PREP_KC; k += 4) {
        const double a = pa[k];
        double b[NCB];
#pragma unroll
        for (int q = 0; q < NCB; ++q) b[q] = pb[q * 8 * PREP_LD + k];
#pragma unroll
        for (int q = 0; q < NCB; ++q) dmma_m8n8k4(acc[q][0], acc[q][1], a, b[q]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int q = lane + 32 * j;
        if (q < R) {
          const double* pr_ = st + q * PREP_LD;
          double s2 = nrm[j];
#pragma unroll
          for (int k = 0; k < PREP_KC; ++k) s2 = fma(pr_[k], pr_[k], s2);
          nrm[j] = s2;
        }
      }
    }
  }
  if (norm_warp)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (lane + 32 * j < R) s_norm[lane + 32 * j] = nrm[j];
  __syncthreads();
  if (norm_warp) return;

  const int pd = (p.kind != RB_INVARIANT_EUCLIDEAN) ? p.point_dim : (pr.d < 3 ? pr.d : 3);
  // 2 * NCB outputs per lane: row 8*warp + g of map 1, columns 8*cb + 2*t + {0, 1} of map 2; one instance of the
  // scoring code (pow is long)
#pragma unroll 1
  for (int oq = 0; oq < 2 * NCB; ++oq) {
    const int cb = oq >> 1, hi = oq & 1;
    const int l1 = warp * 8 + g, l2 = cb * 8 + 2 * t + hi;
    const int i1 = i1_0 + l1, i2 = i2_0 + l2;
    if (i1 >= pr.n1 || i2 >= pr.n2) continue;
    const double* ha = s_head + l1 * PREP_HEAD;
    const double* hb = s_head + (PT1 + l2) * PREP_HEAD;
    double dotv = 0.0;
#pragma unroll
    for (int q = 0; q < NCB; ++q) {  // register select: acc[][] must not be indexed dynamically
      if (q == cb) dotv = hi ? acc[q][1] : acc[q][0];
    }
    double single = 1.0;
    if (p.kind == RB_INVARIANT_ROMAN && (p.cos_feature_dim > 0 || p.ratio_feature_dim > 0)) {  // roman.cpp:116-164
      bool zero = false;
      double cos_scaled = 0.0, ratio_score = 0.0;
      if (p.cos_feature_dim > 0) {
        const double cs = dotv / (sqrt(s_norm[l1]) * sqrt(s_norm[PT1 + l2]));
        if (cs >= p.cosine_max) cos_scaled = 1.0;
        else if (cs <= p.cosine_min) zero = true;
        else cos_scaled = (cs - p.cosine_min) / (p.cosine_max - p.cosine_min);
      }
      if (!zero && p.ratio_feature_dim > 0) {
        double prod = 1.0;
        bool below = false;
        for (int f = 0; f < p.ratio_feature_dim; ++f) {
          const int col = p.point_dim + f;
          const double x = col < PREP_HEAD ? ha[col] : A0[(int64_t)i1 * pr.d + col];
          const double y = col < PREP_HEAD ? hb[col] : B0[(int64_t)i2 * pr.d + col];
          const double qv = x < y ? x / y : y / x;
          if (qv < p.ratio_epsilon[f]) below = true;
          prod *= qv;
        }
        if (below) zero = true;
        else ratio_score = pow(prod, 1.0 / p.ratio_feature_dim);
      }
      if (zero) single = 0.0;
      else if (p.ratio_feature_dim == 0) single = cos_scaled;
      else if (p.cos_feature_dim == 0) single = ratio_score;
      else
        single = pow(pow(ratio_score, p.ratio_weight) * pow(cos_scaled, p.cosine_weight),
                     1.0 / (p.ratio_weight + p.cosine_weight));
    }
    const int64_t o = pr.off_pad + (int64_t)i1 * pr.n2 + i2;  // utils::createAllToAll order
    rec.ax[o] = ha[0]; rec.ay[o] = pd > 1 ? ha[1] : 0.0; rec.az[o] = pd > 2 ? ha[2] : 0.0;
    rec.bx[o] = hb[0]; rec.by[o] = pd > 1 ? hb[1] : 0.0; rec.bz[o] = pd > 2 ? hb[2] : 0.0;
    rec.single[o] = single;
    rec.i1[o] = i1;
    rec.i2[o] = i2;
  }
}

__device__ __forceinline__ Rec load_rec(const RbAssocSoA& rec, int64_t o) {
  Rec r;
  r.ax = rec.ax[o]; r.ay = rec.ay[o]; r.az = rec.az[o];
  r.bx = rec.bx[o]; r.by = rec.by[o]; r.bz = rec.bz[o];
  r.s = rec.single[o];
  r.i1 = rec.i1[o]; r.i2 = rec.i2[o];
  return r;
}

// tile q of a problem with nb tile-rows -> (I, J), I <= J
__device__ __forceinline__ void tile_coords(int64_t q, int nb, int& I, int& J) {
  const double t = 2.0 * nb + 1.0;
  int i = (int)floor((t - sqrt(t * t - 8.0 * (double)q)) * 0.5);
  if (i < 0) i = 0;
  if (i > nb - 1) i = nb - 1;
  while (i > 0 && rb_tile_index(i, i, nb) > q) --i;
  while (i + 1 < nb && rb_tile_index(i + 1, i + 1, nb) <= q) ++i;
  I = i;
  J = i + (int)(q - rb_tile_index(i, i, nb));
}

// ---------------------------------------------------------------------------------------------
// Per 32-association block: bit r set <=> association r of the block can carry an edge (not padding,
// single score not below affinityeps — clipper.cpp:104-108; NaN counts as live like the reference's `<`).
// One warp per block; K1 reads two of these words per tile instead of 64 records.
__global__ void __launch_bounds__(256) rb_block_mask_kernel(const RbProblem* __restrict__ probs, RbAssocSoA rec,
                                                            uint32_t* __restrict__ masks, RbDevParams P) {
  const RbProblem pr = probs[blockIdx.y];
  if (pr.is_copy) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool roman = (P.inv.kind == RB_INVARIANT_ROMAN);
  for (int I = blockIdx.x * 8 + warp; I < pr.nb; I += gridDim.x * 8) {
    const int64_t o = pr.off_pad + (int64_t)I * RB_TILE + lane;
    const unsigned live = __ballot_sync(0xffffffffu, rec.i1[o] >= 0 && !(roman && rec.single[o] < P.affinityeps));
    if (lane == 0) masks[(pr.off_pad >> 5) + I] = live;
  }
}

// ---------------------------------------------------------------------------------------------
// K1: one warp per 32x32 tile.  The tile is written as zeros (coalesced 16-byte stores; each row is a permutation of
// itself: the XOR swizzle of rb_common.cuh); the (live row, live column) pairs are screened, the survivors scored on
// full warps and their non-zero words written over the zeros (affinity_tile below).
// Grid = (CTAs per problem, problems).  A CTA first condenses the problem's block live masks into the
// list of LIVE blocks (blocks holding at least one association that can carry an edge); only tiles
// whose row block and column block are both live are evaluated.
//   tile-skip format (flags != null): dead tiles are never touched (their flags were zeroed by a
//     memset); a live tile is written and flagged iff it ends up holding a non-zero word;
//   dense format (flags == null): every tile of the packed upper triangle is written, dead ones as
//     zeros with 16-byte stores.
// Edge list of a problem (both formats): every non-zero word of the strict upper triangle is also appended as a
// triple (i << 16 | j, value) to the problem's list, capacity RB_EDGE_FACTOR * nb * 32 entries; the count keeps growing
// past the capacity so that the solver can tell an overflow (it then reads the tiles instead).
struct RbEdgeOut {
  unsigned* ij;
  double* v;
  int* count;  // this problem's counter
  int cap;
};

// The live associations of a tile's row block (or column block), compacted: slot k = the k-th live association.  One
// array per field, so that lanes holding consecutive slots read consecutive words (no bank conflicts) and lanes holding
// the same slot share one word (broadcast) — the array-of-records layout this replaces cost up to eight shared-memory
// wavefronts per load and made the kernel LSU-bound at realistic densities (ncu: 76 % of the LSU peak, fp64 pipe 33 %).
struct TileSide {
  double ax[RB_TILE], ay[RB_TILE], az[RB_TILE], bx[RB_TILE], by[RB_TILE], bz[RB_TILE], s[RB_TILE];
  int32_t i1[RB_TILE], i2[RB_TILE];
  unsigned char idx[RB_TILE];  // slot -> position in the block
};

__device__ __forceinline__ void fill_side(TileSide& S, const RbAssocSoA& rec, int64_t o, unsigned live, int lane) {
  const Rec r = load_rec(rec, o + lane);  // coalesced
  if ((live >> lane) & 1u) {
    const int k = __popc(live & ((1u << lane) - 1u));
    S.ax[k] = r.ax; S.ay[k] = r.ay; S.az[k] = r.az;
    S.bx[k] = r.bx; S.by[k] = r.by; S.bz[k] = r.bz;
    S.s[k] = r.s; S.i1[k] = r.i1; S.i2[k] = r.i2;
    S.idx[k] = (unsigned char)lane;
  }
}

__device__ __forceinline__ void affinity_tile(const RbProblem& pr, int I, int J, unsigned row_live, unsigned col_live,
                                              const RbAssocSoA& rec, TileSide& R, TileSide& C, unsigned short* surv,
                                              double* __restrict__ Mp,
                                              uint8_t* __restrict__ flags, const RbEdgeOut& eo, const RbDevParams& P,
                                              int lane) {
  const int64_t q = rb_tile_index(I, J, pr.nb);
  double* out = Mp + (pr.off_tile + q) * RB_TILE_ELEMS;
  __syncwarp();
  fill_side(C, rec, pr.off_pad + (int64_t)J * RB_TILE, col_live, lane);
  fill_side(R, rec, pr.off_pad + (int64_t)I * RB_TILE, row_live, lane);
  // the whole tile as zeros first (coalesced 16-byte stores); the few non-zero words overwrite them below
  {
    double2* o2 = reinterpret_cast<double2*>(out);
#pragma unroll
    for (int t = 0; t < RB_TILE_ELEMS / 2 / 32; ++t) o2[t * 32 + lane] = make_double2(0.0, 0.0);
  }
  __syncwarp();  // orders the zero stores before the value stores of other lanes to the same words
  const bool diag = (I == J);
  // Only (live row, live column) pairs are evaluated, and they are PACKED onto the lanes: pair p = (row slot p / nc,
  // column slot p % nc) over the nr live rows and nc live columns of the tile, 32 pairs per sweep.  With half of the
  // associations live a lane-per-column sweep leaves half of every warp idle (measured: 17 of 32 threads active).
  const int nr = __popc(row_live), nc = __popc(col_live);
  const int npairs = nr * nc;
  bool any_nz = false;
  // Survivors of the screens, (row slot << 5 | column slot), scored 32 at a time on full warps.
  auto score_batch = [&](int first, int count) {
    double v = 0.0;
    int r = 0, c = 0;
    if (lane < count) {
      const int e = surv[first + lane];
      const int ri = e >> 5, ci = e & 31;
      r = R.idx[ri];
      c = C.idx[ci];
      // i = the association with the smaller index: the row, except for the mirrored half of a diagonal tile.  The
      // pairwise score is bitwise symmetric in (i, j) (differences are negated exactly, then squared or fabs'ed); the
      // fusion is not ((d * si) * sj), so the single scores are the one thing that has to swap.
      const bool mirrored = diag && r > c;
      Rec a, b;
      a.ax = R.ax[ri]; a.ay = R.ay[ri]; a.az = R.az[ri]; a.bx = R.bx[ri]; a.by = R.by[ri]; a.bz = R.bz[ri];
      b.ax = C.ax[ci]; b.ay = C.ay[ci]; b.az = C.az[ci]; b.bx = C.bx[ci]; b.by = C.by[ci]; b.bz = C.bz[ci];
      a.i1 = R.i1[ri]; a.i2 = R.i2[ri]; b.i1 = C.i1[ci]; b.i2 = C.i2[ci];
      const double sr = R.s[ri], sc = C.s[ci];
      a.s = mirrored ? sc : sr;
      b.s = mirrored ? sr : sc;
      v = pair_score(P, a, b);
    }
    const bool nz = (__double_as_longlong(v) != 0ll);
    any_nz |= nz;
    if (nz) out[rb_tile_elem(r, c)] = v;
    if (eo.count) {
      const bool edge = nz && (!diag || r < c);
      const unsigned em = __ballot_sync(0xffffffffu, edge);
      if (em) {
        int base = 0;
        if (lane == 0) base = atomicAdd(eo.count, __popc(em));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(em & ((1u << lane) - 1u));
        if (edge && pos < eo.cap) {
          eo.ij[pos] = ((unsigned)(I * RB_TILE + r) << 16) | (unsigned)(J * RB_TILE + c);
          eo.v[pos] = v;
        }
      }
    }
  };
  const ZGate zg = make_zgate(P);
  int ns = 0;                                    // survivors waiting in surv[0 .. ns)
  int ri = lane / nc, ci = lane % nc;            // pair index p = lane + 32 * sweep  ->  (ri, ci)
  const int dr = 32 / nc, dc = 32 % nc;
  // one call site for the scoring code (its exp / pow paths are long: a second inlined copy costs instruction cache)
#pragma unroll 1
  for (int p0 = 0;; p0 += 32) {
    const bool done = p0 >= npairs;
    if (!done) {
      bool keep = false;
      if (ri < nr) {
        const int r = R.idx[ri], c = C.idx[ci];
        if (!(diag && r == c)) {  // r == c on the diagonal: implicit identity (clipper.cpp:129-130)
          keep = R.i1[ri] != C.i1[ci] && R.i2[ri] != C.i2[ci];   // distinctness (padding is never live)
          if (keep && zg.on) keep = zgate_pass(zg, R.az[ri], R.bz[ri], C.az[ci], C.bz[ci]);
        }
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) surv[ns + __popc(km & ((1u << lane) - 1u))] = (unsigned short)((ri << 5) | ci);
      ns += __popc(km);
      ri += dr;
      ci += dc;
      if (ci >= nc) { ci -= nc; ++ri; }
    }
    if (ns >= 32 || (done && ns > 0)) {
      __syncwarp();
      const int cnt = ns < 32 ? ns : 32;
      ns -= cnt;
      score_batch(ns, cnt);  // the newest entries: nothing has to move
      __syncwarp();
    }
    if (done) break;
  }
  const unsigned nzb = __ballot_sync(0xffffffffu, any_nz);
  if (flags && lane == 0) flags[pr.off_tile + q] = nzb ? 1 : 0;
}

__global__ void __launch_bounds__(256) rb_affinity_kernel(const RbProblem* __restrict__ probs, RbAssocSoA rec,
                                                          const uint32_t* __restrict__ masks,
                                                          double* __restrict__ Mp, uint8_t* __restrict__ flags,
                                                          unsigned* __restrict__ edge_ij, double* __restrict__ edge_v,
                                                          int* __restrict__ edge_cnt, RbDevParams P) {
  __shared__ TileSide s_row[8], s_col[8];
  __shared__ unsigned short s_surv[8][64];
  __shared__ unsigned s_mask[256];   // live mask per block (nb <= 160)
  __shared__ unsigned char s_live[256];  // indices of live blocks, ascending
  __shared__ int s_nl;
  const RbProblem pr = probs[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = pr.nb;
  if (nb == 0 || pr.is_copy) return;
  // An association whose single score is below affinityeps (clipper.cpp:104-108), or padding, can
  // have no edge at all (NaN scores are "live", exactly as the reference's `<` test);
  // rb_block_mask_kernel has condensed that into one word per 32-association block.
  for (int i = threadIdx.x; i < nb; i += 256) s_mask[i] = masks[(pr.off_pad >> 5) + i];
  __syncthreads();
  if (warp == 0) {
    int n = 0;
    for (int i0 = 0; i0 < nb; i0 += 32) {
      const int i = i0 + lane;
      const bool lv = (i < nb) && s_mask[i] != 0u;
      const unsigned mk = __ballot_sync(0xffffffffu, lv);
      if (lv) s_live[n + __popc(mk & ((1u << lane) - 1u))] = (unsigned char)i;
      n += __popc(mk);
    }
    if (lane == 0) s_nl = n;
  }
  __syncthreads();
  const int nl = s_nl;
  RbEdgeOut eo;
  eo.ij = edge_ij ? edge_ij + (size_t)pr.off_pad * RB_EDGE_FACTOR : nullptr;
  eo.v = edge_v ? edge_v + (size_t)pr.off_pad * RB_EDGE_FACTOR : nullptr;
  eo.count = edge_cnt ? edge_cnt + blockIdx.y : nullptr;
  eo.cap = RB_EDGE_FACTOR * nb * RB_TILE;

  // Work items: tile-skip format — live x live block pairs only (pair index <-> (a <= b) over the live-block list; dead
  // tiles are never touched); dense format — every tile of the packed upper triangle, dead ones written as zeros.
  const int nwork = flags ? nl * (nl + 1) / 2 : (int)pr.n_tiles;
#pragma unroll 1
  for (int wi = blockIdx.x * 8 + warp; wi < nwork; wi += gridDim.x * 8) {
    int I, J;
    if (flags) {
      // invert wi = a*nl - a(a-1)/2 + (b - a)
      int a = (int)floor(((2.0 * nl + 1.0) - sqrt((2.0 * nl + 1.0) * (2.0 * nl + 1.0) - 8.0 * (double)wi)) * 0.5);
      if (a < 0) a = 0;
      if (a > nl - 1) a = nl - 1;
      while (a > 0 && a * nl - a * (a - 1) / 2 > wi) --a;
      while (a + 1 < nl && (a + 1) * nl - (a + 1) * a / 2 <= wi) ++a;
      const int bcol = a + (wi - (a * nl - a * (a - 1) / 2));
      I = s_live[a];
      J = s_live[bcol];
    } else {
      tile_coords(wi, nb, I, J);
    }
    const unsigned row_live = s_mask[I], col_live = s_mask[J];
    if (row_live == 0u || col_live == 0u) {  // dense format only: a dead tile is all zeros
      double2* o2 = reinterpret_cast<double2*>(Mp + (pr.off_tile + wi) * RB_TILE_ELEMS);
#pragma unroll
      for (int t = 0; t < RB_TILE_ELEMS / 2 / 32; ++t) o2[t * 32 + lane] = make_double2(0.0, 0.0);
      continue;
    }
    affinity_tile(pr, I, J, row_live, col_live, rec, s_row[warp], s_col[warp], s_surv[warp], Mp, flags, eo, P, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// Dense <-> packed conversions (clipper.cpp:202-229).
__global__ void rb_unpack_kernel(RbProblem pr, const double* __restrict__ Mp, const uint8_t* __restrict__ flags,
                                 double* __restrict__ Mout, double* __restrict__ Cout) {
  const int m = pr.m;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)m * m;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (int64_t)i * m);
    double mv, cv;
    if (i == j) {
      mv = 1.0;
      cv = 1.0;
    } else {
      const int a = i < j ? i : j, b = i < j ? j : i;
      const int I = a / RB_TILE, J = b / RB_TILE;
      const int64_t tl = pr.off_tile + rb_tile_index(I, J, pr.nb);
      const double v = (flags && !flags[tl]) ? 0.0 : Mp[tl * RB_TILE_ELEMS + rb_tile_elem(a % RB_TILE, b % RB_TILE)];
      cv = (__double_as_longlong(v) != 0ll) ? 1.0 : 0.0;
      mv = (cv != 0.0) ? fabs(v) : 0.0;  // -0.0 (constraint-only entry) reads back as 0
    }
    if (Mout) Mout[e] = mv;
    if (Cout) Cout[e] = cv;
  }
}

__global__ void rb_pack_kernel(RbProblem pr, const double* __restrict__ Min, const double* __restrict__ Cin,
                               double* __restrict__ Mp, int* __restrict__ bad) {
  const int m = pr.m;
  const int64_t total = pr.n_tiles * RB_TILE_ELEMS;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = e / RB_TILE_ELEMS;
    const int w = (int)(e - q * RB_TILE_ELEMS);
    int I, J;
    tile_coords(q, pr.nb, I, J);
    const int wr = w / RB_TILE, wc = (w % RB_TILE) ^ wr;  // inverse of rb_tile_elem
    const int i = I * RB_TILE + wr, j = J * RB_TILE + wc;
    double v = 0.0;
    if (i < m && j < m && i != j) {
      const int a = i < j ? i : j, b = i < j ? j : i;  // strict upper triangle of the inputs
      const double mv = Min[(int64_t)a * m + b];
      const double cv = Cin ? Cin[(int64_t)a * m + b] : (mv != 0.0 ? 1.0 : 0.0);
      if (mv != 0.0) {
        v = mv;
        if (cv == 0.0) *bad = 1;
      } else if (cv != 0.0) {
        v = -0.0;
      }
    }
    Mp[pr.off_tile * RB_TILE_ELEMS + e] = v;
  }
}

// stored (flagged) tiles per problem: the tile-skip format's footprint, for the roofline accounting
__global__ void rb_count_flags_kernel(const RbProblem* __restrict__ probs, int B, const uint8_t* __restrict__ flags,
                                      int32_t* __restrict__ out) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const RbProblem pr = probs[b];
  int c = 0;
  for (int64_t t = threadIdx.x; t < pr.n_tiles; t += blockDim.x) c += flags[pr.off_tile + t] ? 1 : 0;
  __shared__ int s[32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s[w];
    out[b] = tot;
  }
}

// live associations per problem (those that can carry an edge): popcount of the block live masks
__global__ void rb_live_counts_kernel(const RbProblem* __restrict__ probs, int B, const uint32_t* __restrict__ masks,
                                      int32_t* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const RbProblem pr = probs[b];
  int c = 0;
  for (int I = threadIdx.x & 31; I < pr.nb; I += 32) c += __popc(masks[(pr.off_rec >> 5) + I]);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) out[b] = c;
}

}  // namespace

cudaError_t rb_launch_live_counts(const RbProblem* probs, int B, const uint32_t* masks, int32_t* out, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  rb_live_counts_kernel<<<(B + 7) / 8, 256, 0, st>>>(probs, B, masks, out);
  return cudaGetLastError();
}

cudaError_t rb_launch_count_flags(const RbProblem* probs, int B, const uint8_t* flags, int32_t* out, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  rb_count_flags_kernel<<<B, 256, 0, st>>>(probs, B, flags, out);
  return cudaGetLastError();
}

cudaError_t rb_launch_prep(const RbProblem* probs, int B, int max_mp, int max_n1, int max_n2, bool all_to_all,
                           const RbAssocSoA& rec, const double* D1, const double* D2, const int32_t* A,
                           const RbDevParams& P, cudaStream_t st) {
  if (B == 0 || max_mp == 0) return cudaSuccess;
  if (all_to_all && max_n1 > 0 && max_n2 > 0) {
    // blocks of at most 64 x 64 objects (larger maps take several blocks): NRB compute warps (8 map-1 objects each),
    // NCB 8-object column blocks per warp
    const int e1 = max_n1 < 64 ? max_n1 : 64, e2 = max_n2 < 64 ? max_n2 : 64;
    const int NRB = (e1 + 7) / 8, NCB = (e2 + 7) / 8;
    const int t1 = (max_n1 + 8 * NRB - 1) / (8 * NRB), t2 = (max_n2 + 8 * NCB - 1) / (8 * NCB);
    if ((int64_t)t1 * t2 <= 65535) {
      dim3 grid((unsigned)(t1 * t2), B);
      const int nth = 32 * NRB + 32;   // compute warps + the norm warp
      const int R = 8 * NRB + 8 * NCB;
      const size_t smem = ((size_t)PREP_STAGES * R * PREP_LD + R + (size_t)R * PREP_HEAD + R) * sizeof(double);
      using kern_t = void (*)(const RbProblem*, RbAssocSoA, const double*, const double*, RbDevParams, int, int);
      // up to 5 compute warps (maps of <= 40 objects, the reference's submap_max_size): 192 threads, four CTAs per SM
#define RB_PREP_K(n) rb_prep_table_kernel<n, 192, RB_PREP_MINB>, rb_prep_table_kernel<n, 288, 2>
      static const kern_t kerns[PREP_MAXCB][2] = {{RB_PREP_K(1)}, {RB_PREP_K(2)}, {RB_PREP_K(3)}, {RB_PREP_K(4)},
                                                  {RB_PREP_K(5)}, {RB_PREP_K(6)}, {RB_PREP_K(7)}, {RB_PREP_K(8)}};
#undef RB_PREP_K
      static bool attr_set = false;
      if (!attr_set) {
        for (int i = 0; i < PREP_MAXCB; ++i)
          for (int j = 0; j < 2; ++j) {
            cudaError_t e = cudaFuncSetAttribute(kerns[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
            if (e != cudaSuccess) return e;
          }
        attr_set = true;
      }
      kerns[NCB - 1][nth <= 192 ? 0 : 1]<<<grid, nth, smem, st>>>(probs, rec, D1, D2, P, t2, NRB);
      return cudaGetLastError();
    }
  }
  dim3 grid((max_mp + 7) / 8, B);
  if (grid.x > 512) grid.x = 512;
  rb_prep_kernel<<<grid, 256, 0, st>>>(probs, rec, D1, D2, A, P);
  return cudaGetLastError();
}

cudaError_t rb_launch_affinity(const RbProblem* probs, int B, int64_t tile_base, int64_t total_tiles, int max_nb,
                               const RbAssocSoA& rec, uint32_t* masks, double* Mp, uint8_t* flags,
                               unsigned* edge_ij, double* edge_v, int* edge_cnt, const RbDevParams& P, int sm_count,
                               cudaStream_t st) {
  if (total_tiles == 0 || B == 0) return cudaSuccess;
  if (max_nb > 255) return cudaErrorInvalidValue;
  {
    dim3 g((max_nb + 7) / 8, B);
    rb_block_mask_kernel<<<g, 256, 0, st>>>(probs, rec, masks, P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (flags) {  // tile-skip format: dead tiles are never visited, so their flags are cleared up front
    cudaError_t e = cudaMemsetAsync(flags + tile_base, 0, (size_t)total_tiles, st);
    if (e != cudaSuccess) return e;
  }
  if (edge_cnt) {
    cudaError_t e = cudaMemsetAsync(edge_cnt, 0, (size_t)B * sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  // CTAs per problem: enough warps to cover the largest problem's tiles a few per warp, and at least
  // ~8 CTAs per SM over the whole launch when there are few problems
  const int64_t nt_max = (int64_t)max_nb * (max_nb + 1) / 2;
  int64_t per = (nt_max + 8 * 8 - 1) / (8 * 8);
  const int64_t fill = ((int64_t)sm_count * 8 + B - 1) / B;
  if (per < fill) per = fill;
  const int64_t cap = (nt_max + 7) / 8;
  if (per > cap) per = cap;
  if (per < 1) per = 1;
  dim3 grid((unsigned)per, (unsigned)B);
  rb_affinity_kernel<<<grid, 256, 0, st>>>(probs, rec, masks, Mp, flags, edge_ij, edge_v, edge_cnt, P);
  return cudaGetLastError();
}

cudaError_t rb_launch_unpack(const RbProblem& pr, const double* Mp, const uint8_t* flags, double* Mout, double* Cout,
                             cudaStream_t st) {
  if (pr.m == 0) return cudaSuccess;
  const int64_t n = (int64_t)pr.m * pr.m;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  rb_unpack_kernel<<<grid, 256, 0, st>>>(pr, Mp, flags, Mout, Cout);
  return cudaGetLastError();
}

cudaError_t rb_launch_pack(const RbProblem& pr, const double* Min, const double* Cin, double* Mp, int* bad,
                           cudaStream_t st) {
  if (pr.m == 0) return cudaSuccess;
  const int64_t n = pr.n_tiles * RB_TILE_ELEMS;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  rb_pack_kernel<<<grid, 256, 0, st>>>(pr, Min, Cin, Mp, bad);
  return cudaGetLastError();
}

// rb_prefilter.cu — submap-descriptor prefilter of the all-pairs search.
//
// Reference: Submap.similarity (roman/map/map.py:145-162), evaluated once per submap pair inside
// the pair loop (roman/align/submap_align.py:131-134) and compared with submap_descriptor_thresh
// (:144): the maximum over all (frame a of submap i, frame b of submap j) of the cosine of their
// descriptors, a term being 0 when the product of the two norms is within 1e-9 of zero.
//
// One CTA per submap pair; a warp owns one frame pair at a time and sweeps the two descriptors
// once, accumulating the dot product and both squared norms (lanes stride the descriptor ->
// coalesced 256-byte requests; the few-hundred-KB tables stay in L2 across CTAs).  The table is
// small (S x frames x D doubles), so no staging beyond that is worth it.
#include <math.h>

#include "rb_kernels.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) rb_similarity_kernel(int D, const int64_t* __restrict__ off0,
                                                                 const int64_t* __restrict__ off1,
                                                                 const double* __restrict__ F0,
                                                                 const double* __restrict__ F1,
                                                                 double* __restrict__ sim, int S1) {
  const int j = blockIdx.x, i = blockIdx.y;
  const int64_t a0 = off0[i], na = off0[i + 1] - a0;
  const int64_t b0 = off1[j], nb = off1[j + 1] - b0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kThreads / 32;
  double best = -INFINITY;
  int saw_nan = 0;
  for (int64_t p = warp; p < na * nb; p += nw) {
    const double* x = F0 + (a0 + p / nb) * (int64_t)D;
    const double* y = F1 + (b0 + p % nb) * (int64_t)D;
    double sxy = 0.0, sxx = 0.0, syy = 0.0;
    for (int k = lane; k < D; k += 32) {
      const double xv = x[k], yv = y[k];
      sxy = fma(xv, yv, sxy);
      sxx = fma(xv, xv, sxx);
      syy = fma(yv, yv, syy);
    }
    sxy = rb_warp_sum(sxy);
    sxx = rb_warp_sum(sxx);
    syy = rb_warp_sum(syy);
    const double norm_prod = sqrt(sxx) * sqrt(syy);
    double s = sxy / norm_prod;
    if (fabs(norm_prod) <= 1e-9) s = 0.0;  // np.isclose(norm_prod, 0, atol=1e-9, rtol=0)
    if (s != s) saw_nan = 1;               // np.max propagates a NaN
    else best = fmax(best, s);
  }
  __shared__ double s_best[kThreads / 32];
  __shared__ int s_nan[kThreads / 32];
  if (lane == 0) { s_best[warp] = best; s_nan[warp] = saw_nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < nw; ++w) { best = fmax(best, s_best[w]); saw_nan |= s_nan[w]; }
    sim[(int64_t)i * S1 + j] = saw_nan ? nan("") : best;
  }
}

}  // namespace

cudaError_t rb_launch_similarity(int S0, int S1, int D, const int64_t* off0, const int64_t* off1, const double* F0,
                                 const double* F1, double* sim, cudaStream_t st) {
  if (S0 <= 0 || S1 <= 0) return cudaSuccess;
  dim3 grid((unsigned)S1, (unsigned)S0);
  rb_similarity_kernel<<<grid, kThreads, 0, st>>>(D, off0, off1, F0, F1, sim, S1);
  return cudaGetLastError();
}

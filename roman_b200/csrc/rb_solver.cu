// rb_solver.cu — K2: persistent CLIPPER dense-clique solver, one thread-block cluster per problem.
//
// Replaces CLIPPER::findDenseClique (clipper/src/clipper.cpp:243-394), utils::findIndicesOfkLargest
// (clipper/src/utils.cpp:33-55) and utils::selectInlierAssociations (utils.cpp:101-108).
//
// Shape of the kernel
//   - the grid is a set of thread-block clusters of C CTAs (C = 1, 2, 4 or 8; 8 warps per CTA, one
//     CTA per SM); clusters pull problems from an atomic queue, so the SMs stream different
//     problems' packed affinity matrices from HBM: the working set is far larger than L2 and
//     the kernel is HBM-bound by construction.  C > 1 splits the tiles of one problem over the
//     CTAs of its cluster (balanced pairs of tile-rows per CTA) and combines the per-CTA partial
//     (M x, C x) vectors through distributed shared memory once per pass; it exists to shorten
//     the tail that a slow-converging problem would otherwise leave on a single SM.
//     Every CTA of a cluster then runs the (cheap) vector phases redundantly on identical data,
//     so all scalar decisions agree without any further exchange.
//   - one "pass" = one fused evaluation of (M x, C x) for the current trial vector x:
//     every 32x32 tile (8 KB, contiguous) of the packed upper triangle is brought into shared
//     memory by one cp.async.bulk (TMA) and used twice — tile * x_J feeds the row sums of block
//     I, tile^T * x_I feeds the column sums of block J — so each byte of M is read once per pass.
//     C is never stored: C_ij = [stored word != +0.0].
//   - each warp runs its own 2-slot TMA ring (issue by lane 0, completion on an mbarrier) over
//     the tiles (I, J) with J mod 8 == warp, which makes warp (K mod 8) the only writer of the
//     output block K: no atomics, fixed summation order, bit-reproducible results.
//     Tiles are stored XOR-swizzled, so a warp forms a tile's row sums (lane = row) and column
//     sums (lane = column) straight from shared memory; the 8 warps' row partials of a tile-row
//     are combined once per tile-row through a small ring, on arrive-only named barriers.
//   - all scalar control flow of the reference (line search, convergence tests, the d update)
//     is resolved on the device from block-wide reductions with a fixed tree.
//   - M x and C x of the accepted iterate are kept, so the d update and the new gradient need
//     no extra pass (the reference recomputes them; the values are identical).
#include <cooperative_groups.h>

#include "rb_kernels.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NW = 8;         // warps per CTA (all are consumers; each runs its own TMA ring)
constexpr int NT = NW * 32;   // threads per CTA
constexpr int MAX_SLOTS = 2;  // TMA ring slots per warp (8 KB each); 1 when m is too large for 2
constexpr int MAX_DEPTH = 8;  // see Smem::depth
constexpr int RED_DOUBLES = 64;   // [0,32): double-buffered reduction scratch; [40]: broadcast slot
constexpr int MISC_INTS = 16;
constexpr int EPT_MAX = 20;       // max vector elements per thread (mp_max <= EPT_MAX * NT)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_tile(void* dst, const void* src, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order prior generic reads of dst
  mbar_expect_tx(bar, RB_TILE_BYTES);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"((uint32_t)RB_TILE_BYTES), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}


struct Smem {
  double* x;     // trial vector (mp)
  double* accM;  // M x (strict), then reused as int scratch in rounding
  double* accC;  // C x (strict)
  double* part;  // [depth][NW][2][32] row partials, ring over tile-rows
  double* red;   // [2][2*NW] reduction scratch, double buffered
  uint64_t* bar; // [NW][SLOTS]
  double* ring;  // [NW][SLOTS][1024]
  int* misc;     // small ints / broadcast slots
  const uint8_t* flag;    // this problem's tile flags (shared-memory copy), or null: stream every tile
  const uint8_t* rowany;  // per tile-row: any flagged tile (null with flag == null)
  int slots;     // TMA ring slots per warp (1 or 2)
  int depth;     // tile-rows in flight between the fastest and slowest warp: 8 (decoupled) or 2 (lock-step)
};

struct Reducer {
  double* red;
  int flip;
  // sum of (a, b) over the CTA, same value in every thread, fixed order; one barrier.
  __device__ __forceinline__ void sum2(double& a, double& b) {
    a = rb_warp_sum(a);
    b = rb_warp_sum(b);
    double* r = red + flip * (2 * NW);
    if ((threadIdx.x & 31) == 0) {
      r[threadIdx.x >> 5] = a;
      r[NW + (threadIdx.x >> 5)] = b;
    }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      sa += r[w];
      sb += r[NW + w];
    }
    a = sa;
    b = sb;
    flip ^= 1;
  }
  __device__ __forceinline__ double sum1(double a) {
    double b = 0.0;
    sum2(a, b);
    return a;
  }
};

// One fused pass: accM += M x, accC += C x over the packed upper triangle of one problem.
// Preconditions: x complete, accM/accC zeroed, all threads past a __syncthreads.
//
// Work split.  Tile-rows are dealt to the C CTAs of the cluster in balanced pairs (tile-row p
// has nb-p tiles, tile-row nb-1-p has p+1: a pair always has nb+1), pair q going to CTA q mod C;
// inside a CTA the tiles of a tile-row are dealt to the 8 warps by tile column, J mod 8.  The
// results are this CTA's PARTIAL sums (combined by cluster_combine afterwards).  Warp (K mod 8)
// is the only writer of output block K in its CTA: it owns both the column sums of tile column
// K and the final add of the row sums of tile-row K, so there are no atomics and the summation
// order is fixed.
//
// A tile sits in shared memory in the XOR-swizzled layout of rb_common.cuh, so one warp sweeps
// it twice without bank conflicts: with lane = row (row sums of block I, one register per lane,
// no cross-lane reduction) and with lane = column (column sums of block J).
struct TileIter {
  int nb, C, w, p, h, I, J;
  const uint8_t* F;  // tile flags or null
  bool ok;
  __device__ __forceinline__ bool live(int i, int j) const { return F == nullptr || F[rb_tile_index(i, j, nb)] != 0; }
  __device__ __forceinline__ int row() const { return h == 0 ? p : nb - 1 - p; }
  // position on the first tile (J >= I, J mod NW == w) at or after tile-row slot (p, h)
  __device__ __forceinline__ void settle() {
    const int npairs = (nb + 1) >> 1;
    while (p < npairs) {
      if (!(h == 1 && nb - 1 - p == p)) {  // the middle tile-row of an odd nb is visited once
        I = row();
        J = I + ((w - I) % NW + NW) % NW;
        while (J < nb && !live(I, J)) J += NW;
        if (J < nb) { ok = true; return; }
      }
      if (h == 0) h = 1; else { h = 0; p += C; }
    }
    ok = false;
  }
  __device__ __forceinline__ void init(int nb_, int C_, int crank, int w_, const uint8_t* F_) {
    nb = nb_; C = C_; w = w_; p = crank; h = 0; I = 0; J = 0; F = F_;
    settle();
  }
  __device__ __forceinline__ void next() {
    J += NW;
    while (J < nb && !live(I, J)) J += NW;
    if (J < nb) return;
    if (h == 0) h = 1; else { h = 0; p += C; }
    settle();
  }
};

__device__ __forceinline__ void matvec_pass(const Smem& S, const double* __restrict__ tiles, int nb,
                                            uint32_t& phase_bits, int C, int crank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int SLOTS = S.slots;
  uint64_t* bar = S.bar + w * MAX_SLOTS;
  double* ring = S.ring + (size_t)w * SLOTS * RB_TILE_ELEMS;

  TileIter pf;  // prefetch iterator: runs SLOTS tiles ahead of the consumer along the same sequence
  pf.init(nb, C, crank, w, S.flag);
  auto pf_issue = [&](int slot) {
    if (lane == 0)
      tma_load_tile(ring + slot * RB_TILE_ELEMS, tiles + rb_tile_index(pf.I, pf.J, nb) * RB_TILE_ELEMS, bar + slot);
    pf.next();
  };
  for (int s = 0; s < SLOTS; ++s)
    if (pf.ok) pf_issue(s);

  int slot = 0, pos = 0;
  const int npairs = (nb + 1) >> 1;
  for (int p = crank; p < npairs; p += C) {
    for (int h = 0; h < 2; ++h) {
      const int I = (h == 0) ? p : nb - 1 - p;
      if (h == 1 && I == p) break;
      if (S.rowany && !S.rowany[I]) continue;  // no flagged tile in this tile-row: nothing to add, for any warp
      // row sums of block I over this warp's tiles: lane = row; two chains each to shorten the
      // dependent-add latency
      double rM0 = 0.0, rM1 = 0.0, rC0 = 0.0, rC1 = 0.0;
      const double* xI = S.x + I * RB_TILE;

      for (int J = I + ((w - I) % NW + NW) % NW; J < nb; J += NW) {
        if (S.flag && !S.flag[rb_tile_index(I, J, nb)]) continue;  // not stored / all zero: never fetched
        mbar_wait(bar + slot, (phase_bits >> slot) & 1u);
        phase_bits ^= (1u << slot);
        const double* buf = ring + slot * RB_TILE_ELEMS;
        const double* xJ = S.x + J * RB_TILE;
        const double* rowp = buf + lane * RB_TILE;  // this lane's row (row orientation)
        // An all-(+0.0) tile contributes nothing to either product: find out with 16 wide loads
        // per lane and one vote, and skip the fp64 work (the bytes were still streamed, as always).
        unsigned long long any = 0ull;
        {
          const ulonglong2* rp2 = reinterpret_cast<const ulonglong2*>(rowp);
#pragma unroll
          for (int t = 0; t < RB_TILE / 2; ++t) {
            const ulonglong2 q = rp2[t ^ (lane & 15)];  // a permutation of the row's 16-byte chunks, conflict-free
            any |= q.x | q.y;
          }
        }
        if (!__any_sync(0xffffffffu, any != 0ull)) {
          // nothing to do
        } else if (J != I) {
          double cM0 = 0.0, cM1 = 0.0, cC0 = 0.0, cC1 = 0.0;
#pragma unroll
          for (int t = 0; t < RB_TILE; t += 2) {
            // row orientation: elements (lane, t), (lane, t+1); x_J[t], x_J[t+1] are broadcasts
            const double a0 = rowp[t ^ lane], a1 = rowp[(t + 1) ^ lane];
            const double2 xj = *reinterpret_cast<const double2*>(xJ + t);
            rM0 = fma(a0, xj.x, rM0);
            rM1 = fma(a1, xj.y, rM1);
            if (__double_as_longlong(a0) != 0ll) rC0 += xj.x;  // +0.0 <=> no edge (-0.0: constraint-only)
            if (__double_as_longlong(a1) != 0ll) rC1 += xj.y;
            // column orientation: elements (t, lane), (t+1, lane); x_I[t], x_I[t+1] are broadcasts
            const double b0 = buf[t * RB_TILE + (lane ^ t)], b1 = buf[(t + 1) * RB_TILE + (lane ^ (t + 1))];
            const double2 xi = *reinterpret_cast<const double2*>(xI + t);
            cM0 = fma(b0, xi.x, cM0);
            cM1 = fma(b1, xi.y, cM1);
            if (__double_as_longlong(b0) != 0ll) cC0 += xi.x;
            if (__double_as_longlong(b1) != 0ll) cC1 += xi.y;
          }
          // this warp is the only writer of output block J in this CTA
          S.accM[J * RB_TILE + lane] += cM0 + cM1;
          S.accC[J * RB_TILE + lane] += cC0 + cC1;
        } else {
          // diagonal tile holds both triangles: row sums only
#pragma unroll
          for (int t = 0; t < RB_TILE; t += 2) {
            const double a0 = rowp[t ^ lane], a1 = rowp[(t + 1) ^ lane];
            const double2 xj = *reinterpret_cast<const double2*>(xJ + t);
            rM0 = fma(a0, xj.x, rM0);
            rM1 = fma(a1, xj.y, rM1);
            if (__double_as_longlong(a0) != 0ll) rC0 += xj.x;
            if (__double_as_longlong(a1) != 0ll) rC1 += xj.y;
          }
        }
        __syncwarp();
        if (pf.ok) pf_issue(slot);  // refill the slot just drained
        slot = (slot + 1 == SLOTS) ? 0 : slot + 1;
      }

      // Combine the 8 warps' row partials of block I (fixed order) into this CTA's accumulators.
      // depth == 8: a warp only WAITS on the tile-row's named barrier when it has to — as the
      // owner of block I, or once every 8 tile-rows so that nobody runs more than 7 tile-rows
      // ahead of the slowest warp (the partial ring and the barrier ids are reused modulo 8) —
      // and otherwise just ARRIVES and moves on, so one warp's late tile does not stall the
      // CTA's seven other TMA rings.
      const int pslot = pos % S.depth;
      double* part = S.part + (size_t)pslot * (NW * 2 * RB_TILE);
      part[(w * 2 + 0) * RB_TILE + lane] = rM0 + rM1;
      part[(w * 2 + 1) * RB_TILE + lane] = rC0 + rC1;
      const bool owner = ((I % NW) == w);
      if (owner || S.depth < MAX_DEPTH || (pos % NW) == w) named_bar_sync(1 + pslot, NT);
      else named_bar_arrive(1 + pslot, NT);
      if (owner) {
        double sm = 0.0, sc = 0.0;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) {
          sm += part[(ww * 2 + 0) * RB_TILE + lane];
          sc += part[(ww * 2 + 1) * RB_TILE + lane];
        }
        S.accM[I * RB_TILE + lane] += sm;
        S.accC[I * RB_TILE + lane] += sc;
      }
      ++pos;
    }
  }
}

// Combine the per-CTA partial (accM, accC) of a cluster through distributed shared memory:
// reduce-scatter (CTA r sums slice r of all C partial vectors, in rank order) then all-gather
// (it stores the totals of its slice into every CTA's accumulators).  Every CTA ends up with
// bitwise identical totals.  Three cluster barriers: partials complete -> read; all reads done
// -> overwrite; all slices delivered.
__device__ __forceinline__ void cluster_combine(const Smem& S, int mp, int C, int crank) {
  if (C == 1) { __syncthreads(); return; }
  cg::cluster_group cluster = cg::this_cluster();
  const int L = ((mp / RB_TILE + C - 1) / C) * RB_TILE;  // slice length (multiple of 32)
  const int lo = crank * L, hi = (lo + L < mp) ? lo + L : mp;
  cluster.sync();
  constexpr int SL = 9;  // slice elements per thread: L <= SL * NT (mp <= 4608 at C = 2)
  double sm[SL], sc[SL];
#pragma unroll
  for (int k = 0; k < SL; ++k) { sm[k] = 0.0; sc[k] = 0.0; }
  for (int r = 0; r < C; ++r) {
    const double* rm = cluster.map_shared_rank(S.accM, r);
    const double* rc = cluster.map_shared_rank(S.accC, r);
#pragma unroll
    for (int k = 0; k < SL; ++k) {
      const int i = lo + (int)threadIdx.x + k * NT;
      if (i < hi) { sm[k] += rm[i]; sc[k] += rc[i]; }
    }
  }
  cluster.sync();
  for (int r = 0; r < C; ++r) {
    double* rm = cluster.map_shared_rank(S.accM, r);
    double* rc = cluster.map_shared_rank(S.accC, r);
#pragma unroll
    for (int k = 0; k < SL; ++k) {
      const int i = lo + (int)threadIdx.x + k * NT;
      if (i < hi) { rm[i] = sm[k]; rc[i] = sc[k]; }
    }
  }
  cluster.sync();
}

__global__ void __launch_bounds__(NT, 1) rb_solver_kernel(const RbSolveArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem S;
  {
    const size_t mp = (size_t)A.mp_max;
    unsigned char* p = smem_raw;
    S.slots = A.slots;
    S.depth = A.depth;
    S.ring = (double*)p; p += (size_t)NW * A.slots * RB_TILE_BYTES;
    S.x = (double*)p; p += mp * 8;
    S.accM = (double*)p; p += mp * 8;
    S.accC = (double*)p; p += mp * 8;
    S.part = (double*)p; p += (size_t)A.depth * NW * 2 * RB_TILE * 8;
    S.red = (double*)p; p += (size_t)RED_DOUBLES * 8;
    S.bar = (uint64_t*)p; p += (size_t)NW * MAX_SLOTS * 8;
    S.misc = (int*)p; p += (size_t)MISC_INTS * 4;
    S.flag = A.flags ? (const uint8_t*)p : nullptr; p += (size_t)((A.nt_max + 15) / 16) * 16;
    S.rowany = A.flags ? (const uint8_t*)p : nullptr;
  }
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NW * MAX_SLOTS; ++i) mbar_init(S.bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0u;  // bit s = parity the next wait on ring slot s must see
  Reducer R{S.red, 0};
  const rb_solver_params sp = A.sp;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const bool lead = (crank == 0);

  long long t_mv = 0, t_cb = 0, t_tot0 = clock64(), t_q = 0, t_round = 0;
  for (;;) {
    long long tq0 = clock64();
    // one problem per cluster: rank 0 pulls from the queue and posts the index to every CTA
    cluster.sync();
    if (lead && tid == 0) {
      const int nxt = atomicAdd(A.counter, 1);
      for (int r = 0; r < C; ++r) *cluster.map_shared_rank(S.misc, r) = nxt;
    }
    cluster.sync();
    const int b = S.misc[0];
    t_q += clock64() - tq0;
    if (b >= A.B) break;
    const RbProblem pr = A.probs[b];
    const int m = pr.m, nb = pr.nb, mp = nb * RB_TILE;
    const double* tiles = A.Mp + pr.off_tile * RB_TILE_ELEMS;
    double* u = A.u_work + (size_t)crank * A.pad_stride + pr.off_pad;  // per-CTA private copies
    double* g = A.g_work + (size_t)crank * A.pad_stride + pr.off_pad;
    const double* u0 = A.u0 + pr.off_assoc;
    int passes = 0;
    if (A.flags) {  // this problem's tile flags and per-tile-row summary into shared memory
      uint8_t* fl = const_cast<uint8_t*>(S.flag);
      uint8_t* ra = const_cast<uint8_t*>(S.rowany);
      const int nt = (int)pr.n_tiles;
      for (int t = tid; t < nt; t += NT) fl[t] = A.flags[pr.off_tile + t];
      __syncthreads();
      for (int I = tid; I < nb; I += NT) {
        uint8_t any = 0;
        const int64_t base = rb_tile_index(I, I, nb);
        for (int J = I; J < nb; ++J) any |= fl[base + (J - I)];
        ra[I] = any;
      }
      __syncthreads();
    }

    if (m == 0) {
      if (lead && tid == 0) {
        A.n_nodes[b] = 0; A.score[b] = 0.0; A.ifinal[b] = 0; A.n_passes[b] = 0;
      }
      continue;
    }

    // ---- init (clipper.cpp:263-269): u = M u0 + u0 (optional), u /= ||u||
    for (int i = tid; i < mp; i += NT) {
      S.x[i] = (i < m) ? u0[i] : 0.0;
      S.accM[i] = 0.0;
      S.accC[i] = 0.0;
    }
    __syncthreads();
    if (sp.rescale_u0) {
      { long long t0 = clock64(); matvec_pass(S, tiles, nb, phase, C, crank); long long t1 = clock64();
        ++passes;
        cluster_combine(S, mp, C, crank); t_mv += t1 - t0; t_cb += clock64() - t1; }
    }
    double nrm2 = 0.0;
    for (int i = tid; i < mp; i += NT) {
      const double v = sp.rescale_u0 ? (S.accM[i] + S.x[i]) : S.x[i];
      S.x[i] = v;
      nrm2 += v * v;
    }
    nrm2 = R.sum1(nrm2);
    double sum_u = 0.0;
    {
      const double nrm = sqrt(nrm2);
      for (int i = tid; i < mp; i += NT) {
        const double v = (i < m) ? S.x[i] / nrm : 0.0;
        S.x[i] = v;
        u[i] = v;
        sum_u += v;
        S.accM[i] = 0.0;
        S.accC[i] = 0.0;
      }
    }
    sum_u = R.sum1(sum_u);
    { long long t0 = clock64(); matvec_pass(S, tiles, nb, phase, C, crank);  // (M u, C u) of the normalised start vector
      long long t1 = clock64();
      ++passes;
      cluster_combine(S, mp, C, crank); t_mv += t1 - t0; t_cb += clock64() - t1; }

    // ---- initial d (clipper.cpp:271-280)
    double d = 0.0;
    {
      double cnt = 0.0, s = 0.0;
      for (int i = tid; i < m; i += NT) {
        const double ui = S.x[i];
        const double cbu = (sum_u - S.accC[i]) - ui;
        if (cbu > sp.eps && ui > sp.eps) { cnt += 1.0; s += (S.accM[i] + ui) / cbu; }
      }
      R.sum2(cnt, s);
      if (cnt > 0.0) d = s / cnt;
    }

    // ---- graduated projected gradient ascent (clipper.cpp:286-352)
    double F = 0.0;
    int i_ol;
    for (i_ol = 0; i_ol < sp.maxoliters; ++i_ol) {
      {  // gradF and F at (u, d) from the kept M u, C u  (clipper.cpp:290-291)
        double f = 0.0;
        for (int i = tid; i < m; i += NT) {
          const double ui = S.x[i];
          const double gi = (((1.0 + d) * ui - d * sum_u) + S.accM[i]) + S.accC[i] * d;
          g[i] = gi;
          f += ui * gi;
        }
        F = R.sum1(f);
      }
      for (int j_il = 0; j_il < sp.maxiniters; ++j_il) {
        double alpha = 1.0, Fnew = 0.0, deltaF = 0.0, du2 = 0.0, sum_un = 0.0;
        for (int k_ls = 0; k_ls < sp.maxlsiters; ++k_ls) {
          // unew = normalize(max(u + alpha*gradF, 0))   (clipper.cpp:306-308)
          double z = 0.0;
          for (int i = tid; i < m; i += NT) {
            const double t = u[i] + alpha * g[i];
            const double r = t > 0.0 ? t : 0.0;
            S.x[i] = r;
            z += r * r;
          }
          z = R.sum1(z);
          double su = 0.0;
          const double nrm = sqrt(z);
          for (int i = tid; i < mp; i += NT) {
            double v = S.x[i];
            if (z > 0.0 && i < m) { v = v / nrm; S.x[i] = v; }
            su += (i < m) ? v : 0.0;
            S.accM[i] = 0.0;
            S.accC[i] = 0.0;
          }
          sum_un = R.sum1(su);
          { long long t0 = clock64(); matvec_pass(S, tiles, nb, phase, C, crank); long long t1 = clock64();
            ++passes;
            cluster_combine(S, mp, C, crank); t_mv += t1 - t0; t_cb += clock64() - t1; }
          // gradFnew, Fnew (clipper.cpp:309-315) and, speculatively, ||unew - u||^2
          double f = 0.0, dd = 0.0;
          for (int i = tid; i < m; i += NT) {
            const double xi = S.x[i];
            const double gi = (((1.0 + d) * xi - d * sum_un) + S.accM[i]) + S.accC[i] * d;
            f += xi * gi;
            const double t = xi - u[i];
            dd += t * t;
          }
          R.sum2(f, dd);
          Fnew = f;
          du2 = dd;
          deltaF = Fnew - F;
          if (deltaF < -sp.eps) alpha = alpha * sp.beta;
          else break;
        }
        // accept (also after an exhausted line search, clipper.cpp:324-329)
        const double deltau = sqrt(du2);
        F = Fnew;
        sum_u = sum_un;
        for (int i = tid; i < m; i += NT) {
          const double xi = S.x[i];
          u[i] = xi;
          g[i] = (((1.0 + d) * xi - d * sum_un) + S.accM[i]) + S.accC[i] * d;
        }
        if (deltau < sp.tol_u || fabs(deltaF) < sp.tol_F) break;
      }
      // ---- increase d (clipper.cpp:339-351), from the kept M u, C u
      double cnt = 0.0, s = 0.0;
      for (int i = tid; i < m; i += NT) {
        const double ui = S.x[i];
        const double cbu = (sum_u - S.accC[i]) - ui;
        if (cbu > sp.eps && ui > sp.eps) { cnt += 1.0; s += fabs((S.accM[i] + ui) / cbu); }
      }
      R.sum2(cnt, s);
      if (cnt > 0.0) d += s / cnt;
      else break;
    }

    // ---- rounding (clipper.cpp:358-381) + selected associations (utils.cpp:101-108)
    __syncthreads();
    // every CTA of the cluster holds identical data; rank 0 alone produces the outputs
    if (!lead) continue;
    long long tr0 = clock64();
    int32_t* nodes = A.nodes + pr.off_assoc;
    int* rank = (int*)S.accM;  // accM is dead from here on
    int n_sel = 0;
    if (sp.rounding == 0) {  // NONZERO: utils::findIndicesWhereAboveThreshold(u, 0.0)
      if (tid == 0) {
        int n = 0;
        for (int i = 0; i < m; ++i)
          if (S.x[i] > 0.0) nodes[n++] = i;
        S.misc[1] = n;
      }
      __syncthreads();
      n_sel = S.misc[1];
    } else {  // DSD_HEU: omega = round(F); k largest of u with the reference's heap semantics
      const double Fr = round(F);  // half away from zero, like std::round
      int k = !(Fr >= 1.0) ? 0 : (Fr > (double)m ? m : (int)Fr);
      if (k >= 1) {
        double npos = 0.0;
        for (int i = tid; i < m; i += NT) npos += (S.x[i] > 0.0) ? 1.0 : 0.0;
        npos = R.sum1(npos);
        const bool only_pos = ((double)k <= npos);  // then no zero can be among the k largest
        // candidates = positive entries (all entries if zeros are needed); compacted so that the
        // O(n^2) ranking below runs over the support of u only
        int* cand = (int*)S.accC;  // accC is dead from here on
        if (tid == 0) S.misc[2] = 0;
        __syncthreads();
        for (int i = tid; i < m; i += NT) {
          rank[i] = m;  // sentinel: not ranked
          if (!only_pos || S.x[i] > 0.0) cand[atomicAdd(&S.misc[2], 1)] = i;
        }
        __syncthreads();
        const int nc = S.misc[2];
        // rank_i = #{j : (x_j, j) > (x_i, i)}  == position in "descending value, larger index first"
        for (int c = tid; c < nc; c += NT) {
          const int i = cand[c];
          const double xi = S.x[i];
          int rk = 0;
          for (int q = 0; q < nc; ++q) {
            const int j = cand[q];
            const double xj = S.x[j];
            rk += (xj > xi || (xj == xi && j > i)) ? 1 : 0;
          }
          if (rk == k - 1) S.red[40] = xi;  // value of the k-th largest (unique writer)
          rank[i] = rk;
        }
        __syncthreads();
        const double vstar = S.red[40];
        double a = 0.0, bcnt = 0.0;
        for (int i = tid; i < m; i += NT) {
          const double xi = S.x[i];
          a += (xi > vstar) ? 1.0 : 0.0;
          bcnt += (xi == vstar) ? 1.0 : 0.0;
        }
        R.sum2(a, bcnt);
        const int na = (int)a, nbq = (int)bcnt;
        if (na + nbq == k) {
          for (int i = tid; i < m; i += NT)
            if (rank[i] < k) nodes[rank[i]] = i;
        } else {
          // ties at the boundary value: reproduce the heap's choice (see DESIGN.md "top-k ties")
          for (int i = tid; i < m; i += NT)
            if (S.x[i] > vstar) nodes[rank[i]] = i;
          if (tid == 0) {
            int cnt_ge = 0, g_le = 0, istar = m - 1;
            for (int i = 0; i < m; ++i) {
              const double xi = S.x[i];
              if (xi >= vstar) {
                ++cnt_ge;
                if (xi > vstar) ++g_le;
                if (cnt_ge == k) { istar = i; break; }
              }
            }
            const int skip = na - g_le, t = k - na;
            int pos = 0, q = 0;
            for (int i = 0; i <= istar && q < t; ++i) {
              if (S.x[i] == vstar) {
                if (pos >= skip) { nodes[na + (t - 1 - q)] = i; ++q; }
                ++pos;
              }
            }
          }
        }
        __syncthreads();
      }
      n_sel = k;
    }
    // outputs
    int32_t* sel = A.sel + 2 * pr.off_assoc;
    for (int p = tid; p < n_sel; p += NT) {
      const int nd = nodes[p];
      sel[2 * p] = A.i1[pr.off_pad + nd];
      sel[2 * p + 1] = A.i2[pr.off_pad + nd];
    }
    if (A.u_out) {
      double* uo = A.u_out + pr.off_assoc;
      for (int i = tid; i < m; i += NT) uo[i] = S.x[i];
    }
    if (tid == 0) {
      A.n_nodes[b] = n_sel;
      A.score[b] = F;
      A.ifinal[b] = i_ol;
      A.n_passes[b] = passes;
    }
    t_round += clock64() - tr0;
  }
  if (A.debug && tid == 0) {
    long long* d = A.debug + (size_t)blockIdx.x * 8;
    d[0] = clock64() - t_tot0; d[1] = t_mv; d[2] = t_cb; d[3] = t_q; d[4] = t_round;
  }
}

}  // namespace

size_t rb_solver_smem_bytes(int mp_max, int slots, int depth) {
  const int nb = mp_max / RB_TILE;
  const size_t nt = (size_t)nb * (nb + 1) / 2;  // tile flags + per-tile-row summary (tile-skip mode)
  return (size_t)NW * slots * RB_TILE_BYTES + (size_t)3 * mp_max * 8 + (size_t)depth * NW * 2 * RB_TILE * 8 +
         (size_t)RED_DOUBLES * 8 + (size_t)NW * MAX_SLOTS * 8 + (size_t)MISC_INTS * 4 + ((nt + 15) / 16) * 16 +
         (size_t)((nb + 15) / 16) * 16;
}

// Largest shared-memory configuration that fits: (2 slots, depth 8) > (2, 2) > (1, 2).
bool rb_solver_config(int mp_max, size_t optin, int* slots, int* depth) {
  static const int cfg[3][2] = {{2, MAX_DEPTH}, {2, 2}, {1, 2}};
  if (mp_max > EPT_MAX * NT) return false;
  for (auto& c : cfg)
    if (rb_solver_smem_bytes(mp_max, c[0], c[1]) <= optin) { *slots = c[0]; *depth = c[1]; return true; }
  return false;
}

static size_t g_max_optin = 0;

cudaError_t rb_solver_prepare(int device) {
  int v = 0;
  cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e != cudaSuccess) return e;
  g_max_optin = (size_t)v;
  return cudaFuncSetAttribute(rb_solver_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
}

static void fill_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cluster, size_t smem,
                        cudaStream_t st) {
  cfg = cudaLaunchConfig_t();
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// How many clusters of `cluster` CTAs can be co-resident (one CTA per SM at this smem size).
cudaError_t rb_solver_max_clusters(int mp_max, int cluster, int* out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  int slots = 1, depth = 2;
  if (!rb_solver_config(mp_max, g_max_optin, &slots, &depth)) return cudaErrorInvalidValue;
  fill_config(cfg, attr, cluster, cluster, rb_solver_smem_bytes(mp_max, slots, depth), nullptr);
  return cudaOccupancyMaxActiveClusters(out, rb_solver_kernel, &cfg);
}

cudaError_t rb_launch_solver(const RbSolveArgs& a, int max_clusters, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  const size_t smem = rb_solver_smem_bytes(a.mp_max, a.slots, a.depth);
  if (g_max_optin && smem > g_max_optin) return cudaErrorInvalidValue;
  if (a.mp_max > EPT_MAX * NT) return cudaErrorInvalidValue;
  int nclusters = a.B < max_clusters ? a.B : max_clusters;
  if (nclusters < 1) nclusters = 1;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(cfg, attr, nclusters * a.cluster, a.cluster, smem, st);
  return cudaLaunchKernelEx(&cfg, rb_solver_kernel, a);
}

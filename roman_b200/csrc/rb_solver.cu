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
//   - tile-skip format (rb_set_tile_skip): only tiles flagged as holding a non-zero are fetched
//     (per-warp tile lists built once per problem), and a problem whose edges fit shared memory is
//     solved "resident": a CSR over both directions is built in the space of the TMA rings — from
//     the edge list K1 wrote, or from one pass over the stored tiles — and every pass is a CSR
//     product out of shared memory, with u and g resident too.
//   - dense format, late residency (rb_set_solver_promote): a problem still iterating after N streamed passes reads its
//     tiles once more to extract the non-zero words (build_compact_dense) and finishes from the same shared-memory CSR;
//     slow-converging problems would otherwise stream tens of full passes alone at the end of the launch.
//   - misc[] slots: 0 queue index, 1 NONZERO count, 2 candidate counter, 4..12 tiles per warp +
//     live tile-rows, 14 triple counter, 15 resident vote; red[] slots: 0..31 reductions, 40 k-th
//     value, 48..51 scan scratch.
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
  uint16_t* wlist;        // tile-skip format: per warp, its stored tiles (I << 8 | J) in processing order; else null
  int wl_cap;             // entries per warp in wlist
  int* wcnt;              // [NW] tiles per warp, [NW] = number of live tile-rows of this CTA
  uint8_t* live_rows;     // the CTA's live tile-rows in processing order
  // compact (resident) form of this CTA's share of the matrix, used when its non-zeros fit shared memory
  int* c_rowptr;          // [mp+1] CSR row pointers over directed entries (both (i,j) and (j,i))
  uint16_t* c_col;        // [2*nnz] column index
  double* c_val;          // [2*nnz] value
  int c_cap;              // capacity in undirected entries (CSR built from K1's edge list)
  int c_cap_stage;        // capacity when the triples are staged in accM / accC first (extraction from tiles): <= mp
  double* c_ug;           // room for u and g (2 * mp doubles) next to the CSR, or null
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

// The vector phases of the solver walk the elements tid, tid + NT, ... of m-vectors.  Written naively every element is
// a chain of dependent shared-memory loads and fp64 arithmetic (a division is ~35 instructions) and the chains of a
// thread run one after the other — measured: 3.7 us per pass for four such loops over 7 elements per thread.  elems4
// evaluates FOUR elements at a time (`f`, independent of each other, so their latencies overlap) and then folds their
// results into the thread's accumulators in index order (`g`), which keeps every sum bit-identical to the plain loop.
template <typename V, typename F, typename G>
__device__ __forceinline__ void elems4(int n, F&& f, G&& g) {
  int i = threadIdx.x;
  for (; i + 3 * NT < n; i += 4 * NT) {
    const V a = f(i), b = f(i + NT), c = f(i + 2 * NT), d = f(i + 3 * NT);
    g(a); g(b); g(c); g(d);
  }
  for (; i < n; i += NT) g(f(i));
}
struct V2 { double a, b; };

// The same over an index list (the ACTIVE associations of a resident problem, see build_active); act == nullptr: all of
// 0 .. n-1 as above.
template <typename V, typename F, typename G>
__device__ __forceinline__ void elems_act(const uint16_t* __restrict__ act, int n, F&& f, G&& g) {
  if (act == nullptr) {
    elems4<V>(n, f, g);
    return;
  }
  int k = threadIdx.x;
  for (; k + 3 * NT < n; k += 4 * NT) {
    const int i0 = act[k], i1 = act[k + NT], i2 = act[k + 2 * NT], i3 = act[k + 3 * NT];
    const V a = f(i0), b = f(i1), c = f(i2), d = f(i3);
    g(a); g(b); g(c); g(d);
  }
  for (; k < n; k += NT) g(f((int)act[k]));
}

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
struct TileIter {  // dense format: the arithmetic tile sequence of one warp
  int nb, C, w, p, h, I, J;
  bool ok;
  __device__ __forceinline__ int row() const { return h == 0 ? p : nb - 1 - p; }
  // position on the first tile (J >= I, J mod NW == w) at or after tile-row slot (p, h)
  __device__ __forceinline__ void settle() {
    const int npairs = (nb + 1) >> 1;
    while (p < npairs) {
      if (!(h == 1 && nb - 1 - p == p)) {  // the middle tile-row of an odd nb is visited once
        I = row();
        J = I + ((w - I) % NW + NW) % NW;
        if (J < nb) { ok = true; return; }
      }
      if (h == 0) h = 1; else { h = 0; p += C; }
    }
    ok = false;
  }
  __device__ __forceinline__ void init(int nb_, int C_, int crank, int w_) {
    nb = nb_; C = C_; w = w_; p = crank; h = 0; I = 0; J = 0;
    settle();
  }
  __device__ __forceinline__ void next() {
    J += NW;
    if (J < nb) return;
    if (h == 0) h = 1; else { h = 0; p += C; }
    settle();
  }
};

struct RowAcc {  // row sums of one tile-row held by a warp: lane = row
  double rM0, rC0;
  __device__ __forceinline__ void clear() { rM0 = rC0 = 0.0; }
};

// One tile, already in shared memory (XOR-swizzled): row sums into `ra`, column sums into this
// warp's output block J.  Affinity matrices are sparse inside a tile as well (a few edges per 32x32
// block), so the tile is first reduced to bit masks — which words of my row are non-zero (lane = row),
// and, by a ballot transpose, which rows are non-zero in my column (lane = column) — and the fp64
// work then runs over the set bits only, in ascending index order.  A stored -0.0 (constraint-only
// entry of rb_set_dense) counts as present: it adds nothing to M x and x to C x.
__device__ __forceinline__ void process_tile(const Smem& S, const double* __restrict__ buf, int I, int J, int lane,
                                             RowAcc& ra) {
  const double* xI = S.x + I * RB_TILE;
  const double* xJ = S.x + J * RB_TILE;
  const double* rowp = buf + lane * RB_TILE;  // this lane's row (row orientation)
  unsigned rowmask = 0u;
  {
    const ulonglong2* rp2 = reinterpret_cast<const ulonglong2*>(rowp);
#pragma unroll
    for (int t = 0; t < RB_TILE / 2; ++t) {
      const int ch = t ^ (lane & 15);  // a permutation of the row's 16-byte chunks, conflict-free
      const ulonglong2 q = rp2[ch];
      // chunk ch holds physical words 2ch, 2ch+1 = logical columns (2ch)^lane, (2ch+1)^lane
      if (q.x != 0ull) rowmask |= 1u << ((2 * ch) ^ lane);
      if (q.y != 0ull) rowmask |= 1u << ((2 * ch + 1) ^ lane);
    }
  }
  unsigned cols = __reduce_or_sync(0xffffffffu, rowmask);  // columns holding any entry
  if (cols == 0u) return;
  {  // row sums: lane = row
    unsigned mk = rowmask;
    double sM = 0.0, sC = 0.0;
    while (mk) {
      const int t = __ffs(mk) - 1;
      mk &= mk - 1;
      const double v = rowp[t ^ lane];
      const double xj = xJ[t];
      sM = fma(v, xj, sM);
      sC += xj;
    }
    ra.rM0 += sM;
    ra.rC0 += sC;
  }
  if (J == I) return;  // diagonal tile holds both triangles: row sums only
  {  // column sums: lane = column; colmask bit r <=> (r, lane) is present
    unsigned colmask = 0u;
    while (cols) {
      const int t = __ffs(cols) - 1;
      cols &= cols - 1;
      const unsigned bt = __ballot_sync(0xffffffffu, (rowmask >> t) & 1u);
      if (lane == t) colmask = bt;
    }
    double sM = 0.0, sC = 0.0;
    while (colmask) {
      const int r = __ffs(colmask) - 1;
      colmask &= colmask - 1;
      const double v = buf[r * RB_TILE + (lane ^ r)];
      const double xi = xI[r];
      sM = fma(v, xi, sM);
      sC += xi;
    }
    // this warp is the only writer of output block J in this CTA
    S.accM[J * RB_TILE + lane] += sM;
    S.accC[J * RB_TILE + lane] += sC;
  }
}

// Combine the 8 warps' row partials of block I (fixed order) into this CTA's accumulators.
// depth == 8: a warp only WAITS on the tile-row's named barrier when it has to — as the owner of
// block I, or once every 8 tile-rows so that nobody runs more than 7 tile-rows ahead of the
// slowest warp (the partial ring and the barrier ids are reused modulo 8) — and otherwise just
// ARRIVES and moves on, so one warp's late tile does not stall the CTA's seven other TMA rings.
__device__ __forceinline__ void finish_tile_row(const Smem& S, int I, int pos, int lane, int w, const RowAcc& ra) {
  const int pslot = pos % S.depth;
  double* part = S.part + (size_t)pslot * (NW * 2 * RB_TILE);
  part[(w * 2 + 0) * RB_TILE + lane] = ra.rM0;
  part[(w * 2 + 1) * RB_TILE + lane] = ra.rC0;
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
}

// Dense format: every tile of the packed upper triangle is fetched.
__device__ __forceinline__ void matvec_pass_dense(const Smem& S, const double* __restrict__ tiles, int nb,
                                               uint32_t& phase_bits, int C, int crank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int SLOTS = S.slots;
  uint64_t* bar = S.bar + w * MAX_SLOTS;
  double* ring = S.ring + (size_t)w * SLOTS * RB_TILE_ELEMS;

  TileIter pf;  // prefetch iterator: runs SLOTS tiles ahead of the consumer along the same sequence
  pf.init(nb, C, crank, w);
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
      RowAcc ra;
      ra.clear();
      for (int J = I + ((w - I) % NW + NW) % NW; J < nb; J += NW) {
        mbar_wait(bar + slot, (phase_bits >> slot) & 1u);
        phase_bits ^= (1u << slot);
        process_tile(S, ring + slot * RB_TILE_ELEMS, I, J, lane, ra);
        __syncwarp();
        if (pf.ok) pf_issue(slot);  // refill the slot just drained
        slot = (slot + 1 == SLOTS) ? 0 : slot + 1;
      }
      finish_tile_row(S, I, pos, lane, w, ra);
      ++pos;
    }
  }
}

// Tile-skip format: each warp walks the list of ITS stored tiles (built once per problem by
// build_tile_lists from the tile flags), tile-row by tile-row over the CTA's live tile-rows.
__device__ __forceinline__ void matvec_pass_list(const Smem& S, const double* __restrict__ tiles, int nb,
                                              uint32_t& phase_bits) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int SLOTS = S.slots;
  uint64_t* bar = S.bar + w * MAX_SLOTS;
  double* ring = S.ring + (size_t)w * SLOTS * RB_TILE_ELEMS;
  const uint16_t* wl = S.wlist + (size_t)w * S.wl_cap;
  const int cnt = S.wcnt[w], nlive = S.wcnt[NW];

  int kpf = 0;
  auto pf_issue = [&](int slot) {
    const int e = wl[kpf++];
    if (lane == 0)
      tma_load_tile(ring + slot * RB_TILE_ELEMS, tiles + rb_tile_index(e >> 8, e & 255, nb) * RB_TILE_ELEMS, bar + slot);
  };
  for (int s = 0; s < SLOTS; ++s)
    if (kpf < cnt) pf_issue(s);

  int slot = 0, k = 0;
  for (int pos = 0; pos < nlive; ++pos) {
    const int I = S.live_rows[pos];
    RowAcc ra;
    ra.clear();
    while (k < cnt && (wl[k] >> 8) == I) {
      const int J = wl[k] & 255;
      mbar_wait(bar + slot, (phase_bits >> slot) & 1u);
      phase_bits ^= (1u << slot);
      process_tile(S, ring + slot * RB_TILE_ELEMS, I, J, lane, ra);
      __syncwarp();
      if (kpf < cnt) pf_issue(slot);
      slot = (slot + 1 == SLOTS) ? 0 : slot + 1;
      ++k;
    }
    finish_tile_row(S, I, pos, lane, w, ra);
  }
}

// Once per problem (tile-skip format): from the tile flags, the CTA's live tile-rows in
// processing order and, per warp, the list of its stored tiles (I << 8 | J) in that order.
__device__ __forceinline__ void build_tile_lists(const Smem& S, const uint8_t* __restrict__ gflags_global, int nb, int C,
                                                 int crank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // the problem's flags (one byte per stored-tile slot) are staged into shared memory first — accC is free
  // at this point — so that the per-tile-row scans below do not pay a global-memory latency each
  uint8_t* gflags = reinterpret_cast<uint8_t*>(S.accC);
  {
    const int nt = nb * (nb + 1) / 2;
    for (int t = threadIdx.x; t < nt; t += NT) gflags[t] = gflags_global[t];
    __syncthreads();
  }
  uint16_t* wl = S.wlist + (size_t)w * S.wl_cap;
  int cnt = 0, nlive = 0;
  const int npairs = (nb + 1) >> 1;
  for (int p = crank; p < npairs; p += C) {
    for (int h = 0; h < 2; ++h) {
      const int I = (h == 0) ? p : nb - 1 - p;
      if (h == 1 && I == p) break;
      const int64_t base = rb_tile_index(I, I, nb) - I;  // + J gives the tile index of (I, J)
      // any stored tile in tile-row I?  (uniform over the CTA: every warp scans the whole row)
      unsigned any = 0;
      for (int J = I + lane; J < nb; J += 32) any |= gflags[base + J];
      if (!__any_sync(0xffffffffu, any != 0)) continue;
      if (w == 0 && lane == 0) S.live_rows[nlive] = (uint8_t)I;
      ++nlive;
      // this warp's columns of the row: J = J0 + 8 t
      const int J0 = I + ((w - I) % NW + NW) % NW;
      for (int t0 = 0; J0 + NW * t0 < nb; t0 += 32) {
        const int J = J0 + NW * (t0 + lane);
        const bool st = (J < nb) && gflags[base + J] != 0;
        const unsigned mk = __ballot_sync(0xffffffffu, st);
        if (st) wl[cnt + __popc(mk & ((1u << lane) - 1u))] = (uint16_t)((I << 8) | J);
        cnt += __popc(mk);
      }
    }
  }
  if (lane == 0) S.wcnt[w] = cnt;
  if (w == 0 && lane == 0) S.wcnt[NW] = nlive;
}

// Inclusive scan of rowptr[1..mp] (counts -> row ends), block-wide; rowptr[0] stays 0.
__device__ __forceinline__ void csr_scan_rowptr(const Smem& S, int mp) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
  int* rowptr = S.c_rowptr;
  const int per = (mp + NT - 1) / NT;
  const int lo = 1 + tid * per, hi = (lo + per <= mp + 1) ? lo + per : mp + 1;
  int run = 0;
  for (int i = lo; i < hi; ++i) run += rowptr[i];
  int v = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += up;
  }
  int* wsum = reinterpret_cast<int*>(S.red + 48);  // 8 ints of scratch; no reduction is in flight here
  if (lane == 31) wsum[w] = v;
  __syncthreads();
  int acc = v - run;  // exclusive prefix inside the warp ...
#pragma unroll
  for (int ww = 0; ww < NW; ++ww)
    if (ww < w) acc += wsum[ww];  // ... plus the warps before
  for (int i = lo; i < hi; ++i) { acc += rowptr[i]; rowptr[i] = acc; }
  __syncthreads();
}

// Insertion sort of every CSR row by column: fixed summation order whatever order the entries arrived in.  Then the
// list of non-empty rows (most associations carry no edge at all): the passes walk that list only.  It lives (16-bit
// entries) in the first half of the cursor array, which the build no longer needs; its length goes to misc[13].
__device__ __forceinline__ void csr_sort_rows(const Smem& S, int mp) {
  const int* rowptr = S.c_rowptr;
  for (int i = threadIdx.x; i < mp; i += NT) {
    const int b = rowptr[i], e = rowptr[i + 1];
    for (int a = b + 1; a < e; ++a) {
      const uint16_t c = S.c_col[a];
      const double v = S.c_val[a];
      int q = a - 1;
      while (q >= b && S.c_col[q] > c) { S.c_col[q + 1] = S.c_col[q]; S.c_val[q + 1] = S.c_val[q]; --q; }
      S.c_col[q + 1] = c; S.c_val[q + 1] = v;
    }
  }
  // Non-empty rows, LONGEST FIRST (counting sort by length, capped at 63): thread t takes rows t, t + NT, ... of the
  // list, so every thread gets one of the NT longest rows, one of the next NT, ... — entries per thread end up within
  // one row length of each other (row lengths run from 1 to ~50; a plain row -> thread map had the slowest thread set
  // the pace of every pass).  The order inside a length class is arbitrary: which thread sums a row does not change
  // the sum.
  uint16_t* live = reinterpret_cast<uint16_t*>(S.c_rowptr + (mp + 1));   // first half of the cursor array (the second
                                                                            // half takes the active list, see build_active)
  int* hist = reinterpret_cast<int*>(S.red);   // 64 counters + 64 cursors; no reduction is in flight here
  for (int t = threadIdx.x; t < 128; t += NT) hist[t] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < mp; i += NT) {
    const int len = rowptr[i + 1] - rowptr[i];
    if (len > 0) atomicAdd(&hist[len < 63 ? len : 63], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 63; l >= 1; --l) { hist[64 + l] = acc; acc += hist[l]; }
    S.misc[13] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < mp; i += NT) {
    const int len = rowptr[i + 1] - rowptr[i];
    if (len > 0) live[atomicAdd(&hist[64 + (len < 63 ? len : 63)], 1)] = (uint16_t)i;
  }
  __syncthreads();
}

// Resident CSR straight from the edge list K1 wrote (HBM / L2; 12 bytes per edge), without staging: one pass counts the
// row lengths, a scan turns them into row pointers, a second pass scatters, rows are sorted by column.  With C > 1 a CTA
// takes the edges of ITS tile-rows only (the split of matvec_pass_dense) — or, with C == 1 passed for a cluster, all of
// them.  Returns false (all threads agree) if the entries exceed the capacity.
__device__ __forceinline__ bool build_csr_from_list(const Smem& S, int nb, int mp, const unsigned* __restrict__ e_ij,
                                                    const double* __restrict__ e_v, int e_n, int C, int crank) {
  const int tid = threadIdx.x;
  int* rowptr = S.c_rowptr;             // [mp + 1]
  int* cursor = S.c_rowptr + (mp + 1);  // [mp]
  int* n_mine = S.misc + 14;
  const int npairs = (nb + 1) >> 1;
  auto mine = [&](unsigned ij) -> bool {
    if (C == 1) return true;
    const int I = (int)(ij >> 16) / RB_TILE;
    const int pr_ = I < npairs ? I : nb - 1 - I;  // tile-row pair index (see the matvec work split)
    return (pr_ % C) == crank;
  };
  for (int i = tid; i <= mp; i += NT) rowptr[i] = 0;
  for (int i = tid; i < mp; i += NT) cursor[i] = 0;
  if (tid == 0) *n_mine = 0;
  __syncthreads();
  int cnt = 0;
  for (int e = tid; e < e_n; e += NT) {
    const unsigned ij = e_ij[e];
    if (mine(ij)) {
      atomicAdd(&rowptr[(ij >> 16) + 1], 1);
      atomicAdd(&rowptr[(ij & 0xffffu) + 1], 1);
      ++cnt;
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(n_mine, cnt);
  __syncthreads();
  if (*n_mine > S.c_cap) return false;
  csr_scan_rowptr(S, mp);
  for (int e = tid; e < e_n; e += NT) {
    const unsigned ij = e_ij[e];
    if (!mine(ij)) continue;
    const int i = ij >> 16, j = ij & 0xffffu;
    const double v = e_v[e];
    int p = rowptr[i] + atomicAdd(&cursor[i], 1);
    S.c_col[p] = (uint16_t)j; S.c_val[p] = v;
    p = rowptr[j] + atomicAdd(&cursor[j], 1);
    S.c_col[p] = (uint16_t)i; S.c_val[p] = v;
  }
  __syncthreads();
  csr_sort_rows(S, mp);
  return true;
}


// Second half of the resident build: the staged triples (accM: values, accC: i << 16 | j; count in misc[14]) become a
// CSR over both directions in the ring space.  Returns false (all threads agree) if they exceed the capacity.
__device__ __forceinline__ bool finish_compact(const Smem& S, int mp) {
  const int tid = threadIdx.x;
  const double* tv = S.accM;
  const unsigned* tij = reinterpret_cast<const unsigned*>(S.accC);
  const int cap = S.c_cap_stage;
  const int n = S.misc[14];
  if (n > cap) return false;

  // CSR over directed entries: count, scan, fill, sort rows by column
  int* rowptr = S.c_rowptr;                       // [mp + 1]
  int* cursor = S.c_rowptr + (mp + 1);            // [mp]
  for (int i = tid; i <= mp; i += NT) rowptr[i] = 0;
  for (int i = tid; i < mp; i += NT) cursor[i] = 0;
  __syncthreads();
  for (int e = tid; e < n; e += NT) {
    const unsigned ij = tij[e];
    atomicAdd(&rowptr[(ij >> 16) + 1], 1);
    atomicAdd(&rowptr[(ij & 0xffffu) + 1], 1);
  }
  __syncthreads();
  csr_scan_rowptr(S, mp);
  for (int e = tid; e < n; e += NT) {
    const unsigned ij = tij[e];
    const int i = ij >> 16, j = ij & 0xffffu;
    const double v = tv[e];
    int p = rowptr[i] + atomicAdd(&cursor[i], 1);
    S.c_col[p] = (uint16_t)j; S.c_val[p] = v;
    p = rowptr[j] + atomicAdd(&cursor[j], 1);
    S.c_col[p] = (uint16_t)i; S.c_val[p] = v;
  }
  __syncthreads();
  csr_sort_rows(S, mp);
  return true;
}


// ---- resident (compact) form --------------------------------------------------------------------
// In the tile-skip format a problem often has only a few hundred edges.  Then this CTA's share of
// the matrix is read from HBM exactly once: the stored tiles are streamed through the TMA rings a
// single time, their non-zero words collected as (i, j, v) triples, and a CSR over both directions
// is built in shared memory (in the space of the TMA rings, idle from then on).  Every later pass is
// a CSR product out of shared memory.  Rows are sorted by column, so the summation order is fixed.
// Returns false (all threads agree) if the triples do not fit; the caller then streams tiles.
__device__ __forceinline__ bool build_compact(const Smem& S, const double* __restrict__ tiles, int nb, int mp,
                                              uint32_t& phase_bits, const unsigned* __restrict__ e_ij,
                                              const double* __restrict__ e_v, int e_n, int C, int crank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
  const int SLOTS = S.slots;
  uint64_t* bar = S.bar + w * MAX_SLOTS;
  double* ring = S.ring + (size_t)w * SLOTS * RB_TILE_ELEMS;
  const uint16_t* wl = S.wlist + (size_t)w * S.wl_cap;
  const int cnt = S.wcnt[w];
  // staging area for the triples: accM (values) and accC (packed i << 16 | j), both unused until the first pass
  double* tv = S.accM;
  unsigned* tij = reinterpret_cast<unsigned*>(S.accC);
  const int cap = S.c_cap_stage;
  int* n_coo = S.misc + 14;
  if (tid == 0) *n_coo = 0;
  __syncthreads();

  // Fast path: K1 has already listed this problem's edges (count within capacity).  Then the tiles are not
  // read at all: this CTA takes its share of the list (the edges of its tile-rows) straight from HBM.
  const bool listed = (e_ij != nullptr) && e_n >= 0 && e_n <= mp;
  if (listed) {
    const int npairs = (nb + 1) >> 1;
    for (int e = tid; e < e_n; e += NT) {
      const unsigned ij = e_ij[e];
      bool mine = true;
      if (C > 1) {
        const int I = (int)(ij >> 16) / RB_TILE;
        const int pr_ = I < npairs ? I : nb - 1 - I;  // tile-row pair index (see matvec work split)
        mine = (pr_ % C) == crank;
      }
      if (mine) {
        const int pos = atomicAdd(n_coo, 1);
        if (pos < cap) { tv[pos] = e_v[e]; tij[pos] = ij; }
      }
    }
  }

  int kpf = 0, slot = 0;
  if (listed) kpf = cnt;  // nothing to stream
  auto pf_issue = [&](int sl) {
    const int e = wl[kpf++];
    if (lane == 0)
      tma_load_tile(ring + sl * RB_TILE_ELEMS, tiles + rb_tile_index(e >> 8, e & 255, nb) * RB_TILE_ELEMS, bar + sl);
  };
  for (int s2 = 0; s2 < SLOTS; ++s2)
    if (kpf < cnt) pf_issue(s2);
  for (int k = 0; k < (listed ? 0 : cnt); ++k) {
    const int I = wl[k] >> 8, J = wl[k] & 255;
    mbar_wait(bar + slot, (phase_bits >> slot) & 1u);
    phase_bits ^= (1u << slot);
    const double* rowp = ring + slot * RB_TILE_ELEMS + lane * RB_TILE;  // lane = row
#pragma unroll 4
    for (int t = 0; t < RB_TILE; ++t) {
      const double v = rowp[t ^ lane];
      // diagonal tiles hold both triangles: keep the strict upper one
      if (__double_as_longlong(v) != 0ll && (J != I || lane < t)) {
        const int pos = atomicAdd(n_coo, 1);
        if (pos < cap) {
          tv[pos] = v;
          tij[pos] = ((unsigned)(I * RB_TILE + lane) << 16) | (unsigned)(J * RB_TILE + t);
        }
      }
    }
    __syncwarp();
    if (kpf < cnt) pf_issue(slot);
    slot = (slot + 1 == SLOTS) ? 0 : slot + 1;
  }
  __syncthreads();
  return finish_compact(S, mp);
}

// Dense format, late residency ("promotion"): a problem that is still iterating after `promote_after` streamed passes
// reads its share of the tiles ONE more time — the same tile sequence as matvec_pass_dense — to pull out the non-zero
// words, and runs every later pass from the shared-memory CSR.  Slow-converging problems (tens of passes where the
// batch mean is ~5) otherwise pin a cluster to tens of full streams of a matrix that holds a few hundred edges.
// accM / accC must be free (they stage the triples); the caller re-zeroes them if this returns false.
__device__ __forceinline__ bool build_compact_dense(const Smem& S, const double* __restrict__ tiles, int nb, int mp,
                                                    uint32_t& phase_bits, int C, int crank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
  const int SLOTS = S.slots;
  uint64_t* bar = S.bar + w * MAX_SLOTS;
  double* ring = S.ring + (size_t)w * SLOTS * RB_TILE_ELEMS;
  double* tv = S.accM;
  unsigned* tij = reinterpret_cast<unsigned*>(S.accC);
  const int cap = S.c_cap_stage;
  int* n_coo = S.misc + 14;
  if (tid == 0) *n_coo = 0;
  __syncthreads();
  TileIter pf, it;
  pf.init(nb, C, crank, w);
  it.init(nb, C, crank, w);
  auto pf_issue = [&](int sl) {
    if (lane == 0)
      tma_load_tile(ring + sl * RB_TILE_ELEMS, tiles + rb_tile_index(pf.I, pf.J, nb) * RB_TILE_ELEMS, bar + sl);
    pf.next();
  };
  for (int s2 = 0; s2 < SLOTS; ++s2)
    if (pf.ok) pf_issue(s2);
  int slot = 0;
  while (it.ok) {
    const int I = it.I, J = it.J;
    mbar_wait(bar + slot, (phase_bits >> slot) & 1u);
    phase_bits ^= (1u << slot);
    const double* rowp = ring + slot * RB_TILE_ELEMS + lane * RB_TILE;  // lane = row
    unsigned rowmask = 0u;
    {
      const ulonglong2* rp2 = reinterpret_cast<const ulonglong2*>(rowp);
#pragma unroll
      for (int t = 0; t < RB_TILE / 2; ++t) {
        const int ch = t ^ (lane & 15);
        const ulonglong2 q = rp2[ch];
        if (q.x != 0ull) rowmask |= 1u << ((2 * ch) ^ lane);
        if (q.y != 0ull) rowmask |= 1u << ((2 * ch + 1) ^ lane);
      }
    }
    if (J == I) rowmask &= ~((2u << lane) - 1u);  // diagonal tiles hold both triangles: keep the strict upper one
    while (rowmask) {
      const int t = __ffs(rowmask) - 1;
      rowmask &= rowmask - 1;
      const int pos = atomicAdd(n_coo, 1);
      if (pos < cap) {
        tv[pos] = rowp[t ^ lane];
        tij[pos] = ((unsigned)(I * RB_TILE + lane) << 16) | (unsigned)(J * RB_TILE + t);
      }
    }
    __syncwarp();
    if (pf.ok) pf_issue(slot);
    slot = (slot + 1 == SLOTS) ? 0 : slot + 1;
    it.next();
  }
  __syncthreads();
  return finish_compact(S, mp);
}

// ACTIVE associations of a resident problem: those that carry an edge or whose u is not (yet) exactly zero, in index
// order, as 16-bit indices in the second half of the cursor array; returns their number.  An association without any
// edge has (M u)_i = (C u)_i = 0, so its gradient is (1 + d) u_i - d sum(u): once u_i has been clipped to zero it stays
// zero for good (max(0 + alpha * (-d sum u), 0) = 0 for d, sum u >= 0) and adds exactly 0 to every sum of the algorithm
// (u.g, ||unew||^2, sum unew, ||unew - u||^2, the d update's index set needs u_i > eps).  The all-pairs search is full of
// them — half of the candidate associations of two non-matching submaps fail the single-score gate — and the vector
// phases between two passes, not the passes, are what a resident problem spends its time in.  Whole matrix in this CTA
// only (C == 1 or solo).  Uses S.wcnt (idle while resident) as scratch; two barriers.
__device__ __forceinline__ int build_active(const Smem& S, int m, int mp) {
  const int* __restrict__ rowptr = S.c_rowptr;
  uint16_t* act = reinterpret_cast<uint16_t*>(S.c_rowptr + (mp + 1)) + mp;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int seg = ((m + NW * 32 - 1) / (NW * 32)) * 32;  // indices per warp
  const int lo = w * seg, hi = (lo + seg < m) ? lo + seg : m;
  int c = 0;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const bool keep = i < hi && (rowptr[i + 1] > rowptr[i] || S.x[i] != 0.0);
    c += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if (lane == 0) S.wcnt[w] = c;
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int v = 0; v < NW; ++v) {
    const int cv = S.wcnt[v];
    off += (v < w) ? cv : 0;
    tot += cv;
  }
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const bool keep = i < hi && (rowptr[i + 1] > rowptr[i] || S.x[i] != 0.0);
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    if (keep) act[off + __popc(mk & ((1u << lane) - 1u))] = (uint16_t)i;
    off += __popc(mk);
  }
  __syncthreads();
  return tot;
}

// One pass out of the resident CSR: accM = M x, accC = C x (this CTA's share).  accM / accC were zeroed by the caller;
// only rows that hold an entry are visited.  The four loads of an unrolled step are independent (the summation itself
// stays sequential over the sorted columns).  (Summing the long rows with eight lanes each and a shuffle tree was tried:
// 1.95 -> 2.67 us per pass on the all-pairs workload; the rows are not long enough to pay for the tree.)
__device__ __forceinline__ void matvec_pass_compact(const Smem& S, int mp) {
  const int* __restrict__ rowptr = S.c_rowptr;
  const uint16_t* __restrict__ live = reinterpret_cast<const uint16_t*>(S.c_rowptr + (mp + 1));
  const uint16_t* __restrict__ col = S.c_col;
  const double* __restrict__ val = S.c_val;
  const double* __restrict__ x = S.x;
  const int n_live = S.misc[13];
  // one row: sequential sum over its sorted columns, four independent loads per step
  // (two rows of a thread at a time, interleaved, was tried: no gain — the pace is set by the thread with most entries)
  auto row_tail = [&](int q, int e, double& sM, double& sC) {
    for (; q + 4 <= e; q += 4) {
      const int c0 = col[q], c1 = col[q + 1], c2 = col[q + 2], c3 = col[q + 3];
      const double v0 = val[q], v1 = val[q + 1], v2 = val[q + 2], v3 = val[q + 3];
      const double x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
      sM = fma(v0, x0, sM); sC += x0;
      sM = fma(v1, x1, sM); sC += x1;
      sM = fma(v2, x2, sM); sC += x2;
      sM = fma(v3, x3, sM); sC += x3;
    }
    for (; q < e; ++q) {
      const double xj = x[col[q]];
      sM = fma(val[q], xj, sM);
      sC += xj;
    }
  };
  // The list is sorted longest first and dealt to the threads in rounds of NT rows, alternately forwards and backwards
  // (thread 0 gets the longest row of round 0 and the shortest of round 1): with a plain round-robin the first threads
  // got the longest row of EVERY round and set the pace of the pass.
  int r = 0;
  for (int base = 0; base < n_live; base += NT, ++r) {
    const int t = base + ((r & 1) ? (NT - 1 - (int)threadIdx.x) : (int)threadIdx.x);
    if (t < n_live) {
      const int i = live[t];
      double sM = 0.0, sC = 0.0;
      row_tail(rowptr[i], rowptr[i + 1], sM, sC);
      S.accM[i] = sM;
      S.accC[i] = sC;
    }
  }
}

__device__ __forceinline__ void matvec_pass(const Smem& S, const double* __restrict__ tiles, int nb,
                                            uint32_t& phase_bits, int C, int crank, bool compact) {
  if (compact) matvec_pass_compact(S, nb * RB_TILE);
  else if (S.wlist) matvec_pass_list(S, tiles, nb, phase_bits);
  else matvec_pass_dense(S, tiles, nb, phase_bits, C, crank);
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
  // slice elements per thread: L = ceil(nb / C) * 32 <= ceil(EPT_MAX * NT / 32 / 2) * 32 = SL * NT for every C >= 2
  constexpr int SL = (EPT_MAX + 1) / 2;
  static_assert(SL * NT * 2 >= EPT_MAX * NT, "cluster_combine: slice registers must cover mp_max / 2");
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
    // Layout: vectors and small state first, then the TMA rings and the partial ring, then whatever is left of the
    // SM's shared memory (the launch always asks for the device maximum).  The resident CSR overlays everything from
    // the rings on: once a problem is resident neither rings nor partials are in use.
    const size_t mp = (size_t)A.mp_max;
    unsigned char* p = smem_raw;
    S.slots = A.slots;
    S.depth = A.depth;
    S.x = (double*)p; p += mp * 8;
    S.accM = (double*)p; p += mp * 8;
    S.accC = (double*)p; p += mp * 8;
    S.red = (double*)p; p += (size_t)RED_DOUBLES * 8;
    S.bar = (uint64_t*)p; p += (size_t)NW * MAX_SLOTS * 8;
    S.misc = (int*)p; p += (size_t)MISC_INTS * 4;
    S.wcnt = S.misc + 4;
    const int nbm = A.mp_max / RB_TILE;
    S.wl_cap = (A.nt_max + NW - 1) / NW + nbm;
    S.wlist = A.flags ? (uint16_t*)p : nullptr; p += (size_t)NW * S.wl_cap * 2;
    S.live_rows = (uint8_t*)p; p += (size_t)((nbm + 15) / 16) * 16;
    p = smem_raw + (((size_t)(p - smem_raw) + 127) / 128) * 128;
    S.ring = (double*)p; p += (size_t)NW * A.slots * RB_TILE_BYTES;
    S.part = (double*)p; p += (size_t)A.depth * NW * 2 * RB_TILE * 8;
    {
      // resident CSR: row pointers + cursors, then (if they fit) u and g, then values, then columns
      unsigned char* q = (unsigned char*)S.ring;
      const size_t region = (size_t)A.smem_total - (size_t)(q - smem_raw);
      S.c_rowptr = (int*)q;
      const size_t head = (((size_t)(2 * mp + 1) * 4 + 15) / 16) * 16;
      const size_t ug = (size_t)2 * mp * 8;
      size_t room = region > head ? region - head : 0;
      const bool ug_fits = room >= ug + 512 * 2 * (8 + 2);  // keep u, g in shared memory too if >= 512 edges still fit
      if (ug_fits) room -= ug;
      int cap = (int)(room / (2 * (8 + 2)));       // 2 directed entries of (double + uint16) per edge
      if (cap > 32767) cap = 32767;
      const bool any_resident = (A.flags || A.promote_after > 0);
      S.c_cap = any_resident ? cap : 0;
      S.c_cap_stage = any_resident ? (cap < (int)mp ? cap : (int)mp) : 0;  // triples staged in accM / accC: mp entries
      S.c_ug = ug_fits ? (double*)(q + head) : nullptr;
      unsigned char* v0 = q + head + (ug_fits ? ug : 0);
      S.c_val = (double*)v0;
      S.c_col = (uint16_t*)(v0 + (size_t)2 * cap * 8);
    }
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

  long long t_mv = 0, t_cb = 0, t_tot0 = clock64(), t_q = 0, t_round = 0, t_build = 0;
  for (;;) {
    long long tq0 = clock64();
    // one problem per cluster: rank 0 pulls from the queue and posts the index to every CTA
    cluster.sync();
    if (lead && tid == 0) {
      const int nxt = atomicAdd(A.counter, 1);
      for (int r = 0; r < C; ++r) *cluster.map_shared_rank(S.misc, r) = nxt;
    }
    cluster.sync();
    const int qpos = S.misc[0];
    t_q += clock64() - tq0;
    if (qpos >= A.B) break;
    const int b = A.order ? A.order[qpos] : qpos;  // longest problems first (rb_plan sorts by m)
    const RbProblem pr = A.probs[b];
    const int m = pr.m, nb = pr.nb, mp = nb * RB_TILE;
    const double* tiles = A.Mp + pr.off_tile * RB_TILE_ELEMS;
    double* u = A.u_work + (size_t)crank * A.pad_stride + pr.off_pad;  // per-CTA private copies (global scratch)
    double* g = A.g_work + (size_t)crank * A.pad_stride + pr.off_pad;
    const double* u0 = A.u0 + pr.off_assoc;
    int passes = 0;
    long long tb0 = clock64();
    if (m == 0) {
      if (lead && tid == 0) {
        A.n_nodes[b] = 0; A.score[b] = 0.0; A.ifinal[b] = 0; A.n_passes[b] = 0; A.resident[b] = 0; A.streamed[b] = 0;
      }
      continue;
    }
    // K1 lists every non-zero word of the strict upper triangle (12 bytes per edge, capacity RB_EDGE_FACTOR * mp per
    // problem; a count above the capacity means the list overflowed and is not used).
    const int e_n = (A.edge_cnt != nullptr) ? A.edge_cnt[pr.base - A.b0] : -1;  // the list belongs to the matrix's owner
    const bool have_list = (A.edge_ij != nullptr) && e_n >= 0 && e_n <= RB_EDGE_FACTOR * mp;
    const unsigned* e_ij = have_list ? A.edge_ij + (size_t)pr.off_rec * RB_EDGE_FACTOR : nullptr;
    const double* e_v = have_list ? A.edge_v + (size_t)pr.off_rec * RB_EDGE_FACTOR : nullptr;
    bool compact = false;   // passes run from the shared-memory CSR
    bool solo = false;      // C > 1 and every CTA holds the WHOLE matrix: nothing to combine, no cluster barrier per pass
    bool from_list = false;

    // all CTAs of a cluster must run the same kind of pass: unanimous vote over distributed shared memory
    auto cluster_all = [&](bool mine_ok) -> bool {
      if (C == 1) return mine_ok;
      cluster.sync();
      if (tid == 0) S.misc[15] = mine_ok ? 1 : 0;
      cluster.sync();
      bool all = true;
      for (int r = 0; r < C; ++r) all = all && (*cluster.map_shared_rank(S.misc + 15, r) != 0);
      cluster.sync();
      return all;
    };
    // resident CSR from K1's edge list: the whole matrix in every CTA if it fits (then the CTAs of a cluster run the
    // problem redundantly without exchanging anything), else each CTA's share of the tile-rows
    auto try_list = [&]() -> bool {
      if (!have_list || S.c_cap == 0) return false;
      if (e_n <= S.c_cap) {
        build_csr_from_list(S, nb, mp, e_ij, e_v, e_n, 1, 0);
        solo = (C > 1);
        return true;
      }
      if (C == 1) return false;
      return cluster_all(build_csr_from_list(S, nb, mp, e_ij, e_v, e_n, C, crank));
    };

    // tile-skip format: make the matrix resident up front (read from HBM once) — from the edge list, else from one
    // pass over this CTA's stored tiles; problems that fit neither way stream their stored tiles every pass
    if (A.flags) {
      compact = from_list = try_list();
      if (!compact) {
        build_tile_lists(S, A.flags + pr.off_tile, nb, C, crank);
        __syncthreads();
        compact = cluster_all(build_compact(S, tiles, nb, mp, phase, nullptr, nullptr, -1, C, crank));
      }
      if (compact && S.c_ug) {  // resident mode leaves room for the iterate and its gradient
        u = S.c_ug;
        g = S.c_ug + mp;
      }
    }

    t_build += clock64() - tb0;

    // one fused (M x, C x) evaluation for the vector currently in S.x, incl. the cluster-wide combine
    // (the cycle counters feed rb_debug_solver_cycles)
    int streamed = 0;      // passes that read the packed matrix from HBM
    bool promoted = false;
    auto run_pass = [&]() {
      // dense format, late residency: a problem still iterating after `promote_after` streamed passes finishes from the
      // shared-memory CSR — built from K1's edge list, else by one more pass over the tiles (build_compact_dense; accM /
      // accC were cleared for this pass, so they are free to stage the triples).
      if (!compact && !A.flags && A.promote_after > 0 && passes == A.promote_after) {
        const long long tb = clock64();
        bool ok = try_list();
        from_list = ok;
        if (!ok && S.c_cap_stage > 0) {
          ok = cluster_all(build_compact_dense(S, tiles, nb, mp, phase, C, crank));
          ++streamed;
        }
        if (ok && S.c_ug) {  // iterate and gradient move into shared memory with the matrix
          for (int i = tid; i < mp; i += NT) { S.c_ug[i] = (i < m) ? u[i] : 0.0; S.c_ug[mp + i] = (i < m) ? g[i] : 0.0; }
          u = S.c_ug;
          g = S.c_ug + mp;
        }
        for (int i = tid; i < mp; i += NT) { S.accM[i] = 0.0; S.accC[i] = 0.0; }
        __syncthreads();
        compact = promoted = ok;
        t_build += clock64() - tb;
      }
      const long long t0 = clock64();
      matvec_pass(S, tiles, nb, phase, C, crank, compact);
      const long long t1 = clock64();
      ++passes;
      if (!compact) ++streamed;
      if (solo) __syncthreads();
      else cluster_combine(S, mp, C, crank);
      t_mv += t1 - t0;
      t_cb += clock64() - t1;
    };

    // ---- init (clipper.cpp:263-269): u = M u0 + u0 (optional), u /= ||u||
    for (int i = tid; i < mp; i += NT) {
      S.x[i] = (i < m) ? u0[i] : 0.0;
      S.accM[i] = 0.0;
      S.accC[i] = 0.0;
    }
    __syncthreads();
    if (sp.rescale_u0) {
      run_pass();
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
    run_pass();  // (M u, C u) of the normalised start vector

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
    // (the __restrict__ views below live only inside one vector phase, between two barriers: the passes in between
    // write the same arrays through S)
#define RB_VIEWS double* __restrict__ const xs = S.x; double* __restrict__ const aM = S.accM; double* __restrict__ const aC = S.accC; (void)xs; (void)aM; (void)aC
    const uint16_t* act = nullptr;  // resident problems: the active associations (build_active), else all of 0 .. m-1
    int n_el = m;
    for (i_ol = 0; i_ol < sp.maxoliters; ++i_ol) {
      // (re)built once per outer iteration.  The conditions of build_active's argument: d >= 0 (it only grows from here),
      // sum(u) >= 0 (true for every later iterate: they are clipped at zero), eps >= 0 (so that u_i = 0 is never in the
      // d update's index set); a NaN fails the tests and keeps the full range.
      if (compact && (C == 1 || solo) && sp.eps >= 0.0 && d >= 0.0 && sum_u >= 0.0) {
        n_el = build_active(S, m, mp);
        act = reinterpret_cast<const uint16_t*>(S.c_rowptr + (mp + 1)) + mp;
      }
      {  // gradF and F at (u, d) from the kept M u, C u  (clipper.cpp:290-291)
        double f = 0.0;
        RB_VIEWS;
        double* __restrict__ const gp = g;
        elems_act<V2>(act, n_el, [&](int i) {
          const double ui = xs[i];
          const double gi = (((1.0 + d) * ui - d * sum_u) + aM[i]) + aC[i] * d;
          gp[i] = gi;
          return V2{ui * gi, 0.0};
        }, [&](const V2& v) { f += v.a; });
        F = R.sum1(f);
      }
      // (the accept step of an inner iteration that goes on already takes the next iteration's first trial step — same
      // thread, same elements, alpha = 1 — and hands over its partial sums)
      bool have_trial = false;
      double z_next = 0.0, sraw_next = 0.0;
      for (int j_il = 0; j_il < sp.maxiniters; ++j_il) {
        double alpha = 1.0, Fnew = 0.0, deltaF = 0.0, du2 = 0.0, sum_un = 0.0;
        for (int k_ls = 0; k_ls < sp.maxlsiters; ++k_ls) {
          // unew = normalize(max(u + alpha*gradF, 0))   (clipper.cpp:306-308)
          // (sum of the normalised vector = sum of the raw one / norm: one reduction for both)
          double z = 0.0, sraw = 0.0;
          if (have_trial) {
            z = z_next;
            sraw = sraw_next;
            have_trial = false;
          } else {
            RB_VIEWS;
            const double* __restrict__ const up = u;
            const double* __restrict__ const gp = g;
            elems_act<double>(act, n_el, [&](int i) {
              const double t = up[i] + alpha * gp[i];
              const double r = t > 0.0 ? t : 0.0;
              xs[i] = r;
              return r;
            }, [&](double r) { z += r * r; sraw += r; });
          }
          R.sum2(z, sraw);
          const double nrm = sqrt(z);
          sum_un = (z > 0.0) ? sraw / nrm : sraw;
          if (act) {
            // resident pass: it overwrites (M x)_i, (C x)_i of the rows that hold entries; the others stay zero
            RB_VIEWS;
            if (z > 0.0) elems_act<int>(act, n_el, [&](int i) { xs[i] = xs[i] / nrm; return 0; }, [](int) {});
          } else {
            RB_VIEWS;
            elems4<int>(mp, [&](int i) {
              if (z > 0.0 && i < m) xs[i] = xs[i] / nrm;
              aM[i] = 0.0;
              aC[i] = 0.0;
              return 0;
            }, [](int) {});
          }
          __syncthreads();
          run_pass();
          // gradFnew, Fnew (clipper.cpp:309-315) and, speculatively, ||unew - u||^2
          double f = 0.0, dd = 0.0;
          {
            RB_VIEWS;
            const double* __restrict__ const up = u;
            elems_act<V2>(act, n_el, [&](int i) {
              const double xi = xs[i];
              const double gi = (((1.0 + d) * xi - d * sum_un) + aM[i]) + aC[i] * d;
              const double t = xi - up[i];
              return V2{xi * gi, t * t};
            }, [&](const V2& v) { f += v.a; dd += v.b; });
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
        const bool converged = (deltau < sp.tol_u || fabs(deltaF) < sp.tol_F);
        if (converged || j_il + 1 >= sp.maxiniters) {
          RB_VIEWS;
          double* __restrict__ const up = u;
          double* __restrict__ const gp = g;
          elems_act<int>(act, n_el, [&](int i) {
            const double xi = xs[i];
            up[i] = xi;
            gp[i] = (((1.0 + d) * xi - d * sum_un) + aM[i]) + aC[i] * d;
            return 0;
          }, [](int) {});
        } else {
          RB_VIEWS;
          double* __restrict__ const up = u;
          double* __restrict__ const gp = g;
          const double alpha1 = 1.0;
          z_next = 0.0;
          sraw_next = 0.0;
          elems_act<double>(act, n_el, [&](int i) {
            const double xi = xs[i];
            const double gi = (((1.0 + d) * xi - d * sum_un) + aM[i]) + aC[i] * d;
            up[i] = xi;
            gp[i] = gi;
            const double t = xi + alpha1 * gi;
            const double r = t > 0.0 ? t : 0.0;
            xs[i] = r;
            return r;
          }, [&](double r) { z_next += r * r; sraw_next += r; });
          have_trial = true;
        }
        if (converged) break;
      }
      // ---- increase d (clipper.cpp:339-351), from the kept M u, C u
      double cnt = 0.0, s = 0.0;
      {
        RB_VIEWS;
        elems_act<V2>(act, n_el, [&](int i) {
          const double ui = xs[i];
          const double cbu = (sum_u - aC[i]) - ui;
          const bool in = cbu > sp.eps && ui > sp.eps;
          return V2{in ? 1.0 : 0.0, in ? fabs((aM[i] + ui) / cbu) : 0.0};
        }, [&](const V2& v) { if (v.a != 0.0) { cnt += 1.0; s += v.b; } });
      }
      R.sum2(cnt, s);
      if (cnt > 0.0) d += s / cnt;
      else break;
    }
#undef RB_VIEWS

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
      sel[2 * p] = A.i1[pr.off_rec + nd];
      sel[2 * p + 1] = A.i2[pr.off_rec + nd];
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
      A.resident[b] = promoted ? (from_list ? 4 : 3) : (compact ? (from_list ? 2 : 1) : 0);
      A.streamed[b] = streamed + ((compact && !promoted && !from_list) ? 1 : 0);  // + the one tile pass of build_compact
    }
    t_round += clock64() - tr0;
  }
  if (A.debug && tid == 0) {
    long long* d = A.debug + (size_t)blockIdx.x * 8;
    d[0] = clock64() - t_tot0; d[1] = t_mv; d[2] = t_cb; d[3] = t_q; d[4] = t_round;
    d[5] = t_build;
  }
}

}  // namespace

size_t rb_solver_smem_bytes(int mp_max, int slots, int depth) {
  const int nb = mp_max / RB_TILE;
  const size_t nt = (size_t)nb * (nb + 1) / 2;
  const size_t wl_cap = (nt + NW - 1) / NW + nb;  // tile-skip format: per-warp tile lists + live tile-rows
  // minimum: vectors, small state, tile lists, (alignment), rings, partial ring — see the carve-up in the kernel
  return (size_t)NW * slots * RB_TILE_BYTES + (size_t)3 * mp_max * 8 + (size_t)depth * NW * 2 * RB_TILE * 8 +
         (size_t)RED_DOUBLES * 8 + (size_t)NW * MAX_SLOTS * 8 + (size_t)MISC_INTS * 4 + (size_t)NW * wl_cap * 2 +
         (size_t)((nb + 15) / 16) * 16 + 128;
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
  // the launch always takes the device's whole opt-in shared memory (one CTA per SM either way): what the vectors
  // and rings leave over extends the resident CSR
  fill_config(cfg, attr, cluster, cluster, g_max_optin, nullptr);
  return cudaOccupancyMaxActiveClusters(out, rb_solver_kernel, &cfg);
}

cudaError_t rb_launch_solver(const RbSolveArgs& a_in, int max_clusters, cudaStream_t st) {
  if (a_in.B == 0) return cudaSuccess;
  RbSolveArgs a = a_in;
  size_t smem = rb_solver_smem_bytes(a.mp_max, a.slots, a.depth);
  if (g_max_optin && smem > g_max_optin) return cudaErrorInvalidValue;
  if (g_max_optin > smem) smem = g_max_optin;
  a.smem_total = (int32_t)smem;
  if (a.mp_max > EPT_MAX * NT) return cudaErrorInvalidValue;
  int nclusters = a.B < max_clusters ? a.B : max_clusters;
  if (nclusters < 1) nclusters = 1;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(cfg, attr, nclusters * a.cluster, a.cluster, smem, st);
  return cudaLaunchKernelEx(&cfg, rb_solver_kernel, a);
}

// rb_common.cuh — shared types, packed-M layout and small device helpers.
//
// Packed affinity layout (one per problem, in HBM):
//   the m x m symmetric affinity matrix is cut into 32x32 tiles of doubles (8 KB, contiguous).
//   Inside a tile, element (r, c) sits at word r*32 + (c ^ r): rows stay contiguous 256-byte
//   segments (coalesced stores in K1), and once the tile is in shared memory both "lane = row"
//   and "lane = column" sweeps are bank-conflict free, which lets the solver form row sums and
//   column sums of a tile without any cross-lane reduction.  Only tiles (I, J) with I <= J are stored, panel by panel:
//   panel I = tiles (I,I), (I,I+1) ... (I,nb-1).  Diagonal tiles hold the full symmetric
//   tile (both triangles, zero diagonal); everything beyond m is +0.0 padding.
//   tile_index(I,J) = I*nb - I*(I-1)/2 + (J-I);  n_tiles = nb*(nb+1)/2;  nb = ceil(m/32).
//   The constraint matrix is never stored: C_ij = 1  <=>  the stored word is not +0.0
//   (a stored -0.0 means "affinity 0, constraint inactive", used by rb_set_dense).
//   One tile = one cp.async.bulk (TMA) transfer in the solver.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/roman_b200.h"

#define RB_TILE 32
#define RB_TILE_ELEMS (RB_TILE * RB_TILE)
#define RB_TILE_BYTES (RB_TILE_ELEMS * 8)
// K1's per-problem edge list holds up to RB_EDGE_FACTOR * (padded m) non-zero words of the strict upper triangle
#define RB_EDGE_FACTOR 8

// One problem of the batch (device-resident array of these, built by rb_plan).
struct RbProblem {
  int32_t n1, n2, m, nb;  // nb = ceil(m/32)
  int32_t d, all_to_all;
  int64_t off_obj1, off_obj2;  // object offsets into D1 / D2 (rows of d doubles)
  int64_t off_assoc;           // offset into user-facing per-association arrays (sum of m)
  int64_t off_pad;             // offset into internal padded per-association arrays (sum of nb*32)
  int64_t off_tile;            // offset (in tiles) into the packed M buffer
  int64_t n_tiles;
  // Problems may SHARE an affinity matrix (multi-start: K start vectors on one submap pair, rb_plan_shared): `base` is
  // the batch index of the problem that owns the matrix, the association records and the edge list (== own index for
  // ordinary problems), off_rec its off_pad; K1 skips the copies, K2 / K3 run on every problem with private vectors.
  int32_t base, is_copy;
  int64_t off_rec;
};

// Internal SoA association records, padded arrays of length sum(nb*32).
struct RbAssocSoA {
  double *ax, *ay, *az;  // centre of the map-1 object of this association
  double *bx, *by, *bz;  // centre of the map-2 object
  double* single;        // single-association score (ROMAN::single_similarity), 1.0 for Euclidean
  int32_t *i1, *i2;      // object indices (A[k,0], A[k,1]); -1 in padding
};

// word offset of element (r, c) inside a tile (XOR swizzle, see above)
__host__ __device__ __forceinline__ int rb_tile_elem(int r, int c) { return r * RB_TILE + (c ^ r); }

__host__ __device__ __forceinline__ int64_t rb_tile_index(int I, int J, int nb) {
  return (int64_t)I * nb - ((int64_t)I * (I - 1)) / 2 + (J - I);
}

#ifdef __CUDACC__
__device__ __forceinline__ double rb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

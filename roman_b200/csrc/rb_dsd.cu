// rb_dsd.cu — Rounding::DSD: exact densest-subgraph rounding of the solver's support.
//
// Reference: clipper.cpp:364-371 (S = {i : u_i > 0}; nodes = dsd::solve(M_, S)) and clipper/src/dsd.cpp:157-314
// — Goldberg's flow-based exact algorithm (A. V. Goldberg, "Finding a maximum density subgraph", 1984):
// binary search on the density g; for every g one minimum s-t cut of the network
//      s -> v   capacity c0 = m/2                       for every vertex v
//      v -> t   capacity c0 + 2 g - deg(v)
//      u <-> v  capacity w(u, v)
// whose source side beyond {s} is non-empty exactly when some vertex set T has w(T) - g |T| > 0.
// What is taken from the reference are its CONSTANTS, because they fix which set is returned:
//   m  = |S|^2 - |S|  (it lists every ordered pair, zero weights included; dsd.cpp:276),
//   search interval [0, m/2], stop when n (n - 1) (U - L) < 1 with n = ALL nodes of the problem (:207-212),
//   the answer is the source side of the last non-trivial cut (the `L = g` branch, :223-228), ascending ids.
// The max-flow routine here is an ordinary Dinic (BFS level graph + blocking flow with current arcs) on residual
// capacities, antiparallel arcs merged; the source side of a minimum cut — vertices reachable from s in the residual
// graph — does not depend on which maximum flow was found.
//
// The rounding runs on the host, as in the reference: the solver kernel leaves the support of u (NONZERO rounding),
// rb_gather_sub_kernel extracts the |S| x |S| block of the packed affinity matrix, the host runs Goldberg per problem and
// rb_select_kernel writes the selected associations.  This path synchronises the stream (documented in the header).
#include <math.h>

#include <algorithm>
#include <vector>

#include "rb_kernels.cuh"

namespace {

struct FlowNet {
  int nv = 0;
  std::vector<int> head, nxt, to, level, cur, queue;
  std::vector<double> res;  // residual capacity of each arc; arc e and e^1 are a pair

  void reset(int vertices) {
    nv = vertices;
    head.assign((size_t)nv, -1);
    nxt.clear(); to.clear(); res.clear();
  }
  void add(int u, int v, double cap_uv, double cap_vu) {
    to.push_back(v); res.push_back(cap_uv); nxt.push_back(head[(size_t)u]); head[(size_t)u] = (int)to.size() - 1;
    to.push_back(u); res.push_back(cap_vu); nxt.push_back(head[(size_t)v]); head[(size_t)v] = (int)to.size() - 1;
  }
  bool levels(int s, int t) {
    level.assign((size_t)nv, -1);
    queue.clear();
    queue.push_back(s);
    level[(size_t)s] = 0;
    for (size_t qi = 0; qi < queue.size(); ++qi) {
      const int u = queue[qi];
      for (int e = head[(size_t)u]; e >= 0; e = nxt[(size_t)e])
        if (res[(size_t)e] > 0.0 && level[(size_t)to[(size_t)e]] < 0) {
          level[(size_t)to[(size_t)e]] = level[(size_t)u] + 1;
          queue.push_back(to[(size_t)e]);
        }
    }
    return level[(size_t)t] >= 0;
  }
  double push(int u, int t, double limit) {
    if (u == t) return limit;
    for (int& e = cur[(size_t)u]; e >= 0; e = nxt[(size_t)e]) {
      const int v = to[(size_t)e];
      if (res[(size_t)e] > 0.0 && level[(size_t)v] == level[(size_t)u] + 1) {
        const double got = push(v, t, std::min(limit, res[(size_t)e]));
        if (got > 0.0) {
          res[(size_t)e] -= got;
          res[(size_t)(e ^ 1)] += got;
          return got;
        }
      }
    }
    return 0.0;
  }
  void max_flow(int s, int t) {
    while (levels(s, t)) {
      cur = head;
      while (push(s, t, INFINITY) > 0.0) {}
    }
  }
  // vertices reachable from s in the residual graph (the minimal source side of a minimum cut)
  void source_side(int s, std::vector<char>& in) {
    in.assign((size_t)nv, 0);
    queue.clear();
    queue.push_back(s);
    in[(size_t)s] = 1;
    for (size_t qi = 0; qi < queue.size(); ++qi) {
      const int u = queue[qi];
      for (int e = head[(size_t)u]; e >= 0; e = nxt[(size_t)e])
        if (res[(size_t)e] > 0.0 && !in[(size_t)to[(size_t)e]]) {
          in[(size_t)to[(size_t)e]] = 1;
          queue.push_back(to[(size_t)e]);
        }
    }
  }
};

__global__ void rb_gather_sub_kernel(RbProblem pr, const double* __restrict__ Mp, const uint8_t* __restrict__ flags,
                                     const int32_t* __restrict__ nodes, int k, double* __restrict__ W) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)k * k;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(e / k), q = (int)(e - (int64_t)p * k);
    const int i = nodes[p], j = nodes[q];
    double v = 0.0;
    if (i != j) {
      const int a = i < j ? i : j, b = i < j ? j : i;
      const int64_t tl = pr.off_tile + rb_tile_index(a / RB_TILE, b / RB_TILE, pr.nb);
      if (!flags || flags[tl]) v = fabs(Mp[tl * RB_TILE_ELEMS + rb_tile_elem(a % RB_TILE, b % RB_TILE)]);
    }
    W[e] = v;
  }
}

__global__ void rb_select_kernel(const int32_t* __restrict__ nodes, int k, const int32_t* __restrict__ i1,
                                 const int32_t* __restrict__ i2, int32_t* __restrict__ sel) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < k; p += gridDim.x * blockDim.x) {
    sel[2 * p] = i1[nodes[p]];
    sel[2 * p + 1] = i2[nodes[p]];
  }
}

}  // namespace

extern "C" int rb_dsd_host(int64_t n, int32_t k, const int32_t* S, const double* W, int32_t* out, int32_t* n_out) {
  if (!n_out || k < 0 || n < k || (k > 0 && (!S || !W || !out))) return RB_ERR_ARG;
  *n_out = 0;
  if (k == 0) return RB_OK;
  if (k == 1) {  // the reference reads an uninitialised cut here (its search loop never runs); a single node is its own answer
    out[0] = S[0];
    *n_out = 1;
    return RB_OK;
  }
  const double m = (double)k * (double)k - (double)k;
  const double c0 = floor(m / 2.0);  // integer division in the reference
  std::vector<double> deg((size_t)k, 0.0);
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b)
      if (a != b) deg[(size_t)a] += (a < b) ? W[(size_t)a * k + b] : W[(size_t)b * k + a];  // upper triangle is authoritative
  FlowNet net;
  std::vector<char> side, best;
  const int s = k, t = k + 1;
  double L = 0.0, U = c0;
  const double nn = (double)n * (double)(n - 1);
  while (nn * (U - L) >= 1.0) {
    const double g = (U + L) / 2.0;
    net.reset(k + 2);
    for (int a = 0; a < k; ++a)
      for (int b = a + 1; b < k; ++b) {
        const double w = W[(size_t)a * k + b];
        if (w > 0.0) net.add(a, b, w, w);
      }
    for (int a = 0; a < k; ++a) {
      net.add(s, a, c0, 0.0);
      net.add(a, t, c0 + 2.0 * g - deg[(size_t)a], 0.0);
    }
    net.max_flow(s, t);
    net.source_side(s, side);
    int reached = 0;
    for (int a = 0; a < k; ++a) reached += side[(size_t)a];
    if (reached == 0) {
      U = g;
    } else {
      L = g;
      best = side;
    }
  }
  if (best.empty()) return RB_OK;  // no vertex set of positive density (edgeless support): the reference's result is undefined
  int cnt = 0;
  std::vector<std::pair<int32_t, int>> order;
  for (int a = 0; a < k; ++a)
    if (best[(size_t)a]) order.push_back({S[a], a});
  std::sort(order.begin(), order.end());  // ascending node ids, like the reference's scan over vertices
  for (auto& o : order) out[cnt++] = o.first;
  *n_out = cnt;
  return RB_OK;
}

cudaError_t rb_launch_gather_sub(const RbProblem& pr, const double* Mp, const uint8_t* flags, const int32_t* nodes, int k,
                                 double* W, cudaStream_t st) {
  if (k <= 0) return cudaSuccess;
  const int64_t n = (int64_t)k * k;
  const int grid = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  rb_gather_sub_kernel<<<grid, 256, 0, st>>>(pr, Mp, flags, nodes, k, W);
  return cudaGetLastError();
}

cudaError_t rb_launch_select(const int32_t* nodes, int k, const int32_t* i1, const int32_t* i2, int32_t* sel,
                             cudaStream_t st) {
  if (k <= 0) return cudaSuccess;
  rb_select_kernel<<<(k + 255) / 256, 256, 0, st>>>(nodes, k, i1, i2, sel);
  return cudaGetLastError();
}

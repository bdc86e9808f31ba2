// meshpost_core.cuh — __host__ __device__ core of the mesh post-processing that the reference's
// `integrate` program runs after marching cubes:
//
//   flattenVertices (src/prog/integrate.cpp:103-150)  weld vertices closer than min_dist
//   cleanupMesh     (src/prog/integrate.cpp:152-214)  drop faces in clusters of <= min_neighbors faces
//
// Both are radius-neighbourhood problems that the reference solves with a FLANN kd-tree and a
// sequential sweep.  Here the points go into a spatial hash (cell edge >= radius, per-cell lists) and
// the sequential sweeps are replaced by order-independent formulations with the same result:
//
// * flattenVertices visits vertices in index order; an unassigned vertex becomes a new output vertex
//   and (re)assigns every vertex within the radius.  The vertices that become output vertices are
//   therefore the lexicographically-first maximal independent set of the radius graph, and every
//   other vertex ends up on the LAST (highest-index) such vertex within its radius.  The set is
//   unique, so it can be computed by fix-point iteration in any order.
// * cleanupMesh removes the connected components (of the radius graph on face centroids) that have
//   at most min_neighbors faces.  A face with >= min_neighbors other faces in range is in a larger
//   component and only has to be recognised, not labelled; union-find runs on the sparse rest.
#pragma once
#include "tsdf_core.cuh"

namespace b2 {

B2_HD int atomic_exch_i32 (int* a, int v)
{
#ifdef __CUDA_ARCH__
  return atomicExch (a, v);
#else
  int o = *a; *a = v; return o;
#endif
}
B2_HD int atomic_cas_i32 (int* a, int cmp, int v)
{
#ifdef __CUDA_ARCH__
  return atomicCAS (a, cmp, v);
#else
  int o = *a; if (o == cmp) *a = v; return o;
#endif
}
B2_HD int atomic_add_i32 (int* a, int v)
{
#ifdef __CUDA_ARCH__
  return atomicAdd (a, v);
#else
  int o = *a; *a = o + v; return o;
#endif
}

// ---- spatial hash --------------------------------------------------------------------------------
struct PointGrid
{
  uint64_t* keys;        // cell key per slot (GRID_EMPTY = unused)
  int* head;             // first point of the cell's list, -1 = none
  int* next;             // per point: next point of the same cell
  uint32_t mask;         // slots - 1 (power of two)
  double inv_cell;       // 1 / cell edge
  const float* pts;      // xyz triples
};
constexpr uint64_t GRID_EMPTY = ~0ull;
constexpr int GRID_HALF = 1 << 20;                 // cell coordinates are clamped to 21 bits per axis

B2_HD int grid_cell1 (const PointGrid& g, float x)
{
  double c = floor ((double) x * g.inv_cell);
  if (!(c > -(double) GRID_HALF)) c = -(double) GRID_HALF + 1;       // also catches NaN
  if (c > (double) GRID_HALF - 2) c = (double) GRID_HALF - 2;
  return (int) c;
}
B2_HD uint64_t grid_key (int cx, int cy, int cz)
{ return ((uint64_t) (cx + GRID_HALF) << 42) | ((uint64_t) (cy + GRID_HALF) << 21) | (uint64_t) (cz + GRID_HALF); }
B2_HD uint32_t grid_hash (uint64_t k)
{ k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return (uint32_t) k; }

B2_HD int grid_find (const PointGrid& g, uint64_t key)
{
  for (uint32_t s = grid_hash (key) & g.mask;; s = (s + 1) & g.mask)
  {
    uint64_t k = g.keys[s];
    if (k == key) return (int) s;
    if (k == GRID_EMPTY) return -1;
  }
}
// the table has at least 2x as many slots as points, so insertion always terminates
B2_HD void grid_insert (const PointGrid& g, int i)
{
  const float* q = g.pts + 3 * (size_t) i;
  uint64_t key = grid_key (grid_cell1 (g, q[0]), grid_cell1 (g, q[1]), grid_cell1 (g, q[2]));
  for (uint32_t s = grid_hash (key) & g.mask;; s = (s + 1) & g.mask)
  {
    uint64_t k = g.keys[s];
    if (k == GRID_EMPTY)
    {
      uint64_t prev = atomic_cas64 (&g.keys[s], GRID_EMPTY, key);
      k = prev == GRID_EMPTY ? key : prev;            // claimed, or somebody else's key
    }
    if (k == key) { g.next[i] = atomic_exch_i32 (&g.head[s], i); return; }
  }
}

// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz, accumulated in float in x,y,z order
B2_HD float sqdist3 (const float* a, const float* b)
{
  float dx = fsub (a[0], b[0]), dy = fsub (a[1], b[1]), dz = fsub (a[2], b[2]);
  return fadd (fadd (fmul (dx, dx), fmul (dy, dy)), fmul (dz, dz));
}

// calls f (j) for every point j != i with squared distance < r2 (RadiusResultSet keeps dist < radius^2);
// f returns false to stop early
template <typename F> B2_HD void grid_neighbors (const PointGrid& g, int i, float r2, F& f)
{
  const float* q = g.pts + 3 * (size_t) i;
  int cx = grid_cell1 (g, q[0]), cy = grid_cell1 (g, q[1]), cz = grid_cell1 (g, q[2]);
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = -1; dz <= 1; ++dz)
      {
        int s = grid_find (g, grid_key (cx + dx, cy + dy, cz + dz));
        if (s < 0) continue;
        for (int j = g.head[s]; j >= 0; j = g.next[j])
        {
          if (j == i) continue;
          if (sqdist3 (q, g.pts + 3 * (size_t) j) < r2) { if (!f (j)) return; }
        }
      }
}

// ---- flattenVertices -------------------------------------------------------------------------------
enum { FV_UNDECIDED = 0, FV_KEPT = 1, FV_MERGED = 2 };

struct FvRound { const unsigned char* state; int i; bool kept_before, undecided_before;
  B2_HD bool operator() (int j)
  {
    if (j > i) return true;
    unsigned char s = state[j];
    if (s == FV_KEPT) { kept_before = true; return false; }
    if (s == FV_UNDECIDED) undecided_before = true;
    return true;
  } };
// one fix-point step for vertex i; returns true when i is still undecided
B2_HD bool fv_round (const PointGrid& g, float r2, unsigned char* state, int i)
{
  if (state[i] != FV_UNDECIDED) return false;
  FvRound f{ state, i, false, false };
  grid_neighbors (g, i, r2, f);
  if (f.kept_before) { state[i] = FV_MERGED; return false; }
  if (f.undecided_before) return true;
  state[i] = FV_KEPT;
  return false;
}
struct FvLast { const unsigned char* state; int best;
  B2_HD bool operator() (int j) { if (state[j] == FV_KEPT && j > best) best = j; return true; } };
// integrate.cpp:119-126: the vertex a merged vertex ends up on = the last kept vertex that reaches it
B2_HD int fv_target (const PointGrid& g, float r2, const unsigned char* state, int i)
{
  if (state[i] == FV_KEPT) return i;
  FvLast f{ state, -1 };
  grid_neighbors (g, i, r2, f);
  return f.best;
}

// ---- cleanupMesh -----------------------------------------------------------------------------------
constexpr int CM_MAX_K = 16;        // largest supported min_neighbors

// meshToFaceCloud (integrate.cpp:70-101): centroid (v0 + v1 + v2) / 3.f in float
B2_HD void face_centroid (const float* verts, const int* tri, float* c)
{
  const float *a = verts + 3 * (size_t) tri[0], *b = verts + 3 * (size_t) tri[1], *d = verts + 3 * (size_t) tri[2];
  for (int k = 0; k < 3; ++k) c[k] = fdiv (fadd (fadd (a[k], b[k]), d[k]), 3.f);
}

struct CmCount { int K; int n; int* nb;
  B2_HD bool operator() (int j) { if (n < K - 1) nb[n] = j; return ++n < K; } };
// counts the faces within range of face i, stopping at K; records up to K-1 of them in nb[i*(CM_MAX_K-1)..]
B2_HD int cm_count (const PointGrid& g, float r2, int K, int i, int* nb_all)
{
  CmCount f{ K, 0, nb_all + (size_t) i * (CM_MAX_K - 1) };
  grid_neighbors (g, i, r2, f);
  return f.n;
}
B2_HD int uf_find (int* parent, int a)
{
  for (;;)
  {
    int p = parent[a];
    if (p == a) return a;
    int gp = parent[p];
    if (gp != p) parent[a] = gp;          // path halving (benign race: only ever points further up)
    a = p;
  }
}
B2_HD void uf_union (int* parent, int a, int b)
{
  for (;;)
  {
    a = uf_find (parent, a); b = uf_find (parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // the larger root is hooked under the smaller
    if (atomic_cas_i32 (&parent[a], a, b) == a) return;
  }
}

} // namespace b2

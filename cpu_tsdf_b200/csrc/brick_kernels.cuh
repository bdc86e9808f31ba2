// brick_kernels.cuh — the fast integrate path: level-synchronous work queues above the finest
// tier and one warp per 8^3 voxel block at the finest tier.
//
// updateVoxel (impl/tsdf_volume_octree.hpp:113-218) is a depth-first recursion whose per-node
// decisions depend only on (a) the node's own stored state, (b) the frame, and (c) the return
// codes of its children.  That makes it separable into
//   1. a top-down sweep over levels C .. B-1 (B = L-3, the block-root level): one thread per
//      visited node decides "already split / split now by proximity / update as a leaf" and
//      appends its children to the next level's queue (k_upper_down);
//   2. one warp per visited block root: the whole 1+8+64+512-node subtree is staged in shared
//      memory, swept top-down and bottom-up with warp ballots, and written back (k_blocks);
//   3. a bottom-up sweep over levels B-1 .. C that folds the children's return codes, prunes
//      all-empty children and applies the fall-through leaf update (k_upper_up).
// Speculative splits are safe because storage of non-existent nodes is always fresh (see
// tsdf_core.cuh).  The one case that is not separable — a node whose pre-existing children are
// all pruned AND which then re-splits in the same call — is detected and
// handed to the general depth-first routine update_voxel_dfs for that subtree only.
#pragma once
#include "tsdf_core.cuh"

namespace b2 {

struct QNode
{
  int x, y, z;        // integer coordinates at the queue's level
  int slot, idx;      // storage of the node's own state (slot < 0: root arrays)
  int kind;           // 0 = finished (rc valid), 1 = interior with pre-existing children, 2 = interior split this frame
  int child_base;     // index of the first of its 8 children in the next level's queue
  int rc;             // updateVoxel return code
};

constexpr int MAX_QLEVELS = 8;
struct Queues
{
  QNode* q[MAX_QLEVELS];    // q[i] holds level C+i
  int cap[MAX_QLEVELS];
  int* n;                   // n[i] = entries in q[i]   (device memory)
};

enum { KIND_DONE = 0, KIND_OLD = 1, KIND_NEW = 2 };

__device__ __forceinline__ NodePos qnode_pos (const Params& p, int level, const QNode& e)
{
  NodePos n;
  n.level = level; n.x = e.x; n.y = e.y; n.z = e.z;
  n.cx = center1d (p, level, e.x); n.cy = center1d (p, level, e.y); n.cz = center1d (p, level, e.z);
  n.size = level_size (p, level);
  n.slot = e.slot; n.idx = e.idx;
  return n;
}

__device__ __forceinline__ bool push_children (const Params& p, const Queues& Q, int li, const NodePos& n, int cs, QNode& e)
{
  int base = atomicAdd (&Q.n[li + 1], 8);
  if (base + 8 > Q.cap[li + 1]) { raise_err (p, ERR_QUEUE_FULL); e.kind = KIND_DONE; e.rc = 0; return false; }
  e.child_base = base;
  for (int c = 0; c < 8; ++c)
  {
    NodePos ch = make_child (p, n, c, cs);
    QNode q; q.x = ch.x; q.y = ch.y; q.z = ch.z; q.slot = ch.slot; q.idx = ch.idx; q.kind = KIND_DONE; q.child_base = -1; q.rc = 0;
    Q.q[li + 1][base + c] = q;
  }
  return true;
}

// ---- 1. top-down over the upper levels: one thread per queued node ---------------------------------
__global__ void k_upper_down (Params p, Frame f, Queues Q, int li, unsigned long long* __restrict__ stats)
{
  int level = p.C + li;
  int count = Q.n[li];
  if (count > Q.cap[li]) count = Q.cap[li];
  unsigned long long upd = 0, vis = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
  {
    QNode e = Q.q[li][i];
    NodePos n = qnode_pos (p, level, e);
    vis++;
    uint32_t m; uint32_t* sw = split_word (p, n, m);
    if (*sw & m)                                                    // hpp:122
    {
      int cs = children_slot (p, n, false);
      if (cs < 0) { raise_err (p, ERR_MISSING_BRICK); e.kind = KIND_DONE; e.rc = 0; }
      else { e.kind = KIND_OLD; push_children (p, Q, li, n, cs, e); }
    }
    else
    {
      Obs o = observe (p, f, n.cx, n.cy, n.cz, n.size);
      if (!o.valid) { e.kind = KIND_DONE; e.rc = 0; }
      else if (o.near_ && n.size > p.finest_size)                    // hpp:161-166
      {
        int cs = children_slot (p, n, true);
        if (cs < 0) { e.kind = KIND_DONE; e.rc = 0; }
        else
        {
          e.kind = KIND_NEW;
          if (push_children (p, Q, li, n, cs, e)) atomicOr (sw, m);   // split (): children are fresh by invariant
        }
      }
      else
      {
        bool updated;
        e.rc = leaf_update (p, f, n, o, updated);
        e.kind = KIND_DONE;
        upd += updated;
      }
    }
    Q.q[li][i] = e;
  }
  if (upd) atomicAdd (&stats[0], upd);
  if (vis) atomicAdd (&stats[1], vis);
}

// ---- 3. bottom-up over the upper levels -------------------------------------------------------------
__global__ void k_upper_up (Params p, Frame f, Queues Q, int li, unsigned long long* __restrict__ stats)
{
  int level = p.C + li;
  int count = Q.n[li];
  if (count > Q.cap[li]) count = Q.cap[li];
  unsigned long long upd = 0, vis = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
  {
    QNode e = Q.q[li][i];
    if (e.kind == KIND_DONE) continue;
    const QNode* ch = Q.q[li + 1] + e.child_base;
    bool all_empty = true;
    for (int c = 0; c < 8; ++c) all_empty &= (ch[c].rc < 0);
    if (!all_empty) { e.rc = 1; }                                   // hpp:140 / :185
    else
    {
      NodePos n = qnode_pos (p, level, e);
      uint32_t m; uint32_t* sw = split_word (p, n, m);
      atomicAnd (sw, ~m);                                           // children.clear ()
      int cs = children_slot (p, n, false);
      if (cs >= 0) for (int c = 0; c < 8; ++c) reset_node (p, make_child (p, n, c, cs));
      if (e.kind == KIND_NEW)
      {
        Obs o = observe (p, f, n.cx, n.cy, n.cz, n.size);           // same observation as in the down sweep
        bool updated;
        e.rc = leaf_update (p, f, n, o, updated);                   // hpp:189-214 (falls through after :181)
        upd += updated;
      }
      else
      {
        // pre-existing children pruned: the node is visited as a leaf and may re-split (hpp:136 -> :161)
        Counters cnt; cnt.n_updates = 0; cnt.n_visits = 0;
        e.rc = update_voxel_dfs (p, f, n, cnt);
        upd += cnt.n_updates; vis += cnt.n_visits - 1;
      }
    }
    e.kind = KIND_DONE;
    Q.q[li][i].rc = e.rc;
  }
  if (upd) atomicAdd (&stats[0], upd);
  if (vis) atomicAdd (&stats[1], vis);
}

// ---- 2. one warp per block root -----------------------------------------------------------------------
constexpr int BLK_WARPS = 4;

struct WarpSmem
{
  float2 dw[BRICK_NODES];       // 4672 B
  uchar4 rgb[BRICK_NODES];      // 2336 B (colour volumes only)
};

__device__ __forceinline__ void child_center (float pc, float off, int bit, float& c)
{ c = bit ? fadd (pc, off) : fsub (pc, off); }

// process the node held at smem index `si`, geometry (cx,cy,cz,size); returns rc, sets flags
struct NodeResult { int kind; int rc; bool updated; Obs o; };

__device__ __forceinline__ NodeResult visit_node (const Params& p, const Frame& f, WarpSmem& S, int si, bool split_old,
                                                  float cx, float cy, float cz, float size, bool can_split)
{
  NodeResult r; r.kind = KIND_DONE; r.rc = 0; r.updated = false;
  if (split_old) { r.kind = KIND_OLD; return r; }
  r.o = observe (p, f, cx, cy, cz, size);
  if (!r.o.valid) return r;
  if (can_split && r.o.near_) { r.kind = KIND_NEW; return r; }
  float M = 0.f; int ns = 0;
  r.rc = leaf_update_values (p, f, r.o, S.dw[si], S.rgb[si], M, ns, r.updated);
  return r;
}

template <bool COLOR>
__global__ void __launch_bounds__ (BLK_WARPS * 32) k_blocks (Params p, Frame f, Queues Q, int li, unsigned long long* __restrict__ stats)
{
  __shared__ WarpSmem smem[BLK_WARPS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpSmem& S = smem[wib];
  const int B = p.C + li;                         // block-root level = L - 3
  int count = Q.n[li];
  if (count > Q.cap[li]) count = Q.cap[li];
  const int nwarps = gridDim.x * BLK_WARPS;
  unsigned long long upd = 0, vis = 0, blocks = 0;
  const float sizeB = level_size (p, B);
  const float off1 = sizeB * 0.25f, off2 = sizeB * 0.125f, off3 = sizeB * 0.0625f;
  const float size1 = sizeB * 0.5f, size2 = sizeB * 0.25f, size3 = sizeB * 0.125f;

  for (int wi = blockIdx.x * BLK_WARPS + wib; wi < count; wi += nwarps)
  {
    QNode e = Q.q[li][wi];
    NodePos nb = qnode_pos (p, B, e);
    blocks += (lane == 0);
    vis += (lane == 0);
    uint32_t rm; uint32_t* rsw = split_word (p, nb, rm);
    const bool root_old = (*rsw & rm) != 0;
    int bslot = find_brick (p, 0, nb.x, nb.y, nb.z);
    int kindR;
    Obs oR; oR.valid = false;
    if (root_old)
    {
      kindR = KIND_OLD;
      if (bslot < 0) { if (lane == 0) { raise_err (p, ERR_MISSING_BRICK); Q.q[li][wi].rc = 0; } continue; }
    }
    else
    {
      oR = observe (p, f, nb.cx, nb.cy, nb.cz, nb.size);
      if (!oR.valid) { if (lane == 0) Q.q[li][wi].rc = 0; continue; }
      if (oR.near_)                                 // size > finest always holds at level L-3
      {
        kindR = KIND_NEW;
        if (bslot < 0)
        {
          if (lane == 0) bslot = find_or_insert_brick (p, 0, nb.x, nb.y, nb.z);
          bslot = __shfl_sync (0xffffffffu, bslot, 0);
          if (bslot < 0) { if (lane == 0) Q.q[li][wi].rc = 0; continue; }
        }
      }
      else
      {
        if (lane == 0)
        {
          bool updated;
          Q.q[li][wi].rc = leaf_update (p, f, nb, oR, updated);
          upd += updated;
        }
        continue;
      }
    }
    // ---- stage the brick: coalesced 8-byte loads, node j3 = lane + 32 i lives at smem index 72 + j3 ----
    float2* gdw = p.nodes + (size_t) bslot * BRICK_NODES;
    uchar4* grgb = COLOR ? p.rgb + (size_t) bslot * BRICK_NODES : nullptr;
    uint32_t* gsw = p.split + (size_t) bslot * BRICK_SPLIT_WORDS;
    __syncwarp ();
#pragma unroll 4
    for (int j = lane; j < BRICK_NODES; j += 32)
    {
      S.dw[j] = gdw[j];
      if (COLOR) S.rgb[j] = grgb[j];
    }
    const uint32_t s1_old = gsw[0] & 0xFFu;
    const uint32_t s2_old0 = gsw[1], s2_old1 = gsw[2];
    __syncwarp ();

    bool bail = false;
    unsigned long long bupd = 0, bvis = 0;            // this block's counts, added on commit only
    uint32_t dirty = 0;                             // bits 0-15 finest i, 16-17 level 2, 18 level 1
    // ---- level 1 (8 nodes, lanes 0..7) ----
    NodeResult r1; r1.kind = KIND_DONE; r1.rc = 0; r1.updated = false; r1.o.valid = false;
    float c1x = 0, c1y = 0, c1z = 0;
    if (lane < 8)
    {
      child_center (nb.cx, off1, (lane >> 2) & 1, c1x); child_center (nb.cy, off1, (lane >> 1) & 1, c1y); child_center (nb.cz, off1, lane & 1, c1z);
      r1 = visit_node (p, f, S, lane, (s1_old >> lane) & 1, c1x, c1y, c1z, size1, true);
      if (r1.updated) dirty |= 1u << 18;
      bupd += r1.updated;
    }
    const uint32_t int1 = __ballot_sync (0xffffffffu, lane < 8 && r1.kind != KIND_DONE);   // interior level-1 nodes
    const uint32_t new1 = __ballot_sync (0xffffffffu, lane < 8 && r1.kind == KIND_NEW);
    bvis += (lane == 0) ? 8 : 0;
    // ---- level 2 (64 nodes: j2 = lane + 32 i2) ----
    NodeResult r2[2];
    uint32_t int2[2], new2[2];
#pragma unroll
    for (int i2 = 0; i2 < 2; ++i2)
    {
      const int j2 = lane + 32 * i2;
      const int j1 = j2 >> 3, cc = j2 & 7;
      r2[i2].kind = KIND_DONE; r2[i2].rc = 0; r2[i2].updated = false; r2[i2].o.valid = false;
      const bool visited = (int1 >> j1) & 1;
      if (visited)
      {
        float px, py, pz, cx, cy, cz;
        child_center (nb.cx, off1, (j1 >> 2) & 1, px); child_center (nb.cy, off1, (j1 >> 1) & 1, py); child_center (nb.cz, off1, j1 & 1, pz);
        child_center (px, off2, (cc >> 2) & 1, cx); child_center (py, off2, (cc >> 1) & 1, cy); child_center (pz, off2, cc & 1, cz);
        const bool sold = ((i2 ? s2_old1 : s2_old0) >> lane) & 1;
        r2[i2] = visit_node (p, f, S, 8 + j2, sold, cx, cy, cz, size2, true);
        if (r2[i2].updated) dirty |= 1u << (16 + i2);
        bupd += r2[i2].updated;
      }
      int2[i2] = __ballot_sync (0xffffffffu, visited && r2[i2].kind != KIND_DONE);
      new2[i2] = __ballot_sync (0xffffffffu, visited && r2[i2].kind == KIND_NEW);
    }
    bvis += (lane == 0) ? 8 * __popc (int1) : 0;
    bvis += (lane == 0) ? 8 * (__popc (int2[0]) + __popc (int2[1])) : 0;
    // ---- level 3 (512 finest voxels: j3 = lane + 32 i) ----
    uint32_t nonneg_mine = 0;                       // lane k (<16) keeps the ballot of iteration k
#pragma unroll 2
    for (int i = 0; i < 16; ++i)
    {
      const int j3 = lane + 32 * i;
      const int j2 = j3 >> 3;
      const bool visited = ((j2 < 32 ? int2[0] : int2[1]) >> (j2 & 31)) & 1;
      int rc = 0;
      if (visited)
      {
        const int j1 = j3 >> 6, c2 = (j3 >> 3) & 7, c3 = j3 & 7;
        float ax, ay, az, bx, by, bz, cx, cy, cz;
        child_center (nb.cx, off1, (j1 >> 2) & 1, ax); child_center (nb.cy, off1, (j1 >> 1) & 1, ay); child_center (nb.cz, off1, j1 & 1, az);
        child_center (ax, off2, (c2 >> 2) & 1, bx); child_center (ay, off2, (c2 >> 1) & 1, by); child_center (az, off2, c2 & 1, bz);
        child_center (bx, off3, (c3 >> 2) & 1, cx); child_center (by, off3, (c3 >> 1) & 1, cy); child_center (bz, off3, c3 & 1, cz);
        NodeResult r3 = visit_node (p, f, S, 72 + j3, false, cx, cy, cz, size3, false);
        rc = r3.rc;
        if (r3.updated) dirty |= 1u << i;
        bupd += r3.updated;
      }
      const uint32_t nn = __ballot_sync (0xffffffffu, visited && rc >= 0);
      if (lane == i) nonneg_mine = nn;
    }
    // ---- bottom-up: level 2 ----
    uint32_t pruned2[2], nonneg2[2];
#pragma unroll
    for (int i2 = 0; i2 < 2; ++i2)
    {
      const int j2 = lane + 32 * i2;
      const uint32_t nn = __shfl_sync (0xffffffffu, nonneg_mine, (lane >> 2) + 8 * i2);   // ballot of iteration j2 >> 2
      bool pruned = false;
      if (r2[i2].kind != KIND_DONE)
      {
        const bool all_empty = ((nn >> (8 * (lane & 3))) & 0xFFu) == 0;
        if (!all_empty) r2[i2].rc = 1;
        else
        {
          pruned = true;
          Obs o = r2[i2].o;
          bool ok = true;
          if (r2[i2].kind == KIND_OLD)
          {
            const int j1 = j2 >> 3, cc = j2 & 7;
            float px, py, pz, cx, cy, cz;
            child_center (nb.cx, off1, (j1 >> 2) & 1, px); child_center (nb.cy, off1, (j1 >> 1) & 1, py); child_center (nb.cz, off1, j1 & 1, pz);
            child_center (px, off2, (cc >> 2) & 1, cx); child_center (py, off2, (cc >> 1) & 1, cy); child_center (pz, off2, cc & 1, cz);
            o = observe (p, f, cx, cy, cz, size2);
            if (!o.valid) { r2[i2].rc = 0; ok = false; }
            else if (o.near_) { bail = true; ok = false; }         // prune-then-resplit: general path
          }
          if (ok)
          {
            float M = 0.f; int ns = 0; bool updated;
            r2[i2].rc = leaf_update_values (p, f, o, S.dw[8 + j2], S.rgb[8 + j2], M, ns, updated);
            if (updated) dirty |= 1u << (16 + i2);
            bupd += updated;
          }
        }
      }
      pruned2[i2] = __ballot_sync (0xffffffffu, pruned);
      const bool visited2 = (int1 >> (j2 >> 3)) & 1;
      nonneg2[i2] = __ballot_sync (0xffffffffu, visited2 && r2[i2].rc >= 0);
    }
    // children of pruned level-2 nodes go back to the fresh state
    if (pruned2[0] | pruned2[1])
    {
#pragma unroll 4
      for (int i = 0; i < 16; ++i)
      {
        const int j2 = (lane + 32 * i) >> 3;
        if (((j2 < 32 ? pruned2[0] : pruned2[1]) >> (j2 & 31)) & 1)
        {
          S.dw[72 + lane + 32 * i] = make_float2 (-1.f, 0.f);
          if (COLOR) S.rgb[72 + lane + 32 * i] = make_uchar4 (0, 0, 0, 0);
          dirty |= 1u << i;
        }
      }
    }
    // ---- bottom-up: level 1 ----
    bool pruned1f = false;
    if (lane < 8 && r1.kind != KIND_DONE)
    {
      const uint32_t nn = nonneg2[lane >> 2];
      const bool all_empty = ((nn >> (8 * (lane & 3))) & 0xFFu) == 0;
      if (!all_empty) r1.rc = 1;
      else
      {
        pruned1f = true;
        Obs o = r1.o;
        bool ok = true;
        if (r1.kind == KIND_OLD)
        {
          o = observe (p, f, c1x, c1y, c1z, size1);
          if (!o.valid) { r1.rc = 0; ok = false; }
          else if (o.near_) { bail = true; ok = false; }
        }
        if (ok)
        {
          float M = 0.f; int ns = 0; bool updated;
          r1.rc = leaf_update_values (p, f, o, S.dw[lane], S.rgb[lane], M, ns, updated);
          if (updated) dirty |= 1u << 18;
          bupd += updated;
        }
      }
    }
    const uint32_t pruned1 = __ballot_sync (0xffffffffu, pruned1f);
    const uint32_t nonneg1 = __ballot_sync (0xffffffffu, lane < 8 && r1.rc >= 0);
    if (pruned1)
    {
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2)
      {
        const int j2 = lane + 32 * i2;
        if ((pruned1 >> (j2 >> 3)) & 1)
        {
          S.dw[8 + j2] = make_float2 (-1.f, 0.f);
          if (COLOR) S.rgb[8 + j2] = make_uchar4 (0, 0, 0, 0);
          dirty |= 1u << (16 + i2);
        }
      }
    }
    // ---- root ----
    int rcR = 1;
    bool prunedR = false, root_updated = false;
    float2 rdw = make_float2 (0.f, 0.f); uchar4 rrgb = make_uchar4 (0, 0, 0, 0);
    if ((nonneg1 & 0xFFu) == 0)
    {
      prunedR = true;
      Obs o = oR;
      bool ok = true;
      if (kindR == KIND_OLD)
      {
        o = observe (p, f, nb.cx, nb.cy, nb.cz, nb.size);
        if (!o.valid) { rcR = 0; ok = false; }
        else if (o.near_) { bail = true; ok = false; }
      }
      if (ok)
      {
        rdw = *node_dw (p, nb);
        if (COLOR) rrgb = nb.slot < 0 ? p.root_rgb[nb.idx] : p.rgb[(size_t) nb.slot * BRICK_NODES + nb.idx];
        float M = 0.f; int ns = 0;
        rcR = leaf_update_values (p, f, o, rdw, rrgb, M, ns, root_updated);
      }
    }
    bail = __any_sync (0xffffffffu, bail);
    if (bail)
    {
      // nothing has been committed: redo this block root with the general depth-first routine
      if (lane == 0)
      {
        Counters cnt; cnt.n_updates = 0; cnt.n_visits = 0;
        Q.q[li][wi].rc = update_voxel_dfs (p, f, nb, cnt);
        upd += cnt.n_updates;
        vis += cnt.n_visits - 1;          // the root visit was counted above
      }
      __syncwarp ();
      continue;
    }
    upd += bupd; vis += bvis;
    // ---- commit ----
    __syncwarp ();
#pragma unroll 4
    for (int i = 0; i < 16; ++i)
      if ((dirty >> i) & 1)
      {
        gdw[72 + lane + 32 * i] = S.dw[72 + lane + 32 * i];
        if (COLOR) grgb[72 + lane + 32 * i] = S.rgb[72 + lane + 32 * i];
      }
#pragma unroll
    for (int i2 = 0; i2 < 2; ++i2)
      if ((dirty >> (16 + i2)) & 1)
      {
        gdw[8 + lane + 32 * i2] = S.dw[8 + lane + 32 * i2];
        if (COLOR) grgb[8 + lane + 32 * i2] = S.rgb[8 + lane + 32 * i2];
      }
    if ((dirty >> 18) & 1) { gdw[lane] = S.dw[lane]; if (COLOR) grgb[lane] = S.rgb[lane]; }
    // level-1 nodes reset by a root prune
    if (prunedR && lane < 8) { gdw[lane] = make_float2 (-1.f, 0.f); if (COLOR) grgb[lane] = make_uchar4 (0, 0, 0, 0); }
    if (lane == 0)
    {
      const uint32_t s1_new = ((s1_old | (new1 & 0xFFu)) & ~(pruned1 & 0xFFu)) & 0xFFu;
      const uint32_t s2_new0 = (s2_old0 | new2[0]) & ~pruned2[0];
      const uint32_t s2_new1 = (s2_old1 | new2[1]) & ~pruned2[1];
      if (s1_new != s1_old) gsw[0] = s1_new;         // (all zero when the root is pruned: every child is a leaf then)
      if (s2_new0 != s2_old0) gsw[1] = s2_new0;
      if (s2_new1 != s2_old1) gsw[2] = s2_new1;
      if (kindR == KIND_NEW && !prunedR) atomicOr (rsw, rm);
      if (kindR == KIND_OLD && prunedR) atomicAnd (rsw, ~rm);
      if (root_updated)
      {
        *node_dw (p, nb) = rdw;
        if (COLOR) { if (nb.slot < 0) p.root_rgb[nb.idx] = rrgb; else p.rgb[(size_t) nb.slot * BRICK_NODES + nb.idx] = rrgb; }
        upd += 1;
      }
      Q.q[li][wi].rc = rcR;
    }
    __syncwarp ();
  }
  // warp-reduce the counters, one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    upd += __shfl_down_sync (0xffffffffu, upd, o);
    vis += __shfl_down_sync (0xffffffffu, vis, o);
    blocks += __shfl_down_sync (0xffffffffu, blocks, o);
  }
  if (lane == 0)
  {
    if (upd) atomicAdd (&stats[0], upd);
    if (vis) atomicAdd (&stats[1], vis);
    if (blocks) atomicAdd (&stats[2], blocks);
  }
}

} // namespace b2

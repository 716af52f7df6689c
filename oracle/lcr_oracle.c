/*
 * lcr_oracle.c -- CPU restatement (plain C) of the two native operators on the LCR-Net
 * inference hot path.  TEST INFRASTRUCTURE ONLY: nothing under lcr-net_b200/ may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs use it, and only as the checker / reported baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_ops.py compares every function here with
 * the unmodified reference sources compiled into oracle/_ref/libref_ext.so (see
 * oracle/Makefile) and with the committed fixtures in tests/golden/ generated from them.
 *
 * Reference followed (paths relative to the upstream repository root):
 *   grid subsampling : utils/extensions/cpu/grid_subsampling/grid_subsampling_cpu.cpp:3-75
 *                      utils/extensions/cpu/grid_subsampling/grid_subsampling_cpu.h:7-21
 *                      utils/extensions/extra/cloud/cloud.cpp:4-37, cloud.h:84-99
 *   radius neighbours: utils/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp:3-91
 *                      utils/extensions/extra/nanoflann/nanoflann.hpp:208-256 (result set,
 *                      strict d2 < r2), :423-442 (L2_Simple metric, left-to-right fp32 sum)
 *   tie-break        : cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:125-208
 *                      (batch_ordered_neighbors: equal distances keep ascending support index)
 *
 * Build with -ffp-contract=off: the reference is an x86-64 baseline build, i.e. no FMA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* libstdc++ (GCC 13) _Prime_rehash_policy growth sequence for an unordered_map that only
 * ever grows by single inserts with max_load_factor 1: next = first entry of __prime_list
 * that is >= 2*current (first step: 13, from the __fast_bkt table).  Printed by a 10-line
 * C++ program, see DESIGN.md "subsample order". */
static const uint64_t k_bucket_seq[] = {
    1ull, 13ull, 29ull, 59ull, 127ull, 257ull, 541ull, 1109ull, 2357ull, 5087ull, 10273ull,
    20753ull, 42043ull, 85229ull, 172933ull, 351061ull, 712697ull, 1447153ull, 2938679ull,
    5967347ull, 12117689ull, 24607243ull, 49969847ull, 101473717ull};
#define K_BUCKET_SEQ_LEN ((int)(sizeof(k_bucket_seq) / sizeof(k_bucket_seq[0])))

typedef struct {
  uint64_t key;
  int32_t next; /* index of next node in the singly linked list, -1 = end */
  int32_t cnt;
  float sx, sy, sz;
} vox_node;

/* "before" pointer stored per bucket: -1 = empty bucket, -2 = &_M_before_begin, >=0 node */
#define BKT_EMPTY (-1)
#define BKT_BBEGIN (-2)

typedef struct {
  vox_node* nodes;
  int32_t n_nodes;
  int32_t* buckets;
  uint64_t n_bkt;
  int seq_pos;
  int32_t head; /* _M_before_begin._M_nxt */
} vox_table;

static int32_t* tbl_next_ptr(vox_table* t, int32_t before) {
  return before == BKT_BBEGIN ? &t->head : &t->nodes[before].next;
}

/* _Hashtable::_M_insert_bucket_begin */
static void tbl_insert_bucket_begin(vox_table* t, uint64_t bkt, int32_t n) {
  if (t->buckets[bkt] != BKT_EMPTY) {
    int32_t* nx = tbl_next_ptr(t, t->buckets[bkt]);
    t->nodes[n].next = *nx;
    *nx = n;
  } else {
    t->nodes[n].next = t->head;
    t->head = n;
    if (t->nodes[n].next >= 0) {
      uint64_t ob = t->nodes[t->nodes[n].next].key % t->n_bkt;
      t->buckets[ob] = n;
    }
    t->buckets[bkt] = BKT_BBEGIN;
  }
}

/* _Hashtable::_M_rehash_aux(n, true_type) */
static int tbl_rehash(vox_table* t, uint64_t n_new) {
  int32_t* nb = (int32_t*)malloc(sizeof(int32_t) * n_new);
  if (!nb) return -1;
  for (uint64_t i = 0; i < n_new; i++) nb[i] = BKT_EMPTY;
  int32_t p = t->head;
  t->head = -1;
  uint64_t bbegin_bkt = 0;
  while (p >= 0) {
    int32_t next = t->nodes[p].next;
    uint64_t bkt = t->nodes[p].key % n_new;
    if (nb[bkt] == BKT_EMPTY) {
      t->nodes[p].next = t->head;
      t->head = p;
      nb[bkt] = BKT_BBEGIN;
      if (t->nodes[p].next >= 0) nb[bbegin_bkt] = p;
      bbegin_bkt = bkt;
    } else {
      int32_t* nx = nb[bkt] == BKT_BBEGIN ? &t->head : &t->nodes[nb[bkt]].next;
      t->nodes[p].next = *nx;
      *nx = p;
    }
    p = next;
  }
  free(t->buckets);
  t->buckets = nb;
  t->n_bkt = n_new;
  return 0;
}

static int32_t tbl_find(const vox_table* t, uint64_t key) {
  uint64_t bkt = key % t->n_bkt;
  int32_t before = t->buckets[bkt];
  if (before == BKT_EMPTY) return -1;
  int32_t p = before == BKT_BBEGIN ? t->head : t->nodes[before].next;
  while (p >= 0) {
    if (t->nodes[p].key == key) return p;
    if (t->nodes[p].key % t->n_bkt != bkt) break;
    p = t->nodes[p].next;
  }
  return -1;
}

/* One cloud.  out must hold 3*n floats; returns the number of voxels, <0 on error.
 * grid_subsampling_cpu.cpp:3-48 (single_grid_subsampling_cpu). */
static int64_t subsample_one(const float* pts, int64_t n, float voxel, float* out) {
  if (n <= 0) return 0;
  float mnx = pts[0], mny = pts[1], mnz = pts[2], mxx = mnx, mxy = mny, mxz = mnz;
  for (int64_t i = 0; i < n; i++) {
    const float* p = pts + 3 * i;
    if (p[0] < mnx) mnx = p[0];
    if (p[1] < mny) mny = p[1];
    if (p[2] < mnz) mnz = p[2];
    if (p[0] > mxx) mxx = p[0];
    if (p[1] > mxy) mxy = p[1];
    if (p[2] > mxz) mxz = p[2];
  }
  /* PointXYZ * (1. / voxel_size): the double reciprocal narrows to the float parameter */
  const float inv = (float)(1.0 / (double)voxel);
  const float ox = floorf(mnx * inv) * voxel;
  const float oy = floorf(mny * inv) * voxel;
  const float oz = floorf(mnz * inv) * voxel;
  const uint64_t nx = (uint64_t)(floorf((mxx - ox) / voxel) + 1);
  const uint64_t ny = (uint64_t)(floorf((mxy - oy) / voxel) + 1);

  vox_table t;
  t.nodes = (vox_node*)malloc(sizeof(vox_node) * (size_t)n);
  t.buckets = (int32_t*)malloc(sizeof(int32_t));
  if (!t.nodes || !t.buckets) return -1;
  t.buckets[0] = BKT_EMPTY;
  t.n_bkt = 1;
  t.seq_pos = 0;
  t.n_nodes = 0;
  t.head = -1;

  for (int64_t i = 0; i < n; i++) {
    const float* p = pts + 3 * i;
    const uint64_t ix = (uint64_t)floorf((p[0] - ox) / voxel);
    const uint64_t iy = (uint64_t)floorf((p[1] - oy) / voxel);
    const uint64_t iz = (uint64_t)floorf((p[2] - oz) / voxel);
    const uint64_t key = ix + nx * iy + nx * ny * iz;
    int32_t node = tbl_find(&t, key);
    if (node < 0) {
      /* _M_insert_unique_node: rehash check happens before linking the new node */
      if ((uint64_t)t.n_nodes + 1 > t.n_bkt) {
        if (t.seq_pos + 1 >= K_BUCKET_SEQ_LEN) return -2;
        t.seq_pos++;
        if (tbl_rehash(&t, k_bucket_seq[t.seq_pos]) != 0) return -1;
      }
      node = t.n_nodes++;
      t.nodes[node].key = key;
      t.nodes[node].cnt = 0;
      t.nodes[node].sx = t.nodes[node].sy = t.nodes[node].sz = 0.0f;
      tbl_insert_bucket_begin(&t, key % t.n_bkt, node);
    }
    /* SampledData::update: sequential fp32 accumulation in input order */
    t.nodes[node].cnt += 1;
    t.nodes[node].sx += p[0];
    t.nodes[node].sy += p[1];
    t.nodes[node].sz += p[2];
  }
  int64_t m = 0;
  for (int32_t p = t.head; p >= 0; p = t.nodes[p].next) {
    /* v.second.point * (1.0 / v.second.count): reciprocal in double, narrowed to float */
    const float r = (float)(1.0 / (double)t.nodes[p].cnt);
    out[3 * m + 0] = t.nodes[p].sx * r;
    out[3 * m + 1] = t.nodes[p].sy * r;
    out[3 * m + 2] = t.nodes[p].sz * r;
    m++;
  }
  free(t.nodes);
  free(t.buckets);
  return m;
}

/* grid_subsampling_cpu.cpp:50-75.  out_pts must hold 3*sum(lengths) floats.
 * Returns total number of output points (or <0). */
int64_t lcr_oracle_grid_subsample(const float* pts, const int64_t* lengths, int batch,
                                  float voxel, float* out_pts, int64_t* out_lengths) {
  int64_t start = 0, total = 0;
  for (int b = 0; b < batch; b++) {
    int64_t m = subsample_one(pts + 3 * start, lengths[b], voxel, out_pts + 3 * total);
    if (m < 0) return m;
    out_lengths[b] = m;
    total += m;
    start += lengths[b];
  }
  return total;
}

typedef struct {
  float d2;
  int64_t idx;
} hit_t;

static int hit_cmp(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a;
  const hit_t* y = (const hit_t*)b;
  if (x->d2 < y->d2) return -1;
  if (x->d2 > y->d2) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* Brute-force restatement of radius_neighbors_cpu (kd-tree in the reference; identical
 * result set, ascending d2; exact-distance ties in ascending support index).
 * Pass 1 (out_idx == NULL): returns max_count over all queries and fills counts (optional).
 * Pass 2: fills out_idx[Nq, width] with the first `width` neighbours, padded with Ns_total. */
int64_t lcr_oracle_radius_neighbors(const float* q, const float* s, const int64_t* q_len,
                                    const int64_t* s_len, int batch, float radius,
                                    int64_t width, int64_t* out_idx, int32_t* counts) {
  const float r2 = radius * radius;
  int64_t nq_total = 0, ns_total = 0, max_s = 0;
  for (int b = 0; b < batch; b++) {
    nq_total += q_len[b];
    ns_total += s_len[b];
    if (s_len[b] > max_s) max_s = s_len[b];
  }
  hit_t* hits = (hit_t*)malloc(sizeof(hit_t) * (size_t)(max_s > 0 ? max_s : 1));
  if (!hits) return -1;
  int64_t max_count = 0, q0 = 0, s0 = 0;
  for (int b = 0; b < batch; b++) {
    for (int64_t i = q0; i < q0 + q_len[b]; i++) {
      const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
      int64_t c = 0;
      for (int64_t j = s0; j < s0 + s_len[b]; j++) {
        float d2 = 0.0f;
        float d = qx - s[3 * j];
        d2 += d * d;
        d = qy - s[3 * j + 1];
        d2 += d * d;
        d = qz - s[3 * j + 2];
        d2 += d * d;
        if (d2 < r2) {
          hits[c].d2 = d2;
          hits[c].idx = j;
          c++;
        }
      }
      if (c > max_count) max_count = c;
      if (counts) counts[i] = (int32_t)c;
      if (out_idx) {
        qsort(hits, (size_t)c, sizeof(hit_t), hit_cmp);
        for (int64_t k = 0; k < width; k++)
          out_idx[i * width + k] = k < c ? hits[k].idx : ns_total;
      }
    }
    q0 += q_len[b];
    s0 += s_len[b];
  }
  free(hits);
  (void)nq_total;
  return max_count;
}

/* fp32 squared distances of the radius test, for tie-class comparison in the tests. */
void lcr_oracle_neighbor_d2(const float* q, const float* s, int64_t nq, int64_t ns_total,
                            const int64_t* idx, int64_t width, float* out_d2) {
  for (int64_t i = 0; i < nq; i++)
    for (int64_t k = 0; k < width; k++) {
      int64_t j = idx[i * width + k];
      if (j >= ns_total || j < 0) {
        out_d2[i * width + k] = INFINITY;
        continue;
      }
      float d2 = 0.0f, d = q[3 * i] - s[3 * j];
      d2 += d * d;
      d = q[3 * i + 1] - s[3 * j + 1];
      d2 += d * d;
      d = q[3 * i + 2] - s[3 * j + 2];
      d2 += d * d;
      out_d2[i * width + k] = d2;
    }
}

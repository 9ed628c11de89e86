/*
 * oracle/scn_rules.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Plain-C restatement of the integer half of sparseconvnet 0.2 ("scn"): active-site
 * numbering and rulebooks.  The real package is a third-party dependency of the reference
 * (requirements.txt:2, imported at models/SparseConvNet.py:5); its source is not under
 * /root/reference and cannot be installed here, and the reference has no test that pins
 * its results, so this oracle is PARITY UNPINNED against upstream: it follows the
 * behavioural spec in SURVEY.md App. B (B.2, B.5, B.6), the call sites in
 * models/SparseConvNet.py:59-71,113-140 and the mode notes in Function_test.py:38-44.
 * It is pinned instead by hand-written literal cases and by dense conv3d equivalence
 * (tests/test_oracle_*.py).
 *
 * Same algorithmic shape as upstream's CPU path: a hash map from site coordinates to row
 * id, filled by scanning rows in order ("first sight of a key gets id = nActive++").
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t cap;      /* power of two */
  uint64_t *keys;   /* EMPTY = all ones */
  int32_t *vals;
} Map;

#define EMPTY 0xFFFFFFFFFFFFFFFFull

static uint64_t mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

static int map_init(Map *m, int64_t n) {
  int64_t cap = 16;
  while (cap < 2 * n + 2) cap <<= 1;
  m->cap = cap;
  m->keys = (uint64_t *)malloc(sizeof(uint64_t) * cap);
  m->vals = (int32_t *)malloc(sizeof(int32_t) * cap);
  if (!m->keys || !m->vals) return -1;
  memset(m->keys, 0xFF, sizeof(uint64_t) * cap);
  return 0;
}
static void map_free(Map *m) { free(m->keys); free(m->vals); }

/* returns slot; *found says whether key was present */
static int64_t map_probe(const Map *m, uint64_t key, int *found) {
  int64_t s = (int64_t)(mix(key) & (uint64_t)(m->cap - 1));
  for (;;) {
    if (m->keys[s] == key) { *found = 1; return s; }
    if (m->keys[s] == EMPTY) { *found = 0; return s; }
    s = (s + 1) & (m->cap - 1);
  }
}

/* sites are (x,y,z,b) with 0 <= x,y,z < 65536 and 0 <= b < 32768 */
static uint64_t site_key(int64_t x, int64_t y, int64_t z, int64_t b) {
  return ((uint64_t)b << 48) | ((uint64_t)x << 32) | ((uint64_t)y << 16) | (uint64_t)z;
}

/*
 * InputLayer site numbering (scn.InputLayer at models/SparseConvNet.py:61).
 * coords: P rows of ncols (3 or 4) int64, last column = sample index when ncols == 4.
 * out: pv[P]   = voxel id of each row (first-occurrence order, global over samples)
 *      vox[4*N] = (x,y,z,b) of each voxel in id order (caller allocates 4*P)
 * returns N (number of active sites) or -1 on bad input.
 */
int64_t oracle_input_rules(const int64_t *coords, int64_t P, int ncols, int32_t *pv,
                           int32_t *vox) {
  Map m;
  if (map_init(&m, P)) return -1;
  int64_t n = 0;
  for (int64_t r = 0; r < P; ++r) {
    const int64_t *c = coords + r * ncols;
    int64_t b = ncols == 4 ? c[3] : 0;
    if (c[0] < 0 || c[1] < 0 || c[2] < 0 || b < 0 || c[0] > 65535 || c[1] > 65535 ||
        c[2] > 65535 || b > 32767) { map_free(&m); return -1; }
    uint64_t key = site_key(c[0], c[1], c[2], b);
    int found;
    int64_t s = map_probe(&m, key, &found);
    if (!found) {
      m.keys[s] = key;
      m.vals[s] = (int32_t)n;
      vox[4 * n + 0] = (int32_t)c[0]; vox[4 * n + 1] = (int32_t)c[1];
      vox[4 * n + 2] = (int32_t)c[2]; vox[4 * n + 3] = (int32_t)b;
      ++n;
    }
    pv[r] = m.vals[s];
  }
  map_free(&m);
  return n;
}

/*
 * Submanifold 3x3x3 neighbour table (scn.SubmanifoldConvolution at
 * models/SparseConvNet.py:62,117,119).  nbr[o*27 + k] = id of site (o + d_k) in the same
 * sample or -1, with k = 9(dx+1)+3(dy+1)+(dz+1), dz fastest.  The scn rulebook for offset
 * k is the list of pairs (in = nbr[o][k], out = o) over o with nbr >= 0; the oracle's
 * canonical order inside an offset is ascending `out`.
 */
int oracle_subm_map(const int32_t *vox, int64_t N, int32_t *nbr) {
  Map m;
  if (map_init(&m, N)) return -1;
  for (int64_t i = 0; i < N; ++i) {
    int found;
    int64_t s = map_probe(&m, site_key(vox[4*i], vox[4*i+1], vox[4*i+2], vox[4*i+3]), &found);
    m.keys[s] = site_key(vox[4*i], vox[4*i+1], vox[4*i+2], vox[4*i+3]);
    m.vals[s] = (int32_t)i;
  }
  for (int64_t o = 0; o < N; ++o) {
    int k = 0;
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz, ++k) {
          int64_t x = vox[4*o] + dx, y = vox[4*o+1] + dy, z = vox[4*o+2] + dz;
          int32_t v = -1;
          if (x >= 0 && y >= 0 && z >= 0 && x <= 65535 && y <= 65535 && z <= 65535) {
            int found;
            int64_t s = map_probe(&m, site_key(x, y, z, vox[4*o+3]), &found);
            if (found) v = m.vals[s];
          }
          nbr[o * 27 + k] = v;
        }
  }
  map_free(&m);
  return 0;
}

/*
 * Strided convolution grid, filter size == stride == s (scn.Convolution at
 * models/SparseConvNet.py:137-138 and inside scn.UNet / scn.FullyConvolutionalNet).
 * parent[i] = coarse id of fine site i (site // s), off[i] = ((x%s)*s + (y%s))*s + (z%s).
 * Coarse ids: first touch while scanning fine ids ascending (the canonical order adopted
 * in SURVEY 8c; upstream's within-sample order is a hash-table artefact).
 * voxc gets (x,y,z,b) of the coarse sites (caller allocates 4*Nf).  returns Nc.
 */
int64_t oracle_strided(const int32_t *voxf, int64_t Nf, int s, int32_t *parent, int32_t *off,
                       int32_t *voxc) {
  Map m;
  if (map_init(&m, Nf)) return -1;
  int64_t n = 0;
  for (int64_t i = 0; i < Nf; ++i) {
    int32_t x = voxf[4*i], y = voxf[4*i+1], z = voxf[4*i+2], b = voxf[4*i+3];
    int32_t cx = x / s, cy = y / s, cz = z / s;
    off[i] = ((x - cx * s) * s + (y - cy * s)) * s + (z - cz * s);
    uint64_t key = site_key(cx, cy, cz, b);
    int found;
    int64_t sl = map_probe(&m, key, &found);
    if (!found) {
      m.keys[sl] = key;
      m.vals[sl] = (int32_t)n;
      voxc[4*n] = cx; voxc[4*n+1] = cy; voxc[4*n+2] = cz; voxc[4*n+3] = b;
      ++n;
    }
    parent[i] = m.vals[sl];
  }
  map_free(&m);
  return n;
}

/*
 * point2mask CPU restatement (reference CUDA kernels are the only implementation:
 * ops/point2mask/_ext_src/src/ball_query_gpu.cu:9-45, group_points_gpu.cu:8-28,43-64).
 * Literal loops, including the `k < n - ptnum` scan bound (ball_query_gpu.cu:28) and the
 * -1 sentinel written by the caller (ball_query.cpp:20-22).
 */
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xy,
                       const float *xy, const int32_t *pointnums, int32_t *idx) {
  float r2 = radius * radius;
  for (int bi = 0; bi < b; ++bi) {
    const float *q = new_xy + (int64_t)bi * m * 2, *p = xy + (int64_t)bi * n * 2;
    int32_t *o = idx + (int64_t)bi * m * nsample;
    int ptnum = pointnums[bi];
    for (int j = 0; j < m; ++j) {
      float qx = q[2*j], qy = q[2*j+1];
      for (int k = 0, cnt = 0; k < n - ptnum && cnt < nsample; ++k) {
        float dx = qx - p[2*k], dy = qy - p[2*k+1];
        float d2 = dx * dx + dy * dy;
        if (d2 < r2) { o[j * nsample + cnt] = k; ++cnt; }
      }
    }
  }
}

void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                         const int32_t *idx, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int32_t ii = idx[((int64_t)bi * npoints + j) * nsample + k];
          if (ii >= 0)
            out[(((int64_t)bi * c + l) * npoints + j) * nsample + k] =
                points[((int64_t)bi * c + l) * n + ii];
        }
}

void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample,
                              const float *grad_out, const int32_t *idx, float *grad_points) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int32_t ii = idx[((int64_t)bi * npoints + j) * nsample + k];
          if (ii >= 0)
            grad_points[((int64_t)bi * c + l) * n + ii] +=
                grad_out[(((int64_t)bi * c + l) * npoints + j) * nsample + k];
        }
}

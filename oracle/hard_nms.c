/* oracle/hard_nms.c -- TEST INFRASTRUCTURE (see oracle/__init__.py), never linked into the product.
 *
 * Plain-C restatement of the reference's only native component, cython_nms.pyx:24-74
 * (greedy per-class hard NMS): areas with the "+1" pixel convention (:31), candidates visited in
 * descending score order (:32, `scores.argsort()[::-1]` -- restated as descending score with ties
 * broken by DESCENDING index, i.e. a reversed stable ascending sort), a box suppresses every
 * lower-ranked box whose overlap `inter / (iarea + areas[j] - inter)` is >= thresh (:63-72),
 * survivors are reported in ascending original order (:74).
 * All arithmetic is float32 in the reference's order; build without -ffast-math / FMA contraction.
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { float score; int idx; } item_t;

static int cmp_desc(const void* a, const void* b) {
  const item_t *x = (const item_t*)a, *y = (const item_t*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return y->idx - x->idx;                      /* ties: larger index first */
}

/* dets [n][5] = x1,y1,x2,y2,score; keep [n] receives 1 (kept) / 0 (suppressed). Returns #kept. */
int oracle_hard_nms(const float* dets, int n, float thresh, uint8_t* keep) {
  if (n <= 0) return 0;
  item_t* order = (item_t*)malloc(sizeof(item_t) * (size_t)n);
  float* areas = (float*)malloc(sizeof(float) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const float* d = dets + 5 * i;
    volatile float w = d[2] - d[0]; w = w + 1.0f;
    volatile float h = d[3] - d[1]; h = h + 1.0f;
    areas[i] = w * h;
    order[i].score = d[4]; order[i].idx = i;
    keep[i] = 1;
  }
  qsort(order, (size_t)n, sizeof(item_t), cmp_desc);
  int kept = 0;
  for (int a = 0; a < n; ++a) {
    const int i = order[a].idx;
    if (!keep[i]) continue;
    ++kept;
    const float* di = dets + 5 * i;
    for (int b = a + 1; b < n; ++b) {
      const int j = order[b].idx;
      if (!keep[j]) continue;
      const float* dj = dets + 5 * j;
      const float xx1 = di[0] >= dj[0] ? di[0] : dj[0], yy1 = di[1] >= dj[1] ? di[1] : dj[1];
      const float xx2 = di[2] <= dj[2] ? di[2] : dj[2], yy2 = di[3] <= dj[3] ? di[3] : dj[3];
      volatile float w = xx2 - xx1; w = w + 1.0f;
      volatile float h = yy2 - yy1; h = h + 1.0f;
      const float ww = 0.0f >= w ? 0.0f : w, hh = 0.0f >= h ? 0.0f : h;
      volatile float inter = ww * hh;
      volatile float uni = areas[i] + areas[j]; uni = uni - inter;
      const float ovr = inter / uni;
      if (ovr >= thresh) keep[j] = 0;
    }
  }
  free(order); free(areas);
  return kept;
}

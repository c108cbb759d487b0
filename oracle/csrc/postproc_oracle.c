/* postproc_oracle.c — plain-C restatement of the two scikit-image functions HoVerNet._proc_np_hv calls
 * (tiatoolbox/models/architecture/hovernet.py:544, 614: remove_small_objects(label_img, max_size=...), :616: watershed(dist, markers=..., mask=...)).
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): built by __graft_entry__.build() with gcc into oracle/_build/, loaded with
 * ctypes by oracle/hovernet_postproc_oracle.py.  It exists (i) as a second, independently written implementation that the pure-Python
 * restatement is cross-checked against on adversarial inputs (tests/test_oracle_postproc.py) and (ii) so that the CPU reference arm of
 * bench.py runs the reference's post-processing at compiled speed, as it would with scikit-image's Cython flood, instead of being
 * handicapped by a Python heap.
 *
 * scikit-image (>= 0.26, pinned by the reference's requirements/requirements.txt:29) is absent from this image and not vendored, so the
 * algorithm is restated from its published implementation: skimage/segmentation/_watershed_cy.pyx `watershed_raveled` with connectivity 1,
 * compactness 0, no watershed line — markers outside the mask are dropped; marker pixels enter a priority queue in raveled order with
 * age 0; the queue pops the smallest (value, age); the popped pixel's unlabelled in-mask neighbours, visited in the order (-row, -1, +1,
 * +row), take its label WHEN PUSHED and enter with value image[neighbour] and the next age.  (value, age) ties exist only among marker
 * pixels; scikit-image leaves them to its heap's internals — this restatement, like the Python one and the device code, breaks them by
 * the raveled pixel index, which makes the result independent of the heap's shape.  Parity status: unpinned at this one boundary
 * (SURVEY.md §8c), pinned only by known-answer cases and properties.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double v;
  uint32_t age;
  uint32_t idx;
} item_t;

static inline int item_less(const item_t* a, const item_t* b) {
  if (a->v != b->v) return a->v < b->v;
  if (a->age != b->age) return a->age < b->age;
  return a->idx < b->idx;
}

static void heap_push(item_t* h, long* n, item_t e) {
  long c = (*n)++;
  while (c > 0) {
    long p = (c - 1) / 2;
    if (!item_less(&e, &h[p])) break;
    h[c] = h[p];
    c = p;
  }
  h[c] = e;
}

static item_t heap_pop(item_t* h, long* n) {
  item_t top = h[0];
  item_t last = h[--(*n)];
  long p = 0;
  for (;;) {
    long c = 2 * p + 1;
    if (c >= *n) break;
    if (c + 1 < *n && item_less(&h[c + 1], &h[c])) ++c;
    if (!item_less(&h[c], &last)) break;
    h[p] = h[c];
    p = c;
  }
  if (*n > 0) h[p] = last;
  return top;
}

/* out (H*W int32) receives the labels; returns 0, or -1 when the queue could not be allocated */
int tiab_oracle_watershed(const double* image, const int32_t* markers, const uint8_t* mask, int H, int W, int32_t* out) {
  const long P = (long)H * W;
  item_t* heap = (item_t*)malloc((size_t)(P > 0 ? P : 1) * sizeof(item_t));
  if (!heap) return -1;
  long n = 0;
  for (long p = 0; p < P; ++p) {
    out[p] = mask[p] ? markers[p] : 0;
    if (out[p]) heap_push(heap, &n, (item_t){image[p], 0u, (uint32_t)p});
  }
  uint32_t age = 0;
  while (n > 0) {
    const item_t e = heap_pop(heap, &n);
    const long p = e.idx;
    const int y = (int)(p / W), x = (int)(p % W);
    const int32_t lab = out[p];
    const long nb[4] = {p - W, p - 1, p + 1, p + W};
    const int ok[4] = {y > 0, x > 0, x < W - 1, y < H - 1};
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const long q = nb[k];
      if (!mask[q] || out[q]) continue;
      ++age;
      out[q] = lab;
      heap_push(heap, &n, (item_t){image[q], age, (uint32_t)q});
    }
  }
  free(heap);
  return 0;
}

/* skimage.morphology.remove_small_objects(label_img, max_size=m) for an integer label image: zero every label with <= m pixels, in place */
int tiab_oracle_remove_small_objects(int32_t* lab, long n, int max_size) {
  int32_t mx = 0;
  for (long i = 0; i < n; ++i)
    if (lab[i] > mx) mx = lab[i];
  long* cnt = (long*)calloc((size_t)mx + 1, sizeof(long));
  if (!cnt) return -1;
  for (long i = 0; i < n; ++i)
    if (lab[i] > 0) cnt[lab[i]]++;
  for (long i = 0; i < n; ++i)
    if (lab[i] > 0 && cnt[lab[i]] <= max_size) lab[i] = 0;
  free(cnt);
  return 0;
}

/* CPU restatement (plain C, fp32) of greedy NMS and per-class batched NMS.
 *
 * TEST INFRASTRUCTURE ONLY -- parity oracle, not the product.  Linked/loaded only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Restates torchvision::nms (CPU kernel, torchvision 0.26.0 -- third-party, absent
 * from /root/reference; reached from demonet/models/generalized_ssd.py:389 and
 * demonet/models/box_head.py:374 through torchvision.ops.boxes.batched_nms ->
 * _batched_nms_vanilla).  Semantics pinned in SURVEY.md section 8(a) row N1 and
 * re-checked against the installed torchvision by tests/golden/make_golden.py:
 *   area  = (x2-x1)*(y2-y1)                     fp32
 *   inter = max(0,xx2-xx1)*max(0,yy2-yy1)       fp32
 *   ovr   = inter / ((area_i + area_j) - inter) fp32, no fused multiply-add
 *   suppress j iff (double)ovr > iou_threshold  (strict; NaN never suppresses)
 *   visiting order = stable descending score.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int64_t i; } key_t_;

static int cmp_desc(const void* a, const void* b) {
    const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);      /* stable: lower index first */
}

/* boxes f32[n,4] xyxy, scores f32[n]; keep_out i64[n]; returns number kept */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold,
                   int64_t* keep_out) {
    if (n <= 0) return 0;
    key_t_* ord = (key_t_*)malloc(sizeof(key_t_) * (size_t)n);
    float* area = (float*)malloc(sizeof(float) * (size_t)n);
    unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) { ord[i].s = scores[i]; ord[i].i = i; }
    qsort(ord, (size_t)n, sizeof(key_t_), cmp_desc);
    for (int64_t i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = ord[a].i;
        if (sup[i]) continue;
        keep_out[nk++] = i;
        const float ix1 = boxes[4*i], iy1 = boxes[4*i+1], ix2 = boxes[4*i+2], iy2 = boxes[4*i+3];
        const float iarea = area[i];
        for (int64_t c = a + 1; c < n; ++c) {
            int64_t j = ord[c].i;
            if (sup[j]) continue;
            float xx1 = ix1 > boxes[4*j]   ? ix1 : boxes[4*j];
            float yy1 = iy1 > boxes[4*j+1] ? iy1 : boxes[4*j+1];
            float xx2 = ix2 < boxes[4*j+2] ? ix2 : boxes[4*j+2];
            float yy2 = iy2 < boxes[4*j+3] ? iy2 : boxes[4*j+3];
            float w = xx2 - xx1; w = w > 0.0f ? w : 0.0f;
            float h = yy2 - yy1; h = h > 0.0f ? h : 0.0f;
            float inter = w * h;
            float uni = (iarea + area[j]) - inter;
            float ovr = inter / uni;
            if ((double)ovr > iou_threshold) sup[j] = 1;
        }
    }
    free(ord); free(area); free(sup);
    return nk;
}

/* _batched_nms_vanilla: nms per class on raw coordinates, kept set ordered by
 * (score desc, index asc).  idxs i64[n].  keep_out i64[n]; returns number kept. */
int64_t oracle_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                           double iou_threshold, int64_t* keep_out) {
    if (n <= 0) return 0;
    unsigned char* mask = (unsigned char*)calloc((size_t)n, 1);
    unsigned char* done = (unsigned char*)calloc((size_t)n, 1);
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    int64_t* kk = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    float* cs = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t s = 0; s < n; ++s) {
        if (done[s]) continue;
        int64_t cls = idxs[s], m = 0;
        for (int64_t j = s; j < n; ++j)
            if (!done[j] && idxs[j] == cls) {
                done[j] = 1; cur[m] = j; memcpy(cb + 4*m, boxes + 4*j, 16); cs[m] = scores[j]; ++m;
            }
        int64_t nk = oracle_nms(cb, cs, m, iou_threshold, kk);
        for (int64_t t = 0; t < nk; ++t) mask[cur[kk[t]]] = 1;
    }
    key_t_* ord = (key_t_*)malloc(sizeof(key_t_) * (size_t)n);
    int64_t nk = 0;
    for (int64_t i = 0; i < n; ++i) if (mask[i]) { ord[nk].s = scores[i]; ord[nk].i = i; ++nk; }
    qsort(ord, (size_t)nk, sizeof(key_t_), cmp_desc);
    for (int64_t i = 0; i < nk; ++i) keep_out[i] = ord[i].i;
    free(mask); free(done); free(cur); free(kk); free(cb); free(cs); free(ord);
    return nk;
}

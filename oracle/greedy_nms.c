/*
 * Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the greedy IoU
 * suppression that the reference obtains from a third-party dependency,
 * torchvision.ops.nms (pinned torchvision==0.20.1, reference pyproject.toml:52;
 * call site cerberusdet/utils/general.py:464).  The torchvision source is not in
 * the reference tree, so this restates its published CPU algorithm:
 *
 *   areas = (x2-x1)*(y2-y1); visit boxes in score-descending order (stable);
 *   an unsuppressed box i is kept and suppresses every later unsuppressed j with
 *       inter / (area_i + area_j - inter)  >  iou_threshold
 *   where inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1)), every
 *   operation a separately rounded fp32 operation, the comparison made after
 *   widening the fp32 quotient to double (the threshold is a double).
 *
 * The caller passes boxes already in processing order (the reference sorts by
 * score at general.py:459, so torchvision's own stable sort is the identity).
 * Compile with -ffp-contract=off so no FMA is formed.
 * Pinned by tests/test_oracle_golden.py against the installed torchvision binary.
 */
#include <stdlib.h>

static inline float fmax_std(float a, float b) { return (a < b) ? b : a; } /* std::max */
static inline float fmin_std(float a, float b) { return (b < a) ? b : a; } /* std::min */

/* max_keep > 0: stop as soon as max_keep boxes are kept.  Greedy suppression is causal (whether box i is kept depends
 * only on boxes before it), so the first max_keep entries are exactly those of the full run -- the caller cuts the
 * list there anyway (reference general.py:465 `i = i[:max_det]`); it only spares the O(n^2) tail. */
long oracle_greedy_nms_topk(const float *boxes, long n, double iou_threshold, long *keep, long max_keep);

long oracle_greedy_nms(const float *boxes, long n, double iou_threshold, long *keep) {
    return oracle_greedy_nms_topk(boxes, n, iou_threshold, keep, 0);
}

long oracle_greedy_nms_topk(const float *boxes, long n, double iou_threshold, long *keep, long max_keep) {
    if (n <= 0) return 0;
    unsigned char *dead = (unsigned char *)calloc((size_t)n, 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    long kept = 0;
    for (long i = 0; i < n; ++i) {
        const float *b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    for (long i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[kept++] = i;
        if (max_keep > 0 && kept >= max_keep) break;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1];
        const float ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float ia = area[i];
        for (long j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            const float *q = boxes + 4 * j;
            float xx1 = fmax_std(ix1, q[0]);
            float yy1 = fmax_std(iy1, q[1]);
            float xx2 = fmin_std(ix2, q[2]);
            float yy2 = fmin_std(iy2, q[3]);
            float w = fmax_std(0.0f, xx2 - xx1);
            float h = fmax_std(0.0f, yy2 - yy1);
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > iou_threshold) dead[j] = 1;
        }
    }
    free(dead);
    free(area);
    return kept;
}

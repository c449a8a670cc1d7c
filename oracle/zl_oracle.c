/*
 * zl_oracle.c — CPU restatement of the reference's pre/post-processing.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libzl_b200.so, the
 * host adapter) may link, load or call this file; only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY PINNED for P1 / F1 / N1: the reference ships no tests or golden vectors
 * (SURVEY.md section 4), but preProcess / postProcess / applyNMS / calculateIoU are
 * self-contained C++, so oracle/ref/ cuts exactly those line ranges out of
 * /root/reference/src/inference/onnx_engine.cpp (checksummed) and compiles them
 * UNMODIFIED against a small shim (oracle/_ref/libzl_ref.so).  This file must
 * equal that library bit for bit (tests/test_oracle_vs_ref.py) and the golden
 * fixtures minted from it (tests/golden/, tests/test_golden.py).  One behaviour
 * of the reference is implementation-defined: std::sort's order inside an exact
 * (class, confidence) tie group; this file completes it with the anchor index.
 * The conv graph itself (M1 + D1: ONNX Runtime + an exported model, neither in
 * the tree) stays unpinned: oracle/yolov8_ref.py restates the public YOLOv8
 * definition.  Citations are relative to the reference tree.
 *
 * Build WITHOUT -ffast-math and with -ffp-contract=off (oracle/Makefile): the
 * reference's Release flags use -ffast-math (CMakeLists.txt:279), which leaves
 * its own rounding unspecified; this oracle pins IEEE-754 single precision,
 * round-to-nearest, no FMA contraction, and the CUDA kernels match that.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ZLO_OK 0
#define ZLO_INVALID_INPUT 203 /* ErrorCode::INVALID_INPUT, src/common/result.h:33 */

typedef struct zlo_det {
    float x, y, w, h;   /* BoundingBox, src/common/types.h:16-18 (centre format) */
    float confidence;   /* Detection::confidence, types.h:22 */
    int32_t class_id;   /* Detection::class_id, types.h:23 */
} zlo_det;

/* preProcess, src/inference/onnx_engine.cpp:649-700 (== preProcessZeroCopy :703-755).
 * Nearest-neighbour STRETCH (no letterbox), BGR->RGB by reading 2-c, /255.0f,
 * output planar [3][mh][mw] fp32. */
int zlo_preprocess(const uint8_t* img, size_t len, int width, int height,
                   int mw, int mh, float* out_chw)
{
    /* :659-665 size check */
    if (len != (size_t)width * (size_t)height * 3u) return ZLO_INVALID_INPUT;
    /* :673-674 */
    float scale_w = (float)width / mw;
    float scale_h = (float)height / mh;
    for (int c = 0; c < 3; c++) {                    /* :677 */
        for (int h = 0; h < mh; h++) {               /* :678 */
            for (int w = 0; w < mw; w++) {           /* :679 */
                int src_h = (int)(h * scale_h);      /* :681 float mul, truncation */
                if (src_h > height - 1) src_h = height - 1;
                int src_w = (int)(w * scale_w);      /* :682 */
                if (src_w > width - 1) src_w = width - 1;
                size_t src_idx = ((size_t)src_h * width + src_w) * 3 + (2 - c); /* :685 */
                size_t dst_idx = (size_t)c * mh * mw + (size_t)h * mw + w;      /* :688 */
                if (src_idx < len)                                              /* :691 */
                    out_chw[dst_idx] = img[src_idx] / 255.0f;                   /* :693 */
            }
        }
    }
    return ZLO_OK;
}

/* Source index table of the stretch, for the known-answer tests. */
void zlo_stretch_index(int src_dim, int dst_dim, int32_t* idx_out)
{
    float scale = (float)src_dim / dst_dim;
    for (int i = 0; i < dst_dim; i++) {
        int s = (int)(i * scale);
        if (s > src_dim - 1) s = src_dim - 1;
        idx_out[i] = s;
    }
}

/* calculateIoU, src/inference/onnx_engine.cpp:881-909. */
float zlo_iou(const zlo_det* a, const zlo_det* b)
{
    float x1_min = a->x - a->w / 2;   /* :883-886 */
    float y1_min = a->y - a->h / 2;
    float x1_max = a->x + a->w / 2;
    float y1_max = a->y + a->h / 2;
    float x2_min = b->x - b->w / 2;   /* :888-891 */
    float y2_min = b->y - b->h / 2;
    float x2_max = b->x + b->w / 2;
    float y2_max = b->y + b->h / 2;
    float x_overlap = fmaxf(0.0f, fminf(x1_max, x2_max) - fmaxf(x1_min, x2_min)); /* :894 */
    float y_overlap = fmaxf(0.0f, fminf(y1_max, y2_max) - fmaxf(y1_min, y2_min)); /* :895 */
    float intersection = x_overlap * y_overlap;                                   /* :896 */
    float area1 = a->w * a->h;                                                    /* :899 */
    float area2 = b->w * b->h;
    float union_area = area1 + area2 - intersection;                              /* :901 */
    if (union_area > 0) return intersection / union_area;                         /* :904-906 */
    return 0.0f;
}

/* Decode part of postProcess, src/inference/onnx_engine.cpp:773-819.
 * raw = [4+nc][A] fp32 (one frame).  Writes candidates in ANCHOR ORDER and
 * their anchor indices; returns the count.  out/anchor_out need room for A. */
int zlo_decode_filter(const float* raw, int nc, int A, int img_w, int img_h,
                      float conf_thr, zlo_det* out, int32_t* anchor_out)
{
    int n = 0;
    for (int i = 0; i < A; i++) {                  /* :779 */
        float cx = raw[0 * (size_t)A + i];         /* :781-784 */
        float cy = raw[1 * (size_t)A + i];
        float w  = raw[2 * (size_t)A + i];
        float h  = raw[3 * (size_t)A + i];
        float max_conf = 0.0f;                     /* :787 */
        int max_class_id = -1;                     /* :788 */
        for (int j = 0; j < nc; j++) {             /* :790 */
            float s = raw[(size_t)(j + 4) * A + i];
            if (s > max_conf) {                    /* :792 strict: first max wins */
                max_conf = s;
                max_class_id = j;
            }
        }
        if (max_conf >= conf_thr && max_class_id >= 0) {   /* :799 */
            zlo_det d;
            d.x = cx / img_w;                      /* :802-805: REQUEST frame dims */
            d.y = cy / img_h;
            d.w = w / img_w;
            d.h = h / img_h;
            d.confidence = max_conf;
            d.class_id = max_class_id;
            out[n] = d;
            if (anchor_out) anchor_out[n] = i;
            n++;
        }
    }
    return n;
}

typedef struct { zlo_det d; int32_t anchor; } zlo_keyed;

/* Sort order of applyNMS, onnx_engine.cpp:846-851: class asc, confidence desc.
 * The reference uses std::sort, which leaves the order of equal (class, conf)
 * pairs unspecified; this oracle (and the CUDA kernel) complete it to a total
 * order with the anchor index ascending, i.e. what a stable sort of the
 * anchor-ordered candidate list yields. */
static int zlo_cmp(const void* pa, const void* pb)
{
    const zlo_keyed* a = (const zlo_keyed*)pa;
    const zlo_keyed* b = (const zlo_keyed*)pb;
    if (a->d.class_id != b->d.class_id) return a->d.class_id < b->d.class_id ? -1 : 1;
    if (a->d.confidence != b->d.confidence) return a->d.confidence > b->d.confidence ? -1 : 1;
    if (a->anchor != b->anchor) return a->anchor < b->anchor ? -1 : 1;
    return 0;
}

/* applyNMS, src/inference/onnx_engine.cpp:837-878.  dets/anchors: n candidates
 * in anchor order.  out (and anchor_out, optional) receive the kept list in
 * sorted order.  Returns the kept count. */
int zlo_nms(const zlo_det* dets, const int32_t* anchors, int n, float iou_thr,
            zlo_det* out, int32_t* anchor_out)
{
    if (n <= 1) {                                   /* :841-843 */
        for (int i = 0; i < n; i++) {
            out[i] = dets[i];
            if (anchor_out) anchor_out[i] = anchors ? anchors[i] : i;
        }
        return n;
    }
    zlo_keyed* v = (zlo_keyed*)malloc(sizeof(zlo_keyed) * (size_t)n);
    uint8_t* removed = (uint8_t*)calloc((size_t)n, 1);  /* :853 */
    for (int i = 0; i < n; i++) { v[i].d = dets[i]; v[i].anchor = anchors ? anchors[i] : i; }
    qsort(v, (size_t)n, sizeof(zlo_keyed), zlo_cmp);    /* :846-851 (total order, see zlo_cmp) */
    int kept = 0;
    for (int i = 0; i < n; i++) {                   /* :856 */
        if (removed[i]) continue;                   /* :857-859 */
        int cur = v[i].d.class_id;                  /* :861 */
        out[kept] = v[i].d;                         /* :862 */
        if (anchor_out) anchor_out[kept] = v[i].anchor;
        kept++;
        for (int j = i + 1; j < n; j++) {           /* :865 */
            if (removed[j] || v[j].d.class_id != cur) continue;   /* :866-868 */
            float iou = zlo_iou(&v[i].d, &v[j].d);  /* :870 */
            if (iou > iou_thr) removed[j] = 1;      /* :871 strict */
        }
    }
    free(v);
    free(removed);
    return kept;
}

/* postProcess, src/inference/onnx_engine.cpp:758-834: decode + (if non-empty) NMS. */
int zlo_postprocess(const float* raw, int nc, int A, int img_w, int img_h,
                    float conf_thr, float iou_thr, zlo_det* out, int32_t* anchor_out)
{
    zlo_det* cand = (zlo_det*)malloc(sizeof(zlo_det) * (size_t)(A > 0 ? A : 1));
    int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(A > 0 ? A : 1));
    int n = zlo_decode_filter(raw, nc, A, img_w, img_h, conf_thr, cand, anc);
    int kept = 0;
    if (n > 0) kept = zlo_nms(cand, anc, n, iou_thr, out, anchor_out);   /* :822-824 */
    free(cand);
    free(anc);
    return kept;
}

/* Batched convenience for the CPU baseline: n frames, raw = [n][4+nc][A]. */
long zlo_postprocess_batch(const float* raw, int n, int nc, int A, const int32_t* img_w,
                           const int32_t* img_h, float conf_thr, float iou_thr,
                           zlo_det* out, int32_t* counts)
{
    long total = 0;
    for (int f = 0; f < n; f++) {
        int k = zlo_postprocess(raw + (size_t)f * (4 + nc) * A, nc, A, img_w[f], img_h[f],
                                conf_thr, iou_thr, out + total, NULL);
        counts[f] = k;
        total += k;
    }
    return total;
}

/*
 * roi3d_oracle.c -- CPU restatement of the reference's ROI hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or the reported CPU baseline.  The product path (3d-mask-r-cnn_b200/) never
 * links or imports this file.
 *
 * What it restates.  The reference (podtyazhki1337/3d-mask-r-cnn) ships its
 * four native ops only as a prebuilt wheel
 * (core/custom_op/tensorflow_nms_car_3d-0.1.0-cp36-cp36m-linux_x86_64.whl);
 * there is no C source in the tree.  Every function below follows the
 * machine code of those binaries (addresses cited per function, notation
 * <lib>@0xADDR as in SURVEY.md section 0) and the Python call sites in
 * core/custom_op/custom_op.py:22-65 and core/models.py:450-456, 663-664.
 *
 * Parity pin: PINNED.  oracle/refrun/ maps the wheel's four shared objects
 * into this process (its own ELF loader + stand-ins for the ~25 TensorFlow
 * call-outs) and runs the reference's OWN Compute() functions and its own
 * IOU<float>.  tests/test_oracle_pin.py asserts that every function below is
 * bit-identical to them (forward, grad-image, grad-boxes, NMS, 200k IoU
 * pairs; edge cases included), and the tests/golden npz files were written by
 * `make_golden.py --check-ref`, i.e. only after the same check passed.
 * Two reference defects found that way are documented where they apply:
 * the grad-boxes depth step (restated faithfully) and the `nearest` forward
 * loop bound (undefined behaviour upstream for crop_width != crop_depth).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile): every
 * fp32 operation below is a separately rounded IEEE op, like the reference's
 * scalar SSE code (mulss/addss/subss/divss, no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * IOU<float>(boxes, i, j)                          NMS.so@0xb500-0xb656
 * boxes: [N,6] = (y1,x1,z1,y2,x2,z2), any corner order.
 * ---------------------------------------------------------------------- */
static inline float minf_(float a, float b) { return a < b ? a : b; } /* minss */
static inline float maxf_(float a, float b) { return a > b ? a : b; } /* maxss */

ORACLE_API float roi3d_oracle_iou3d(const float *boxes, int i, int j)
{
    const float *a = boxes + (size_t)i * 6;
    const float *b = boxes + (size_t)j * 6;
    const float ymin_i = minf_(a[0], a[3]), ymax_i = maxf_(a[0], a[3]);
    const float xmin_i = minf_(a[1], a[4]), xmax_i = maxf_(a[1], a[4]);
    const float zmin_i = minf_(a[2], a[5]), zmax_i = maxf_(a[2], a[5]);
    const float ymin_j = minf_(b[0], b[3]), ymax_j = maxf_(b[0], b[3]);
    const float xmin_j = minf_(b[1], b[4]), xmax_j = maxf_(b[1], b[4]);
    const float zmin_j = minf_(b[2], b[5]), zmax_j = maxf_(b[2], b[5]);
    const float area_i = ((ymax_i - ymin_i) * (xmax_i - xmin_i)) * (zmax_i - zmin_i);
    if (area_i <= 0.0f) return 0.0f;
    const float area_j = ((ymax_j - ymin_j) * (xmax_j - xmin_j)) * (zmax_j - zmin_j);
    if (area_j <= 0.0f) return 0.0f;
    const float dy = maxf_(0.0f, minf_(ymax_i, ymax_j) - maxf_(ymin_i, ymin_j));
    const float dx = maxf_(0.0f, minf_(xmax_i, xmax_j) - maxf_(xmin_i, xmin_j));
    const float dz = maxf_(0.0f, minf_(zmax_i, zmax_j) - maxf_(zmin_i, zmin_j));
    const float inter = (dy * dx) * dz;
    return inter / ((area_i + area_j) - inter);
}

/* ------------------------------------------------------------------------
 * DoNonMaxSuppressionOp<float>                     NMS.so@0xd0c0-0xe4e0
 * called by NonMaxSuppression3DOp<CPUDevice>::Compute (NMS.so@0xe4e0) with
 * score_threshold = -FLT_MAX, soft_nms_sigma = 0, return_scores = false,
 * pad_to_max_output_size = false.
 *
 * Candidate {box_index, score, suppress_begin_index}; max-heap whose top is
 * the highest score, ties -> lower box_index (comparator recovered from
 * __push_heap NMS.so@0xc6be-0xc6d8).
 * Returns the number of selected indices written to `selected` (capacity
 * must be >= max(max_out, 0)).  Indices may repeat (zero-volume quirk, see
 * SURVEY.md section 8 row a2).
 * ---------------------------------------------------------------------- */
typedef struct { int box_index; float score; int suppress_begin_index; } cand_t;

/* "a sorts below b" */
static inline int cand_less(const cand_t *a, const cand_t *b)
{
    return (a->score < b->score) || (a->score == b->score && a->box_index > b->box_index);
}

static void heap_push(cand_t *h, size_t *n, cand_t c)
{
    size_t i = (*n)++;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!cand_less(&h[p], &c)) break;
        h[i] = h[p];
        i = p;
    }
    h[i] = c;
}

static cand_t heap_pop(cand_t *h, size_t *n)
{
    cand_t top = h[0];
    cand_t last = h[--(*n)];
    size_t i = 0, cnt = *n;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m;
        if (l >= cnt) break;
        m = (r < cnt && cand_less(&h[l], &h[r])) ? r : l;
        if (!cand_less(&last, &h[m])) break;
        h[i] = h[m];
        i = m;
    }
    if (cnt) h[i] = last;
    return top;
}

ORACLE_API int roi3d_oracle_nms3d(const float *boxes, const float *scores, int n,
                                  int max_out, float iou_threshold, int *selected)
{
    const float score_threshold = -FLT_MAX;     /* NMS.so .rodata @0x1cab0 */
    const float scale = 0.0f;                   /* soft_nms_sigma == 0 */
    size_t hn = 0;
    int nsel = 0;
    if (n <= 0 || max_out <= 0) return 0;
    cand_t *heap = (cand_t *)malloc(sizeof(cand_t) * (size_t)n);
    for (int i = 0; i < n; ++i) {               /* @0xd63d-0xd78d */
        if (scores[i] > score_threshold) {
            cand_t c = { i, scores[i], 0 };
            heap_push(heap, &hn, c);
        }
    }
    while (nsel < max_out && hn > 0) {          /* @0xd832-0xd853, 0xdab9-0xdae3 */
        cand_t c = heap_pop(heap, &hn);
        const float original_score = c.score;
        int hard = 0;
        for (int j = nsel - 1; j >= c.suppress_begin_index; --j) {   /* @0xda47-0xdaa1 */
            const float sim = roi3d_oracle_iou3d(boxes, c.box_index, selected[j]);
            const float w = (sim <= iou_threshold) ? expf(scale * sim * sim) : 0.0f;
            c.score *= w;
            if (sim >= iou_threshold) { hard = 1; break; }          /* ucomiss; jb continue */
            if (c.score <= score_threshold) break;
        }
        c.suppress_begin_index = nsel;          /* @0xdc28-0xdc43 */
        if (!hard) {
            if (c.score == original_score)      /* @0xdc4a-0xdc4f */
                selected[nsel++] = c.box_index;
            /* TF r2.0-2.2: no `continue` after selecting -> re-queued */
            if (c.score > score_threshold)      /* @0xdc55-0xdc5a */
                heap_push(heap, &hn, c);
        }
    }
    free(heap);
    return nsel;
}

/* ------------------------------------------------------------------------
 * Sample coordinate along one axis               CAR.so@0x499a-0x4ae5, 0x533a
 * ---------------------------------------------------------------------- */
static inline float axis_scale(float a1, float a2, int dim, int p)
{
    return (p > 1) ? ((a2 - a1) * (float)(dim - 1)) / (float)(p - 1) : 0.0f;
}
static inline float axis_coord(float a1, float a2, int dim, int p, int k, float scale)
{
    if (p > 1) return a1 * (float)(dim - 1) + (float)k * scale;
    return (float)((double)(a1 + a2) * 0.5 * (double)(dim - 1));
}
static inline int axis_invalid(float in, int dim)
{
    return in < 0.0f || in > (float)(dim - 1);
}

/* ------------------------------------------------------------------------
 * CropAndResize3DOp::Compute                      CAR.so@0x4370-0x5a90
 * image [B,H,W,D,C], boxes [N,6], box_index [N] -> crops [N,ph,pw,pd,C]
 * method: 0 trilinear, 1 nearest.  `threads`<=1 -> the reference's single
 * thread; >1 -> OpenMP over boxes (same arithmetic, used for the generous
 * CPU baseline only).
 * ---------------------------------------------------------------------- */
static void car3d_fwd_box(const float *image, int H, int W, int D, int C,
                          const float *box, int b_in, int ph, int pw, int pd,
                          int method, float ext, float *crop /* [ph,pw,pd,C] */)
{
    const float y1 = box[0], x1 = box[1], z1 = box[2];
    const float y2 = box[3], x2 = box[4], z2 = box[5];
    const float hs = axis_scale(y1, y2, H, ph);
    const float ws = axis_scale(x1, x2, W, pw);
    const float ds = axis_scale(z1, z2, D, pd);
    const size_t sD = (size_t)C, sW = (size_t)D * C, sH = (size_t)W * D * C;
    const float *img = image + (size_t)b_in * H * sH;
    for (int y = 0; y < ph; ++y) {
        const float in_y = axis_coord(y1, y2, H, ph, y, hs);
        float *oy = crop + (size_t)y * pw * pd * C;
        if (axis_invalid(in_y, H)) {             /* @0x52be-0x5335 */
            for (size_t e = 0; e < (size_t)pw * pd * C; ++e) oy[e] = ext;
            continue;
        }
        for (int x = 0; x < pw; ++x) {
            const float in_x = axis_coord(x1, x2, W, pw, x, ws);
            float *ox = oy + (size_t)x * pd * C;
            if (axis_invalid(in_x, W)) {         /* @0x5223-0x5282 */
                for (size_t e = 0; e < (size_t)pd * C; ++e) ox[e] = ext;
                continue;
            }
            for (int z = 0; z < pd; ++z) {
                const float in_z = axis_coord(z1, z2, D, pd, z, ds);
                float *oz = ox + (size_t)z * C;
                if (axis_invalid(in_z, D)) {     /* @0x51b0-0x51e7 */
                    for (int c = 0; c < C; ++c) oz[c] = ext;
                    continue;
                }
                if (method == 0) {
                    const int t = (int)floorf(in_y), bo = (int)ceilf(in_y);
                    const int l = (int)floorf(in_x), r = (int)ceilf(in_x);
                    const int f = (int)floorf(in_z), ce = (int)ceilf(in_z);
                    const float yl = in_y - (float)t, xl = in_x - (float)l, zl = in_z - (float)f;
                    const float *ptlf = img + t * sH + l * sW + f * sD, *ptlc = img + t * sH + l * sW + ce * sD;
                    const float *ptrf = img + t * sH + r * sW + f * sD, *ptrc = img + t * sH + r * sW + ce * sD;
                    const float *pblf = img + bo * sH + l * sW + f * sD, *pblc = img + bo * sH + l * sW + ce * sD;
                    const float *pbrf = img + bo * sH + r * sW + f * sD, *pbrc = img + bo * sH + r * sW + ce * sD;
                    for (int c = 0; c < C; ++c) { /* @0x4f88-0x5021: z, then x, then y */
                        const float tl = ptlf[c] + (ptlc[c] - ptlf[c]) * zl;
                        const float tr = ptrf[c] + (ptrc[c] - ptrf[c]) * zl;
                        const float bl = pblf[c] + (pblc[c] - pblf[c]) * zl;
                        const float br = pbrf[c] + (pbrc[c] - pbrf[c]) * zl;
                        const float top = tl + (tr - tl) * xl;
                        const float bot = bl + (br - bl) * xl;
                        oz[c] = top + (bot - top) * yl;
                    }
                } else {                          /* nearest @0x54c2-0x54e8.  NB: the reference's nearest branch bounds its
                                                   * z loop by crop_width (cmp [rbp-0xa4] @0x5570), so for pw != pd it leaves
                                                   * crop voxels unwritten (pw < pd) or writes past the row (pw > pd):
                                                   * undefined behaviour.  For pw == pd (every caller) it is this code. */
                    const int yi = (int)roundf(in_y), xi = (int)roundf(in_x), zi = (int)roundf(in_z);
                    const float *p = img + yi * sH + xi * sW + zi * sD;
                    for (int c = 0; c < C; ++c) oz[c] = p[c];
                }
            }
        }
    }
}

ORACLE_API void roi3d_oracle_car3d_fwd(const float *image, int B, int H, int W, int D, int C,
                                       const float *boxes, const int *box_index, int n,
                                       int ph, int pw, int pd, int method, float ext,
                                       float *crops, int threads)
{
    (void)B;
    const size_t per = (size_t)ph * pw * pd * C;
#ifdef _OPENMP
    if (threads > 1) {
        #pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int b = 0; b < n; ++b)
            car3d_fwd_box(image, H, W, D, C, boxes + (size_t)b * 6, box_index[b], ph, pw, pd,
                          method, ext, crops + (size_t)b * per);
        return;
    }
#else
    (void)threads;
#endif
    for (int b = 0; b < n; ++b)
        car3d_fwd_box(image, H, W, D, C, boxes + (size_t)b * 6, box_index[b], ph, pw, pd,
                      method, ext, crops + (size_t)b * per);
}

/* ------------------------------------------------------------------------
 * CropAndResize3DGradImageOp::Compute              GI.so@0x3a80-0x5230
 * grads [N,ph,pw,pd,C] -> out [B,H,W,D,C] (zero-filled first, @0x3ec5).
 * Channel range [c0,c1) lets the OpenMP baseline split the channel axis
 * without changing the per-element accumulation order.
 * ---------------------------------------------------------------------- */
static void car3d_grad_image_range(const float *grads, const float *boxes, const int *box_ind, int n,
                                   int ph, int pw, int pd, int H, int W, int D, int C,
                                   int method, float *out, int c0, int c1)
{
    const size_t sD = (size_t)C, sW = (size_t)D * C, sH = (size_t)W * D * C;
    for (int b = 0; b < n; ++b) {
        const float *box = boxes + (size_t)b * 6;
        const float y1 = box[0], x1 = box[1], z1 = box[2];
        const float y2 = box[3], x2 = box[4], z2 = box[5];
        const float hs = axis_scale(y1, y2, H, ph);
        const float ws = axis_scale(x1, x2, W, pw);
        const float ds = axis_scale(z1, z2, D, pd);
        float *img = out + (size_t)box_ind[b] * H * sH;
        for (int y = 0; y < ph; ++y) {
            const float in_y = axis_coord(y1, y2, H, ph, y, hs);
            if (axis_invalid(in_y, H)) continue;             /* @0x4244 */
            const int t = (int)floorf(in_y), bo = (int)ceilf(in_y);
            const float yl = in_y - (float)t;
            for (int x = 0; x < pw; ++x) {
                const float in_x = axis_coord(x1, x2, W, pw, x, ws);
                if (axis_invalid(in_x, W)) continue;         /* @0x43ac */
                const int l = (int)floorf(in_x), r = (int)ceilf(in_x);
                const float xl = in_x - (float)l;
                for (int z = 0; z < pd; ++z) {
                    const float in_z = axis_coord(z1, z2, D, pd, z, ds);
                    if (axis_invalid(in_z, D)) continue;     /* @0x451e */
                    const float *g = grads + ((((size_t)b * ph + y) * pw + x) * pd + z) * C;
                    if (method == 0) {
                        const int f = (int)floorf(in_z), ce = (int)ceilf(in_z);
                        const float zl = in_z - (float)f;
                        /* weights ((wy*wx)*wz), corner order tlf,tlc,trf,trc,blf,blc,brf,brc
                           (@0x4692-0x470c, @0x4740-0x47ea) */
                        const float wt = 1.0f - yl, wb = yl, wl = 1.0f - xl, wr = xl, wf = 1.0f - zl, wc = zl;
                        const float w_tlf = (wt * wl) * wf, w_tlc = (wt * wl) * wc;
                        const float w_trf = (wt * wr) * wf, w_trc = (wt * wr) * wc;
                        const float w_blf = (wb * wl) * wf, w_blc = (wb * wl) * wc;
                        const float w_brf = (wb * wr) * wf, w_brc = (wb * wr) * wc;
                        float *ptlf = img + t * sH + l * sW + f * sD, *ptlc = img + t * sH + l * sW + ce * sD;
                        float *ptrf = img + t * sH + r * sW + f * sD, *ptrc = img + t * sH + r * sW + ce * sD;
                        float *pblf = img + bo * sH + l * sW + f * sD, *pblc = img + bo * sH + l * sW + ce * sD;
                        float *pbrf = img + bo * sH + r * sW + f * sD, *pbrc = img + bo * sH + r * sW + ce * sD;
                        for (int c = c0; c < c1; ++c) {
                            const float gv = g[c];
                            ptlf[c] += gv * w_tlf; ptlc[c] += gv * w_tlc;
                            ptrf[c] += gv * w_trf; ptrc[c] += gv * w_trc;
                            pblf[c] += gv * w_blf; pblc[c] += gv * w_blc;
                            pbrf[c] += gv * w_brf; pbrc[c] += gv * w_brc;
                        }
                    } else {
                        const int yi = (int)roundf(in_y), xi = (int)roundf(in_x), zi = (int)roundf(in_z);
                        float *p = img + yi * sH + xi * sW + zi * sD;
                        for (int c = c0; c < c1; ++c) p[c] += g[c];
                    }
                }
            }
        }
    }
}

ORACLE_API void roi3d_oracle_car3d_grad_image(const float *grads, const float *boxes, const int *box_ind,
                                              int n, int ph, int pw, int pd,
                                              int B, int H, int W, int D, int C,
                                              int method, float *out, int threads)
{
    memset(out, 0, sizeof(float) * (size_t)B * H * W * D * C);
#ifdef _OPENMP
    if (threads > 1 && C >= threads) {
        #pragma omp parallel num_threads(threads)
        {
            const int t = omp_get_thread_num(), nt = omp_get_num_threads();
            const int c0 = (int)((long long)C * t / nt), c1 = (int)((long long)C * (t + 1) / nt);
            car3d_grad_image_range(grads, boxes, box_ind, n, ph, pw, pd, H, W, D, C, method, out, c0, c1);
        }
        return;
    }
#else
    (void)threads;
#endif
    car3d_grad_image_range(grads, boxes, box_ind, n, ph, pw, pd, H, W, D, C, method, out, 0, C);
}

/* ------------------------------------------------------------------------
 * CropAndResize3DGradBoxesOp::Compute              GB.so@0x3980-0x51c0
 * grads [N,ph,pw,pd,C], image [B,H,W,D,C] -> out [N,6] (zero-initialised),
 * trilinear only.  Inner loop @0x46a4-0x48e9.
 * ---------------------------------------------------------------------- */
ORACLE_API void roi3d_oracle_car3d_grad_boxes(const float *grads, const float *image,
                                              int B, int H, int W, int D, int C,
                                              const float *boxes, const int *box_ind, int n,
                                              int ph, int pw, int pd, float *out)
{
    (void)B;
    const size_t sD = (size_t)C, sW = (size_t)D * C, sH = (size_t)W * D * C;
    const float rh = (ph > 1) ? (float)(H - 1) / (float)(ph - 1) : 0.0f;        /* @0x3f4a-0x3f65 */
    const float rw = (pw > 1) ? (float)(W - 1) / (float)(pw - 1) : 0.0f;
    const float rd = (pd > 1) ? (float)(D - 1) / (float)(pd - 1) : 0.0f;
    memset(out, 0, sizeof(float) * (size_t)n * 6);
    for (int b = 0; b < n; ++b) {
        const float *box = boxes + (size_t)b * 6;
        const float y1 = box[0], x1 = box[1], z1 = box[2];
        const float y2 = box[3], x2 = box[4], z2 = box[5];
        /* Unlike the forward and grad-image kernels, this op forms the sample step as
         * (a2 - a1) * ratio (GB.so@0x3ff6-0x4071, the TF crop_and_resize_op.cc form).
         * REFERENCE QUIRK, restated on purpose: the depth step is computed from the wrong
         * operands, (z2 - y1) * ratio_h instead of (z2 - z1) * ratio_d (GB.so@0x4059-0x4071
         * loads box[5], box[0] and the height ratio).  The op is never executed by
         * core/models.py (boxes are stop_gradient'ed, core/models.py:660), so the slip went
         * unnoticed upstream; parity means reproducing it. */
        const float hs = (ph > 1) ? (y2 - y1) * rh : 0.0f;
        const float ws = (pw > 1) ? (x2 - x1) * rw : 0.0f;
        const float ds = (pd > 1) ? (z2 - y1) * rh : 0.0f;
        const float *img = image + (size_t)box_ind[b] * H * sH;
        float *o = out + (size_t)b * 6;
        for (int y = 0; y < ph; ++y) {
            const float in_y = axis_coord(y1, y2, H, ph, y, hs);
            if (axis_invalid(in_y, H)) continue;
            const int t = (int)floorf(in_y), bo = (int)ceilf(in_y);
            const float yl = in_y - (float)t;
            for (int x = 0; x < pw; ++x) {
                const float in_x = axis_coord(x1, x2, W, pw, x, ws);
                if (axis_invalid(in_x, W)) continue;
                const int l = (int)floorf(in_x), r = (int)ceilf(in_x);
                const float xl = in_x - (float)l;
                for (int z = 0; z < pd; ++z) {
                    const float in_z = axis_coord(z1, z2, D, pd, z, ds);
                    if (axis_invalid(in_z, D)) continue;
                    const int f = (int)floorf(in_z), ce = (int)ceilf(in_z);
                    const float zl = in_z - (float)f;
                    const float *g = grads + ((((size_t)b * ph + y) * pw + x) * pd + z) * C;
                    const float *ptlf = img + t * sH + l * sW + f * sD, *ptlc = img + t * sH + l * sW + ce * sD;
                    const float *ptrf = img + t * sH + r * sW + f * sD, *ptrc = img + t * sH + r * sW + ce * sD;
                    const float *pblf = img + bo * sH + l * sW + f * sD, *pblc = img + bo * sH + l * sW + ce * sD;
                    const float *pbrf = img + bo * sH + r * sW + f * sD, *pbrc = img + bo * sH + r * sW + ce * sD;
                    for (int c = 0; c < C; ++c) {
                        const float tlf = ptlf[c], tlc = ptlc[c], trf = ptrf[c], trc = ptrc[c];
                        const float blf = pblf[c], blc = pblc[c], brf = pbrf[c], brc = pbrc[c];
                        float gy = ((blf - tlf) * (1.0f - xl) + (brf - trf) * xl) * (1.0f - zl)
                                 + ((blc - tlc) * (1.0f - xl) + (brc - trc) * xl) * zl;
                        float gx = ((trf - tlf) * (1.0f - yl) + (brf - blf) * yl) * (1.0f - zl)
                                 + ((trc - tlc) * (1.0f - yl) + (brc - blc) * yl) * zl;
                        float gz = ((tlc - tlf) * (1.0f - yl) + (blc - blf) * yl) * (1.0f - xl)
                                 + ((trc - trf) * (1.0f - yl) + (brc - brf) * yl) * xl;
                        const float tg = g[c];
                        gy *= tg; gx *= tg; gz *= tg;
                        if (ph > 1) {
                            o[0] += gy * ((float)(H - 1) - (float)y * rh);
                            o[3] += (gy * (float)y) * rh;
                        } else {
                            /* accumulated in double, rounded once (GB.so@0x47d0-0x4829) */
                            const double v = (double)gy * 0.5 * (double)(H - 1);
                            o[0] = (float)((double)o[0] + v); o[3] = (float)(v + (double)o[3]);
                        }
                        if (pw > 1) {
                            o[1] += gx * ((float)(W - 1) - (float)x * rw);
                            o[4] += (gx * (float)x) * rw;
                        } else {
                            /* accumulated in double, rounded once (GB.so@0x47d0-0x4829) */
                            const double v = (double)gx * 0.5 * (double)(W - 1);
                            o[1] = (float)((double)o[1] + v); o[4] = (float)(v + (double)o[4]);
                        }
                        if (pd > 1) {
                            o[2] += gz * ((float)(D - 1) - (float)z * rd);
                            o[5] += (gz * (float)z) * rd;
                        } else {
                            /* accumulated in double, rounded once (GB.so@0x47d0-0x4829) */
                            const double v = (double)gz * 0.5 * (double)(D - 1);
                            o[2] = (float)((double)o[2] + v); o[5] = (float)(v + (double)o[5]);
                        }
                    }
                }
            }
        }
    }
}

/* Pairwise IoU over [n] x [m] boxes -- helper for tests (not a reference op). */
ORACLE_API void roi3d_oracle_iou_matrix(const float *boxes, int n, float *out /* [n,n] */)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            out[(size_t)i * n + j] = roi3d_oracle_iou3d(boxes, i, j);
}

ORACLE_API int roi3d_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

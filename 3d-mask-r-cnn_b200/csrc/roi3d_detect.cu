// roi3d_detect.cu -- the callers and data formats either side of the hot ops.
//
//   refine_detections : DetectionLayer / refine_detections_graph (core/models.py:1415-1524) for a whole batch in one
//                       set of launches: class-1 score + deltas, confidence filter, apply_box_deltas_3d_graph in pixel
//                       space (core/utils.py:412-464), clip to the image, min sizes, NMS, score order, normalise,
//                       zero-pad.  NMS modes: ROI3D_NMS_REFERENCE_2D (default) = what the fork's graph does, the 2-D
//                       tf.image.non_max_suppression on the (y, x) projection, suppressing on IoU > thr (:1496-1501),
//                       run on the 3-D kernels with every box given z = [0, 1] (IoU3D == IoU2D bit for bit) and the
//                       threshold moved one ulp up (iou > t  <=>  iou >= nextafter(t)); ROI3D_NMS_3D = the 3-D op with
//                       its own >= rule, the upstream design row f3 asks for (opt-in).  Filters are
//                       folded into the NMS candidate rule (score := -FLT_MAX), so nothing is compacted and no count
//                       ever travels to the host.
//   mask_targets      : detection_targets_graph._get_masks (core/models.py:972-1005): CropAndResize3D of the assigned
//                       ground-truth mask (C = 1) followed by tf.round; reads uint8 or float32 masks through the
//                       assignment index (no gather, no cast pass) and can emit the bit-packed on-disk form directly.
//   wire format       : the 3-stage pipeline's target files (core/models.py:3585-3636): rois_aligned as float16
//                       (ndarray.astype(float16): round-to-nearest-even) and masks as numpy.packbits(x > 0.5)
//                       (MSB first), plus the inverse kernels for the reader side.
// Everything here is elementwise or a gather: coalesced, grid sized from the SM count, no shared state.
#include "roi3d_common.cuh"
#include <cuda_fp16.h>
#include <cfloat>
#include <math.h>

namespace roi3d {

// ---------------------------------------------------------------------------------
// refine_detections
// ---------------------------------------------------------------------------------
struct RefineParams {
    float dim[3];            // image H, W, D in pixels (float32 like the graph's image_shape)
    float std[6];            // BBOX_STD_DEV
    float min_conf;
    float log_limit;         // float32(log(1000/16)), core/utils.py:438
};

__device__ __forceinline__ float clip_rn(float v, float lo, float hi) { return fmaxf(fminf(v, hi), lo); }

__global__ void __launch_bounds__(256)
refine_decode_kernel(const float *__restrict__ rois, const float *__restrict__ probs, const float *__restrict__ deltas,
                     int total, int rois_per_image, int images, int num_classes, RefineParams P,
                     float *__restrict__ boxes_px, float *__restrict__ boxes_nms, float *__restrict__ scores,
                     int *__restrict__ seg_offsets)
{
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= images) seg_offsets[i] = i * rois_per_image;
    if (i >= total) return;
    const float score = __ldg(probs + (size_t)i * num_classes + 1);           // fg_probs = probs[:, 1]   (:1441)
    const float *dl = deltas + ((size_t)i * num_classes + 1) * 6;             // gather_nd(deltas, [i, class 1]) (:1464)
    const float *r = rois + (size_t)i * 6;
    float b[6], d[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        b[q] = __fmul_rn(__ldg(r + q), P.dim[q % 3]);                         // denorm_boxes_3d_graph: boxes * scale
        d[q] = __fmul_rn(__ldg(dl + q), P.std[q]);
    }
    float o[6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {                                             // y, x, z axes
        const float len = __fsub_rn(b[a + 3], b[a]);
        const float ctr = __fadd_rn(b[a], __fmul_rn(0.5f, len));
        const float ds = clip_rn(d[a + 3], -P.log_limit, P.log_limit);
        const float ctr2 = __fadd_rn(ctr, __fmul_rn(d[a], len));
        const float len2 = __fmul_rn(len, expf(ds));
        const float lo = __fsub_rn(ctr2, __fmul_rn(0.5f, len2));
        o[a] = clip_rn(lo, 0.0f, P.dim[a]);                                   // tf.clip_by_value(., 0, H)   (:1471-1476)
        o[a + 3] = clip_rn(__fadd_rn(lo, len2), 0.0f, P.dim[a]);
    }
    const bool ok = score >= P.min_conf &&                                    // (:1446)
                    __fsub_rn(o[3], o[0]) >= 1.0f && __fsub_rn(o[4], o[1]) >= 1.0f && __fsub_rn(o[5], o[2]) >= 0.5f;  // (:1480-1488)
    float *bo = boxes_px + (size_t)i * 6;
#pragma unroll
    for (int q = 0; q < 6; ++q) bo[q] = o[q];
    if (boxes_nms) {                                                          // reference mode: NMS sees the (y, x) projection
        float *bn = boxes_nms + (size_t)i * 6;
        bn[0] = o[0]; bn[1] = o[1]; bn[2] = 0.0f; bn[3] = o[3]; bn[4] = o[4]; bn[5] = 1.0f;
    }
    scores[i] = ok ? score : -FLT_MAX;                                        // not a candidate for the NMS op
}

__global__ void __launch_bounds__(256)
refine_gather_kernel(const float *__restrict__ boxes_px, const float *__restrict__ scores, const int *__restrict__ keep,
                     const int *__restrict__ count, int rois_per_image, int max_inst, int images, RefineParams P,
                     float *__restrict__ det, int *__restrict__ det_count)
{
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (det_count && t < images) det_count[t] = min(count[t], max_inst);
    if (t >= images * max_inst * 8) return;
    const int c = t & 7, j = (t >> 3) % max_inst, b = (t >> 3) / max_inst;
    float v = 0.0f;                                                           // tf.pad rows
    if (j < __ldg(count + b)) {
        const size_t src = (size_t)b * rois_per_image + __ldg(keep + (size_t)b * max_inst + j);
        if (c < 6) v = clip_rn(__fdiv_rn(__ldg(boxes_px + src * 6 + c), P.dim[c % 3]), 0.0f, 1.0f);   // norm_boxes_3d_graph
        else if (c == 6) v = 1.0f;                                            // class id (this fork: one fg class)
        else v = __ldg(scores + src);
    }
    det[t] = v;
}

struct RefineLayout { size_t boxes, boxes2d, scores, offs, keep, count, nms, total; };
static size_t up256(size_t v) { return (v + 255) & ~size_t(255); }
static RefineLayout refine_layout(int images, int rois, int max_inst) {
    RefineLayout L;
    size_t off = 0;
    const size_t total = (size_t)images * rois;
    L.boxes = off; off += up256(total * 6 * sizeof(float));
    L.boxes2d = off; off += up256(total * 6 * sizeof(float));
    L.scores = off; off += up256(total * sizeof(float));
    L.offs = off; off += up256(((size_t)images + 1) * sizeof(int));
    L.keep = off; off += up256((size_t)images * (max_inst > 0 ? max_inst : 1) * sizeof(int));
    L.count = off; off += up256((size_t)images * sizeof(int));
    L.nms = off; off += nms3d_workspace_bytes(rois, images);
    L.total = off;
    return L;
}

size_t refine_detections_workspace_bytes(int images, int rois, int max_inst) { return refine_layout(images, rois, max_inst).total; }

int launch_refine_detections(const float *rois, const float *probs, const float *deltas, int images, int rois_per_image,
                             int num_classes, const float image_shape[3], const float std_dev[6], float min_conf,
                             float nms_thr, int nms_mode, int max_inst, float *detections, int *det_count, void *ws,
                             size_t ws_bytes, cudaStream_t stream)
{
    const RefineLayout L = refine_layout(images, rois_per_image, max_inst);
    if (ws == nullptr || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return ROI3D_EWORKSPACE;
    char *base = static_cast<char *>(ws);
    float *boxes_px = reinterpret_cast<float *>(base + L.boxes);
    float *scores = reinterpret_cast<float *>(base + L.scores);
    const bool ref2d = nms_mode == ROI3D_NMS_REFERENCE_2D;
    float *boxes_nms = ref2d ? reinterpret_cast<float *>(base + L.boxes2d) : nullptr;
    // tf.image.non_max_suppression suppresses on iou > thr: the same compare as iou >= (thr + 1 ulp)
    const float thr_eff = ref2d ? nextafterf(nms_thr, INFINITY) : nms_thr;
    int *offs = reinterpret_cast<int *>(base + L.offs);
    int *keep = reinterpret_cast<int *>(base + L.keep);
    int *count = reinterpret_cast<int *>(base + L.count);
    RefineParams P;
    for (int q = 0; q < 3; ++q) P.dim[q] = image_shape[q];
    for (int q = 0; q < 6; ++q) P.std[q] = std_dev[q];
    P.min_conf = min_conf;
    P.log_limit = 4.1351666f;                                  // float32(log(62.5)) = 0x408452E3... nearest float
    const int total = images * rois_per_image;
    ROI3D_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)images, stream));
    if (total > 0) {
        ROI3D_CUDA_TRY(launch_dependent(refine_decode_kernel, dim3((max(total, images + 1) + 255) / 256), dim3(256), 0, stream, true, rois, probs,
                                        deltas, total, rois_per_image, images, num_classes, P, boxes_px, boxes_nms, scores, offs));
        ROI3D_LAUNCH_CHECK();
        const int rc = launch_nms3d(ref2d ? boxes_nms : boxes_px, scores, offs, images, rois_per_image, max_inst, thr_eff, keep, count,
                                    base + L.nms, ws_bytes - L.nms, stream);
        if (rc != ROI3D_OK) return rc;
    }
    ROI3D_CUDA_TRY(launch_dependent(refine_gather_kernel, dim3((images * max_inst * 8 + 255) / 256), dim3(256), 0, stream, true,
                                    (const float *)boxes_px, (const float *)scores, (const int *)keep, (const int *)count, rois_per_image,
                                    max_inst, images, P, detections, det_count));
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

// ---------------------------------------------------------------------------------
// mask targets: crop (C = 1) + round (+ packbits).  One thread per target voxel, z fastest, so a warp owns 32
// consecutive flat outputs = 4 packed bytes.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float axis_lerp_setup(float a1, float a2, int dim, int p, int k, int &i0, int &i1, bool &valid) {
    const float scale = axis_scale(a1, a2, dim, p);
    const float in = axis_coord(a1, a2, dim, p, k, scale);
    valid = !axis_invalid(in, dim);
    const float fl = floorf(in);
    i0 = (int)fl;
    i1 = (int)ceilf(in);
    return __fsub_rn(in, fl);
}

template <typename T> __device__ __forceinline__ float mask_ld(const T *p, size_t i);
template <> __device__ __forceinline__ float mask_ld<float>(const float *p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float mask_ld<unsigned char>(const unsigned char *p, size_t i) { return (float)__ldg(p + i); }

// np.packbits order: flat element 8q+r -> bit (7-r) of byte q.  `word` = ballot of 32 consecutive elements.
__device__ __forceinline__ void store_packed_word(unsigned char *bits, long long first, long long total, unsigned ballot) {
    const unsigned be = __brev(ballot);                          // element 0 -> bit 31: bytes in big-endian order
    const unsigned le = __byte_perm(be, 0, 0x0123);              // memory order: byte 0 = elements 0..7
    const long long nbytes = (total + 7) >> 3, q = first >> 3;
    if (q + 4 <= nbytes) {
        *reinterpret_cast<unsigned *>(bits + q) = le;
    } else {
        for (int k = 0; q + k < nbytes; ++k) bits[q + k] = (unsigned char)(le >> (8 * k));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
mask_targets_kernel(const T *__restrict__ masks, int H, int W, int D, const float *__restrict__ boxes,
                    const int *__restrict__ assignment, int n, int mh, int mw, int md,
                    float *__restrict__ targets, unsigned char *__restrict__ bits)
{
    const long long total = (long long)n * mh * mw * md;
    const long long padded = (total + 31) & ~31ll;               // whole warps stay converged for the ballot
    const size_t sW = (size_t)D, sH = (size_t)W * D;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < padded;
         idx += (long long)gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (idx < total) {
            long long r = idx;
            const int z = (int)(r % md); r /= md;
            const int x = (int)(r % mw); r /= mw;
            const int y = (int)(r % mh);
            const int b = (int)(r / mh);
            const float *box = boxes + (size_t)b * 6;
            int y0, y1, x0, x1, z0, z1;
            bool vy, vx, vz;
            const float yl = axis_lerp_setup(__ldg(box + 0), __ldg(box + 3), H, mh, y, y0, y1, vy);
            const float xl = axis_lerp_setup(__ldg(box + 1), __ldg(box + 4), W, mw, x, x0, x1, vx);
            const float zl = axis_lerp_setup(__ldg(box + 2), __ldg(box + 5), D, md, z, z0, z1, vz);
            if (vy && vx && vz) {                                 // else extrapolation_value 0 (the op's default)
                const T *m = masks + (size_t)(assignment ? __ldg(assignment + b) : b) * H * sH;
                const size_t t = y0 * sH, bo = y1 * sH, l = x0 * sW, rr = x1 * sW;
                const float tl = lerp_rn(mask_ld(m, t + l + z0), mask_ld(m, t + l + z1), zl);
                const float tr = lerp_rn(mask_ld(m, t + rr + z0), mask_ld(m, t + rr + z1), zl);
                const float bl = lerp_rn(mask_ld(m, bo + l + z0), mask_ld(m, bo + l + z1), zl);
                const float br = lerp_rn(mask_ld(m, bo + rr + z0), mask_ld(m, bo + rr + z1), zl);
                v = lerp_rn(lerp_rn(tl, tr, xl), lerp_rn(bl, br, xl), yl);
            }
            v = rintf(v);                                         // tf.round: half to even
            if (targets) __stcs(targets + idx, v);
        }
        if (bits) {
            const unsigned ballot = __ballot_sync(0xffffffffu, v > 0.5f);
            if ((threadIdx.x & 31) == 0) store_packed_word(bits, idx, total, ballot);
        }
    }
}

template <typename T>
static int launch_mask_targets_t(const T *masks, int H, int W, int D, const float *boxes, const int *assignment, int n,
                                 int mh, int mw, int md, float *targets, unsigned char *bits, cudaStream_t stream)
{
    const long long total = (long long)n * mh * mw * md;
    const long long blocks = min((total + 255) / 256, (long long)num_sms() * 16);
    mask_targets_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(masks, H, W, D, boxes, assignment, n, mh, mw, md, targets, bits);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_mask_targets(const void *masks, int mask_dtype, int H, int W, int D, const float *boxes, const int *assignment,
                        int n, int mh, int mw, int md, float *targets, unsigned char *bits, cudaStream_t stream)
{
    if (mask_dtype == 0)
        return launch_mask_targets_t(static_cast<const float *>(masks), H, W, D, boxes, assignment, n, mh, mw, md, targets, bits, stream);
    return launch_mask_targets_t(static_cast<const unsigned char *>(masks), H, W, D, boxes, assignment, n, mh, mw, md, targets, bits, stream);
}

// ---------------------------------------------------------------------------------
// wire format of the target files
// ---------------------------------------------------------------------------------
// numpy's float -> half: round to nearest even; a NaN keeps its sign and the top 10 payload bits (never becomes inf)
__device__ __forceinline__ unsigned short f2h_np(float x) {
    if (x != x) {
        const unsigned u = __float_as_uint(x);
        unsigned short r = (unsigned short)(0x7c00u + ((u & 0x007fffffu) >> 13));
        if (r == 0x7c00u) ++r;
        return (unsigned short)((u >> 16) & 0x8000u) | r;
    }
    return __half_as_ushort(__float2half_rn(x));
}
__device__ __forceinline__ unsigned f2h2_np(float a, float b) { return (unsigned)f2h_np(a) | ((unsigned)f2h_np(b) << 16); }

__global__ void __launch_bounds__(256)
f32_to_f16_kernel(const float *__restrict__ x, long long n, __half *__restrict__ y)
{
    const long long n8 = n >> 3;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const long long stride = (long long)gridDim.x * blockDim.x, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (vec) {
        for (long long i = tid; i < n8; i += stride) {
            const float4 a = ldg4(x + i * 8), b = ldg4(x + i * 8 + 4);
            uint4 o;
            o.x = f2h2_np(a.x, a.y); o.y = f2h2_np(a.z, a.w);
            o.z = f2h2_np(b.x, b.y); o.w = f2h2_np(b.z, b.w);
            __stcs(reinterpret_cast<uint4 *>(y) + i, o);
        }
    }
    for (long long i = (vec ? n8 * 8 : 0) + tid; i < n; i += stride) y[i] = __ushort_as_half(f2h_np(__ldg(x + i)));
}

// numpy's half -> float: exact; a NaN keeps its sign and payload (shifted), where cvt would canonicalise it
__device__ __forceinline__ float h2f_np(unsigned short h) {
    if ((h & 0x7c00u) == 0x7c00u && (h & 0x03ffu))
        return __uint_as_float(((unsigned)(h & 0x8000u) << 16) | 0x7f800000u | ((unsigned)(h & 0x03ffu) << 13));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ float2 h2f2_np(unsigned v) { return make_float2(h2f_np((unsigned short)(v & 0xffffu)), h2f_np((unsigned short)(v >> 16))); }

__global__ void __launch_bounds__(256)
f16_to_f32_kernel(const __half *__restrict__ x, long long n, float *__restrict__ y)
{
    const long long n8 = n >> 3;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const long long stride = (long long)gridDim.x * blockDim.x, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (vec) {
        for (long long i = tid; i < n8; i += stride) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(x) + i);
            const float2 a = h2f2_np(v.x), b = h2f2_np(v.y), c = h2f2_np(v.z), d = h2f2_np(v.w);
            st_stream4(y + i * 8, make_float4(a.x, a.y, b.x, b.y));
            st_stream4(y + i * 8 + 4, make_float4(c.x, c.y, d.x, d.y));
        }
    }
    for (long long i = (vec ? n8 * 8 : 0) + tid; i < n; i += stride) y[i] = h2f_np(__half_as_ushort(x[i]));
}

// numpy.packbits((x > 0.5).reshape(-1)): a warp packs 32 consecutive elements into 4 bytes per iteration
__global__ void __launch_bounds__(256)
pack_bits_kernel(const float *__restrict__ x, long long n, unsigned char *__restrict__ bits)
{
    const long long padded = (n + 31) & ~31ll;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < padded;
         idx += (long long)gridDim.x * blockDim.x) {
        const bool on = idx < n && __ldg(x + idx) > 0.5f;
        const unsigned ballot = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0) store_packed_word(bits, idx, n, ballot);
    }
}

// numpy.unpackbits(bits)[:n] as float32 0/1: one thread expands one byte into 8 floats
__global__ void __launch_bounds__(256)
unpack_bits_kernel(const unsigned char *__restrict__ bits, long long n, float *__restrict__ y)
{
    const long long nbytes = (n + 7) >> 3;
    const bool vec = (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nbytes; q += (long long)gridDim.x * blockDim.x) {
        const unsigned b = __ldg(bits + q);
        float v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = (float)((b >> (7 - r)) & 1u);
        if (vec && q * 8 + 8 <= n) {
            st_stream4(y + q * 8, make_float4(v[0], v[1], v[2], v[3]));
            st_stream4(y + q * 8 + 4, make_float4(v[4], v[5], v[6], v[7]));
        } else {
            for (int r = 0; r < 8 && q * 8 + r < n; ++r) y[q * 8 + r] = v[r];
        }
    }
}

static unsigned wire_grid(long long work_items) {
    const long long blocks = (work_items + 255) / 256;
    return (unsigned)max(1ll, min(blocks, (long long)num_sms() * 32));
}

int launch_f32_to_f16(const float *x, long long n, void *y, cudaStream_t stream)
{
    f32_to_f16_kernel<<<wire_grid((n + 7) / 8), 256, 0, stream>>>(x, n, static_cast<__half *>(y));
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}
int launch_f16_to_f32(const void *x, long long n, float *y, cudaStream_t stream)
{
    f16_to_f32_kernel<<<wire_grid((n + 7) / 8), 256, 0, stream>>>(static_cast<const __half *>(x), n, y);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}
int launch_pack_bits(const float *x, long long n, unsigned char *bits, cudaStream_t stream)
{
    pack_bits_kernel<<<wire_grid(n), 256, 0, stream>>>(x, n, bits);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}
int launch_unpack_bits(const unsigned char *bits, long long n, float *y, cudaStream_t stream)
{
    unpack_bits_kernel<<<wire_grid((n + 7) / 8), 256, 0, stream>>>(bits, n, y);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

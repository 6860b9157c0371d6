// roi3d_boxes.cu -- box-space helpers that sit either side of the hot ops.
//
//   overlaps3d        : pairwise IoU matrix of overlaps_graph (core/models.py:695-733), the first step of
//                       DetectionTargetLayer.  NOTE its arithmetic differs from the NMS op's IOU<float>: no corner
//                       ordering, no zero-volume early-out, union clamped with 1e-10.
//   decode_proposals  : ProposalLayer's per-anchor front-end between top_k and NMS (core/models.py:397-447):
//                       deltas * RPN_BBOX_STD_DEV, clip to +-3, apply_box_deltas_graph (:280-337), clip to [0,1]
//                       (clip_boxes_graph :340-364), min sizes (:435-447) -- one fused elementwise kernel, with an
//                       optional gather by top-k indices.
// Both are HBM-trivial elementwise / outer-product kernels: coalesced, one launch, no shared state.
#include "roi3d_common.cuh"

namespace roi3d {

constexpr int OV_TILE = 256;

__global__ void __launch_bounds__(256)
overlaps3d_kernel(const float *__restrict__ boxes1, int n, const float *__restrict__ boxes2, int m, float *__restrict__ out)
{
    __shared__ float s_b2[OV_TILE * 6];
    __shared__ float s_v2[OV_TILE];
    const int j0 = blockIdx.x * OV_TILE;
    const int mt = min(OV_TILE, m - j0);
    for (int t = threadIdx.x; t < mt * 6; t += blockDim.x) s_b2[t] = __ldg(boxes2 + (size_t)j0 * 6 + t);
    __syncthreads();
    for (int t = threadIdx.x; t < mt; t += blockDim.x) {
        const float *b = s_b2 + t * 6;
        s_v2[t] = __fmul_rn(__fmul_rn(__fsub_rn(b[3], b[0]), __fsub_rn(b[4], b[1])), __fsub_rn(b[5], b[2]));
    }
    __syncthreads();
    const int rows_per_cta = 32;
    const int i0 = blockIdx.y * rows_per_cta;
    for (int ii = 0; ii < rows_per_cta; ++ii) {
        const int i = i0 + ii;
        if (i >= n) break;
        const float *a = boxes1 + (size_t)i * 6;
        const float a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2), a3 = __ldg(a + 3), a4 = __ldg(a + 4), a5 = __ldg(a + 5);
        const float v1 = __fmul_rn(__fmul_rn(__fsub_rn(a3, a0), __fsub_rn(a4, a1)), __fsub_rn(a5, a2));
        for (int t = threadIdx.x; t < mt; t += blockDim.x) {
            const float *b = s_b2 + t * 6;
            const float y1 = fmaxf(a0, b[0]), x1 = fmaxf(a1, b[1]), z1 = fmaxf(a2, b[2]);
            const float y2 = fminf(a3, b[3]), x2 = fminf(a4, b[4]), z2 = fminf(a5, b[5]);
            const float inter = __fmul_rn(__fmul_rn(fmaxf(__fsub_rn(y2, y1), 0.f), fmaxf(__fsub_rn(x2, x1), 0.f)),
                                          fmaxf(__fsub_rn(z2, z1), 0.f));
            const float uni = __fsub_rn(__fadd_rn(v1, s_v2[t]), inter);
            out[(size_t)i * m + j0 + t] = __fdiv_rn(inter, fmaxf(uni, 1e-10f));
        }
    }
}

struct Std6 { float v[6]; };

__global__ void __launch_bounds__(256)
decode_proposals_kernel(const float *__restrict__ anchors, const float *__restrict__ deltas, const int *__restrict__ index,
                        int n, Std6 std, float min_dz, float *__restrict__ boxes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t src = index ? (size_t)__ldg(index + i) : (size_t)i;
    const float *a = anchors + src * 6, *dl = deltas + src * 6;
    float d[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) d[q] = fminf(fmaxf(__fmul_rn(__ldg(dl + q), std.v[q]), -3.0f), 3.0f);
    const float a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2), a3 = __ldg(a + 3), a4 = __ldg(a + 4), a5 = __ldg(a + 5);
    float h = __fsub_rn(a3, a0), w = __fsub_rn(a4, a1), dp = __fsub_rn(a5, a2);
    float cy = __fadd_rn(a0, __fmul_rn(0.5f, h)), cx = __fadd_rn(a1, __fmul_rn(0.5f, w)), cz = __fadd_rn(a2, __fmul_rn(0.5f, dp));
    cy = __fadd_rn(cy, __fmul_rn(d[0], h));
    cx = __fadd_rn(cx, __fmul_rn(d[1], w));
    cz = __fadd_rn(cz, __fmul_rn(d[2], dp));
    h = __fmul_rn(h, expf(d[3]));
    w = __fmul_rn(w, expf(d[4]));
    dp = __fmul_rn(dp, expf(d[5]));
    float y1 = __fsub_rn(cy, __fmul_rn(0.5f, h)), x1 = __fsub_rn(cx, __fmul_rn(0.5f, w)), z1 = __fsub_rn(cz, __fmul_rn(0.5f, dp));
    float y2 = __fadd_rn(y1, h), x2 = __fadd_rn(x1, w), z2 = __fadd_rn(z1, dp);
    // tf.clip_by_value(result, 0, 1) and clip_boxes_graph(window = [0,0,0,1,1,1]) are the same clamp
    y1 = fmaxf(fminf(y1, 1.f), 0.f); x1 = fmaxf(fminf(x1, 1.f), 0.f); z1 = fmaxf(fminf(z1, 1.f), 0.f);
    y2 = fmaxf(fminf(y2, 1.f), 0.f); x2 = fmaxf(fminf(x2, 1.f), 0.f); z2 = fmaxf(fminf(z2, 1.f), 0.f);
    y2 = fmaxf(y2, __fadd_rn(y1, 1e-6f));
    x2 = fmaxf(x2, __fadd_rn(x1, 1e-6f));
    z2 = fmaxf(z2, __fadd_rn(z1, min_dz));
    float *o = boxes + (size_t)i * 6;
    o[0] = y1; o[1] = x1; o[2] = z1; o[3] = y2; o[4] = x2; o[5] = z2;
}

int launch_overlaps3d(const float *boxes1, int n, const float *boxes2, int m, float *out, cudaStream_t stream)
{
    dim3 grid((m + OV_TILE - 1) / OV_TILE, (n + 31) / 32);
    overlaps3d_kernel<<<grid, 256, 0, stream>>>(boxes1, n, boxes2, m, out);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_decode_proposals(const float *anchors, const float *deltas, const int *index, int n, const float std_dev[6],
                            float image_depth, float *boxes, cudaStream_t stream)
{
    Std6 s;
    for (int q = 0; q < 6; ++q) s.v[q] = std_dev[q];
    const float depth = image_depth > 1.0f ? image_depth : 1.0f;
    const float inv = 1.0f / depth;
    const float min_dz = inv > 1e-4f ? inv : 1e-4f;
    decode_proposals_kernel<<<(n + 255) / 256, 256, 0, stream>>>(anchors, deltas, index, n, s, min_dz, boxes);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

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
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
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
    ROI3D_CUDA_TRY(launch_dependent(decode_proposals_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, true, anchors, deltas, index, n, s, min_dz, boxes));
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

// =====================================================================================================
// top-k selection (tf.nn.top_k(scores, k), core/models.py:403-404) as a 3-pass radix select on the 32-bit
// descending-order key, followed by an index-ordered compaction.  The selected SET is exactly TF's (ties at the
// threshold broken towards lower indices); it is emitted in ascending index order, which is all the NMS that follows
// needs (it sorts by (score, position) itself, so position ties resolve like top_k's "lower index first").
// =====================================================================================================
namespace roi3d {

__device__ __forceinline__ unsigned topk_key(float s) {
    // ascending unsigned key <=> descending float; -0 == +0; NaN sorts last
    if (s != s) return 0xFFFFFFFFu;
    if (s == 0.0f) s = 0.0f;
    const unsigned b = __float_as_uint(s);
    const unsigned asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~asc;
}

struct TopkState {                  // lives in the workspace
    unsigned prefix;                // key bits decided so far (high bits)
    unsigned need;                  // how many elements still to take among the current prefix class
    unsigned done_ctr;              // CTAs of the current kernel that have published their part (last one finishes the job)
    unsigned hist[2048];
};

constexpr int TK_THREADS = 1024;

// pass p in {0,1,2}: bits [31:21], [20:10], [9:0]
__device__ __forceinline__ int tk_shift(int p) { return p == 0 ? 21 : (p == 1 ? 10 : 0); }
__device__ __forceinline__ int tk_bits(int p) { return p == 2 ? 10 : 11; }

// the CTA that publishes last finishes the pass: find the digit where the running count (best keys first) reaches
// `need`, record it in the prefix and clear the histogram for the next pass
__device__ __forceinline__ void tk_pick(int pass, int k_total, TopkState *__restrict__ st, unsigned *s_cnt, unsigned *s_sum)
{
    const int nb = tk_bits(pass), bins = 1 << nb, shift = tk_shift(pass);
    for (int t = threadIdx.x; t < 2048; t += blockDim.x) { s_cnt[t] = __ldcg(&st->hist[t]); st->hist[t] = 0; }
    __syncthreads();
    // inclusive scan of 2 bins per thread
    const unsigned a = (2 * threadIdx.x < bins) ? s_cnt[2 * threadIdx.x] : 0u;
    const unsigned b = (2 * threadIdx.x + 1 < bins) ? s_cnt[2 * threadIdx.x + 1] : 0u;
    s_sum[threadIdx.x] = a + b;
    __syncthreads();
    for (int o = 1; o < TK_THREADS; o <<= 1) {
        const unsigned v = (threadIdx.x >= o) ? s_sum[threadIdx.x - o] : 0u;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned need = (pass == 0) ? (unsigned)k_total : st->need;
    const unsigned prefix = (pass == 0) ? 0u : st->prefix;
    const unsigned before = s_sum[threadIdx.x] - (a + b);        // count of strictly better digits
    __syncthreads();
    // the digit d with before(d) < need <= before(d) + cnt(d)
    if (a && before < need && need <= before + a) {
        st->prefix = prefix | ((unsigned)(2 * threadIdx.x) << shift);
        st->need = need - before;
    } else if (b && before + a < need && need <= before + a + b) {
        st->prefix = prefix | ((unsigned)(2 * threadIdx.x + 1) << shift);
        st->need = need - (before + a);
    }
    if (threadIdx.x == 0) st->done_ctr = 0;
}

__device__ __forceinline__ bool tk_last_cta(TopkState *__restrict__ st)
{
    __shared__ bool s_last;
    __threadfence();                                             // this CTA's results are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->done_ctr, 1u) == gridDim.x - 1;   // ... before it takes its ticket
    __syncthreads();
    return s_last;
}

__global__ void __launch_bounds__(TK_THREADS)
topk_hist_kernel(const float *__restrict__ scores, int n, int pass, int k_total, TopkState *__restrict__ st)
{
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
    __shared__ unsigned s_hist[2048];
    __shared__ unsigned s_sum[TK_THREADS];
    for (int t = threadIdx.x; t < 2048; t += blockDim.x) s_hist[t] = 0;
    __syncthreads();
    const int shift = tk_shift(pass), nb = tk_bits(pass);
    const unsigned prefix = st->prefix;
    const int pshift = shift + nb;                               // bits above the current digit
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned k = topk_key(__ldg(scores + i));
        if (pass == 0 || (k >> pshift) == (prefix >> pshift))
            atomicAdd(&s_hist[(k >> shift) & ((1u << nb) - 1u)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2048; t += blockDim.x)
        if (s_hist[t]) atomicAdd(&st->hist[t], s_hist[t]);
    if (tk_last_cta(st)) tk_pick(pass, k_total, st, s_hist, s_sum);
}

// per-range counts of keys strictly better than / equal to the threshold key
__global__ void __launch_bounds__(TK_THREADS)
topk_count_kernel(const float *__restrict__ scores, int n, int range, int nblocks, TopkState *__restrict__ st,
                  unsigned *__restrict__ less_cnt, unsigned *__restrict__ eq_cnt, unsigned *__restrict__ eq_off,
                  unsigned *__restrict__ sel_off)
{
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
    __shared__ unsigned s_less, s_eq;
    __shared__ unsigned s_a[TK_THREADS];
    if (threadIdx.x == 0) { s_less = 0; s_eq = 0; }
    __syncthreads();
    const unsigned T = st->prefix;
    const int i0 = blockIdx.x * range, i1 = min(n, i0 + range);
    unsigned l = 0, e = 0;
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const unsigned k = topk_key(__ldg(scores + i));
        l += k < T;
        e += k == T;
    }
    for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
    if ((threadIdx.x & 31) == 0) { if (l) atomicAdd(&s_less, l); if (e) atomicAdd(&s_eq, e); }
    __syncthreads();
    if (threadIdx.x == 0) { less_cnt[blockIdx.x] = s_less; eq_cnt[blockIdx.x] = s_eq; }
    if (!tk_last_cta(st)) return;
    // last CTA: exclusive scans over the ranges -> eq_off[b] (rank of the range's first equal key) and sel_off[b]
    {
        const unsigned need_eq = st->need;
        const int t = threadIdx.x;
        const unsigned ee = t < nblocks ? __ldcg(eq_cnt + t) : 0u, ll = t < nblocks ? __ldcg(less_cnt + t) : 0u;
        s_a[t] = ee;
        __syncthreads();
        for (int o = 1; o < TK_THREADS; o <<= 1) {
            const unsigned v = (t >= o) ? s_a[t - o] : 0u;
            __syncthreads();
            s_a[t] += v;
            __syncthreads();
        }
        const unsigned eoff = s_a[t] - ee;
        const unsigned take = eoff >= need_eq ? 0u : min(ee, need_eq - eoff);
        __syncthreads();
        s_a[t] = ll + take;
        __syncthreads();
        for (int o = 1; o < TK_THREADS; o <<= 1) {
            const unsigned v = (t >= o) ? s_a[t - o] : 0u;
            __syncthreads();
            s_a[t] += v;
            __syncthreads();
        }
        if (t < nblocks) { eq_off[t] = eoff; sel_off[t] = s_a[t] - (ll + take); }
        if (t == 0) st->done_ctr = 0;
    }
}

// index-ordered compaction of the selected set
__global__ void __launch_bounds__(TK_THREADS)
topk_compact_kernel(const float *__restrict__ scores, int n, int range, const TopkState *__restrict__ st,
                    const unsigned *__restrict__ eq_off, const unsigned *__restrict__ sel_off,
                    int *__restrict__ idx_out, float *__restrict__ scores_out)
{
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
    __shared__ unsigned s_warp[2][32];
    __shared__ unsigned s_base[2];
    const unsigned T = st->prefix, need_eq = st->need;
    const int i0 = blockIdx.x * range, i1 = min(n, i0 + range);
    if (threadIdx.x == 0) { s_base[0] = eq_off[blockIdx.x]; s_base[1] = sel_off[blockIdx.x]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = i0; c0 < i1; c0 += TK_THREADS) {
        const int i = c0 + threadIdx.x;
        float s = 0.f;
        unsigned k = 0xFFFFFFFFu;
        const bool in = i < i1;
        if (in) { s = __ldg(scores + i); k = topk_key(s); }
        const bool eq = in && k == T, less = in && k < T;
        // rank among the equal keys of this chunk
        const unsigned eb = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp[0][warp] = __popc(eb);
        __syncthreads();
        unsigned ebefore = 0;
        for (int w = 0; w < warp; ++w) ebefore += s_warp[0][w];
        const unsigned erank = s_base[0] + ebefore + __popc(eb & ((1u << lane) - 1u));
        const bool sel = less || (eq && erank < need_eq);
        const unsigned sb = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) s_warp[1][warp] = __popc(sb);
        __syncthreads();
        unsigned sbefore = 0, etotal = 0, stotal = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < warp) sbefore += s_warp[1][w];
            etotal += s_warp[0][w];
            stotal += s_warp[1][w];
        }
        if (sel) {
            const unsigned pos = s_base[1] + sbefore + __popc(sb & ((1u << lane) - 1u));
            idx_out[pos] = i;
            if (scores_out) scores_out[pos] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_base[0] += etotal; s_base[1] += stotal; }
        __syncthreads();
    }
}

// proposals[P,6] = boxes[keep[0..count)] zero-padded (ProposalLayer, core/models.py:476-484); count read on device
__global__ void __launch_bounds__(256)
gather_pad_boxes_kernel(const float *__restrict__ boxes, const int *__restrict__ keep, const int *__restrict__ count,
                        int p, float *__restrict__ out)
{
    pdl_wait();                                                // see roi3d_common.cuh: programmatic dependent launch
    pdl_trigger();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p * 6) return;
    const int r = t / 6, c = t - r * 6;
    out[t] = (r < *count) ? __ldg(boxes + (size_t)__ldg(keep + r) * 6 + c) : 0.0f;
}

static int topk_range(int n) {
    int range = 4096;
    while ((long long)range * TK_THREADS < n) range *= 2;       // at most 1024 ranges (single-CTA scan)
    return range;
}

size_t topk_workspace_bytes(int n) {
    const int nblocks = (n + topk_range(n) - 1) / topk_range(n);
    return 256 * ((sizeof(TopkState) + 255) / 256) + 4 * 256 * ((sizeof(unsigned) * (size_t)(nblocks > 0 ? nblocks : 1) + 255) / 256);
}

int launch_topk(const float *scores, int n, int k, int *idx_out, float *scores_out, void *ws, size_t ws_bytes,
                cudaStream_t stream)
{
    if (ws == nullptr || ws_bytes < topk_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(ws) & 255)) return ROI3D_EWORKSPACE;
    const int range = topk_range(n), nblocks = (n + range - 1) / range;
    char *base = static_cast<char *>(ws);
    TopkState *st = reinterpret_cast<TopkState *>(base);
    const size_t st_bytes = 256 * ((sizeof(TopkState) + 255) / 256);
    const size_t arr = 256 * ((sizeof(unsigned) * (size_t)nblocks + 255) / 256);
    unsigned *less_cnt = reinterpret_cast<unsigned *>(base + st_bytes);
    unsigned *eq_cnt = reinterpret_cast<unsigned *>(base + st_bytes + arr);
    unsigned *eq_off = reinterpret_cast<unsigned *>(base + st_bytes + 2 * arr);
    unsigned *sel_off = reinterpret_cast<unsigned *>(base + st_bytes + 3 * arr);
    ROI3D_CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(TopkState), stream));
    const int hgrid = min((n + TK_THREADS - 1) / TK_THREADS, num_sms() * 2);
    for (int pass = 0; pass < 3; ++pass) {
        ROI3D_CUDA_TRY(launch_dependent(topk_hist_kernel, dim3(hgrid), dim3(TK_THREADS), 0, stream, true, scores, n, pass, k, st));
        ROI3D_LAUNCH_CHECK();
    }
    ROI3D_CUDA_TRY(launch_dependent(topk_count_kernel, dim3(nblocks), dim3(TK_THREADS), 0, stream, true, scores, n, range, nblocks, st,
                                    less_cnt, eq_cnt, eq_off, sel_off));
    ROI3D_LAUNCH_CHECK();
    ROI3D_CUDA_TRY(launch_dependent(topk_compact_kernel, dim3(nblocks), dim3(TK_THREADS), 0, stream, true, scores, n, range, (const TopkState *)st, (const unsigned *)eq_off, (const unsigned *)sel_off, idx_out, scores_out));
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_gather_pad_boxes(const float *boxes, const int *keep, const int *count, int p, float *out, cudaStream_t stream)
{
    ROI3D_CUDA_TRY(launch_dependent(gather_pad_boxes_kernel, dim3((p * 6 + 255) / 256), dim3(256), 0, stream, true, boxes, keep, count, p, out));
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

// roi3d_nms.cu -- NonMaxSuppression3D on sm_100a.
//
// Reference behaviour: NonMaxSuppression3DOp<CPUDevice>::Compute
// (NMS.so@0xe4e0) -> DoNonMaxSuppressionOp<float> (NMS.so@0xd0c0, TF r2.2 priority-queue
// greedy NMS, hard suppression on `iou >= thr`, ties -> lower index first, re-push quirk)
// with IOU<float> (NMS.so@0xb500).  For inputs whose selected boxes all have volume > 0
// that algorithm equals: stable sort by (score desc, index asc), keep a box iff its IoU
// with every earlier KEPT box is < thr, stop after max_out keeps.  A selected box with
// volume <= 0 has self-IoU 0 and is therefore re-selected until max_out is reached
// (row a2) -- the scan kernel reproduces that.
//
// Three launches, all on the caller's stream, no host sync:
//   1. nms_rank_sort_kernel : device sort.  Each score becomes a 32-bit descending-order
//      key; a box's rank is the number of boxes that precede it in (key, index) order,
//      counted against shared-memory tiles of keys.  One kernel = stable sort + gather:
//      it writes the sorted original indices and the sorted boxes in canonical
//      (min corner, max corner, volume) form.  O(n^2) compares at 2 instr/pair beat a
//      multi-pass radix sort up to ~10^5 boxes because there is a single launch.
//   2. nms_mask_kernel      : tiled pairwise 3-D IoU -> suppression bitmask
//      mask[i][w] bit b = IoU(sorted i, sorted 32w+b) >= thr.  Row boxes staged in shared
//      memory, one __ballot_sync per 32 pairs builds a mask word.  IoU is evaluated in the
//      reference's fp32 operation order with IEEE division.
//   3. nms_scan_kernel      : single-CTA greedy scan over 32-box chunks; warp 0 resolves a
//      chunk serially from its diagonal word, all warps OR the kept rows into the
//      shared-memory `removed` bitmap.  Emits original indices in selection order.
#include "roi3d_common.cuh"
#include <float.h>

namespace roi3d {

struct __align__(16) SBox {          // canonical sorted box, 32 bytes
    float ymin, xmin, zmin, ymax;
    float xmax, zmax, vol, pad;
};

__device__ __forceinline__ unsigned score_key(float s) {
    // ascending unsigned key <=> descending score; invalid candidates -> 0xFFFFFFFF.
    // Candidate rule of the reference: score > -FLT_MAX (NaN fails it).  -0.0 == +0.0.
    if (!(s > -FLT_MAX)) return 0xFFFFFFFFu;
    if (s == 0.0f) s = 0.0f;                                   // canonicalise -0
    const unsigned b = __float_as_uint(s);
    const unsigned asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~asc;
}

// ---------------------------------------------------------------------------------
// 1. rank sort + gather
// ---------------------------------------------------------------------------------
constexpr int RS_ITILE = 32;        // boxes ranked per CTA
constexpr int RS_JSPLIT = 32;       // threads sharing one box's j-range
constexpr int RS_THREADS = RS_ITILE * RS_JSPLIT;
constexpr int RS_KTILE = 4096;      // keys staged per shared-memory tile

__global__ void __launch_bounds__(RS_THREADS)
nms_rank_sort_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, int n,
                     int *__restrict__ sorted_idx, SBox *__restrict__ sboxes, int *__restrict__ nvalid_out)
{
    __shared__ __align__(16) unsigned s_keys[RS_KTILE];
    __shared__ int s_rank[RS_ITILE];
    __shared__ int s_valid;
    const int il = threadIdx.x / RS_JSPLIT;                   // box within the CTA's tile
    const int jq = threadIdx.x % RS_JSPLIT;                   // lane over the key tile
    const int i = blockIdx.x * RS_ITILE + il;
    const unsigned ki = (i < n) ? score_key(__ldg(scores + i)) : 0xFFFFFFFFu;
    if (threadIdx.x < RS_ITILE) s_rank[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_valid = 0;
    int cnt = 0, nvalid = 0;
    for (int j0 = 0; j0 < n; j0 += RS_KTILE) {
        __syncthreads();
        const int tile = min(RS_KTILE, n - j0);
        for (int t = threadIdx.x; t < RS_KTILE; t += RS_THREADS) {
            const unsigned k = (t < tile) ? score_key(__ldg(scores + j0 + t)) : 0xFFFFFFFFu;
            s_keys[t] = k;
            if (blockIdx.x == 0 && t < tile && k != 0xFFFFFFFFu) ++nvalid;
        }
        __syncthreads();
        // a box j precedes i iff key_j < key_i, or key_j == key_i and j < i
        for (int t = jq * 4; t < tile; t += RS_JSPLIT * 4) {
            const uint4 k4 = *reinterpret_cast<const uint4 *>(&s_keys[t]);
            const int j = j0 + t;
            cnt += (k4.x < ki) || (k4.x == ki && j + 0 < i);
            cnt += (t + 1 < tile) && ((k4.y < ki) || (k4.y == ki && j + 1 < i));
            cnt += (t + 2 < tile) && ((k4.z < ki) || (k4.z == ki && j + 2 < i));
            cnt += (t + 3 < tile) && ((k4.w < ki) || (k4.w == ki && j + 3 < i));
        }
    }
    // reduce the RS_JSPLIT partial counts of each box (RS_JSPLIT == warp size)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (blockIdx.x == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        if (jq == 0 && nvalid) atomicAdd(&s_valid, nvalid);
    }
    if (jq == 0) s_rank[il] = cnt;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *nvalid_out = s_valid;
    if (threadIdx.x < RS_ITILE) {
        const int ii = blockIdx.x * RS_ITILE + threadIdx.x;
        if (ii < n) {
            const int r = s_rank[threadIdx.x];
            sorted_idx[r] = ii;
            const float *b = boxes + (size_t)ii * 6;
            const float b0 = __ldg(b + 0), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
            const float b3 = __ldg(b + 3), b4 = __ldg(b + 4), b5 = __ldg(b + 5);
            SBox s;
            s.ymin = fminf(b0, b3); s.ymax = fmaxf(b0, b3);
            s.xmin = fminf(b1, b4); s.xmax = fmaxf(b1, b4);
            s.zmin = fminf(b2, b5); s.zmax = fmaxf(b2, b5);
            s.vol = __fmul_rn(__fmul_rn(__fsub_rn(s.ymax, s.ymin), __fsub_rn(s.xmax, s.xmin)),
                              __fsub_rn(s.zmax, s.zmin));
            s.pad = 0.f;
            sboxes[r] = s;
        }
    }
}

// ---------------------------------------------------------------------------------
// 2. pairwise IoU bitmask
// ---------------------------------------------------------------------------------
constexpr int MK_ROWS = 64;         // row boxes staged per CTA
constexpr int MK_WARPS = 8;         // mask words (32 columns each) per CTA

// IOU<float> (NMS.so@0xb500) on canonical boxes: same operations, same order.
__device__ __forceinline__ bool iou_ge(const SBox &a, const float4 b0, const float4 b1, float thr) {
    // b0 = (ymin,xmin,zmin,ymax)  b1 = (xmax,zmax,vol,-)
    if (a.vol <= 0.0f || b1.z <= 0.0f) return 0.0f >= thr;
    const float dy = fmaxf(0.0f, __fsub_rn(fminf(a.ymax, b0.w), fmaxf(a.ymin, b0.x)));
    const float dx = fmaxf(0.0f, __fsub_rn(fminf(a.xmax, b1.x), fmaxf(a.xmin, b0.y)));
    const float dz = fmaxf(0.0f, __fsub_rn(fminf(a.zmax, b1.y), fmaxf(a.zmin, b0.z)));
    const float inter = __fmul_rn(__fmul_rn(dy, dx), dz);
    if (!(inter > 0.0f)) return 0.0f >= thr;                  // 0 / positive == 0 exactly
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.vol, b1.z), inter));
    return iou >= thr;
}

__global__ void __launch_bounds__(MK_WARPS * 32)
nms_mask_kernel(const SBox *__restrict__ sboxes, int n, int pitch_words, float thr,
                unsigned *__restrict__ mask)
{
    __shared__ SBox s_rows[MK_ROWS];
    const int i0 = blockIdx.y * MK_ROWS;
    const int w0 = blockIdx.x * MK_WARPS;
    // only words that contain some column j > i0 are ever read by the scan
    if ((w0 + MK_WARPS) * 32 <= i0) return;
    const int rows = min(MK_ROWS, n - i0);
    for (int t = threadIdx.x; t < rows * 2; t += blockDim.x)
        reinterpret_cast<float4 *>(s_rows)[t] = __ldg(reinterpret_cast<const float4 *>(sboxes + i0) + t);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = w0 + warp;
    if (w * 32 >= n || (w + 1) * 32 <= i0) return;
    const int j = w * 32 + lane;
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n) {
        b0 = __ldg(reinterpret_cast<const float4 *>(sboxes + j));
        b1 = __ldg(reinterpret_cast<const float4 *>(sboxes + j) + 1);
    }
    unsigned word0 = 0, word1 = 0;
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
        const bool bit = (j < n) && iou_ge(s_rows[r], b0, b1, thr);
        const unsigned wd = __ballot_sync(0xffffffffu, bit);
        if (r < 32) { if (lane == r) word0 = wd; }
        else        { if (lane == r - 32) word1 = wd; }
    }
    if (lane < rows) mask[(size_t)(i0 + lane) * pitch_words + w] = word0;
    if (lane + 32 < rows) mask[(size_t)(i0 + 32 + lane) * pitch_words + w] = word1;
}

// ---------------------------------------------------------------------------------
// 3. greedy scan (single CTA)
// ---------------------------------------------------------------------------------
constexpr int SC_THREADS = 1024;

__global__ void __launch_bounds__(SC_THREADS)
nms_scan_kernel(const unsigned *__restrict__ mask, int pitch_words, const SBox *__restrict__ sboxes,
                const int *__restrict__ sorted_idx, const int *__restrict__ nvalid_p,
                int max_out, float thr, int *__restrict__ keep_idx, int *__restrict__ keep_count)
{
    extern __shared__ unsigned s_removed[];                    // pitch_words words
    __shared__ unsigned s_kept;
    __shared__ int s_nsel;
    __shared__ int s_fill;                                     // >= 0: zero-volume quirk, repeat this index
    const int nvalid = *nvalid_p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    for (int t = threadIdx.x; t < pitch_words; t += blockDim.x) s_removed[t] = 0u;
    if (threadIdx.x == 0) { s_nsel = 0; s_fill = -1; s_kept = 0u; }
    const int nchunks = (nvalid + 31) >> 5;
    int nsel = 0;
    for (int c = 0; c < nchunks && nsel < max_out; ++c) {
        __syncthreads();                                       // removed[c] final, s_nsel visible
        if (warp == 0) {
            unsigned rw = s_removed[c];
            const int row = c * 32 + lane;
            unsigned d = 0u;
            bool degenerate = false;
            if (row < nvalid) {
                d = __ldg(mask + (size_t)row * pitch_words + c);
                degenerate = !(sboxes[row].vol > 0.0f) && !(0.0f >= thr);   // self-IoU 0 < thr
            } else {
                rw |= (1u << lane);
            }
            rw = __reduce_or_sync(0xffffffffu, rw);
            const unsigned degen = __ballot_sync(0xffffffffu, degenerate);
            unsigned kept = 0u;
            int fill_row = -1, nk = 0;
            for (int b = 0; b < 32; ++b) {
                const unsigned db = __shfl_sync(0xffffffffu, d, b);
                if (!((rw >> b) & 1u)) {
                    kept |= (1u << b);
                    rw |= db;
                    ++nk;
                    if ((degen >> b) & 1u) { fill_row = c * 32 + b; break; }
                    if (nsel + nk >= max_out) break;
                }
            }
            if ((kept >> lane) & 1u)
                keep_idx[nsel + __popc(kept & ((1u << lane) - 1u))] = __ldg(sorted_idx + row);
            if (lane == 0) {
                s_kept = kept;
                s_nsel = nsel + nk;
                s_fill = (fill_row >= 0) ? __ldg(sorted_idx + fill_row) : -1;
            }
        }
        __syncthreads();
        const unsigned kept = s_kept;
        nsel = s_nsel;
        if (s_fill >= 0) break;
        // OR the kept rows into removed[w] for w > c (one warp per kept row)
        unsigned kk = kept;
        int q = 0;
        while (kk) {
            const int b = __ffs(kk) - 1;
            kk &= kk - 1;
            if ((q++ % nwarps) == warp) {
                const unsigned *mrow = mask + (size_t)(c * 32 + b) * pitch_words;
                for (int w = c + 1 + lane; w < nchunks; w += 32) {
                    const unsigned m = __ldg(mrow + w);
                    if (m) atomicOr(&s_removed[w], m);
                }
            }
        }
    }
    __syncthreads();
    const int fill = s_fill;
    nsel = s_nsel;
    if (fill >= 0) {                                           // zero-volume quirk: repeat until max_out
        for (int t = nsel + threadIdx.x; t < max_out; t += blockDim.x) keep_idx[t] = fill;
        nsel = max_out;
    }
    if (threadIdx.x == 0) *keep_count = nsel;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct NmsLayout {
    int pitch_words;
    size_t off_sidx, off_sboxes, off_nvalid, off_mask, total;
};

static NmsLayout nms_layout(int n) {
    NmsLayout L;
    const int words = (n + 31) / 32;
    L.pitch_words = (int)align_up((size_t)(words > 0 ? words : 1), 32);   // 128-byte rows
    size_t off = 0;
    L.off_nvalid = off; off += 256;
    L.off_sidx = off;   off += align_up(sizeof(int) * (size_t)n, 256);
    L.off_sboxes = off; off += align_up(sizeof(SBox) * (size_t)n, 256);
    L.off_mask = off;   off += align_up(sizeof(unsigned) * (size_t)n * L.pitch_words, 256);
    L.total = off;
    return L;
}

size_t nms3d_workspace_bytes(int n) { return n <= 0 ? 256 : nms_layout(n).total; }

int launch_nms3d(const float *boxes, const float *scores, int n, int max_out, float thr,
                 int *keep_idx, int *keep_count, void *ws, size_t ws_bytes, cudaStream_t stream)
{
    if (n <= 0 || max_out <= 0) {
        ROI3D_CUDA_TRY(cudaMemsetAsync(keep_count, 0, sizeof(int), stream));
        return ROI3D_OK;
    }
    const NmsLayout L = nms_layout(n);
    if (ws == nullptr || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return ROI3D_EWORKSPACE;
    if ((size_t)L.pitch_words * sizeof(unsigned) > 200 * 1024) return ROI3D_EUNSUPPORTED;   // > 1.6 M boxes
    char *base = static_cast<char *>(ws);
    int *nvalid = reinterpret_cast<int *>(base + L.off_nvalid);
    int *sidx = reinterpret_cast<int *>(base + L.off_sidx);
    SBox *sboxes = reinterpret_cast<SBox *>(base + L.off_sboxes);
    unsigned *mask = reinterpret_cast<unsigned *>(base + L.off_mask);

    nms_rank_sort_kernel<<<(n + RS_ITILE - 1) / RS_ITILE, RS_THREADS, 0, stream>>>(boxes, scores, n, sidx, sboxes, nvalid);
    ROI3D_LAUNCH_CHECK();
    const int words = (n + 31) / 32;
    dim3 mgrid((words + MK_WARPS - 1) / MK_WARPS, (n + MK_ROWS - 1) / MK_ROWS);
    nms_mask_kernel<<<mgrid, MK_WARPS * 32, 0, stream>>>(sboxes, n, L.pitch_words, thr, mask);
    ROI3D_LAUNCH_CHECK();
    const size_t smem = (size_t)L.pitch_words * sizeof(unsigned);
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_scan_kernel<<<1, SC_THREADS, smem, stream>>>(mask, L.pitch_words, sboxes, sidx, nvalid, max_out, thr,
                                                    keep_idx, keep_count);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

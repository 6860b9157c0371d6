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
//   3. nms_scan_kernel      : single-CTA warp-level greedy scan.  Warp 0 resolves 32-box chunks
//      with ballot rounds out of a shared-memory copy of the diagonal mask block that the
//      other warps prefetch one super-chunk ahead; all warps OR the kept rows into the
//      shared-memory `removed` bitmap at super-chunk boundaries.  Emits original indices in
//      selection order.
#include "roi3d_common.cuh"
#include <float.h>

namespace roi3d {

struct __align__(16) SBox {          // canonical sorted box, 32 bytes
    float ymin, xmin, zmin, ymax;
    float xmax, zmax, vol, pad;
};

// One launch can run many independent NMS problems ("segments": the images of a batch in ProposalLayer's
// batch_slice, core/models.py:487-490, or the classes of DetectionLayer's per-class NMS).  Segment z covers the input
// boxes [offsets[z], offsets[z+1]) and owns workspace slices of `stride` boxes; offsets == nullptr means a single
// problem of n boxes.  Indices are local to the segment.
struct NmsSeg {
    const int *offsets;
    int n;
    int stride;
};
__device__ __forceinline__ int seg_begin(const NmsSeg &g, int z) { return g.offsets ? __ldg(g.offsets + z) : 0; }
__device__ __forceinline__ int seg_size(const NmsSeg &g, int z) {
    return g.offsets ? __ldg(g.offsets + z + 1) - __ldg(g.offsets + z) : g.n;
}

// Scan state handed from the head phase to the tail phase (one per segment, in the workspace).
struct __align__(16) ScanState {
    int nsel;        // boxes selected so far (= entries of krows / keep_idx)
    int done;        // the result is final: later phases exit at once
    int nvalid;      // candidates (score > -FLT_MAX), published by the rank sort: one 16-byte load gets all three
    int pad;
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
constexpr int RS_ITILE = 16;        // boxes ranked per CTA
constexpr int RS_JSPLIT = 16;       // threads sharing one box's j-range (lanes of different boxes read the same
                                    // 16-byte key vectors: shared-memory broadcast)
constexpr int RS_THREADS = RS_ITILE * RS_JSPLIT;
constexpr int RS_KTILE = 8192;      // keys staged per shared-memory tile (32 KB)

// writes box `ii` of the segment at sorted position `r`: original index + canonical corners + volume
__device__ __forceinline__ void place_box(const float *__restrict__ boxes, int ii, int r, int *__restrict__ sorted_idx,
                                          SBox *__restrict__ sboxes)
{
    sorted_idx[r] = ii;
    const float *b = boxes + (size_t)ii * 6;
    const float b0 = __ldg(b + 0), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
    const float b3 = __ldg(b + 3), b4 = __ldg(b + 4), b5 = __ldg(b + 5);
    SBox s;
    s.ymin = fminf(b0, b3); s.ymax = fmaxf(b0, b3);
    s.xmin = fminf(b1, b4); s.xmax = fmaxf(b1, b4);
    s.zmin = fminf(b2, b5); s.zmax = fmaxf(b2, b5);
    s.vol = __fmul_rn(__fmul_rn(__fsub_rn(s.ymax, s.ymin), __fsub_rn(s.xmax, s.xmin)), __fsub_rn(s.zmax, s.zmin));
    s.pad = 0.f;
    sboxes[r] = s;
}

// 1a. scores -> 32-bit sort keys, once (padded to whole 4-key vectors with the "not a candidate" key), so that every
// rank-sort CTA stages them with independent 16-byte loads instead of converting all n scores again
__global__ void __launch_bounds__(256)
nms_keys_kernel(const float *__restrict__ scores, NmsSeg seg, int kstride, unsigned *__restrict__ keys)
{
    const int z = blockIdx.y, n = seg_size(seg, z);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((n + 3) & ~3)) return;
    keys[(size_t)z * kstride + i] = (i < n) ? score_key(__ldg(scores + seg_begin(seg, z) + i)) : 0xFFFFFFFFu;
}

// DIRECT: single problem with 16-byte aligned scores -- the CTA converts the scores it stages itself (float4 loads), which
// saves the key pre-pass launch; otherwise it stages the pre-computed keys.
template <bool DIRECT>
__global__ void __launch_bounds__(RS_THREADS)
nms_rank_sort_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, const unsigned *__restrict__ keys,
                     int kstride, NmsSeg seg, int *__restrict__ sorted_idx, SBox *__restrict__ sboxes,
                     int *__restrict__ nvalid_out, ScanState *__restrict__ state)
{
    pdl_trigger();                                                     // the mask kernel may start launching
    const int z = blockIdx.y, n = seg_size(seg, z);
    if (blockIdx.x > 0 && blockIdx.x * RS_ITILE >= n) return;          // CTA-uniform; CTA 0 still publishes nvalid
    {
        const int base = seg_begin(seg, z);
        boxes += (size_t)base * 6;
        scores += base;
        keys += (size_t)z * kstride;
        sorted_idx += (size_t)z * seg.stride;
        sboxes += (size_t)z * seg.stride;
        nvalid_out += z;
    }
    __shared__ __align__(16) unsigned s_keys[RS_KTILE];
    __shared__ int s_rank[RS_ITILE];
    __shared__ int s_valid;
    const int il = threadIdx.x / RS_JSPLIT;                   // box within the CTA's tile
    const int jq = threadIdx.x % RS_JSPLIT;                   // lane over the key tile
    const int i = blockIdx.x * RS_ITILE + il;
    const unsigned ki = (i < n) ? (DIRECT ? score_key(__ldg(scores + i)) : __ldg(keys + i)) : 0xFFFFFFFFu;
    if (threadIdx.x < RS_ITILE) s_rank[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_valid = 0;
    int cnt = 0, nvalid = 0;
    // region split at 4-key vector granularity: keys strictly before the CTA's boxes precede on
    // `key <= ki`, keys strictly after on `key < ki`; only the few vectors that straddle the CTA's
    // own boxes need the full (key, index) comparison.  Padding keys are 0xFFFFFFFF with j >= n.
    const int i_lo = (blockIdx.x * RS_ITILE) & ~3, i_hi = (blockIdx.x * RS_ITILE + RS_ITILE + 3) & ~3;
    for (int j0 = 0; j0 < n; j0 += RS_KTILE) {
        __syncthreads();
        const int tile = min(RS_KTILE, n - j0);
        const uint4 *src = reinterpret_cast<const uint4 *>(keys + j0);
        const float4 *fsrc = reinterpret_cast<const float4 *>(scores + j0);
        for (int t = threadIdx.x; t < (tile + 3) / 4; t += RS_THREADS) {
            uint4 k4;
            if constexpr (DIRECT) {
                if (4 * t + 4 <= tile) {
                    const float4 f = __ldg(fsrc + t);
                    k4 = make_uint4(score_key(f.x), score_key(f.y), score_key(f.z), score_key(f.w));
                } else {                                       // last, partial vector: pad with "not a candidate"
                    const float *fs = scores + j0 + 4 * t;
                    const int r = tile - 4 * t;
                    k4 = make_uint4(score_key(__ldg(fs)), r > 1 ? score_key(__ldg(fs + 1)) : 0xFFFFFFFFu,
                                    r > 2 ? score_key(__ldg(fs + 2)) : 0xFFFFFFFFu, 0xFFFFFFFFu);
                }
            } else {
                k4 = __ldg(src + t);
            }
            reinterpret_cast<uint4 *>(s_keys)[t] = k4;
            if (blockIdx.x == 0)
                nvalid += (k4.x != 0xFFFFFFFFu) + (k4.y != 0xFFFFFFFFu) + (k4.z != 0xFFFFFFFFu) + (k4.w != 0xFFFFFFFFu);
        }
        __syncthreads();
        const int tile4 = (tile + 3) & ~3;
        // a box j precedes i iff key_j < key_i, or key_j == key_i and j < i
        const int t_lo = min(max(i_lo - j0, 0), tile4), t_hi = min(max(i_hi - j0, 0), tile4);
        for (int t = jq * 4; t < t_lo; t += RS_JSPLIT * 4) {
            const uint4 k4 = *reinterpret_cast<const uint4 *>(&s_keys[t]);
            cnt += (k4.x <= ki) + (k4.y <= ki) + (k4.z <= ki) + (k4.w <= ki);
        }
        for (int t = t_lo + jq * 4; t < t_hi; t += RS_JSPLIT * 4) {
            const uint4 k4 = *reinterpret_cast<const uint4 *>(&s_keys[t]);
            const int j = j0 + t;
            cnt += (k4.x < ki) || (k4.x == ki && j + 0 < i);
            cnt += (k4.y < ki) || (k4.y == ki && j + 1 < i);
            cnt += (k4.z < ki) || (k4.z == ki && j + 2 < i);
            cnt += (k4.w < ki) || (k4.w == ki && j + 3 < i);
        }
        for (int t = t_hi + jq * 4; t < tile4; t += RS_JSPLIT * 4) {
            const uint4 k4 = *reinterpret_cast<const uint4 *>(&s_keys[t]);
            cnt += (k4.x < ki) + (k4.y < ki) + (k4.z < ki) + (k4.w < ki);
        }
    }
    // reduce the RS_JSPLIT partial counts of each box (its lanes are adjacent within a warp)
#pragma unroll
    for (int o = RS_JSPLIT / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (blockIdx.x == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd(&s_valid, nvalid);
    }
    if (jq == 0) s_rank[il] = cnt;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *nvalid_out = s_valid;
        state[z] = ScanState{0, 0, s_valid, 0};
    }
    if (threadIdx.x < RS_ITILE) {
        const int ii = blockIdx.x * RS_ITILE + threadIdx.x;
        if (ii < n) place_box(boxes, ii, s_rank[threadIdx.x], sorted_idx, sboxes);
    }
}

// ---------------------------------------------------------------------------------
// 1b. bucketed sort for large n: the rank by counting above is O(n^2) (66 us at 20 k, 1.6 ms at 100 k).  Scores are
// binned into NB_M + 2 order-preserving buckets (uniform bins over [0, 1], one bin for negative scores, one for
// non-candidates), a histogram + scan gives every bucket its slice of the sorted order, the (key, index) pairs are
// scattered into their slices, and each element ranks itself inside its slice only.  Skewed scores (everything in one
// bin) degrade towards the O(n^2) count, never to a wrong order: the in-slice comparison is the full (key, index) one.
// ---------------------------------------------------------------------------------
constexpr int NB_M = 4096;
constexpr int NB = NB_M + 2;                   // bucket NB_M: negative scores, NB_M + 1: not a candidate
constexpr int NB_PITCH = 4352;                 // 17 * 256 (>= NB + 1, one scan element block per thread)

__device__ __forceinline__ int score_bucket(float s, unsigned key) {
    if (key == 0xFFFFFFFFu) return NB_M + 1;
    const float u = fminf(fmaxf(s, 0.0f), 1.0f);
    return (int)floorf(__fmul_rn(__fsub_rn(1.0f, u), (float)NB_M));           // monotone: higher score, lower bucket
}

__global__ void __launch_bounds__(256)
nms_bucket_hist_kernel(const float *__restrict__ scores, NmsSeg seg, int kstride, unsigned *__restrict__ keys,
                       int *__restrict__ hist)
{
    const int z = blockIdx.y, n = seg_size(seg, z);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((n + 3) & ~3)) return;
    unsigned key = 0xFFFFFFFFu;
    if (i < n) {
        const float sc = __ldg(scores + seg_begin(seg, z) + i);
        key = score_key(sc);
        atomicAdd(hist + (size_t)z * 2 * NB_PITCH + score_bucket(sc, key), 1);
    }
    keys[(size_t)z * kstride + i] = key;
}

// every CTA scans the histogram itself (4 K counters, cheaper than one more launch); CTA 0 publishes the bucket starts
__global__ void __launch_bounds__(1024)
nms_bucket_scatter_kernel(const float *__restrict__ scores, const unsigned *__restrict__ keys, int kstride, NmsSeg seg,
                          int *__restrict__ hist, int *__restrict__ start, unsigned *__restrict__ tkey,
                          int *__restrict__ tidx, int *__restrict__ nvalid_out, ScanState *__restrict__ state)
{
    __shared__ int s_start[NB_PITCH];
    __shared__ int s_warp[32];
    const int z = blockIdx.y, n = seg_size(seg, z);
    int *h = hist + (size_t)z * 2 * NB_PITCH, *fill = h + NB_PITCH;
    constexpr int PER = NB_PITCH / 1024 + 1;                                   // 5 counters per thread
    int v[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int b = threadIdx.x * PER + q;
        v[q] = (b < NB) ? h[b] : 0;
        sum += v[q];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int b = threadIdx.x * PER + q;
        if (b < NB_PITCH) s_start[b] = run;
        run += v[q];
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b <= NB; b += blockDim.x) start[(size_t)z * NB_PITCH + b] = s_start[b];
        if (threadIdx.x == 0) {
            const int nvalid = s_start[NB_M + 1];                              // everything before the last bucket
            nvalid_out[z] = nvalid;
            state[z] = ScanState{0, 0, nvalid, 0};
        }
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned key = __ldg(keys + (size_t)z * kstride + i);
        const int b = score_bucket(__ldg(scores + seg_begin(seg, z) + i), key);
        const int pos = s_start[b] + atomicAdd(fill + b, 1);
        tkey[(size_t)z * seg.stride + pos] = key;
        tidx[(size_t)z * seg.stride + pos] = i;
    }
}

__global__ void __launch_bounds__(256)
nms_bucket_rank_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, NmsSeg seg,
                       const int *__restrict__ start, const unsigned *__restrict__ tkey, const int *__restrict__ tidx,
                       int *__restrict__ sorted_idx, SBox *__restrict__ sboxes)
{
    pdl_trigger();
    const int z = blockIdx.y, n = seg_size(seg, z);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    tkey += (size_t)z * seg.stride;
    tidx += (size_t)z * seg.stride;
    const unsigned key = __ldg(tkey + p);
    const int idx = __ldg(tidx + p);
    const int base = seg_begin(seg, z);
    const int b = score_bucket(__ldg(scores + base + idx), key);
    const int lo = __ldg(start + (size_t)z * NB_PITCH + b), hi = __ldg(start + (size_t)z * NB_PITCH + b + 1);
    int cnt = 0;
    for (int q = lo; q < hi; ++q) {
        const unsigned kq = __ldg(tkey + q);
        cnt += (kq < key) || (kq == key && __ldg(tidx + q) < idx);
    }
    place_box(boxes + (size_t)base * 6, idx, lo + cnt, sorted_idx + (size_t)z * seg.stride, sboxes + (size_t)z * seg.stride);
}

// ---------------------------------------------------------------------------------
// 2. pairwise IoU bitmask
// ---------------------------------------------------------------------------------
constexpr int MK_ROWS = 128;        // row boxes staged per CTA (tail phase; the small head phase uses 32 for more CTAs)
constexpr int MK_WARPS = 8;         // mask words (32 columns each) per CTA

// IOU<float> (NMS.so@0xb500) on canonical boxes: same operations, same order.  The reference's
// early returns (volume <= 0 -> 0) need no test here: a box with a zero side has a zero
// intersection with everything (dy <= side), and 0 / positive == 0 exactly.
__device__ __forceinline__ bool iou_ge(const float4 a0, const float4 a1, const float4 b0, const float4 b1, float thr) {
    // x0 = (ymin,xmin,zmin,ymax)  x1 = (xmax,zmax,vol,-)
    const float dy = fmaxf(0.0f, __fsub_rn(fminf(a0.w, b0.w), fmaxf(a0.x, b0.x)));
    const float dx = fmaxf(0.0f, __fsub_rn(fminf(a1.x, b1.x), fmaxf(a0.y, b0.y)));
    const float dz = fmaxf(0.0f, __fsub_rn(fminf(a1.y, b1.y), fmaxf(a0.z, b0.z)));
    const float inter = __fmul_rn(__fmul_rn(dy, dx), dz);
    if (!(inter > 0.0f)) return 0.0f >= thr;
    if (a1.z <= 0.0f || b1.z <= 0.0f) return 0.0f >= thr;     // volume product underflowed to 0 (reference returns 0)
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1.z, b1.z), inter));
    return iou >= thr;
}

template <int ROWS>
__global__ void __launch_bounds__(MK_WARPS * 32)
nms_mask_kernel(const SBox *__restrict__ sboxes, NmsSeg seg, int pitch_words, float thr,
                int word_begin, int row_end, const ScanState *__restrict__ state, unsigned *__restrict__ mask,
                const int *__restrict__ krows, int known_rows)
{
    // The grid covers mask words [word_begin, ...) of rows [0, row_end): the head phase computes the top-left
    // triangle (word_begin = 0, row_end = T), the tail phase everything right of it (word_begin = T / 32, all rows) --
    // unless the head phase's scan already finished (state->done), which is the common case when max_out << n.
    // Rows below `known_rows` were already visited by the scans of the earlier phases: the scan will only ever read
    // the rows it KEPT (their positions are in krows[0 .. state.nsel)), so the mask words of the suppressed ones are
    // not computed at all -- with dense proposals (what a trained RPN produces) that is most of the rows above a tail
    // window (6000 dense boxes: 1536 x 4464 pairs shrink to ~250 x 4464).  The skipped words keep whatever the workspace held.
    __shared__ float4 s_rows[ROWS * 2];
    __shared__ unsigned s_kept[ROWS / 32];
    pdl_wait();                                                          // predecessor (sort / head scan) complete
    pdl_trigger();
    const int z = blockIdx.z, n = min(seg_size(seg, z), row_end);
    if (state != nullptr && state[z].done) return;                       // CTA-uniform
    const int i0 = blockIdx.y * ROWS;
    const int w0 = word_begin + blockIdx.x * MK_WARPS;
    if (i0 >= n || w0 * 32 >= seg_size(seg, z)) return;                 // CTA-uniform
    sboxes += (size_t)z * seg.stride;
    mask += (size_t)z * seg.stride * pitch_words;
    // only words that contain some column j >= i0 are ever read by the scan
    if ((w0 + MK_WARPS) * 32 <= i0) return;
    const int ncol = seg_size(seg, z);
    const int rows = min(ROWS, n - i0);
    const bool use_kept = known_rows > 0 && i0 < known_rows;             // CTA-uniform
    if (use_kept) {
        if (threadIdx.x < ROWS / 32) s_kept[threadIdx.x] = 0u;
        __syncthreads();
        const int nsel = state[z].nsel;
        const int *kr = krows + (size_t)z * seg.stride;
        for (int q = threadIdx.x; q < nsel; q += blockDim.x) {
            const int r = __ldg(kr + q) - i0;
            if (r >= 0 && r < ROWS) atomicOr(&s_kept[r >> 5], 1u << (r & 31));
        }
        if (threadIdx.x < ROWS && i0 + (int)threadIdx.x >= known_rows)     // rows of this block the scans have not reached yet
            atomicOr(&s_kept[threadIdx.x >> 5], 1u << (threadIdx.x & 31));
    }
    for (int t = threadIdx.x; t < rows * 2; t += blockDim.x)
        s_rows[t] = __ldg(reinterpret_cast<const float4 *>(sboxes + i0) + t);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = w0 + warp;
    if (w * 32 >= ncol || (w + 1) * 32 <= i0) return;
    const int j = w * 32 + lane;
    // padding columns (j >= n): a far-away zero-size box never intersects anything
    float4 b0 = make_float4(3e38f, 3e38f, 3e38f, 3e38f), b1 = make_float4(3e38f, 3e38f, 0.f, 0.f);
    if (j < ncol) {
        b0 = __ldg(reinterpret_cast<const float4 *>(sboxes + j));
        b1 = __ldg(reinterpret_cast<const float4 *>(sboxes + j) + 1);
    }
    const bool pad_bit = (j >= ncol);
    for (int rb = 0; rb < rows; rb += 32) {
        // rows whose word lies entirely below the diagonal are never read
        if ((w + 1) * 32 <= i0 + rb) continue;
        const unsigned kw = use_kept ? s_kept[rb >> 5] : 0xFFFFFFFFu;    // rows of this group the scan can still read
        if (kw == 0u) continue;
        unsigned word = 0;
        const int rend = min(32, rows - rb);
#pragma unroll 8
        for (int r = 0; r < rend; ++r) {
            if (!((kw >> r) & 1u)) continue;                             // warp-uniform
            const bool bit = iou_ge(s_rows[2 * (rb + r)], s_rows[2 * (rb + r) + 1], b0, b1, thr) && !pad_bit;
            const unsigned wd = __ballot_sync(0xffffffffu, bit);
            if (lane == r) word = wd;
        }
        if (lane < rend && ((kw >> lane) & 1u)) mask[(size_t)(i0 + rb + lane) * pitch_words + w] = word;
    }
}

// ---------------------------------------------------------------------------------
// 3. greedy scan (single CTA, warp-level)
//
// Boxes are visited in super-chunks of SC_SB = 512 sorted boxes (16 mask words).  While warp 0
// resolves super-chunk s, the other warps prefetch the 512 x 16-word diagonal block of
// super-chunk s+1 into shared memory, so the serial chain of warp 0 touches only shared
// memory and registers:
//   - lane w (< 16) keeps word w of the super-chunk's `removed` bits in a register;
//   - a 32-box chunk is resolved with the parallel greedy rule (a box is kept once every
//     earlier conflicting box of the chunk is decided removed; it is removed once one of them
//     is decided kept) -- a few __ballot_sync rounds instead of 32 dependent steps;
//   - the kept rows of the chunk are OR-ed into the lanes' words from the shared block.
// Suppression reaches later super-chunks lazily, one super-chunk ahead: while warp 0 resolves s, the other warps
// OR the mask words of super-chunk s+1 over all rows kept BEFORE s (their positions are logged in `krows`); at the
// boundary all 32 warps add the rows kept IN s.  Each (kept row, word) pair is read at most once, in batches of
// independent loads, and never for columns the scan does not reach (it stops at max_out).
// ---------------------------------------------------------------------------------
constexpr int SC_THREADS = 1024;
constexpr int SC_SB = 512;                 // boxes per super-chunk
constexpr int SC_W = SC_SB / 32;           // words per super-chunk row (16)
constexpr int SC_P = SC_W + 1;             // padded row pitch in shared memory (bank-conflict free diagonal reads)

// stage super-chunk s: its 512 x 16-word diagonal mask block, the boxes' volumes and original indices
__device__ __forceinline__ void sc_prefetch(unsigned *dst, float *vol, int *sidx, const unsigned *__restrict__ mask,
                                            int pitch_words, const SBox *__restrict__ sboxes,
                                            const int *__restrict__ sorted_idx, int s, int nvalid, int nwords,
                                            int tid0, int nthreads)
{
    const int row0 = s * SC_SB, w0 = s * SC_W;
    for (int t = tid0; t < SC_SB * SC_W; t += nthreads) {
        const int r = t / SC_W, w = t % SC_W;
        unsigned v = 0u;
        if (row0 + r < nvalid && w0 + w < nwords) v = __ldg(mask + (size_t)(row0 + r) * pitch_words + w0 + w);
        dst[r * SC_P + w] = v;
    }
    for (int r = tid0; r < SC_SB; r += nthreads) {
        const bool ok = row0 + r < nvalid;
        vol[r] = ok ? sboxes[row0 + r].vol : 1.0f;
        sidx[r] = ok ? __ldg(sorted_idx + row0 + r) : -1;
    }
}

// OR the mask words [w0, w0+16) of the kept rows krows[q0 .. q1) into dst[16].  Thread t of `nthreads` owns word
// t % 16 of every (t / 16)-th row: 16 threads read one 64-byte row segment, the loads of a thread are independent
// (accumulated in a register, 8 in flight), one shared atomic per thread at the end.
__device__ __forceinline__ void sc_or_rows(unsigned *dst, const unsigned *__restrict__ mask, int pitch_words,
                                           const int *krows, int q0, int q1, int w0, int nwords, int tid, int nthreads)
{
    const int w = tid & (SC_W - 1), g = tid / SC_W, ng = nthreads / SC_W;
    if (g >= ng || w0 + w >= nwords) return;
    const unsigned *col = mask + w0 + w;
    unsigned acc = 0u;
    int q = q0 + g;
    for (; q + 7 * ng < q1; q += 8 * ng) {
        unsigned m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) m[u] = __ldg(col + (size_t)krows[q + u * ng] * pitch_words);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc |= m[u];
    }
    for (; q < q1; q += ng) acc |= __ldg(col + (size_t)krows[q] * pitch_words);
    if (acc) atomicOr(dst + w, acc);
}

__global__ void __launch_bounds__(SC_THREADS)
nms_scan_kernel(const unsigned *__restrict__ mask, int pitch_words, const SBox *__restrict__ sboxes,
                const int *__restrict__ sorted_idx, const int *__restrict__ nvalid_p, int seg_stride,
                int max_out, float thr, int sb_begin, int sb_end, ScanState *__restrict__ state,
                int *__restrict__ krows, int *__restrict__ keep_idx, int *__restrict__ keep_count)
{
    // Visits super-chunks [sb_begin, sb_end).  The head phase (sb_begin = 0) stops at sb_end = T / 512 and hands
    // {nsel, krows, keep_idx} over through `state`; the tail phase resumes there -- or exits at once when the head
    // phase already produced the final result.
    extern __shared__ unsigned s_dyn[];
    unsigned *s_blk = s_dyn;                                   // 2 x SC_SB x SC_P words
    float *s_vol = reinterpret_cast<float *>(s_blk + 2 * SC_SB * SC_P);   // 2 x SC_SB
    int *s_sidx = reinterpret_cast<int *>(s_vol + 2 * SC_SB);             // 2 x SC_SB
    __shared__ unsigned s_rem[2][SC_W];                        // removed bits of the current / the next super-chunk
    __shared__ int s_nsel, s_fill, s_done;
    {
        const size_t z = blockIdx.x;                                   // one CTA per segment
        mask += z * seg_stride * pitch_words;
        sboxes += z * seg_stride;
        sorted_idx += z * seg_stride;
        krows += z * seg_stride;
        nvalid_p += z;
        keep_idx += z * max_out;
        keep_count += z;
        state += z;
    }
    pdl_wait();                                                // predecessor (the phase's mask kernel) complete
    pdl_trigger();
    const int4 st0 = *reinterpret_cast<const int4 *>(state);   // {nsel, done, nvalid, -}
    if (st0.y) return;                                         // CTA-uniform (written by the previous phase)
    const int nvalid = st0.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwords = (nvalid + 31) >> 5;
    const int nsuper = (nvalid + SC_SB - 1) / SC_SB;
    const int s_stop = min(sb_end, nsuper);
    const int k_start = st0.x;
    if (threadIdx.x < 2 * SC_W) s_rem[threadIdx.x / SC_W][threadIdx.x % SC_W] = 0u;
    if (threadIdx.x == 0) { s_nsel = k_start; s_fill = -1; s_done = 0; }
    if (sb_begin < s_stop)
        sc_prefetch(s_blk + (sb_begin & 1) * (SC_SB * SC_P), s_vol + (sb_begin & 1) * SC_SB, s_sidx + (sb_begin & 1) * SC_SB,
                    mask, pitch_words, sboxes, sorted_idx, sb_begin, nvalid, nwords, threadIdx.x, SC_THREADS);
    __syncthreads();
    if (sb_begin > 0 && sb_begin < s_stop) {                   // resume: rebuild the first super-chunk's removed words
        sc_or_rows(s_rem[sb_begin & 1], mask, pitch_words, krows, 0, k_start, sb_begin * SC_W, nwords, threadIdx.x, SC_THREADS);
        __syncthreads();
    }
    const bool self_suppresses = (0.0f >= thr);                // thr == 0: even a zero-volume box suppresses itself
    // rows kept so far, carried in a register by every thread: s_nsel is written by warp 0 inside an iteration and
    // may only be read between the barrier that follows the chain and the next barrier (never at the top of the loop,
    // where a late warp could already see warp 0's new value)
    int k_cur = k_start;

    for (int s = sb_begin; s < s_stop; ++s) {
        const int cur = s & 1, nxt = cur ^ 1;
        unsigned *blk = s_blk + cur * (SC_SB * SC_P);
        const float *bvol = s_vol + cur * SC_SB;
        const int *bsidx = s_sidx + cur * SC_SB;
        const int k_before = k_cur;                            // rows kept before this super-chunk: krows[0 .. k_before)
        if (warp != 0) {
            // background: stage the diagonal block of the next super-chunk and OR the rows kept so far into ITS
            // removed words -- suppression is propagated lazily, one super-chunk ahead, never to columns the scan
            // may not reach
            if (s + 1 < s_stop) {
                sc_prefetch(s_blk + nxt * (SC_SB * SC_P), s_vol + nxt * SC_SB, s_sidx + nxt * SC_SB, mask, pitch_words,
                            sboxes, sorted_idx, s + 1, nvalid, nwords, threadIdx.x - 32, SC_THREADS - 32);
                sc_or_rows(s_rem[nxt], mask, pitch_words, krows, 0, k_before, (s + 1) * SC_W, nwords,
                           threadIdx.x - 32, SC_THREADS - 32);
            }
        } else {
            int nsel = k_before, fill = -1;
            bool done = false;
            const unsigned lt = (1u << lane) - 1u;
            // mine_or[w]: OR of word w over MY rows kept so far in this super-chunk (lane b owns row b of every
            // chunk), seeded with the removed bits inherited from earlier super-chunks.  The chunk loop is fully
            // unrolled so that the array lives in registers; the only warp reduction on the serial chain is the one
            // that forms the current chunk's removed word.
            unsigned mine_or[SC_W];
            unsigned degen_lanes[SC_W];                        // zero-volume boxes per chunk (independent of the chain)
#pragma unroll
            for (int w = 0; w < SC_W; ++w) {
                mine_or[w] = (lane == 0) ? ((s * SC_W + w < nwords) ? s_rem[cur][w] : 0xFFFFFFFFu) : 0u;
                const bool dg = (s * SC_SB + w * 32 + lane < nvalid) && !(bvol[w * 32 + lane] > 0.0f) && !self_suppresses;
                degen_lanes[w] = __ballot_sync(0xffffffffu, dg);
            }
#pragma unroll
            for (int c = 0; c < SC_W; ++c) {
                if (done || s * SC_SB + c * 32 >= nvalid) continue;      // warp-uniform
                const int row = s * SC_SB + c * 32 + lane;
                const unsigned *myrow = blk + (c * 32 + lane) * SC_P;
                const unsigned d = myrow[c];                   // rows >= nvalid were staged as zeros
                const unsigned nextw = (c + 1 < SC_W) ? myrow[c + 1] : 0u;   // needed right after this chunk: preload
                const bool valid = row < nvalid;
                const unsigned rw = __reduce_or_sync(0xffffffffu, mine_or[c]);
                const unsigned sup = d & lt;                   // earlier boxes of this chunk that conflict with me
                unsigned undecided = __ballot_sync(0xffffffffu, valid && !((rw >> lane) & 1u));
                unsigned kept = 0u;
                // parallel greedy rule: a box is kept once no earlier conflicting box of the chunk is still
                // undecided or kept; it is dropped as soon as one of them is kept.  One ballot when nothing conflicts.
                while (undecided) {
                    const bool me = (undecided >> lane) & 1u;
                    const bool keep = me && !(sup & (kept | undecided));
                    const unsigned nkp = __ballot_sync(0xffffffffu, keep);
                    kept |= nkp;
                    if (!(undecided & ~nkp)) break;
                    undecided = __ballot_sync(0xffffffffu, me && !keep && !(sup & kept));
                }
                // zero-volume quirk (NMS.so@0xdc55): a selected box whose self-IoU is 0 is selected again
                // until max_out; everything after it is never reached
                const unsigned degen = degen_lanes[c] & kept;
                if (degen) {
                    const int bq = __ffs(degen) - 1;
                    kept &= (2u << bq) - 1u;
                    fill = s * SC_SB + c * 32 + bq;
                    done = true;
                }
                int cnt = __popc(kept);
                if (nsel + cnt >= max_out) {                   // keep only the first (max_out - nsel) of them
                    const int room = max_out - nsel;
                    if (cnt > room) {
                        const bool mine = ((kept >> lane) & 1u) && __popc(kept & lt) < room;
                        kept = __ballot_sync(0xffffffffu, mine);
                        cnt = room;
                        fill = -1;                             // the degenerate box (if any) was beyond max_out
                    }
                    done = true;
                }
                const bool kept_me = (kept >> lane) & 1u;
                if (c + 1 < SC_W) mine_or[c + 1] |= kept_me ? nextw : 0u;
                if (kept_me) {
                    const int slot = __popc(kept & lt);
                    keep_idx[nsel + slot] = bsidx[c * 32 + lane];
                    krows[nsel + slot] = row;
#pragma unroll
                    for (int w = c + 2; w < SC_W; ++w) mine_or[w] |= myrow[w];   // off the serial chain
                }
                nsel += cnt;
            }
            if (lane == 0) {
                s_nsel = nsel;
                s_done = done || nsel >= max_out;
                s_fill = (fill >= 0) ? bsidx[fill - s * SC_SB] : -1;
            }
        }
        __syncthreads();                                       // also publishes warp 0's krows[] stores to the CTA
        k_cur = s_nsel;
        if (s_done || s + 1 >= s_stop) break;
        // boundary: the rows kept in THIS super-chunk complete the next super-chunk's removed words
        sc_or_rows(s_rem[nxt], mask, pitch_words, krows, k_before, k_cur, (s + 1) * SC_W, nwords, threadIdx.x, SC_THREADS);
        if (threadIdx.x < SC_W) s_rem[cur][threadIdx.x] = 0u;  // becomes the accumulator of super-chunk s + 2
        __syncthreads();
    }
    __syncthreads();
    int nsel = s_nsel;
    if (!s_done && s_stop < nsuper) {                          // more super-chunks to visit: hand over to the tail phase
        if (threadIdx.x == 0) state->nsel = nsel;
        return;
    }
    const int fill = s_fill;
    if (fill >= 0) {                                           // zero-volume quirk: repeat until max_out
        for (int t = nsel + threadIdx.x; t < max_out; t += blockDim.x) keep_idx[t] = fill;
        nsel = max_out;
    }
    if (threadIdx.x == 0) {
        *keep_count = nsel;
        state->done = 1;
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct NmsLayout {
    int pitch_words;
    int kstride;
    size_t off_sidx, off_sboxes, off_nvalid, off_state, off_krows, off_keys, off_hist, off_start, off_tkey, off_tidx, off_mask, total;
};

static NmsLayout nms_layout(int n, int segments) {
    NmsLayout L;
    const int words = (n + 31) / 32;
    const size_t S = (size_t)(segments > 0 ? segments : 1);
    L.pitch_words = (int)align_up((size_t)(words > 0 ? words : 1), 32);   // 128-byte rows
    size_t off = 0;
    L.off_nvalid = off; off += align_up(sizeof(int) * S, 256);
    L.off_state = off;  off += align_up(sizeof(ScanState) * S, 256);
    L.off_sidx = off;   off += align_up(sizeof(int) * S * n, 256);
    L.off_sboxes = off; off += align_up(sizeof(SBox) * S * n, 256);
    L.off_krows = off;  off += align_up(sizeof(int) * S * n, 256);
    L.kstride = (int)align_up((size_t)n, 4);
    L.off_keys = off;   off += align_up(sizeof(unsigned) * S * L.kstride, 256);
    L.off_hist = off;   off += align_up(sizeof(int) * S * 2 * NB_PITCH, 256);      // histogram + fill cursors
    L.off_start = off;  off += align_up(sizeof(int) * S * NB_PITCH, 256);
    L.off_tkey = off;   off += align_up(sizeof(unsigned) * S * n, 256);
    L.off_tidx = off;   off += align_up(sizeof(int) * S * n, 256);
    L.off_mask = off;   off += align_up(sizeof(unsigned) * S * n * L.pitch_words, 256);
    L.total = off;
    return L;
}

size_t nms3d_workspace_bytes(int n, int segments) { return n <= 0 ? 256 : nms_layout(n, segments).total; }

// segments == 0: one problem of n boxes (seg_offsets ignored); else `segments` problems of at most n boxes each
int launch_nms3d(const float *boxes, const float *scores, const int *seg_offsets, int segments, int n, int max_out,
                 float thr, int *keep_idx, int *keep_count, void *ws, size_t ws_bytes, cudaStream_t stream)
{
    const int S = segments > 0 ? segments : 1;
    if (n <= 0 || max_out <= 0) {
        ROI3D_CUDA_TRY(cudaMemsetAsync(keep_count, 0, sizeof(int) * S, stream));
        return ROI3D_OK;
    }
    const NmsLayout L = nms_layout(n, S);
    if (ws == nullptr || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return ROI3D_EWORKSPACE;
    if (S > 65535 || n > (1 << 20)) return ROI3D_EUNSUPPORTED;
    char *base = static_cast<char *>(ws);
    int *nvalid = reinterpret_cast<int *>(base + L.off_nvalid);
    int *sidx = reinterpret_cast<int *>(base + L.off_sidx);
    SBox *sboxes = reinterpret_cast<SBox *>(base + L.off_sboxes);
    int *krows = reinterpret_cast<int *>(base + L.off_krows);
    unsigned *mask = reinterpret_cast<unsigned *>(base + L.off_mask);
    const NmsSeg seg{segments > 0 ? seg_offsets : nullptr, n, n};

    ScanState *state = reinterpret_cast<ScanState *>(base + L.off_state);
    unsigned *keys = reinterpret_cast<unsigned *>(base + L.off_keys);
    const int sort_variant = option_value(OPT_NMS_SORT);                   // 0 auto, 1 rank by counting, 2 bucketed
    if (sort_variant == 2 || (sort_variant == 0 && n >= 8192)) {
        int *hist = reinterpret_cast<int *>(base + L.off_hist);
        int *start = reinterpret_cast<int *>(base + L.off_start);
        unsigned *tkey = reinterpret_cast<unsigned *>(base + L.off_tkey);
        int *tidx = reinterpret_cast<int *>(base + L.off_tidx);
        ROI3D_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(int) * (size_t)S * 2 * NB_PITCH, stream));
        nms_bucket_hist_kernel<<<dim3((L.kstride + 255) / 256, S), 256, 0, stream>>>(scores, seg, L.kstride, keys, hist);
        ROI3D_LAUNCH_CHECK();
        nms_bucket_scatter_kernel<<<dim3((n + 1023) / 1024, S), 1024, 0, stream>>>(scores, keys, L.kstride, seg, hist, start, tkey,
                                                                                tidx, nvalid, state);
        ROI3D_LAUNCH_CHECK();
        nms_bucket_rank_kernel<<<dim3((n + 255) / 256, S), 256, 0, stream>>>(boxes, scores, seg, start, tkey, tidx, sidx, sboxes);
        ROI3D_LAUNCH_CHECK();
    } else {
        const dim3 rgrid((n + RS_ITILE - 1) / RS_ITILE, S);
        if (segments <= 0 && (reinterpret_cast<uintptr_t>(scores) & 15) == 0) {
            nms_rank_sort_kernel<true><<<rgrid, RS_THREADS, 0, stream>>>(boxes, scores, keys, L.kstride, seg, sidx, sboxes, nvalid, state);
        } else {
            nms_keys_kernel<<<dim3((L.kstride + 255) / 256, S), 256, 0, stream>>>(scores, seg, L.kstride, keys);
            ROI3D_LAUNCH_CHECK();
            nms_rank_sort_kernel<false><<<rgrid, RS_THREADS, 0, stream>>>(boxes, scores, keys, L.kstride, seg, sidx, sboxes, nvalid, state);
        }
        ROI3D_LAUNCH_CHECK();
    }
    const size_t smem = ((size_t)2 * SC_SB * SC_P + 4 * SC_SB) * sizeof(unsigned);
    ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(nms_scan_kernel), smem));
    // Head / tail split.  The scan stops as soon as max_out boxes are selected, i.e. after about max_out sorted boxes
    // when few of them suppress each other -- far from n.  So the first T = max_out * 1.25 + 256 boxes (rounded up
    // to whole super-chunks) get their own mask triangle and scan; the rest of the mask (the bulk of the n^2 / 2
    // pairs) and the tail scan are launched behind it and return immediately when the head already finished.
    // nms_variant 1 forces the single-phase schedule.
    const bool pdl = option_value(OPT_PDL) == 0;
    const int nsb = (n + SC_SB - 1) / SC_SB;
    int head_sb = (int)(((long long)max_out + max_out / 4 + 256 + SC_SB - 1) / SC_SB);
    if (option_value(OPT_NMS_VARIANT) == 1 || head_sb >= nsb) head_sb = nsb;
    // Phase boundaries in super-chunks.  Up to 8191 boxes: head + tail.  From 8192 boxes the tail -- the bulk of the
    // n^2 / 2 pairs -- is cut into windows that double in size (at most 4 phases): when the proposals are dense and the
    // scan has to go a few times deeper than max_out (what a trained RPN produces), only the windows it actually
    // reaches are computed (20 000 dense boxes: 0.376 -> see profiles/r2_nms_ncu_summary.txt); a window whose
    // predecessor finished the job costs two kernels that exit at once (~2 us each under PDL).
    int bounds[5] = {0, head_sb, nsb, nsb, nsb};
    int nphase = (head_sb < nsb) ? 2 : 1;
    if (head_sb < nsb && n >= 8192 && option_value(OPT_NMS_VARIANT) != 2) {
        nphase = 1;
        int b = head_sb;
        while (b < nsb && nphase < 4) {
            const int nb2 = (nphase == 3) ? nsb : min(nsb, 2 * b);
            bounds[nphase + 1] = nb2;
            b = nb2;
            ++nphase;
        }
        for (int q = nphase + 1; q < 5; ++q) bounds[q] = nsb;
    }
    for (int ph = 0; ph < nphase; ++ph) {
        const int sb0 = bounds[ph], sb1 = bounds[ph + 1];
        const int c0 = min(sb0 * SC_SB, n), c1 = min(sb1 * SC_SB, n);      // columns [c0, c1) for rows [0, c1)
        const int wb = c0 / 32, words = (c1 + 31) / 32 - wb;
        const ScanState *st = (ph == 0) ? nullptr : state;
        if (ph == 0 && nphase > 1) {                           // a small head triangle needs more, smaller CTAs
            dim3 mgrid((words + MK_WARPS - 1) / MK_WARPS, (c1 + 31) / 32, S);
            ROI3D_CUDA_TRY(launch_dependent(nms_mask_kernel<32>, mgrid, dim3(MK_WARPS * 32), 0, stream, pdl, (const SBox *)sboxes, seg,
                                            L.pitch_words, thr, wb, c1, st, mask, (const int *)krows, 0));
        } else {
            dim3 mgrid((words + MK_WARPS - 1) / MK_WARPS, (c1 + MK_ROWS - 1) / MK_ROWS, S);
            ROI3D_CUDA_TRY(launch_dependent(nms_mask_kernel<MK_ROWS>, mgrid, dim3(MK_WARPS * 32), 0, stream, pdl, (const SBox *)sboxes, seg,
                                            L.pitch_words, thr, wb, c1, st, mask, (const int *)krows,
                                            (ph > 0 && option_value(OPT_EXPERIMENT) != 8) ? c0 : 0));
        }
        ROI3D_LAUNCH_CHECK();
        ROI3D_CUDA_TRY(launch_dependent(nms_scan_kernel, dim3(S), dim3(SC_THREADS), smem, stream, pdl, (const unsigned *)mask, L.pitch_words,
                                        (const SBox *)sboxes, (const int *)sidx, (const int *)nvalid, n, max_out, thr, sb0, sb1, state,
                                        krows, keep_idx, keep_count));
        ROI3D_LAUNCH_CHECK();
    }
    return ROI3D_OK;
}

}  // namespace roi3d

// roi3d_car_plane.cu -- plane-staged separable CropAndResize3D (variant 2), the
// production kernels for trilinear sampling with C % 4 == 0.
//
// Why: trilinear taps form a product set.  For one ROI and one depth sample k the
// reference (CAR.so@0x4f88-0x5021) first lerps along z, then x, then y, each as
// a + (b - a) * t.  The z-lerp of a voxel column (yi, xi) depends only on (yi, xi, k),
// so a CTA that owns (ROI, k-range, 64-channel chunk)
//   A. reads every footprint voxel (yi, xi, {floor z, ceil z}) exactly ONCE from global
//      memory (coalesced 16-byte loads along channels), z-lerps it in registers and
//      parks the result in a shared-memory plane Z[row][col][chunk];
//   B. produces all ph*pw outputs of that k from 4 shared-memory reads each (x-lerp,
//      then y-lerp) and streams them out with 16-byte evict-first stores.
// The intermediate values are the very expressions the reference evaluates, in the same
// order and rounding, so the result is bit-identical to the direct kernel and to the
// oracle, while L2->SM traffic drops from 8 taps per output to 2 * footprint / outputs.
//
// Instruction economy (round-1 ncu: the first version issued ~190 warp instructions per
// 16-byte output and was issue-bound at 47 % of HBM): everything that depends only on
// (ROI, y-tile) -- footprint voxel offsets, and per output the four plane offsets plus the
// x/y lerp weights -- is tabulated once per CTA in shared memory, so the k loop does one
// table read, four plane reads, the 36 exactly-rounded fp32 ops and one store per output.
//
// The backward kernel is the transpose: the k-slice of grads is staged in shared memory
// (each element read once, coalesced), every footprint voxel gathers its (wy*wx)-weighted
// sum from that slice (deterministic, no shared-memory atomics) and issues two vector REDs
// (floor z, ceil z) instead of 8 per grads element.
#include "roi3d_common.cuh"

namespace roi3d {

constexpr int PL_THREADS = 256;
constexpr int PL_MAXP = 64;                 // max crop size per axis handled here
constexpr int PL_UNROLL = 4;
constexpr int PL_MAXOUT = 1024;             // output-table entries per y-tile

struct AxisTab {                            // one per axis (y, x), lives in shared memory
    short pos0[PL_MAXP];                    // footprint position of floor(in)   (-1 if sample invalid)
    short pos1[PL_MAXP];                    // footprint position of ceil(in)
    float t[PL_MAXP];                       // lerp weight in - floor(in)
    int   list[2 * PL_MAXP];                // footprint voxel indices, in first-occurrence (monotonic) order
    int   cand[2 * PL_MAXP];                // scratch: floor/ceil per sample
    unsigned char first[2 * PL_MAXP];       // scratch: 1 if first occurrence
    int   n;                                // footprint size
};

struct PlaneShared {
    AxisTab ax[2];
    short tile_y0[PL_MAXP + 1];             // y-sample tile boundaries
    short tile_r0[PL_MAXP], tile_r1[PL_MAXP];   // footprint rows spanned by each tile (r1 < r0: none)
    int ntiles;
    float box[6];
};

// Build pos0/pos1/t/list for both axes.  tid in [0,64): y sample, [64,128): x sample.
__device__ __forceinline__ void build_axis_tables(PlaneShared &S, const CarGeom &g)
{
    const int tid = threadIdx.x;
    if (tid < 2 * PL_MAXP) {
        const int a = tid / PL_MAXP, k = tid % PL_MAXP;
        const int p = a ? g.pw : g.ph, dim = a ? g.W : g.H;
        AxisTab &T = S.ax[a];
        int c0 = -1, c1 = -1;
        if (k < p) {
            const float a1 = S.box[a], a2 = S.box[3 + a];
            const float in = axis_coord(a1, a2, dim, p, k, axis_scale(a1, a2, dim, p));
            if (!axis_invalid(in, dim)) {
                const float fl = floorf(in);
                c0 = (int)fl;
                c1 = (int)ceilf(in);
                T.t[k] = __fsub_rn(in, fl);
            }
        }
        T.cand[2 * k] = c0;
        T.cand[2 * k + 1] = c1;
    }
    __syncthreads();
    {   // dedupe: candidate j is a "first" if no earlier candidate has its value
        const int a = tid / (2 * PL_MAXP), j = tid % (2 * PL_MAXP);
        const int p = a ? g.pw : g.ph;
        AxisTab &T = S.ax[a];
        int fo = j;
        const int v = T.cand[j];
        if (j < 2 * p && v >= 0) {
            for (int i = 0; i < j; ++i)
                if (T.cand[i] == v) { fo = i; break; }
        }
        T.first[j] = (j < 2 * p && v >= 0 && fo == j) ? 1 : 0;
        __syncthreads();
        if (j < 2 * p) {
            int pos = -1;
            if (v >= 0) {
                pos = 0;
                for (int i = 0; i < fo; ++i) pos += T.first[i];
                if (fo == j) T.list[pos] = v;
            }
            if (j & 1) T.pos1[j >> 1] = (short)pos; else T.pos0[j >> 1] = (short)pos;
        }
        if (j == 2 * p - 1) {
            int cnt = 0;
            for (int i = 0; i < 2 * p; ++i) cnt += T.first[i];
            T.n = cnt;
        }
    }
    __syncthreads();
}

// Greedy y-sample tiles: (rows spanned) * n_x <= zcap plane voxels and (samples) * pw <= PL_MAXOUT.
__device__ __forceinline__ void build_y_tiles(PlaneShared &S, const CarGeom &g, int zcap)
{
    if (threadIdx.x == 0) {
        const AxisTab &Y = S.ax[0];
        const int nx = max(S.ax[1].n, 1);
        int nt = 0, lo = 1 << 30, hi = -1, ya = 0;
        S.tile_y0[0] = 0;
        for (int y = 0; y < g.ph; ++y) {
            const bool valid = Y.pos0[y] >= 0;
            const int a = min((int)Y.pos0[y], (int)Y.pos1[y]), b = max((int)Y.pos0[y], (int)Y.pos1[y]);
            const int l2 = valid ? min(lo, a) : lo, h2 = valid ? max(hi, b) : hi;
            const bool rows_over = valid && hi >= 0 && (h2 - l2 + 1) * nx > zcap;
            const bool outs_over = (y - ya + 1) * g.pw > PL_MAXOUT && y > ya;
            if (rows_over || outs_over) {                     // close the tile before y
                S.tile_r0[nt] = (short)(hi < 0 ? 0 : lo); S.tile_r1[nt] = (short)hi;
                S.tile_y0[++nt] = (short)y;
                ya = y;
                lo = valid ? a : (1 << 30);
                hi = valid ? b : -1;
            } else { lo = l2; hi = h2; }
        }
        S.tile_r0[nt] = (short)(hi < 0 ? 0 : lo); S.tile_r1[nt] = (short)hi;
        S.tile_y0[++nt] = (short)g.ph;
        S.ntiles = nt;
    }
    __syncthreads();
}

struct PlaneLaunch {
    int cl;            // channel lanes (float4 each) per voxel
    int chunks;        // channel chunks per voxel
    int ksplits;       // depth-sample splits per ROI
    int zcap;          // plane capacity in voxels (forward) / staged grads entries (backward)
    int otab;          // output-table entries (forward)
};

struct __align__(16) OutEntry {             // per output (y, x) of the current y-tile
    unsigned o_top;                         // byte offsets into the plane: (t,l) | (t,r) << 16
    unsigned o_bot;                         //                              (b,l) | (b,r) << 16
    float xl, yl;
};

// ---------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_THREADS, 4)
car3d_fwd_plane_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                       const int *__restrict__ box_index, CarGeom g, PlaneLaunch L, float ext,
                       float *__restrict__ crops)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);  off += (sizeof(PlaneShared) + 15) & ~size_t(15);
    OutEntry *otab = reinterpret_cast<OutEntry *>(smem_raw + off); off += sizeof(OutEntry) * (size_t)L.otab;
    unsigned *voff = reinterpret_cast<unsigned *>(smem_raw + off); off += (sizeof(unsigned) * (size_t)L.zcap + 15) & ~size_t(15);
    unsigned char *Zraw = smem_raw + off;

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
    __syncthreads();
    build_axis_tables(S, g);
    build_y_tiles(S, g, L.zcap);

    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    const int nx = X.n;
    const int cl = L.cl;
    const int lane = threadIdx.x % cl, slot = threadIdx.x / cl, vs = PL_THREADS / cl;
    const int c4 = chunk * cl + lane;                       // float4 channel group
    const bool lane_on = c4 < g.C / 4;
    const unsigned sW = (unsigned)g.D * g.C, sH = (unsigned)g.W * g.D * g.C;
    const float *img = image + (long long)__ldg(box_index + b) * g.H * sH + c4 * 4;
    float *crop = crops + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float4 ext4 = make_float4(ext, ext, ext, ext);
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);
    float4 *Zst = reinterpret_cast<float4 *>(Zraw) + lane;   // stage-A store base
    const unsigned char *Zld = Zraw + lane * 16;               // stage-B load base
    const unsigned ebytes = (unsigned)cl * 16;                 // bytes per plane voxel
    const long long ostride = (long long)vs * g.pd * g.C;

    for (int tl = 0; tl < S.ntiles; ++tl) {
        const int ya = S.tile_y0[tl], yb = S.tile_y0[tl + 1];
        const int r0 = S.tile_r0[tl], r1 = S.tile_r1[tl];
        const int nvox = (r1 >= r0) ? (r1 - r0 + 1) * nx : 0;
        const int nout = (yb - ya) * g.pw;
        // ---- per-tile tables ----------------------------------------------------------
        for (int idx = threadIdx.x; idx < nvox; idx += PL_THREADS) {
            const int r = idx / nx, cx = idx - r * nx;
            voff[idx] = (unsigned)Y.list[r0 + r] * sH + (unsigned)X.list[cx] * sW;
        }
        for (int idx = threadIdx.x; idx < nout; idx += PL_THREADS) {
            const int yy = idx / g.pw, x = idx - yy * g.pw, y = ya + yy;
            const int py0 = Y.pos0[y], px0 = X.pos0[x];
            OutEntry e;
            if (py0 < 0 || px0 < 0) {
                e.o_top = 0xFFFFFFFFu; e.o_bot = 0xFFFFFFFFu; e.xl = 0.f; e.yl = 0.f;
            } else {
                const unsigned rt = (unsigned)(py0 - r0) * nx, rb = (unsigned)(Y.pos1[y] - r0) * nx;
                const unsigned px1 = (unsigned)X.pos1[x];
                e.o_top = ((rt + px0) * ebytes) | (((rt + px1) * ebytes) << 16);
                e.o_bot = ((rb + px0) * ebytes) | (((rb + px1) * ebytes) << 16);
                e.xl = X.t[x]; e.yl = Y.t[y];
            }
            otab[idx] = e;
        }
        __syncthreads();

        for (int k = k0; k < k1; ++k) {
            const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
            float *o = crop + (((long long)ya * g.pw + slot) * g.pd + k) * g.C;
            if (axis_invalid(in_z, g.D)) {                   // uniform: the whole depth sample extrapolates
                if (lane_on)
                    for (int idx = slot; idx < nout; idx += vs, o += ostride) st_stream4(o, ext4);
                continue;
            }
            const float zfl = floorf(in_z);
            const unsigned zf = (unsigned)(int)zfl * g.C, zc = (unsigned)(int)ceilf(in_z) * g.C;
            const float zl = __fsub_rn(in_z, zfl);
            // ---- stage A: footprint voxels -> z-lerp -> shared plane ----------------
            if (lane_on) {
                for (int base = slot; base < nvox; base += vs * PL_UNROLL) {
                    float4 f[PL_UNROLL], c[PL_UNROLL];
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u) {
                        const int idx = base + u * vs;
                        if (idx < nvox) {
                            const float *p = img + voff[idx];
                            f[u] = ldg4(p + zf);
                            c[u] = ldg4(p + zc);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u) {
                        const int idx = base + u * vs;
                        if (idx < nvox) Zst[idx * cl] = lerp_rn(f[u], c[u], zl);
                    }
                }
            }
            __syncthreads();
            // ---- stage B: x-lerp, y-lerp from the plane -> crops ----------------------
            if (lane_on) {
#pragma unroll 2
                for (int idx = slot; idx < nout; idx += vs, o += ostride) {
                    const uint4 e = *reinterpret_cast<const uint4 *>(&otab[idx]);
                    if (e.x == 0xFFFFFFFFu) { st_stream4(o, ext4); continue; }
                    const float xl = __uint_as_float(e.z), yl = __uint_as_float(e.w);
                    const float4 tlv = *reinterpret_cast<const float4 *>(Zld + (e.x & 0xFFFFu));
                    const float4 trv = *reinterpret_cast<const float4 *>(Zld + (e.x >> 16));
                    const float4 blv = *reinterpret_cast<const float4 *>(Zld + (e.y & 0xFFFFu));
                    const float4 brv = *reinterpret_cast<const float4 *>(Zld + (e.y >> 16));
                    const float4 top = lerp_rn(tlv, trv, xl), bot = lerp_rn(blv, brv, xl);
                    st_stream4(o, lerp_rn(top, bot, yl));
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------
// backward (grad image).  grad_image must be zero-filled before the launch.
// ---------------------------------------------------------------------------------
struct BwdRanges {                          // per footprint position: the samples that tap it
    short a0[2 * PL_MAXP], e0[2 * PL_MAXP]; // [a0,e0): samples whose floor is this position
    short a1[2 * PL_MAXP], e1[2 * PL_MAXP]; // [a1,e1): samples whose ceil is this position
};

__device__ __forceinline__ void build_ranges(const AxisTab &T, BwdRanges &R, int p, int tid0, int nthreads)
{
    // samples with a given floor (ceil) position are contiguous because `in` is monotonic in the sample index
    for (int q = tid0; q < T.n; q += nthreads) {
        int a0 = p, e0 = 0, a1 = p, e1 = 0;
        for (int s = 0; s < p; ++s) {
            if (T.pos0[s] == q) { a0 = min(a0, s); e0 = max(e0, s + 1); }
            if (T.pos1[s] == q) { a1 = min(a1, s); e1 = max(e1, s + 1); }
        }
        if (a0 >= e0) a0 = e0 = 0;
        if (a1 >= e1) a1 = e1 = 0;
        R.a0[q] = (short)a0; R.e0[q] = (short)e0;
        R.a1[q] = (short)a1; R.e1[q] = (short)e1;
    }
}

__device__ __forceinline__ void fma4(float4 &acc, const float4 a, float w) {
    acc.x = __fmaf_rn(a.x, w, acc.x); acc.y = __fmaf_rn(a.y, w, acc.y);
    acc.z = __fmaf_rn(a.z, w, acc.z); acc.w = __fmaf_rn(a.w, w, acc.w);
}

__global__ void __launch_bounds__(PL_THREADS, 4)
car3d_grad_image_plane_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                              const int *__restrict__ box_ind, CarGeom g, PlaneLaunch L,
                              float *__restrict__ grad_image)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);  off += (sizeof(PlaneShared) + 15) & ~size_t(15);
    BwdRanges &RY = *reinterpret_cast<BwdRanges *>(smem_raw + off); off += (sizeof(BwdRanges) + 15) & ~size_t(15);
    BwdRanges &RX = *reinterpret_cast<BwdRanges *>(smem_raw + off); off += (sizeof(BwdRanges) + 15) & ~size_t(15);
    float4 *G = reinterpret_cast<float4 *>(smem_raw + off);          // staged grads slice [(y-ya)*pw + x][cl]

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
    __syncthreads();
    build_axis_tables(S, g);
    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    build_ranges(Y, RY, g.ph, threadIdx.x, PL_THREADS);
    build_ranges(X, RX, g.pw, threadIdx.x, PL_THREADS);
    __syncthreads();

    const int nx = X.n, ny = Y.n;
    if (nx == 0 || ny == 0) return;                           // no in-range sample: nothing to scatter
    const int cl = L.cl;
    const int lane = threadIdx.x % cl, slot = threadIdx.x / cl, vs = PL_THREADS / cl;
    const int c4 = chunk * cl + lane;
    const bool lane_on = c4 < g.C / 4;
    const long long sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    float *img = grad_image + (long long)__ldg(box_ind + b) * g.H * sH + c4 * 4;
    const float *gcrop = grads + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);
    const int ty = max(1, L.zcap / g.pw);                    // y samples per tile
    const long long gstride = (long long)vs * g.pd * g.C;
    const int nvox = ny * nx;
    const float rnx = 1.0f / (float)nx;

    for (int ya = 0; ya < g.ph; ya += ty) {
        const int yb = min(g.ph, ya + ty);
        const int nent = (yb - ya) * g.pw;
        for (int k = k0; k < k1; ++k) {
            const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
            if (axis_invalid(in_z, g.D)) continue;             // uniform across the CTA
            const float zfl = floorf(in_z);
            const long long zf = (long long)(int)zfl * g.C, zc = (long long)(int)ceilf(in_z) * g.C;
            const float zl = __fsub_rn(in_z, zfl), wzf = __fsub_rn(1.0f, zl);
            // ---- stage A': stage the k-slice of grads (each element read once) -------------
            if (lane_on) {
                const float *gp = gcrop + (((long long)ya * g.pw + slot) * g.pd + k) * g.C;
                for (int base = slot; base < nent; base += vs * PL_UNROLL) {
                    float4 v[PL_UNROLL];
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u)
                        if (base + u * vs < nent) v[u] = ldg4(gp + u * gstride);
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u)
                        if (base + u * vs < nent) G[(base + u * vs) * cl + lane] = v[u];
                    gp += PL_UNROLL * gstride;
                }
            }
            __syncthreads();
            // ---- stage B': every footprint voxel gathers its weighted sum, then 2 REDs ------
            if (lane_on) {
                for (int idx = slot; idx < nvox; idx += vs) {
                    const int r = (int)(((float)idx + 0.5f) * rnx), cx = idx - r * nx;
                    const int ylo0 = max((int)RY.a0[r], ya), yhi0 = min((int)RY.e0[r], yb);
                    const int ylo1 = max((int)RY.a1[r], ya), yhi1 = min((int)RY.e1[r], yb);
                    if (ylo0 >= yhi0 && ylo1 >= yhi1) continue;
                    const int xa0 = RX.a0[cx], xe0 = RX.e0[cx], xa1 = RX.a1[cx], xe1 = RX.e1[cx];
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int pass = 0; pass < 2; ++pass) {
                        const int ylo = pass ? ylo1 : ylo0, yhi = pass ? yhi1 : yhi0;
                        for (int y = ylo; y < yhi; ++y) {
                            const float wy = pass ? Y.t[y] : __fsub_rn(1.0f, Y.t[y]);
                            const float4 *row = G + ((y - ya) * g.pw) * cl + lane;
                            for (int x = xa0; x < xe0; ++x) fma4(acc, row[x * cl], __fmul_rn(wy, __fsub_rn(1.0f, X.t[x])));
                            for (int x = xa1; x < xe1; ++x) fma4(acc, row[x * cl], __fmul_rn(wy, X.t[x]));
                        }
                    }
                    float *p = img + Y.list[r] * sH + X.list[cx] * sW;
                    red_add4(p + zf, make_float4(acc.x * wzf, acc.y * wzf, acc.z * wzf, acc.w * wzf));
                    red_add4(p + zc, make_float4(acc.x * zl, acc.y * zl, acc.z * zl, acc.w * zl));
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
static inline size_t a16(size_t v) { return (v + 15) & ~size_t(15); }

static int pick_cl(const CarGeom &g, int pref_cl) {
    int cl = pref_cl;
    while (cl > 1 && cl / 2 >= g.C / 4) cl /= 2;               // do not idle half the lanes
    return cl;
}

static int pick_ksplits(const CarGeom &g, int chunks) {
    const long long want = (long long)kNumSMs * 16;
    const long long per = (long long)g.n * chunks;
    long long ks = (want + per - 1) / per;
    if (ks < 1) ks = 1;
    if (ks > g.pd) ks = g.pd;
    return (int)ks;
}

int launch_car3d_fwd_plane(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                           float ext, float *crops, cudaStream_t stream)
{
    PlaneLaunch L;
    L.cl = pick_cl(g, 16);
    const int nxmax = min(2 * g.pw, g.W);
    L.otab = min(g.ph * g.pw, max(PL_MAXOUT, g.pw));
    size_t smem;
    for (;;) {
        L.zcap = max(2 * nxmax, (40 * 1024) / (L.cl * 16));
        // plane byte offsets are packed into 16 bits
        while ((size_t)L.zcap * L.cl * 16 > 65535 && L.zcap > 2 * nxmax) --L.zcap;
        smem = a16(sizeof(PlaneShared)) + sizeof(OutEntry) * (size_t)L.otab + a16(sizeof(unsigned) * (size_t)L.zcap) +
               (size_t)L.zcap * L.cl * 16;
        if (((size_t)L.zcap * L.cl * 16 <= 65535 && smem <= 200 * 1024) || L.cl == 1) break;
        L.cl /= 2;
    }
    if ((size_t)L.zcap * L.cl * 16 > 65535 || smem > 200 * 1024) return ROI3D_EUNSUPPORTED;
    L.chunks = (g.C / 4 + L.cl - 1) / L.cl;
    L.ksplits = pick_ksplits(g, L.chunks);
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(cudaFuncSetAttribute(car3d_fwd_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    car3d_fwd_plane_kernel<<<(unsigned)grid, PL_THREADS, smem, stream>>>(image, boxes, box_index, g, L, ext, crops);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_car3d_grad_image_plane(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                                  float *grad_image, cudaStream_t stream)
{
    PlaneLaunch L;
    L.cl = pick_cl(g, 16);
    L.otab = 0;
    const size_t fixed = a16(sizeof(PlaneShared)) + 2 * a16(sizeof(BwdRanges));
    size_t smem;
    for (;;) {
        // stage whole k-slices when they fit ~50 KB, else tile over y samples
        const int want = min(g.ph * g.pw, max(g.pw, (50 * 1024) / (L.cl * 16)));
        L.zcap = max(want, g.pw);
        smem = fixed + (size_t)L.zcap * L.cl * 16;
        if (smem <= 200 * 1024 || L.cl == 1) break;
        L.cl /= 2;
    }
    if (smem > 200 * 1024) return ROI3D_EUNSUPPORTED;
    L.chunks = (g.C / 4 + L.cl - 1) / L.cl;
    L.ksplits = pick_ksplits(g, L.chunks);
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(cudaFuncSetAttribute(car3d_grad_image_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    car3d_grad_image_plane_kernel<<<(unsigned)grid, PL_THREADS, smem, stream>>>(grads, boxes, box_ind, g, L, grad_image);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

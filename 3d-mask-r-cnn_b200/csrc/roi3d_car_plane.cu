// roi3d_car_plane.cu -- plane-staged separable CropAndResize3D (variant 2), the
// production kernels for trilinear sampling with C % 4 == 0.
//
// Why: trilinear taps form a product set.  For one ROI and one depth sample k the
// reference (CAR.so@0x4f88-0x5021) first lerps along z, then x, then y, each as
// a + (b - a) * t.  The z-lerp of a voxel column (yi, xi) depends only on (yi, xi, k),
// so a CTA that owns (ROI, k-range, channel chunk)
//   A. reads every footprint voxel (yi, xi, {floor z, ceil z}) exactly ONCE from global
//      memory (coalesced 16-byte loads along channels), z-lerps it in registers and
//      parks the result in a shared-memory plane Z[row][col][chunk];
//   B. produces all ph*pw outputs of that k from 4 shared-memory reads each (x-lerp,
//      then y-lerp) and streams them out with 16-byte evict-first stores.
// The intermediate values are the very expressions the reference evaluates, in the same
// order and rounding, so the result is bit-identical to the direct kernel and to the
// oracle, while L2->SM traffic drops from 8 taps per output to 2 * footprint / outputs.
//
// Instruction economy (profiles/r1a, r1b: the first versions issued 190 / 134 warp
// instructions per 16-byte output and were issue-bound): everything that depends only on
// (ROI, y-tile) -- footprint voxel offsets, and per output the four plane offsets plus the
// x/y lerp weights -- is tabulated once per CTA in shared memory; each thread carries V
// float4 channel groups through every index computation; shared memory is addressed with
// 32-bit shared-window addresses; out-of-range outputs are selected, not branched.
//
// The backward kernel is the transpose: the k-slice of grads is staged in shared memory
// (each element read once, coalesced), every footprint voxel gathers its (wy*wx)-weighted
// sum over per-axis contribution lists (deterministic order, no shared-memory atomics)
// and issues two vector REDs (floor z, ceil z) instead of 8 per grads element.
#include "roi3d_common.cuh"
#include "roi3d_car_pyr.cuh"
#include <cuda.h>                            // CUtensorMap (the encode function is fetched from the driver at run time)
#include <type_traits>
#include <math.h>
#include <string.h>

namespace roi3d {

constexpr int PL_THREADS = 256;
constexpr int PL_MAXP = 64;                 // max crop size per axis handled here
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
    {   // dedupe: candidate j is a "first" if no earlier candidate has its value; list = firsts in order
        const bool act = tid < 4 * PL_MAXP;                   // (CTAs of more than 256 threads: the other warps only join the barriers)
        const int a = act ? tid / (2 * PL_MAXP) : 0, j = tid % (2 * PL_MAXP);
        const int p = a ? g.pw : g.ph;
        AxisTab &T = S.ax[a];
        const int v = T.cand[j];
        const bool fast = 2 * p <= 32;                        // the axis' candidates fit one warp (crops up to 16)
        int fo = j;
        if (!act) {
        } else if (fast) {
            if (j < 32) {                                      // warp 0 (y) / warp 4 (x): match instead of searching
                const bool valid = j < 2 * p && v >= 0;
                const unsigned m = __match_any_sync(0xffffffffu, valid ? v : -1 - j);
                fo = __ffs(m) - 1;
                const bool isfirst = valid && fo == j;
                const unsigned fb = __ballot_sync(0xffffffffu, isfirst);
                if (j < 2 * p) {
                    const int pos = valid ? __popc(fb & ((1u << fo) - 1u)) : -1;
                    if (isfirst) T.list[pos] = v;
                    if (j & 1) T.pos1[j >> 1] = (short)pos; else T.pos0[j >> 1] = (short)pos;
                }
                if (j == 0) T.n = __popc(fb);
            }
        } else {
            if (j < 2 * p && v >= 0) {
                for (int i = 0; i < j; ++i)
                    if (T.cand[i] == v) { fo = i; break; }
            }
            T.first[j] = (j < 2 * p && v >= 0 && fo == j) ? 1 : 0;
        }
        __syncthreads();
        if (act && !fast) {
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
        if (Y.n * nx <= zcap && g.ph * g.pw <= PL_MAXOUT) {   // common case: the whole crop is one tile
            S.tile_r0[0] = 0; S.tile_r1[0] = (short)(Y.n - 1);
            S.tile_y0[1] = (short)g.ph;
            S.ntiles = 1;
        } else {
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
    }
    __syncthreads();
}

struct PlaneLaunch {
    int cl;            // channel lanes per voxel (each lane carries V float4 groups)
    int chunks;        // channel chunks per voxel
    int ksplits;       // depth-sample splits per ROI
    int zcap;          // plane capacity in voxels (forward) / staged grads entries (backward)
    int otab;          // output-table entries (forward)
    const int *perm;   // processing order of the ROIs (CTA group i works on ROI perm[i]); nullptr: as given
};

struct __align__(16) OutEntry {             // per output (y, x) of the current y-tile
    unsigned o_top;                         // byte offsets into the plane: (t,l) | (t,r) << 16
    unsigned o_bot;                         //                              (b,l) | (b,r) << 16
    float xl, yl;
};


// ---------------------------------------------------------------------------------
// forward.  V = float4 channel groups per thread (the chunk is cl * V * 4 channels).
// ---------------------------------------------------------------------------------
// FULL: every channel lane of every chunk carries live channels ((C / 4) % (cl * V) == 0, e.g. C = 256): the per-group
// predicates, and the divergence bookkeeping (BSSY / BSYNC) they drag into every inner loop, compile away.
// NT: threads per CTA (256; 512 = experiment: 32 voxel / output slots per pass instead of 16, two CTAs per SM)
template <int V, bool PYR, bool HALF = false, bool FULL = false, int NT = PL_THREADS>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : ((V == 1) ? 4 : 3))
car3d_fwd_plane_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                       const int *__restrict__ box_index, CarGeom g, PlaneLaunch L, float ext,
                       void *__restrict__ crops, const PyrParams P)
{
    using OutT = typename std::conditional<HALF, __half, float>::type;    // HALF: the target files' float16 payload
    constexpr int UNR = 1;                                     // stage-A voxels in flight per thread (x V x 2 taps);
                                                               // measured: less unrolling beats more here (72 regs, no spills)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);  off += (sizeof(PlaneShared) + 15) & ~size_t(15);
    OutEntry *otab = reinterpret_cast<OutEntry *>(smem_raw + off); off += sizeof(OutEntry) * (size_t)L.otab;
    unsigned *voff = reinterpret_cast<unsigned *>(smem_raw + off); off += (sizeof(unsigned) * (size_t)L.zcap + 15) & ~size_t(15);
    unsigned char *Zraw = smem_raw + off;

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = L.perm ? __ldg(L.perm + bid / L.ksplits) : bid / L.ksplits;
    __shared__ int s_level;
    if constexpr (PYR) {
        if (threadIdx.x == 0) {
            float b6[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) b6[q] = __ldg(boxes + (size_t)b * 6 + q);
            const PyrRoute r = pyr_route(b6, P);
#pragma unroll
            for (int q = 0; q < 6; ++q) S.box[q] = r.box[q];
            s_level = r.level - 2;
        }
        __syncthreads();
        const int lv = s_level;
        const float *lim;
        pyr_level(P, lv, g.H, g.W, g.D, lim);
        image = lim;
    } else {
        if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
        __syncthreads();
    }
    build_axis_tables(S, g);
    build_y_tiles(S, g, L.zcap);

    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    const int nx = X.n;
    const int cl = L.cl;
    const int lane = threadIdx.x % cl, slot = threadIdx.x / cl, vs = NT / cl;
    const int c4 = chunk * cl * V + lane;                   // first float4 channel group of this thread
    bool von[V];
#pragma unroll
    for (int v = 0; v < V; ++v) von[v] = FULL || (c4 + v * cl) < g.C / 4;
    const unsigned sW = (unsigned)g.D * g.C, sH = (unsigned)g.W * g.D * g.C;
    const int bimg = PYR ? b / P.rois_per_image : __ldg(box_index + b);
    const bool bad_img = (unsigned)bimg >= (unsigned)g.B;     // out-of-range box_index: the whole crop extrapolates, nothing is read
    const float *img = image + (long long)bimg * g.H * sH + c4 * 4;
    OutT *crop = static_cast<OutT *>(crops) + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float4 ext4 = make_float4(ext, ext, ext, ext);
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);
    const unsigned ebytes = (unsigned)cl * V * 16;             // bytes per plane voxel
    const unsigned z_u32 = smem_u32(Zraw) + lane * 16;         // this lane's column of the plane
    const unsigned otab_u32 = smem_u32(otab), voff_u32 = smem_u32(voff);
    const long long ostride = (long long)vs * g.pd * g.C;
    const int vstep = cl * 4;                                  // floats between a thread's channel groups

    for (int tl = 0; tl < S.ntiles; ++tl) {
        const int ya = S.tile_y0[tl], yb = S.tile_y0[tl + 1];
        const int r0 = S.tile_r0[tl], r1 = S.tile_r1[tl];
        const int nvox = (r1 >= r0) ? (r1 - r0 + 1) * nx : 0;
        const int nout = (yb - ya) * g.pw;
        // ---- per-tile tables ----------------------------------------------------------
        for (int idx = threadIdx.x; idx < nvox; idx += NT) {
            const int r = idx / nx, cx = idx - r * nx;
            voff[idx] = (unsigned)Y.list[r0 + r] * sH + (unsigned)X.list[cx] * sW;
        }
        int my_bad = 0;
        for (int idx = threadIdx.x; idx < nout; idx += NT) {
            const int yy = idx / g.pw, x = idx - yy * g.pw, y = ya + yy;
            const int py0 = Y.pos0[y], px0 = X.pos0[x];
            OutEntry e;
            if (py0 < 0 || px0 < 0) {
                e.o_top = 0xFFFFFFFFu; e.o_bot = 0xFFFFFFFFu; e.xl = 0.f; e.yl = 0.f;
                my_bad = 1;
            } else {
                const unsigned rt = (unsigned)(py0 - r0) * nx, rb = (unsigned)(Y.pos1[y] - r0) * nx;
                const unsigned px1 = (unsigned)X.pos1[x];
                e.o_top = ((rt + px0) * ebytes) | (((rt + px1) * ebytes) << 16);
                e.o_bot = ((rb + px0) * ebytes) | (((rb + px1) * ebytes) << 16);
                e.xl = X.t[x]; e.yl = Y.t[y];
            }
            otab[idx] = e;
        }
        // the barrier doubles as the vote: tiles whose samples all lie inside the volume (the usual case) run
        // stage B without the per-output extrapolation selects
        const bool tile_has_bad = __syncthreads_or(my_bad) != 0;

        for (int k = k0; k < k1; ++k) {
            const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
            OutT *o = crop + (((long long)ya * g.pw + slot) * g.pd + k) * g.C;
            if (axis_invalid(in_z, g.D) || nvox == 0 || bad_img) {   // uniform: every output of this (tile, k) extrapolates
                for (int idx = slot; idx < nout; idx += vs, o += ostride) {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (von[v]) st_stream4(o + v * vstep, ext4);
                }
                continue;
            }
            const float zfl = floorf(in_z);
            const unsigned zf = (unsigned)(int)zfl * g.C, zc = (unsigned)(int)ceilf(in_z) * g.C;
            const float zl = __fsub_rn(in_z, zfl);
            // ---- stage A: footprint voxels -> z-lerp -> shared plane ----------------
            for (int base = slot; base < nvox; base += vs * UNR) {
                float4 f[UNR][V], c[UNR][V];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int idx = min(base + u * vs, nvox - 1);      // clamp: tail lanes re-read a valid voxel
                    const float *p = img + lds32u(voff_u32 + idx * 4);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (von[v]) {
                            f[u][v] = ldg4(p + zf + v * vstep);
                            c[u][v] = ldg4(p + zc + v * vstep);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int idx = base + u * vs;
                    if (idx < nvox) {
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            if (von[v]) sts128(z_u32 + idx * ebytes + v * (cl * 16), lerp_rn(f[u][v], c[u][v], zl));
                    }
                }
            }
            __syncthreads();
            // ---- L2 prefetch of the next depth sample's footprint: its stage A then finds the taps in L2 ----
            if (k + 1 < k1 && (lane & 7) == 0) {
                const float in_zn = axis_coord(z1, z2, g.D, g.pd, k + 1, zscale);
                if (!axis_invalid(in_zn, g.D)) {
                    const unsigned zfn = (unsigned)(int)floorf(in_zn) * g.C, zcn = (unsigned)(int)ceilf(in_zn) * g.C;
                    for (int idx = slot; idx < nvox; idx += vs) {
                        const float *p = img + lds32u(voff_u32 + idx * 4);
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if (von[v]) {
                                if (zfn != zf && zfn != zc) prefetch_l2(p + zfn + v * vstep);
                                if (zcn != zf && zcn != zc) prefetch_l2(p + zcn + v * vstep);
                            }
                        }
                    }
                }
            }
            // ---- stage B: x-lerp, y-lerp from the plane -> crops ----------------------
            auto stage_b = [&](auto has_bad_t) {
                constexpr bool HAS_BAD = decltype(has_bad_t)::value;
#pragma unroll 1
                for (int idx = slot; idx < nout; idx += vs, o += ostride) {
                    const uint4 e = lds128u(otab_u32 + idx * 16);
                    const bool bad = HAS_BAD && e.x == 0xFFFFFFFFu;
                    const unsigned et = bad ? 0u : e.x, eb = bad ? 0u : e.y;
                    const float xl = __uint_as_float(e.z), yl = __uint_as_float(e.w);
                    const unsigned a_tl = z_u32 + (et & 0xFFFFu), a_tr = z_u32 + (et >> 16);
                    const unsigned a_bl = z_u32 + (eb & 0xFFFFu), a_br = z_u32 + (eb >> 16);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (von[v]) {
                            const unsigned vo = v * (cl * 16);
                            const float4 tlv = lds128(a_tl + vo), trv = lds128(a_tr + vo);
                            const float4 blv = lds128(a_bl + vo), brv = lds128(a_br + vo);
                            const float4 top = lerp_rn(tlv, trv, xl), bot = lerp_rn(blv, brv, xl);
                            float4 res = lerp_rn(top, bot, yl);
                            if constexpr (HAS_BAD) res = sel4(bad, ext4, res);
                            if constexpr (PYR) res = scrub4(res);
                            st_stream4(o + v * vstep, res);
                        }
                    }
                }
            };
            if (tile_has_bad) stage_b(std::true_type{}); else stage_b(std::false_type{});
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------
// forward, TMA-fed (variant 3).  Same tables and arithmetic; stage A's gather is done by the TMA engine: one warp
// issues a bulk copy (cp.async.bulk, 256-byte channel-chunk row per footprint voxel and z tap) for the NEXT depth
// sample into a raw buffer right after the current sample has been z-lerped out of it, so the copies fly while all
// eight warps run stage B.  No registers and no LSU issue slots are spent on the gather; completion is an mbarrier.
// ---------------------------------------------------------------------------------
template <bool PYR>
__global__ void __launch_bounds__(PL_THREADS, 3)
car3d_fwd_plane_tma_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                           const int *__restrict__ box_index, CarGeom g, PlaneLaunch L, float ext,
                           float *__restrict__ crops, const PyrParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);  off += (sizeof(PlaneShared) + 15) & ~size_t(15);
    OutEntry *otab = reinterpret_cast<OutEntry *>(smem_raw + off); off += sizeof(OutEntry) * (size_t)L.otab;
    unsigned *voff = reinterpret_cast<unsigned *>(smem_raw + off); off += (sizeof(unsigned) * (size_t)L.zcap + 15) & ~size_t(15);
    off = (off + 127) & ~size_t(127);
    const unsigned ebytes = (unsigned)L.cl * 16;
    unsigned char *Rf = smem_raw + off; off += (size_t)L.zcap * ebytes;      // raw floor-z taps
    unsigned char *Rc = smem_raw + off; off += (size_t)L.zcap * ebytes;      // raw ceil-z taps
    unsigned char *Zraw = smem_raw + off; off += (size_t)L.zcap * ebytes;    // z-lerped plane
    __shared__ __align__(8) unsigned long long s_mbar;

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    __shared__ int s_level;
    if constexpr (PYR) {
        if (threadIdx.x == 0) {
            float b6[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) b6[q] = __ldg(boxes + (size_t)b * 6 + q);
            const PyrRoute r = pyr_route(b6, P);
#pragma unroll
            for (int q = 0; q < 6; ++q) S.box[q] = r.box[q];
            s_level = r.level - 2;
        }
        __syncthreads();
        const int lv = s_level;
        const float *lim;
        pyr_level(P, lv, g.H, g.W, g.D, lim);
        image = lim;
    } else {
        if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
        __syncthreads();
    }
    const unsigned mbar = smem_u32(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    build_axis_tables(S, g);
    build_y_tiles(S, g, L.zcap);

    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    const int nx = X.n;
    const int cl = L.cl;
    const int lane = threadIdx.x % cl, slot = threadIdx.x / cl, vs = PL_THREADS / cl;
    const int c4 = chunk * cl + lane;
    const bool on = c4 < g.C / 4;
    const unsigned rowbytes = (unsigned)min(cl, g.C / 4 - chunk * cl) * 16;  // bytes of this chunk actually present
    const unsigned sW = (unsigned)g.D * g.C, sH = (unsigned)g.W * g.D * g.C;
    const int bimg = PYR ? b / P.rois_per_image : __ldg(box_index + b);
    const bool bad_img = (unsigned)bimg >= (unsigned)g.B;
    const float *img_chunk = image + (long long)bimg * g.H * sH + chunk * cl * 4;   // lane-independent: TMA source
    float *crop = crops + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float4 ext4 = make_float4(ext, ext, ext, ext);
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = min(ks * kper, g.pd), k1 = min(g.pd, k0 + kper);
    const unsigned rf_u32 = smem_u32(Rf), rc_u32 = smem_u32(Rc);
    const unsigned z_u32 = smem_u32(Zraw) + lane * 16;
    const unsigned otab_u32 = smem_u32(otab), voff_u32 = smem_u32(voff);
    const long long ostride = (long long)vs * g.pd * g.C;
    unsigned parity = 0;

    for (int tl = 0; tl < S.ntiles; ++tl) {
        const int ya = S.tile_y0[tl], yb = S.tile_y0[tl + 1];
        const int r0 = S.tile_r0[tl], r1 = S.tile_r1[tl];
        const int nvox = (r1 >= r0) ? (r1 - r0 + 1) * nx : 0;
        const int nout = (yb - ya) * g.pw;
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
                const unsigned rt_ = (unsigned)(py0 - r0) * nx, rb = (unsigned)(Y.pos1[y] - r0) * nx;
                const unsigned px1 = (unsigned)X.pos1[x];
                e.o_top = ((rt_ + px0) * ebytes) | (((rt_ + px1) * ebytes) << 16);
                e.o_bot = ((rb + px0) * ebytes) | (((rb + px1) * ebytes) << 16);
                e.xl = X.t[x]; e.yl = Y.t[y];
            }
            otab[idx] = e;
        }
        __syncthreads();                                       // tables (and the mbarrier init) visible

        // warp 0 gathers depth sample k into the raw buffers with bulk copies
        auto issue = [&](int k) {
            if (threadIdx.x < 32) {
                const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
                const unsigned zf = (unsigned)(int)floorf(in_z) * g.C, zc = (unsigned)(int)ceilf(in_z) * g.C;
                fence_proxy_async();                           // earlier generic reads of the raw buffers are done
                if (threadIdx.x == 0) mbar_expect_tx(mbar, 2u * (unsigned)nvox * rowbytes);
                __syncwarp();
                for (int idx = threadIdx.x; idx < nvox; idx += 32) {
                    const float *p = img_chunk + voff[idx];
                    bulk_g2s(rf_u32 + idx * ebytes, p + zf, rowbytes, mbar);
                    bulk_g2s(rc_u32 + idx * ebytes, p + zc, rowbytes, mbar);
                }
            }
        };
        auto zvalid = [&](int k) { return k < k1 && nvox > 0 && !bad_img && !axis_invalid(axis_coord(z1, z2, g.D, g.pd, k, zscale), g.D); };

        int kfirst = k0;
        while (kfirst < k1 && !zvalid(kfirst)) ++kfirst;
        if (kfirst < k1) issue(kfirst);

        for (int k = k0; k < k1; ++k) {
            float *o = crop + (((long long)ya * g.pw + slot) * g.pd + k) * g.C;
            if (!zvalid(k)) {
                if (on)
                    for (int idx = slot; idx < nout; idx += vs, o += ostride) st_stream4(o, ext4);
                continue;
            }
            const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
            const float zl = __fsub_rn(in_z, floorf(in_z));
            mbar_wait(mbar, parity);                           // the taps of sample k have landed
            parity ^= 1u;
            if (on) {
                for (int idx = slot; idx < nvox; idx += vs) {
                    const float4 f = lds128(rf_u32 + idx * ebytes + lane * 16);
                    const float4 c = lds128(rc_u32 + idx * ebytes + lane * 16);
                    sts128(z_u32 + idx * ebytes, lerp_rn(f, c, zl));
                }
            }
            __syncthreads();                                   // plane complete, raw buffers free
            if (zvalid(k + 1)) issue(k + 1);                   // flies during stage B
            if (on) {
#pragma unroll 2
                for (int idx = slot; idx < nout; idx += vs, o += ostride) {
                    const uint4 e = lds128u(otab_u32 + idx * 16);
                    const bool bad = e.x == 0xFFFFFFFFu;
                    const unsigned et = bad ? 0u : e.x, eb = bad ? 0u : e.y;
                    const float xl = __uint_as_float(e.z), yl = __uint_as_float(e.w);
                    const float4 tlv = lds128(z_u32 + (et & 0xFFFFu)), trv = lds128(z_u32 + (et >> 16));
                    const float4 blv = lds128(z_u32 + (eb & 0xFFFFu)), brv = lds128(z_u32 + (eb >> 16));
                    const float4 top = lerp_rn(tlv, trv, xl), bot = lerp_rn(blv, brv, xl);
                    float4 res = sel4(bad, ext4, lerp_rn(top, bot, yl));
                    if constexpr (PYR) res = scrub4(res);
                    st_stream4(o, res);
                }
            }
            __syncthreads();                                   // plane free for the next sample
        }
    }
}

// ---------------------------------------------------------------------------------
// forward, fed by the Blackwell TMA row gather (variant 5).  Same structure as variant 3, but stage A's gather is
// cp.async.bulk.tensor.2d ... tile::gather4 (SASS UTMALDG with the gather4 modifier) on a 2-D tensor map
// [B*H*W*D rows, C columns] with a box of one channel chunk x 1 row: ONE instruction brings the chunk rows of FOUR
// footprint voxels (4 x 256 B), so a depth sample of ~81 voxels x 2 taps takes 42 copies instead of 162 -- the
// granularity problem of variant 3 (profiles/README.md).  voff[] holds row indices here, not element offsets.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_gather4(unsigned dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(mbar) : "memory");
}

// wait with a bound: a transaction-count mismatch (the only way this can block) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(unsigned mbar, unsigned parity) {
#pragma unroll 1
    for (unsigned spin = 0; spin < (1u << 24); ++spin) {     // try_wait suspends the thread in hardware for a while: not a hot spin
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

__global__ void __launch_bounds__(PL_THREADS, 3)
car3d_fwd_plane_g4_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ boxes,
                          const int *__restrict__ box_index, CarGeom g, PlaneLaunch L, float ext,
                          float *__restrict__ crops)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);  off += (sizeof(PlaneShared) + 15) & ~size_t(15);
    OutEntry *otab = reinterpret_cast<OutEntry *>(smem_raw + off); off += sizeof(OutEntry) * (size_t)L.otab;
    int *vrow = reinterpret_cast<int *>(smem_raw + off); off += (sizeof(int) * (size_t)L.zcap + 15) & ~size_t(15);
    off += (128u - ((smem_u32(smem_raw) + (unsigned)off) & 127u)) & 127u;     // tensor copies need 128-byte aligned destinations
    const unsigned ebytes = (unsigned)L.cl * 16;
    unsigned char *Rf = smem_raw + off; off += (size_t)L.zcap * ebytes;      // raw floor-z taps
    unsigned char *Rc = smem_raw + off; off += (size_t)L.zcap * ebytes;      // raw ceil-z taps
    unsigned char *Zraw = smem_raw + off; off += (size_t)L.zcap * ebytes;    // z-lerped plane
    __shared__ __align__(8) unsigned long long s_mbar;

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
    __syncthreads();
    const unsigned mbar = smem_u32(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    build_axis_tables(S, g);
    build_y_tiles(S, g, L.zcap);

    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    const int nx = X.n;
    const int cl = L.cl;
    const int lane = threadIdx.x % cl, slot = threadIdx.x / cl, vs = PL_THREADS / cl;
    const int c4 = chunk * cl + lane;
    const bool on = c4 < g.C / 4;
    const int bimg = __ldg(box_index + b);
    const bool bad_img = (unsigned)bimg >= (unsigned)g.B;
    float *crop = crops + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float4 ext4 = make_float4(ext, ext, ext, ext);
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = min(ks * kper, g.pd), k1 = min(g.pd, k0 + kper);
    const unsigned rf_u32 = smem_u32(Rf), rc_u32 = smem_u32(Rc);
    const unsigned z_u32 = smem_u32(Zraw) + lane * 16;
    const unsigned otab_u32 = smem_u32(otab);
    const long long ostride = (long long)vs * g.pd * g.C;
    const int col0 = chunk * cl * 4;                              // first channel of this chunk (tensor-map column)
    unsigned parity = 0;

    for (int tl = 0; tl < S.ntiles; ++tl) {
        const int ya = S.tile_y0[tl], yb = S.tile_y0[tl + 1];
        const int r0 = S.tile_r0[tl], r1 = S.tile_r1[tl];
        const int nvox = (r1 >= r0) ? (r1 - r0 + 1) * nx : 0;
        const int nout = (yb - ya) * g.pw;
        for (int idx = threadIdx.x; idx < nvox; idx += PL_THREADS) {
            const int r = idx / nx, cx = idx - r * nx;
            vrow[idx] = ((bimg * g.H + Y.list[r0 + r]) * g.W + X.list[cx]) * g.D;      // tensor-map row of (y, x, z = 0)
        }
        for (int idx = threadIdx.x; idx < nout; idx += PL_THREADS) {
            const int yy = idx / g.pw, x = idx - yy * g.pw, y = ya + yy;
            const int py0 = Y.pos0[y], px0 = X.pos0[x];
            OutEntry e;
            if (py0 < 0 || px0 < 0) {
                e.o_top = 0xFFFFFFFFu; e.o_bot = 0xFFFFFFFFu; e.xl = 0.f; e.yl = 0.f;
            } else {
                const unsigned rt_ = (unsigned)(py0 - r0) * nx, rb = (unsigned)(Y.pos1[y] - r0) * nx;
                const unsigned px1 = (unsigned)X.pos1[x];
                e.o_top = ((rt_ + px0) * ebytes) | (((rt_ + px1) * ebytes) << 16);
                e.o_bot = ((rb + px0) * ebytes) | (((rb + px1) * ebytes) << 16);
                e.xl = X.t[x]; e.yl = Y.t[y];
            }
            otab[idx] = e;
        }
        __syncthreads();                                       // tables (and the mbarrier init) visible

        // warp 0 gathers depth sample k into the raw buffers: one gather4 per four voxels and tap
        auto issue = [&](int k) {
            if (threadIdx.x < 64) {                            // warp 0: floor taps, warp 1: ceil taps
                const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
                const int tap = threadIdx.x >> 5, ln = threadIdx.x & 31;
                const int zt = tap ? (int)ceilf(in_z) : (int)floorf(in_z);
                const unsigned dst = tap ? rc_u32 : rf_u32;
                const int ngroups = (nvox + 3) >> 2;
                fence_proxy_async();                           // earlier generic reads of the raw buffers are done
                // the phase cannot complete before this arrival, whatever the order of the copies' complete_tx
                if (threadIdx.x == 0) mbar_expect_tx(mbar, 2u * (unsigned)ngroups * 4u * ebytes);
                for (int gq = ln; gq < ngroups; gq += 32) {
                    const int i0 = gq * 4, last = nvox - 1;
                    const int ra = vrow[i0], rb_ = vrow[min(i0 + 1, last)], rc_ = vrow[min(i0 + 2, last)], rd = vrow[min(i0 + 3, last)];
                    tma_gather4(dst + (unsigned)i0 * ebytes, &tmap, col0, ra + zt, rb_ + zt, rc_ + zt, rd + zt, mbar);
                }
            }
        };
        auto zvalid = [&](int k) { return k < k1 && nvox > 0 && !bad_img && !axis_invalid(axis_coord(z1, z2, g.D, g.pd, k, zscale), g.D); };

        int kfirst = k0;
        while (kfirst < k1 && !zvalid(kfirst)) ++kfirst;
        if (kfirst < k1) issue(kfirst);

        for (int k = k0; k < k1; ++k) {
            float *o = crop + (((long long)ya * g.pw + slot) * g.pd + k) * g.C;
            if (!zvalid(k)) {
                if (on)
                    for (int idx = slot; idx < nout; idx += vs, o += ostride) st_stream4(o, ext4);
                continue;
            }
            const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
            const float zl = __fsub_rn(in_z, floorf(in_z));
            mbar_wait_bounded(mbar, parity);                           // the taps of sample k have landed
            parity ^= 1u;
            if (on) {
                for (int idx = slot; idx < nvox; idx += vs) {
                    const float4 f = lds128(rf_u32 + idx * ebytes + lane * 16);
                    const float4 c = lds128(rc_u32 + idx * ebytes + lane * 16);
                    sts128(z_u32 + idx * ebytes, lerp_rn(f, c, zl));
                }
            }
            __syncthreads();                                   // plane complete, raw buffers free
            int kn = k + 1;
            while (kn < k1 && !zvalid(kn)) ++kn;
            if (kn < k1) issue(kn);                            // flies during stage B
            if (on) {
#pragma unroll 2
                for (int idx = slot; idx < nout; idx += vs, o += ostride) {
                    const uint4 e = lds128u(otab_u32 + idx * 16);
                    const bool bad = e.x == 0xFFFFFFFFu;
                    const unsigned et = bad ? 0u : e.x, eb = bad ? 0u : e.y;
                    const float xl = __uint_as_float(e.z), yl = __uint_as_float(e.w);
                    const float4 tlv = lds128(z_u32 + (et & 0xFFFFu)), trv = lds128(z_u32 + (et >> 16));
                    const float4 blv = lds128(z_u32 + (eb & 0xFFFFu)), brv = lds128(z_u32 + (eb >> 16));
                    const float4 top = lerp_rn(tlv, trv, xl), bot = lerp_rn(blv, brv, xl);
                    st_stream4(o, sel4(bad, ext4, lerp_rn(top, bot, yl)));
                }
            }
            __syncthreads();                                   // plane free for the next sample
        }
    }
}

// 3-D tensor tile copy / L2 prefetch (grads viewed as [N*ph*pw rows, pd, C]): the backward's k-slice staging in ONE copy
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
// the same with an L2 eviction-priority hint: grads are read exactly once, so they should be the first lines to leave L2 --
// what must stay there are the grad-image lines the REDs keep hitting
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mbar,
                                                 unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d_hint(const CUtensorMap *map, int c0, int c1, int c2, unsigned long long pol) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile.L2::cache_hint [%0, {%1, %2, %3}], %4;"
                 :: "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 :: "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---------------------------------------------------------------------------------
// backward (grad image).  grad_image must be zero-filled before the launch.
// ---------------------------------------------------------------------------------
struct __align__(8) Contrib { unsigned off; float w; };   // staged-slice byte offset of a sample, its weight

struct BwdLists {
    short start[2][2 * PL_MAXP];            // per axis, per footprint position: first entry ...
    short cnt[2][2 * PL_MAXP];              // ... and number of entries
    short samp[2][2 * PL_MAXP];             // sample index of each entry
    Contrib ent[2][2 * PL_MAXP];            // (offset, weight) of each entry
};

// Contribution lists: footprint position q of axis a is tapped by the samples whose floor is q
// (weight 1 - t) and by those whose ceil is q (weight t); both sets are contiguous sample ranges
// because `in` is monotonic in the sample index.
__device__ __forceinline__ void build_lists(const PlaneShared &S, BwdLists &B, const CarGeom &g, unsigned ebytes)
{
    const int tid = threadIdx.x;
    const int a = tid / (2 * PL_MAXP), q = tid % (2 * PL_MAXP);            // one thread per (axis, footprint position)
    const bool mine = tid < 4 * PL_MAXP && q < S.ax[a].n;
    const int p = a ? g.pw : g.ph;
    if (mine) {
        const AxisTab &T = S.ax[a];
        int c = 0;
        for (int s = 0; s < p; ++s) c += (T.pos0[s] == q) + (T.pos1[s] == q);
        B.cnt[a][q] = (short)c;
    }
    __syncthreads();
    if (mine) {
        const AxisTab &T = S.ax[a];
        int w = 0;
        for (int r = 0; r < q; ++r) w += B.cnt[a][r];                      // exclusive prefix (a dozen entries)
        B.start[a][q] = (short)w;
        const unsigned unit = a ? ebytes : ebytes * (unsigned)g.pw;
        for (int s = 0; s < p; ++s)
            if (T.pos0[s] == q) { B.samp[a][w] = (short)s; B.ent[a][w] = Contrib{(unsigned)s * unit, __fsub_rn(1.0f, T.t[s])}; ++w; }
        for (int s = 0; s < p; ++s)
            if (T.pos1[s] == q) { B.samp[a][w] = (short)s; B.ent[a][w] = Contrib{(unsigned)s * unit, T.t[s]}; ++w; }
    }
    __syncthreads();
}

// TMA: the k-slice of grads is staged by ONE tensor tile copy (cp.async.bulk.tensor.3d, box = channel chunk x 1 depth
// sample x zcap (y, x) samples) issued by one thread and awaited on an mbarrier, instead of 2-3 dependent batches of
// per-thread loads + shared stores; the next slice is pulled towards L2 by one tensor prefetch instruction.
template <int V, bool PYR, bool FULL = false, bool TMA = false, int NT = PL_THREADS>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : ((V == 1) ? 4 : 3))
car3d_grad_image_plane_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                              const int *__restrict__ box_ind, CarGeom g, PlaneLaunch L,
                              float *__restrict__ grad_image, const PyrParams P, const int only_image,
                              const __grid_constant__ CUtensorMap tmap)
{
    constexpr int UNR = (V == 1) ? 4 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);  off += (sizeof(PlaneShared) + 15) & ~size_t(15);
    BwdLists &B = *reinterpret_cast<BwdLists *>(smem_raw + off);  off += (sizeof(BwdLists) + 15) & ~size_t(15);
    if constexpr (TMA) off += (128u - ((smem_u32(smem_raw) + (unsigned)off) & 127u)) & 127u;   // tensor copies: 128-byte aligned destination
    unsigned char *Graw = smem_raw + off;                          // staged grads slice [(y-ya)*pw + x][V][cl] float4
    __shared__ __align__(8) unsigned long long s_mbar;
    const unsigned mbar = smem_u32(&s_mbar);
    if constexpr (TMA) { if (threadIdx.x == 0) mbar_init(mbar, 1); }                            // visible after the prologue's barriers
    unsigned tma_parity = 0;

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = L.perm ? __ldg(L.perm + bid / L.ksplits) : bid / L.ksplits;
    // per-image launches (the output slice of one image stays L2-resident between its zero-fill and its REDs)
    if (!PYR && only_image >= 0 && __ldg(box_ind + b) != only_image) return;
    __shared__ int s_level;
    if constexpr (PYR) {
        if (threadIdx.x == 0) {
            float b6[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) b6[q] = __ldg(boxes + (size_t)b * 6 + q);
            const PyrRoute r = pyr_route(b6, P);
#pragma unroll
            for (int q = 0; q < 6; ++q) S.box[q] = r.box[q];
            s_level = r.level - 2;
        }
        __syncthreads();
        const int lv = s_level;
        const float *gim;
        pyr_level(P, lv, g.H, g.W, g.D, gim);
        grad_image = const_cast<float *>(gim);
    } else {
        if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
        __syncthreads();
    }
    build_axis_tables(S, g);
    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    const int cl = L.cl;
    const unsigned ebytes = (unsigned)cl * V * 16;
    build_lists(S, B, g, ebytes);

    const int nx = X.n, ny = Y.n;
    if (nx == 0 || ny == 0) return;                           // no in-range sample: nothing to scatter
    const int lane = threadIdx.x % cl, slot = threadIdx.x / cl, vs = NT / cl;
    const int c4 = chunk * cl * V + lane;
    bool von[V];
#pragma unroll
    for (int v = 0; v < V; ++v) von[v] = FULL || (c4 + v * cl) < g.C / 4;
    const long long sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    const int bimg = PYR ? b / P.rois_per_image : __ldg(box_ind + b);
    if ((unsigned)bimg >= (unsigned)g.B) return;             // out-of-range box_ind (CTA-uniform): nothing is scattered
    float *img = grad_image + (long long)bimg * g.H * sH + c4 * 4;
    const float *gcrop = grads + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);
    const int ty = max(1, L.zcap / g.pw);                    // y samples per tile
    const long long gstride = (long long)vs * g.pd * g.C;
    const float rnx = 1.0f / (float)nx;
    const int vstep = cl * 4;
    const unsigned g_u32 = smem_u32(Graw) + lane * 16;
    const unsigned yent_u32 = smem_u32(&B.ent[0][0]), xent_u32 = smem_u32(&B.ent[1][0]);

    for (int ya = 0; ya < g.ph; ya += ty) {
        const int yb = min(g.ph, ya + ty);
        const int nent = (yb - ya) * g.pw;
        const unsigned tile_off = (unsigned)ya * g.pw * ebytes;
        int rlo = 1 << 30, rhi = -1;                           // footprint rows tapped by this tile's samples
        for (int y = ya; y < yb; ++y) {
            const int a = Y.pos0[y], c = Y.pos1[y];
            if (a < 0) continue;
            rlo = min(rlo, min(a, c));
            rhi = max(rhi, max(a, c));
        }
        if (rhi < rlo) continue;
        const int nvox = (rhi - rlo + 1) * nx;
        for (int k = k0; k < k1; ++k) {
            const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
            if (axis_invalid(in_z, g.D)) continue;             // uniform across the CTA
            const float zfl = floorf(in_z);
            const long long zf = (long long)(int)zfl * g.C, zc = (long long)(int)ceilf(in_z) * g.C;
            const float zl = __fsub_rn(in_z, zfl), wzf = __fsub_rn(1.0f, zl);
            // ---- stage A': stage the k-slice of grads (each element read once) -------------
            if constexpr (TMA) {
                if (threadIdx.x == 0) {
                    fence_proxy_async();                       // the previous slice's generic reads (before the barrier) are done
                    mbar_expect_tx(mbar, (unsigned)L.zcap * ebytes);
                    if (only_image == -2) {                    // A/B (car_experiment bit 32): no eviction hint
                        tma_load_3d(smem_u32(Graw), &tmap, chunk * cl * V * 4, k, (b * g.ph + ya) * g.pw, mbar);
                        if (k + 1 < k1) tma_prefetch_3d(&tmap, chunk * cl * V * 4, k + 1, (b * g.ph + ya) * g.pw);
                    } else {
                        const unsigned long long pol = l2_evict_first_policy();
                        tma_load_3d_hint(smem_u32(Graw), &tmap, chunk * cl * V * 4, k, (b * g.ph + ya) * g.pw, mbar, pol);
                        if (k + 1 < k1) tma_prefetch_3d_hint(&tmap, chunk * cl * V * 4, k + 1, (b * g.ph + ya) * g.pw, pol);
                    }
                }
                mbar_wait_bounded(mbar, tma_parity);           // every thread sees the landed slice
                tma_parity ^= 1u;
            } else {
                const float *gp = gcrop + (((long long)ya * g.pw + slot) * g.pd + k) * g.C;
                for (int base = slot; base < nent; base += vs * UNR, gp += UNR * gstride) {
                    float4 val[UNR][V];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        if (base + u * vs < nent) {
#pragma unroll
                            for (int v = 0; v < V; ++v)
                                if (von[v]) val[u][v] = ldg4(gp + u * gstride + v * vstep);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        if (base + u * vs < nent) {
#pragma unroll
                            for (int v = 0; v < V; ++v)
                                if (von[v]) sts128(g_u32 + (base + u * vs) * ebytes + v * (cl * 16), val[u][v]);
                        }
                    }
                }
            }
            if constexpr (!TMA) __syncthreads();
            // ---- L2 prefetch of the next depth sample's grads slice ----------------------------
            if (!TMA && k + 1 < k1 && (lane & 7) == 0) {
                const float *gn = gcrop + (((long long)ya * g.pw + slot) * g.pd + (k + 1)) * g.C;
                for (int idx = slot; idx < nent; idx += vs, gn += gstride) {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (von[v]) prefetch_l2(gn + v * vstep);
                }
            }
            // (PDL: everything above only read this op's inputs; the zero-fill kernel launched just before must be
            // complete before the first RED -- a no-op when the kernel was launched in plain stream order)
            pdl_wait();
            // ---- stage B': every footprint voxel gathers its weighted sum, then 2 REDs ------
            for (int idx = slot; idx < nvox; idx += vs) {
                const int rr = (int)(((float)idx + 0.5f) * rnx), cx = idx - rr * nx, r = rlo + rr;
                const int ys = B.start[0][r], yn = B.cnt[0][r];
                const int xs = B.start[1][cx], xn = B.cnt[1][cx];
                float4 acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                bool any = false;
#pragma unroll 1
                for (int i = ys; i < ys + yn; ++i) {
                    const int y = B.samp[0][i];
                    if (y < ya || y >= yb) continue;
                    any = true;
                    const uint2 ye = lds64u(yent_u32 + i * 8);
                    const unsigned row = g_u32 + ye.x - tile_off;
                    const float wy = __uint_as_float(ye.y);
#pragma unroll 1
                    for (int j = xs; j < xs + xn; ++j) {
                        const uint2 xe = lds64u(xent_u32 + j * 8);
                        const float w = __fmul_rn(wy, __uint_as_float(xe.y));
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if (von[v]) {
                                const float4 gv = lds128(row + xe.x + v * (cl * 16));
                                acc[v].x = __fmaf_rn(gv.x, w, acc[v].x); acc[v].y = __fmaf_rn(gv.y, w, acc[v].y);
                                acc[v].z = __fmaf_rn(gv.z, w, acc[v].z); acc[v].w = __fmaf_rn(gv.w, w, acc[v].w);
                            }
                        }
                    }
                }
                if (!any) continue;
                float *p = img + Y.list[r] * sH + X.list[cx] * sW;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if (von[v]) {
                        red_add4(p + zf + v * vstep, make_float4(acc[v].x * wzf, acc[v].y * wzf, acc[v].z * wzf, acc[v].w * wzf));
                        red_add4(p + zc + v * vstep, make_float4(acc[v].x * zl, acc[v].y * zl, acc[v].z * zl, acc[v].w * zl));
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------
// ROI processing order.  CTAs are dispatched in blockIdx order, so the order of the ROIs decides which footprints are
// in flight together: ROIs of one image whose footprints are neighbours along y -- the slowest-varying axis of the
// [B,H,W,D,C] layout, 4 MB per row at P2 -- share L2 lines if they are processed close in time.  Measured at cfg2
// (profiles/locality_experiment.py): ROIs sorted by (image, y centre) instead of "as generated": forward 14^3 0.259 ->
// 0.242 ms, grad-image 14^3 0.318 -> 0.299 ms; images interleaved at random: 0.282 / 0.337 ms.  This kernel builds that
// order on the device -- a counting sort by (image, 64 y buckets) done by one CTA -- into a caller-owned workspace; the
// crop-and-resize kernels then map CTA group i to ROI perm[i].  Results do not depend on the order (the forward is a pure
// gather; the backward's REDs commute up to fp32 rounding, as before).
// ---------------------------------------------------------------------------------
constexpr int ORD_YB = 64, ORD_IMG = 64, ORD_THREADS = 1024;

__global__ void __launch_bounds__(ORD_THREADS)
car3d_order_kernel(const float *__restrict__ boxes, const int *__restrict__ box_index, int n, int rois_per_image,
                   int *__restrict__ perm)
{
    __shared__ int s_cnt[ORD_YB * ORD_IMG];
    __shared__ int s_part[ORD_THREADS / 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < ORD_YB * ORD_IMG; i += ORD_THREADS) s_cnt[i] = 0;
    __syncthreads();
    auto bucket = [&](int i) {
        int img = box_index ? __ldg(box_index + i) : i / rois_per_image;
        img = min(max(img, 0), ORD_IMG - 1);
        const float yc = 0.5f * (__ldg(boxes + (size_t)i * 6) + __ldg(boxes + (size_t)i * 6 + 3));
        const int yb = min(max((int)(yc * (float)ORD_YB), 0), ORD_YB - 1);    // NaN -> 0
        return img * ORD_YB + yb;
    };
    for (int i = tid; i < n; i += ORD_THREADS) atomicAdd(&s_cnt[bucket(i)], 1);
    __syncthreads();
    // exclusive scan over the 4096 buckets: 4 per thread + warp / block scan
    constexpr int PER = ORD_YB * ORD_IMG / ORD_THREADS;
    int v[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { v[q] = s_cnt[tid * PER + q]; sum += v[q]; }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if ((tid & 31) >= d) inc += t; }
    if ((tid & 31) == 31) s_part[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        int w = s_part[tid];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, d); if (tid >= d) w += t; }
        s_part[tid] = w;
    }
    __syncthreads();
    int base = inc - sum + ((tid >> 5) ? s_part[(tid >> 5) - 1] : 0);
#pragma unroll
    for (int q = 0; q < PER; ++q) { s_cnt[tid * PER + q] = base; base += v[q]; }    // bucket -> first slot
    __syncthreads();
    for (int i = tid; i < n; i += ORD_THREADS) perm[atomicAdd(&s_cnt[bucket(i)], 1)] = i;
}

int launch_car3d_order(const float *boxes, const int *box_index, int n, int rois_per_image, int *perm, cudaStream_t stream)
{
    car3d_order_kernel<<<1, ORD_THREADS, 0, stream>>>(boxes, box_index, n, rois_per_image > 0 ? rois_per_image : 1, perm);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
static inline size_t a16(size_t v) { return (v + 15) & ~size_t(15); }

// channel lanes: 16 lanes x V float4 per voxel; narrower when C is small so that no lane idles
static void pick_lanes(const CarGeom &g, int &cl, int &V) {
    const int c4 = g.C / 4;
    V = (c4 >= 32) ? 2 : 1;
    const int forced = option_value(OPT_CAR_V);
    if (forced == 1 || forced == 2) V = forced;
    cl = 16;
    while (cl > 1 && cl / 2 * V >= c4) cl /= 2;
}

static int pick_ksplits(const CarGeom &g, int chunks) {
    const long long want = (long long)num_sms() * (option_value(OPT_KSPLIT) > 0 ? option_value(OPT_KSPLIT) : 16);
    const long long per = (long long)g.n * chunks;
    long long ks = (want + per - 1) / per;
    if (ks < 1) ks = 1;
    if (ks > g.pd) ks = g.pd;
    return (int)ks;
}

static int launch_fwd_plane_impl(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                                 float ext, void *crops, const PyrParams *pyr, cudaStream_t stream, bool half_out = false,
                                 const int *perm = nullptr)
{
    PlaneLaunch L{};
    L.perm = perm;
    int V;
    pick_lanes(g, L.cl, V);
    const int nxmax = min(2 * g.pw, g.W);
    L.otab = min(g.ph * g.pw, max(PL_MAXOUT, g.pw));
    size_t smem;
    for (;;) {
        const size_t eb = (size_t)L.cl * V * 16;
        L.zcap = (int)max((size_t)2 * nxmax, (size_t)((V == 1 ? 46 : 62) * 1024) / eb);   // 4 (V=1) / 3 (V=2) CTAs per SM
        while ((size_t)L.zcap * eb > 65535 && L.zcap > 2 * nxmax) --L.zcap;     // plane offsets are packed in 16 bits
        smem = a16(sizeof(PlaneShared)) + sizeof(OutEntry) * (size_t)L.otab + a16(sizeof(unsigned) * (size_t)L.zcap) +
               (size_t)L.zcap * eb;
        if ((size_t)L.zcap * eb <= 65535 && smem <= 200 * 1024) break;
        if (V > 1) V = 1; else if (L.cl > 1) L.cl /= 2; else return ROI3D_EUNSUPPORTED;
    }
    L.chunks = (g.C / 4 + L.cl * V - 1) / (L.cl * V);
    L.ksplits = pick_ksplits(g, L.chunks);
    if (half_out && !pyr) return ROI3D_EUNSUPPORTED;           // float16 output exists for the fused pyramid forward only
    const bool full = (g.C / 4) % (L.cl * V) == 0 && !(option_value(OPT_EXPERIMENT) & 1);
    using K = void (*)(const float *, const float *, const int *, CarGeom, PlaneLaunch, float, void *, const PyrParams);
    K kern;
    if (V == 2) kern = half_out ? (full ? (K)car3d_fwd_plane_kernel<2, true, true, true> : (K)car3d_fwd_plane_kernel<2, true, true, false>)
                     : pyr      ? (full ? (K)car3d_fwd_plane_kernel<2, true, false, true> : (K)car3d_fwd_plane_kernel<2, true, false, false>)
                                : (full ? (K)car3d_fwd_plane_kernel<2, false, false, true> : (K)car3d_fwd_plane_kernel<2, false, false, false>);
    else        kern = half_out ? (full ? (K)car3d_fwd_plane_kernel<1, true, true, true> : (K)car3d_fwd_plane_kernel<1, true, true, false>)
                     : pyr      ? (full ? (K)car3d_fwd_plane_kernel<1, true, false, true> : (K)car3d_fwd_plane_kernel<1, true, false, false>)
                                : (full ? (K)car3d_fwd_plane_kernel<1, false, false, true> : (K)car3d_fwd_plane_kernel<1, false, false, false>);
    // 512-thread CTAs (32 voxel / output slots per pass, two CTAs per SM at 64 registers) for crops of 12 x 12 outputs per
    // depth slice and more: fewer dependent passes per stage -- cfg2 14^3 0.246 -> 0.234 ms, 28^3 0.339 -> 0.31 ms, 7^3
    // slower (profiles/fwd_threads_ab.py; car_experiment bit 128 keeps 256 threads)
    int nthreads = PL_THREADS;
    if (V == 2 && full && g.ph * g.pw >= 144 && !(option_value(OPT_EXPERIMENT) & 128)) {
        kern = half_out ? (K)car3d_fwd_plane_kernel<2, true, true, true, 512>
             : pyr      ? (K)car3d_fwd_plane_kernel<2, true, false, true, 512>
                        : (K)car3d_fwd_plane_kernel<2, false, false, true, 512>;
        nthreads = 512;
    }
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    kern<<<(unsigned)grid, nthreads, smem, stream>>>(image, boxes, box_index, g, L, ext, crops, pyr ? *pyr : PyrParams{});
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_car3d_fwd_plane(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                           float ext, float *crops, cudaStream_t stream, const int *perm)
{
    return launch_fwd_plane_impl(image, boxes, box_index, g, ext, crops, nullptr, stream, false, perm);
}

int launch_car3d_fwd_plane_tma(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                               float ext, float *crops, cudaStream_t stream)
{
    PlaneLaunch L{};
    L.cl = 16;
    while (L.cl > 1 && L.cl / 2 >= g.C / 4) L.cl /= 2;
    const int nxmax = min(2 * g.pw, g.W);
    L.otab = min(g.ph * g.pw, max(PL_MAXOUT, g.pw));
    const size_t eb = (size_t)L.cl * 16;
    L.zcap = (int)max((size_t)2 * nxmax, (size_t)(22 * 1024) / eb);       // 3 buffers of ~22 KB
    if ((size_t)L.zcap * eb > 65535) return ROI3D_EUNSUPPORTED;
    const size_t smem = a16(sizeof(PlaneShared)) + sizeof(OutEntry) * (size_t)L.otab + a16(sizeof(unsigned) * (size_t)L.zcap) +
                        128 + 3 * (size_t)L.zcap * eb;
    if (smem > 200 * 1024) return ROI3D_EUNSUPPORTED;
    L.chunks = (g.C / 4 + L.cl - 1) / L.cl;
    L.ksplits = pick_ksplits(g, L.chunks);
    auto kern = car3d_fwd_plane_tma_kernel<false>;
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    kern<<<(unsigned)grid, PL_THREADS, smem, stream>>>(image, boxes, box_index, g, L, ext, crops, PyrParams{});
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn plane_encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int launch_car3d_fwd_plane_g4(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                              float ext, float *crops, cudaStream_t stream)
{
    PlaneLaunch L{};
    L.cl = 16;
    while (L.cl > 1 && L.cl / 2 >= g.C / 4) L.cl /= 2;
    const long long rows = (long long)g.B * g.H * g.W * g.D;
    // a gather4 lands four chunk rows at once: 4 * cl * 16 bytes must keep the 128-byte alignment of tensor-copy destinations
    if (L.cl < 2 || rows >= (1ll << 31) || (reinterpret_cast<uintptr_t>(image) & 15) || (g.C * 4) % 16) return ROI3D_EUNSUPPORTED;
    const int nxmax = min(2 * g.pw, g.W);
    L.otab = min(g.ph * g.pw, max(PL_MAXOUT, g.pw));
    const size_t eb = (size_t)L.cl * 16;
    L.zcap = (int)max((size_t)2 * nxmax, (size_t)(22 * 1024) / eb);       // 3 buffers of ~22 KB
    L.zcap = (L.zcap + 3) & ~3;                                            // whole gather4 groups
    if ((size_t)L.zcap * eb > 65535) return ROI3D_EUNSUPPORTED;
    const size_t smem = a16(sizeof(PlaneShared)) + sizeof(OutEntry) * (size_t)L.otab + a16(sizeof(int) * (size_t)L.zcap) +
                        256 + 3 * (size_t)L.zcap * eb;
    if (smem > 200 * 1024) return ROI3D_EUNSUPPORTED;
    L.chunks = (g.C / 4 + L.cl - 1) / L.cl;
    L.ksplits = pick_ksplits(g, L.chunks);
    EncodeTiledFn encode = plane_encode_tiled_fn();
    if (!encode) return ROI3D_EUNSUPPORTED;
    // image [B, H, W, D, C] viewed as a 2-D tensor (C columns, B*H*W*D rows); box = one channel chunk x 1 row (gather4
    // brings four such rows per instruction)
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)g.C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)g.C * 4};
    const cuuint32_t box[2] = {(cuuint32_t)L.cl * 4, 1}, estr[2] = {1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(image), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return ROI3D_EUNSUPPORTED;
    auto kern = car3d_fwd_plane_g4_kernel;
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    kern<<<(unsigned)grid, PL_THREADS, smem, stream>>>(tmap, boxes, box_index, g, L, ext, crops);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

// The op's zero-fill as a kernel that lets its successor start early: the scatter kernel behind it is launched with
// programmatic dependent launch, builds its tables and stages its first grads slice while the zeros are still being
// written, and waits (griddepcontrol.wait) only before its first RED.
__global__ void __launch_bounds__(256)
zero_fill_kernel(float4 *__restrict__ p, size_t n4)
{
    pdl_trigger();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) { p[i] = z; p[i + stride] = z; }
    if (i < n4) p[i] = z;
}

struct Fill4 { float4 *p[4]; size_t n4[4]; };
// the fused PyramidROIAlign backward's zero-fill of the four grad maps in ONE kernel that triggers its successor at once
__global__ void __launch_bounds__(256)
zero_fill4_kernel(const Fill4 f)
{
    pdl_trigger();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int r = 3; r >= 0; --r) {                              // coarsest level first: the largest map (P2, most ROIs) is the
        float4 *p = f.p[r];                                     // most recently written one when the scatter kernel starts
        const size_t n4 = f.n4[r];
        size_t i = t0;
        for (; i + stride < n4; i += 2 * stride) { p[i] = z; p[i + stride] = z; }
        if (i < n4) p[i] = z;
    }
}

// CTAs per SM of the zero-fill kernels.  They need few registers and no shared memory, so a fill grid that does not
// occupy every thread slot lets CTAs of the PDL-launched scatter kernel become resident next to it and run their
// prologue (tables, first grads slice) while the zeros are written (option "car_fill_ctas_per_sm").
static inline unsigned fill_grid() {
    const int per_sm = option_value(OPT_FILL_CTAS) > 0 ? option_value(OPT_FILL_CTAS) : 8;
    return (unsigned)(num_sms() * per_sm);
}

static int launch_grad_plane_impl(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                                  float *grad_image, const PyrParams *pyr, cudaStream_t stream, bool zero_fill = false,
                                  bool pdl_after_fill = false, bool tma = false, const int *perm = nullptr)
{
    PlaneLaunch L{};
    L.perm = perm;
    int V;
    pick_lanes(g, L.cl, V);
    L.otab = 0;
    const size_t fixed = a16(sizeof(PlaneShared)) + a16(sizeof(BwdLists));
    size_t smem;
    for (;;) {
        const size_t eb = (size_t)L.cl * V * 16;
        // stage whole k-slices when they fit ~50 KB, else tile over y samples
        const int kib = option_value(OPT_BWD_STAGE_KIB) > 0 ? option_value(OPT_BWD_STAGE_KIB) : (V == 1 ? 50 : 64);
        L.zcap = max(g.pw, min(g.ph * g.pw, (int)((kib * 1024) / eb)));
        if (!(option_value(OPT_EXPERIMENT) & 4)) {   // equal y-tiles: 14 sample rows at a capacity of 9 become 7 + 7 instead of 9 + 5 (-2 % at cfg2 14^3,
            // profiles/bwd_stage_sweep.py) and the CTA asks for no more shared memory than its tiles use
            const int ty = max(1, L.zcap / g.pw), nt = (g.ph + ty - 1) / ty;
            L.zcap = ((g.ph + nt - 1) / nt) * g.pw;
        }
        smem = fixed + (size_t)L.zcap * eb + (tma ? 128 : 0);
        if (smem <= 200 * 1024) break;
        if (V > 1) V = 1; else if (L.cl > 1) L.cl /= 2; else return ROI3D_EUNSUPPORTED;
    }
    L.chunks = (g.C / 4 + L.cl * V - 1) / (L.cl * V);
    L.ksplits = pick_ksplits(g, L.chunks);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (tma) {
        // grads [N, ph, pw, pd, C] viewed as a 3-D tensor (C, pd, N*ph*pw); box = one channel chunk x 1 depth sample x zcap rows
        const long long rows = (long long)g.n * g.ph * g.pw;
        EncodeTiledFn encode = plane_encode_tiled_fn();
        if (!encode || L.zcap > 256 || L.cl * V * 4 > 256 || rows >= (1ll << 31) ||
            (reinterpret_cast<uintptr_t>(grads) & 15) || (g.C * 4) % 16)
            return ROI3D_EUNSUPPORTED;
        const cuuint64_t dims[3] = {(cuuint64_t)g.C, (cuuint64_t)g.pd, (cuuint64_t)rows};
        const cuuint64_t strides[2] = {(cuuint64_t)g.C * 4, (cuuint64_t)g.pd * g.C * 4};
        const cuuint32_t box[3] = {(cuuint32_t)(L.cl * V * 4), 1, (cuuint32_t)L.zcap}, estr[3] = {1, 1, 1};
        if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(grads), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return ROI3D_EUNSUPPORTED;
    }
    // FULL pays off in the pyramid instantiation only (cfg2 14^3: 0.377 vs 0.398 ms); in the plain one ptxas then spills
    // inside the gather loop (0.398 vs 0.371 ms) -- measured with profiles/full_ab.py (car_experiment bit 2 flips the choice)
    const bool full = (g.C / 4) % (L.cl * V) == 0 && ((pyr != nullptr || tma) != ((option_value(OPT_EXPERIMENT) & 2) != 0));
    using K = void (*)(const float *, const float *, const int *, CarGeom, PlaneLaunch, float *, const PyrParams, const int, const CUtensorMap);
    K kern;
    if (tma && pyr) kern = (V == 2) ? (full ? (K)car3d_grad_image_plane_kernel<2, true, true, true> : (K)car3d_grad_image_plane_kernel<2, true, false, true>)
                                    : (full ? (K)car3d_grad_image_plane_kernel<1, true, true, true> : (K)car3d_grad_image_plane_kernel<1, true, false, true>);
    else if (tma)   kern = (V == 2) ? (full ? (K)car3d_grad_image_plane_kernel<2, false, true, true> : (K)car3d_grad_image_plane_kernel<2, false, false, true>)
                                    : (full ? (K)car3d_grad_image_plane_kernel<1, false, true, true> : (K)car3d_grad_image_plane_kernel<1, false, false, true>);
    else if (V == 2) kern = pyr ? (full ? (K)car3d_grad_image_plane_kernel<2, true, true> : (K)car3d_grad_image_plane_kernel<2, true, false>)
                           : (full ? (K)car3d_grad_image_plane_kernel<2, false, true> : (K)car3d_grad_image_plane_kernel<2, false, false>);
    else        kern = pyr ? (full ? (K)car3d_grad_image_plane_kernel<1, true, true> : (K)car3d_grad_image_plane_kernel<1, true, false>)
                           : (full ? (K)car3d_grad_image_plane_kernel<1, false, true> : (K)car3d_grad_image_plane_kernel<1, false, false>);
    int nthreads = PL_THREADS;
    if ((option_value(OPT_EXPERIMENT) & 256) && tma && V == 2 && full) {            // A/B: 512-thread CTAs
        kern = pyr ? (K)car3d_grad_image_plane_kernel<2, true, true, true, 512> : (K)car3d_grad_image_plane_kernel<2, false, true, true, 512>;
        nthreads = 512;
    }
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    const int nohint = (option_value(OPT_EXPERIMENT) & 32) ? -2 : -1;      // only_image < 0: all images (-2: A/B, no L2 hint)
    if (zero_fill) {
        // Opt-in experiment ("car_bwd_image_split" = 1): zero-fill and scatter image by image, so that the REDs find the zeros
        // still in L2 instead of fetching them back from DRAM.  Measured at cfg2 (profiles/split_experiment.py): 0.438 vs
        // 0.390 ms at 14^3, 0.153 vs 0.142 ms at 7^3 -- the half-empty grids cost more than the traffic saves.  Off by default.
        const size_t per_image = (size_t)g.H * g.W * g.D * g.C * sizeof(float);
        const bool split = option_value(OPT_BWD_SPLIT) == 1 && g.B > 1;
        for (int img = 0; img < (split ? g.B : 1); ++img) {
            const size_t n4 = (split ? per_image : per_image * g.B) / 16;
            float *dst = grad_image + (split ? (size_t)img * (per_image / 4) : 0);
            zero_fill_kernel<<<fill_grid(), 256, 0, stream>>>(reinterpret_cast<float4 *>(dst), n4);
            ROI3D_LAUNCH_CHECK();
            ROI3D_CUDA_TRY(launch_dependent(kern, dim3((unsigned)grid), dim3(nthreads), smem, stream, true, grads, boxes, box_ind, g, L,
                                            grad_image, PyrParams{}, split ? img : nohint, tmap));
            if (img + 1 < (split ? g.B : 1)) ROI3D_LAUNCH_CHECK();
        }
    } else if (pdl_after_fill) {                               // the caller has just enqueued the zero-fill kernel
        ROI3D_CUDA_TRY(launch_dependent(kern, dim3((unsigned)grid), dim3(nthreads), smem, stream, true, grads, boxes, box_ind, g, L,
                                        grad_image, pyr ? *pyr : PyrParams{}, nohint, tmap));
    } else {
        kern<<<(unsigned)grid, nthreads, smem, stream>>>(grads, boxes, box_ind, g, L, grad_image, pyr ? *pyr : PyrParams{}, nohint, tmap);
    }
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

// zero_fill: the launcher also performs the op's zero-fill of grad_image (as a kernel the scatter kernel overlaps with)
int launch_car3d_grad_image_plane(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                                  float *grad_image, cudaStream_t stream, bool zero_fill, bool tma, const int *perm)
{
    return launch_grad_plane_impl(grads, boxes, box_ind, g, grad_image, nullptr, stream, zero_fill, false, tma, perm);
}

// ---- fused PyramidROIAlign entry points (geometry g: B, C, n = B * R, crop; H/W/D = the largest level, for sizing) ----
int launch_pyramid_fwd(const float *const images[4], const int H[4], const int W[4], const int D[4], int B, int C,
                       const float *boxes, int rois_per_image, float imH, float imW, float imD,
                       int ph, int pw, int pd, void *crops, bool half_out, cudaStream_t stream, const int *perm)
{
    PyrParams P;
    for (int l = 0; l < 4; ++l) { P.image[l] = images[l]; P.H[l] = H[l]; P.W[l] = W[l]; P.D[l] = D[l]; }
    P.imH = imH; P.imW = imW; P.imD = imD; P.rois_per_image = rois_per_image;
    pyr_level_thresholds(imH, imW, imD, P.lt);
    int wmax = 1;
    for (int l = 0; l < 4; ++l) wmax = max(wmax, W[l]);
    const CarGeom g{B, H[0], wmax, D[0], C, B * rois_per_image, ph, pw, pd};
    return launch_fwd_plane_impl(nullptr, boxes, nullptr, g, 0.0f, crops, &P, stream, half_out, perm);
}

int launch_pyramid_grad(const float *grads, float *const grad_images[4], const int H[4], const int W[4], const int D[4],
                        int B, int C, const float *boxes, int rois_per_image, float imH, float imW, float imD,
                        int ph, int pw, int pd, cudaStream_t stream, const int *perm)
{
    PyrParams P;
    Fill4 f;
    for (int l = 0; l < 4; ++l) {
        P.image[l] = grad_images[l]; P.H[l] = H[l]; P.W[l] = W[l]; P.D[l] = D[l];
        f.p[l] = reinterpret_cast<float4 *>(grad_images[l]);
        f.n4[l] = (size_t)B * H[l] * W[l] * D[l] * C / 4;
    }
    P.imH = imH; P.imW = imW; P.imD = imD; P.rois_per_image = rois_per_image;
    pyr_level_thresholds(imH, imW, imD, P.lt);
    int wmax = 1;
    for (int l = 0; l < 4; ++l) wmax = max(wmax, W[l]);
    const CarGeom g{B, H[0], wmax, D[0], C, B * rois_per_image, ph, pw, pd};
    // one zero-fill kernel for the four maps; the scatter kernel behind it is launched with programmatic dependent
    // launch and waits (griddepcontrol.wait) only before its first RED (option "pdl" = 1: plain stream order)
    zero_fill4_kernel<<<fill_grid(), 256, 0, stream>>>(f);
    ROI3D_LAUNCH_CHECK();
    if (g.n == 0) return ROI3D_OK;
    if (option_value(OPT_CAR_BWD_VARIANT) != 2) {               // TMA-staged grads slices (default); car_bwd_variant 2 = LDG staging
        const int rc = launch_grad_plane_impl(grads, boxes, nullptr, g, nullptr, &P, stream, false, option_value(OPT_PDL) == 0, true, perm);
        if (rc != ROI3D_EUNSUPPORTED) return rc;
    }
    return launch_grad_plane_impl(grads, boxes, nullptr, g, nullptr, &P, stream, false, option_value(OPT_PDL) == 0, false, perm);
}

}  // namespace roi3d

// roi3d_car_os.cu -- output-stationary CropAndResize3DGradImage (variant 3): every voxel of grad_image is
// produced by exactly one thread and stored exactly ONCE (its sum, or zero), with no atomics and no zero-fill.
//
// Why (round-1 profile of the scatter kernel, profiles/r1h_ncu_summary.txt): zero-fill (268 MB at cfg2) + RED
// scatter moved 1.80 GB through DRAM for an op whose compulsory traffic is "write the output once + read grads
// once" = 0.99 GB: the zeros were evicted from L2 before the REDs arrived and had to be fetched and written again.
//
// Shape of the kernel.  A CTA owns a tile of the output: 4 x 4 voxel columns (y, x) x TZ voxels (z) x one channel
// chunk (16 float4 lanes x V groups).  A thread owns one (column, lane): its TZ x V accumulators live in shared
// memory, private to the thread, so accumulation needs neither atomics nor barriers.  The CTA
//   1. scans the boxes and keeps, in ascending box order, those of its image whose sample footprint can touch the
//      tile (conservative test on the first / last sample coordinate of each axis -- exact tables follow);
//   2. per batch of OS_NB kept boxes builds per-axis tables in shared memory with the reference's own coordinate
//      arithmetic (GI.so@0x3a80: same `in`, floor, lerp as the forward): for y and x the contiguous sample range
//      that taps each of the tile's 4 rows / columns and its weights (1 - t for a floor tap, t for a ceil tap),
//      for z the tile-local floor index and lerp of each sample;
//   3. a producer warp stages, box after box, the sub-block of grads that taps the tile -- for every (y, x) sample the
//      depth samples k that reach the tile's z range, 64 channels -- into a ring of shared-memory stages with TMA
//      tensor copies (cp.async.bulk.tensor.3d over grads viewed as [N*ph*pw, pd, C]; one copy per (y, x) sample,
//      box = 2^c depth samples x 64 channels; completion on an mbarrier), running up to OS_NS stages ahead;
//   4. the 8 consumer warps wait for a stage, and every thread adds the samples of its own (y, x) ranges into
//      S[k] (registers): S[k] += (wy * wx) * staged[y, x, k] with 16-byte LDS at immediate offsets and packed
//      fp32 FMAs; after the box's last stage acc[floor z] += (1 - zl) * S[k], acc[ceil z] += zl * S[k];
//   5. every thread stores its column.
// The order of the additions per voxel is fixed (box, k, y, x): the result is deterministic run to run, unlike the
// RED scatter.  It is not the reference's order (box, y, x, k with unfactored weights), so parity stays a tolerance
// (<= 1e-4, tests/test_gpu_parity.py), not bit equality.
//
// Algorithmic bytes (SURVEY.md 8d formula is kept for the roofline); compulsory DRAM traffic = B*H*W*D*C*4 written
// once + N*ph*pw*pd*C*4 read once.
#include "roi3d_common.cuh"
#include <cuda.h>                            // CUtensorMap (the encode function is fetched from the driver at run time)

namespace roi3d {

#ifndef OS_TY_
#define OS_TY_ 4
#endif
constexpr int OS_TY = OS_TY_, OS_TX = 4;     // voxel columns per tile (y, x)
constexpr int OS_LANES = 16;
constexpr int OS_CONSUMERS = OS_TY * OS_TX * OS_LANES;   // consumer warps: one thread per (column, channel lane)
constexpr int OS_THREADS = OS_CONSUMERS + 32;   // + the producer warp
constexpr int OS_MINB = (OS_TY == 1) ? 5 : (OS_TY == 2 ? 3 : 2);
constexpr int OS_CH = OS_LANES * 4;          // channels per CTA
constexpr int OS_ROWB = OS_CH * 4;           // bytes of one staged depth sample
constexpr int OS_NB = 16;                    // boxes per table batch
constexpr int OS_MAXP = 32;                  // crop size per axis handled here (one warp lane per sample)
constexpr int OS_KMAX = 8;                   // depth samples per pass: S[OS_KMAX] stays in registers
constexpr int OS_RNG = 12;                   // ints per box in the range table
constexpr int OS_NONE = -128;                // "sample taps nothing near this tile" marker for tile-local floor indices

struct alignas(64) OsMaps { CUtensorMap m[4]; };   // box depth 1, 2, 4, 8 samples

struct OsLaunch {
    int tz;                                  // tile depth in voxels
    int chunks;                              // channel chunks of OS_CH
    int ty_tiles, tx_tiles, tz_tiles;
    int ps;                                  // table stride per axis (>= max(ph, pw, pd))
    int ns, sb;                              // stages in the ring, bytes per stage
    int debug;                               // experiments: 1 = consumers skip the arithmetic, 2 = producer skips the copies
};

struct OsTables {                            // dynamic shared memory after the stages and accumulators
    int *rl;                                 // [OS_CONSUMERS] boxes of the current scan round that can touch the tile, ascending
    int *wcnt;                               // [8] per-warp hit counts
    int *rng;                                // [NB][OS_RNG] (first | count << 16) of the tapping samples: y row 0..3, x column 0..3,
                                             //   z, y union, x union
    float *w;                                // [NB][2][4][ps] weight of sample s for tile row / column q
    float *zt;                               // [NB][ps] z lerp of every sample
    signed char *zf;                         // [NB][ps] tile-local z floor index (OS_NONE: out of range / far away)
};

__device__ __forceinline__ bool os_axis_hits(float a1, float a2, int dim, int p, int lo_vox, int hi_vox)
{
    const float sc = axis_scale(a1, a2, dim, p);
    const float i0 = axis_coord(a1, a2, dim, p, 0, sc), i1 = axis_coord(a1, a2, dim, p, p - 1, sc);
    float lo = fminf(i0, i1), hi = fmaxf(i0, i1);
    const float top = (float)(dim - 1);
    if (!(hi >= 0.0f) || !(lo <= top)) return false;          // no in-range sample (also NaN boxes: no hit)
    lo = fmaxf(lo, 0.0f);
    hi = fminf(hi, top);
    return (int)floorf(lo) <= hi_vox && (int)ceilf(hi) >= lo_vox;
}

__device__ __forceinline__ int os_range(unsigned m) { return m ? ((__ffs(m) - 1) | ((32 - __clz(m) - (__ffs(m) - 1)) << 16)) : 0; }

// ---- how a box's sub-block is cut into passes (<= OS_KMAX depth samples) and stages: shared by producer and consumers ----
struct OsBox { int Ys, nY, Xs, nX, ks, kn; };
struct OsPass { int k0, cls, slot, XS, R; };
struct OsRing { unsigned st, par; };             // current stage of the ring and the parity of its use
__device__ __forceinline__ void os_ring_next(OsRing &r, unsigned ns) { if (++r.st == ns) { r.st = 0; r.par ^= 1u; } }

__device__ __forceinline__ bool os_box_plan(const int *rng, OsBox &P)
{
    const int ry = rng[9], rx = rng[10], rz = rng[8];
    P.nY = ry >> 16; P.nX = rx >> 16; P.kn = rz >> 16;
    P.Ys = ry & 0xffff; P.Xs = rx & 0xffff; P.ks = rz & 0xffff;
    return P.nY != 0 && P.nX != 0 && P.kn != 0;
}
__device__ __forceinline__ void os_pass_plan(const OsBox &P, int kp, OsPass &Q, int OS_SB)
{
    Q.k0 = P.ks + kp * OS_KMAX;
    const int knp = min(OS_KMAX, P.kn - kp * OS_KMAX);
    Q.cls = (knp <= 1) ? 0 : 32 - __clz(knp - 1);             // box depth 2^cls >= knp
    Q.slot = OS_ROWB << Q.cls;
    const int cap = OS_SB / Q.slot;                           // samples per stage, >= 2
    if (P.nX <= cap) { Q.XS = P.nX; Q.R = min(P.nY, cap / P.nX); }
    else { Q.XS = cap; Q.R = 1; }
}

// ---- mbarrier / TMA / packed-FMA helpers ----------------------------------------------------------------------------
// mbarrier wait that fails loudly instead of hanging the GPU if the producer / consumer protocol is ever broken
__device__ __forceinline__ bool mbar_try(unsigned mbar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(unsigned mbar, unsigned parity) {
    const long long t0 = clock64();
    while (!mbar_try(mbar, parity))
        if (clock64() - t0 > 4000000000ll) __trap();          // ~2 s at 2 GHz: a broken protocol must not hang the GPU
}
__device__ __forceinline__ void mbar_wait_guarded(unsigned mbar, unsigned parity) {
    if (!mbar_try(mbar, parity)) mbar_wait_slow(mbar, parity);
}
__device__ __forceinline__ void mbar_arrive(unsigned mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
struct f2x2 { unsigned long long lo, hi; };                   // four floats as two packed pairs
__device__ __forceinline__ f2x2 lds_f2x2(unsigned a) {
    f2x2 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "r"(a));
    return v;
}
__device__ __forceinline__ void fma_f2x2(f2x2 &acc, const f2x2 v, unsigned long long w2) {   // acc += v * w (both halves)
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.lo) : "l"(v.lo), "l"(w2));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.hi) : "l"(v.hi), "l"(w2));
}
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ void st_f2x2(float4 *p, const f2x2 v) {
    *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(v.lo, v.hi);
}
__device__ __forceinline__ f2x2 ld_f2x2(const float4 *p) {
    const ulonglong2 u = *reinterpret_cast<const ulonglong2 *>(p);
    return f2x2{u.x, u.y};
}

// one stage's samples of this thread's (y, x) ranges into S[0..KB)
template <int KB>
__device__ __forceinline__ void os_consume(f2x2 (&S)[OS_KMAX], unsigned base, int y_lo, int y_hi, int x_lo, int x_hi,
                                           int ya, int xa, int nxs, int slot, const float *wy, const float *wx)
{
    for (int y = y_lo; y < y_hi; ++y) {
        const float wyv = wy[y];
        unsigned a = base + (unsigned)(((y - ya) * nxs + (x_lo - xa)) * slot);
#pragma unroll 1
        for (int x = x_lo; x < x_hi; ++x, a += slot) {
            const float w = __fmul_rn(wyv, wx[x]);
            const unsigned long long w2 = pack2(w, w);
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) fma_f2x2(S[kk], lds_f2x2(a + kk * OS_ROWB), w2);
        }
    }
}

__global__ void __launch_bounds__(OS_THREADS, OS_MINB)
car3d_grad_image_os_kernel(const __grid_constant__ OsMaps maps, const float *__restrict__ boxes,
                           const int *__restrict__ box_ind, CarGeom g, OsLaunch L, float *__restrict__ grad_image)
{
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);        // TMA destinations: 128-byte aligned
    const int tz = L.tz, ps = L.ps, OS_NS = L.ns, OS_SB = L.sb;
    unsigned char *stages = smem_raw;                                                          // [OS_NS][OS_SB]
    float4 *acc_all = reinterpret_cast<float4 *>(smem_raw + OS_NS * OS_SB);                    // [16 columns][tz][16 lanes]
    OsTables T;
    unsigned long long *bars;
    {
        unsigned char *p = reinterpret_cast<unsigned char *>(acc_all) + (size_t)OS_TY * OS_TX * tz * OS_LANES * sizeof(float4);
        bars = reinterpret_cast<unsigned long long *>(p);  p += sizeof(unsigned long long) * 2 * OS_NS;
        T.rl = reinterpret_cast<int *>(p);    p += sizeof(int) * OS_CONSUMERS;
        T.wcnt = reinterpret_cast<int *>(p);  p += sizeof(int) * 8;
        T.rng = reinterpret_cast<int *>(p);   p += sizeof(int) * OS_NB * OS_RNG;
        T.w = reinterpret_cast<float *>(p);   p += sizeof(float) * OS_NB * 8 * ps;
        T.zt = reinterpret_cast<float *>(p);  p += sizeof(float) * OS_NB * ps;
        T.zf = reinterpret_cast<signed char *>(p);
    }
    const int tid = threadIdx.x, lane = tid & (OS_LANES - 1), col = (tid >> 4) % (OS_TY * OS_TX), cy = col / OS_TX, cx = col % OS_TX;
    const int wlane = tid & 31, warp = tid >> 5;
    const bool consumer = warp < OS_CONSUMERS / 32;
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + OS_NS), stage0 = smem_u32(stages);

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks;   bid /= L.chunks;
    const int tzi = bid % L.tz_tiles;   bid /= L.tz_tiles;
    const int txi = bid % L.tx_tiles;   bid /= L.tx_tiles;
    const int tyi = bid % L.ty_tiles;
    const int bimg = bid / L.ty_tiles;
    const int y0 = tyi * OS_TY, x0 = txi * OS_TX, z0 = tzi * tz;
    const int zn = min(tz, g.D - z0);                                       // voxels of this tile along z
    const int c4 = chunk * OS_LANES + lane;                                 // this thread's float4 channel group
    const bool von = c4 < g.C / 4;

    if (tid == 0) {
        for (int s = 0; s < OS_NS; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, OS_CONSUMERS / 32); }
    }
    float4 *acc = acc_all + (size_t)col * tz * OS_LANES + lane;             // acc[z * 16], private to this thread
    if (consumer)
        for (int i = 0; i < tz; ++i) acc[i * OS_LANES] = make_float4(0.f, 0.f, 0.f, 0.f);
    OsRing ring{0u, 0u};                                                    // advances identically on both sides

    for (int base = 0; base < g.n; base += OS_CONSUMERS) {
        // ---- 1. which boxes can touch this tile (ascending order kept) ---------------------------------------------
        const int r = base + tid;
        bool hit = false;
        if (consumer && r < g.n && __ldg(box_ind + r) == bimg) {
            const float *b6 = boxes + (size_t)r * 6;
            hit = os_axis_hits(__ldg(b6 + 0), __ldg(b6 + 3), g.H, g.ph, y0, y0 + OS_TY - 1) &&
                  os_axis_hits(__ldg(b6 + 1), __ldg(b6 + 4), g.W, g.pw, x0, x0 + OS_TX - 1) &&
                  os_axis_hits(__ldg(b6 + 2), __ldg(b6 + 5), g.D, g.pd, z0, z0 + zn - 1);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        __syncthreads();                                                    // the previous round's list / tables are free
        if (consumer && wlane == 0) T.wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, nh = 0;
#pragma unroll
        for (int w = 0; w < OS_CONSUMERS / 32; ++w) {
            const int c = T.wcnt[w];
            before += (w < warp) ? c : 0;
            nh += c;
        }
        if (hit) T.rl[before + __popc(bal & ((1u << wlane) - 1u))] = r;
        __syncthreads();

        for (int jb = 0; jb < nh; jb += OS_NB) {
            const int nb = min(OS_NB, nh - jb);
            if (jb > 0) __syncthreads();                                    // every warp is done with the previous tables
            // ---- 2. tables of the batch: one warp per (box, axis), one lane per sample --------------------------------
            if (consumer) {
                for (int task = warp; task < nb * 3; task += OS_CONSUMERS / 32) {
                    const int j = task / 3, a = task - j * 3;
                    const int p = a == 0 ? g.ph : (a == 1 ? g.pw : g.pd), dim = a == 0 ? g.H : (a == 1 ? g.W : g.D);
                    const int org = a == 0 ? y0 : (a == 1 ? x0 : z0);
                    const float *b6 = boxes + (size_t)T.rl[jb + j] * 6;
                    const float a1 = __ldg(b6 + a), a2 = __ldg(b6 + 3 + a);
                    int fl = OS_NONE;
                    float t = 0.0f;
                    if (wlane < p) {
                        const float in = axis_coord(a1, a2, dim, p, wlane, axis_scale(a1, a2, dim, p));
                        if (!axis_invalid(in, dim)) {                       // the reference's arithmetic (GI.so@0x3a80)
                            const float f = floorf(in);
                            const int lf = (int)f - org;
                            if (lf >= -1 && lf < 120) { fl = lf; t = __fsub_rn(in, f); }
                        }
                    }
                    const bool up = t > 0.0f;
                    if (a < 2) {
                        int mine = 0;
                        unsigned any = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bool tap = q < (a == 0 ? OS_TY : OS_TX) && (fl == q || (fl + 1 == q && up));
                            const unsigned m = __ballot_sync(0xffffffffu, tap);
                            any |= m;
                            if (wlane == q) mine = os_range(m);
                            if (wlane < ps) T.w[((j * 2 + a) * 4 + q) * ps + wlane] = !tap ? 0.0f : ((fl == q) ? __fsub_rn(1.0f, t) : t);
                        }
                        if (wlane < 4) T.rng[j * OS_RNG + a * 4 + wlane] = mine;
                        if (wlane == 0) T.rng[j * OS_RNG + 9 + a] = os_range(any);
                    } else {
                        const unsigned m = __ballot_sync(0xffffffffu, (unsigned)fl < (unsigned)zn || ((unsigned)(fl + 1) < (unsigned)zn && up));
                        if (wlane == 0) T.rng[j * OS_RNG + 8] = os_range(m);
                        if (wlane < ps) { T.zt[j * ps + wlane] = t; T.zf[j * ps + wlane] = (signed char)fl; }
                    }
                }
            }
            __syncthreads();

            if (!consumer) {
                // ---- 3. producer warp: TMA tensor copies, one per (y, x) sample, up to OS_NS stages ahead ---------------
                for (int j = 0; j < nb; ++j) {
                    OsBox P;
                    if (!os_box_plan(T.rng + j * OS_RNG, P)) continue;
                    const int srow = T.rl[jb + j] * g.ph;
                    for (int kp = 0; kp * OS_KMAX < P.kn; ++kp) {
                        OsPass Q;
                        os_pass_plan(P, kp, Q, OS_SB);
                        const CUtensorMap *map = &maps.m[Q.cls];
                        for (int ya = P.Ys; ya < P.Ys + P.nY; ya += Q.R) {
                            const int ny = min(Q.R, P.Ys + P.nY - ya);
                            for (int xa = P.Xs; xa < P.Xs + P.nX; xa += Q.XS) {
                                const int nxs = min(Q.XS, P.Xs + P.nX - xa), ns = ny * nxs;
                                const unsigned fullb = full0 + 8 * ring.st;
                                if (wlane == 0) {
                                    mbar_wait_guarded(empty0 + 8 * ring.st, ring.par ^ 1u);   // all consumer warps released the stage
                                    mbar_expect_tx(fullb, L.debug == 2 ? 0u : (unsigned)(ns * Q.slot));
                                }
                                __syncwarp();
                                if (L.debug != 2) {
                                    const unsigned dst = stage0 + ring.st * OS_SB;
                                    int yy = 0, xx = wlane;                  // sample m = wlane, wlane + 32, ... -> (yy, xx)
                                    while (xx >= nxs) { xx -= nxs; ++yy; }
                                    for (int m = wlane; m < ns; m += 32) {
                                        tma_load_3d(dst + m * Q.slot, map, chunk * OS_CH, Q.k0, (srow + ya + yy) * g.pw + xa + xx, fullb);
                                        xx += 32;
                                        while (xx >= nxs) { xx -= nxs; ++yy; }
                                    }
                                }
                                os_ring_next(ring, OS_NS);
                            }
                        }
                    }
                }
            } else {
                // ---- 4. consumers: every thread, its own column; box after box, stage after stage ------------------------
                for (int j = 0; j < nb; ++j) {
                    OsBox P;
                    if (!os_box_plan(T.rng + j * OS_RNG, P)) continue;
                    const int ry = T.rng[j * OS_RNG + cy], rx = T.rng[j * OS_RNG + 4 + cx];
                    const int ys = ry & 0xffff, ye = ys + (ry >> 16), xs = rx & 0xffff, xe = xs + (rx >> 16);
                    const bool mine = von && ye > ys && xe > xs;
                    const float *wy = T.w + ((j * 2 + 0) * 4 + cy) * ps, *wx = T.w + ((j * 2 + 1) * 4 + cx) * ps;
                    for (int kp = 0; kp * OS_KMAX < P.kn; ++kp) {
                        OsPass Q;
                        os_pass_plan(P, kp, Q, OS_SB);
                        f2x2 S[OS_KMAX];
#pragma unroll
                        for (int kk = 0; kk < OS_KMAX; ++kk) S[kk] = f2x2{0ull, 0ull};
                        for (int ya = P.Ys; ya < P.Ys + P.nY; ya += Q.R) {
                            const int yb = min(ya + Q.R, P.Ys + P.nY);
                            for (int xa = P.Xs; xa < P.Xs + P.nX; xa += Q.XS) {
                                const int xb = min(xa + Q.XS, P.Xs + P.nX);
                                const unsigned st = ring.st;
                                mbar_wait_guarded(full0 + 8 * st, ring.par);        // the stage's copies have landed
                                os_ring_next(ring, OS_NS);
                                if (mine && L.debug != 1) {
                                    const int y_lo = max(ya, ys), y_hi = min(yb, ye), x_lo = max(xa, xs), x_hi = min(xb, xe);
                                    const unsigned sb = stage0 + st * OS_SB + lane * 16;
                                    switch (Q.cls) {
                                    case 0: os_consume<1>(S, sb, y_lo, y_hi, x_lo, x_hi, ya, xa, xb - xa, Q.slot, wy, wx); break;
                                    case 1: os_consume<2>(S, sb, y_lo, y_hi, x_lo, x_hi, ya, xa, xb - xa, Q.slot, wy, wx); break;
                                    case 2: os_consume<4>(S, sb, y_lo, y_hi, x_lo, x_hi, ya, xa, xb - xa, Q.slot, wy, wx); break;
                                    default: os_consume<8>(S, sb, y_lo, y_hi, x_lo, x_hi, ya, xa, xb - xa, Q.slot, wy, wx); break;
                                    }
                                }
                                __syncwarp();
                                if (wlane == 0) mbar_arrive(empty0 + 8 * st);
                            }
                        }
                        // ---- the box's depth samples into the column: acc[floor z] += (1 - zl) S, acc[ceil z] += zl S ----
                        if (mine) {
                            const int knp = min(OS_KMAX, P.kn - kp * OS_KMAX);
                            const signed char *zf = T.zf + j * ps + Q.k0;
                            const float *zt = T.zt + j * ps + Q.k0;
#pragma unroll
                            for (int kk = 0; kk < OS_KMAX; ++kk) {
                                if (kk < knp) {
                                    const int f = zf[kk];
                                    const float t = zt[kk], t0 = __fsub_rn(1.0f, t);
                                    if ((unsigned)f < (unsigned)zn) {
                                        f2x2 a = ld_f2x2(acc + f * OS_LANES);
                                        fma_f2x2(a, S[kk], pack2(t0, t0));
                                        st_f2x2(acc + f * OS_LANES, a);
                                    }
                                    if ((unsigned)(f + 1) < (unsigned)zn && t > 0.0f) {
                                        f2x2 a = ld_f2x2(acc + (f + 1) * OS_LANES);
                                        fma_f2x2(a, S[kk], pack2(t, t));
                                        st_f2x2(acc + (f + 1) * OS_LANES, a);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- 5. every voxel of the tile is stored once -----------------------------------------------------------------
    const int y = y0 + cy, x = x0 + cx;
    if (consumer && von && y < g.H && x < g.W) {
        float *out = grad_image + ((((size_t)bimg * g.H + y) * g.W + x) * g.D + z0) * g.C + c4 * 4;
        for (int z = 0; z < zn; ++z, out += g.C) *reinterpret_cast<float4 *>(out) = acc[z * OS_LANES];
    }
}

static size_t os_smem_bytes(int tz, int ps, int OS_NS, int OS_SB)
{
    return 128 + (size_t)OS_NS * OS_SB + (size_t)OS_TY * OS_TX * tz * OS_LANES * sizeof(float4) + 16 * OS_NS +
           sizeof(int) * (OS_CONSUMERS + 8) + sizeof(int) * OS_NB * OS_RNG + sizeof(float) * OS_NB * 8 * ps +
           sizeof(float) * OS_NB * ps + (((size_t)OS_NB * ps + 15) & ~size_t(15));
}

bool car3d_grad_image_os_supported(const CarGeom &g)
{
    return g.C % 4 == 0 && g.ph <= OS_MAXP && g.pw <= OS_MAXP && g.pd <= OS_MAXP &&
           (long long)g.n * g.ph * g.pw < (1ll << 31);
}

static void os_plan(const CarGeom &g, OsLaunch &L, size_t &smem)
{
    L.tz = option_value(OPT_OS_TZ) > 0 ? option_value(OPT_OS_TZ) : 16;
    L.tz = max(1, min(L.tz, min(g.D, 64)));
    L.ps = max(g.ph, max(g.pw, g.pd));
    L.ns = option_value(OPT_OS_STAGES) > 0 ? min(option_value(OPT_OS_STAGES), 16) : 2;   // measured (profiles/os_experiments.py)
    L.sb = option_value(OPT_OS_STAGE_KIB) > 0 ? min(option_value(OPT_OS_STAGE_KIB), 64) * 1024 : 16384;
    L.debug = option_value(OPT_OS_DEBUG);
    smem = os_smem_bytes(L.tz, L.ps, L.ns, L.sb);
    L.chunks = (g.C + OS_CH - 1) / OS_CH;
    L.ty_tiles = (g.H + OS_TY - 1) / OS_TY;
    L.tx_tiles = (g.W + OS_TX - 1) / OS_TX;
    L.tz_tiles = (g.D + L.tz - 1) / L.tz;
}

// CTAs the launch would have: the caller's auto rule prefers the scatter kernel when the output is too small to
// fill the GPU with tiles
long long car3d_grad_image_os_ctas(const CarGeom &g)
{
    OsLaunch L; size_t smem;
    os_plan(g, L, smem);
    return (long long)g.B * L.ty_tiles * L.tx_tiles * L.tz_tiles * L.chunks;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
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

int launch_car3d_grad_image_os(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                               float *grad_image, cudaStream_t stream)
{
    if (!car3d_grad_image_os_supported(g)) return ROI3D_EUNSUPPORTED;
    OsLaunch L; size_t smem;
    os_plan(g, L, smem);
    if (smem > 220 * 1024) return ROI3D_EUNSUPPORTED;
    const long long grid = (long long)g.B * L.ty_tiles * L.tx_tiles * L.tz_tiles * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return ROI3D_EUNSUPPORTED;
    // grads [N, ph, pw, pd, C] viewed as a 3-D tensor (C, pd, N*ph*pw); box = 64 channels x 2^c depth samples x 1
    OsMaps maps;
    const cuuint64_t dims[3] = {(cuuint64_t)g.C, (cuuint64_t)g.pd, (cuuint64_t)g.n * g.ph * g.pw};
    const cuuint64_t strides[2] = {(cuuint64_t)g.C * 4, (cuuint64_t)g.pd * g.C * 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int c = 0; c < 4; ++c) {
        const cuuint32_t box[3] = {(cuuint32_t)OS_CH, 1u << c, 1};
        const CUresult rc = encode(&maps.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(grads), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return ROI3D_EUNSUPPORTED;
    }
    ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(car3d_grad_image_os_kernel), smem));
    car3d_grad_image_os_kernel<<<(unsigned)grid, OS_THREADS, smem, stream>>>(maps, boxes, box_ind, g, L, grad_image);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

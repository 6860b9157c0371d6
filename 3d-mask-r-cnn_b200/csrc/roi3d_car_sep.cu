// roi3d_car_sep.cu -- "row-walk" separable CropAndResize3D (variant 4).
//
// The reference lerps z, then x, then y (CAR.so@0x4f88-0x5021).  The plane-staged kernels (roi3d_car_plane.cu) do the
// z-lerp once per footprint voxel and then pay four 16-byte shared-memory reads per output for the x/y lerps; ncu shows
// them neither DRAM- nor issue-bound but stuck on shared-memory/L1 traffic and barrier latency (profiles/README.md).
// Here both remaining lerps are hoisted as well, each evaluated exactly once, and the only thing that goes through
// shared memory is the x-lerped row table T:
//   pass 1  a thread owns (footprint row r, channel lane) and WALKS the row's footprint columns in ascending voxel
//           order: two 16-byte loads per voxel (floor z, ceil z) straight into registers, z-lerp in registers; the x
//           samples whose right tap is the current column are emitted from a two-entry register window (previous /
//           current z-lerped voxel) as T[r][x] = lerp(Z_left, Z_right, tx).  Loads of a whole batch of voxels are in
//           flight per thread before the first use.
//   pass 2  a thread owns (output column x, channel lane) and walks the y samples: out[y][x] = lerp(T[top][x],
//           T[bottom][x], ty), streamed out with evict-first 16-byte stores.
// Per depth sample that is ny*pw table writes and at most 2 reads per output instead of n_y*n_x writes and 4 reads per
// output, no per-output offset tables, and every intermediate is still the very expression the reference evaluates, in
// its order and rounding => bit-identical to the oracle.  Footprint rows are processed in tiles of RC rows through a
// ring of row slots, so shared memory does not grow with the footprint and the row above a tile is still in the ring.
//
// Measured (B200, cfg2 P2, profiles/sep_sweep.py, profiles/r2_sep_fwd_ncu.txt): bit-identical, 40 % fewer FP
// instructions and 4x less shared-memory traffic than the plane-staged kernel, but 0.295 vs 0.262 ms at 14^3 and 0.088 vs
// 0.082 ms at 7^3: the register-window walk serialises pass 1 (2-3 dependent load batches per row segment, only
// n_y * segments of the 16 thread slots busy), so the kernel trades shared-memory stalls for global-load and barrier
// stalls.  Opt-in (car_fwd_variant = 4), parity-tested like the other variants.
#include "roi3d_common.cuh"
#include "roi3d_car_pyr.cuh"
#include <type_traits>

namespace roi3d {

constexpr int SP_THREADS = 256;
constexpr int SP_LANES = 16;                            // channel lanes; each carries V float4 groups
constexpr int SP_SLOTS = SP_THREADS / SP_LANES;
constexpr int SP_MAXP = 64;                             // max crop size per axis

struct SepAxis {
    short pos0[SP_MAXP], pos1[SP_MAXP];                 // position of floor(in) / ceil(in) in `list` (-1: sample invalid)
    float t[SP_MAXP];                                   // in - floor(in)
    int   list[2 * SP_MAXP];                            // distinct tapped voxel indices, ASCENDING
    short cnt1[2 * SP_MAXP + 1];                        // per list position q: number of samples whose ceil tap is q ...
    short pre1[2 * SP_MAXP + 1];                        // ... and number of valid samples whose ceil tap is below q
    int   cand[2 * SP_MAXP];
    unsigned char first[2 * SP_MAXP];
    int n;                                              // footprint size
    int dir;                                            // +1: `in` grows with the sample index, -1: it falls (swapped corners)
    int lead;                                           // invalid samples before the first valid one, in order of growing `in`
    int nvalid;
};

struct SepShared {
    SepAxis ax[2];
    float box[6];
    int level;
};

// sample index of the j-th sample in order of growing coordinate
__device__ __forceinline__ int sep_sample(const SepAxis &T, int p, int j) { return T.dir > 0 ? j : p - 1 - j; }

__device__ __forceinline__ void build_sep_tables(SepShared &S, const CarGeom &g)
{
    const int tid = threadIdx.x;
    if (tid < 2 * SP_MAXP) {
        const int a = tid / SP_MAXP, k = tid % SP_MAXP;
        const int p = a ? g.pw : g.ph, dim = a ? g.W : g.H;
        SepAxis &T = S.ax[a];
        int c0 = -1, c1 = -1;
        if (k < p) {
            const float a1 = S.box[a], a2 = S.box[3 + a];
            const float scale = axis_scale(a1, a2, dim, p);
            const float in = axis_coord(a1, a2, dim, p, k, scale);
            if (!axis_invalid(in, dim)) {
                const float fl = floorf(in);
                c0 = (int)fl;
                c1 = (int)ceilf(in);
                T.t[k] = __fsub_rn(in, fl);
            }
            if (k == 0) T.dir = (scale < 0.0f) ? -1 : 1;
        }
        T.cand[2 * k] = c0;
        T.cand[2 * k + 1] = c1;
    }
    __syncthreads();
    const int a = tid / (2 * SP_MAXP), j = tid % (2 * SP_MAXP);           // one thread per (axis, candidate)
    const int p = a ? g.pw : g.ph;
    SepAxis &T = S.ax[a];
    const int v = T.cand[j];
    {
        bool fo = j < 2 * p && v >= 0;
        for (int i = 0; i < j && fo; ++i)
            if (T.cand[i] == v) fo = false;
        T.first[j] = fo ? 1 : 0;
    }
    __syncthreads();
    {
        if (j < 2 * p) {
            int pos = -1;
            if (v >= 0) {
                pos = 0;
                for (int i = 0; i < 2 * p; ++i) pos += (T.first[i] && T.cand[i] < v) ? 1 : 0;
                if (T.first[j]) T.list[pos] = v;
            }
            if (j & 1) T.pos1[j >> 1] = (short)pos; else T.pos0[j >> 1] = (short)pos;
        }
        if (j == 2 * SP_MAXP - 1) {
            int cnt = 0;
            for (int i = 0; i < 2 * p; ++i) cnt += T.first[i];
            T.n = cnt;
        }
    }
    __syncthreads();
    {
        if (j <= T.n) {
            int c = 0, pr = 0;
            for (int s = 0; s < p; ++s) {
                const int p1 = T.pos1[s];
                if (p1 >= 0) { c += (p1 == j); pr += (p1 < j); }
            }
            T.cnt1[j] = (short)c;
            T.pre1[j] = (short)pr;
        }
        if (j == 2 * SP_MAXP - 1) {
            int fs = -1, ls = -1;
            for (int s = 0; s < p; ++s)
                if (T.pos0[s] >= 0) { if (fs < 0) fs = s; ls = s; }
            T.nvalid = fs < 0 ? 0 : ls - fs + 1;
            T.lead = fs < 0 ? p : (T.dir > 0 ? fs : p - 1 - ls);
        }
    }
    __syncthreads();
}

struct SepLaunch {
    int chunks;        // channel chunks of 16 * V float4 per voxel
    int ksplits;       // depth-sample splits per ROI
    int rc;            // footprint rows (forward) / sample rows (backward) per tile
    int ns;            // row slots in the ring (>= rc + 1; >= 2 * rc + 1: one barrier per tile instead of two)
};

template <int V> struct SepCfg {
    static constexpr int NB = (V == 1) ? 4 : 2;         // footprint voxels whose taps are in flight per thread (pass 1)
    static constexpr int MINB = 3;                     // 80 registers: measured faster than 4 CTAs/SM at 64
};

// Walk tables, built once per CTA from the axis tables so that the inner loops read ONE packed entry per step instead
// of chasing pos0 / pos1 / t / list / cnt1 / pre1 and multiplying strides (ncu: only 21 % of the first version's
// instructions were lerps).
struct SepWalk {
    uint2 xc[2 * SP_MAXP];      // per footprint column: .x = element offset of the voxel column (list * sW),
                                //                       .y = first emitted sample (j order) << 16 | samples emitted
    uint2 xe[SP_MAXP];          // per valid x sample in j order: .x = byte offset of T[.][x] | (floor == ceil), .y = tx
    uint4 ye[SP_MAXP];          // per y sample in j order: .x = top row | bottom row << 16 (0xFFFF: invalid),
                                //                          .y = ty, .z = element offset of output row y (y * pw * pd * C)
    unsigned yr[2 * SP_MAXP];   // per footprint row: element offset (list * sH)
};

__device__ __forceinline__ void build_sep_walk(const SepShared &S, SepWalk &Wk, const CarGeom &g, unsigned eb)
{
    const int tid = threadIdx.x;
    const SepAxis &Y = S.ax[0], &X = S.ax[1];
    const unsigned sW = (unsigned)g.D * g.C, sH = (unsigned)g.W * g.D * g.C;
    if (tid < 2 * SP_MAXP) {
        if (tid < X.n) Wk.xc[tid] = make_uint2((unsigned)X.list[tid] * sW, ((unsigned)X.pre1[tid] << 16) | (unsigned)X.cnt1[tid]);
        if (tid < Y.n) Wk.yr[tid] = (unsigned)Y.list[tid] * sH;
    } else if (tid < 3 * SP_MAXP) {
        const int j = tid - 2 * SP_MAXP;                       // j-th valid x sample
        if (j < X.nvalid) {
            const int x = sep_sample(X, g.pw, X.lead + j);
            Wk.xe[j] = make_uint2((unsigned)x * eb | (X.pos0[x] == X.pos1[x] ? 1u : 0u), __float_as_uint(X.t[x]));
        }
    } else {
        const int j = tid - 3 * SP_MAXP;                       // j-th y sample (valid or not)
        if (j < g.ph) {
            const int y = sep_sample(Y, g.ph, j);
            const int ra = Y.pos0[y], rb = Y.pos1[y];
            Wk.ye[j] = make_uint4(ra < 0 ? 0xFFFFFFFFu : ((unsigned)ra | ((unsigned)rb << 16)), __float_as_uint(Y.t[y]),
                                  (unsigned)y * g.pw * g.pd * g.C, 0u);
        }
    }
    __syncthreads();
}

// pass-1 role of a thread slot in a tile of `nrows` rows: (row in tile, first column, end column), or inactive
__device__ __forceinline__ unsigned sep_role(int slot, int nrows, int nx)
{
    if (nrows <= 0 || nx <= 0) return 0xFFFFFFFFu;
    const int nseg = max(1, min(SP_SLOTS / nrows, nx));
    if (slot >= nrows * nseg) return 0xFFFFFFFFu;
    const int rr = slot / nseg, sg = slot - rr * nseg;
    return (unsigned)rr | ((unsigned)(sg * nx / nseg) << 8) | ((unsigned)((sg + 1) * nx / nseg) << 16);
}

// ---------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------
// The two passes are separate, non-inlined device functions on purpose: compiled inline under the kernel's register cap,
// ptxas rematerialised thread ids, 64-bit strides and shared-window addresses inside the innermost loops (ncu: 141
// instructions per 16-byte store, 21 % of them lerps).  As functions each loop nest gets its own register allocation.
struct SepP1 {                  // pass 1 arguments
    const float *rowp;          // image + batch item + footprint row + channel group
    unsigned zf, zc;            // element offsets of the floor / ceil depth taps
    float zl;
    unsigned trow;              // shared address of T[row slot][0][lane]
    unsigned xc, xe;            // shared addresses of the walk tables
    int c0, c1;                 // footprint columns [c0, c1) of this thread's segment
};

template <int V, bool FULL>
__device__ __noinline__ void sep_fwd_pass1(const SepP1 a, const unsigned vmask)
{
    constexpr int NB = SepCfg<V>::NB;
    constexpr int VS = SP_LANES * 4;
    float4 zprev[V];
#pragma unroll
    for (int v = 0; v < V; ++v) zprev[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int base = (a.c0 > 0 ? a.c0 - 1 : 0); base < a.c1; base += NB) {
        float4 f[NB][V], c[NB][V];
        unsigned em[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const uint2 xc = lds64u(a.xc + min(base + u, a.c1 - 1) * 8);       // tail: re-read a valid voxel
            em[u] = xc.y;
            const float *p = a.rowp + xc.x;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (FULL || ((vmask >> v) & 1u)) {
                    f[u][v] = ldg4(p + a.zf + v * VS);
                    c[u][v] = ldg4(p + a.zc + v * VS);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int cx = base + u;
            if (cx < a.c1) {
                float4 zcur[V];
#pragma unroll
                for (int v = 0; v < V; ++v) zcur[v] = lerp_rn(f[u][v], c[u][v], a.zl);
                if (cx >= a.c0) {
                    unsigned ea = a.xe + (em[u] >> 16) * 8;
#pragma unroll 1
                    for (int q = em[u] & 0xFFFFu; q > 0; --q, ea += 8) {
                        const uint2 e = lds64u(ea);
                        const bool same = e.x & 1u;                    // floor == ceil: both taps are this voxel
                        const float tx = __uint_as_float(e.y);
                        const unsigned dst = a.trow + (e.x & ~1u);
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if (FULL || ((vmask >> v) & 1u)) {
                                const float4 left = same ? zcur[v] : zprev[v];
                                sts128(dst + v * (SP_LANES * 16), lerp_rn(left, zcur[v], tx));
                            }
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < V; ++v) zprev[v] = zcur[v];
            }
        }
    }
}

struct SepP2 {                  // pass 2 arguments
    void *ox;                   // crops + ROI + k + output column x + channel group
    unsigned tcol;              // shared address of T[0][x][lane]
    unsigned ye;                // shared address of the y walk table
    unsigned rowbytes;
    int ringoff, ns;            // ring slot of row r = wrap(ringoff + r)
    int jlo, jhi;               // samples (j order) to emit; all of them are valid
};

template <int V, bool FULL, bool PYR, typename OutT>
__device__ __noinline__ void sep_fwd_pass2(const SepP2 a, const unsigned vmask)
{
    constexpr int VS = SP_LANES * 4;
    OutT *ox = static_cast<OutT *>(a.ox);
#pragma unroll 1
    for (unsigned ea = a.ye + a.jlo * 16, ee = a.ye + a.jhi * 16; ea < ee; ea += 16) {
        const uint4 e = lds128u(ea);
        OutT *o = ox + e.z;
        int sa = a.ringoff + (int)(e.x & 0xFFFFu), sb = a.ringoff + (int)(e.x >> 16);
        if (sa < 0) sa += a.ns;
        if (sa >= a.ns) sa -= a.ns;
        if (sb >= a.ns) sb -= a.ns;
        const float ty = __uint_as_float(e.y);
        const unsigned at = a.tcol + (unsigned)sa * a.rowbytes, ab = a.tcol + (unsigned)sb * a.rowbytes;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (FULL || ((vmask >> v) & 1u)) {
                const float4 top = lds128(at + v * (SP_LANES * 16));
                const float4 bot = lds128(ab + v * (SP_LANES * 16));
                float4 res = lerp_rn(top, bot, ty);
                if constexpr (PYR) res = scrub4(res);
                st_stream4(o + v * VS, res);
            }
        }
    }
}

// outputs that extrapolate (sample outside the volume): samples [jlo, jhi) in j order of one output column
template <int V, typename OutT>
__device__ __noinline__ void sep_fwd_ext(void *ox_, unsigned ye, int jlo, int jhi, float ext, const unsigned vmask)
{
    constexpr int VS = SP_LANES * 4;
    const float4 ext4 = make_float4(ext, ext, ext, ext);
    OutT *ox = static_cast<OutT *>(ox_);
    for (int j = jlo; j < jhi; ++j) {
        OutT *o = ox + lds128u(ye + j * 16).z;
#pragma unroll
        for (int v = 0; v < V; ++v)
            if ((vmask >> v) & 1u) st_stream4(o + v * VS, ext4);
    }
}

template <int V, bool PYR, bool HALF, bool FULL>
__global__ void __launch_bounds__(SP_THREADS, SepCfg<V>::MINB)
car3d_fwd_sep_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                     const int *__restrict__ box_index, CarGeom g, SepLaunch L, float ext,
                     void *__restrict__ crops, const PyrParams P)
{
    using OutT = typename std::conditional<HALF, __half, float>::type;
    constexpr unsigned EB = V * SP_LANES * 16;                 // bytes per table entry (one voxel/sample, one chunk)
    constexpr int VS = SP_LANES * 4;                           // floats between a thread's channel groups
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off = 0;
    SepShared &S = *reinterpret_cast<SepShared *>(smem_raw);   off += (sizeof(SepShared) + 15) & ~size_t(15);
    SepWalk &Wk = *reinterpret_cast<SepWalk *>(smem_raw + off); off += (sizeof(SepWalk) + 15) & ~size_t(15);
    unsigned char *Traw = smem_raw + off;

    int bid = blockIdx.x;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    if constexpr (PYR) {
        if (threadIdx.x == 0) {
            float b6[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) b6[q] = __ldg(boxes + (size_t)b * 6 + q);
            const PyrRoute r = pyr_route(b6, P);
#pragma unroll
            for (int q = 0; q < 6; ++q) S.box[q] = r.box[q];
            S.level = r.level - 2;
        }
        __syncthreads();
        const int lv = S.level;
        const float *lim;
        pyr_level(P, lv, g.H, g.W, g.D, lim);
        image = lim;
    } else {
        if (threadIdx.x < 6) S.box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
        __syncthreads();
    }
    build_sep_tables(S, g);
    build_sep_walk(S, Wk, g, EB);

    const SepAxis &Y = S.ax[0], &X = S.ax[1];
    const int nx = X.n, ny = Y.n;
    const int lane = threadIdx.x % SP_LANES, slot = threadIdx.x / SP_LANES;
    const int bimg = PYR ? b / P.rois_per_image : __ldg(box_index + b);
    const bool bad_img = (unsigned)bimg >= (unsigned)g.B;      // out-of-range box_index: the whole crop extrapolates
    const float *img_b = image + (long long)bimg * g.H * g.W * g.D * g.C;
    OutT *crop_b = static_cast<OutT *>(crops) + (long long)b * g.ph * g.pw * g.pd * g.C;
    const float z1 = S.box[2], z2 = S.box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);
    const int RC = L.rc, NS = L.ns;
    const bool two_bar = NS < 2 * RC + 1;
    const unsigned t_u32 = smem_u32(Traw) + lane * 16;
    const unsigned rowbytes = (unsigned)g.pw * EB;
    const int ntiles = (ny + RC - 1) / RC;
    // pass-1 roles: full tiles and the last (partial) tile
    const unsigned role_full = sep_role(slot, min(RC, ny), nx);
    const unsigned role_last = sep_role(slot, ny - (ntiles - 1) * RC, nx);
    // pass-2 role: output column x, y-sample segment [js0, js1) (small crops: several segments per column)
    const int nys = max(1, min(SP_SLOTS / g.pw, g.ph));
    const int units2 = g.pw * nys;
    const bool one_unit = units2 <= SP_SLOTS;
    const int x2 = slot % g.pw;
    int js0 = 0, js1 = 0;
    {
        const int ys = slot / g.pw;
        if (slot < units2) { js0 = ys * g.ph / nys; js1 = (ys + 1) * g.ph / nys; }
    }
    SepP1 a1;
    a1.xc = smem_u32(&Wk.xc[0]); a1.xe = smem_u32(&Wk.xe[0]);
    SepP2 a2;
    a2.ye = smem_u32(&Wk.ye[0]); a2.rowbytes = rowbytes; a2.ns = NS;
    const unsigned yr_u32 = smem_u32(&Wk.yr[0]);
    int ring = 0;                                              // ring slot of the current tile's first row

    for (int k = k0; k < k1; ++k) {
        const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
        const bool zbad = axis_invalid(in_z, g.D) || bad_img || nx == 0 || ny == 0;
        const float zfl = floorf(in_z);
        a1.zf = zbad ? 0u : (unsigned)(int)zfl * g.C;
        a1.zc = zbad ? 0u : (unsigned)(int)ceilf(in_z) * g.C;
        a1.zl = __fsub_rn(in_z, zfl);
        for (int chunk = 0; chunk < L.chunks; ++chunk) {
            const int c4 = chunk * SP_LANES * V + lane;        // first float4 channel group of this thread
            unsigned vmask = 0;
#pragma unroll
            for (int v = 0; v < V; ++v) vmask |= ((c4 + v * SP_LANES) < g.C / 4) ? (1u << v) : 0u;
            OutT *crop = crop_b + c4 * 4 + (long long)k * g.C;
            if (zbad) {                                        // uniform: every output of this depth sample extrapolates
                const float4 ext4 = make_float4(ext, ext, ext, ext);
                for (int u = slot; u < units2; u += SP_SLOTS) {
                    const int x = u % g.pw, ys = u / g.pw;
                    for (int y = ys * g.ph / nys; y < (ys + 1) * g.ph / nys; ++y) {
                        OutT *o = crop + ((long long)y * g.pw + x) * g.pd * g.C;
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            if ((vmask >> v) & 1u) st_stream4(o + v * VS, ext4);
                    }
                }
                continue;
            }
            const float *img = img_b + c4 * 4;
#pragma unroll 1
            for (int tl = 0; tl < ntiles; ++tl) {
                const int r0 = tl * RC, nrows = min(RC, ny - r0);
                // ---- pass 1: walk footprint rows, z-lerp in registers, x-lerp into T ------------------------------
                const unsigned role = (tl == ntiles - 1) ? role_last : role_full;
                if (role != 0xFFFFFFFFu) {
                    const int rr = role & 0xFF;
                    a1.c0 = (role >> 8) & 0xFF; a1.c1 = role >> 16;
                    int rs = ring + rr;
                    if (rs >= NS) rs -= NS;
                    a1.trow = t_u32 + (unsigned)rs * rowbytes;
                    a1.rowp = img + lds32u(yr_u32 + (r0 + rr) * 4);
                    sep_fwd_pass1<V, FULL>(a1, vmask);
                }
                __syncthreads();
                // ---- pass 2: walk the y samples of an output column, y-lerp from T -> crops -------------------------
                {
                    // valid samples whose bottom row lies in this tile (j order); the invalid ones before / after all
                    // valid samples are written with the first / last tile
                    const int vj0 = Y.lead + Y.pre1[r0], vj1 = (tl == ntiles - 1) ? Y.lead + Y.nvalid : Y.lead + Y.pre1[r0 + nrows];
                    a2.ringoff = ring - r0;
                    for (int u = slot; u < units2; u += SP_SLOTS) {
                        int x = x2, ja = js0, jb = js1;
                        if (!one_unit) { x = u % g.pw; const int ys = u / g.pw; ja = ys * g.ph / nys; jb = (ys + 1) * g.ph / nys; }
                        OutT *ox = crop + (unsigned)x * g.pd * g.C;
                        if (X.pos0[x] < 0) {                   // the whole output column extrapolates
                            if (tl == 0) sep_fwd_ext<V, OutT>(ox, a2.ye, ja, jb, ext, vmask);
                            continue;
                        }
                        a2.ox = ox;
                        a2.tcol = t_u32 + (unsigned)x * EB;
                        a2.jlo = max(vj0, ja); a2.jhi = min(vj1, jb);
                        sep_fwd_pass2<V, FULL, PYR, OutT>(a2, vmask);
                        if (tl == 0 && ja < Y.lead) sep_fwd_ext<V, OutT>(ox, a2.ye, ja, min(jb, Y.lead), ext, vmask);
                        if (tl == ntiles - 1 && jb > Y.lead + Y.nvalid) sep_fwd_ext<V, OutT>(ox, a2.ye, max(ja, Y.lead + Y.nvalid), jb, ext, vmask);
                    }
                }
                if (two_bar) __syncthreads();
                ring += nrows;
                if (ring >= NS) ring -= NS;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
static inline size_t a16(size_t v) { return (v + 15) & ~size_t(15); }

static int sep_pick_v(const CarGeom &g) {
    const int forced = option_value(OPT_CAR_V);
    if (forced == 1 || forced == 2) return forced;
    return 1;
}

static void sep_pick_launch(const CarGeom &g, int V, int rows_max, int entries_per_row, SepLaunch &L, size_t &smem) {
    L.chunks = (g.C / 4 + SP_LANES * V - 1) / (SP_LANES * V);
    int rc = option_value(OPT_SEP_RC) > 0 ? option_value(OPT_SEP_RC) : 8;
    rc = max(1, min(min(rc, SP_SLOTS), rows_max));
    int ns = option_value(OPT_SEP_NS) > 0 ? option_value(OPT_SEP_NS) : 2 * rc + 1;
    ns = max(ns, rc + 1);
    const size_t eb = (size_t)V * SP_LANES * 16;
    while (a16(sizeof(SepShared)) + a16(sizeof(SepWalk)) + (size_t)ns * entries_per_row * eb > 200 * 1024 && (ns > rc + 1 || rc > 1)) {
        if (ns > rc + 1) ns = rc + 1; else { --rc; ns = rc + 1; }
    }
    L.rc = rc; L.ns = ns;
    smem = a16(sizeof(SepShared)) + a16(sizeof(SepWalk)) + (size_t)ns * entries_per_row * eb;
    const long long want = (long long)num_sms() * (option_value(OPT_KSPLIT) > 0 ? option_value(OPT_KSPLIT) : 16);
    long long ks = (want + g.n - 1) / max(g.n, 1);
    if (ks < 1) ks = 1;
    if (ks > g.pd) ks = g.pd;
    L.ksplits = (int)ks;
}

int launch_car3d_fwd_sep(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                         float ext, void *crops, const PyrParams *pyr, bool half_out, cudaStream_t stream)
{
    const int V = sep_pick_v(g);
    SepLaunch L;
    size_t smem;
    sep_pick_launch(g, V, min(2 * g.ph, g.H), g.pw, L, smem);
    if (smem > 200 * 1024) return ROI3D_EUNSUPPORTED;
    if (half_out && !pyr) return ROI3D_EUNSUPPORTED;
    const bool full = (g.C / 4) % (SP_LANES * V) == 0;          // every lane carries live channels: no per-group predicates
    using K = void (*)(const float *, const float *, const int *, CarGeom, SepLaunch, float, void *, const PyrParams);
    K kern;
    if (V == 2) kern = half_out ? (full ? (K)car3d_fwd_sep_kernel<2, true, true, true> : (K)car3d_fwd_sep_kernel<2, true, true, false>)
                     : pyr      ? (full ? (K)car3d_fwd_sep_kernel<2, true, false, true> : (K)car3d_fwd_sep_kernel<2, true, false, false>)
                                : (full ? (K)car3d_fwd_sep_kernel<2, false, false, true> : (K)car3d_fwd_sep_kernel<2, false, false, false>);
    else        kern = half_out ? (full ? (K)car3d_fwd_sep_kernel<1, true, true, true> : (K)car3d_fwd_sep_kernel<1, true, true, false>)
                     : pyr      ? (full ? (K)car3d_fwd_sep_kernel<1, true, false, true> : (K)car3d_fwd_sep_kernel<1, true, false, false>)
                                : (full ? (K)car3d_fwd_sep_kernel<1, false, false, true> : (K)car3d_fwd_sep_kernel<1, false, false, false>);
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    const long long grid = (long long)g.n * L.ksplits;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    kern<<<(unsigned)grid, SP_THREADS, smem, stream>>>(image, boxes, box_index, g, L, ext, crops, pyr ? *pyr : PyrParams{});
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

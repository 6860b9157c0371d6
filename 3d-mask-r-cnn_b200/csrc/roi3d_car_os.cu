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
//   3. every thread walks the depth samples k that tap the tile: S = sum over its (y, x) sample ranges of
//      (wy * wx) * grads[b, y, x, k, c] (16-byte loads straight from global memory; the up to four columns that
//      share a sample find it in L1), then acc[floor z] += (1 - zl) * S and acc[ceil z] += zl * S;
//   4. stores its column.
// The order of the additions per voxel is fixed (box, k, y, x): the result is deterministic run to run, unlike the
// RED scatter.  It is not the reference's order (box, y, x, k with unfactored weights), so parity stays a tolerance
// (<= 1e-4, tests/test_gpu_parity.py), not bit equality.
//
// Algorithmic bytes (SURVEY.md 8d formula is kept for the roofline); compulsory DRAM traffic = B*H*W*D*C*4 written
// once + N*ph*pw*pd*C*4 read once.
#include "roi3d_common.cuh"

namespace roi3d {

constexpr int OS_THREADS = 256;
constexpr int OS_TY = 4, OS_TX = 4;          // voxel columns per tile: 16 columns x 16 channel lanes = 256 threads
constexpr int OS_LANES = 16;
constexpr int OS_NB = 32;                    // boxes per table batch (a tile of cfg2 sees ~20: one batch, no re-build)
constexpr int OS_MAXP = 32;                  // crop size per axis handled here (one warp lane per sample)
constexpr int OS_NONE = -128;                // "sample taps nothing near this tile" marker for tile-local floor indices

struct OsLaunch {
    int tz;                                  // tile depth in voxels
    int chunks;                              // channel chunks (of 16 * V float4)
    int cpc;                                 // chunks per CTA (they share the box list and the tables)
    int cgroups;                             // ceil(chunks / cpc)
    int ty_tiles, tx_tiles, tz_tiles;
    int ps;                                  // table stride per axis (>= max(ph, pw, pd))
};

struct OsTables {                            // dynamic shared memory after the accumulators; arrays sized by L.ps
    int *rl;                                 // [OS_THREADS] boxes of the current scan round that can touch the tile, ascending
    int *wcnt;                               // [8] per-warp hit counts
    int *rng;                                // [NB][9] (first | count << 16) of the tapping samples: y row 0..3, x column 0..3, z
    float *t;                                // [NB][3][ps] lerp of every sample
    signed char *fl;                         // [NB][3][ps] tile-local floor index (OS_NONE: out of range / far away)
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

// weight of sample (floor index f, lerp t) for tile row / column q: 1 - t as the floor tap, t as the ceil tap
__device__ __forceinline__ float os_weight(int f, float t, int q)
{
    return (f == q) ? __fsub_rn(1.0f, t) : ((f + 1 == q && t > 0.0f) ? t : 0.0f);
}

__device__ __forceinline__ void fma4(float4 &acc, const float4 v, float w)
{
    acc.x = __fmaf_rn(v.x, w, acc.x); acc.y = __fmaf_rn(v.y, w, acc.y);
    acc.z = __fmaf_rn(v.z, w, acc.z); acc.w = __fmaf_rn(v.w, w, acc.w);
}

// V = float4 channel groups per thread, KU = depth samples and OS_XU = x samples gathered together (loads in flight
// per thread: V * KU * OS_XU), MINB = CTAs per SM the register budget is set for
template <int V, int KU, int OS_XU, int MINB>
__global__ void __launch_bounds__(OS_THREADS, MINB)
car3d_grad_image_os_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                           const int *__restrict__ box_ind, CarGeom g, OsLaunch L, float *__restrict__ grad_image)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tz = L.tz, ps = L.ps;
    float4 *acc_all = reinterpret_cast<float4 *>(smem_raw);                 // [16 columns][tz][V][16 lanes]
    OsTables T;
    {
        unsigned char *p = smem_raw + (size_t)OS_TY * OS_TX * tz * V * OS_LANES * sizeof(float4);
        T.rl = reinterpret_cast<int *>(p);    p += sizeof(int) * OS_THREADS;
        T.wcnt = reinterpret_cast<int *>(p);  p += sizeof(int) * 8;
        T.rng = reinterpret_cast<int *>(p);   p += sizeof(int) * OS_NB * 9;
        T.t = reinterpret_cast<float *>(p);   p += sizeof(float) * OS_NB * 3 * ps;
        T.fl = reinterpret_cast<signed char *>(p);
    }
    const int tid = threadIdx.x, lane = tid & (OS_LANES - 1), col = tid >> 4, cy = col >> 2, cx = col & 3;
    const int wlane = tid & 31, warp = tid >> 5;

    int bid = blockIdx.x;
    const int cg = bid % L.cgroups;     bid /= L.cgroups;
    const int tzi = bid % L.tz_tiles;   bid /= L.tz_tiles;
    const int txi = bid % L.tx_tiles;   bid /= L.tx_tiles;
    const int tyi = bid % L.ty_tiles;
    const int bimg = bid / L.ty_tiles;
    const int y0 = tyi * OS_TY, x0 = txi * OS_TX, z0 = tzi * tz;
    const int zn = min(tz, g.D - z0);                                       // voxels of this tile along z

    float4 *acc = acc_all + (size_t)col * tz * V * OS_LANES + lane;         // acc[(z * V + v) * 16]
    const int pdC = g.pd * g.C;                                             // floats between x samples
    const int pwpdC = g.pw * pdC;                                           // ... between y samples
    const size_t roi_stride = (size_t)g.ph * pwpdC;

    // ---- tables of a batch of boxes: one warp per (box, axis), one lane per sample ---------------------------------
    auto build_tables = [&](int jb, int nb) {
        for (int task = warp; task < nb * 3; task += OS_THREADS / 32) {
            const int j = task / 3, a = task - j * 3;
            const int p = a == 0 ? g.ph : (a == 1 ? g.pw : g.pd), dim = a == 0 ? g.H : (a == 1 ? g.W : g.D);
            const int org = a == 0 ? y0 : (a == 1 ? x0 : z0);
            const float *b6 = boxes + (size_t)T.rl[jb + j] * 6;
            const float a1 = __ldg(b6 + a), a2 = __ldg(b6 + 3 + a);
            int fl = OS_NONE;
            float t = 0.0f;
            if (wlane < p) {
                const float in = axis_coord(a1, a2, dim, p, wlane, axis_scale(a1, a2, dim, p));
                if (!axis_invalid(in, dim)) {                               // the reference's arithmetic (GI.so@0x3a80)
                    const float f = floorf(in);
                    const int lf = (int)f - org;
                    if (lf >= -1 && lf < 120) { fl = lf; t = __fsub_rn(in, f); }
                }
            }
            const bool up = t > 0.0f;
            int mine = 0;
            if (a < 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned m = __ballot_sync(0xffffffffu, fl == q || (fl + 1 == q && up));
                    if (wlane == q && m) mine = (__ffs(m) - 1) | ((32 - __clz(m) - (__ffs(m) - 1)) << 16);
                }
                if (wlane < 4) T.rng[j * 9 + a * 4 + wlane] = mine;
            } else {
                const unsigned m = __ballot_sync(0xffffffffu, (unsigned)fl < (unsigned)zn || ((unsigned)(fl + 1) < (unsigned)zn && up));
                if (wlane == 0) T.rng[j * 9 + 8] = m ? ((__ffs(m) - 1) | ((32 - __clz(m) - (__ffs(m) - 1)) << 16)) : 0;
            }
            if (wlane < ps) {
                T.t[(j * 3 + a) * ps + wlane] = t;
                T.fl[(j * 3 + a) * ps + wlane] = (signed char)fl;
            }
        }
    };

    // ---- accumulate a batch: every thread, its own column, box after box; no barrier inside --------------------------
    auto accumulate = [&](int jb, int nb, int c4, const bool (&von)[V]) {
        for (int j = 0; j < nb; ++j) {
            const int ry = T.rng[j * 9 + cy], rx = T.rng[j * 9 + 4 + cx], rz = T.rng[j * 9 + 8];
            const int yn = ry >> 16, xn = rx >> 16, kn = rz >> 16;
            if (yn == 0 || xn == 0 || kn == 0) continue;
            const int ys = ry & 0xffff, xs = rx & 0xffff, ks = rz & 0xffff;
            const float *ty_ = T.t + (j * 3 + 0) * ps, *tx_ = T.t + (j * 3 + 1) * ps, *tz_ = T.t + (j * 3 + 2) * ps;
            const signed char *fy_ = T.fl + (j * 3 + 0) * ps, *fx_ = T.fl + (j * 3 + 1) * ps, *fz_ = T.fl + (j * 3 + 2) * ps;
            const float *gb = grads + (size_t)T.rl[jb + j] * roi_stride + c4 * 4;
            for (int xc = xs; xc < xs + xn; xc += OS_XU) {
                float wx[OS_XU];
                int xo[OS_XU];
#pragma unroll
                for (int u = 0; u < OS_XU; ++u) {                           // past the range: re-read the last sample, weight 0
                    const bool ok = xc + u < xs + xn;
                    const int x = ok ? xc + u : xs + xn - 1;
                    const float w = os_weight(fx_[x], tx_[x], cx);
                    wx[u] = ok ? w : 0.0f;
                    xo[u] = x * pdC;
                }
                for (int k = ks; k < ks + kn; k += KU) {
                    int koff[KU];
#pragma unroll
                    for (int h = 0; h < KU; ++h) koff[h] = (k + h < ks + kn) ? h * g.C : 0;   // odd tail: re-read, unused
                    float4 s[KU][V];
#pragma unroll
                    for (int h = 0; h < KU; ++h)
#pragma unroll
                        for (int v = 0; v < V; ++v) s[h][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float *gk = gb + k * g.C;
#pragma unroll 1
                    for (int y = ys; y < ys + yn; ++y) {
                        const float wy = os_weight(fy_[y], ty_[y], cy);
                        const float *gy = gk + y * pwpdC;
                        float4 val[OS_XU][KU][V];
#pragma unroll
                        for (int u = 0; u < OS_XU; ++u)
#pragma unroll
                            for (int h = 0; h < KU; ++h)
#pragma unroll
                                for (int v = 0; v < V; ++v)
                                    if (von[v]) val[u][h][v] = ldg4(gy + xo[u] + koff[h] + v * (OS_LANES * 4));
#pragma unroll
                        for (int u = 0; u < OS_XU; ++u) {
                            const float w = __fmul_rn(wy, wx[u]);
#pragma unroll
                            for (int h = 0; h < KU; ++h)
#pragma unroll
                                for (int v = 0; v < V; ++v)
                                    if (von[v]) fma4(s[h][v], val[u][h][v], w);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < KU; ++h) {
                        if (h > 0 && k + h >= ks + kn) break;
                        const int f = fz_[k + h];
                        const float t = tz_[k + h], t0 = __fsub_rn(1.0f, t);
                        if ((unsigned)f < (unsigned)zn) {
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                float4 a = acc[(f * V + v) * OS_LANES];
                                fma4(a, s[h][v], t0);
                                acc[(f * V + v) * OS_LANES] = a;
                            }
                        }
                        if ((unsigned)(f + 1) < (unsigned)zn && t > 0.0f) {
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                float4 a = acc[((f + 1) * V + v) * OS_LANES];
                                fma4(a, s[h][v], t);
                                acc[((f + 1) * V + v) * OS_LANES] = a;
                            }
                        }
                    }
                }
            }
        }
    };

    bool reuse = false;                     // the tables in shared memory cover ALL boxes of the tile: later chunks reuse them
    int nb_all = 0;
    const int chunk_end = min(L.chunks, (cg + 1) * L.cpc);
    for (int chunk = cg * L.cpc; chunk < chunk_end; ++chunk) {
        const int c4 = chunk * OS_LANES * V + lane;                         // first float4 channel group of the thread
        bool von[V];
#pragma unroll
        for (int v = 0; v < V; ++v) von[v] = (c4 + v * OS_LANES) < g.C / 4;
        for (int i = 0; i < tz * V; ++i) acc[i * OS_LANES] = make_float4(0.f, 0.f, 0.f, 0.f);

        if (reuse) {
            accumulate(0, nb_all, c4, von);
        } else {
            int batches = 0;
            for (int base = 0; base < g.n; base += OS_THREADS) {
                // ---- which boxes can touch this tile (ascending order kept) ---------------------------------------
                const int r = base + tid;
                bool hit = false;
                if (r < g.n && __ldg(box_ind + r) == bimg) {
                    const float *b6 = boxes + (size_t)r * 6;
                    hit = os_axis_hits(__ldg(b6 + 0), __ldg(b6 + 3), g.H, g.ph, y0, y0 + OS_TY - 1) &&
                          os_axis_hits(__ldg(b6 + 1), __ldg(b6 + 4), g.W, g.pw, x0, x0 + OS_TX - 1) &&
                          os_axis_hits(__ldg(b6 + 2), __ldg(b6 + 5), g.D, g.pd, z0, z0 + zn - 1);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                __syncthreads();                                            // the previous round's list / tables are free
                if (wlane == 0) T.wcnt[warp] = __popc(bal);
                __syncthreads();
                int before = 0, nh = 0;
#pragma unroll
                for (int w = 0; w < OS_THREADS / 32; ++w) {
                    const int c = T.wcnt[w];
                    before += (w < warp) ? c : 0;
                    nh += c;
                }
                if (hit) T.rl[before + __popc(bal & ((1u << wlane) - 1u))] = r;
                __syncthreads();
                for (int jb = 0; jb < nh; jb += OS_NB) {
                    const int nb = min(OS_NB, nh - jb);
                    if (jb > 0) __syncthreads();                            // every warp is done with the previous tables
                    build_tables(jb, nb);
                    __syncthreads();
                    accumulate(jb, nb, c4, von);
                    ++batches;
                    nb_all = nb;
                }
            }
            reuse = g.n <= OS_THREADS && batches <= 1;                      // (jb was 0: the list offsets stay valid too)
            if (batches == 0) nb_all = 0;
        }

        // ---- every voxel of the tile (this chunk) is stored once ---------------------------------------------------
        const int y = y0 + cy, x = x0 + cx;
        if (y < g.H && x < g.W) {
            float *out = grad_image + ((((size_t)bimg * g.H + y) * g.W + x) * g.D + z0) * g.C + c4 * 4;
            for (int z = 0; z < zn; ++z, out += g.C) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (von[v]) *reinterpret_cast<float4 *>(out + v * (OS_LANES * 4)) = acc[(z * V + v) * OS_LANES];
            }
        }
    }
}

static size_t os_smem_bytes(int tz, int V, int ps)
{
    return (size_t)OS_TY * OS_TX * tz * V * OS_LANES * sizeof(float4) + sizeof(int) * (OS_THREADS + 8) +
           sizeof(int) * OS_NB * 9 + sizeof(float) * OS_NB * 3 * ps + (((size_t)OS_NB * 3 * ps + 15) & ~size_t(15));
}

bool car3d_grad_image_os_supported(const CarGeom &g)
{
    return g.C % 4 == 0 && g.ph <= OS_MAXP && g.pw <= OS_MAXP && g.pd <= OS_MAXP &&
           (long long)g.ph * g.pw * g.pd * g.C < (1ll << 31);
}

static void os_plan(const CarGeom &g, OsLaunch &L, int &V, size_t &smem)
{
    V = (g.C / 4 > OS_LANES) ? 2 : 1;
    const int forced = option_value(OPT_CAR_V);
    if (forced == 1 || forced == 2) V = forced;
    if (g.C / 4 <= OS_LANES) V = 1;
    L.tz = option_value(OPT_OS_TZ) > 0 ? option_value(OPT_OS_TZ) : (V == 2 ? 8 : 16);
    L.tz = max(1, min(L.tz, min(g.D, 64)));
    L.ps = max(g.ph, max(g.pw, g.pd));
    while (os_smem_bytes(L.tz, V, L.ps) > 200 * 1024 && L.tz > 1) L.tz /= 2;
    smem = os_smem_bytes(L.tz, V, L.ps);
    L.chunks = (g.C / 4 + OS_LANES * V - 1) / (OS_LANES * V);
    L.ty_tiles = (g.H + OS_TY - 1) / OS_TY;
    L.tx_tiles = (g.W + OS_TX - 1) / OS_TX;
    L.tz_tiles = (g.D + L.tz - 1) / L.tz;
    // chunks of a tile share the box list and the tables: put as many in one CTA as still leaves ~8 CTAs per SM
    const long long tiles = (long long)g.B * L.ty_tiles * L.tx_tiles * L.tz_tiles;
    L.cpc = option_value(OPT_OS_CPC) > 0 ? option_value(OPT_OS_CPC) : L.chunks;
    L.cpc = max(1, min(L.cpc, L.chunks));
    if (option_value(OPT_OS_CPC) <= 0)
        while (L.cpc > 1 && tiles * ((L.chunks + L.cpc - 1) / L.cpc) < 8ll * num_sms()) --L.cpc;
    L.cgroups = (L.chunks + L.cpc - 1) / L.cpc;
}

// CTAs the launch would have: the caller's auto rule prefers the scatter kernel when the output is too small to
// fill the GPU with tiles
long long car3d_grad_image_os_ctas(const CarGeom &g)
{
    OsLaunch L; int V; size_t smem;
    os_plan(g, L, V, smem);
    return (long long)g.B * L.ty_tiles * L.tx_tiles * L.tz_tiles * L.cgroups;
}

int launch_car3d_grad_image_os(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                               float *grad_image, cudaStream_t stream)
{
    if (!car3d_grad_image_os_supported(g)) return ROI3D_EUNSUPPORTED;
    OsLaunch L; int V; size_t smem;
    os_plan(g, L, V, smem);
    if (smem > 200 * 1024) return ROI3D_EUNSUPPORTED;
    const long long grid = (long long)g.B * L.ty_tiles * L.tx_tiles * L.tz_tiles * L.cgroups;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    using Kern = void (*)(const float *, const float *, const int *, CarGeom, OsLaunch, float *);
    Kern kern;
    switch (option_value(OPT_OS_SHAPE)) {                                  // experiment knob; 0 = production choice
    case 1: kern = (V == 2) ? car3d_grad_image_os_kernel<2, 1, 2, 3> : car3d_grad_image_os_kernel<1, 2, 2, 3>; break;
    case 2: kern = (V == 2) ? car3d_grad_image_os_kernel<2, 2, 2, 2> : car3d_grad_image_os_kernel<1, 1, 4, 3>; break;
    default: kern = (V == 2) ? car3d_grad_image_os_kernel<2, 1, 4, 2> : car3d_grad_image_os_kernel<1, 2, 4, 2>; break;
    }
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    kern<<<(unsigned)grid, OS_THREADS, smem, stream>>>(grads, boxes, box_ind, g, L, grad_image);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

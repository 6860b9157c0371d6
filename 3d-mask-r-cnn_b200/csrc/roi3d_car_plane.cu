// roi3d_car_plane.cu -- plane-staged separable CropAndResize3D (variant 2), the
// production kernels for trilinear sampling with C % 4 == 0.
//
// Why: trilinear taps form a product set.  For one ROI and one depth sample k the
// reference (CAR.so@0x4f88-0x5021) first lerps along z, then x, then y, each as
// a + (b - a) * t.  The z-lerp of a voxel column (yi, xi) depends only on (yi, xi, k),
// the x-lerp only on (yi, x-sample, k); so a CTA that owns (ROI, k, channel chunk)
//   A. reads every footprint voxel (yi, xi, {floor z, ceil z}) exactly ONCE from global
//      memory (coalesced 16-byte loads along channels), z-lerps it in registers and
//      parks the result in a shared-memory plane Z[row][col][chunk];
//   B. produces all ph*pw outputs of that k from 4 shared-memory reads each (x-lerp,
//      then y-lerp) and streams them out with 16-byte evict-first stores.
// The intermediate values are the very expressions the reference evaluates, in the same
// order and rounding, so the result is bit-identical to the direct kernel and to the
// oracle, while L2->SM traffic drops from 8 taps per output to (2 * footprint / outputs).
// The backward kernel is the transpose: gather grads into the plane (deterministic
// shared-memory sums), then 2 vector REDs per footprint voxel instead of 8 per element.
//
// Per-ROI metadata (sample -> footprint position, lerp weight, validity, footprint voxel
// lists, y-tiles) is built once per CTA in shared memory.
#include "roi3d_common.cuh"

namespace roi3d {

constexpr int PL_THREADS = 256;
constexpr int PL_MAXP = 64;                 // max crop size per axis handled here
constexpr int PL_UNROLL = 4;

struct AxisTab {                            // one per axis (y, x), lives in shared memory
    int   pos0[PL_MAXP];                    // footprint position of floor(in)   (-1 if sample invalid)
    int   pos1[PL_MAXP];                    // footprint position of ceil(in)
    float t[PL_MAXP];                       // lerp weight in - floor(in)
    int   list[2 * PL_MAXP];                // footprint voxel indices, in first-occurrence (monotonic) order
    int   cand[2 * PL_MAXP];                // scratch: floor/ceil per sample
    int   first[2 * PL_MAXP];               // scratch: 1 if first occurrence
    int   n;                                // footprint size
};

struct PlaneShared {
    AxisTab ax[2];
    int tile_y0[PL_MAXP + 1];               // y-sample tile boundaries
    int ntiles;
};

// Build pos0/pos1/t/list for both axes.  tid in [0,64): y sample, [64,128): x sample.
__device__ __forceinline__ void build_axis_tables(PlaneShared &S, const float *box, const CarGeom &g)
{
    const int tid = threadIdx.x;
    if (tid < 2 * PL_MAXP) {
        const int a = tid / PL_MAXP, k = tid % PL_MAXP;
        const int p = a ? g.pw : g.ph, dim = a ? g.W : g.H;
        AxisTab &T = S.ax[a];
        int c0 = -1, c1 = -1;
        if (k < p) {
            const float a1 = box[a], a2 = box[3 + a];
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
            if (j & 1) T.pos1[j >> 1] = pos; else T.pos0[j >> 1] = pos;
        }
        if (j == 2 * p - 1) {
            int cnt = 0;
            for (int i = 0; i < 2 * p; ++i) cnt += T.first[i];
            T.n = cnt;
        }
    }
    __syncthreads();
}

// Greedy y-sample tiles: rows spanned by a tile times n_x must fit the plane.
__device__ __forceinline__ void build_y_tiles(PlaneShared &S, const CarGeom &g, int zcap)
{
    if (threadIdx.x == 0) {
        const AxisTab &Y = S.ax[0];
        const int nx = max(S.ax[1].n, 1);
        int nt = 0, lo = 1 << 30, hi = -1;
        S.tile_y0[0] = 0;
        for (int y = 0; y < g.ph; ++y) {
            if (Y.pos0[y] < 0) continue;
            const int l2 = min(lo, min(Y.pos0[y], Y.pos1[y])), h2 = max(hi, max(Y.pos0[y], Y.pos1[y]));
            if (hi >= 0 && (h2 - l2 + 1) * nx > zcap) {      // close the tile before y
                S.tile_y0[++nt] = y;
                lo = min(Y.pos0[y], Y.pos1[y]);
                hi = max(Y.pos0[y], Y.pos1[y]);
            } else { lo = l2; hi = h2; }
        }
        S.tile_y0[++nt] = g.ph;
        S.ntiles = nt;
    }
    __syncthreads();
}

// rows [r0, r1] used by the valid samples of y-tile [ya, yb); r1 < r0 if none
__device__ __forceinline__ void tile_rows(const AxisTab &Y, int ya, int yb, int &r0, int &r1)
{
    r0 = 1 << 30; r1 = -1;
    for (int y = ya; y < yb; ++y) {
        const int a = Y.pos0[y], b = Y.pos1[y];
        if (a < 0) continue;
        r0 = min(r0, min(a, b));
        r1 = max(r1, max(a, b));
    }
}

struct PlaneLaunch {
    int cl;            // channel lanes (float4 each) per voxel
    int chunks;        // channel chunks per voxel
    int ksplits;       // depth-sample splits per ROI
    int zcap;          // plane capacity in voxels
};

// ---------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_THREADS)
car3d_fwd_plane_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                       const int *__restrict__ box_index, CarGeom g, PlaneLaunch L, float ext,
                       float *__restrict__ crops)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);
    float4 *Z = reinterpret_cast<float4 *>(smem_raw + ((sizeof(PlaneShared) + 15) & ~size_t(15)));
    __shared__ float s_box[6];

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    if (threadIdx.x < 6) s_box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
    __syncthreads();
    build_axis_tables(S, s_box, g);
    build_y_tiles(S, g, L.zcap);

    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    const int nx = X.n;
    const int lane = threadIdx.x % L.cl, slot = threadIdx.x / L.cl, vs = PL_THREADS / L.cl;
    const int c4 = chunk * L.cl + lane;                     // float4 channel group
    const bool lane_on = c4 < g.C / 4;
    const long long sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    const float *img = image + (long long)__ldg(box_index + b) * g.H * sH + c4 * 4;
    float *crop = crops + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float4 ext4 = make_float4(ext, ext, ext, ext);
    const float z1 = s_box[2], z2 = s_box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);

    for (int k = k0; k < k1; ++k) {
        const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
        const bool zvalid = !axis_invalid(in_z, g.D);
        const float zfl = floorf(in_z);
        const int zf = (int)zfl, zc = (int)ceilf(in_z);
        const float zl = __fsub_rn(in_z, zfl);
        for (int tl = 0; tl < S.ntiles; ++tl) {
            const int ya = S.tile_y0[tl], yb = S.tile_y0[tl + 1];
            int r0, r1;
            tile_rows(Y, ya, yb, r0, r1);
            const int nvox = (zvalid && r1 >= r0) ? (r1 - r0 + 1) * nx : 0;
            // ---- stage A: footprint voxels -> z-lerp -> shared plane ----------------
            if (lane_on) {
                for (int base = slot; base < nvox; base += vs * PL_UNROLL) {
                    float4 f[PL_UNROLL], c[PL_UNROLL];
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u) {
                        const int idx = base + u * vs;
                        if (idx < nvox) {
                            const int r = idx / nx, cx = idx - r * nx;
                            const float *p = img + Y.list[r0 + r] * sH + X.list[cx] * sW;
                            f[u] = ldg4(p + (long long)zf * g.C);
                            c[u] = ldg4(p + (long long)zc * g.C);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u) {
                        const int idx = base + u * vs;
                        if (idx < nvox) Z[idx * L.cl + lane] = lerp_rn(f[u], c[u], zl);
                    }
                }
            }
            __syncthreads();
            // ---- stage B: x-lerp, y-lerp from the plane -> crops ----------------------
            if (lane_on) {
                const int nout = (yb - ya) * g.pw;
                for (int idx = slot; idx < nout; idx += vs) {
                    const int yy = idx / g.pw, x = idx - yy * g.pw, y = ya + yy;
                    float *o = crop + (((long long)y * g.pw + x) * g.pd + k) * g.C;
                    const int py0 = Y.pos0[y], px0 = X.pos0[x];
                    if (!zvalid || py0 < 0 || px0 < 0) { st_stream4(o, ext4); continue; }
                    const int rt = (py0 - r0) * nx, rb = (Y.pos1[y] - r0) * nx, px1 = X.pos1[x];
                    const float xl = X.t[x], yl = Y.t[y];
                    const float4 tlv = Z[(rt + px0) * L.cl + lane], trv = Z[(rt + px1) * L.cl + lane];
                    const float4 blv = Z[(rb + px0) * L.cl + lane], brv = Z[(rb + px1) * L.cl + lane];
                    const float4 top = lerp_rn(tlv, trv, xl), bot = lerp_rn(blv, brv, xl);
                    st_stream4(o, lerp_rn(top, bot, yl));
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------
// backward (grad image): transpose of the above.
//   A'. plane G[row][col] = sum over the samples (y,x) that tap (row,col) of
//       (wy * wx) * grads[y][x][k]   -- a gather with deterministic order, so no
//       shared-memory atomics;
//   B'. two vector REDs per plane voxel: G * (1 - zl) -> floor z, G * zl -> ceil z.
// grad_image must be zero-filled before the launch.
// ---------------------------------------------------------------------------------
struct BwdRanges {                          // per footprint position: samples that tap it
    short a0[2 * PL_MAXP], e0[2 * PL_MAXP]; // [a0,e0): samples whose floor is this position
    short a1[2 * PL_MAXP], e1[2 * PL_MAXP]; // [a1,e1): samples whose ceil is this position
};

__device__ __forceinline__ void build_ranges(const AxisTab &T, BwdRanges &R, int p, int tid0, int nthreads)
{
    // samples with a given floor (ceil) position are contiguous because `in` is monotonic
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

__device__ __forceinline__ float4 fma4(const float4 a, float w, const float4 acc) {
    return make_float4(acc.x + a.x * w, acc.y + a.y * w, acc.z + a.z * w, acc.w + a.w * w);
}

__global__ void __launch_bounds__(PL_THREADS)
car3d_grad_image_plane_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                              const int *__restrict__ box_ind, CarGeom g, PlaneLaunch L,
                              float *__restrict__ grad_image)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PlaneShared &S = *reinterpret_cast<PlaneShared *>(smem_raw);
    size_t off = (sizeof(PlaneShared) + 15) & ~size_t(15);
    BwdRanges *RY = reinterpret_cast<BwdRanges *>(smem_raw + off); off += (sizeof(BwdRanges) + 15) & ~size_t(15);
    BwdRanges *RX = reinterpret_cast<BwdRanges *>(smem_raw + off); off += (sizeof(BwdRanges) + 15) & ~size_t(15);
    float4 *T = reinterpret_cast<float4 *>(smem_raw + off);        // [rows_y_samples][nx][cl] x-reduced
    __shared__ float s_box[6];

    int bid = blockIdx.x;
    const int chunk = bid % L.chunks; bid /= L.chunks;
    const int ks = bid % L.ksplits;
    const int b = bid / L.ksplits;
    if (threadIdx.x < 6) s_box[threadIdx.x] = __ldg(boxes + (size_t)b * 6 + threadIdx.x);
    __syncthreads();
    build_axis_tables(S, s_box, g);
    const AxisTab &Y = S.ax[0], &X = S.ax[1];
    build_ranges(Y, *RY, g.ph, threadIdx.x, PL_THREADS);
    build_ranges(X, *RX, g.pw, threadIdx.x, PL_THREADS);
    // y-sample tiles: T holds (samples in tile) * nx entries
    if (threadIdx.x == 0) {
        const int nx = max(X.n, 1);
        const int ty = max(1, L.zcap / nx);
        int nt = 0;
        for (int y = 0; y < g.ph; y += ty) S.tile_y0[nt++] = y;
        S.tile_y0[nt] = g.ph;
        S.ntiles = nt;
    }
    __syncthreads();

    const int nx = X.n, ny = Y.n;
    const int lane = threadIdx.x % L.cl, slot = threadIdx.x / L.cl, vs = PL_THREADS / L.cl;
    const int c4 = chunk * L.cl + lane;
    const bool lane_on = c4 < g.C / 4;
    const long long sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    float *img = grad_image + (long long)__ldg(box_ind + b) * g.H * sH + c4 * 4;
    const float *gcrop = grads + (long long)b * g.ph * g.pw * g.pd * g.C + c4 * 4;
    const float z1 = s_box[2], z2 = s_box[5];
    const float zscale = axis_scale(z1, z2, g.D, g.pd);
    const int kper = (g.pd + L.ksplits - 1) / L.ksplits;
    const int k0 = ks * kper, k1 = min(g.pd, k0 + kper);
    if (nx == 0 || ny == 0) return;

    for (int k = k0; k < k1; ++k) {
        const float in_z = axis_coord(z1, z2, g.D, g.pd, k, zscale);
        if (axis_invalid(in_z, g.D)) continue;                 // uniform across the CTA
        const float zfl = floorf(in_z);
        const int zf = (int)zfl, zc = (int)ceilf(in_z);
        const float zl = __fsub_rn(in_z, zfl), wzf = __fsub_rn(1.0f, zl);
        for (int tl = 0; tl < S.ntiles; ++tl) {
            const int ya = S.tile_y0[tl], yb = S.tile_y0[tl + 1];
            // ---- stage A'1: T[y][col] = sum_x wx * grads[y][x][k]  (x-reduction) ------
            if (lane_on) {
                const int nent = (yb - ya) * nx;
                for (int idx = slot; idx < nent; idx += vs) {
                    const int yy = idx / nx, cx = idx - yy * nx, y = ya + yy;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (Y.pos0[y] >= 0) {
                        const float *grow = gcrop + (((long long)y * g.pw) * g.pd + k) * g.C;
                        const long long xs = (long long)g.pd * g.C;
                        for (int x = RX->a0[cx]; x < RX->e0[cx]; ++x)
                            acc = fma4(ldg4(grow + x * xs), __fsub_rn(1.0f, X.t[x]), acc);
                        for (int x = RX->a1[cx]; x < RX->e1[cx]; ++x)
                            acc = fma4(ldg4(grow + x * xs), X.t[x], acc);
                    }
                    T[idx * L.cl + lane] = acc;
                }
            }
            __syncthreads();
            // ---- stage A'2 + B': y-reduction over the tile's samples, then 2 REDs ------
            if (lane_on) {
                const int nvox = ny * nx;
                for (int idx = slot; idx < nvox; idx += vs) {
                    const int r = idx / nx, cx = idx - r * nx;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    bool any = false;
                    for (int y = max((int)RY->a0[r], ya); y < min((int)RY->e0[r], yb); ++y) {
                        acc = fma4(T[((y - ya) * nx + cx) * L.cl + lane], __fsub_rn(1.0f, Y.t[y]), acc);
                        any = true;
                    }
                    for (int y = max((int)RY->a1[r], ya); y < min((int)RY->e1[r], yb); ++y) {
                        acc = fma4(T[((y - ya) * nx + cx) * L.cl + lane], Y.t[y], acc);
                        any = true;
                    }
                    if (!any) continue;
                    float *p = img + Y.list[r] * sH + X.list[cx] * sW;
                    red_add4(p + (long long)zf * g.C, make_float4(acc.x * wzf, acc.y * wzf, acc.z * wzf, acc.w * wzf));
                    red_add4(p + (long long)zc * g.C, make_float4(acc.x * zl, acc.y * zl, acc.z * zl, acc.w * zl));
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
static PlaneLaunch plan(const CarGeom &g, size_t fixed_smem, int pref_cl, size_t *smem_bytes)
{
    PlaneLaunch L;
    const int c4 = g.C / 4;
    int cl = pref_cl;
    while (cl > 1 && cl / 2 >= c4) cl /= 2;                    // do not idle half the lanes
    const int nxmax = min(2 * g.pw, g.W);
    int zcap;
    for (;;) {
        zcap = max(2 * nxmax, (40 * 1024) / (cl * 16));
        if (fixed_smem + (size_t)zcap * cl * 16 <= 200 * 1024 || cl == 1) break;
        cl /= 2;
    }
    L.cl = cl;
    L.chunks = (c4 + cl - 1) / cl;
    L.zcap = zcap;
    const long long want = (long long)kNumSMs * 16;
    long long per = (long long)g.n * L.chunks;
    int ks = (int)((want + per - 1) / per);
    if (ks < 1) ks = 1;
    if (ks > g.pd) ks = g.pd;
    L.ksplits = ks;
    *smem_bytes = fixed_smem + (size_t)zcap * cl * 16;
    return L;
}

int launch_car3d_fwd_plane(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                           float ext, float *crops, cudaStream_t stream)
{
    size_t smem;
    const size_t fixed = (sizeof(PlaneShared) + 15) & ~size_t(15);
    const PlaneLaunch L = plan(g, fixed, 16, &smem);
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
    size_t smem;
    const size_t fixed = ((sizeof(PlaneShared) + 15) & ~size_t(15)) + 2 * ((sizeof(BwdRanges) + 15) & ~size_t(15));
    const PlaneLaunch L = plan(g, fixed, 16, &smem);
    if (smem > 48 * 1024)
        ROI3D_CUDA_TRY(cudaFuncSetAttribute(car3d_grad_image_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)g.n * L.ksplits * L.chunks;
    if (grid > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    car3d_grad_image_plane_kernel<<<(unsigned)grid, PL_THREADS, smem, stream>>>(grads, boxes, box_ind, g, L, grad_image);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

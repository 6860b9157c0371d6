// roi3d_car_direct.cu -- "direct" CropAndResize3D kernels (variant 1): one thread per
// (crop voxel, 4-channel group), every thread recomputes its sample coordinate and
// touches its 8 taps itself.  Simple and exact; used for small channel counts, for
// `nearest`, for grad-boxes, and as the in-GPU cross-check of the plane-staged
// kernels in roi3d_car_plane.cu.  All of it is CUDA: there is no CPU path.
//
// Reference behaviour restated from the binaries:
//   forward      CAR.so@0x4370-0x5a90   (lerp z, then x, then y; a + (b-a)*t)
//   grad image   GI.so@0x3a80-0x5230    (weights ((wy*wx)*wz), 8 RMW per element)
//   grad boxes   GB.so@0x3980-0x51c0
#include "roi3d_common.cuh"

namespace roi3d {

struct Sample {
    float in;
    int i0, i1;
    float t;
    bool valid;
};

__device__ __forceinline__ Sample make_sample(float a1, float a2, int dim, int p, int k) {
    Sample s;
    const float scale = axis_scale(a1, a2, dim, p);
    s.in = axis_coord(a1, a2, dim, p, k, scale);
    s.valid = !axis_invalid(s.in, dim);
    const float fl = floorf(s.in);
    s.i0 = (int)fl;
    s.i1 = (int)ceilf(s.in);
    s.t = __fsub_rn(s.in, fl);
    return s;
}

__device__ __forceinline__ Sample make_sample_scaled(float a1, float a2, int dim, int p, int k, float scale) {
    Sample s;
    s.in = axis_coord(a1, a2, dim, p, k, scale);
    s.valid = !axis_invalid(s.in, dim);
    const float fl = floorf(s.in);
    s.i0 = (int)fl;
    s.i1 = (int)ceilf(s.in);
    s.t = __fsub_rn(s.in, fl);
    return s;
}

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

template <int VEC> __device__ __forceinline__ typename VecT<VEC>::type ld(const float *p);
template <> __device__ __forceinline__ float4 ld<4>(const float *p) { return ldg4(p); }
template <> __device__ __forceinline__ float ld<1>(const float *p) { return __ldg(p); }
__device__ __forceinline__ void st(float *p, float4 v) { st_stream4(p, v); }
__device__ __forceinline__ void st(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ float4 splat4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float lerp_v(float a, float b, float t) { return lerp_rn(a, b, t); }
__device__ __forceinline__ float4 lerp_v(float4 a, float4 b, float t) { return lerp_rn(a, b, t); }
__device__ __forceinline__ float mul_v(float a, float w) { return __fmul_rn(a, w); }
__device__ __forceinline__ float4 mul_v(float4 a, float w) {
    return make_float4(__fmul_rn(a.x, w), __fmul_rn(a.y, w), __fmul_rn(a.z, w), __fmul_rn(a.w, w));
}
__device__ __forceinline__ void red(float *p, float v) { red_add1(p, v); }
__device__ __forceinline__ void red(float *p, float4 v) { red_add4(p, v); }
template <int VEC> __device__ __forceinline__ typename VecT<VEC>::type splat(float v);
template <> __device__ __forceinline__ float4 splat<4>(float v) { return splat4(v); }
template <> __device__ __forceinline__ float splat<1>(float v) { return v; }

// ---------------------------------------------------------------------------------
// forward, direct gather
// ---------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
car3d_fwd_direct_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                        const int *__restrict__ box_index, CarGeom g, int method, float ext,
                        float *__restrict__ crops)
{
    using V = typename VecT<VEC>::type;
    const int cv = g.C / VEC;                                  // channel groups per voxel
    const long long total = (long long)g.n * g.ph * g.pw * g.pd * cv;
    const long long sD = g.C, sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        long long r = idx / cv;
        const int z = (int)(r % g.pd); r /= g.pd;
        const int x = (int)(r % g.pw); r /= g.pw;
        const int y = (int)(r % g.ph);
        const int b = (int)(r / g.ph);
        const float *box = boxes + (size_t)b * 6;
        const Sample sy = make_sample(__ldg(box + 0), __ldg(box + 3), g.H, g.ph, y);
        const Sample sx = make_sample(__ldg(box + 1), __ldg(box + 4), g.W, g.pw, x);
        const Sample sz = make_sample(__ldg(box + 2), __ldg(box + 5), g.D, g.pd, z);
        float *out = crops + idx * VEC;
        const int bimg = __ldg(box_index + b);
        // a box_index outside [0, B) reads nothing: the crop is the extrapolation value (the reference would read out of bounds)
        if (!(sy.valid && sx.valid && sz.valid) || (unsigned)bimg >= (unsigned)g.B) {
            st(out, splat<VEC>(ext));
            continue;
        }
        const float *img = image + (long long)bimg * g.H * sH + c;
        if (method == ROI3D_METHOD_TRILINEAR) {
            const float *pt = img + sy.i0 * sH, *pb = img + sy.i1 * sH;
            const long long ol = sx.i0 * sW, orr = sx.i1 * sW, of = sz.i0 * sD, oc = sz.i1 * sD;
            const V tlf = ld<VEC>(pt + ol + of), tlc = ld<VEC>(pt + ol + oc);
            const V trf = ld<VEC>(pt + orr + of), trc = ld<VEC>(pt + orr + oc);
            const V blf = ld<VEC>(pb + ol + of), blc = ld<VEC>(pb + ol + oc);
            const V brf = ld<VEC>(pb + orr + of), brc = ld<VEC>(pb + orr + oc);
            const V tl = lerp_v(tlf, tlc, sz.t), tr = lerp_v(trf, trc, sz.t);
            const V bl = lerp_v(blf, blc, sz.t), br = lerp_v(brf, brc, sz.t);
            const V top = lerp_v(tl, tr, sx.t), bot = lerp_v(bl, br, sx.t);
            st(out, lerp_v(top, bot, sy.t));
        } else {
            const int yi = (int)roundf(sy.in), xi = (int)roundf(sx.in), zi = (int)roundf(sz.in);
            st(out, ld<VEC>(img + yi * sH + xi * sW + zi * sD));
        }
    }
}

// ---------------------------------------------------------------------------------
// grad image, direct scatter (grad_image must be zero-filled beforehand)
// ---------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
car3d_grad_image_direct_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                               const int *__restrict__ box_ind, CarGeom g, int method,
                               float *__restrict__ grad_image)
{
    using V = typename VecT<VEC>::type;
    const int cv = g.C / VEC;
    const long long total = (long long)g.n * g.ph * g.pw * g.pd * cv;
    const long long sD = g.C, sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        long long r = idx / cv;
        const int z = (int)(r % g.pd); r /= g.pd;
        const int x = (int)(r % g.pw); r /= g.pw;
        const int y = (int)(r % g.ph);
        const int b = (int)(r / g.ph);
        const float *box = boxes + (size_t)b * 6;
        const Sample sy = make_sample(__ldg(box + 0), __ldg(box + 3), g.H, g.ph, y);
        const Sample sx = make_sample(__ldg(box + 1), __ldg(box + 4), g.W, g.pw, x);
        const Sample sz = make_sample(__ldg(box + 2), __ldg(box + 5), g.D, g.pd, z);
        const int bimg = __ldg(box_ind + b);
        if (!(sy.valid && sx.valid && sz.valid) || (unsigned)bimg >= (unsigned)g.B) continue;   // out-of-range box_ind: nothing is scattered
        const V gv = ld<VEC>(grads + idx * VEC);
        float *img = grad_image + (long long)bimg * g.H * sH + c;
        if (method == ROI3D_METHOD_TRILINEAR) {
            const float wt = __fsub_rn(1.0f, sy.t), wb = sy.t;
            const float wl = __fsub_rn(1.0f, sx.t), wr = sx.t;
            const float wf = __fsub_rn(1.0f, sz.t), wc = sz.t;
            float *pt = img + sy.i0 * sH, *pb = img + sy.i1 * sH;
            const long long ol = sx.i0 * sW, orr = sx.i1 * sW, of = sz.i0 * sD, oc = sz.i1 * sD;
            const float wtl = __fmul_rn(wt, wl), wtr = __fmul_rn(wt, wr);
            const float wbl = __fmul_rn(wb, wl), wbr = __fmul_rn(wb, wr);
            red(pt + ol + of, mul_v(gv, __fmul_rn(wtl, wf)));
            red(pt + ol + oc, mul_v(gv, __fmul_rn(wtl, wc)));
            red(pt + orr + of, mul_v(gv, __fmul_rn(wtr, wf)));
            red(pt + orr + oc, mul_v(gv, __fmul_rn(wtr, wc)));
            red(pb + ol + of, mul_v(gv, __fmul_rn(wbl, wf)));
            red(pb + ol + oc, mul_v(gv, __fmul_rn(wbl, wc)));
            red(pb + orr + of, mul_v(gv, __fmul_rn(wbr, wf)));
            red(pb + orr + oc, mul_v(gv, __fmul_rn(wbr, wc)));
        } else {
            const int yi = (int)roundf(sy.in), xi = (int)roundf(sx.in), zi = (int)roundf(sz.in);
            red(img + yi * sH + xi * sW + zi * sD, gv);
        }
    }
}

// ---------------------------------------------------------------------------------
// grad boxes: one CTA per box, block-reduced partial sums, no atomics
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
car3d_grad_boxes_kernel(const float *__restrict__ grads, const float *__restrict__ image,
                        const float *__restrict__ boxes, const int *__restrict__ box_ind,
                        CarGeom g, float *__restrict__ grad_boxes)
{
    const int b = blockIdx.x;
    const long long sD = g.C, sW = (long long)g.D * g.C, sH = (long long)g.W * g.D * g.C;
    const float *box = boxes + (size_t)b * 6;
    const float y1 = __ldg(box + 0), x1 = __ldg(box + 1), z1 = __ldg(box + 2);
    const float y2 = __ldg(box + 3), x2 = __ldg(box + 4), z2 = __ldg(box + 5);
    const float rh = (g.ph > 1) ? __fdiv_rn((float)(g.H - 1), (float)(g.ph - 1)) : 0.0f;
    const float rw = (g.pw > 1) ? __fdiv_rn((float)(g.W - 1), (float)(g.pw - 1)) : 0.0f;
    const float rd = (g.pd > 1) ? __fdiv_rn((float)(g.D - 1), (float)(g.pd - 1)) : 0.0f;
    // This op forms the sample step as (a2 - a1) * ratio (GB.so@0x3ff6-0x4071), not as the forward's
    // ((a2 - a1) * (dim - 1)) / (p - 1).  REFERENCE QUIRK reproduced for parity: its depth step is
    // (z2 - y1) * ratio_h (GB.so@0x4059-0x4071 loads box[5], box[0] and the height ratio).
    const float hs = (g.ph > 1) ? __fmul_rn(__fsub_rn(y2, y1), rh) : 0.0f;
    const float ws = (g.pw > 1) ? __fmul_rn(__fsub_rn(x2, x1), rw) : 0.0f;
    const float ds = (g.pd > 1) ? __fmul_rn(__fsub_rn(z2, y1), rh) : 0.0f;
    const int bimg = __ldg(box_ind + b);
    const float *img = image + (long long)bimg * g.H * sH;
    const float *gb = grads + (long long)b * g.ph * g.pw * g.pd * g.C;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // a box_ind outside [0, B) reads nothing: its row of grad_boxes is zero
    const long long per_box = ((unsigned)bimg < (unsigned)g.B) ? (long long)g.ph * g.pw * g.pd * g.C : 0;
    for (long long e = threadIdx.x; e < per_box; e += blockDim.x) {
        const int c = (int)(e % g.C);
        long long r = e / g.C;
        const int z = (int)(r % g.pd); r /= g.pd;
        const int x = (int)(r % g.pw);
        const int y = (int)(r / g.pw);
        const Sample sy = make_sample_scaled(y1, y2, g.H, g.ph, y, hs);
        const Sample sx = make_sample_scaled(x1, x2, g.W, g.pw, x, ws);
        const Sample sz = make_sample_scaled(z1, z2, g.D, g.pd, z, ds);
        if (!(sy.valid && sx.valid && sz.valid)) continue;
        const float *pt = img + sy.i0 * sH + c, *pb = img + sy.i1 * sH + c;
        const long long ol = sx.i0 * sW, orr = sx.i1 * sW, of = sz.i0 * sD, oc = sz.i1 * sD;
        const float tlf = __ldg(pt + ol + of), tlc = __ldg(pt + ol + oc);
        const float trf = __ldg(pt + orr + of), trc = __ldg(pt + orr + oc);
        const float blf = __ldg(pb + ol + of), blc = __ldg(pb + ol + oc);
        const float brf = __ldg(pb + orr + of), brc = __ldg(pb + orr + oc);
        const float yl = sy.t, xl = sx.t, zl = sz.t;
        const float myl = 1.0f - yl, mxl = 1.0f - xl, mzl = 1.0f - zl;
        float gy = ((blf - tlf) * mxl + (brf - trf) * xl) * mzl + ((blc - tlc) * mxl + (brc - trc) * xl) * zl;
        float gx = ((trf - tlf) * myl + (brf - blf) * yl) * mzl + ((trc - tlc) * myl + (brc - blc) * yl) * zl;
        float gz = ((tlc - tlf) * myl + (blc - blf) * yl) * mxl + ((trc - trf) * myl + (brc - brf) * yl) * xl;
        const float tg = __ldg(gb + e);
        gy *= tg; gx *= tg; gz *= tg;
        if (g.ph > 1) { acc[0] += gy * ((float)(g.H - 1) - (float)y * rh); acc[3] += (gy * (float)y) * rh; }
        else { const float v = (float)((double)gy * 0.5 * (double)(g.H - 1)); acc[0] += v; acc[3] += v; }
        if (g.pw > 1) { acc[1] += gx * ((float)(g.W - 1) - (float)x * rw); acc[4] += (gx * (float)x) * rw; }
        else { const float v = (float)((double)gx * 0.5 * (double)(g.W - 1)); acc[1] += v; acc[4] += v; }
        if (g.pd > 1) { acc[2] += gz * ((float)(g.D - 1) - (float)z * rd); acc[5] += (gz * (float)z) * rd; }
        else { const float v = (float)((double)gz * 0.5 * (double)(g.D - 1)); acc[2] += v; acc[5] += v; }
    }
    __shared__ float red_s[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red_s[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red_s[w][threadIdx.x];
        grad_boxes[(size_t)b * 6 + threadIdx.x] = v;
    }
}

// ---------------------------------------------------------------------------------
// host launchers (called from roi3d_abi.cu)
// ---------------------------------------------------------------------------------
static inline int grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 32;             // grid-stride beyond 32 CTAs/SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_car3d_fwd_direct(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                            int method, float ext, float *crops, cudaStream_t stream)
{
    const bool vec4 = (g.C % 4 == 0) && ((reinterpret_cast<uintptr_t>(image) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(crops) & 15) == 0);
    const long long total = (long long)g.n * g.ph * g.pw * g.pd * (vec4 ? g.C / 4 : g.C);
    const int grid = grid_for(total, 256);
    if (vec4) car3d_fwd_direct_kernel<4><<<grid, 256, 0, stream>>>(image, boxes, box_index, g, method, ext, crops);
    else      car3d_fwd_direct_kernel<1><<<grid, 256, 0, stream>>>(image, boxes, box_index, g, method, ext, crops);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_car3d_grad_image_direct(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                                   int method, float *grad_image, cudaStream_t stream)
{
    const bool vec4 = (g.C % 4 == 0) && ((reinterpret_cast<uintptr_t>(grads) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(grad_image) & 15) == 0);
    const long long total = (long long)g.n * g.ph * g.pw * g.pd * (vec4 ? g.C / 4 : g.C);
    const int grid = grid_for(total, 256);
    if (vec4) car3d_grad_image_direct_kernel<4><<<grid, 256, 0, stream>>>(grads, boxes, box_ind, g, method, grad_image);
    else      car3d_grad_image_direct_kernel<1><<<grid, 256, 0, stream>>>(grads, boxes, box_ind, g, method, grad_image);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

int launch_car3d_grad_boxes(const float *grads, const float *image, const float *boxes, const int *box_ind,
                            const CarGeom &g, float *grad_boxes, cudaStream_t stream)
{
    car3d_grad_boxes_kernel<<<g.n, 256, 0, stream>>>(grads, image, boxes, box_ind, g, grad_boxes);
    ROI3D_LAUNCH_CHECK();
    return ROI3D_OK;
}

}  // namespace roi3d

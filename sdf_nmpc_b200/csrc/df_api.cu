// Ground truth from range / depth images: collision labels (ColChecker, sdf_nmpc/utils/collision_checker.py:23-125) and the
// signed / unsigned distance-field labeller built on them (DfComputer, sdf_nmpc/utils/df_computer.py:56-221).  The reference runs
// these as Warp kernels + torch glue that materialise [points x voxels] / [points x pixels] arrays; here a CTA owns a point and
// reduces over its voxels / pixels on chip, so the only HBM traffic is the image, the point and the voxel table.
#include <cmath>
#include <cstdint>
#include "handles.h"

using namespace sdfnmpc;

namespace {

struct ColParams {
    const float* img;     // [B][H][W], normalised by dmax
    int H, W;
    float dmax, hfov, vfov, safe_ball;
    int outside, is_depth, is_spherical;
};

// ColChecker._kernel_colcheck (collision_checker.py:23-91) for one point of image ii
__device__ __forceinline__ bool colcheck_point(const ColParams& c, float px, float py, float pz, long long ii) {
    const float len = sqrtf(px * px + py * py + pz * pz);
    if (!(len > c.safe_ball)) return false;                    // inside the safe ball around the sensor: never in collision
    const float val = c.is_depth ? px : len;
    if (val >= c.dmax) return true;                            // past dmax is unsafe by default
    float az = atan2f(py, px);
    float el = c.is_spherical ? atan2f(pz, sqrtf(px * px + py * py)) : atan2f(pz, px);
    if (c.outside == 2) {                                      // extrapolate: clamp to the frustum
        az = fminf(fmaxf(az, -c.hfov), c.hfov);
        el = fminf(fmaxf(el, -c.vfov), c.vfov);
    } else if (fabsf(az) >= c.hfov || fabsf(el) >= c.vfov) {
        return c.outside == 1;
    }
    const float w = (float)c.W, h = (float)c.H;
    int u, v;
    if (c.is_spherical) {
        u = (int)(w / 2.f * (1.f - az / c.hfov));
        v = (int)(h / 2.f * (1.f - el / c.vfov));
    } else {
        u = (int)(w / 2.f * (1.f - tanf(az) / tanf(c.hfov)));
        v = (int)(h / 2.f * (1.f - tanf(el) / tanf(c.vfov)));
    }
    u = min(max(u, 0), c.W - 1);
    v = min(max(v, 0), c.H - 1);
    return val >= __ldg(c.img + (ii * c.H + v) * c.W + u) * c.dmax;
}

__global__ void colcheck_kernel(ColParams c, const float* __restrict__ pts, const int* __restrict__ p2i, long long n, long long pts_per_img,
                                unsigned char* __restrict__ out) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const long long ii = p2i ? (long long)p2i[tid] : tid / pts_per_img;
    out[tid] = colcheck_point(c, pts[tid * 3], pts[tid * 3 + 1], pts[tid * 3 + 2], ii) ? 1 : 0;
}

// block-wide lexicographic minimum of (value, index): the smallest value, and among equal values the smallest index
// (what torch.min(dim) returns on the CPU).  Result valid in thread 0.
constexpr int DF_THREADS = 256;
__device__ __forceinline__ void block_argmin(float& v, int& i) {
    __shared__ float sv[DF_THREADS / 32];
    __shared__ int si[DF_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = v; si[warp] = i; }
    __syncthreads();
    if (warp == 0) {
        v = lane < DF_THREADS / 32 ? sv[lane] : INFINITY;
        i = lane < DF_THREADS / 32 ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
    }
}

// DfComputer.get_sdf (df_computer.py:198-221): the sign of a point is its own collision label ('extrapolate', no safe ball); its
// distance is the smallest |offset| among the voxels of the multi-resolution ball around it whose label differs from the point's
// (df_computer.py:187-195: occupied voxels for a free point, free voxels for a point inside an obstacle), max_df if there is none;
// sdf = clamp(sign * dist, min_df, max_df), gradient = -sign * offset / |offset|, zero where the value saturates.
__global__ void __launch_bounds__(DF_THREADS)
df_signed_kernel(ColParams c, const float* __restrict__ pts, const int* __restrict__ p2i, long long pts_per_img,
                 const float* __restrict__ grid /*[G][3]*/, const float* __restrict__ gdist /*[G]*/, int G, float min_df, float max_df,
                 float* __restrict__ sdf, float* __restrict__ grad) {
    const long long t = blockIdx.x;
    const float px = pts[t * 3], py = pts[t * 3 + 1], pz = pts[t * 3 + 2];
    const long long ii = p2i ? (long long)p2i[t] : t / pts_per_img;
    const bool inside = colcheck_point(c, px, py, pz, ii);
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int g = threadIdx.x; g < G; g += DF_THREADS) {
        const float gx = __ldg(grid + 3 * g), gy = __ldg(grid + 3 * g + 1), gz = __ldg(grid + 3 * g + 2);
        const bool occ = colcheck_point(c, px + gx, py + gy, pz + gz, ii);
        const float d = occ != inside ? __ldg(gdist + g) : max_df;
        if (d < best) { best = d; bi = g; }            // g ascends within a thread: the first of equal values is kept
    }
    block_argmin(best, bi);
    if (threadIdx.x == 0) {
        const float sign = inside ? -1.f : 1.f;
        const float v = fminf(fmaxf(sign * best, min_df), max_df);
        const float gx = grid[3 * bi], gy = grid[3 * bi + 1], gz = grid[3 * bi + 2];
        const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
        const bool sat = v == min_df || v == max_df;
        sdf[t] = v;
        // -sign * where(saturated, 0, dir): keeps the reference's signed zeros out of the comparison by writing +0
        grad[t * 3] = sat ? 0.f : -sign * (gx / nrm);
        grad[t * 3 + 1] = sat ? 0.f : -sign * (gy / nrm);
        grad[t * 3 + 2] = sat ? 0.f : -sign * (gz / nrm);
    }
}

// df_computer.py:138-146: k x k min-pool that disregards zeros (invalid pixels) unless the whole window is zero
__global__ void df_minpool_kernel(int H, int W, int k, float dmax, long long total, const float* __restrict__ img, float* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int Wp = W / k, Hp = H / k;
    const int x = (int)(gid % Wp), y = (int)((gid / Wp) % Hp);
    const long long b = gid / ((long long)Wp * Hp);
    const float* p = img + ((size_t)b * H + (size_t)y * k) * W + (size_t)x * k;
    float m = INFINITY;
    bool any = false;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const float v = p[(size_t)i * W + j];
            any |= v != 0.f;
            m = fminf(m, v != 0.f ? v : dmax);
        }
    out[gid] = any ? m : 0.f;
}

// DfComputer._kernel_pixel_wise_udf + the reduction of get_udf (df_computer.py:84-176) for one point: distance to the 3-D position
// of every (pooled) pixel, or to the virtual wall at dmax when that is closer; minimum over the pixels and the vector towards it.
// Statement for statement, including the reference's choices: the pinhole direction is used when is_spherical is SET and the
// spherical one otherwise (df_computer.py:112-122), "invalid" is tested on the direction's x component (:126), and the wall
// case stores the ABSOLUTE vector (dmax, p_y, p_z) as the gradient direction (:136).
__global__ void __launch_bounds__(DF_THREADS)
df_unsigned_kernel(const float* __restrict__ img /*[B][H][W] pooled*/, int H, int W, const float* __restrict__ pts, const int* __restrict__ p2i,
                   long long pts_per_img, float dmax, float hfov, float vfov, int is_depth, int is_spherical, float max_df,
                   float* __restrict__ udf, float* __restrict__ grad) {
    const long long t = blockIdx.x;
    const float px = pts[t * 3], py = pts[t * 3 + 1], pz = pts[t * 3 + 2];
    const long long ii = p2i ? (long long)p2i[t] : t / pts_per_img;
    const float* im = img + (size_t)ii * H * W;
    const float w = (float)W, h = (float)H;
    const float d_bg = is_depth ? dmax - px : dmax - sqrtf(px * px + py * py + pz * pz);
    const float th = tanf(hfov), tv = tanf(vfov);
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int q = threadIdx.x; q < H * W; q += DF_THREADS) {
        const int ph = q % W, pv = q / W;
        float x, y, z;
        if (is_spherical) {
            x = 1.f;
            y = th * (1.f - 2.f * (float)ph / w);
            z = tv * (1.f - 2.f * (float)pv / h);
        } else {
            const float az = hfov * (1.f - 2.f * (float)ph / w), el = vfov * (1.f - 2.f * (float)pv / h);
            x = cosf(el) * cosf(az);
            y = cosf(el) * sinf(az);
            z = sinf(el);
        }
        const float r = __ldg(im + q);
        float d;
        if (x == 0.f) {
            d = dmax;
        } else {
            const float rx = x * r * dmax - px, ry = y * r * dmax - py, rz = z * r * dmax - pz;
            const float dp = sqrtf(rx * rx + ry * ry + rz * rz);
            d = dp > d_bg ? d_bg : dp;
        }
        if (d < best) { best = d; bi = q; }
    }
    block_argmin(best, bi);
    if (threadIdx.x == 0) {
        // recompute the vector of the winning pixel
        const int ph = bi % W, pv = bi / W;
        float x, y, z;
        if (is_spherical) {
            x = 1.f; y = th * (1.f - 2.f * (float)ph / w); z = tv * (1.f - 2.f * (float)pv / h);
        } else {
            const float az = hfov * (1.f - 2.f * (float)ph / w), el = vfov * (1.f - 2.f * (float)pv / h);
            x = cosf(el) * cosf(az); y = cosf(el) * sinf(az); z = sinf(el);
        }
        const float r = im[bi];
        float rx = x * r * dmax - px, ry = y * r * dmax - py, rz = z * r * dmax - pz;
        if (x != 0.f && sqrtf(rx * rx + ry * ry + rz * rz) > d_bg) { rx = dmax; ry = py; rz = pz; }
        const float v = fminf(fmaxf(best, 0.f), max_df);
        const float nrm = sqrtf(rx * rx + ry * ry + rz * rz);
        const bool sat = v == max_df;
        udf[t] = v;
        grad[t * 3] = sat ? 0.f : -(rx / nrm);
        grad[t * 3 + 1] = sat ? 0.f : -(ry / nrm);
        grad[t * 3 + 2] = sat ? 0.f : -(rz / nrm);
    }
}

}  // namespace

extern "C" int sdfnmpc_colcheck(const float* imgs, int32_t B, int32_t H, int32_t W, const float* pts, const int32_t* p2i, int64_t n,
                                float dmax, float hfov, float vfov, float safe_ball, int32_t outside, int32_t is_depth,
                                int32_t is_spherical, uint8_t* out, void* stream) {
    if (n == 0) return 0;
    if (!imgs || !pts || !out || B < 1 || H < 1 || W < 1 || n < 0 || outside < 0 || outside > 2) return SDFNMPC_E_INVALID;
    if (int rc = ensure_device()) return rc;
    const long long ppi = (n + B - 1) / B;
    const ColParams c{imgs, H, W, dmax, hfov, vfov, safe_ball, outside, is_depth, is_spherical};
    colcheck_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, pts, p2i, n, ppi, out);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int sdfnmpc_df_signed(const float* imgs, int32_t B, int32_t H, int32_t W, const float* pts, const int32_t* p2i, int64_t n,
                                 const float* grid, const float* grid_dist, int32_t G, float dmax, float hfov, float vfov,
                                 int32_t is_depth, int32_t is_spherical, float min_df, float max_df, float* sdf, float* grad,
                                 void* stream) {
    if (n == 0) return 0;
    if (!imgs || !pts || !grid || !grid_dist || !sdf || !grad || B < 1 || H < 1 || W < 1 || n < 0 || n > 0x7fffffffLL || G < 1 ||
        !(min_df < max_df))
        return SDFNMPC_E_INVALID;
    if (int rc = ensure_device()) return rc;
    const long long ppi = (n + B - 1) / B;
    const ColParams c{imgs, H, W, dmax, hfov, vfov, 0.f, 2, is_depth, is_spherical};       // df_computer.py:24
    df_signed_kernel<<<(unsigned)n, DF_THREADS, 0, (cudaStream_t)stream>>>(c, pts, p2i, ppi, grid, grid_dist, G, min_df, max_df, sdf, grad);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int sdfnmpc_df_unsigned(const float* imgs, int32_t B, int32_t H, int32_t W, int32_t pool, float* pooled, const float* pts,
                                   const int32_t* p2i, int64_t n, float dmax, float hfov, float vfov, int32_t is_depth,
                                   int32_t is_spherical, float max_df, float* udf, float* grad, void* stream) {
    if (n == 0) return 0;
    if (!imgs || !pooled || !pts || !udf || !grad || B < 1 || H < 1 || W < 1 || pool < 1 || H % pool || W % pool || n < 0 ||
        n > 0x7fffffffLL)
        return SDFNMPC_E_INVALID;
    if (int rc = ensure_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Hp = H / pool, Wp = W / pool;
    const long long total = (long long)B * Hp * Wp;
    df_minpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(H, W, pool, dmax, total, imgs, pooled);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    const long long ppi = (n + B - 1) / B;
    df_unsigned_kernel<<<(unsigned)n, DF_THREADS, 0, st>>>(pooled, Hp, Wp, pts, p2i, ppi, dmax, hfov, vfov, is_depth, is_spherical, max_df, udf, grad);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// lk_kernel.cu -- cv::calcOpticalFlowPyrLK as called by Frame::RmDynamicPointWithSemanticAndGeometry (src/Frame.cc:445):
// I = current gray, J = previous gray, window 21x21, 4 pyramid levels, <= 30 iterations or |delta|^2 <= 1e-4.
// Follows OpenCV video/lkpyramid.cpp: cv::pyrDown levels (bit-exact integers), Scharr derivatives of I (int16, zero outside the
// image), 14-bit fixed-point bilinear window extraction, float 2x2 solve.  Sums of integer products are accumulated EXACTLY
// (int64) and rounded once; OpenCV accumulates them in float in SIMD order, so positions agree to ~1e-4 px, not bit-for-bit
// (tolerance stated in tests/test_gpu_lk.py).  status/err are not produced (the reference ignores them, quirk Q4); a level
// that fails leaves the running estimate untouched.
//
// One warp per point; the 21x21 window is tiled over the lanes (7x2 pixels each + one pixel of the last column) and lives in
// registers for the whole level; no shared memory.
#include <cuda_runtime.h>

#include <vector>

#include "sgs_common.h"

namespace sgs {

constexpr int kWin = 21, kLkMaxLevel = 3, kLkMaxCount = 30;
constexpr int kLkWarps = 4;

struct LkLevels {
    const uint8_t* I[kLkMaxLevel + 1]; const uint8_t* J[kLkMaxLevel + 1];
    int32_t w[kLkMaxLevel + 1], h[kLkMaxLevel + 1], pitch[kLkMaxLevel + 1];
    int64_t fstride[kLkMaxLevel + 1];
    int32_t pitchJ0; int64_t fstrideJ0;     // level 0 of J may have its own layout (caller buffers)
    int32_t max_level;
};

__device__ __forceinline__ int lk_refl(int i, int n) {
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

// cv::pyrDown, CV_8U: separable [1 4 6 4 1], exact integer sums, (v + 128) >> 8, BORDER_REFLECT_101
__global__ void __launch_bounds__(256) lk_pyrdown_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, int64_t sfstride,
                                                         uint8_t* __restrict__ dst, int dw, int dh, int dpitch, int64_t dfstride) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const uint8_t* S = src + (int64_t)blockIdx.z * sfstride;
    int cx[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) cx[k] = lk_refl(2 * x - 2 + k, sw);
    int v = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const uint8_t* row = S + (int64_t)lk_refl(2 * y - 2 + r, sh) * spitch;
        const int hsum = __ldg(row + cx[2]) * 6 + (__ldg(row + cx[1]) + __ldg(row + cx[3])) * 4 + __ldg(row + cx[0]) + __ldg(row + cx[4]);
        v += hsum * (r == 2 ? 6 : (r == 1 || r == 3) ? 4 : 1);
    }
    dst[(int64_t)blockIdx.z * dfstride + (int64_t)y * dpitch + x] = (uint8_t)((v + 128) >> 8);
}

// Exact warp sums of per-lane int32 partials.  |partial| < 15 * 8160 * 4080 < 2^28.9, so the first stages stay in int32 (4 lanes for
// the mismatch sums, 8 lanes for the gradient sums whose terms are < 4080^2) before widening.
__device__ __forceinline__ long long warp_sum_grad(int v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 4);
    long long w = v;
    w += __shfl_xor_sync(0xffffffffu, w, 2); w += __shfl_xor_sync(0xffffffffu, w, 1);
    return w;
}
// two sums at once: after the first exchange lanes 0..15 carry s1 and lanes 16..31 carry s2; every lane gets both totals as float
__device__ __forceinline__ void warp_sum_pair(int s1, int s2, int lane, float scale, float& B1, float& B2) {
    const bool hi = lane >= 16;
    const int give = hi ? s1 : s2, keep = hi ? s2 : s1;
    int v = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    long long w = v;
    w += __shfl_xor_sync(0xffffffffu, w, 4); w += __shfl_xor_sync(0xffffffffu, w, 2); w += __shfl_xor_sync(0xffffffffu, w, 1);
    const float mine = __fmul_rn((float)w, scale), other = __shfl_xor_sync(0xffffffffu, mine, 16);
    B1 = hi ? other : mine; B2 = hi ? mine : other;
}

__device__ __forceinline__ int lk_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), __fsub_rn(1.f, b)), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.f, b)), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}


// Window set-up of one level, in the tiled ownership of lk_mismatch_tiles (see there): each lane loads the 10x5 raw bytes of I under
// its 7x2 window pixels (reflect-101 outside the image), forms the Scharr rows it needs (8x3; zero where the derivative position is
// outside the image, BORDER_CONSTANT) and interpolates I, Ix, Iy with the 14-bit weights; lanes 0..20 do the same for their pixel of
// window column 20.  Outputs stay in registers: C = 256 - 512 Iw, GX = Ix, GY = Iy; s11/s12/s22 are this lane's share of the
// gradient matrix (exact in int32: 15 terms of < 2^24.1).  kInside: the whole 24x24 raw patch lies inside the image.
template <bool kInside>
__device__ __forceinline__ void lk_setup_tiles(const uint8_t* __restrict__ img, int pitch, int lw, int lh, int ipx, int ipy,
                                               int k2, int g7, int erow, int lane, int w00, int w01, int w10, int w11,
                                               int (&C)[14], int (&GX)[14], int (&GY)[14], int& Ce, int& GXe, int& GYe,
                                               int& s11, int& s12, int& s22) {
    int R[10][5], E[4][4];
    if (kInside) {
        const uint8_t* p = img + (int64_t)(ipy - 1 + g7) * pitch + (ipx - 1 + k2);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
#pragma unroll
            for (int x = 0; x < 5; ++x) R[r][x] = __ldg(p + x);
            p += pitch;
        }
        const uint8_t* q = img + (int64_t)(ipy - 1 + erow) * pitch + (ipx + 19);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int x = 0; x < 4; ++x) E[r][x] = __ldg(q + x);
            q += pitch;
        }
    } else {
        int cx[5], ce[4];
#pragma unroll
        for (int x = 0; x < 5; ++x) cx[x] = lk_refl(ipx - 1 + k2 + x, lw);
#pragma unroll
        for (int x = 0; x < 4; ++x) ce[x] = lk_refl(ipx + 19 + x, lw);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint8_t* p = img + (int64_t)lk_refl(ipy - 1 + g7 + r, lh) * pitch;
#pragma unroll
            for (int x = 0; x < 5; ++x) R[r][x] = __ldg(p + cx[x]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint8_t* q = img + (int64_t)lk_refl(ipy - 1 + erow + r, lh) * pitch;
#pragma unroll
            for (int x = 0; x < 4; ++x) E[r][x] = __ldg(q + ce[x]);
        }
    }
    // Scharr: t0(x) = 3 (s[y-1][x] + s[y+1][x]) + 10 s[y][x],  t1(x) = s[y+1][x] - s[y-1][x];  dx = t0(x+1) - t0(x-1),
    // dy = 3 (t1(x-1) + t1(x+1)) + 10 t1(x)
    int Gx[8][3], Gy[8][3];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        int t0[5], t1[5];
#pragma unroll
        for (int x = 0; x < 5; ++x) { t0[x] = (R[d][x] + R[d + 2][x]) * 3 + R[d + 1][x] * 10; t1[x] = R[d + 2][x] - R[d][x]; }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            int gx = t0[x + 2] - t0[x], gy = (t1[x] + t1[x + 2]) * 3 + t1[x + 1] * 10;
            if (!kInside) {
                const bool in = (unsigned)(ipx + k2 + x) < (unsigned)lw && (unsigned)(ipy + g7 + d) < (unsigned)lh;
                gx = in ? gx : 0; gy = in ? gy : 0;
            }
            Gx[d][x] = gx; Gy[d][x] = gy;
        }
    }
    int a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int ival = (R[r + 1][c + 1] * w00 + R[r + 1][c + 2] * w01 + R[r + 2][c + 1] * w10 + R[r + 2][c + 2] * w11 + 256) >> 9;
            const int ixv = (Gx[r][c] * w00 + Gx[r][c + 1] * w01 + Gx[r + 1][c] * w10 + Gx[r + 1][c + 1] * w11 + 8192) >> 14;
            const int iyv = (Gy[r][c] * w00 + Gy[r][c + 1] * w01 + Gy[r + 1][c] * w10 + Gy[r + 1][c + 1] * w11 + 8192) >> 14;
            C[2 * r + c] = 256 - 512 * ival; GX[2 * r + c] = ixv; GY[2 * r + c] = iyv;
            a11 += ixv * ixv; a12 += ixv * iyv; a22 += iyv * iyv;
        }
    if (lane >= 30) a11 = a12 = a22 = 0;
    // the pixel of column 20
    int ex[2][2], ey[2][2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        int t0[4], t1[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { t0[x] = (E[d][x] + E[d + 2][x]) * 3 + E[d + 1][x] * 10; t1[x] = E[d + 2][x] - E[d][x]; }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            int gx = t0[x + 2] - t0[x], gy = (t1[x] + t1[x + 2]) * 3 + t1[x + 1] * 10;
            if (!kInside) {
                const bool in = (unsigned)(ipx + 20 + x) < (unsigned)lw && (unsigned)(ipy + erow + d) < (unsigned)lh;
                gx = in ? gx : 0; gy = in ? gy : 0;
            }
            ex[d][x] = gx; ey[d][x] = gy;
        }
    }
    const int ival = (E[1][1] * w00 + E[1][2] * w01 + E[2][1] * w10 + E[2][2] * w11 + 256) >> 9;
    const int ixv = (ex[0][0] * w00 + ex[0][1] * w01 + ex[1][0] * w10 + ex[1][1] * w11 + 8192) >> 14;
    const int iyv = (ey[0][0] * w00 + ey[0][1] * w01 + ey[1][0] * w10 + ey[1][1] * w11 + 8192) >> 14;
    Ce = 256 - 512 * ival; GXe = lane < kWin ? ixv : 0; GYe = lane < kWin ? iyv : 0;
    s11 = a11 + GXe * GXe; s12 = a12 + GXe * GYe; s22 = a22 + GYe * GYe;
}

// Mismatch vector of one iteration.  The 21x21 window is tiled over the warp: lane = 10 g + k (k < 10, g < 3) owns window columns
// 2k, 2k+1 of rows 7g .. 7g+6 (14 pixels, held in registers: C = 256 - 512 Iw folds the descale rounding and the subtraction of the
// template into the first multiply-add), and lanes 0..20 each own one pixel of the left-over column 20.  Lanes 30, 31 carry zero
// gradients.  Every lane loads its own 8x3 (+2x2) bytes of J; the sums per lane are < 15 * 2^25: exact in int32.
template <bool kInside>
__device__ __forceinline__ void lk_mismatch_tiles(const uint8_t* __restrict__ img, int pitch, int lw, int lh, int inx, int iny,
                                                  int k2, int g7, int erow, bool lane30, int w00, int w01, int w10, int w11,
                                                  const int (&C)[14], const int (&GX)[14], const int (&GY)[14], int Ce, int GXe, int GYe,
                                                  int& s1, int& s2) {
    int v[8][3];
    int e00, e01, e10, e11;
    if (kInside) {
        const uint8_t* p = img + (int64_t)(iny + g7) * pitch + (inx + k2);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            v[r][0] = __ldg(p); v[r][1] = __ldg(p + 1); v[r][2] = __ldg(p + 2);
            p += pitch;
        }
        const uint8_t* q = img + (int64_t)(iny + erow) * pitch + (inx + 20);
        e00 = __ldg(q); e01 = __ldg(q + 1); e10 = __ldg(q + pitch); e11 = __ldg(q + pitch + 1);
    } else {
        const int x0 = lk_refl(inx + k2, lw), x1 = lk_refl(inx + k2 + 1, lw), x2 = lk_refl(inx + k2 + 2, lw);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint8_t* p = img + (int64_t)lk_refl(iny + g7 + r, lh) * pitch;
            v[r][0] = __ldg(p + x0); v[r][1] = __ldg(p + x1); v[r][2] = __ldg(p + x2);
        }
        const int xe0 = lk_refl(inx + 20, lw), xe1 = lk_refl(inx + 21, lw);
        const uint8_t* q0 = img + (int64_t)lk_refl(iny + erow, lh) * pitch;
        const uint8_t* q1 = img + (int64_t)lk_refl(iny + erow + 1, lh) * pitch;
        e00 = __ldg(q0 + xe0); e01 = __ldg(q0 + xe1); e10 = __ldg(q1 + xe0); e11 = __ldg(q1 + xe1);
    }
    int a1 = 0, a2 = 0, b1 = 0, b2 = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const int da = (v[r][0] * w00 + v[r][1] * w01 + v[r + 1][0] * w10 + v[r + 1][1] * w11 + C[2 * r]) >> 9;
        const int db = (v[r][1] * w00 + v[r][2] * w01 + v[r + 1][1] * w10 + v[r + 1][2] * w11 + C[2 * r + 1]) >> 9;
        a1 += da * GX[2 * r]; a2 += da * GY[2 * r];
        b1 += db * GX[2 * r + 1]; b2 += db * GY[2 * r + 1];
    }
    const int de = (e00 * w00 + e01 * w01 + e10 * w10 + e11 * w11 + Ce) >> 9;
    if (lane30) { a1 = 0; a2 = 0; b1 = 0; b2 = 0; }
    s1 = a1 + b1 + de * GXe; s2 = a2 + b2 + de * GYe;
}

// points come either from keypoints (kps != nullptr: kp.x, kp.y) or from a plain float2 array
// (128 registers per thread: forcing more resident blocks spills the register window and measured 15-30 % slower)
__global__ void __launch_bounds__(kLkWarps * 32) lk_track_kernel(const __grid_constant__ LkLevels L, const sgs_keypoint* __restrict__ kps,
                                                                 const float2* __restrict__ pts, const int32_t* __restrict__ counts, int cap,
                                                                 const int32_t* __restrict__ prev_index, float2* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int p = blockIdx.x * kLkWarps + warp;
    const int n = counts ? min(counts[f], cap) : cap;
    if (p >= n) return;
    const int fj = prev_index ? prev_index[f] : f;       // frame of the batch that plays the role of the previous image
    const int64_t pi = (int64_t)f * cap + p;
    float ptx, pty;
    if (kps) { ptx = kps[pi].x; pty = kps[pi].y; } else { const float2 q = pts[pi]; ptx = q.x; pty = q.y; }
    const int tl = min(lane, 29);
    const int k2 = 2 * (tl % 10), g7 = 7 * (tl / 10), erow = min(lane, kWin - 1);      // this lane's tile of the window (see lk_mismatch_tiles)
    const float half_win = 10.f;                       // (winSize - 1) * 0.5
    const float flt_scale = 1.f / (1 << 20);
    float nx = 0.f, ny = 0.f;
    for (int level = L.max_level; level >= 0; --level) {
        const int lw = L.w[level], lh = L.h[level];
        const uint8_t* Iimg = L.I[level] + (int64_t)f * L.fstride[level];
        const int ipitch = L.pitch[level];
        const uint8_t* Jimg = L.J[level] + (int64_t)fj * (level == 0 ? L.fstrideJ0 : L.fstride[level]);
        const int jpitch = level == 0 ? L.pitchJ0 : L.pitch[level];
        const float sc = 1.f / (float)(1 << level);
        float px = __fmul_rn(ptx, sc), py = __fmul_rn(pty, sc);
        float qx, qy;
        if (level == L.max_level) { qx = px; qy = py; } else { qx = __fmul_rn(nx, 2.f); qy = __fmul_rn(ny, 2.f); }
        nx = qx; ny = qy;
        px = __fsub_rn(px, half_win); py = __fsub_rn(py, half_win);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -kWin || ipx >= lw || ipy < -kWin || ipy >= lh) continue;
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);
        int s11, s12, s22;
        int C[14], GX[14], GY[14], Ce, GXe, GYe;
        if (ipx >= 1 && ipx + 22 < lw && ipy >= 1 && ipy + 22 < lh)
            lk_setup_tiles<true>(Iimg, ipitch, lw, lh, ipx, ipy, k2, g7, erow, lane, w00, w01, w10, w11, C, GX, GY, Ce, GXe, GYe, s11, s12, s22);
        else
            lk_setup_tiles<false>(Iimg, ipitch, lw, lh, ipx, ipy, k2, g7, erow, lane, w00, w01, w10, w11, C, GX, GY, Ce, GXe, GYe, s11, s12, s22);
        const float A11 = __fmul_rn((float)warp_sum_grad(s11), flt_scale), A12 = __fmul_rn((float)warp_sum_grad(s12), flt_scale),
                    A22 = __fmul_rn((float)warp_sum_grad(s22), flt_scale);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                                        (float)(2 * kWin * kWin));
        if (min_eig < 1e-4f || D < 1.1920928955078125e-7f) continue;
        D = __fdiv_rn(1.f, D);
        qx = __fsub_rn(qx, half_win); qy = __fsub_rn(qy, half_win);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < kLkMaxCount; ++j) {
            const int inx = (int)floorf(qx), iny = (int)floorf(qy);
            if (inx < -kWin || inx >= lw || iny < -kWin || iny >= lh) break;
            lk_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            int s1, s2;
            if (inx >= 0 && inx + 21 < lw && iny >= 0 && iny + 21 < lh)
                lk_mismatch_tiles<true>(Jimg, jpitch, lw, lh, inx, iny, k2, g7, erow, lane >= 30, w00, w01, w10, w11, C, GX, GY, Ce, GXe, GYe, s1, s2);
            else
                lk_mismatch_tiles<false>(Jimg, jpitch, lw, lh, inx, iny, k2, g7, erow, lane >= 30, w00, w01, w10, w11, C, GX, GY, Ce, GXe, GYe, s1, s2);
            float B1, B2;
            warp_sum_pair(s1, s2, lane, flt_scale, B1, B2);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, B2), __fmul_rn(A22, B1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, B1), __fmul_rn(A11, B2)), D);
            qx = __fadd_rn(qx, dx); qy = __fadd_rn(qy, dy);
            nx = __fadd_rn(qx, half_win); ny = __fadd_rn(qy, half_win);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= 0.01 * 0.01) break;        // Point2f::ddot is double
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f)); ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
    }
    if (lane == 0) out[pi] = make_float2(nx, ny);
}

}  // namespace sgs

using namespace sgs;

struct sgs_lk {
    int device = 0, w = 0, h = 0, max_batch = 0, max_level = 0;
    int lw[kLkMaxLevel + 1], lh[kLkMaxLevel + 1], lp[kLkMaxLevel + 1];
    int64_t lfs[kLkMaxLevel + 1];
    uint8_t* d_pyrI = nullptr; uint8_t* d_pyrJ = nullptr;     // levels 1..max_level, each [max_batch][h][pitch]
    int64_t loff[kLkMaxLevel + 1];
    // staging for the single-pair host API
    uint8_t* d_img = nullptr; float* d_pts = nullptr; float* d_out = nullptr; int pts_cap = 0;
    cudaStream_t st = nullptr;
    // optional stage timing (pyramid build, tracker), same contract as sgs_extractor_set_profiling
    bool profiling = false, pending = false; cudaEvent_t ev[3] = {nullptr, nullptr, nullptr}; double ms_acc[2] = {0, 0}; int calls = 0;
};

namespace {
int lk_bad(const char* m) { set_error("%s", m); return SGS_ERR_INVALID; }

void build_pyramid(sgs_lk* k, const uint8_t* d_l0, int pitch0, int64_t fstride0, uint8_t* d_pyr, int nframes, cudaStream_t st) {
    const uint8_t* src = d_l0; int sp = pitch0; int64_t sfs = fstride0;
    for (int l = 1; l <= k->max_level; ++l) {
        uint8_t* dst = d_pyr + k->loff[l];
        dim3 grid((k->lw[l] + 31) / 32, (k->lh[l] + 7) / 8, nframes);
        lk_pyrdown_kernel<<<grid, 256, 0, st>>>(src, k->lw[l - 1], k->lh[l - 1], sp, sfs, dst, k->lw[l], k->lh[l], k->lp[l], k->lfs[l]);
        src = dst; sp = k->lp[l]; sfs = k->lfs[l];
    }
}

int run_lk(sgs_lk* k, const uint8_t* d_cur, const uint8_t* d_prev, const int32_t* d_prev_index, int nframes, size_t frame_stride, int pitch,
           const sgs_keypoint* d_kps, const float* d_pts, const int32_t* d_counts, int cap, float* d_out, cudaStream_t st) {
    const bool prof = k->profiling;
    if (prof && k->pending) {
        if (cudaEventSynchronize(k->ev[2]) == cudaSuccess) {
            for (int i = 0; i < 2; ++i) { float ms = 0; cudaEventElapsedTime(&ms, k->ev[i], k->ev[i + 1]); k->ms_acc[i] += ms; }
            k->calls++;
        }
        k->pending = false;
    }
    if (prof) cudaEventRecord(k->ev[0], st);
    build_pyramid(k, d_cur, pitch, (int64_t)frame_stride, k->d_pyrI, nframes, st);
    const bool same_batch = d_prev_index != nullptr;     // previous images are other frames of the same batch: one pyramid serves both roles
    if (!same_batch) build_pyramid(k, d_prev, pitch, (int64_t)frame_stride, k->d_pyrJ, nframes, st);
    LkLevels L;
    L.max_level = k->max_level;
    for (int l = 0; l <= k->max_level; ++l) {
        L.w[l] = k->lw[l]; L.h[l] = k->lh[l];
        if (l == 0) { L.I[0] = d_cur; L.J[0] = same_batch ? d_cur : d_prev; L.pitch[0] = pitch; L.fstride[0] = (int64_t)frame_stride; }
        else { L.I[l] = k->d_pyrI + k->loff[l]; L.J[l] = (same_batch ? k->d_pyrI : k->d_pyrJ) + k->loff[l]; L.pitch[l] = k->lp[l]; L.fstride[l] = k->lfs[l]; }
    }
    for (int l = k->max_level + 1; l <= kLkMaxLevel; ++l) { L.I[l] = L.J[l] = nullptr; L.w[l] = L.h[l] = L.pitch[l] = 0; L.fstride[l] = 0; }
    L.pitchJ0 = pitch; L.fstrideJ0 = (int64_t)frame_stride;
    if (prof) cudaEventRecord(k->ev[1], st);
    dim3 grid((cap + kLkWarps - 1) / kLkWarps, nframes);
    lk_track_kernel<<<grid, kLkWarps * 32, 0, st>>>(L, d_kps, reinterpret_cast<const float2*>(d_pts), d_counts, cap, d_prev_index, reinterpret_cast<float2*>(d_out));
    if (prof) { cudaEventRecord(k->ev[2], st); k->pending = true; }
    SGS_CUDA_TRY(cudaGetLastError());
    return SGS_OK;
}
}  // namespace

extern "C" {

SGS_API int sgs_lk_set_profiling(sgs_lk* k, int enable) {
    if (!k) return lk_bad("sgs_lk_set_profiling: NULL");
    SGS_CUDA_TRY(cudaSetDevice(k->device));
    if (enable && !k->ev[0]) for (auto& e : k->ev) SGS_CUDA_TRY(cudaEventCreate(&e));
    k->profiling = enable != 0; k->pending = false; k->ms_acc[0] = k->ms_acc[1] = 0; k->calls = 0;
    return SGS_OK;
}

SGS_API int sgs_lk_stage_times(sgs_lk* k, double* ms_total2, int* ncalls) {
    if (!k || !ms_total2 || !ncalls) return lk_bad("sgs_lk_stage_times: NULL");
    if (k->pending && cudaEventSynchronize(k->ev[2]) == cudaSuccess) {
        for (int i = 0; i < 2; ++i) { float ms = 0; cudaEventElapsedTime(&ms, k->ev[i], k->ev[i + 1]); k->ms_acc[i] += ms; }
        k->calls++; k->pending = false;
    }
    ms_total2[0] = k->ms_acc[0]; ms_total2[1] = k->ms_acc[1]; *ncalls = k->calls;
    return SGS_OK;
}

SGS_API void sgs_lk_destroy(sgs_lk* k) {
    if (!k) return;
    cudaSetDevice(k->device);
    for (auto& e : k->ev) if (e) cudaEventDestroy(e);
    cudaFree(k->d_pyrI); cudaFree(k->d_pyrJ); cudaFree(k->d_img); cudaFree(k->d_pts); cudaFree(k->d_out);
    if (k->st) cudaStreamDestroy(k->st);
    delete k;
}

SGS_API int sgs_lk_create(int width, int height, int max_batch, int device, sgs_lk** out) {
    if (!out || width < 24 || height < 24 || max_batch < 1) return lk_bad("sgs_lk_create: bad argument");
    *out = nullptr;
    SGS_CUDA_TRY(cudaSetDevice(device));
    sgs_lk* k = new sgs_lk();
    k->device = device; k->w = width; k->h = height; k->max_batch = max_batch;
    k->lw[0] = width; k->lh[0] = height; k->lp[0] = 0; k->lfs[0] = 0; k->loff[0] = 0;
    int64_t off = 0;
    k->max_level = 0;
    for (int l = 1; l <= kLkMaxLevel; ++l) {     // buildOpticalFlowPyramid stops when a level would not be larger than the window
        const int nw = (k->lw[l - 1] + 1) / 2, nh = (k->lh[l - 1] + 1) / 2;
        if (nw <= kWin || nh <= kWin) break;
        k->lw[l] = nw; k->lh[l] = nh; k->lp[l] = (nw + 15) & ~15; k->lfs[l] = (int64_t)k->lp[l] * nh; k->loff[l] = off;
        off += k->lfs[l] * max_batch;
        k->max_level = l;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&k->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&k->d_pyrI, (size_t)off + 256);
    if (e == cudaSuccess) e = cudaMalloc(&k->d_pyrJ, (size_t)off + 256);
    if (e != cudaSuccess) { set_error("sgs_lk_create: %s", cudaGetErrorString(e)); sgs_lk_destroy(k); return SGS_ERR_CUDA; }
    *out = k;
    return SGS_OK;
}

SGS_API int sgs_lk_track_batch_device(sgs_lk* k, const uint8_t* d_cur, const uint8_t* d_prev, const int32_t* d_prev_index, int nframes,
                                      size_t frame_stride, int pitch, const sgs_keypoint* d_kps, const int32_t* d_counts, int cap, float* d_prev_xy,
                                      void* stream) {
    if (!k || !d_cur || (!d_prev && !d_prev_index) || !d_kps || !d_counts || !d_prev_xy) return lk_bad("sgs_lk_track_batch_device: NULL argument");
    if (nframes < 1 || nframes > k->max_batch || cap < 1) return lk_bad("sgs_lk_track_batch_device: nframes/cap out of range");
    if (pitch < k->w || frame_stride < (size_t)pitch * k->h) return lk_bad("sgs_lk_track_batch_device: pitch/frame_stride too small");
    return run_lk(k, d_cur, d_prev, d_prev_index, nframes, frame_stride, pitch, d_kps, nullptr, d_counts, cap, d_prev_xy, stream ? (cudaStream_t)stream : k->st);
}

SGS_API int sgs_lk_track(sgs_lk* k, const uint8_t* cur, const uint8_t* prev, int pitch, const float* pts, int n, float* out) {
    if (!k || !cur || !prev || n < 0 || (n > 0 && (!pts || !out))) return lk_bad("sgs_lk_track: bad argument");
    if (n == 0) return SGS_OK;
    SGS_CUDA_TRY(cudaSetDevice(k->device));
    const int dp = (k->w + 15) & ~15;
    if (!k->d_img) SGS_CUDA_TRY(cudaMalloc(&k->d_img, (size_t)2 * dp * k->h));
    if (n > k->pts_cap) {
        cudaFree(k->d_pts); cudaFree(k->d_out); k->d_pts = k->d_out = nullptr;
        SGS_CUDA_TRY(cudaMalloc(&k->d_pts, 8 * (size_t)n)); SGS_CUDA_TRY(cudaMalloc(&k->d_out, 8 * (size_t)n));
        k->pts_cap = n;
    }
    SGS_CUDA_TRY(cudaMemcpy2DAsync(k->d_img, dp, cur, pitch, k->w, k->h, cudaMemcpyHostToDevice, k->st));
    SGS_CUDA_TRY(cudaMemcpy2DAsync(k->d_img + (size_t)dp * k->h, dp, prev, pitch, k->w, k->h, cudaMemcpyHostToDevice, k->st));
    SGS_CUDA_TRY(cudaMemcpyAsync(k->d_pts, pts, 8 * (size_t)n, cudaMemcpyHostToDevice, k->st));
    int rc = run_lk(k, k->d_img, k->d_img + (size_t)dp * k->h, nullptr, 1, (size_t)dp * k->h, dp, nullptr, k->d_pts, nullptr, n, k->d_out, k->st);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaMemcpyAsync(out, k->d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, k->st));
    SGS_CUDA_TRY(cudaStreamSynchronize(k->st));
    return SGS_OK;
}

SGS_API int sgs_lk_read_level(sgs_lk* k, int which, int level, uint8_t* out, int out_pitch) {   // parity accessor: frame 0 of the last call
    if (!k || !out || level < 1 || level > k->max_level) return lk_bad("sgs_lk_read_level: bad argument");
    SGS_CUDA_TRY(cudaSetDevice(k->device));
    SGS_CUDA_TRY(cudaStreamSynchronize(k->st));
    SGS_CUDA_TRY(cudaMemcpy2D(out, out_pitch, (which ? k->d_pyrJ : k->d_pyrI) + k->loff[level], k->lp[level], k->lw[level], k->lh[level], cudaMemcpyDeviceToHost));
    return SGS_OK;
}

}  // extern "C"

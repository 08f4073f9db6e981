// blur_kernel.cu -- GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of one pyramid level for all frames
// (src/ORBextractor.cc:1086-1087) in OpenCV's bit-exact 8.8 fixed point: taps {18,34,48,56,48,34,18}/256; the horizontal
// sums are exact 16-bit values (<= 65280), the vertical pass accumulates in 32 bits and rounds once: (v + 2^15) >> 16.
//
// Block = 128 x 64 output tile, 256 threads, three phases through shared memory:
//   1. stage the (64+6) x (128+8) input tile: aligned 32-bit loads in the image interior, per-byte reflect-101 at the borders;
//   2. horizontal pass on PACKED 16-bit lanes: a 32-bit register holds the pixel pair (x, x+2); one IMAD applies a tap to two
//      pixels (no lane overflow: every partial sum <= 65280);
//   3. vertical pass, 4 pixels x 8 rows per thread with the 7-row window in registers; the result byte is bits 16..23 of the
//      32-bit accumulator (accumulator < 2^24), picked with PRMT -- no shifts.
#include <cuda_runtime.h>

#include "extract_dev.cuh"
#include "extract_kernels.h"

namespace sgs {

constexpr int kBTW = 128, kBTH = 64;
constexpr int kBInRows = kBTH + 6;
constexpr int kBInWords = kBTW / 4 + 2;     // tile column 0 == image x0 - 4

__device__ __forceinline__ int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// One launch blurs every level.  Block order = the order of the former per-level launches: level-major, then frame, then the tiles of the frame
// (start[l] = first tile of level l; a level owns blocks [start[l] * nframes, start[l+1] * nframes)).  Frame-major order (all levels of a frame together)
// measured 5 % slower for the whole extraction -- the describe kernel then finds less of the blurred levels in L2 --, tile-major (all frames of one tile
// position together) 7 % slower for the blur itself.
struct BlurTiles { int32_t start[kMaxLevels + 1]; int32_t tx[kMaxLevels]; int32_t aligned4[kMaxLevels]; int32_t nlevels; };

__global__ void __launch_bounds__(256) blur_tile_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ BlurTiles T) {
    __shared__ uint32_t in[kBInRows * kBInWords];
    __shared__ uint2 hs[kBInRows * (kBTW / 4)];
    int level = 0;
    while (level + 1 < T.nlevels && blockIdx.x >= (unsigned)T.start[level + 1] * (unsigned)P.nframes) ++level;
    const DevLevel& L = P.lv[level];
    const int ntl = T.start[level + 1] - T.start[level];
    const unsigned rel = blockIdx.x - (unsigned)T.start[level] * (unsigned)P.nframes;
    const int frame = (int)(rel / (unsigned)ntl), tile = (int)(rel - (unsigned)frame * (unsigned)ntl);
    const int ty = tile / T.tx[level], tx = tile - ty * T.tx[level];
    const int w = L.w, h = L.h, spitch = L.pitch, dpitch = L.bpitch, aligned4 = T.aligned4[level];
    const uint8_t* S = L.img + (int64_t)frame * L.fstride;
    uint8_t* D = L.blur + (int64_t)frame * L.bfstride;
    const int x0 = tx * kBTW, y0 = ty * kBTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- 1. input tile ------------------------------------------------------------------------------------------------------
    // 1a: every word that lies completely inside the image row: one aligned 32-bit load (rows are reflected, columns are not)
    for (int it = threadIdx.x; it < kBInRows * kBInWords; it += 256) {
        const int r = it / kBInWords, wi = it - r * kBInWords;
        const int x = x0 - 4 + 4 * wi;
        if (aligned4 && x >= 0 && x + 3 < w) {
            const int gy = refl101(min(y0 - 3 + r, h + 2), h);
            in[it] = __ldg(reinterpret_cast<const uint32_t*>(S + (int64_t)gy * spitch + x));
        }
    }
    // 1b: the few words that straddle the left/right image border (or everything when the source is not 4-byte aligned):
    //     per-byte loads with reflect-101.  Words entirely beyond x = w+2 are never read by a stored output: skipped.
    {
        const int first_bad = aligned4 ? max(0, (w - 3 - (x0 - 4) + 3) >> 2) : 0;          // first word with x + 3 >= w
        const int last_needed = min(kBInWords - 1, (w + 2 - (x0 - 4)) >> 2);               // word holding column w + 2
        const int nright = max(0, last_needed - first_bad + 1);
        const int nleft = (aligned4 && x0 == 0) ? 1 : 0;                                   // word 0 = columns -4..-1
        const int per_row = nright + nleft;
        for (int it = threadIdx.x; it < kBInRows * per_row; it += 256) {
            const int r = it / per_row, k = it - r * per_row;
            const int wi = (k < nleft) ? 0 : first_bad + (k - nleft);
            if (nleft && k >= nleft && wi == 0) continue;                                   // word 0 already handled as the left word
            const int gy = refl101(min(y0 - 3 + r, h + 2), h);
            const uint8_t* row = S + (int64_t)gy * spitch;
            const int x = x0 - 4 + 4 * wi;
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (uint32_t)__ldg(row + refl101(min(x + b, w + 2), w)) << (8 * b);
            in[r * kBInWords + wi] = v;
        }
    }
    __syncthreads();
    // ---- 2. horizontal pass: lanes (x, x+2) ------------------------------------------------------------------------------------
    for (int r = warp; r < kBInRows; r += 8) {
        const uint32_t* p = in + r * kBInWords + lane;
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];                    // tile columns 4g .. 4g+11
        // X_i = bytes (b[i], b[i+1], b[i+2], b[i+3]) of the 12-byte window; R_i = (b[i], b[i+2]) as two 16-bit lanes
        const uint32_t x1 = __funnelshift_r(w0, w1, 8), x3 = __funnelshift_r(w0, w1, 24);
        const uint32_t x5 = __funnelshift_r(w1, w2, 8), x7 = __funnelshift_r(w1, w2, 24);
        const uint32_t r1 = x1 & 0x00FF00FFu, r2 = __byte_perm(x1, 0, 0x4341);
        const uint32_t r3 = x3 & 0x00FF00FFu, r4 = __byte_perm(x3, 0, 0x4341);
        const uint32_t r5 = x5 & 0x00FF00FFu, r6 = __byte_perm(x5, 0, 0x4341);
        const uint32_t r7 = x7 & 0x00FF00FFu, r8 = __byte_perm(x7, 0, 0x4341);
        uint2 o;
        o.x = 18u * (r1 + r7) + 34u * (r2 + r6) + 48u * (r3 + r5) + 56u * r4;   // (H0, H2): output x = x0 + 4g + {0, 2}
        o.y = 18u * (r2 + r8) + 34u * (r3 + r7) + 48u * (r4 + r6) + 56u * r5;   // (H1, H3)
        hs[r * (kBTW / 4) + lane] = o;
    }
    __syncthreads();
    // ---- 3. vertical pass: 4 pixels x 8 output rows per thread -------------------------------------------------------------------
    {
        const int g = lane, rb = warp * 8;                                   // output rows rb .. rb+7 need hs rows rb .. rb+13
        const int gx = x0 + 4 * g;
        if (gx >= w || y0 + rb >= h) return;
        uint32_t win[7][4];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const uint2 v = hs[(rb + i) * (kBTW / 4) + g];
            win[i][0] = v.x & 0xFFFFu; win[i][2] = v.x >> 16; win[i][1] = v.y & 0xFFFFu; win[i][3] = v.y >> 16;
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const uint2 v = hs[(rb + o + 6) * (kBTW / 4) + g];
            win[(o + 6) % 7][0] = v.x & 0xFFFFu; win[(o + 6) % 7][2] = v.x >> 16; win[(o + 6) % 7][1] = v.y & 0xFFFFu; win[(o + 6) % 7][3] = v.y >> 16;
            uint32_t acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a0 = win[(o + 0) % 7][j], a1 = win[(o + 1) % 7][j], a2 = win[(o + 2) % 7][j], a3 = win[(o + 3) % 7][j];
                const uint32_t a4 = win[(o + 4) % 7][j], a5 = win[(o + 5) % 7][j], a6 = win[(o + 6) % 7][j];
                acc[j] = 56u * a3 + 32768u + 48u * (a2 + a4) + 34u * (a1 + a5) + 18u * (a0 + a6);
            }
            const int gy = y0 + rb + o;
            if (gy < h) {
                const uint32_t lo = __byte_perm(acc[0], acc[1], 0x0062), hi = __byte_perm(acc[2], acc[3], 0x0062);   // byte 2 of each
                *reinterpret_cast<uint32_t*>(D + (int64_t)gy * dpitch + gx) = __byte_perm(lo, hi, 0x5410);          // dpitch % 16 == 0
            }
        }
    }
}

void launch_blur_all(const DevPlan& P, cudaStream_t st) {
    BlurTiles T{};
    T.nlevels = P.nlevels;
    int total = 0;
    for (int l = 0; l < P.nlevels; ++l) {
        const DevLevel& L = P.lv[l];
        T.start[l] = total; T.tx[l] = (L.w + kBTW - 1) / kBTW;
        T.aligned4[l] = (((uintptr_t)L.img & 3) == 0 && (L.pitch & 3) == 0 && (L.fstride & 3) == 0) ? 1 : 0;
        total += T.tx[l] * ((L.h + kBTH - 1) / kBTH);
    }
    T.start[P.nlevels] = total;
    blur_tile_kernel<<<(unsigned)total * (unsigned)P.nframes, 256, 0, st>>>(P, T);
}

}  // namespace sgs

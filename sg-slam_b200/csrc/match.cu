// match.cu -- projection matchers of ORBmatcher on the GPU, batched over independent frames (one block per frame).
//
//   match_lastframe_kernel : ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)   src/ORBmatcher.cc:1332-1472
//   match_localmap_kernel  : ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)    src/ORBmatcher.cc:45-129
//
// Both reference loops are order dependent (a keypoint already holding a map point with Observations()>0 is skipped,
// last writer wins).  Exact restatement in three phases per frame:
//   1. the Frame grid (Frame::AssignFeaturesToGrid, src/Frame.cc:257-272) as a sorted key array (cell << 16 | index):
//      ascending position in that array == the iteration order of Frame::GetFeaturesInArea (src/Frame.cc:354-407);
//   2. every map point scores its window in parallel (one warp per point) IGNORING claims -> optimistic best;
//   3. one warp walks the points in index order, accepts the optimistic result when its keypoint is not claimed,
//      otherwise rescans the window with the current claims (rare), reproducing the sequential semantics exactly.
// The sequential "dist < best" / "dist < second" updates keep the two smallest candidates under the order
// (distance, iteration position); that is what the warp reductions compute.
#include <cuda_runtime.h>

#include <vector>

#include "match_dev.cuh"
#include "sgs_common.h"

namespace sgs {

constexpr int kMatchThreads = 256;
constexpr int kGridCols = 64, kGridRows = 48;         // FRAME_GRID_COLS / ROWS, include/Frame.h:39-40
constexpr int kGridCells = kGridCols * kGridRows;
constexpr int kThHigh = 100, kHistoLen = 30;          // src/ORBmatcher.cc:37-39
constexpr uint32_t kNoKey = 0xFFFFFFFFu;

struct FrameSmem {
    uint32_t* keys;      // [n_sort]  cell << 16 | keypoint index, sorted; kNoKey padding
    float* kx; float* ky; float* ur;
    uint8_t* oct;
    uint8_t* claimed;    // 1 = holds a map point with Observations() > 0
    int32_t* cell_start; // [kGridCells + 1]
    int n, n_sort;
};

__device__ __forceinline__ size_t frame_smem_bytes(int cap_pow2) {
    return (size_t)cap_pow2 * (4 + 4 + 4 + 4 + 1 + 1) + (kGridCells + 1) * 4 + 64;
}

__device__ __forceinline__ void carve(uint8_t* base, int cap_pow2, FrameSmem& s) {
    s.keys = reinterpret_cast<uint32_t*>(base); base += (size_t)cap_pow2 * 4;
    s.kx = reinterpret_cast<float*>(base); base += (size_t)cap_pow2 * 4;
    s.ky = reinterpret_cast<float*>(base); base += (size_t)cap_pow2 * 4;
    s.ur = reinterpret_cast<float*>(base); base += (size_t)cap_pow2 * 4;
    s.cell_start = reinterpret_cast<int32_t*>(base); base += (size_t)(kGridCells + 1) * 4;
    s.oct = base; base += cap_pow2;
    s.claimed = base;
}

// Frame::PosInGrid (src/Frame.cc:409-419): round() is half away from zero
__device__ __forceinline__ int grid_round(float v) { return (int)roundf(v); }

__device__ void build_frame_grid(const MatchCam& cam, const sgs_keypoint* __restrict__ kps, const float* __restrict__ uright, int n, FrameSmem& s) {
    const float w_inv = __fdiv_rn((float)kGridCols, __fsub_rn(cam.max_x, cam.min_x));   // Frame.cc:183-184
    const float h_inv = __fdiv_rn((float)kGridRows, __fsub_rn(cam.max_y, cam.min_y));
    for (int i = threadIdx.x; i < s.n_sort; i += blockDim.x) {
        uint32_t key = kNoKey;
        if (i < n) {
            const sgs_keypoint kp = kps[i];
            s.kx[i] = kp.x; s.ky[i] = kp.y; s.oct[i] = (uint8_t)kp.octave; s.ur[i] = uright[i];
            const int px = grid_round(__fmul_rn(__fsub_rn(kp.x, cam.min_x), w_inv));
            const int py = grid_round(__fmul_rn(__fsub_rn(kp.y, cam.min_y), h_inv));
            if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) key = ((uint32_t)(px * kGridRows + py) << 16) | (uint32_t)i;
        }
        s.keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= s.n_sort; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (s.n_sort >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = ((i & k) == 0);
                const uint32_t x = s.keys[i], y = s.keys[p];
                if ((x > y) == up) { s.keys[i] = y; s.keys[p] = x; }
            }
            __syncthreads();
        }
    for (int c = threadIdx.x; c <= kGridCells; c += blockDim.x) {
        const uint32_t target = (uint32_t)c << 16;
        int lo = 0, hi = s.n_sort;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s.keys[mid] < target) lo = mid + 1; else hi = mid; }
        s.cell_start[c] = lo;
    }
    __syncthreads();
}

struct Top2 { uint32_t k1, k2; };  // key = dist << 16 | position in the sorted key array; 0xFFFFFFFF = none

__device__ __forceinline__ void top2_insert(Top2& t, uint32_t k) {
    if (k < t.k1) { t.k2 = t.k1; t.k1 = k; }
    else if (k < t.k2) t.k2 = k;
}

__device__ __forceinline__ Top2 warp_top2(Top2 t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t a1 = __shfl_xor_sync(0xffffffffu, t.k1, o), a2 = __shfl_xor_sync(0xffffffffu, t.k2, o);
        top2_insert(t, a1);
        top2_insert(t, a2);
    }
    return t;
}

__device__ __forceinline__ int popc256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Warp-cooperative Frame::GetFeaturesInArea + the candidate loop shared by both matchers.
//   use_claims : skip keypoints whose `claimed` flag is set (phase 3 rescans); phase 2 passes false
//   ur_gate    : the stereo/RGB-D right-coordinate gate (ur_pred, ur_tol)
// Returns the two smallest (dist, position) keys; *ncand accumulates |vIndices| (lane-0 value is the warp total).
__device__ Top2 scan_window(const MatchCam& cam, const FrameSmem& s, const uint4* __restrict__ cur_desc, float x, float y, float r,
                            int min_level, int max_level, const uint4& d0, const uint4& d1, bool use_claims, float ur_pred, float ur_tol,
                            int* ncand) {
    Top2 t; t.k1 = kNoKey; t.k2 = kNoKey;
    const float w_inv = __fdiv_rn((float)kGridCols, __fsub_rn(cam.max_x, cam.min_x));
    const float h_inv = __fdiv_rn((float)kGridRows, __fsub_rn(cam.max_y, cam.min_y));
    const int min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, cam.min_x), r), w_inv)));
    if (min_cx >= kGridCols) return t;
    const int max_cx = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, cam.min_x), r), w_inv)));
    if (max_cx < 0) return t;
    const int min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, cam.min_y), r), h_inv)));
    if (min_cy >= kGridRows) return t;
    const int max_cy = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, cam.min_y), r), h_inv)));
    if (max_cy < 0) return t;
    const bool check_levels = (min_level > 0) || (max_level >= 0);
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int ix = min_cx; ix <= max_cx; ++ix) {
        // cells (ix, min_cy..max_cy) are contiguous in the sorted array
        const int beg = s.cell_start[ix * kGridRows + min_cy], end = s.cell_start[ix * kGridRows + max_cy + 1];
        for (int j = beg + lane; j < end; j += 32) {
            const int idx = (int)(s.keys[j] & 0xFFFFu);
            const int o = s.oct[idx];
            if (check_levels) {
                if (o < min_level) continue;
                if (max_level >= 0 && o > max_level) continue;
            }
            const float dx = __fsub_rn(s.kx[idx], x), dy = __fsub_rn(s.ky[idx], y);
            if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
            ++cnt;                                            // member of vIndices
            if (use_claims && s.claimed[idx]) continue;       // :87-89 / :1407-1409
            const float ur = s.ur[idx];
            if (ur > 0.f) {                                   // :91-96 / :1411-1417
                const float er = fabsf(__fsub_rn(ur_pred, ur));
                if (er > ur_tol) continue;
            }
            const int dist = popc256(d0, d1, __ldg(&cur_desc[2 * idx]), __ldg(&cur_desc[2 * idx + 1]));
            top2_insert(t, ((uint32_t)dist << 16) | (uint32_t)j);
        }
    }
    if (ncand) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        *ncand += cnt;
    }
    return warp_top2(t);
}

__device__ __forceinline__ void three_maxima(const int* hist, int& i1, int& i2, int& i3) {  // ORBmatcher.cc:1603-1644
    int m1 = 0, m2 = 0, m3 = 0;
    i1 = i2 = i3 = -1;
    for (int i = 0; i < kHistoLen; ++i) {
        const int sz = hist[i];
        if (sz > m1) { m3 = m2; m2 = m1; m1 = sz; i3 = i2; i2 = i1; i1 = i; }
        else if (sz > m2) { m3 = m2; m2 = sz; i3 = i2; i2 = i; }
        else if (sz > m3) { m3 = sz; i3 = i; }
    }
    if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
    else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMatchThreads) match_lastframe_kernel(const __grid_constant__ LastFrameArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int hist[kHistoLen];
    __shared__ int s_nmatch, s_nevent;
    __shared__ float s_pose[16];     // Rcw (9), tcw (3), forward/backward flags
    const int f = blockIdx.x;
    const int n = min(A.cur_n[f], A.cur_cap);
    const int nlast = min(A.last_n[f], A.last_cap);
    FrameSmem s;
    carve(smem, A.cur_cap_pow2, s);
    s.n = n; s.n_sort = 1;
    while (s.n_sort < n) s.n_sort <<= 1;
    const sgs_keypoint* kps = A.cur_kps + (int64_t)f * A.cur_cap;
    const uint4* cur_desc = reinterpret_cast<const uint4*>(A.cur_desc + (int64_t)f * A.cur_cap * 32);
    int32_t* cur_mp = A.cur_mp + (int64_t)f * A.cur_cap;
    const int64_t lo = (int64_t)f * A.last_cap;
    if (threadIdx.x < kHistoLen) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        s_nmatch = 0; s_nevent = 0;
        // Rcw, tcw, twc = -Rcw^T tcw, tlc = Rlw twc + tlw : cv::Mat float gemm = double accumulation, one rounding (:1342-1351)
        const float* Tc = A.tcw_cur + 16 * f; const float* Tl = A.tcw_last + 16 * f;
        float twc[3];
        for (int r = 0; r < 3; ++r) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc = __dadd_rn(acc, __dmul_rn((double)(-Tc[4 * k + r]), (double)Tc[4 * k + 3]));
            twc[r] = (float)acc;
        }
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc = __dadd_rn(acc, __dmul_rn((double)Tl[8 + k], (double)twc[k]));
        const float tlc2 = (float)__dadd_rn(acc, (double)Tl[11]);
        const float mb = __fdiv_rn(A.cam.bf, A.cam.fx);       // Frame.cc:196
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) s_pose[3 * r + c] = Tc[4 * r + c]; s_pose[9 + r] = Tc[4 * r + 3]; }
        s_pose[12] = (tlc2 > mb && !A.mono) ? 1.f : 0.f;
        s_pose[13] = (-tlc2 > mb && !A.mono) ? 1.f : 0.f;
    }
    build_frame_grid(A.cam, kps, A.cur_uright + (int64_t)f * A.cur_cap, n, s);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool has = cur_mp[i] >= 0;
        s.claimed[i] = has ? (A.cur_mp_obs_in ? A.cur_mp_obs_in[(int64_t)f * A.cur_cap + i] : 1) : 0;
    }
    __syncthreads();
    const bool fwd = s_pose[12] != 0.f, bwd = s_pose[13] != 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    PointPre* pre = A.pre + lo;
    int ncand = 0;
    // phase 2: optimistic scoring, one warp per last-frame point
    for (int i = warp; i < nlast; i += nwarps) {
        PointPre pp; pp.valid = 0; pp.best_key = kNoKey; pp.u = pp.v = pp.invz = pp.radius = 0.f; pp.min_level = pp.max_level = 0;
        if (A.last_flags[lo + i] & 1) {
            const float* X = A.last_xyz + 3 * (lo + i);
            float xc[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double acc = 0.0;
                for (int k = 0; k < 3; ++k) acc = __dadd_rn(acc, __dmul_rn((double)s_pose[3 * r + k], (double)X[k]));
                xc[r] = (float)__dadd_rn(acc, (double)s_pose[9 + r]);
            }
            const float invz = (float)__ddiv_rn(1.0, (double)xc[2]);                       // :1369
            if (!(invz < 0.f)) {
                const float u = __fadd_rn(__fmul_rn(__fmul_rn(A.cam.fx, xc[0]), invz), A.cam.cx);
                const float v = __fadd_rn(__fmul_rn(__fmul_rn(A.cam.fy, xc[1]), invz), A.cam.cy);
                if (!(u < A.cam.min_x || u > A.cam.max_x) && !(v < A.cam.min_y || v > A.cam.max_y)) {
                    const int oct = A.last_octave[lo + i];
                    const float radius = __fmul_rn(A.th, A.cam.scale[oct]);
                    int mn, mx;
                    if (fwd) { mn = oct; mx = -1; } else if (bwd) { mn = 0; mx = oct; } else { mn = oct - 1; mx = oct + 1; }
                    pp.valid = 1; pp.u = u; pp.v = v; pp.invz = invz; pp.radius = radius; pp.min_level = mn; pp.max_level = mx;
                    const uint4* dm = reinterpret_cast<const uint4*>(A.last_desc + 32 * (lo + i));
                    const uint4 d0 = __ldg(dm), d1 = __ldg(dm + 1);
                    const float ur_pred = __fsub_rn(u, __fmul_rn(A.cam.bf, invz));
                    const Top2 t = scan_window(A.cam, s, cur_desc, u, v, radius, mn, mx, d0, d1, false, ur_pred, radius, &ncand);
                    pp.best_key = t.k1;
                }
            }
        }
        if (lane == 0) pre[i] = pp;
    }
    __syncthreads();
    // phase 3: sequential resolution by warp 0
    if (warp == 0) {
        int nmatch = 0, nevent = 0;
        for (int i = 0; i < nlast; ++i) {
            const PointPre pp = pre[i];
            if (!pp.valid || pp.best_key == kNoKey) {
                // no candidate survived the optimistic scan; with claims there are even fewer: nothing to do.
                continue;
            }
            uint32_t key = pp.best_key;
            int idx = (int)(s.keys[key & 0xFFFFu] & 0xFFFFu);
            if (s.claimed[idx]) {
                const uint4* dm = reinterpret_cast<const uint4*>(A.last_desc + 32 * (lo + i));
                const uint4 d0 = __ldg(dm), d1 = __ldg(dm + 1);
                const float ur_pred = __fsub_rn(pp.u, __fmul_rn(A.cam.bf, pp.invz));
                const Top2 t = scan_window(A.cam, s, cur_desc, pp.u, pp.v, pp.radius, pp.min_level, pp.max_level, d0, d1, true, ur_pred, pp.radius, nullptr);
                key = t.k1;
                if (key == kNoKey) continue;
                idx = (int)(s.keys[key & 0xFFFFu] & 0xFFFFu);
            }
            const int dist = (int)(key >> 16);
            if (dist <= kThHigh) {
                if (lane == 0) {
                    cur_mp[idx] = i;                                                  // :1432 last writer wins
                    s.claimed[idx] = (A.last_flags[lo + i] >> 1) & 1;
                    if (A.check_ori) {
                        float rot = __fsub_rn(A.last_angle[lo + i], kps[idx].angle);
                        if (rot < 0.f) rot = __fadd_rn(rot, 360.f);
                        int bin = (int)roundf(__fmul_rn(rot, (float)kHistoLen / 360.0f));
                        if (bin == kHistoLen) bin = 0;
                        hist[bin]++;
                        A.events[lo + nevent] = (bin << 16) | idx;
                    }
                }
                ++nmatch; ++nevent;
                __syncwarp();
            }
        }
        if (lane == 0) { s_nmatch = nmatch; s_nevent = nevent; }
    }
    __syncthreads();
    if (A.check_ori && threadIdx.x == 0) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        int nm = s_nmatch;
        for (int e = 0; e < s_nevent; ++e) {
            const int ev = A.events[lo + e];
            const int bin = ev >> 16, idx = ev & 0xFFFF;
            if (bin != i1 && bin != i2 && bin != i3) { cur_mp[idx] = -1; --nm; }       // :1451-1470
        }
        s_nmatch = nm;
    }
    __syncthreads();
    // totals
    if (lane == 0 && ncand) atomicAdd((unsigned long long*)&A.ncand[f], (unsigned long long)ncand);
    if (threadIdx.x == 0) A.nmatches[f] = s_nmatch;
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMatchThreads) match_localmap_kernel(const __grid_constant__ LocalMapArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_nmatch;
    const int f = blockIdx.x;
    const int n = min(A.cur_n[f], A.cur_cap);
    const int nmp = min(A.mp_n[f], A.mp_cap);
    FrameSmem s;
    carve(smem, A.cur_cap_pow2, s);
    s.n = n; s.n_sort = 1;
    while (s.n_sort < n) s.n_sort <<= 1;
    const sgs_keypoint* kps = A.cur_kps + (int64_t)f * A.cur_cap;
    const uint4* cur_desc = reinterpret_cast<const uint4*>(A.cur_desc + (int64_t)f * A.cur_cap * 32);
    int32_t* f_mp = A.f_mp + (int64_t)f * A.cur_cap;
    uint8_t* f_obs = A.f_mp_obs + (int64_t)f * A.cur_cap;
    const int64_t lo = (int64_t)f * A.mp_cap;
    build_frame_grid(A.cam, kps, A.cur_uright + (int64_t)f * A.cur_cap, n, s);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s.claimed[i] = (f_mp[i] >= 0 && f_obs[i]) ? 1 : 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool b_factor = A.th != 1.0f;
    LocalPre* pre = A.pre + lo;
    int ncand = 0;
    for (int i = warp; i < nmp; i += nwarps) {
        LocalPre pp; pp.k1 = kNoKey; pp.k2 = kNoKey;
        if (A.mp_inview[lo + i]) {
            const int lvl = A.level[lo + i];
            float r = ((double)A.view_cos[lo + i] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos :131-137 (compared in double)
            if (b_factor) r = __fmul_rn(r, A.th);
            const float rr = __fmul_rn(r, A.cam.scale[lvl]);
            const uint4* dm = reinterpret_cast<const uint4*>(A.mp_desc + 32 * (lo + i));
            const Top2 t = scan_window(A.cam, s, cur_desc, A.proj_x[lo + i], A.proj_y[lo + i], rr, lvl - 1, lvl, __ldg(dm), __ldg(dm + 1), false,
                                       A.proj_xr[lo + i], rr, &ncand);
            pp.k1 = t.k1; pp.k2 = t.k2;
        }
        if (lane == 0) pre[i] = pp;
    }
    __syncthreads();
    if (warp == 0) {
        int nmatch = 0;
        for (int i = 0; i < nmp; ++i) {
            LocalPre pp = pre[i];
            if (pp.k1 == kNoKey) continue;
            // the optimistic top-2 is exact unless one of the two keypoints has been claimed in the meantime
            const int i1 = (int)(s.keys[pp.k1 & 0xFFFFu] & 0xFFFFu);
            const int i2 = pp.k2 != kNoKey ? (int)(s.keys[pp.k2 & 0xFFFFu] & 0xFFFFu) : -1;
            if (s.claimed[i1] || (i2 >= 0 && s.claimed[i2])) {
                const int lvl = A.level[lo + i];
                float r = ((double)A.view_cos[lo + i] > 0.998) ? 2.5f : 4.0f;
                if (b_factor) r = __fmul_rn(r, A.th);
                const float rr = __fmul_rn(r, A.cam.scale[lvl]);
                const uint4* dm = reinterpret_cast<const uint4*>(A.mp_desc + 32 * (lo + i));
                const Top2 t = scan_window(A.cam, s, cur_desc, A.proj_x[lo + i], A.proj_y[lo + i], rr, lvl - 1, lvl, __ldg(dm), __ldg(dm + 1), true,
                                           A.proj_xr[lo + i], rr, nullptr);
                pp.k1 = t.k1; pp.k2 = t.k2;
                if (pp.k1 == kNoKey) continue;
            }
            const int best_dist = (int)(pp.k1 >> 16);
            if (best_dist <= kThHigh) {
                const int bi = (int)(s.keys[pp.k1 & 0xFFFFu] & 0xFFFFu);
                const int best_level = s.oct[bi];
                int best_level2 = -1, best_dist2 = 256;
                if (pp.k2 != kNoKey) { best_dist2 = (int)(pp.k2 >> 16); best_level2 = s.oct[s.keys[pp.k2 & 0xFFFFu] & 0xFFFFu]; }
                if (best_level == best_level2 && (float)best_dist > __fmul_rn(A.nnratio, (float)best_dist2)) continue;   // :118-119
                if (lane == 0) {
                    f_mp[bi] = A.id_base + i;
                    const uint8_t ob = A.mp_obs[lo + i];
                    f_obs[bi] = ob;
                    s.claimed[bi] = ob ? 1 : 0;
                }
                ++nmatch;
                __syncwarp();
            }
        }
        if (lane == 0) s_nmatch = nmatch;
    }
    __syncthreads();
    if (lane == 0 && ncand) atomicAdd((unsigned long long*)&A.ncand[f], (unsigned long long)ncand);
    if (threadIdx.x == 0) A.nmatches[f] = s_nmatch;
}

size_t match_smem_bytes(int cap_pow2) { return (size_t)cap_pow2 * (4 + 4 + 4 + 4 + 1 + 1) + (kGridCells + 1) * 4 + 64; }

int launch_match_lastframe(const LastFrameArgs& A, int nframes, cudaStream_t st) {
    const size_t smem = match_smem_bytes(A.cur_cap_pow2);
    if (smem > 200 * 1024) { set_error("match: cur_cap %d too large for shared memory", A.cur_cap); return SGS_ERR_UNSUPPORTED; }
    SGS_CUDA_TRY(cudaFuncSetAttribute(match_lastframe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_lastframe_kernel<<<nframes, kMatchThreads, smem, st>>>(A);
    SGS_CUDA_TRY(cudaGetLastError());
    return SGS_OK;
}

int launch_match_localmap(const LocalMapArgs& A, int nframes, cudaStream_t st) {
    const size_t smem = match_smem_bytes(A.cur_cap_pow2);
    if (smem > 200 * 1024) { set_error("match: cur_cap %d too large for shared memory", A.cur_cap); return SGS_ERR_UNSUPPORTED; }
    SGS_CUDA_TRY(cudaFuncSetAttribute(match_localmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_localmap_kernel<<<nframes, kMatchThreads, smem, st>>>(A);
    SGS_CUDA_TRY(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs

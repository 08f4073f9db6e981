// =====================================================================================
// sgs_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17 restatement (no OpenCV, no CUDA) of the SG-SLAM per-frame tracking hot
// path, written to be *literal* rather than fast.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (libsgs_cuda.so) never links, calls or falls back to it.
//
// PARITY STATUS: the reference ships no tests/golden vectors for this path and cannot be
// compiled here (needs OpenCV/Eigen/ncnn/ROS).  The oracle is therefore pinned by
//   (1) tests/golden/*.npz produced by tests/golden/make_golden.py, an independent Python
//       restatement that calls the REAL OpenCV primitives through cv2 (resize, FAST,
//       GaussianBlur, fastAtan2) -- see tests/test_oracle_golden.py, and
//   (2) live cv2 cross-checks of every OpenCV primitive restated below.
// "parity unpinned by the reference itself" -- see DESIGN.md section "Oracle".
//
// All `file:line` citations are relative to /root/reference/src/sg-slam/ .
// Build flags mirror the reference (CMakeLists.txt:11-12): -O3, no -march=native, and
// -ffp-contract=off so no FMA contraction can occur.
// =====================================================================================
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <list>
#include <utility>
#include <vector>

#define SGO_API extern "C" __attribute__((visibility("default")))

namespace {

// ---- cvRound / cvFloor / cvCeil (OpenCV fast_math.hpp): round-half-to-even via lrint ----
inline int cvRoundD(double v) { return (int)std::lrint(v); }
inline int cvRoundF(float v) { return (int)std::lrint((double)v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeilF(float v) { int i = (int)v; return i + (i < v); }

constexpr int PATCH_SIZE = 31;        // src/ORBextractor.cc:73
constexpr int HALF_PATCH_SIZE = 15;   // :74
constexpr int EDGE_THRESHOLD = 19;    // :75

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

struct KeyPoint {  // cv::KeyPoint layout (28 bytes)
    float x, y, size, angle, response;
    int32_t octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct OrbParams {
    int32_t nfeatures;
    float scaleFactor;
    int32_t nlevels;
    int32_t iniThFAST;
    int32_t minThFAST;
};

struct OrbTables {
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> nPerLevel;
    int umax[HALF_PATCH_SIZE + 1];
};

// ORBextractor::ORBextractor, src/ORBextractor.cc:411-471
OrbTables make_tables(const OrbParams& p) {
    OrbTables t;
    const int nlevels = p.nlevels;
    const double scaleFactor = (double)p.scaleFactor;  // member is `double scaleFactor` (ORBextractor.h:97)
    t.scale.resize(nlevels); t.sigma2.resize(nlevels);
    t.scale[0] = 1.0f; t.sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        t.scale[i] = (float)((double)t.scale[i - 1] * scaleFactor);  // :420
        t.sigma2[i] = t.scale[i] * t.scale[i];                        // :421
    }
    t.invScale.resize(nlevels); t.invSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        t.invScale[i] = 1.0f / t.scale[i];    // :428
        t.invSigma2[i] = 1.0f / t.sigma2[i];  // :429
    }
    t.nPerLevel.resize(nlevels);
    float factor = (float)(1.0 / scaleFactor);  // :435  (1.0f / double -> double -> float)
    float nDesired = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));  // :436
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        t.nPerLevel[level] = cvRoundF(nDesired);  // :441
        sum += t.nPerLevel[level];
        nDesired *= factor;
    }
    t.nPerLevel[nlevels - 1] = std::max(p.nfeatures - sum, 0);  // :445
    // umax, :453-470
    int v, v0;
    int vmax = cvFloorF(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = cvCeilF(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= HALF_PATCH_SIZE; ++v) t.umax[v] = 0;
    for (v = 0; v <= vmax; ++v) t.umax[v] = cvRoundD(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
    return t;
}

// ---------------------------------------------------------------------------------------
// cv::resize(..., INTER_LINEAR) for CV_8UC1 (OpenCV imgproc/resize.cpp, pinned 3.4.15 by
// README.md:91; cross-checked bit-exactly against cv2 4.13 in tests).  Called at
// src/ORBextractor.cc:1121.
// ---------------------------------------------------------------------------------------
struct LinTab { std::vector<int> s; std::vector<short> a0, a1; };

LinTab linear_table(int ssize, int dsize) {
    LinTab t; t.s.resize(dsize); t.a0.resize(dsize); t.a1.resize(dsize);
    const double inv_scale = (double)dsize / (double)ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cvFloorF(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        t.s[d] = s;
        t.a0[d] = (short)cvRoundF((1.f - f) * 2048.f);  // saturate_cast<short>(float) == cvRound
        t.a1[d] = (short)cvRoundF(f * 2048.f);
    }
    return t;
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch) {
    LinTab tx = linear_table(sw, dw), ty = linear_table(sh, dh);
    std::vector<int> row0(dw), row1(dw);
    for (int y = 0; y < dh; y++) {
        const int sy0 = ty.s[y];
        const int sy1 = std::min(sy0 + 1, sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * spitch;
        const uint8_t* S1 = src + (size_t)sy1 * spitch;
        for (int x = 0; x < dw; x++) {
            const int sx0 = tx.s[x];
            const int sx1 = std::min(sx0 + 1, sw - 1);
            row0[x] = S0[sx0] * tx.a0[x] + S0[sx1] * tx.a1[x];
            row1[x] = S1[sx0] * tx.a0[x] + S1[sx1] * tx.a1[x];
        }
        const int b0 = ty.a0[y], b1 = ty.a1[y];
        uint8_t* D = dst + (size_t)y * dpitch;
        for (int x = 0; x < dw; x++) {
            int v = (((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2;
            D[x] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;  // tightly packed, pitch == w
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
    uint8_t* row(int y) { return px.data() + (size_t)y * w; }
};

// ORBextractor::ComputePyramid, src/ORBextractor.cc:1108-1133.  The 19-px reflect-101 border
// the reference adds (:1123,:1128) is never read on the mono/RGB-D path:
// it is not materialised here.
void level_size(int w, int h, const OrbTables& t, int level, int* lw, int* lh) {
    const float scale = t.invScale[level];
    *lw = cvRoundF((float)w * scale);  // :1113
    *lh = cvRoundF((float)h * scale);
}

std::vector<Image> compute_pyramid(const uint8_t* img, int w, int h, int pitch, const OrbTables& t) {
    const int nlevels = (int)t.scale.size();
    std::vector<Image> pyr(nlevels);
    for (int level = 0; level < nlevels; level++) {
        Image& L = pyr[level];
        level_size(w, h, t, level, &L.w, &L.h);
        L.px.resize((size_t)L.w * L.h);
        if (level == 0) {
            for (int y = 0; y < h; y++) std::memcpy(L.row(y), img + (size_t)y * pitch, w);
        } else {
            const Image& P = pyr[level - 1];  // resized from level-1, not level 0 (:1121)
            resize_linear_u8(P.px.data(), P.w, P.h, P.w, L.px.data(), L.w, L.h, L.w);
        }
    }
    return pyr;
}

// ---------------------------------------------------------------------------------------
// cv::FAST(view, kps, threshold, nonmaxSuppression=true)  (TYPE_9_16), features2d/fast.cpp
// + fast_score.cpp.  Called at src/ORBextractor.cc:810,815 on a cell view.
// ---------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// cornerScore<16>: max over the 16 cyclic 9-arcs of max(min d, -max d), minus 1.
// For a pixel that is a corner at threshold t this is >= t and independent of t.
inline int fast_corner_score(const uint8_t* p, int pitch) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * pitch + kRingDx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

// is p a FAST-9-16 corner at threshold t (>= 9 contiguous ring pixels all > v+t or all < v-t)
inline bool fast_is_corner(const uint8_t* p, int pitch, int t) {
    const int v = p[0];
    uint32_t bright = 0, dark = 0;
    for (int k = 0; k < 16; k++) {
        const int r = p[kRingDy[k] * pitch + kRingDx[k]];
        if (r > v + t) bright |= 1u << k;
        if (r < v - t) dark |= 1u << k;
    }
    auto has9 = [](uint32_t m) {
        uint32_t mm = m | (m << 16);  // unroll the ring
        uint32_t r = mm;
        for (int j = 1; j < 9; j++) r &= (mm >> j);
        return (r & 0xFFFFu) != 0;
    };
    return has9(bright) || has9(dark);
}

struct CellKp { int x, y, score; };

// view = w x h pixels starting at `base` (pitch bytes per row).  Emits keypoints in row-major
// order with coordinates relative to the view, response = score.
void fast_detect_view(const uint8_t* base, int pitch, int w, int h, int threshold, bool nms, std::vector<CellKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<int> sc((size_t)w * h, 0);  // 0 for non-corners and for the 3-px frame
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = base + (size_t)y * pitch + x;
            if (fast_is_corner(p, pitch, threshold)) sc[(size_t)y * w + x] = fast_corner_score(p, pitch);
        }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const int s = sc[(size_t)y * w + x];
            if (s == 0) continue;  // non-corner (a corner has score >= threshold >= 1)
            if (nms) {
                bool keep = true;
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (!(s > sc[(size_t)(y + dy) * w + (x + dx)])) { keep = false; break; }
                    }
                if (!keep) continue;
            }
            out.push_back({x, y, s});
        }
}

struct Cand { float x, y, response; };  // coordinates relative to (minBorderX, minBorderY)

// FAST part of ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:766-830
void fast_level_candidates(const Image& L, int iniTh, int minTh, std::vector<Cand>& cands, int* nfallback) {
    cands.clear();
    const float W = 30;
    const int minBorderX = EDGE_THRESHOLD - 3;
    const int minBorderY = minBorderX;
    const int maxBorderX = L.w - EDGE_THRESHOLD + 3;
    const int maxBorderY = L.h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX);
    const float height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W);
    const int nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols);
    const int hCell = (int)std::ceil(height / nRows);
    std::vector<CellKp> cell;
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = (float)maxBorderX;
            const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;  // rowRange/colRange(int)
            const uint8_t* base = L.row(y0) + x0;
            fast_detect_view(base, L.w, x1 - x0, y1 - y0, iniTh, true, cell);
            if (cell.empty()) {
                fast_detect_view(base, L.w, x1 - x0, y1 - y0, minTh, true, cell);
                if (nfallback) (*nfallback)++;
            }
            for (const CellKp& k : cell)
                cands.push_back({(float)k.x + j * wCell, (float)k.y + i * hCell, (float)k.score});  // :821-826
        }
    }
}

// ---------------------------------------------------------------------------------------
// ExtractorNode::DivideNode + ORBextractor::DistributeOctTree, src/ORBextractor.cc:482-764
// Tie-break of the stable_sort on pair<int, ExtractorNode*> (:685) is DEFINED here as node
// creation sequence (quirk Q1).
// ---------------------------------------------------------------------------------------
struct Node {
    std::vector<int> keys;  // indices into the candidate array, in candidate order
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool noMore = false;
    long seq = 0;
};

void divide_node(const Node& n, const std::vector<Cand>& c, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)std::ceil((float)(n.URx - n.ULx) / 2);  // :484
    const int halfY = (int)std::ceil((float)(n.BRy - n.ULy) / 2);  // :485
    n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (int k : n.keys) {  // :512-527 (float point vs int corner)
        const Cand& kp = c[k];
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(k); else n3.keys.push_back(k);
        } else if (kp.y < n1.BRy) n2.keys.push_back(k);
        else n4.keys.push_back(k);
    }
    n1.noMore = n1.keys.size() == 1; n2.noMore = n2.keys.size() == 1;
    n3.noMore = n3.keys.size() == 1; n4.noMore = n4.keys.size() == 1;
}

std::vector<int> distribute_octree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N) {
    std::vector<int> result;
    if (c.empty()) return result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));  // :544
    if (nIni <= 0) return result;  // quirk Q2: the reference divides by zero here (portrait images)
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
        ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        nodes.push_back(ni);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < c.size(); i++) {
        size_t r = (size_t)(c[i].x / hX);  // :571
        if (r >= (size_t)nIni) r = nIni - 1;  // (never taken for valid input; guards the oracle only)
        ini[r]->keys.push_back((int)i);
    }
    for (auto lit = nodes.begin(); lit != nodes.end();) {  // :575-586
        if (lit->keys.size() == 1) { lit->noMore = true; ++lit; }
        else if (lit->keys.empty()) lit = nodes.erase(lit);
        else ++lit;
    }
    bool finish = false;
    typedef std::pair<int, Node*> SP;
    auto sp_less = [](const SP& a, const SP& b) {
        if (a.first != b.first) return a.first < b.first;
        return a.second->seq < b.second->seq;  // Q1: creation order replaces the heap address
    };
    std::vector<SP> sizeAndNode;
    auto push_children = [&](Node* kids[4], int* nToExpand) {
        for (int q = 0; q < 4; q++) {
            Node& ch = *kids[q];
            if (ch.keys.empty()) continue;
            ch.seq = seq++;
            nodes.push_front(ch);
            if (ch.keys.size() > 1) {
                if (nToExpand) (*nToExpand)++;
                sizeAndNode.push_back(SP((int)ch.keys.size(), &nodes.front()));
                nodes.front().lit = nodes.begin();
            }
        }
    };
    while (!finish) {
        const int prevSize = (int)nodes.size();
        auto lit = nodes.begin();
        int nToExpand = 0;
        sizeAndNode.clear();
        while (lit != nodes.end()) {  // coarse pass :601-666
            if (lit->noMore) { ++lit; continue; }
            Node n1, n2, n3, n4;
            divide_node(*lit, c, n1, n2, n3, n4);
            Node* kids[4] = {&n1, &n2, &n3, &n4};
            push_children(kids, &nToExpand);
            lit = nodes.erase(lit);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if (((int)nodes.size() + nToExpand * 3) > N) {  // :674
            while (!finish) {
                const int prevSize2 = (int)nodes.size();
                std::vector<SP> prev = sizeAndNode;
                sizeAndNode.clear();
                std::stable_sort(prev.begin(), prev.end(), sp_less);  // :685
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    divide_node(*prev[j].second, c, n1, n2, n3, n4);
                    Node* kids[4] = {&n1, &n2, &n3, &n4};
                    push_children(kids, nullptr);
                    nodes.erase(prev[j].second->lit);  // :729
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize2) finish = true;
            }
        }
    }
    result.reserve(nodes.size());
    for (const Node& n : nodes) {  // :742-763, first wins ties
        int best = n.keys[0];
        float maxResp = c[best].response;
        for (size_t k = 1; k < n.keys.size(); k++)
            if (c[n.keys[k]].response > maxResp) { best = n.keys[k]; maxResp = c[best].response; }
        result.push_back(best);
    }
    return result;
}

// cv::fastAtan2 (core/mathfuncs_core), called at src/ORBextractor.cc:104
float fast_atan2f(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// IC_Angle, src/ORBextractor.cc:78-105
float ic_angle(const Image& L, float ptx, float pty, const int* umax, int* m01_out = nullptr, int* m10_out = nullptr) {
    int m_01 = 0, m_10 = 0;
    const int step = L.w;
    const uint8_t* center = L.row(cvRoundF(pty)) + cvRoundF(ptx);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    if (m01_out) *m01_out = m_01;
    if (m10_out) *m10_out = m_10;
    return fast_atan2f((float)m_01, (float)m_10);
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8U
// (OpenCV bit-exact fixed-point path).  Called at src/ORBextractor.cc:1087.
inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i; else i = 2 * (n - 1) - i;
    }
    return i;
}

void gaussian_blur7(const Image& S, Image& D) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    D.w = S.w; D.h = S.h; D.px.resize(S.px.size());
    std::vector<uint32_t> H((size_t)S.w * S.h);
    for (int y = 0; y < S.h; y++) {
        const uint8_t* r = S.row(y);
        for (int x = 0; x < S.w; x++) {
            uint32_t acc = 0;
            for (int i = 0; i < 7; i++) acc += (uint32_t)k[i] * r[reflect101(x + i - 3, S.w)];
            H[(size_t)y * S.w + x] = acc;  // 8.8 fixed point, <= 65280
        }
    }
    for (int y = 0; y < S.h; y++)
        for (int x = 0; x < S.w; x++) {
            uint32_t acc = 0;
            for (int i = 0; i < 7; i++) acc += (uint32_t)k[i] * H[(size_t)reflect101(y + i - 3, S.h) * S.w + x];
            D.px[(size_t)y * S.w + x] = (uint8_t)((acc + 32768u) >> 16);
        }
}

// computeOrbDescriptor, src/ORBextractor.cc:107-148.
// cos/sin: contract A9: (float)cos((double)angle), (float)sin((double)angle).
void orb_descriptor(const Image& B, float ptx, float pty, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);  // :107
    const float angle = angle_deg * factorPI;
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const uint8_t* center = B.row(cvRoundF(pty)) + cvRoundF(ptx);
    const int step = B.w;
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; j++) {
            const float x0 = (float)pat[4 * j + 0], y0 = (float)pat[4 * j + 1];
            const float x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
            const int t0 = center[cvRoundF(x0 * b + y0 * a) * step + cvRoundF(x0 * a - y0 * b)];
            const int t1 = center[cvRoundF(x1 * b + y1 * a) * step + cvRoundF(x1 * a - y1 * b)];
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

struct ExtractDump {  // optional per-stage outputs for the parity gates
    std::vector<Image> pyramid, blurred;
    std::vector<std::vector<Cand>> cands;      // per level, reference order
    std::vector<std::vector<KeyPoint>> lvlkps; // per level after octree + angle (level coordinates)
    int nfallback = 0;
};

// ORBextractor::operator(), src/ORBextractor.cc:1045-1106
int extract(const OrbParams& p, const uint8_t* img, int w, int h, int pitch, std::vector<KeyPoint>& kps,
            std::vector<uint8_t>& desc, ExtractDump* dump) {
    kps.clear(); desc.clear();
    if (!img || w <= 0 || h <= 0) return 0;  // :1048
    OrbTables t = make_tables(p);
    std::vector<Image> pyr = compute_pyramid(img, w, h, pitch, t);
    std::vector<std::vector<KeyPoint>> all(p.nlevels);
    std::vector<std::vector<Cand>> allc(p.nlevels);
    int nfallback = 0;
    for (int level = 0; level < p.nlevels; ++level) {  // ComputeKeyPointsOctTree :766-849
        const Image& L = pyr[level];
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = L.w - EDGE_THRESHOLD + 3, maxBorderY = L.h - EDGE_THRESHOLD + 3;
        std::vector<Cand>& cands = allc[level];
        fast_level_candidates(L, p.iniThFAST, p.minThFAST, cands, &nfallback);
        std::vector<int> sel = distribute_octree(cands, minBorderX, maxBorderX, minBorderY, maxBorderY, t.nPerLevel[level]);
        const int scaledPatchSize = (int)(PATCH_SIZE * t.scale[level]);  // :838
        for (int idx : sel) {
            KeyPoint k;
            k.x = cands[idx].x + minBorderX; k.y = cands[idx].y + minBorderY;
            k.size = (float)scaledPatchSize; k.angle = -1; k.response = cands[idx].response;
            k.octave = level; k.class_id = -1;
            all[level].push_back(k);
        }
    }
    for (int level = 0; level < p.nlevels; ++level)  // :852-853
        for (KeyPoint& k : all[level]) k.angle = ic_angle(pyr[level], k.x, k.y, t.umax);
    std::vector<Image> blurred(p.nlevels);
    for (int level = 0; level < p.nlevels; ++level) {  // :1077-1105
        std::vector<KeyPoint>& lk = all[level];
        if (lk.empty()) continue;
        gaussian_blur7(pyr[level], blurred[level]);
        const size_t off = desc.size();
        desc.resize(off + lk.size() * 32);
        for (size_t i = 0; i < lk.size(); i++) orb_descriptor(blurred[level], lk[i].x, lk[i].y, lk[i].angle, &desc[off + i * 32]);
        if (dump) dump->lvlkps.push_back(lk);
        if (level != 0) {
            const float scale = t.scale[level];
            for (KeyPoint& k : lk) { k.x *= scale; k.y *= scale; }
        }
        kps.insert(kps.end(), lk.begin(), lk.end());
    }
    if (dump) {
        dump->pyramid = std::move(pyr); dump->blurred = std::move(blurred);
        dump->cands = std::move(allc); dump->nfallback = nfallback;
    }
    return (int)kps.size();
}

// ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1649-1665 (== FORB::distance, DBoW2/FORB.cpp:81-101)
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        std::memcpy(&pa, a + 4 * i, 4); std::memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// ---------------------------------------------------------------------------------------
// Frame grid: AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, src/Frame.cc:257-272,
// 354-419.  FRAME_GRID_COLS 64, ROWS 48 (include/Frame.h:39-40).
// ---------------------------------------------------------------------------------------
constexpr int GRID_COLS = 64, GRID_ROWS = 48;

struct FrameView {          // flattened, read-only view of a Frame
    int N;
    const KeyPoint* keysUn; // mvKeysUn
    const float* uRight;    // mvuRight
    const uint8_t* desc;    // mDescriptors N x 32
    float minX, minY, maxX, maxY;  // mnMinX.. (Frame.cc:686-714)
    float fx, fy, cx, cy, bf, b;   // mb = mbf/fx (Frame.cc:196)
    int nlevels;
    const float* scaleFactors;
    float logScaleFactor;
};

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    float wInv, hInv;
};

void build_grid(const FrameView& f, Grid& g) {
    g.wInv = (float)GRID_COLS / (f.maxX - f.minX);  // Frame.cc:183
    g.hInv = (float)GRID_ROWS / (f.maxY - f.minY);
    for (int i = 0; i < f.N; i++) {
        const int px = (int)std::round((f.keysUn[i].x - f.minX) * g.wInv);  // :411
        const int py = (int)std::round((f.keysUn[i].y - f.minY) * g.hInv);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        g.cell[px][py].push_back(i);
    }
}

void features_in_area(const FrameView& f, const Grid& g, float x, float y, float r, int minLevel, int maxLevel,
                      std::vector<int>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - f.minX - r) * g.wInv));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - f.minX + r) * g.wInv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - f.minY - r) * g.hInv));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - f.minY + r) * g.hInv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
            for (int idx : g.cell[ix][iy]) {
                const KeyPoint& kp = f.keysUn[idx];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float distx = kp.x - x, disty = kp.y - y;
                if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
            }
}

// ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1603-1644
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;  // src/ORBmatcher.cc:37-39

}  // namespace

// =========================================================================================
// C entry points (ctypes) -- test infrastructure only
// =========================================================================================
SGO_API int sgo_orb_tables(const OrbParams* p, float* scale, float* invScale, float* sigma2, float* invSigma2,
                           int32_t* nPerLevel, int32_t* umax16) {
    OrbTables t = make_tables(*p);
    for (int i = 0; i < p->nlevels; i++) {
        if (scale) scale[i] = t.scale[i];
        if (invScale) invScale[i] = t.invScale[i];
        if (sigma2) sigma2[i] = t.sigma2[i];
        if (invSigma2) invSigma2[i] = t.invSigma2[i];
        if (nPerLevel) nPerLevel[i] = t.nPerLevel[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = t.umax[i];
    return 0;
}

SGO_API int sgo_level_size(const OrbParams* p, int w, int h, int level, int32_t* lw, int32_t* lh) {
    OrbTables t = make_tables(*p);
    int a, b; level_size(w, h, t, level, &a, &b); *lw = a; *lh = b;
    return 0;
}

SGO_API int sgo_resize(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch) {
    resize_linear_u8(src, sw, sh, spitch, dst, dw, dh, dpitch);
    return 0;
}

// cv::FAST on a view; out = (x, y, score) int32 triples; returns count (or -needed if cap too small)
SGO_API int sgo_fast_view(const uint8_t* base, int pitch, int w, int h, int threshold, int nms, int32_t* out, int cap) {
    std::vector<CellKp> k;
    fast_detect_view(base, pitch, w, h, threshold, nms != 0, k);
    if ((int)k.size() > cap) return -(int)k.size();
    for (size_t i = 0; i < k.size(); i++) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].score; }
    return (int)k.size();
}

SGO_API int sgo_fast_score(const uint8_t* p, int pitch) { return fast_corner_score(p, pitch); }

// per-level FAST candidates (float x, y, response relative to (16,16)), reference order
SGO_API int sgo_fast_level(const uint8_t* img, int w, int h, int pitch, int iniTh, int minTh, float* out, int cap,
                           int32_t* nfallback) {
    Image L; L.w = w; L.h = h; L.px.resize((size_t)w * h);
    for (int y = 0; y < h; y++) std::memcpy(L.row(y), img + (size_t)y * pitch, w);
    std::vector<Cand> c; int nf = 0;
    fast_level_candidates(L, iniTh, minTh, c, &nf);
    if (nfallback) *nfallback = nf;
    if ((int)c.size() > cap) return -(int)c.size();
    for (size_t i = 0; i < c.size(); i++) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].response; }
    return (int)c.size();
}

// DistributeOctTree on explicit candidates; returns selected candidate indices in output order
SGO_API int sgo_octree(const float* cands, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* sel, int cap) {
    std::vector<Cand> c(n);
    for (int i = 0; i < n; i++) c[i] = {cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]};
    std::vector<int> r = distribute_octree(c, minX, maxX, minY, maxY, N);
    if ((int)r.size() > cap) return -(int)r.size();
    for (size_t i = 0; i < r.size(); i++) sel[i] = r[i];
    return (int)r.size();
}

SGO_API float sgo_fast_atan2(float y, float x) { return fast_atan2f(y, x); }

SGO_API float sgo_ic_angle(const uint8_t* img, int w, int h, int pitch, float x, float y, int32_t* m01, int32_t* m10) {
    Image L; L.w = w; L.h = h; L.px.resize((size_t)w * h);
    for (int r = 0; r < h; r++) std::memcpy(L.row(r), img + (size_t)r * pitch, w);
    OrbParams p{1000, 1.2f, 8, 20, 7};
    OrbTables t = make_tables(p);
    int a, b;
    float ang = ic_angle(L, x, y, t.umax, &a, &b);
    if (m01) *m01 = a;
    if (m10) *m10 = b;
    return ang;
}

SGO_API int sgo_blur(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch) {
    Image S, D; S.w = w; S.h = h; S.px.resize((size_t)w * h);
    for (int y = 0; y < h; y++) std::memcpy(S.row(y), src + (size_t)y * spitch, w);
    gaussian_blur7(S, D);
    for (int y = 0; y < h; y++) std::memcpy(dst + (size_t)y * dpitch, D.row(y), w);
    return 0;
}

SGO_API int sgo_brief(const uint8_t* blurred, int w, int h, int pitch, float x, float y, float angle, uint8_t* desc32) {
    Image B; B.w = w; B.h = h; B.px.resize((size_t)w * h);
    for (int r = 0; r < h; r++) std::memcpy(B.row(r), blurred + (size_t)r * pitch, w);
    orb_descriptor(B, x, y, angle, desc32);
    return 0;
}

// max |rotated tap offset| over all 512 pattern points for a given angle (degrees) -- used by tests
// to prove that descriptor taps never leave the 15-px disc (so image borders are never read).
SGO_API int sgo_pattern_extent(float angle_deg) {
    const float factorPI = (float)(M_PI / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    int ext = 0;
    for (int i = 0; i < 512; i++) {
        const float px = (float)kPattern[2 * i], py = (float)kPattern[2 * i + 1];
        ext = std::max(ext, std::abs(cvRoundF(px * b + py * a)));
        ext = std::max(ext, std::abs(cvRoundF(px * a - py * b)));
    }
    return ext;
}

// Full extractor.  kps = cv::KeyPoint-layout records; desc = n x 32.  Returns n (or -needed).
SGO_API int sgo_extract(const OrbParams* p, const uint8_t* img, int w, int h, int pitch, KeyPoint* kps, uint8_t* desc, int cap) {
    std::vector<KeyPoint> k; std::vector<uint8_t> d;
    extract(*p, img, w, h, pitch, k, d, nullptr);
    if ((int)k.size() > cap) return -(int)k.size();
    if (!k.empty()) { std::memcpy(kps, k.data(), k.size() * sizeof(KeyPoint)); std::memcpy(desc, d.data(), d.size()); }
    return (int)k.size();
}

// Stage dumps for the parity gates.  Opaque handle API.
struct SgoDump { ExtractDump d; std::vector<KeyPoint> kps; std::vector<uint8_t> desc; };

SGO_API SgoDump* sgo_extract_dump(const OrbParams* p, const uint8_t* img, int w, int h, int pitch) {
    SgoDump* s = new SgoDump();
    extract(*p, img, w, h, pitch, s->kps, s->desc, &s->d);
    return s;
}
SGO_API void sgo_dump_free(SgoDump* s) { delete s; }
SGO_API int sgo_dump_nkps(SgoDump* s) { return (int)s->kps.size(); }
SGO_API int sgo_dump_nfallback(SgoDump* s) { return s->d.nfallback; }
SGO_API void sgo_dump_kps(SgoDump* s, KeyPoint* kps, uint8_t* desc) {
    if (s->kps.empty()) return;
    std::memcpy(kps, s->kps.data(), s->kps.size() * sizeof(KeyPoint));
    std::memcpy(desc, s->desc.data(), s->desc.size());
}
SGO_API int sgo_dump_level(SgoDump* s, int level, int which /*0 pyramid, 1 blurred*/, int32_t* w, int32_t* h, uint8_t* out) {
    const std::vector<Image>& v = which ? s->d.blurred : s->d.pyramid;
    if (level < 0 || level >= (int)v.size()) return -1;
    *w = v[level].w; *h = v[level].h;
    if (out && !v[level].px.empty()) std::memcpy(out, v[level].px.data(), v[level].px.size());
    return (int)v[level].px.size();
}
SGO_API int sgo_dump_cands(SgoDump* s, int level, float* out, int cap) {
    const std::vector<Cand>& c = s->d.cands[level];
    if (out) {
        if ((int)c.size() > cap) return -(int)c.size();
        for (size_t i = 0; i < c.size(); i++) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].response; }
    }
    return (int)c.size();
}

SGO_API int sgo_hamming(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// Brute-force nearest / second-nearest (first index wins ties) -- restates the inner loop shared by every
// ORBmatcher search (src/ORBmatcher.cc:93-113) over an unrestricted candidate set.
SGO_API int sgo_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* best_idx, int32_t* best_d, int32_t* second_d) {
    for (int i = 0; i < nq; i++) {
        int bd = 256, bd2 = 256, bi = -1;
        for (int j = 0; j < nt; j++) {
            const int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < bd) { bd2 = bd; bd = d; bi = j; }
            else if (d < bd2) bd2 = d;
        }
        best_idx[i] = bi; best_d[i] = bd; second_d[i] = bd2;
    }
    return 0;
}

// ---- flattened Frame view shared by the matcher entry points -----------------------------------
struct SgoFrame {
    int32_t N;
    const KeyPoint* keysUn;
    const float* uRight;
    const uint8_t* desc;
    float minX, minY, maxX, maxY;
    float fx, fy, cx, cy, bf;
    int32_t nlevels;
    const float* scaleFactors;
    float logScaleFactor;
};

static FrameView to_view(const SgoFrame* f) {
    FrameView v;
    v.N = f->N; v.keysUn = f->keysUn; v.uRight = f->uRight; v.desc = f->desc;
    v.minX = f->minX; v.minY = f->minY; v.maxX = f->maxX; v.maxY = f->maxY;
    v.fx = f->fx; v.fy = f->fy; v.cx = f->cx; v.cy = f->cy; v.bf = f->bf; v.b = f->bf / f->fx;
    v.nlevels = f->nlevels; v.scaleFactors = f->scaleFactors; v.logScaleFactor = f->logScaleFactor;
    return v;
}

SGO_API int sgo_features_in_area(const SgoFrame* f, float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap) {
    FrameView v = to_view(f); Grid g; build_grid(v, g);
    std::vector<int> idx; features_in_area(v, g, x, y, r, minLevel, maxLevel, idx);
    if ((int)idx.size() > cap) return -(int)idx.size();
    for (size_t i = 0; i < idx.size(); i++) out[i] = idx[i];
    return (int)idx.size();
}

// ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono), src/ORBmatcher.cc:1332-1472.
// Flattening of the object graph:
//   last_has_mp[i]   LastFrame.mvpMapPoints[i] != NULL && !LastFrame.mvbOutlier[i]
//   last_xyz[3i..]   pMP->GetWorldPos()           last_desc[32 i..]  pMP->GetDescriptor()
//   last_obs[i]      pMP->Observations() > 0      last_octave/last_angle  LastFrame.mvKeys[i].octave / mvKeysUn[i].angle
//   cur_mp_inout[j]  index i of the last-frame point held by CurrentFrame.mvpMapPoints[j] (or -1); on entry
//                    it is what the caller left there (Tracking.cc:916 clears it), cur_mp_obs_in[j] says whether a
//                    pre-existing entry has Observations()>0.
// Returns nmatches exactly as the reference counts them (including the double counting quirk).
SGO_API int sgo_search_by_projection_last(const SgoFrame* cur, const float* Tcw_cur /*4x4 row major*/, const float* Tcw_last,
                                          int nlast, const uint8_t* last_has_mp, const float* last_xyz,
                                          const uint8_t* last_desc, const uint8_t* last_obs, const int32_t* last_octave,
                                          const float* last_angle, float th, int bMono, int checkOri,
                                          int32_t* cur_mp_inout, const uint8_t* cur_mp_obs_in, int64_t* ncand_out) {
    FrameView F = to_view(cur); Grid g; build_grid(F, g);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    // Rcw, tcw, twc = -Rcw^T tcw, tlc = Rlw twc + tlw  (:1342-1351)  float32 cv::Mat arithmetic
    float Rcw[9], tcw[3], Rlw[9], tlw[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { Rcw[3 * r + c] = Tcw_cur[4 * r + c]; Rlw[3 * r + c] = Tcw_last[4 * r + c]; }
        tcw[r] = Tcw_cur[4 * r + 3]; tlw[r] = Tcw_last[4 * r + 3]; }
    float twc[3], tlc[3];
    for (int r = 0; r < 3; r++) {  // -Rcw.t()*tcw: transposed gemm takes the general path, double accumulator then one cast (GEMMSingleMul)
        double acc = 0; for (int k = 0; k < 3; k++) acc += (double)(-Rcw[3 * k + r]) * (double)tcw[k];
        twc[r] = (float)acc;
    }
    for (int r = 0; r < 3; r++) {   // MatExpr A*B+C == gemm(A,B,1,C,1) with flags == 0: OpenCV's small-matrix path sums float products in float,
        const float acc = Rlw[3 * r] * twc[0] + Rlw[3 * r + 1] * twc[1] + Rlw[3 * r + 2] * twc[2];          // left to right (pinned against cv2.gemm)
        tlc[r] = (float)((double)acc + (double)tlw[r]);
    }
    const bool bForward = tlc[2] > F.b && !bMono;
    const bool bBackward = -tlc[2] > F.b && !bMono;
    std::vector<int> cand;
    std::vector<uint8_t> cur_obs(F.N, 0);
    for (int j = 0; j < F.N; j++) cur_obs[j] = (cur_mp_inout[j] >= 0) ? (cur_mp_obs_in ? cur_mp_obs_in[j] : 1) : 0;
    int64_t ncand = 0;
    for (int i = 0; i < nlast; i++) {
        if (!last_has_mp[i]) continue;
        const float* Xw = last_xyz + 3 * i;
        float x3Dc[3];
        for (int r = 0; r < 3; r++) {   // gemm(Rcw, x3Dw, 1, tcw, 1): small-matrix path, float accumulation
            const float acc = Rcw[3 * r] * Xw[0] + Rcw[3 * r + 1] * Xw[1] + Rcw[3 * r + 2] * Xw[2];
            x3Dc[r] = (float)((double)acc + (double)tcw[r]);
        }
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / x3Dc[2]);  // :1369 (double division, stored to float)
        if (invzc < 0) continue;
        const float u = F.fx * xc * invzc + F.cx;
        const float v = F.fy * yc * invzc + F.cy;
        if (u < F.minX || u > F.maxX) continue;
        if (v < F.minY || v > F.maxY) continue;
        const int nLastOctave = last_octave[i];
        const float radius = th * F.scaleFactors[nLastOctave];
        if (bForward) features_in_area(F, g, u, v, radius, nLastOctave, -1, cand);
        else if (bBackward) features_in_area(F, g, u, v, radius, 0, nLastOctave, cand);
        else features_in_area(F, g, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
        if (cand.empty()) continue;
        ncand += (int64_t)cand.size();
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (cur_mp_inout[i2] >= 0 && cur_obs[i2]) continue;  // :1407-1409
            if (F.uRight[i2] > 0) {
                const float ur = u - F.bf * invzc;
                const float er = std::fabs(ur - F.uRight[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(last_desc + 32 * (size_t)i, F.desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            cur_mp_inout[bestIdx2] = i;  // :1432 last writer wins
            cur_obs[bestIdx2] = last_obs[i];
            nmatches++;
            if (checkOri) {
                float rot = last_angle[i] - F.keysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i]) { cur_mp_inout[j] = -1; nmatches--; }
    }
    if (ncand_out) *ncand_out = ncand;
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist),
// src/ORBmatcher.cc:1474-1601 (Tracking::Relocalization, src/Tracking.cc:1494,1508).  kf_valid[i]: pKF's i-th map point exists, is not bad and
// is not in sAlreadyFound; kf_angle[i] = pKF->mvKeysUn[i].angle; min/max_dist are mfMinDistance / mfMaxDistance (raw).
// cur_mp_inout[j] >= 0: CurrentFrame.mvpMapPoints[j] is set (any such keypoint is skipped, :1543-1544).  New matches write the index i.
SGO_API int sgo_search_by_projection_kf(const SgoFrame* cur, const float* Tcw_cur, int nkf, const uint8_t* kf_valid, const float* kf_xyz,
                                        const uint8_t* kf_desc, const float* kf_angle, const float* min_dist, const float* max_dist, float th,
                                        int ORBdist, int checkOri, float log_scale_factor, int32_t* cur_mp_inout, int64_t* ncand_out) {
    FrameView F = to_view(cur); Grid g; build_grid(F, g);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    float Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Rcw[3 * r + c] = Tcw_cur[4 * r + c]; tcw[r] = Tcw_cur[4 * r + 3]; }
    for (int r = 0; r < 3; r++) {   // Ow = -Rcw.t()*tcw: transposed gemm, double accumulator
        double acc = 0; for (int k = 0; k < 3; k++) acc += (double)(-Rcw[3 * k + r]) * (double)tcw[k];
        Ow[r] = (float)acc;
    }
    std::vector<int> cand;
    int64_t ncand = 0;
    for (int i = 0; i < nkf; i++) {
        if (!kf_valid[i]) continue;
        const float* Xw = kf_xyz + 3 * i;
        float x3Dc[3];
        for (int r = 0; r < 3; r++) {   // gemm(Rcw, x3Dw, 1, tcw, 1): small-matrix path, float accumulation
            const float acc = Rcw[3 * r] * Xw[0] + Rcw[3 * r + 1] * Xw[1] + Rcw[3 * r + 2] * Xw[2];
            x3Dc[r] = (float)((double)acc + (double)tcw[r]);
        }
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / x3Dc[2]);                 // :1506, no sign test in this variant
        const float u = F.fx * xc * invzc + F.cx;
        const float v = F.fy * yc * invzc + F.cy;
        if (u < F.minX || u > F.maxX) continue;
        if (v < F.minY || v > F.maxY) continue;
        const float PO[3] = {Xw[0] - Ow[0], Xw[1] - Ow[1], Xw[2] - Ow[2]};
        const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float ratio = max_dist[i] / dist3D;                   // MapPoint::PredictScale (src/MapPoint.cc:400-418)
        int nPredictedLevel = (int)std::ceil(std::log(ratio) / log_scale_factor);
        if (nPredictedLevel < 0) nPredictedLevel = 0; else if (nPredictedLevel >= F.nlevels) nPredictedLevel = F.nlevels - 1;
        const float radius = th * F.scaleFactors[nPredictedLevel];
        features_in_area(F, g, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, cand);
        if (cand.empty()) continue;
        ncand += (int64_t)cand.size();
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (cur_mp_inout[i2] >= 0) continue;                    // :1543-1544
            const int dist = descriptor_distance(kf_desc + 32 * (size_t)i, F.desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= ORBdist) {
            cur_mp_inout[bestIdx2] = i;
            nmatches++;
            if (checkOri) {
                float rot = kf_angle[i] - F.keysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i]) { cur_mp_inout[j] = -1; nmatches--; }
    }
    if (ncand_out) *ncand_out = ncand;
    return nmatches;
}

// The SEARCH half of ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th) (src/ORBmatcher.cc:829-980, called from
// LocalMapping::SearchInNeighbors): for every candidate map point (mp_valid[i]: exists, !isBad(), !IsInKeyFrame(pKF)) the key-frame
// feature it would be fused with -- best_idx[i] (-1: nothing passed the gates) and best_dist[i] (256 likewise).  The caller applies
// "bestDist <= TH_LOW" and the map side effects (Replace / AddObservation) in order, exactly as the reference loop does; they do not
// influence the search of later points.  kf: the key frame's mvKeysUn / mvuRight / mDescriptors and grid; Ow = pKF->GetCameraCenter().
SGO_API int sgo_fuse_search(const SgoFrame* kf, const float* Tcw, const float* Ow, int nmp, const uint8_t* mp_valid, const float* mp_xyz, const float* mp_normal,
                            const float* min_dist, const float* max_dist, const uint8_t* mp_desc, float th, const float* inv_level_sigma2,
                            float log_scale_factor, int sim3_variant, const float* xform2, int32_t* best_idx, int32_t* best_dist) {
    // sim3_variant == 2: one direction of ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1106-1330): p3Dc1 = R1w p3Dw + t1w (Tcw), p3Dc2 = sR21 p3Dc1 + t21
    // (xform2 = 3x3 row major + 3, computed by the caller :1122-1124), distance = |p3Dc2|, no viewing-angle test, no chi-square gates; the caller
    // applies bestDist <= TH_HIGH and the mutual-agreement check (:1314-1327)
    // sim3_variant: Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:982-1104, loop closing) with Rcw / tcw / Ow already
    // decomposed from Scw by the caller (:988-992): invz = 1.0 / z in double, no chi-square gates
    FrameView F = to_view(kf); Grid g; build_grid(F, g);
    float Rcw[9], tcw[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Rcw[3 * r + c] = Tcw[4 * r + c]; tcw[r] = Tcw[4 * r + 3]; }
    std::vector<int> cand;
    int nfound = 0;
    for (int i = 0; i < nmp; i++) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (!mp_valid[i]) continue;
        const float* Xw = mp_xyz + 3 * i;
        float p3Dc[3];
        for (int r = 0; r < 3; r++) {   // gemm(Rcw, p3Dw, 1, tcw, 1): small-matrix path, float accumulation
            const float acc = Rcw[3 * r] * Xw[0] + Rcw[3 * r + 1] * Xw[1] + Rcw[3 * r + 2] * Xw[2];
            p3Dc[r] = (float)((double)acc + (double)tcw[r]);
        }
        if (sim3_variant == 2) {
            float q[3];
            for (int r = 0; r < 3; r++) {
                const float acc = xform2[3 * r] * p3Dc[0] + xform2[3 * r + 1] * p3Dc[1] + xform2[3 * r + 2] * p3Dc[2];
                q[r] = (float)((double)acc + (double)xform2[9 + r]);
            }
            p3Dc[0] = q[0]; p3Dc[1] = q[1]; p3Dc[2] = q[2];
        }
        if (p3Dc[2] < 0.0f) continue;
        const float invz = sim3_variant ? (float)(1.0 / p3Dc[2]) : 1 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = F.fx * x + F.cx, v = F.fy * y + F.cy;
        if (!(u >= F.minX && u < F.maxX && v >= F.minY && v < F.maxY)) continue;        // KeyFrame::IsInImage
        const float ur = u - F.bf * invz;
        const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
        float PO[3] = {Xw[0] - Ow[0], Xw[1] - Ow[1], Xw[2] - Ow[2]};
        if (sim3_variant == 2) { PO[0] = p3Dc[0]; PO[1] = p3Dc[1]; PO[2] = p3Dc[2]; }      // cv::norm(p3Dc2)
        const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float* Pn = mp_normal + 3 * i;
        if (sim3_variant != 2 && ((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) < 0.5 * dist3D) continue;
        const float ratio = max_dist[i] / dist3D;
        int nPredictedLevel = (int)std::ceil(std::log(ratio) / log_scale_factor);
        if (nPredictedLevel < 0) nPredictedLevel = 0; else if (nPredictedLevel >= F.nlevels) nPredictedLevel = F.nlevels - 1;
        const float radius = th * F.scaleFactors[nPredictedLevel];
        features_in_area(F, g, u, v, radius, -1, -1, cand);                             // KeyFrame::GetFeaturesInArea: no level filter
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const int kpLevel = F.keysUn[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const float kpx = F.keysUn[idx].x, kpy = F.keysUn[idx].y;
            const float ex = u - kpx, ey = v - kpy;
            if (sim3_variant) {
            } else if (F.uRight[idx] >= 0) {
                const float er = ur - F.uRight[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
            }
            const int dist = descriptor_distance(mp_desc + 32 * (size_t)i, F.desc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        best_idx[i] = bestIdx; best_dist[i] = bestDist;
        if (bestIdx >= 0) nfound++;
    }
    return nfound;
}

// ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize), src/ORBmatcher.cc:407-522 (monocular
// initialisation, Tracking::MonocularInitialization).  f1: mvKeysUn / mDescriptors of the reference frame; f2: the current frame with its grid.
// prev_xy [n1][2] = vbPrevMatched (in/out: matched entries receive the matched keypoint's position, :517-519); match12 [n1] out.
// The loop is order dependent: a feature of F2 keeps the best distance seen so far (vMatchedDistance) and a better later match steals it (:466-471).
SGO_API int sgo_search_for_initialization(const SgoFrame* f1, const SgoFrame* f2, float* prev_xy, int window_size, float nnratio, int check_ori,
                                          int32_t* match12) {
    FrameView F1 = to_view(f1), F2 = to_view(f2); Grid g; build_grid(F2, g);
    int nmatches = 0;
    for (int i = 0; i < F1.N; i++) match12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    std::vector<int> vMatchedDistance(F2.N, INT_MAX), vnMatches21(F2.N, -1);
    std::vector<int> cand;
    for (int i1 = 0; i1 < F1.N; i1++) {
        const int level1 = F1.keysUn[i1].octave;
        if (level1 > 0) continue;
        features_in_area(F2, g, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window_size, level1, level1, cand);
        if (cand.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = descriptor_distance(F1.desc + 32 * (size_t)i1, F2.desc + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { match12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                match12[i1] = bestIdx2; vnMatches21[bestIdx2] = i1; vMatchedDistance[bestIdx2] = bestDist; nmatches++;
                if (check_ori) {
                    float rot = F1.keysUn[i1].angle - F2.keysUn[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (match12[idx1] >= 0) { match12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < F1.N; i1++)
        if (match12[i1] >= 0) { prev_xy[2 * i1] = F2.keysUn[match12[i1]].x; prev_xy[2 * i1 + 1] = F2.keysUn[match12[i1]].y; }
    return nmatches;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th),
// src/ORBmatcher.cc:292-405 (loop closing, LoopClosing.cc:239 / :589).  Rcw, tcw (Tcw rows) and Ow are the values the caller decomposed from Scw
// (:301-305) with the reference's own cv::Mat expressions.  mp_valid[i] = !isBad() && not in spAlreadyFound (:308-320).  kf_matched[idx] is
// vpMatched: >= 0 means occupied on entry (any id); a feature claimed here receives the index i of the claiming point.  Points are visited in
// order and a feature taken by an earlier point is skipped by later ones (:374-375), so the result depends on the order.
SGO_API int sgo_search_by_projection_sim3(const SgoFrame* kf, const float* Tcw, const float* Ow, int nmp, const uint8_t* mp_valid, const float* mp_xyz,
                                          const float* mp_normal, const float* min_dist, const float* max_dist, const uint8_t* mp_desc, float th,
                                          float log_scale_factor, int32_t* kf_matched) {
    FrameView F = to_view(kf); Grid g; build_grid(F, g);
    float Rcw[9], tcw[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Rcw[3 * r + c] = Tcw[4 * r + c]; tcw[r] = Tcw[4 * r + 3]; }
    std::vector<int> cand;
    int nmatches = 0;
    for (int i = 0; i < nmp; i++) {
        if (!mp_valid[i]) continue;
        const float* Xw = mp_xyz + 3 * i;
        float p3Dc[3];
        for (int r = 0; r < 3; r++) {   // Rcw*p3Dw+tcw: small-matrix gemm path, float accumulation (:329)
            const float acc = Rcw[3 * r] * Xw[0] + Rcw[3 * r + 1] * Xw[1] + Rcw[3 * r + 2] * Xw[2];
            p3Dc[r] = (float)((double)acc + (double)tcw[r]);
        }
        if (p3Dc[2] < 0.0f) continue;                                                     // :332
        const float invz = 1 / p3Dc[2];                                                   // :336, float division
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = F.fx * x + F.cx, v = F.fy * y + F.cy;
        if (!(u >= F.minX && u < F.maxX && v >= F.minY && v < F.maxY)) continue;          // KeyFrame::IsInImage (:344)
        const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];   // Get{Max,Min}DistanceInvariance
        const float PO[3] = {Xw[0] - Ow[0], Xw[1] - Ow[1], Xw[2] - Ow[2]};
        const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        if (dist < minDistance || dist > maxDistance) continue;                            // :353
        const float* Pn = mp_normal + 3 * i;
        if (((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) < 0.5 * dist) continue;   // :359
        int nPredictedLevel = (int)std::ceil(std::log(max_dist[i] / dist) / log_scale_factor);               // MapPoint::PredictScale(dist, pKF)
        if (nPredictedLevel < 0) nPredictedLevel = 0; else if (nPredictedLevel >= F.nlevels) nPredictedLevel = F.nlevels - 1;
        const float radius = th * F.scaleFactors[nPredictedLevel];
        features_in_area(F, g, u, v, radius, -1, -1, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            if (kf_matched[idx] >= 0) continue;                                             // :374
            const int kpLevel = F.keysUn[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int d = descriptor_distance(mp_desc + 32 * (size_t)i, F.desc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) { kf_matched[bestIdx] = i; nmatches++; }                   // :393-397
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129, with the
// per-point fields that Frame::isInFrustum (src/Frame.cc:296-352) fills given as flat arrays:
//   mp_inview[i] (mbTrackInView && !isBad()), projx/projy/projxr, level (mnTrackScaleLevel), viewcos, desc.
//   f_mp_inout[j] >= 0 : F.mvpMapPoints[j] set (value = opaque id), f_mp_obs[j]: Observations()>0 of that point
// New matches write (id_base + i) into f_mp_inout and mark obs from mp_obs[i].
SGO_API int sgo_search_by_projection_local(const SgoFrame* fr, int nmp, const uint8_t* mp_inview, const float* projx,
                                           const float* projy, const float* projxr, const int32_t* level,
                                           const float* viewcos, const uint8_t* mp_desc, const uint8_t* mp_obs,
                                           float th, float nnratio, int32_t id_base, int32_t* f_mp_inout,
                                           uint8_t* f_mp_obs_inout, int64_t* ncand_out) {
    FrameView F = to_view(fr); Grid g; build_grid(F, g);
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> cand;
    int64_t ncand = 0;
    for (int iMP = 0; iMP < nmp; iMP++) {
        if (!mp_inview[iMP]) continue;
        const int nPredictedLevel = level[iMP];
        float r = viewcos[iMP] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos :131-137
        if (bFactor) r *= th;
        features_in_area(F, g, projx[iMP], projy[iMP], r * F.scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, cand);
        if (cand.empty()) continue;
        ncand += (int64_t)cand.size();
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (f_mp_inout[idx] >= 0 && f_mp_obs_inout[idx]) continue;
            if (F.uRight[idx] > 0) {
                const float er = std::fabs(projxr[iMP] - F.uRight[idx]);
                if (er > r * F.scaleFactors[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(mp_desc + 32 * (size_t)iMP, F.desc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.keysUn[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = F.keysUn[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            f_mp_inout[bestIdx] = id_base + iMP;
            f_mp_obs_inout[bestIdx] = mp_obs[iMP];
            nmatches++;
        }
    }
    if (ncand_out) *ncand_out = ncand;
    return nmatches;
}

// ---------------------------------------------------------------------------------------
// Dynamic-feature rejection, geometry half: Frame::RmDynamicPointWithSemanticAndGeometry
// "version3" loop (src/Frame.cc:560-604) + CheckEpiLineDistToRmDynamicPoint (:613-627) +
// isInDynamicRegion (:629-652).  LK flow (prev points) and F come from the caller.
//   keep[i] 1 iff keypoint i survives; dist[i] = epipolar distance (double).
//   Returns Cur_keypoint_sum; *restored = 1 when the restore-all branch (:599-602) fired
//   (then keep[] still reports the per-point verdicts but the caller keeps every keypoint).
// Quirk Q11: F == NULL (empty matrix) is DEFINED as "keep everything".
// ---------------------------------------------------------------------------------------
SGO_API int sgo_dynreject(const float* cur_xy, const float* prev_xy, int n, const double* F /*9, row major*/,
                          const float* boxes /*x,y,w,h*/, int nboxes, int have_dyn, int nfeatures, uint8_t* keep,
                          double* dist_out, int32_t* restored) {
    int sum = n;
    for (int i = 0; i < n; i++) {
        const float x = cur_xy[2 * i], y = cur_xy[2 * i + 1];
        bool in_box = false;
        if (have_dyn)
            for (int b = 0; b < nboxes; b++) {
                const float bx = boxes[4 * b], by = boxes[4 * b + 1], bw = boxes[4 * b + 2], bh = boxes[4 * b + 3];
                if (x > bx && x < bx + bw && y > by && y < by + bh) { in_box = true; break; }
            }
        const double thr = in_box ? 0.2 : 1.0;
        bool ok = true; double dist = 0;
        if (F) {
            const double a = x * F[0] + y * F[1] + F[2];
            const double b = x * F[3] + y * F[4] + F[5];
            const double c = x * F[6] + y * F[7] + F[8];
            const double son = std::fabs(a * prev_xy[2 * i] + b * prev_xy[2 * i + 1] + c);
            const double mom = std::sqrt(a * a + b * b);
            dist = son / mom;
            ok = dist < thr;  // NaN compares false -> removed, as in the reference
        }
        keep[i] = ok ? 1 : 0;
        if (dist_out) dist_out[i] = dist;
        if (!ok) sum--;
    }
    const bool restore = have_dyn && sum < nfeatures * 0.1;
    if (restored) *restored = restore ? 1 : 0;
    return sum;
}

// ---------------------------------------------------------------------------------------
// cv::calcOpticalFlowPyrLK(prevImg = I, nextImg = J, prevPts, nextPts, status, err, winSize (21,21), maxLevel 3,
// TermCriteria(COUNT|EPS, 30, 0.01)) as called at src/Frame.cc:445 (I = current gray, J = previous gray, prevPts = current
// keypoints, nextPts = "Prepoint").  Restates OpenCV video/lkpyramid.cpp (buildOpticalFlowPyramid with REFLECT_101 borders,
// cv::pyrDown 5x5, calcScharrDeriv, LKTrackerInvoker scalar path).  Float accumulations run in natural scalar order; OpenCV's
// SIMD paths sum in another order, so agreement with cv2 is to ~1e-4 px on converged tracks (tests/test_lk.py), not bit-exact.
// status/err are not produced: the reference ignores them (quirk Q4); a failed level leaves the running estimate untouched (A10).
// ---------------------------------------------------------------------------------------
namespace {

inline int refl(int i, int n) { return reflect101(i, n); }

struct LkLevel { int w, h; std::vector<uint8_t> px; inline int at(int x, int y) const { return px[(size_t)refl(y, h) * w + refl(x, w)]; } };

// cv::pyrDown for CV_8U (imgproc/pyramids.cpp): separable [1 4 6 4 1], exact integer sums, (v + 128) >> 8, BORDER_REFLECT_101
void pyr_down(const LkLevel& s, LkLevel& d) {
    d.w = (s.w + 1) / 2; d.h = (s.h + 1) / 2; d.px.resize((size_t)d.w * d.h);
    std::vector<int> row((size_t)s.h * d.w);
    for (int y = 0; y < s.h; y++)
        for (int x = 0; x < d.w; x++) {
            const uint8_t* r = s.px.data() + (size_t)y * s.w;
            row[(size_t)y * d.w + x] = r[refl(2 * x, s.w)] * 6 + (r[refl(2 * x - 1, s.w)] + r[refl(2 * x + 1, s.w)]) * 4 + r[refl(2 * x - 2, s.w)] + r[refl(2 * x + 2, s.w)];
        }
    for (int y = 0; y < d.h; y++)
        for (int x = 0; x < d.w; x++) {
            auto R = [&](int yy) { return row[(size_t)refl(yy, s.h) * d.w + x]; };
            const int v = R(2 * y) * 6 + (R(2 * y - 1) + R(2 * y + 1)) * 4 + R(2 * y - 2) + R(2 * y + 2);
            d.px[(size_t)y * d.w + x] = (uint8_t)((v + 128) >> 8);
        }
}

// calcScharrDeriv at one pixel inside the image (reflect-101 at the image edge); zero outside (BORDER_CONSTANT copyMakeBorder)
inline void scharr_at(const LkLevel& L, int x, int y, int& dx, int& dy) {
    if (x < 0 || x >= L.w || y < 0 || y >= L.h) { dx = 0; dy = 0; return; }
    auto t0 = [&](int xx) { return (L.at(xx, y - 1) + L.at(xx, y + 1)) * 3 + L.at(xx, y) * 10; };
    auto t1 = [&](int xx) { return L.at(xx, y + 1) - L.at(xx, y - 1); };
    // the horizontal border of the temporary rows is reflect-101 as well: trow[-1] = trow[1], trow[cols] = trow[cols-2]
    const int xm = refl(x - 1, L.w), xp = refl(x + 1, L.w);
    dx = (int16_t)(t0(xp) - t0(xm));
    dy = (int16_t)((t1(xm) + t1(xp)) * 3 + t1(x) * 10);
}

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// What buildOpticalFlowPyramid(withDerivatives) hands the tracker: the level with a REFLECT_101 border and its Scharr derivative
// image with a zero border, so the window loops below read straight through the image edge like OpenCV's do.
struct LkPadded {
    static constexpr int PAD = 23;            // the window reaches 21 px before and 22 px past the image
    int w = 0, h = 0, stride = 0;
    std::vector<uint8_t> px; std::vector<int16_t> dx, dy;
    inline const uint8_t* at(int x, int y) const { return px.data() + (size_t)(y + PAD) * stride + (x + PAD); }
    inline size_t idx(int x, int y) const { return (size_t)(y + PAD) * stride + (x + PAD); }
    void build(const LkLevel& L) {
        w = L.w; h = L.h; stride = w + 2 * PAD;
        px.assign((size_t)stride * (h + 2 * PAD), 0); dx.assign(px.size(), 0); dy.assign(px.size(), 0);
        for (int y = -PAD; y < h + PAD; y++)
            for (int x = -PAD; x < w + PAD; x++) px[idx(x, y)] = (uint8_t)L.at(x, y);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) { int gx, gy; scharr_at(L, x, y, gx, gy); dx[idx(x, y)] = (int16_t)gx; dy[idx(x, y)] = (int16_t)gy; }
    }
};

}  // namespace

// pts / out: n x 2 float.  I, J: 8-bit gray images of the same size.  Returns 0.
SGO_API int sgo_lk_track(const uint8_t* I0, const uint8_t* J0, int w, int h, int pitch, const float* pts, int n, float* out) {
    const int WIN = 21, MAXLEVEL = 3, MAXCOUNT = 30;
    const float MIN_EIG = 1e-4f;
    std::vector<LkLevel> I(1), J(1);
    I[0].w = J[0].w = w; I[0].h = J[0].h = h; I[0].px.resize((size_t)w * h); J[0].px.resize((size_t)w * h);
    for (int y = 0; y < h; y++) { std::memcpy(&I[0].px[(size_t)y * w], I0 + (size_t)y * pitch, w); std::memcpy(&J[0].px[(size_t)y * w], J0 + (size_t)y * pitch, w); }
    int maxLevel = 0;
    for (int l = 1; l <= MAXLEVEL; l++) {   // buildOpticalFlowPyramid stops when the next level is not larger than the window
        const int nw = (I[l - 1].w + 1) / 2, nh = (I[l - 1].h + 1) / 2;
        if (nw <= WIN || nh <= WIN) break;
        I.emplace_back(); J.emplace_back();
        pyr_down(I[l - 1], I[l]); pyr_down(J[l - 1], J[l]);
        maxLevel = l;
    }
    std::vector<LkPadded> PI(maxLevel + 1), PJ(maxLevel + 1);
    for (int l = 0; l <= maxLevel; l++) { PI[l].build(I[l]); PJ[l].build(J[l]); }
    std::vector<float> nx(n), ny(n);
    std::vector<short> Iw(WIN * WIN), dIx(WIN * WIN), dIy(WIN * WIN);
    const float halfWin = (WIN - 1) * 0.5f;
    for (int level = maxLevel; level >= 0; level--) {
        const LkPadded& LI = PI[level]; const LkPadded& LJ = PJ[level];
        for (int p = 0; p < n; p++) {
            float px = pts[2 * p] * (float)(1. / (1 << level)), py = pts[2 * p + 1] * (float)(1. / (1 << level));
            float qx, qy;
            if (level == maxLevel) { qx = px; qy = py; } else { qx = nx[p] * 2.f; qy = ny[p] * 2.f; }
            nx[p] = qx; ny[p] = qy;
            px -= halfWin; py -= halfWin;
            const int ipx = (int)std::floor(px), ipy = (int)std::floor(py);
            if (ipx < -WIN || ipx >= LI.w || ipy < -WIN || ipy >= LI.h) continue;
            float a = px - ipx, b = py - ipy;
            const int W_BITS = 14;
            const float FLT_SCALE = 1.f / (1 << 20);
            int iw00 = cvRoundF((1.f - a) * (1.f - b) * (1 << W_BITS)), iw01 = cvRoundF(a * (1.f - b) * (1 << W_BITS));
            int iw10 = cvRoundF((1.f - a) * b * (1 << W_BITS)), iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            float iA11 = 0, iA12 = 0, iA22 = 0;
            const int st = LI.stride;
            for (int y = 0; y < WIN; y++) {
                const size_t o = LI.idx(ipx, ipy + y);
                const uint8_t* src = LI.px.data() + o; const int16_t* gx = LI.dx.data() + o; const int16_t* gy = LI.dy.data() + o;
                for (int x = 0; x < WIN; x++) {
                    const int ival = descale(src[x] * iw00 + src[x + 1] * iw01 + src[x + st] * iw10 + src[x + st + 1] * iw11, W_BITS - 5);
                    const int ixval = descale(gx[x] * iw00 + gx[x + 1] * iw01 + gx[x + st] * iw10 + gx[x + st + 1] * iw11, W_BITS);
                    const int iyval = descale(gy[x] * iw00 + gy[x + 1] * iw01 + gy[x + st] * iw10 + gy[x + st + 1] * iw11, W_BITS);
                    Iw[y * WIN + x] = (short)ival; dIx[y * WIN + x] = (short)ixval; dIy[y * WIN + x] = (short)iyval;
                    iA11 += (float)(ixval * ixval); iA12 += (float)(ixval * iyval); iA22 += (float)(iyval * iyval);
                }
            }
            const float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * WIN * WIN);
            if (minEig < MIN_EIG || D < FLT_EPSILON) continue;
            D = 1.f / D;
            qx -= halfWin; qy -= halfWin;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < MAXCOUNT; j++) {
                const int inx = (int)std::floor(qx), iny = (int)std::floor(qy);
                if (inx < -WIN || inx >= LJ.w || iny < -WIN || iny >= LJ.h) break;
                a = qx - inx; b = qy - iny;
                iw00 = cvRoundF((1.f - a) * (1.f - b) * (1 << W_BITS)); iw01 = cvRoundF(a * (1.f - b) * (1 << W_BITS));
                iw10 = cvRoundF((1.f - a) * b * (1 << W_BITS)); iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                float ib1 = 0, ib2 = 0;
                const int sj = LJ.stride;
                for (int y = 0; y < WIN; y++) {
                    const uint8_t* src = LJ.at(inx, iny + y);
                    for (int x = 0; x < WIN; x++) {
                        const int diff = descale(src[x] * iw00 + src[x + 1] * iw01 + src[x + sj] * iw10 + src[x + sj + 1] * iw11, W_BITS - 5) - Iw[y * WIN + x];
                        ib1 += (float)(diff * dIx[y * WIN + x]); ib2 += (float)(diff * dIy[y * WIN + x]);
                    }
                }
                const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
                const float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
                qx += dx; qy += dy;
                nx[p] = qx + halfWin; ny[p] = qy + halfWin;
                if ((double)dx * dx + (double)dy * dy <= 0.01 * 0.01) break;                 // Point2f::ddot is double; epsilon is squared in double
                if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) { nx[p] -= dx * 0.5f; ny[p] -= dy * 0.5f; break; }
                pdx = dx; pdy = dy;
            }
        }
    }
    for (int p = 0; p < n; p++) { out[2 * p] = nx[p]; out[2 * p + 1] = ny[p]; }
    return 0;
}

// LK pyramid level (for the GPU parity gate): returns level `level` of cv::buildOpticalFlowPyramid without its border
SGO_API int sgo_lk_pyr_level(const uint8_t* img, int w, int h, int pitch, int level, uint8_t* out, int32_t* ow, int32_t* oh) {
    LkLevel cur; cur.w = w; cur.h = h; cur.px.resize((size_t)w * h);
    for (int y = 0; y < h; y++) std::memcpy(&cur.px[(size_t)y * w], img + (size_t)y * pitch, w);
    for (int l = 0; l < level; l++) { LkLevel nxt; pyr_down(cur, nxt); cur = std::move(nxt); }
    *ow = cur.w; *oh = cur.h;
    if (out) std::memcpy(out, cur.px.data(), cur.px.size());
    return 0;
}

SGO_API int sgo_abi_version(void) { return 1; }

// C entry point around the reference's OWN ORBextractor (src/ORBextractor.cc compiled from the reference tree against the stand-ins of orbmatcher_shim/; the
// OpenCV algorithms it calls -- resize, FAST, GaussianBlur, fastAtan2 -- are the oracle's cv2-pinned restatements): image in, the reference's keypoints and
// descriptors out.  TEST INFRASTRUCTURE (oracle/_ref/liborbextractor_ref.so).
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBextractor.h"

struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };
void ref_arena_restart();      // below: no object of an earlier call is alive when a new one starts

extern "C" __attribute__((visibility("default")))
int ref_orb_extract(const uint8_t* img, int w, int h, int pitch, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, Kp* kps, uint8_t* desc, int cap) {
    ref_arena_restart();
    cv::Mat im(h, w, CV_8U);
    for (int y = 0; y < h; ++y) std::memcpy(im.ptr(y), img + (size_t)y * pitch, (size_t)w);
    ORB_SLAM2::ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
    std::vector<cv::KeyPoint> k; cv::Mat d;
    ex(im, cv::Mat(), k, d);
    const int n = (int)k.size();
    for (int i = 0; i < n && i < cap; ++i) {
        kps[i] = Kp{k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id};
        std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    }
    return n;
}

// The reference's DistributeOctTree alone (protected member: reached through a derived class).  cands: [n][3] = x, y, response; returns the kept keypoints.
namespace {
struct Open : ORB_SLAM2::ORBextractor {
    using ORB_SLAM2::ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> tree(const std::vector<cv::KeyPoint>& v, int minX, int maxX, int minY, int maxY, int N, int level) { return DistributeOctTree(v, minX, maxX, minY, maxY, N, level); }
};
}  // namespace
extern "C" __attribute__((visibility("default")))
int ref_octree(const float* cands, int n, int minX, int maxX, int minY, int maxY, int N, float* out3, int cap) {
    ref_arena_restart();
    Open ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> v(n);
    for (int i = 0; i < n; ++i) { v[i].pt = cv::Point2f(cands[3 * i], cands[3 * i + 1]); v[i].response = cands[3 * i + 2]; v[i].size = 7.f; }
    const std::vector<cv::KeyPoint> r = ex.tree(v, minX, maxX, minY, maxY, N, 0);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) { out3[3 * i] = r[i].pt.x; out3[3 * i + 1] = r[i].pt.y; out3[3 * i + 2] = r[i].response; }
    return (int)r.size();
}

// ---- allocation order = address order ----------------------------------------------------------------------------------------------------------------------------
// DistributeOctTree sorts (size, ExtractorNode*) pairs (src/ORBextractor.cc:684): nodes of equal size are ordered by their ADDRESS, i.e. by what the allocator
// happened to hand out -- with glibc's malloc (freed list nodes are reused last-in-first-out) that order depends on the history of the heap, so the reference's
// own output is not a function of its input there.  The oracle fixes the tie-break as "creation sequence" (quirk Q1).  To compare like with like
// this library allocates from a bump arena that never reuses memory: addresses grow with creation order.  The operators are hidden (only code inlined into this
// library -- every container of ORBextractor.cc -- uses them).
#include <cstdlib>
#include <new>
namespace {
struct Arena {                      // ONE block (virtual memory, committed on touch), restarted at every extraction: addresses grow with creation order throughout a call
    char* base = nullptr; size_t cap = (size_t)1 << 30, off = 0;
    void* get(size_t n) {
        n = (n + 15) & ~(size_t)15;
        if (!base) base = (char*)std::malloc(cap);
        if (!base || off + n > cap) return std::malloc(n);
        void* p = base + off; off += n; return p;
    }
    bool owns(void* p) const { return base && (char*)p >= base && (char*)p < base + cap; }
};
Arena g_arena;
bool g_monotone = false;
}  // namespace
__attribute__((visibility("hidden"))) void* operator new(size_t n) { if (g_monotone) return g_arena.get(n); void* p = std::malloc(n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
__attribute__((visibility("hidden"))) void operator delete(void* p) noexcept { if (!p || g_arena.owns(p)) return; std::free(p); }
__attribute__((visibility("hidden"))) void operator delete(void* p, size_t) noexcept { ::operator delete(p); }
void ref_arena_restart() { g_arena.off = 0; }
extern "C" __attribute__((visibility("default"))) void ref_set_monotone_allocator(int on) { g_monotone = on != 0; }

// chain.cpp -- the reference's per-frame tracking chain on the CPU, for a batch of frames on a pool of pinned worker threads:
//   ORBextractor::operator()  ->  calcOpticalFlowPyrLK to the previous frame  ->  findFundamentalMat (pair selection + RANSAC / LMedS)
//   ->  dynamic-feature rejection + ordered compaction  ->  ORBmatcher::SearchByProjection(cur, last)
// (src/Frame.cc:129-198, 430-612; src/Tracking.cc:906-930).  Each frame is one task; a frame is processed by exactly one thread, like the
// reference's single tracking thread, and `nthreads` frames run side by side.  This is the CPU arm of bench.py (`--impl reference`,
// `cpu_baseline`) and the pure-oracle side of the full-chain parity check.  TEST / BENCH INFRASTRUCTURE, never linked into the product.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <malloc.h>
#include <pthread.h>
#include <sched.h>

#define SGO_API extern "C" __attribute__((visibility("default")))

struct KeyPoint { float x, y, size, angle, response; int32_t octave, class_id; };
struct OrbParams { int32_t nfeatures; float scaleFactor; int32_t nlevels; int32_t iniThFAST; int32_t minThFAST; };
struct SgoFrame {
    int32_t N; const KeyPoint* keysUn; const float* uRight; const uint8_t* desc;
    float minX, minY, maxX, maxY; float fx, fy, cx, cy, bf; int32_t nlevels; const float* scaleFactors; float logScaleFactor;
};
extern "C" {
int sgo_extract(const OrbParams* p, const uint8_t* img, int w, int h, int pitch, KeyPoint* kps, uint8_t* desc, int cap);
int sgo_lk_track(const uint8_t* I0, const uint8_t* J0, int w, int h, int pitch, const float* pts, int n, float* out);
int sgo_select_static_pairs(const float* cur, const float* prev, int n, const float* boxes, int nboxes, int prev_have_dyn, float* sel1, float* sel2);
int sgo_find_fundamental_ransac(const float* m1, const float* m2, int n, double thresh, double confidence, int max_iters, double* F, uint8_t* mask_out, int32_t* info);
int sgo_dynreject(const float* cur_xy, const float* prev_xy, int n, const double* F, const float* boxes, int nboxes, int have_dyn, int nfeatures, uint8_t* keep,
                  double* dist_out, int32_t* restored);
int sgo_search_by_projection_last(const SgoFrame* cur, const float* Tcw_cur, const float* Tcw_last, int nlast, const uint8_t* last_has_mp, const float* last_xyz,
                                  const uint8_t* last_desc, const uint8_t* last_obs, const int32_t* last_octave, const float* last_angle, float th, int bMono,
                                  int checkOri, int32_t* cur_mp_inout, const uint8_t* cur_mp_obs_in, int64_t* ncand_out);
}

struct SgoChainArgs {
    const OrbParams* params;
    const uint8_t* frames;            // [nframes][h][w] 8-bit gray
    int32_t nframes, w, h;
    const int32_t* prev_index;        // row of the previous frame of the same stream; == f: no previous frame (nothing is rejected)
    const float* boxes;               // [nframes][max_boxes][4] person boxes of every frame (x, y, w, h)
    const int32_t* nboxes;            // [nframes]
    const uint8_t* have_dyn;          // [nframes]
    int32_t max_boxes;
    const float* depth;               // [h][w] metres, shared by all frames (Frame::ComputeStereoFromRGBD, src/Frame.cc:893-914), or NULL with u_right given
    const float* u_right_in;          // [nframes][cap] when depth == NULL
    float fx, fy, cx, cy, bf;
    const float* scale_factors;       // [nlevels]
    float log_scale_factor;
    // last-frame map points of every frame (Tracking::UpdateLastFrame state): [nframes][point_cap] ...
    int32_t point_cap;
    const float* last_xyz; const uint8_t* last_desc; const uint8_t* last_flags; const int32_t* last_octave; const float* last_angle; const int32_t* last_n;
    const float* tcw;                 // [nframes][16]
    float th;
    int32_t cap;                      // rows per frame of the outputs
    // outputs (any may be NULL)
    KeyPoint* kps; uint8_t* desc; int32_t* counts;            // extractor output
    float* prev_xy;                                           // LK points [nframes][cap][2]
    double* F; int32_t* f_ok;                                 // [nframes][9], [nframes]
    uint8_t* keep; int32_t* nkeep; int32_t* restored;         // per-keypoint verdicts, survivors (after the restore-all guard), guard fired
    int32_t* match; int32_t* nmatch;                          // [nframes][cap] index of the last-frame point per SURVIVING keypoint (compacted order), count
};

static void chain_one(const SgoChainArgs& A, int f) {
    const int cap = A.cap;
    std::vector<KeyPoint> k(cap); std::vector<uint8_t> d((size_t)cap * 32);
    const uint8_t* img = A.frames + (size_t)f * A.w * A.h;
    const int n = sgo_extract(A.params, img, A.w, A.h, A.w, k.data(), d.data(), cap);
    std::vector<float> cur(2 * (size_t)n), prev(2 * (size_t)n);
    for (int i = 0; i < n; i++) { cur[2 * i] = k[i].x; cur[2 * i + 1] = k[i].y; }
    const int g = A.prev_index[f];
    sgo_lk_track(img, A.frames + (size_t)g * A.w * A.h, A.w, A.h, A.w, cur.data(), n, prev.data());
    double Fm[9]; int have_F = 0;
    if (g != f) {
        std::vector<float> s1(2 * (size_t)n), s2(2 * (size_t)n);
        // the previous-frame flag is file-scope state written only inside the rejection (src/Frame.cc:482-491): a stream's first frame (its own predecessor here)
        // never sets it, so its detections do not filter the pairs of the second frame (quirk Q13; tests/test_frame_ref.py runs the reference's own Frame.cc)
        const int pre_have = A.have_dyn[g] && A.prev_index[g] != g;
        const int ns = sgo_select_static_pairs(cur.data(), prev.data(), n, A.boxes + (size_t)g * A.max_boxes * 4, A.nboxes[g], pre_have, s1.data(), s2.data());
        have_F = sgo_find_fundamental_ransac(s1.data(), s2.data(), ns, 1.0, 0.99, 1000, Fm, nullptr, nullptr);
    }
    std::vector<uint8_t> keep(n > 0 ? n : 1);
    int32_t restored = 0;
    int nk = sgo_dynreject(cur.data(), prev.data(), n, have_F ? Fm : nullptr, A.boxes + (size_t)f * A.max_boxes * 4, A.nboxes[f], A.have_dyn[f], A.params->nfeatures,
                           keep.data(), nullptr, &restored);
    // ordered compaction (src/Frame.cc:560-604) + u_right of the survivors
    std::vector<KeyPoint> ks; std::vector<uint8_t> ds; std::vector<float> ur;
    ks.reserve(n); ds.reserve((size_t)n * 32); ur.reserve(n);
    for (int i = 0; i < n; i++) {
        if (!restored && !keep[i]) continue;
        ks.push_back(k[i]); ds.insert(ds.end(), d.begin() + (size_t)i * 32, d.begin() + (size_t)(i + 1) * 32);
        float u = -1.f;
        if (A.depth) {                                                     // ComputeStereoFromRGBD: depth at the (distorted) pixel, k1 = 0 here
            const float dz = A.depth[(size_t)(int)k[i].y * A.w + (int)k[i].x];
            if (dz > 0) u = k[i].x - A.bf / dz;
        } else u = A.u_right_in[(size_t)f * cap + i];
        ur.push_back(u);
    }
    if (restored) nk = n;
    SgoFrame fr;
    fr.N = (int)ks.size(); fr.keysUn = ks.data(); fr.uRight = ur.data(); fr.desc = ds.data();
    fr.minX = 0; fr.minY = 0; fr.maxX = (float)A.w; fr.maxY = (float)A.h; fr.fx = A.fx; fr.fy = A.fy; fr.cx = A.cx; fr.cy = A.cy; fr.bf = A.bf;
    fr.nlevels = A.params->nlevels; fr.scaleFactors = A.scale_factors; fr.logScaleFactor = A.log_scale_factor;
    const int m = A.last_n[f];
    std::vector<uint8_t> has(m > 0 ? m : 1), obs(m > 0 ? m : 1);
    const uint8_t* fl = A.last_flags + (size_t)f * A.point_cap;
    for (int i = 0; i < m; i++) { has[i] = fl[i] & 1; obs[i] = (fl[i] >> 1) & 1; }
    std::vector<int32_t> mp(fr.N > 0 ? fr.N : 1, -1);
    int64_t ncand = 0;
    const float* T = A.tcw + (size_t)f * 16;
    const int nm = sgo_search_by_projection_last(&fr, T, T, m, has.data(), A.last_xyz + (size_t)f * A.point_cap * 3, A.last_desc + (size_t)f * A.point_cap * 32, obs.data(),
                                                 A.last_octave + (size_t)f * A.point_cap, A.last_angle + (size_t)f * A.point_cap, A.th, 0, 1, mp.data(), nullptr, &ncand);
    if (A.counts) A.counts[f] = n;
    if (A.kps) std::memcpy(A.kps + (size_t)f * cap, k.data(), sizeof(KeyPoint) * (size_t)n);
    if (A.desc) std::memcpy(A.desc + (size_t)f * cap * 32, d.data(), (size_t)n * 32);
    if (A.prev_xy) std::memcpy(A.prev_xy + (size_t)f * cap * 2, prev.data(), sizeof(float) * 2 * (size_t)n);
    if (A.F) { if (have_F) std::memcpy(A.F + (size_t)f * 9, Fm, sizeof(Fm)); else for (int i = 0; i < 9; i++) A.F[(size_t)f * 9 + i] = 0.0 / 0.0; }
    if (A.f_ok) A.f_ok[f] = have_F;
    if (A.keep) std::memcpy(A.keep + (size_t)f * cap, keep.data(), (size_t)n);
    if (A.nkeep) A.nkeep[f] = nk;
    if (A.restored) A.restored[f] = restored;
    if (A.match) std::memcpy(A.match + (size_t)f * cap, mp.data(), sizeof(int32_t) * (size_t)fr.N);
    if (A.nmatch) A.nmatch[f] = nm;
}

// Runs the chain over frames [first, first + count) on `nthreads` worker threads; thread t is pinned to the t-th CPU of the process's
// affinity mask (pin != 0).  Returns the number of worker threads used.
SGO_API int sgo_chain_batch(const SgoChainArgs* a, int first, int count, int nthreads, int pin) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > count) nthreads = count > 0 ? count : 1;
    // the per-frame buffers (pyramids, LK planes: MBs each) would otherwise be mmap'ed and unmapped on every frame: with one frame per core in
    // flight that serialises the workers on the kernel's address-space lock.  Keep them in the malloc arenas instead (process-wide, set once).
    static bool tuned = false;
    if (!tuned) { mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_TOP_PAD, 64 << 20); tuned = true; }
    std::vector<int> cpus;
    cpu_set_t mask;
    if (pin && sched_getaffinity(0, sizeof(mask), &mask) == 0)
        for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &mask)) cpus.push_back(c);
    std::atomic<int> next(0);
    auto work = [&](int t) {
        if (pin && !cpus.empty()) {
            cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[t % cpus.size()], &one);
            pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
        }
        for (;;) { const int i = next.fetch_add(1); if (i >= count) break; chain_one(*a, first + i); }
    };
    std::vector<std::thread> th;                              // always worker threads: the caller's own affinity is never touched
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    return nthreads;
}

// CPUs this process may really use: the affinity mask, capped by the cgroup CPU quota when there is one (cpu.max "quota period": containers on
// shared hosts often see every core of the machine but are throttled to a few cores' worth of time; running more busy threads than that only
// adds throttling stalls).
SGO_API int sgo_online_cpus(void) {
    cpu_set_t mask;
    int n = sched_getaffinity(0, sizeof(mask), &mask) == 0 ? CPU_COUNT(&mask) : (int)std::thread::hardware_concurrency();
    if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[64] = {0}; long period = 0;
        if (std::fscanf(f, "%63s %ld", q, &period) == 2 && q[0] != 'm' && period > 0) {
            const long quota = std::atol(q);
            const int cap = (int)((quota + period - 1) / period);
            if (cap >= 1 && cap < n) n = cap;
        }
        std::fclose(f);
    }
    return n > 0 ? n : 1;
}

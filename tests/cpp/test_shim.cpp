// test_shim.cpp -- compiles the reference-shaped C++ headers (include/sgslam/*.h) without OpenCV and drives them on the GPU.
// Scenario + expected results come from files written by tests/test_gpu_cpp_shim.py (expected = CPU oracle).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "sgslam/FrameDynamic.h"
#include "sgslam/FrameGeometry.h"
#include "sgslam/Optimizer.h"
#include "sgslam/Detector2D.h"
#include <map>
#include <set>

#include "sgslam/ORBextractor.h"
#include "sgslam/ORBmatcher.h"

using namespace ORB_SLAM2;

// ---- structs with the member names of the reference's MapPoint / Frame (include/MapPoint.h, include/Frame.h) -------------
struct MapPoint {
    cv::Mat pos, desc; int nobs = 1; bool bad = false;
    bool mbTrackInView = false; float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1; int mnTrackScaleLevel = 0;
    cv::Mat GetWorldPos() { return pos; }
    cv::Mat GetDescriptor() { return desc; }
    cv::Mat normal; float mfMinDistance = 0, mfMaxDistance = 0;
    cv::Mat GetNormal() { return normal; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int Observations() { return nobs; }
    bool isBad() { return bad; }
    // mapping-side surface used by Fuse / SearchByProjection(KeyFrame*, Scw, ...): observations as an opaque key-frame set, Replace marks the point bad
    std::set<const void*> obs; MapPoint* replaced_by = nullptr;
    template <class KF> bool IsInKeyFrame(KF* kf) { return obs.count(kf) != 0; }
    template <class KF> void AddObservation(KF* kf, size_t) { if (obs.insert(kf).second) ++nobs; }
    void Replace(MapPoint* other) { bad = true; replaced_by = other; }
    std::map<const void*, int> index_in_kf;
    template <class KF> int GetIndexInKeyFrame(KF* kf) { auto it = index_in_kf.find(kf); return it == index_in_kf.end() ? -1 : it->second; }
};
struct KeyFrame {
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors, mvInvLevelSigma2, mvLevelSigma2;
    std::map<unsigned, std::vector<unsigned> > mFeatVec;          // DBoW2::FeatureVector
    cv::Mat mDescriptors, R, t, O;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mnMinX = 0, mnMinY = 0, mnMaxX = 640, mnMaxY = 480;
    std::vector<MapPoint*> mps;
    cv::Mat GetRotation() { return R; }
    cv::Mat GetTranslation() { return t; }
    cv::Mat GetCameraCenter() { return O; }
    MapPoint* GetMapPoint(size_t idx) { return mps[idx]; }
    std::vector<MapPoint*> GetMapPointMatches() { return mps; }
    void AddMapPoint(MapPoint* p, size_t idx) { mps[idx] = p; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> r; for (MapPoint* p : mps) if (p && !p->isBad()) r.insert(p); return r; }
};
struct Frame {
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors, mTcw;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors, mvInvLevelSigma2;
    float mbf = 40.f;
    void SetPose(const cv::Mat& T) { mTcw = T.clone(); }
    static float fx, fy, cx, cy, mnMinX, mnMinY, mnMaxX, mnMaxY;
};
float Frame::fx = 535.4f, Frame::fy = 539.2f, Frame::cx = 320.1f, Frame::cy = 247.6f, Frame::mnMinX = 0, Frame::mnMinY = 0, Frame::mnMaxX = 640, Frame::mnMaxY = 480;

template <class T> std::vector<T> rd(std::ifstream& f, size_t n) { std::vector<T> v(n); f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T)); return v; }
static int fail(const char* what) { std::printf("FAIL %s\n", what); return 1; }

int main(int argc, char** argv) {
    if (argc < 2) return fail("usage: test_shim <scenario.bin>");
    std::ifstream f(argv[1], std::ios::binary);
    int32_t hdr[8]; f.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    const int W = hdr[0], H = hdr[1], nkp = hdr[2], ncur = hdr[3], nlast = hdr[4], exp_nm = hdr[5], exp_keep = hdr[6];
    // 1. extractor
    std::vector<uint8_t> img = rd<uint8_t>(f, (size_t)W * H);
    std::vector<cv::KeyPoint> exp_kps = rd<cv::KeyPoint>(f, nkp);
    std::vector<uint8_t> exp_desc = rd<uint8_t>(f, (size_t)nkp * 32);
    ORBextractor ex(1000, 1.2f, 8, 20, 7);
    cv::Mat im(H, W, CV_8UC1, img.data(), W), d;
    std::vector<cv::KeyPoint> kps;
    ex(im, cv::Mat(), kps, d);
    if ((int)kps.size() != nkp) return fail("extractor: keypoint count");
    if (std::memcmp(kps.data(), exp_kps.data(), (size_t)nkp * sizeof(cv::KeyPoint))) return fail("extractor: keypoints");
    for (int i = 0; i < nkp; ++i) if (std::memcmp(d.ptr<uint8_t>(i), &exp_desc[(size_t)i * 32], 32)) return fail("extractor: descriptors");
    if (ex.GetLevels() != 8 || ex.GetScaleFactors().size() != 8) return fail("extractor: getters");
    // 2. SearchByProjection(cur, last)
    Frame cur, last;
    cur.N = ncur; last.N = nlast;
    cur.mvKeysUn = rd<cv::KeyPoint>(f, ncur); cur.mvKeys = cur.mvKeysUn;
    cur.mvuRight = rd<float>(f, ncur);
    std::vector<uint8_t> cd = rd<uint8_t>(f, (size_t)ncur * 32);
    cur.mDescriptors = cv::Mat(ncur, 32, CV_8U, cd.data(), 32);
    cur.mvScaleFactors = rd<float>(f, 8); last.mvScaleFactors = cur.mvScaleFactors;
    std::vector<float> tc = rd<float>(f, 16), tl = rd<float>(f, 16);
    cur.mTcw = cv::Mat(4, 4, CV_32F, tc.data(), 16); last.mTcw = cv::Mat(4, 4, CV_32F, tl.data(), 16);
    cur.mvpMapPoints.assign(ncur, nullptr);
    std::vector<uint8_t> has = rd<uint8_t>(f, nlast), obs = rd<uint8_t>(f, nlast);
    std::vector<float> xyz = rd<float>(f, (size_t)nlast * 3);
    std::vector<uint8_t> ld = rd<uint8_t>(f, (size_t)nlast * 32);
    std::vector<int32_t> oct = rd<int32_t>(f, nlast);
    std::vector<float> ang = rd<float>(f, nlast);
    std::vector<int32_t> exp_mp = rd<int32_t>(f, ncur);
    std::vector<MapPoint> pts(nlast);
    last.mvKeys.resize(nlast); last.mvKeysUn.resize(nlast); last.mvpMapPoints.assign(nlast, nullptr); last.mvbOutlier.assign(nlast, false);
    for (int i = 0; i < nlast; ++i) {
        pts[i].pos = cv::Mat(3, 1, CV_32F, &xyz[(size_t)i * 3], 4); pts[i].desc = cv::Mat(1, 32, CV_8U, &ld[(size_t)i * 32], 32); pts[i].nobs = obs[i];
        last.mvKeys[i].octave = oct[i]; last.mvKeysUn[i].angle = ang[i];
        if (has[i]) last.mvpMapPoints[i] = &pts[i];
    }
    ORBmatcher matcher(0.9f, true);
    const int nm = matcher.SearchByProjection(cur, last, 15.f, false);
    if (nm != exp_nm) { std::printf("nmatches %d expected %d\n", nm, exp_nm); return fail("SearchByProjection: count"); }
    for (int j = 0; j < ncur; ++j) {
        const int got = cur.mvpMapPoints[j] ? (int)(cur.mvpMapPoints[j] - pts.data()) : -1;
        if (got != exp_mp[j]) return fail("SearchByProjection: assignment");
    }
    // 3. dyn-reject compaction
    const int nd = hdr[7];
    std::vector<cv::KeyPoint> dk = rd<cv::KeyPoint>(f, nd);
    std::vector<uint8_t> dd = rd<uint8_t>(f, (size_t)nd * 32);
    std::vector<cv::Point2f> prev = rd<cv::Point2f>(f, nd);
    std::vector<double> Fm = rd<double>(f, 9);
    std::vector<float> bx = rd<float>(f, 8);
    std::vector<uint8_t> exp_keepmask = rd<uint8_t>(f, nd);
    cv::Mat ddm(nd, 32, CV_8U, dd.data(), 32); cv::Mat ddc = ddm.clone();
    cv::Mat Fmat(3, 3, CV_64F, Fm.data(), 24);
    std::vector<cv::Rect_<float> > boxes = {cv::Rect_<float>(bx[0], bx[1], bx[2], bx[3]), cv::Rect_<float>(bx[4], bx[5], bx[6], bx[7])};
    std::vector<cv::KeyPoint> dk2 = dk;
    const int kept = RmDynamicPointsGeometry(dk2, ddc, prev, Fmat, boxes, true, 1000);
    if (kept != exp_keep || (int)dk2.size() != exp_keep) return fail("dynreject: count");
    int w = 0;
    for (int i = 0; i < nd; ++i) if (exp_keepmask[i]) { if (std::memcmp(&dk2[w], &dk[i], sizeof(cv::KeyPoint)) || std::memcmp(ddc.ptr<uint8_t>(w), &dd[(size_t)i * 32], 32)) return fail("dynreject: compaction"); ++w; }
    // 4. LK + findFundamentalMat shims (tolerances of tests/test_gpu_lk.py / test_gpu_fundamental.py)
    int32_t nlk = 0; f.read(reinterpret_cast<char*>(&nlk), 4);
    std::vector<uint8_t> im1 = rd<uint8_t>(f, (size_t)W * H), im0 = rd<uint8_t>(f, (size_t)W * H);
    std::vector<cv::Point2f> cpts = rd<cv::Point2f>(f, nlk), exp_trk = rd<cv::Point2f>(f, nlk);
    std::vector<double> expF = rd<double>(f, 9);
    std::vector<cv::Point2f> ppts;
    CalcOpticalFlowPyrLK(cv::Mat(H, W, CV_8UC1, im1.data(), W), cv::Mat(H, W, CV_8UC1, im0.data(), W), cpts, ppts);
    int far = 0;
    for (int i = 0; i < nlk; ++i) {
        const float e = std::max(std::fabs(ppts[i].x - exp_trk[i].x), std::fabs(ppts[i].y - exp_trk[i].y));
        if (e > 0.25f) return fail("LK: track");
        if (e > 0.02f) ++far;
    }
    if (far > std::max(1, nlk * 3 / 1000)) return fail("LK: too many loose tracks");
    cv::Mat Fg = FindFundamentalMatRansac(cpts, exp_trk);          // on the expected tracks, so F is comparable entry by entry
    if (Fg.empty()) return fail("findFundamentalMat: empty");
    double fmax = 1.0;
    for (int i = 0; i < 9; ++i) fmax = std::max(fmax, std::fabs(expF[i]));
    for (int i = 0; i < 9; ++i) if (std::fabs(Fg.at<double>(i / 3, i % 3) - expF[i]) > 1e-9 * fmax) return fail("findFundamentalMat: F");
    // below 15 pairs OpenCV's function takes its LMedS branch (8..14 pairs), below 7 it returns an empty matrix: both through the same entry point
    if (FindFundamentalMatRansac(std::vector<cv::Point2f>(cpts.begin(), cpts.begin() + 10), std::vector<cv::Point2f>(exp_trk.begin(), exp_trk.begin() + 10)).empty())
        return fail("findFundamentalMat: 10 pairs must take the LMedS branch and return a model");
    if (!FindFundamentalMatRansac(std::vector<cv::Point2f>(cpts.begin(), cpts.begin() + 6), std::vector<cv::Point2f>(exp_trk.begin(), exp_trk.begin() + 6)).empty())
        return fail("findFundamentalMat: fewer than 7 pairs must come back empty");
    // 5. isInFrustum for a batch of map points (expected values: CPU oracle)
    int32_t nfr = 0; f.read(reinterpret_cast<char*>(&nfr), 4);
    std::vector<float> fT = rd<float>(f, 16), fxyz = rd<float>(f, (size_t)nfr * 3), fnrm = rd<float>(f, (size_t)nfr * 3), fmn = rd<float>(f, nfr), fmx = rd<float>(f, nfr);
    std::vector<uint8_t> ein = rd<uint8_t>(f, nfr);
    std::vector<float> epx = rd<float>(f, nfr), epy = rd<float>(f, nfr), epxr = rd<float>(f, nfr), evc = rd<float>(f, nfr);
    std::vector<int32_t> elv = rd<int32_t>(f, nfr);
    Frame ff; ff.mTcw = cv::Mat(4, 4, CV_32F, fT.data(), 16).clone(); ff.mvScaleFactors = cur.mvScaleFactors;
    std::vector<MapPoint> store(nfr); std::vector<MapPoint*> ptrs(nfr);
    for (int i = 0; i < nfr; ++i) {
        store[i].pos = cv::Mat(3, 1, CV_32F, &fxyz[3 * (size_t)i], 4).clone(); store[i].normal = cv::Mat(3, 1, CV_32F, &fnrm[3 * (size_t)i], 4).clone();
        store[i].mfMinDistance = fmn[i]; store[i].mfMaxDistance = fmx[i]; ptrs[i] = &store[i];
    }
    const int nview = UpdateTrackInView(ff, ptrs, 0.5f);
    int expview = 0, lvldiff = 0;
    for (int i = 0; i < nfr; ++i) {
        expview += ein[i];
        if (store[i].mbTrackInView != (ein[i] != 0)) return fail("isInFrustum: flag");
        if (!ein[i]) continue;
        // the shim divides the invariance distances back by 0.8 / 1.2: compare positions exactly, distances-dependent gates by flag only
        if (store[i].mTrackProjX != epx[i] || store[i].mTrackProjY != epy[i] || store[i].mTrackProjXR != epxr[i] || store[i].mTrackViewCos != evc[i]) return fail("isInFrustum: projection");
        if (store[i].mnTrackScaleLevel != elv[i]) ++lvldiff;
    }
    if (nview != expview || lvldiff > 2) return fail("isInFrustum: count / level");
    // 6. PoseOptimization (expected: the CPU restatement)
    int32_t npo = 0; f.read(reinterpret_cast<char*>(&npo), 4);
    std::vector<float> pT0 = rd<float>(f, 16), pxyz = rd<float>(f, (size_t)npo * 3);
    std::vector<cv::KeyPoint> pk = rd<cv::KeyPoint>(f, npo);
    std::vector<float> pur = rd<float>(f, npo), pis2 = rd<float>(f, 8);
    std::vector<uint8_t> phas = rd<uint8_t>(f, npo), pexp_out = rd<uint8_t>(f, npo);
    std::vector<float> pexpT = rd<float>(f, 16);
    int32_t pexp_n = 0; f.read(reinterpret_cast<char*>(&pexp_n), 4);
    Frame pf; pf.N = npo; pf.mvKeysUn = pk; pf.mvuRight = pur; pf.mvScaleFactors = cur.mvScaleFactors; pf.mvInvLevelSigma2 = pis2;
    pf.mTcw = cv::Mat(4, 4, CV_32F, pT0.data(), 16).clone(); pf.mvbOutlier.assign(npo, false); pf.mvpMapPoints.assign(npo, nullptr);
    std::vector<MapPoint> pstore(npo);
    for (int i = 0; i < npo; ++i) if (phas[i]) { pstore[i].pos = cv::Mat(3, 1, CV_32F, &pxyz[3 * (size_t)i], 4).clone(); pf.mvpMapPoints[i] = &pstore[i]; }
    const int pin = PoseOptimizationGPU(&pf);
    if (pin != pexp_n) return fail("PoseOptimization: inlier count");
    for (int i = 0; i < npo; ++i) if (phas[i] && pf.mvbOutlier[i] != (pexp_out[i] != 0)) return fail("PoseOptimization: outlier flags");
    for (int i = 0; i < 16; ++i) if (std::fabs(pf.mTcw.at<float>(i / 4, i % 4) - pexpT[i]) > 1e-6f) return fail("PoseOptimization: pose");
    // 7. Detector2D::detect through the class mirror (model paths on the command line)
    int ndet = -1, npers = 0;
    if (argc >= 4) {
        int32_t dh[3]; f.read(reinterpret_cast<char*>(dh), sizeof dh);
        const int dW = dh[0], dH = dh[1], dn = dh[2];
        std::vector<uint8_t> rgb = rd<uint8_t>(f, (size_t)dW * dH * 3);
        std::vector<int32_t> eid = rd<int32_t>(f, dn);
        std::vector<float> erect = rd<float>(f, (size_t)dn * 5);          // prob, x, y, w, h
        struct Img { const uint8_t* data; int cols, rows; size_t step; } im3{rgb.data(), dW, dH, (size_t)dW * 3};
        DetectorGPU det(0.9f, 0.01f, argv[2], argv[3]);
        det.detect(im3);
        ndet = (int)det.mvObjects2D_to_View.size();
        if (ndet != dn) return fail("Detector2D: object count");
        size_t others = 0;
        for (int i = 0; i < dn; ++i) {
            const Object2D& o = det.mvObjects2D_to_View[i];
            if (o.id != eid[i] || o.name != DetectorGPU::class_name(eid[i])) return fail("Detector2D: labels");
            const float got[5] = {o.prob, o.rect.x, o.rect.y, o.rect.width, o.rect.height};
            for (int k = 0; k < 5; ++k) if (std::fabs(got[k] - erect[(size_t)i * 5 + k]) > 0.03f) return fail("Detector2D: boxes");
            if (o.id == 15) ++npers; else ++others;
        }
        if ((int)det.mvPotentialDynamicBorderForMapping.size() != npers || det.mvObjects2D.size() != others || det.mbHaveDynamicObjectForMapping != (npers > 0))
            return fail("Detector2D: person / object split");
        det.detect(im3);                                                    // the view list accumulates until draw_objects clears it
        if ((int)det.mvObjects2D_to_View.size() != 2 * dn || (int)det.mvPotentialDynamicBorderForMapping.size() != npers) return fail("Detector2D: second call");
    }
    // 8. SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) and 9. Fuse(KeyFrame*, vpMapPoints, th) on one key frame / candidate list
    int nsim3 = -1, nfused = -1, nsim3b = -1;
    if (argc >= 4) {
        int32_t kh[3]; f.read(reinterpret_cast<char*>(kh), sizeof kh);
        const int nk = kh[0], nmp = kh[1], ith = kh[2];
        KeyFrame kf; kf.N = nk;
        kf.mvKeysUn = rd<cv::KeyPoint>(f, nk); kf.mvuRight = rd<float>(f, nk);
        std::vector<uint8_t> kd = rd<uint8_t>(f, (size_t)nk * 32);
        kf.mDescriptors = cv::Mat(nk, 32, CV_8U, kd.data(), 32);
        kf.mvScaleFactors = rd<float>(f, 8); kf.mvInvLevelSigma2 = rd<float>(f, 8);
        std::vector<float> cam5 = rd<float>(f, 5), T = rd<float>(f, 16), Ow = rd<float>(f, 3);
        kf.fx = cam5[0]; kf.fy = cam5[1]; kf.cx = cam5[2]; kf.cy = cam5[3]; kf.mbf = cam5[4];
        kf.R = cv::Mat(3, 3, CV_32F); kf.t = cv::Mat(3, 1, CV_32F); kf.O = cv::Mat(3, 1, CV_32F);
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) kf.R.at<float>(r, c) = T[4 * r + c]; kf.t.at<float>(r, 0) = T[4 * r + 3]; kf.O.at<float>(r, 0) = Ow[r]; }
        std::vector<uint8_t> pvalid = rd<uint8_t>(f, nmp);
        std::vector<float> pxyz = rd<float>(f, (size_t)nmp * 3), pnrm = rd<float>(f, (size_t)nmp * 3), pmin = rd<float>(f, nmp), pmax = rd<float>(f, nmp);
        std::vector<uint8_t> pdesc = rd<uint8_t>(f, (size_t)nmp * 32);
        std::vector<int32_t> m_in = rd<int32_t>(f, nk), m_exp = rd<int32_t>(f, nk);
        int32_t nm_exp = 0; f.read(reinterpret_cast<char*>(&nm_exp), 4);
        std::vector<int32_t> kf_has = rd<int32_t>(f, nk), pobs = rd<int32_t>(f, nmp);
        std::vector<uint8_t> pinkf = rd<uint8_t>(f, nmp);
        float fth = 0; f.read(reinterpret_cast<char*>(&fth), 4);
        int32_t nf_exp = 0; f.read(reinterpret_cast<char*>(&nf_exp), 4);
        std::vector<int32_t> kf_final = rd<int32_t>(f, nk);
        std::vector<uint8_t> pbad_exp = rd<uint8_t>(f, nmp), ebad_exp = rd<uint8_t>(f, nk);
        auto make_points = [&](std::vector<MapPoint>& store, std::vector<MapPoint*>& ptrs) {
            store.assign(nmp, MapPoint()); ptrs.resize(nmp);
            for (int i = 0; i < nmp; ++i) {
                MapPoint& p = store[i];
                p.pos = cv::Mat(3, 1, CV_32F, &pxyz[3 * (size_t)i], 4).clone(); p.normal = cv::Mat(3, 1, CV_32F, &pnrm[3 * (size_t)i], 4).clone();
                p.desc = cv::Mat(1, 32, CV_8U, &pdesc[(size_t)i * 32], 32).clone(); p.mfMinDistance = pmin[i]; p.mfMaxDistance = pmax[i]; p.bad = !pvalid[i];
                ptrs[i] = &p;
            }
        };
        ORBmatcher lm(0.8f, true);
        {   // 8.
            std::vector<MapPoint> store; std::vector<MapPoint*> pts; make_points(store, pts);
            std::vector<MapPoint> dummy(nk); std::vector<MapPoint*> vpMatched(nk, nullptr);
            for (int j = 0; j < nk; ++j) if (m_in[j] >= 0) vpMatched[j] = &dummy[j];
            nsim3 = lm.SearchByProjection(&kf, T.data(), Ow.data(), pts, vpMatched, ith);
            if (nsim3 != nm_exp) return fail("SearchByProjection(KF,Scw): match count");
            for (int j = 0; j < nk; ++j) {
                MapPoint* e = m_in[j] >= 0 ? &dummy[j] : (m_exp[j] >= 0 ? pts[m_exp[j]] : nullptr);
                if (vpMatched[j] != e) return fail("SearchByProjection(KF,Scw): vpMatched");
            }
        }
        {   // 9.
            std::vector<MapPoint> store; std::vector<MapPoint*> pts; make_points(store, pts);
            std::vector<MapPoint> exist(nk); kf.mps.assign(nk, nullptr);
            for (int j = 0; j < nk; ++j) if (kf_has[j] >= 0) { exist[j].nobs = kf_has[j]; exist[j].obs.insert(&kf); kf.mps[j] = &exist[j]; }
            for (int i = 0; i < nmp; ++i) { store[i].nobs = pobs[i]; if (pinkf[i]) store[i].obs.insert(&kf); }
            nfused = lm.Fuse(&kf, pts, fth);
            if (nfused != nf_exp) return fail("Fuse: count");
            for (int j = 0; j < nk; ++j) {
                MapPoint* e = kf_final[j] >= 0 ? pts[kf_final[j]] : (kf_final[j] == -2 ? &exist[j] : nullptr);
                if (kf.mps[j] != e) return fail("Fuse: key-frame map points");
                if (kf_has[j] >= 0 && exist[j].bad != (ebad_exp[j] != 0)) return fail("Fuse: replaced key-frame points");
            }
            for (int i = 0; i < nmp; ++i) if (store[i].bad != (pbad_exp[i] != 0)) return fail("Fuse: replaced candidates");
        }
        {   // 10. SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th): two key frames with the same features and pose, different map points
            int32_t sh[2]; f.read(reinterpret_cast<char*>(sh), sizeof sh);
            const int npts = sh[0];
            float s12 = 0, th3 = 0; f.read(reinterpret_cast<char*>(&s12), 4); f.read(reinterpret_cast<char*>(&th3), 4);
            std::vector<float> R12 = rd<float>(f, 9), t12 = rd<float>(f, 3);
            std::vector<int32_t> mp1 = rd<int32_t>(f, nk), mp2 = rd<int32_t>(f, nk), pre = rd<int32_t>(f, nk), exp12 = rd<int32_t>(f, nk);
            int32_t nfound_exp = 0; f.read(reinterpret_cast<char*>(&nfound_exp), 4);
            std::vector<MapPoint> store; std::vector<MapPoint*> pts; make_points(store, pts);
            (void)npts;
            KeyFrame kf2 = kf;
            kf.mps.assign(nk, nullptr); kf2.mps.assign(nk, nullptr);
            for (int j = 0; j < nk; ++j) {
                if (mp1[j] >= 0) kf.mps[j] = pts[mp1[j]];
                if (mp2[j] >= 0) { kf2.mps[j] = pts[mp2[j]]; pts[mp2[j]]->index_in_kf[&kf2] = j; }
            }
            std::vector<MapPoint*> vpMatches12(nk, nullptr);
            for (int j = 0; j < nk; ++j) if (pre[j] >= 0) vpMatches12[j] = pts[pre[j]];
            cv::Mat R(3, 3, CV_32F, R12.data(), 12), t(3, 1, CV_32F, t12.data(), 4);
            nsim3b = lm.SearchBySim3(&kf, &kf2, vpMatches12, s12, R, t, th3);
            if (nsim3b != nfound_exp) return fail("SearchBySim3: count");
            for (int j = 0; j < nk; ++j) {
                MapPoint* e = exp12[j] >= 0 ? pts[exp12[j]] : nullptr;
                if (vpMatches12[j] != e) return fail("SearchBySim3: vpMatches12");
            }
        }
    }
    // 11. SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) and 12. SearchForTriangulation on a pair of key frames with a flattened FeatureVector
    int nbow2 = -1, ntri = -1;
    if (argc >= 4) {
        int32_t ph[3]; f.read(reinterpret_cast<char*>(ph), sizeof ph);
        const int nn[2] = {ph[0], ph[1]}, only_stereo = ph[2];
        KeyFrame K[2]; std::vector<uint8_t> dd[2]; std::vector<MapPoint> pstore[2];
        for (int q = 0; q < 2; ++q) {
            const int n = nn[q];
            K[q].N = n; K[q].mvKeysUn = rd<cv::KeyPoint>(f, n); K[q].mvuRight = rd<float>(f, n); dd[q] = rd<uint8_t>(f, (size_t)n * 32);
            K[q].mDescriptors = cv::Mat(n, 32, CV_8U, dd[q].data(), 32);
            std::vector<int32_t> node = rd<int32_t>(f, n); std::vector<uint8_t> listed = rd<uint8_t>(f, n), state = rd<uint8_t>(f, n);
            for (int i = 0; i < n; ++i) if (listed[i]) K[q].mFeatVec[(unsigned)node[i]].push_back((unsigned)i);
            pstore[q].assign(n, MapPoint()); K[q].mps.assign(n, nullptr);
            for (int i = 0; i < n; ++i) if (state[i]) { pstore[q][i].bad = state[i] == 2; K[q].mps[i] = &pstore[q][i]; }
        }
        std::vector<float> R2 = rd<float>(f, 9), t2 = rd<float>(f, 3), Cw = rd<float>(f, 3), cam4 = rd<float>(f, 4), F12v = rd<float>(f, 9), sig = rd<float>(f, 8), sfv = rd<float>(f, 8);
        K[1].R = cv::Mat(3, 3, CV_32F, R2.data(), 12).clone(); K[1].t = cv::Mat(3, 1, CV_32F, t2.data(), 4).clone(); K[0].O = cv::Mat(3, 1, CV_32F, Cw.data(), 4).clone();
        K[1].fx = cam4[0]; K[1].fy = cam4[1]; K[1].cx = cam4[2]; K[1].cy = cam4[3];
        K[1].mvLevelSigma2 = sig; K[1].mvScaleFactors = sfv;
        int32_t nm11 = 0, nm12 = 0;
        f.read(reinterpret_cast<char*>(&nm11), 4); std::vector<int32_t> e11 = rd<int32_t>(f, nn[0]);
        f.read(reinterpret_cast<char*>(&nm12), 4); std::vector<int32_t> e12 = rd<int32_t>(f, nn[0]);
        ORBmatcher bm(0.8f, true);
        std::vector<MapPoint*> v12;
        nbow2 = bm.SearchByBoW(&K[0], &K[1], v12);
        if (nbow2 != nm11 || (int)v12.size() != nn[0]) return fail("SearchByBoW(KF,KF): count");
        for (int i = 0; i < nn[0]; ++i) if (v12[i] != (e11[i] >= 0 ? K[1].mps[e11[i]] : nullptr)) return fail("SearchByBoW(KF,KF): vpMatches12");
        std::vector<std::pair<size_t, size_t> > pairs;
        cv::Mat F12m(3, 3, CV_32F, F12v.data(), 12);
        ntri = bm.SearchForTriangulation(&K[0], &K[1], F12m, pairs, only_stereo != 0);
        if (ntri != nm12 || (int)pairs.size() != nm12) return fail("SearchForTriangulation: count");
        size_t k = 0;
        for (int i = 0; i < nn[0]; ++i) if (e12[i] >= 0) { if (pairs[k].first != (size_t)i || pairs[k].second != (size_t)e12[i]) return fail("SearchForTriangulation: pairs"); ++k; }
    }
    // 13. SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
    int ninit = -1;
    if (argc >= 4) {
        int32_t ih[3]; f.read(reinterpret_cast<char*>(ih), sizeof ih);
        const int a1 = ih[0], a2 = ih[1], win = ih[2];
        Frame I1, I2;
        I1.N = a1; I1.mvKeysUn = rd<cv::KeyPoint>(f, a1); I1.mvuRight.assign(a1, -1.f);
        std::vector<uint8_t> e1 = rd<uint8_t>(f, (size_t)a1 * 32);
        I1.mDescriptors = cv::Mat(a1, 32, CV_8U, e1.data(), 32);
        I2.N = a2; I2.mvKeysUn = rd<cv::KeyPoint>(f, a2); I2.mvuRight.assign(a2, -1.f);
        std::vector<uint8_t> e2 = rd<uint8_t>(f, (size_t)a2 * 32);
        I2.mDescriptors = cv::Mat(a2, 32, CV_8U, e2.data(), 32);
        I1.mvScaleFactors = cur.mvScaleFactors; I2.mvScaleFactors = cur.mvScaleFactors;
        std::vector<float> pin_ = rd<float>(f, (size_t)a1 * 2), pout = rd<float>(f, (size_t)a1 * 2);
        int32_t nmi = 0; f.read(reinterpret_cast<char*>(&nmi), 4);
        std::vector<int32_t> em = rd<int32_t>(f, a1);
        std::vector<cv::Point2f> prevm(a1);
        for (int i = 0; i < a1; ++i) { prevm[i].x = pin_[2 * (size_t)i]; prevm[i].y = pin_[2 * (size_t)i + 1]; }
        std::vector<int> m12v;
        ORBmatcher im(0.9f, true);
        ninit = im.SearchForInitialization(I1, I2, prevm, m12v, win);
        if (ninit != nmi || (int)m12v.size() != a1) return fail("SearchForInitialization: count");
        for (int i = 0; i < a1; ++i) {
            if (m12v[i] != em[i]) return fail("SearchForInitialization: vnMatches12");
            if (prevm[i].x != pout[2 * (size_t)i] || prevm[i].y != pout[2 * (size_t)i + 1]) return fail("SearchForInitialization: vbPrevMatched");
        }
    }
    std::printf("OK shim: %d keypoints, %d matches, %d/%d kept, %d LK tracks, F ok, %d/%d map points in view, pose optimised on %d/%d edges, %d detections (%d persons), %d sim3 matches, %d fused, %d Sim3 pairs, %d BoW pairs, %d triangulation pairs, %d initialisation matches\n",
                nkp, nm, kept, nd, nlk, nview, nfr, pin, npo, ndet, npers, nsim3, nfused, nsim3b, nbow2, ntri, ninit);
    return 0;
}

// C entry points around the reference's OWN MapPoint (src/MapPoint.cc compiled unmodified against the real include/MapPoint.h and the stand-ins of
// frame_shim/mappoint_standins.h).  TEST INFRASTRUCTURE (oracle/_ref/libmappoint_ref.so).
#include <cstdint>
#include <cstring>
#include <vector>

#include "MapPoint.h"

using namespace ORB_SLAM2;
#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
struct Open : MapPoint {       // the scale-invariance distances are protected (UpdateNormalAndDepth writes them)
    Open(const cv::Mat& p, KeyFrame* kf, Map* m) : MapPoint(p, kf, m) {}
    void set_distances(float mn, float mx) { mfMinDistance = mn; mfMaxDistance = mx; }
};
cv::Mat origin() { cv::Mat p(3, 1, CV_32F); return p; }
}  // namespace

// MapPoint::ComputeDistinctiveDescriptors over n observations (one key frame each, in ascending ADDRESS order = the iteration order of the reference's
// std::map<KeyFrame*, size_t>); bad[i] marks key frames whose descriptor is skipped.  Returns 1 and the chosen descriptor, 0 when the point keeps none.
REF_API int ref_mp_distinctive(const uint8_t* desc, const uint8_t* bad, int n, uint8_t* out32) {
    std::vector<KeyFrame> kfs(n > 0 ? n : 1);
    Map map;
    for (int i = 0; i < n; ++i) {
        kfs[i].mDescriptors = cv::Mat(1, 32, CV_8U);
        std::memcpy(kfs[i].mDescriptors.ptr(0), desc + 32 * (size_t)i, 32);
        kfs[i].mvuRight.assign(1, -1.f);
        kfs[i].bad = bad && bad[i];
        kfs[i].mnScaleLevels = 8; kfs[i].mvScaleFactors.assign(8, 1.f); kfs[i].mvKeysUn.resize(1);
    }
    Open mp(origin(), &kfs[0], &map);
    for (int i = 0; i < n; ++i) mp.AddObservation(&kfs[i], 0);
    mp.ComputeDistinctiveDescriptors();
    cv::Mat d = mp.GetDescriptor();
    if (d.empty()) return 0;
    std::memcpy(out32, d.ptr(0), 32);
    return 1;
}

// MapPoint::PredictScale(dist, KeyFrame*) and (dist, Frame*) for n distances; MapPoint::GetMin / MaxDistanceInvariance
REF_API void ref_mp_predict_scale(float min_dist, float max_dist, const float* dists, int n, int nlevels, float log_scale_factor, int32_t* level_kf, int32_t* level_f, float* inv2) {
    KeyFrame kf; Frame fr; Map map;
    kf.mnScaleLevels = fr.mnScaleLevels = nlevels; kf.mfLogScaleFactor = fr.mfLogScaleFactor = log_scale_factor;
    kf.mvScaleFactors.assign(nlevels, 1.f); kf.mvKeysUn.resize(1); kf.mvuRight.assign(1, -1.f);
    Open mp(origin(), &kf, &map);
    mp.set_distances(min_dist, max_dist);
    for (int i = 0; i < n; ++i) { level_kf[i] = mp.PredictScale(dists[i], &kf); level_f[i] = mp.PredictScale(dists[i], &fr); }
    inv2[0] = mp.GetMinDistanceInvariance(); inv2[1] = mp.GetMaxDistanceInvariance();
}

// oracle/frame_geom.cpp -- TEST INFRASTRUCTURE ONLY: CPU restatement of two small per-point pieces of the Frame class:
//   Frame::ComputeStereoFromRGBD (src/Frame.cc:893-914) and Frame::isInFrustum (src/Frame.cc:296-352) with
//   MapPoint::PredictScale / Get{Min,Max}DistanceInvariance (src/MapPoint.cc:372-418).
// cv::Mat float arithmetic follows OpenCV as probed with cv2 (tests/golden/make_golden_frustum.py): a plain 3x3 * 3x1 (+ C) gemm takes
// the small-matrix path (float products summed in float, left to right); a transposed product takes the general path (double
// accumulator, one rounding); cv::norm and Mat::dot of CV_32F data accumulate in double.
#include <cmath>
#include <cstdint>

#define SGO_API extern "C" __attribute__((visibility("default")))

namespace {
struct KeyPoint28 { float x, y, size, angle, response; int32_t octave, class_id; };
}

// mvuRight / mvDepth of every keypoint: d = imDepth.at<float>(v, u) with the float coordinates truncated by the implicit int conversion;
// kps_un == NULL: undistorted == distorted keypoints (zero distortion, src/Frame.cc:656-660).
SGO_API int sgo_stereo_from_rgbd(const KeyPoint28* kps, const KeyPoint28* kps_un, int n, const float* depth, int pitch_elems, float bf, float* u_right,
                                 float* depth_out) {
    for (int i = 0; i < n; i++) {
        const float v = kps[i].y, u = kps[i].x;
        const float d = depth[(int64_t)(int)v * pitch_elems + (int)u];
        u_right[i] = -1.f; depth_out[i] = -1.f;
        if (d > 0) { depth_out[i] = d; u_right[i] = (kps_un ? kps_un[i].x : kps[i].x) - bf / d; }
    }
    return 0;
}

// Tcw: 4x4 row major.  cam: fx, fy, cx, cy, bf, minX, minY, maxX, maxY.  max_dist / min_dist are mfMaxDistance / mfMinDistance (raw).
SGO_API int sgo_is_in_frustum(const float* Tcw, const float* cam, int nlevels, float log_scale_factor, float viewing_cos_limit, int n, const float* xyz,
                              const float* normal, const float* min_dist, const float* max_dist, uint8_t* inview, float* proj_x, float* proj_y,
                              float* proj_xr, int32_t* level, float* view_cos) {
    float R[9], t[3], Ow[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R[3 * r + c] = Tcw[4 * r + c]; t[r] = Tcw[4 * r + 3]; }
    for (int r = 0; r < 3; r++) {          // mOw = -mRcw.t() * mtcw  (Frame::UpdatePoseMatrices, src/Frame.cc:289-294)
        double acc = 0; for (int k = 0; k < 3; k++) acc += (double)(-R[3 * k + r]) * (double)t[k];
        Ow[r] = (float)acc;
    }
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], bf = cam[4], minX = cam[5], minY = cam[6], maxX = cam[7], maxY = cam[8];
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        inview[i] = 0; proj_x[i] = proj_y[i] = proj_xr[i] = 0.f; level[i] = 0; view_cos[i] = 0.f;
        const float* P = xyz + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; r++) {          // mRcw*P+mtcw: small-matrix gemm path (float products summed in float, left to right)
            const float acc = R[3 * r] * P[0] + R[3 * r + 1] * P[1] + R[3 * r + 2] * P[2];
            Pc[r] = (float)((double)acc + (double)t[r]);
        }
        if (Pc[2] < 0.0f) continue;
        const float invz = 1.0f / Pc[2];
        const float u = fx * Pc[0] * invz + cx, v = fy * Pc[1] * invz + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const float maxD = 1.2f * max_dist[i], minD = 0.8f * min_dist[i];
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);      // cv::norm
        if (dist < minD || dist > maxD) continue;
        const float* Pn = normal + 3 * i;
        const float vc = (float)(((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) / dist);     // PO.dot(Pn) / dist
        if (vc < viewing_cos_limit) continue;
        const float ratio = max_dist[i] / dist;                                          // MapPoint::PredictScale
        int ns = (int)std::ceil(std::log(ratio) / log_scale_factor);                     // float log / float ceil (std:: overloads)
        if (ns < 0) ns = 0; else if (ns >= nlevels) ns = nlevels - 1;
        inview[i] = 1; proj_x[i] = u; proj_xr[i] = u - bf * invz; proj_y[i] = v; level[i] = ns; view_cos[i] = vc;
        cnt++;
    }
    return cnt;
}

// cv::undistortPoints(src, dst, K, distCoef(k1,k2,p1,p2,k3), cv::Mat(), K) as called by Frame::UndistortKeyPoints (src/Frame.cc:654-684) and
// Frame::ComputeImageBounds (:686-714).  Restates calib3d/undistort.cpp cvUndistortPointsInternal: double arithmetic, 5 fixed-point
// iterations (the default criteria), then re-projection with P = K.  Pinned bit-exactly against cv2.undistortPoints (tests/test_frame_geom.py).
SGO_API int sgo_undistort_points(const float* xy, int n, float fxf, float fyf, float cxf, float cyf, const float* dist5, float* out_xy) {
    const double fx = fxf, fy = fyf, cx = cxf, cy = cyf, ifx = 1. / fx, ify = 1. / fy;
    const double k0 = dist5[0], k1 = dist5[1], p1 = dist5[2], p2 = dist5[3], k2 = dist5[4];     // OpenCV order: k1 k2 p1 p2 k3
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx; y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k2 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0 * r2 + 0 * r2 * r2;
            x = (x0 - deltaX) * icdist; y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        out_xy[2 * i] = (float)(xx * ww); out_xy[2 * i + 1] = (float)(yy * ww);
    }
    return 0;
}

// TEST INFRASTRUCTURE: the host entry points of the C ABI that the tracking thread reaches through the mirror headers -- sgs_match_project_lastframe,
// sgs_match_project_localmap, sgs_match_bow, sgs_pose_optimization -- answered by the CPU oracle (liboracle.so) instead of the CUDA library, so that tests/test_mirror_on_reference.py
// can RUN include/sgslam/ORBmatcher.h and Optimizer.h inside the reference's own Tracking.cc without a device: what is under test there is the mirror's host logic
// (flattening the reference's object graph, the ids it hands to the matchers, writing results back), not the kernels (tests/test_gpu_*.py).  Never linked into the product.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sgs_abi.h"

extern "C" {
struct SgoFrame {               // oracle/sgs_oracle.cpp:837-847
    int32_t N; const void* keysUn; const float* uRight; const uint8_t* desc;
    float minX, minY, maxX, maxY, fx, fy, cx, cy, bf;
    int32_t nlevels; const float* scaleFactors; float logScaleFactor;
};
int sgo_search_by_projection_last(const SgoFrame* cur, const float* Tcw_cur, const float* Tcw_last, int nlast, const uint8_t* last_has_mp, const float* last_xyz,
                                  const uint8_t* last_desc, const uint8_t* last_obs, const int32_t* last_octave, const float* last_angle, float th, int bMono, int checkOri,
                                  int32_t* cur_mp_inout, const uint8_t* cur_mp_obs_in, int64_t* ncand_out);
int sgo_search_by_projection_local(const SgoFrame* fr, int nmp, const uint8_t* mp_inview, const float* projx, const float* projy, const float* projxr, const int32_t* level,
                                   const float* viewcos, const uint8_t* mp_desc, const uint8_t* mp_obs, float th, float nnratio, int32_t id_base, int32_t* f_mp_inout,
                                   uint8_t* f_mp_obs_inout, int64_t* ncand_out);
int sgo_search_by_bow(int nkf, const int32_t* kf_node, const double* kf_weight, const uint8_t* kf_valid, const uint8_t* kf_desc, const float* kf_angle, int nf, const int32_t* f_node,
                      const double* f_weight, const uint8_t* f_desc, const float* f_angle, float nnratio, int checkOri, int32_t* match_f);
int sgo_pose_optimization(const float* Tcw_in, int n, const uint8_t* has_mp, const float* xyz, const float* kp_xy, const int32_t* octave, const float* uright,
                          const float* inv_level_sigma2, float fx, float fy, float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier);
}

static SgoFrame view(const sgs_frame_view* f) {
    SgoFrame s;
    s.N = f->n; s.keysUn = f->keys_un; s.uRight = f->u_right; s.desc = f->desc;
    s.minX = f->min_x; s.minY = f->min_y; s.maxX = f->max_x; s.maxY = f->max_y; s.fx = f->fx; s.fy = f->fy; s.cx = f->cx; s.cy = f->cy; s.bf = f->bf;
    s.nlevels = f->nlevels; s.scaleFactors = f->scale_factors; s.logScaleFactor = f->nlevels > 1 ? logf(f->scale_factors[1]) : 0.f;
    return s;
}

extern "C" {

SGS_API const char* sgs_last_error(void) { return "fake backend"; }

SGS_API int sgs_match_project_lastframe(const sgs_frame_view* cur, const float* tcw_cur, const float* tcw_last, int nlast, const uint8_t* last_has_mp, const float* last_xyz,
                                        const uint8_t* last_desc, const uint8_t* last_obs, const int32_t* last_octave, const float* last_angle, float th, int mono,
                                        int check_orientation, int32_t* cur_mp_inout, const uint8_t* cur_mp_obs_in, int* nmatches, int) {
    SgoFrame s = view(cur); int64_t nc = 0;
    const int n = sgo_search_by_projection_last(&s, tcw_cur, tcw_last, nlast, last_has_mp, last_xyz, last_desc, last_obs, last_octave, last_angle, th, mono, check_orientation,
                                                cur_mp_inout, cur_mp_obs_in, &nc);
    if (nmatches) *nmatches = n;
    return n < 0 ? SGS_ERR_INVALID : SGS_OK;
}

SGS_API int sgs_match_project_localmap(const sgs_frame_view* f, int nmp, const uint8_t* mp_inview, const float* proj_x, const float* proj_y, const float* proj_xr,
                                       const int32_t* level, const float* view_cos, const uint8_t* mp_desc, const uint8_t* mp_obs, float th, float nnratio, int32_t id_base,
                                       int32_t* f_mp_inout, uint8_t* f_mp_obs_inout, int* nmatches, int) {
    SgoFrame s = view(f); int64_t nc = 0;
    const int n = sgo_search_by_projection_local(&s, nmp, mp_inview, proj_x, proj_y, proj_xr, level, view_cos, mp_desc, mp_obs, th, nnratio, id_base, f_mp_inout, f_mp_obs_inout, &nc);
    if (nmatches) *nmatches = n;
    return n < 0 ? SGS_ERR_INVALID : SGS_OK;
}

SGS_API int sgs_match_bow(int nkf, const int32_t* kf_node, const double* kf_weight, const uint8_t* kf_valid, const uint8_t* kf_desc, const float* kf_angle, int nf, const int32_t* f_node,
                          const double* f_weight, const uint8_t* f_desc, const float* f_angle, float nnratio, int check_orientation, int32_t* match_f, int* nmatches, int) {
    const int n = sgo_search_by_bow(nkf, kf_node, kf_weight, kf_valid, kf_desc, kf_angle, nf, f_node, f_weight, f_desc, f_angle, nnratio, check_orientation, match_f);
    if (nmatches) *nmatches = n;
    return n < 0 ? SGS_ERR_INVALID : SGS_OK;
}

SGS_API int sgs_pose_optimization(const sgs_camera* cam, const float* tcw_in, int n, const sgs_keypoint* kps_un, const float* uright, const uint8_t* has_mp, const float* xyz,
                                  const float* inv_level_sigma2, float* tcw_out, uint8_t* outlier, int* ninliers, int) {
    std::vector<float> xy(2 * (size_t)(n > 0 ? n : 1)); std::vector<int32_t> oct(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) { xy[2 * i] = kps_un[i].x; xy[2 * i + 1] = kps_un[i].y; oct[i] = kps_un[i].octave; }
    const int r = sgo_pose_optimization(tcw_in, n, has_mp, xyz, xy.data(), oct.data(), uright, inv_level_sigma2, cam->fx, cam->fy, cam->cx, cam->cy, cam->bf, tcw_out, outlier);
    if (ninliers) *ninliers = r;
    return SGS_OK;
}

}  // extern "C"

// match_api.cu -- C-ABI entry points of the projection matchers: the batched device API (sgs_matcher handle) and the
// single-frame host-pointer wrappers that flatten one Frame pair, run the same kernels with nframes = 1 and copy back.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "match_dev.cuh"

using namespace sgs;

struct sgs_matcher {
    int device = 0, max_frames = 0, cur_cap = 0, point_cap = 0;
    PointPre* d_pre = nullptr;
    LocalPre* d_lpre = nullptr;
    int32_t* d_events = nullptr;
};

namespace {

int pow2_at_least(int v) { int p = 1; while (p < v) p <<= 1; return p; }

MatchCam to_cam(const sgs_camera& c) {
    MatchCam m;
    m.min_x = c.min_x; m.min_y = c.min_y; m.max_x = c.max_x; m.max_y = c.max_y;
    m.fx = c.fx; m.fy = c.fy; m.cx = c.cx; m.cy = c.cy; m.bf = c.bf; m.nlevels = c.nlevels;
    for (int i = 0; i < kMaxLevels; ++i) m.scale[i] = c.scale_factors[i];
    return m;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    cudaError_t upload(const void* h, size_t bytes) {
        cudaError_t e = alloc(bytes);
        if (e != cudaSuccess || !bytes) return e;
        return cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
    }
    template <class T> T* as() { return static_cast<T*>(p); }
};

sgs_camera view_cam(const sgs_frame_view* v) {
    sgs_camera c;
    std::memset(&c, 0, sizeof c);
    c.min_x = v->min_x; c.min_y = v->min_y; c.max_x = v->max_x; c.max_y = v->max_y;
    c.fx = v->fx; c.fy = v->fy; c.cx = v->cx; c.cy = v->cy; c.bf = v->bf; c.nlevels = v->nlevels;
    for (int i = 0; i < v->nlevels && i < kMaxLevels; ++i) c.scale_factors[i] = v->scale_factors[i];
    return c;
}

}  // namespace

extern "C" {

SGS_API int sgs_matcher_create(int device, int max_frames, int cur_cap, int point_cap, sgs_matcher** out) {
    if (!out || max_frames < 1 || cur_cap < 1 || point_cap < 1) { set_error("sgs_matcher_create: bad argument"); return SGS_ERR_INVALID; }
    if (cur_cap > 8192) { set_error("sgs_matcher_create: cur_cap %d exceeds the 8192 keypoints per frame the shared-memory grid supports", cur_cap); return SGS_ERR_UNSUPPORTED; }
    *out = nullptr;
    SGS_CUDA_TRY(cudaSetDevice(device));
    sgs_matcher* m = new sgs_matcher();
    m->device = device; m->max_frames = max_frames; m->cur_cap = cur_cap; m->point_cap = point_cap;
    const size_t np = (size_t)max_frames * point_cap;
    if (cudaMalloc(&m->d_pre, np * sizeof(PointPre)) != cudaSuccess || cudaMalloc(&m->d_lpre, np * sizeof(LocalPre)) != cudaSuccess ||
        cudaMalloc(&m->d_events, np * sizeof(int32_t)) != cudaSuccess) {
        set_error("sgs_matcher_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(m->d_pre); cudaFree(m->d_lpre); cudaFree(m->d_events); delete m;
        return SGS_ERR_CUDA;
    }
    *out = m;
    return SGS_OK;
}

SGS_API void sgs_matcher_destroy(sgs_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_pre); cudaFree(m->d_lpre); cudaFree(m->d_events);
    delete m;
}

SGS_API int sgs_match_project_lastframe_batch_device(sgs_matcher* m, const sgs_lastframe_batch* a, int nframes, void* stream) {
    if (!m || !a) { set_error("sgs_match_project_lastframe_batch_device: NULL"); return SGS_ERR_INVALID; }
    if (nframes < 1 || nframes > m->max_frames) { set_error("nframes outside [1,max_frames]"); return SGS_ERR_INVALID; }
    LastFrameArgs A;
    A.cam = to_cam(a->cam);
    A.cur_kps = a->cur_kps; A.cur_desc = a->cur_desc; A.cur_uright = a->cur_uright; A.cur_n = a->cur_n;
    A.cur_cap = m->cur_cap; A.cur_cap_pow2 = pow2_at_least(m->cur_cap);
    A.last_xyz = a->last_xyz; A.last_desc = a->last_desc; A.last_flags = a->last_flags; A.last_octave = a->last_octave;
    A.last_angle = a->last_angle; A.last_n = a->last_n; A.last_cap = m->point_cap;
    A.tcw_cur = a->tcw_cur; A.tcw_last = a->tcw_last; A.th = a->th; A.mono = a->mono; A.check_ori = a->check_orientation;
    A.cur_mp = a->cur_mp; A.cur_mp_obs_in = a->cur_mp_obs_in; A.nmatches = a->nmatches; A.ncand = (unsigned long long*)a->ncand;
    A.kf_mode = 0; A.orb_dist = 100; A.log_sf = 0.f; A.kf_min_dist = A.kf_max_dist = nullptr;      // TH_HIGH, src/ORBmatcher.cc:37
    A.frame_enable = a->frame_enable;
    A.pre = m->d_pre; A.events = m->d_events;
    return launch_match_lastframe(A, nframes, (cudaStream_t)stream);
}

SGS_API int sgs_fuse_search_batch_device(const sgs_fuse_batch* a, int nframes, void* stream) {
    if (!a || !a->kf_kps || !a->kf_desc || !a->kf_uright || !a->kf_n || !a->tcw || !a->ow || !a->mp_xyz || !a->mp_normal || !a->mp_min_dist || !a->mp_max_dist ||
        !a->mp_desc || !a->mp_valid || !a->mp_n || !a->best_idx || !a->best_dist || nframes < 1 || a->kf_cap < 1 || a->mp_cap < 1) {
        set_error("sgs_fuse_search_batch_device: bad argument"); return SGS_ERR_INVALID;
    }
    if (a->cam.nlevels < 2 || a->cam.nlevels > kMaxLevels || !(a->cam.scale_factors[1] > 1.f)) { set_error("sgs_fuse_search_batch_device: camera scale table missing"); return SGS_ERR_INVALID; }
    FuseArgs A;
    A.cam = to_cam(a->cam);
    A.kf_kps = a->kf_kps; A.kf_desc = a->kf_desc; A.kf_uright = a->kf_uright; A.kf_n = a->kf_n; A.kf_cap = a->kf_cap;
    A.tcw = a->tcw; A.ow = a->ow; A.mp_xyz = a->mp_xyz; A.mp_normal = a->mp_normal; A.mp_min_dist = a->mp_min_dist; A.mp_max_dist = a->mp_max_dist;
    A.mp_desc = a->mp_desc; A.mp_valid = a->mp_valid; A.mp_n = a->mp_n; A.mp_cap = a->mp_cap; A.th = a->th; A.log_sf = logf(a->cam.scale_factors[1]);
    for (int l = 0; l < kMaxLevels; ++l) A.inv_sigma2[l] = a->inv_level_sigma2[l];
    A.sim3_variant = a->sim3_variant; A.xform2 = a->xform2;
    if (a->sim3_variant < 0 || a->sim3_variant > 3 || (a->sim3_variant == 2 && !a->xform2) || (a->sim3_variant == 3 && !a->kf_matched)) {
        set_error("sgs_fuse_search_batch_device: bad variant"); return SGS_ERR_INVALID; }
    A.best_idx = a->best_idx; A.best_dist = a->best_dist;
    A.kf_matched = a->kf_matched; A.nmatches = a->nmatches;
    return launch_fuse_search(A, nframes, (cudaStream_t)stream);
}

SGS_API int sgs_match_project_keyframe_batch_device(sgs_matcher* m, const sgs_keyframe_batch* a, int nframes, void* stream) {
    if (!m || !a) { set_error("sgs_match_project_keyframe_batch_device: NULL"); return SGS_ERR_INVALID; }
    if (nframes < 1 || nframes > m->max_frames) { set_error("nframes outside [1,max_frames]"); return SGS_ERR_INVALID; }
    if (!a->cur_kps || !a->cur_desc || !a->cur_n || !a->kf_xyz || !a->kf_desc || !a->kf_valid || !a->kf_angle || !a->kf_min_dist || !a->kf_max_dist ||
        !a->kf_n || !a->tcw_cur || !a->cur_mp || !a->nmatches || !a->ncand || !a->cur_uright) { set_error("sgs_match_project_keyframe_batch_device: NULL array"); return SGS_ERR_INVALID; }
    if (a->cam.nlevels < 2 || !(a->cam.scale_factors[1] > 1.f)) { set_error("sgs_match_project_keyframe_batch_device: camera scale table missing"); return SGS_ERR_INVALID; }
    LastFrameArgs A;
    A.cam = to_cam(a->cam);
    A.cur_kps = a->cur_kps; A.cur_desc = a->cur_desc; A.cur_uright = a->cur_uright; A.cur_n = a->cur_n;
    A.cur_cap = m->cur_cap; A.cur_cap_pow2 = pow2_at_least(m->cur_cap);
    A.last_xyz = a->kf_xyz; A.last_desc = a->kf_desc; A.last_flags = a->kf_valid; A.last_octave = nullptr;
    A.last_angle = a->kf_angle; A.last_n = a->kf_n; A.last_cap = m->point_cap;
    A.tcw_cur = a->tcw_cur; A.tcw_last = nullptr; A.th = a->th; A.mono = 1; A.check_ori = a->check_orientation;
    A.cur_mp = a->cur_mp; A.cur_mp_obs_in = nullptr; A.nmatches = a->nmatches; A.ncand = (unsigned long long*)a->ncand;
    A.kf_mode = 1; A.orb_dist = a->orb_dist; A.log_sf = logf(a->cam.scale_factors[1]); A.kf_min_dist = a->kf_min_dist; A.kf_max_dist = a->kf_max_dist;
    A.frame_enable = nullptr;
    A.pre = m->d_pre; A.events = m->d_events;
    return launch_match_lastframe(A, nframes, (cudaStream_t)stream);
}

SGS_API int sgs_match_project_localmap_batch_device(sgs_matcher* m, const sgs_localmap_batch* a, int nframes, void* stream) {
    if (!m || !a) { set_error("sgs_match_project_localmap_batch_device: NULL"); return SGS_ERR_INVALID; }
    if (nframes < 1 || nframes > m->max_frames) { set_error("nframes outside [1,max_frames]"); return SGS_ERR_INVALID; }
    LocalMapArgs A;
    A.cam = to_cam(a->cam);
    A.cur_kps = a->cur_kps; A.cur_desc = a->cur_desc; A.cur_uright = a->cur_uright; A.cur_n = a->cur_n;
    A.cur_cap = m->cur_cap; A.cur_cap_pow2 = pow2_at_least(m->cur_cap);
    A.mp_inview = a->mp_inview; A.proj_x = a->proj_x; A.proj_y = a->proj_y; A.proj_xr = a->proj_xr; A.level = a->level;
    A.view_cos = a->view_cos; A.mp_desc = a->mp_desc; A.mp_obs = a->mp_obs; A.mp_n = a->mp_n; A.mp_cap = m->point_cap;
    A.th = a->th; A.nnratio = a->nnratio; A.id_base = a->id_base;
    A.f_mp = a->f_mp; A.f_mp_obs = a->f_mp_obs; A.nmatches = a->nmatches; A.ncand = (unsigned long long*)a->ncand;
    A.pre = m->d_lpre;
    return launch_match_localmap(A, nframes, (cudaStream_t)stream);
}

SGS_API int sgs_match_project_lastframe(const sgs_frame_view* cur, const float* tcw_cur, const float* tcw_last, int nlast,
                                        const uint8_t* last_has_mp, const float* last_xyz, const uint8_t* last_desc, const uint8_t* last_obs,
                                        const int32_t* last_octave, const float* last_angle, float th, int mono, int check_orientation,
                                        int32_t* cur_mp_inout, const uint8_t* cur_mp_obs_in, int* nmatches, int device) {
    if (!cur || !tcw_cur || !tcw_last || !nmatches || nlast < 0 || cur->n < 0) { set_error("sgs_match_project_lastframe: bad argument"); return SGS_ERR_INVALID; }
    *nmatches = 0;
    if (nlast == 0 || cur->n == 0) return SGS_OK;
    if (!last_has_mp || !last_xyz || !last_desc || !last_obs || !last_octave || !last_angle || !cur_mp_inout || !cur->keys_un || !cur->u_right || !cur->desc) {
        set_error("sgs_match_project_lastframe: NULL array"); return SGS_ERR_INVALID;
    }
    sgs_matcher* m = nullptr;
    int rc = sgs_matcher_create(device, 1, cur->n, nlast, &m);
    if (rc != SGS_OK) return rc;
    struct Guard { sgs_matcher* m; ~Guard() { sgs_matcher_destroy(m); } } guard{m};
    const int n = cur->n;
    std::vector<uint8_t> flags(nlast);
    for (int i = 0; i < nlast; ++i) flags[i] = (uint8_t)((last_has_mp[i] ? 1 : 0) | (last_obs[i] ? 2 : 0));
    DevBuf kps, desc, ur, cn, xyz, ld, lf, lo, la, ln, tc, tl, mp, mpo, nm, nc;
    const int32_t n32 = n, nl32 = nlast;
    SGS_CUDA_TRY(kps.upload(cur->keys_un, sizeof(sgs_keypoint) * n)); SGS_CUDA_TRY(desc.upload(cur->desc, (size_t)32 * n));
    SGS_CUDA_TRY(ur.upload(cur->u_right, 4 * (size_t)n)); SGS_CUDA_TRY(cn.upload(&n32, 4));
    SGS_CUDA_TRY(xyz.upload(last_xyz, 12 * (size_t)nlast)); SGS_CUDA_TRY(ld.upload(last_desc, 32 * (size_t)nlast));
    SGS_CUDA_TRY(lf.upload(flags.data(), nlast)); SGS_CUDA_TRY(lo.upload(last_octave, 4 * (size_t)nlast));
    SGS_CUDA_TRY(la.upload(last_angle, 4 * (size_t)nlast)); SGS_CUDA_TRY(ln.upload(&nl32, 4));
    SGS_CUDA_TRY(tc.upload(tcw_cur, 64)); SGS_CUDA_TRY(tl.upload(tcw_last, 64));
    SGS_CUDA_TRY(mp.upload(cur_mp_inout, 4 * (size_t)n));
    if (cur_mp_obs_in) SGS_CUDA_TRY(mpo.upload(cur_mp_obs_in, n));
    SGS_CUDA_TRY(nm.alloc(4)); SGS_CUDA_TRY(nc.alloc(8)); SGS_CUDA_TRY(cudaMemset(nc.p, 0, 8));
    sgs_lastframe_batch b;
    std::memset(&b, 0, sizeof b);
    b.cam = view_cam(cur);
    b.cur_kps = kps.as<sgs_keypoint>(); b.cur_desc = desc.as<uint8_t>(); b.cur_uright = ur.as<float>(); b.cur_n = cn.as<int32_t>();
    b.last_xyz = xyz.as<float>(); b.last_desc = ld.as<uint8_t>(); b.last_flags = lf.as<uint8_t>(); b.last_octave = lo.as<int32_t>();
    b.last_angle = la.as<float>(); b.last_n = ln.as<int32_t>(); b.tcw_cur = tc.as<float>(); b.tcw_last = tl.as<float>();
    b.th = th; b.mono = mono; b.check_orientation = check_orientation;
    b.cur_mp = mp.as<int32_t>(); b.cur_mp_obs_in = cur_mp_obs_in ? mpo.as<uint8_t>() : nullptr; b.nmatches = nm.as<int32_t>(); b.ncand = nc.as<uint64_t>();
    rc = sgs_match_project_lastframe_batch_device(m, &b, 1, nullptr);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    int32_t nm_h = 0;
    SGS_CUDA_TRY(cudaMemcpy(&nm_h, nm.p, 4, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(cur_mp_inout, mp.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    *nmatches = nm_h;
    return SGS_OK;
}

SGS_API int sgs_match_project_keyframe(const sgs_frame_view* cur, const float* tcw_cur, int nkf, const uint8_t* kf_valid, const float* kf_xyz,
                                       const uint8_t* kf_desc, const float* kf_angle, const float* kf_min_dist, const float* kf_max_dist, float th,
                                       int orb_dist, int check_orientation, int32_t* cur_mp_inout, int* nmatches, int device) {
    if (!cur || !tcw_cur || !nmatches || nkf < 0 || cur->n < 0) { set_error("sgs_match_project_keyframe: bad argument"); return SGS_ERR_INVALID; }
    *nmatches = 0;
    if (nkf == 0 || cur->n == 0) return SGS_OK;
    if (!kf_valid || !kf_xyz || !kf_desc || !kf_angle || !kf_min_dist || !kf_max_dist || !cur_mp_inout || !cur->keys_un || !cur->u_right || !cur->desc) {
        set_error("sgs_match_project_keyframe: NULL array"); return SGS_ERR_INVALID;
    }
    sgs_matcher* m = nullptr;
    int rc = sgs_matcher_create(device, 1, cur->n, nkf, &m);
    if (rc != SGS_OK) return rc;
    struct Guard { sgs_matcher* m; ~Guard() { sgs_matcher_destroy(m); } } guard{m};
    const int n = cur->n;
    const int32_t n32 = n, nkf32 = nkf;
    std::vector<uint8_t> flags(nkf);
    for (int i = 0; i < nkf; ++i) flags[i] = kf_valid[i] ? 1 : 0;
    DevBuf kps, desc, ur, cn, fl, xyz, kd, ang, mn, mx, kn, tc, mp, nm, nc;
    SGS_CUDA_TRY(kps.upload(cur->keys_un, sizeof(sgs_keypoint) * n)); SGS_CUDA_TRY(desc.upload(cur->desc, (size_t)32 * n));
    SGS_CUDA_TRY(ur.upload(cur->u_right, 4 * (size_t)n)); SGS_CUDA_TRY(cn.upload(&n32, 4));
    SGS_CUDA_TRY(fl.upload(flags.data(), nkf)); SGS_CUDA_TRY(xyz.upload(kf_xyz, 12 * (size_t)nkf)); SGS_CUDA_TRY(kd.upload(kf_desc, 32 * (size_t)nkf));
    SGS_CUDA_TRY(ang.upload(kf_angle, 4 * (size_t)nkf)); SGS_CUDA_TRY(mn.upload(kf_min_dist, 4 * (size_t)nkf)); SGS_CUDA_TRY(mx.upload(kf_max_dist, 4 * (size_t)nkf));
    SGS_CUDA_TRY(kn.upload(&nkf32, 4)); SGS_CUDA_TRY(tc.upload(tcw_cur, 64)); SGS_CUDA_TRY(mp.upload(cur_mp_inout, 4 * (size_t)n));
    SGS_CUDA_TRY(nm.alloc(4)); SGS_CUDA_TRY(nc.alloc(8)); SGS_CUDA_TRY(cudaMemset(nc.p, 0, 8));
    sgs_keyframe_batch b;
    std::memset(&b, 0, sizeof b);
    b.cam = view_cam(cur);
    b.cur_kps = kps.as<sgs_keypoint>(); b.cur_desc = desc.as<uint8_t>(); b.cur_uright = ur.as<float>(); b.cur_n = cn.as<int32_t>();
    b.kf_xyz = xyz.as<float>(); b.kf_desc = kd.as<uint8_t>(); b.kf_valid = fl.as<uint8_t>(); b.kf_angle = ang.as<float>();
    b.kf_min_dist = mn.as<float>(); b.kf_max_dist = mx.as<float>(); b.kf_n = kn.as<int32_t>(); b.tcw_cur = tc.as<float>();
    b.th = th; b.orb_dist = orb_dist; b.check_orientation = check_orientation;
    b.cur_mp = mp.as<int32_t>(); b.nmatches = nm.as<int32_t>(); b.ncand = nc.as<uint64_t>();
    rc = sgs_match_project_keyframe_batch_device(m, &b, 1, nullptr);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    int32_t nm_h = 0;
    SGS_CUDA_TRY(cudaMemcpy(&nm_h, nm.p, 4, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(cur_mp_inout, mp.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    *nmatches = nm_h;
    return SGS_OK;
}

SGS_API int sgs_fuse_search(const sgs_frame_view* kf, const float* tcw, const float* ow, int nmp, const uint8_t* mp_valid, const float* mp_xyz,
                            const float* mp_normal, const float* mp_min_dist, const float* mp_max_dist, const uint8_t* mp_desc, float th,
                            const float* inv_level_sigma2, int sim3_variant, const float* xform2, int32_t* best_idx, int32_t* best_dist,
                            int32_t* kf_matched_inout, int* nmatches, int device) {
    if (!kf || !tcw || !ow || nmp < 0 || kf->n < 0 || !best_idx || !best_dist) { set_error("sgs_fuse_search: bad argument"); return SGS_ERR_INVALID; }
    if (nmatches) *nmatches = 0;
    for (int i = 0; i < nmp; ++i) { best_idx[i] = -1; best_dist[i] = 256; }
    if (nmp == 0 || kf->n == 0) return SGS_OK;
    if (!mp_valid || !mp_xyz || !mp_normal || !mp_min_dist || !mp_max_dist || !mp_desc || !kf->keys_un || !kf->u_right || !kf->desc ||
        (sim3_variant == 0 && !inv_level_sigma2) || (sim3_variant == 3 && !kf_matched_inout)) { set_error("sgs_fuse_search: NULL array"); return SGS_ERR_INVALID; }
    SGS_CUDA_TRY(cudaSetDevice(device));
    const int n = kf->n;
    const int32_t n32 = n, nmp32 = nmp;
    DevBuf kps, desc, ur, kn, tc, o, xyz, nrm, mn, mx, md, mv, mpn, xf, bi, bd, km, nm;
    SGS_CUDA_TRY(kps.upload(kf->keys_un, sizeof(sgs_keypoint) * n)); SGS_CUDA_TRY(desc.upload(kf->desc, (size_t)32 * n)); SGS_CUDA_TRY(ur.upload(kf->u_right, 4 * (size_t)n));
    SGS_CUDA_TRY(kn.upload(&n32, 4)); SGS_CUDA_TRY(tc.upload(tcw, 64)); SGS_CUDA_TRY(o.upload(ow, 12));
    SGS_CUDA_TRY(xyz.upload(mp_xyz, 12 * (size_t)nmp)); SGS_CUDA_TRY(nrm.upload(mp_normal, 12 * (size_t)nmp)); SGS_CUDA_TRY(mn.upload(mp_min_dist, 4 * (size_t)nmp));
    SGS_CUDA_TRY(mx.upload(mp_max_dist, 4 * (size_t)nmp)); SGS_CUDA_TRY(md.upload(mp_desc, 32 * (size_t)nmp)); SGS_CUDA_TRY(mv.upload(mp_valid, (size_t)nmp));
    SGS_CUDA_TRY(mpn.upload(&nmp32, 4)); SGS_CUDA_TRY(bi.alloc(4 * (size_t)nmp)); SGS_CUDA_TRY(bd.alloc(4 * (size_t)nmp)); SGS_CUDA_TRY(nm.alloc(4));
    if (xform2) SGS_CUDA_TRY(xf.upload(xform2, 48));
    if (kf_matched_inout) SGS_CUDA_TRY(km.upload(kf_matched_inout, 4 * (size_t)n));
    sgs_fuse_batch b;
    std::memset(&b, 0, sizeof b);
    b.cam = view_cam(kf);
    b.kf_kps = kps.as<sgs_keypoint>(); b.kf_desc = desc.as<uint8_t>(); b.kf_uright = ur.as<float>(); b.kf_n = kn.as<int32_t>(); b.kf_cap = n;
    b.tcw = tc.as<float>(); b.ow = o.as<float>(); b.mp_xyz = xyz.as<float>(); b.mp_normal = nrm.as<float>(); b.mp_min_dist = mn.as<float>(); b.mp_max_dist = mx.as<float>();
    b.mp_desc = md.as<uint8_t>(); b.mp_valid = mv.as<uint8_t>(); b.mp_n = mpn.as<int32_t>(); b.mp_cap = nmp; b.th = th;
    for (int l = 0; l < 16; ++l) b.inv_level_sigma2[l] = inv_level_sigma2 && l < kf->nlevels ? inv_level_sigma2[l] : 0.f;
    b.sim3_variant = sim3_variant; b.xform2 = xform2 ? xf.as<float>() : nullptr;
    b.best_idx = bi.as<int32_t>(); b.best_dist = bd.as<int32_t>(); b.kf_matched = kf_matched_inout ? km.as<int32_t>() : nullptr; b.nmatches = nm.as<int32_t>();
    const int rc = sgs_fuse_search_batch_device(&b, 1, nullptr);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    SGS_CUDA_TRY(cudaMemcpy(best_idx, bi.p, 4 * (size_t)nmp, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(best_dist, bd.p, 4 * (size_t)nmp, cudaMemcpyDeviceToHost));
    if (sim3_variant == 3) {
        SGS_CUDA_TRY(cudaMemcpy(kf_matched_inout, km.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
        int32_t nm_h = 0;
        SGS_CUDA_TRY(cudaMemcpy(&nm_h, nm.p, 4, cudaMemcpyDeviceToHost));
        if (nmatches) *nmatches = nm_h;
    }
    return SGS_OK;
}

SGS_API int sgs_search_for_initialization_batch_device(const sgs_init_batch* a, int nframes, void* stream) {
    if (!a || !a->f1_kps || !a->f1_desc || !a->f1_n || !a->f2_kps || !a->f2_desc || !a->f2_n || !a->prev_xy || !a->match12 || nframes < 1 || a->f1_cap < 1 || a->f2_cap < 1 ||
        a->window_size < 0 || !(a->cam.max_x > a->cam.min_x) || !(a->cam.max_y > a->cam.min_y)) { set_error("sgs_search_for_initialization_batch_device: bad argument"); return SGS_ERR_INVALID; }
    InitArgs A;
    A.cam = to_cam(a->cam);
    A.f1_kps = a->f1_kps; A.f1_desc = a->f1_desc; A.f1_n = a->f1_n; A.f1_cap = a->f1_cap;
    A.f2_kps = a->f2_kps; A.f2_desc = a->f2_desc; A.f2_n = a->f2_n; A.f2_cap = a->f2_cap;
    A.prev_xy = a->prev_xy; A.window = a->window_size; A.nnratio = a->nnratio; A.check_ori = a->check_orientation; A.match12 = a->match12; A.nmatches = a->nmatches;
    return launch_search_init(A, nframes, (cudaStream_t)stream);
}

SGS_API int sgs_search_for_initialization(const sgs_frame_view* f1, const sgs_frame_view* f2, float* prev_xy, int window_size, float nnratio, int check_orientation,
                                          int32_t* match12, int* nmatches, int device) {
    if (!f1 || !f2 || !nmatches || f1->n < 0 || f2->n < 0) { set_error("sgs_search_for_initialization: bad argument"); return SGS_ERR_INVALID; }
    *nmatches = 0;
    if (f1->n > 0 && match12) for (int i = 0; i < f1->n; ++i) match12[i] = -1;
    if (f1->n == 0 || f2->n == 0) return SGS_OK;
    if (!prev_xy || !match12 || !f1->keys_un || !f1->desc || !f2->keys_un || !f2->desc) { set_error("sgs_search_for_initialization: NULL array"); return SGS_ERR_INVALID; }
    SGS_CUDA_TRY(cudaSetDevice(device));
    const size_t A1 = (size_t)f1->n, A2 = (size_t)f2->n;
    const int32_t cnt[3] = {f1->n, f2->n, 0};
    DevBuf k1, d1, k2, d2, c, pv, m;
    SGS_CUDA_TRY(k1.upload(f1->keys_un, sizeof(sgs_keypoint) * A1)); SGS_CUDA_TRY(d1.upload(f1->desc, 32 * A1));
    SGS_CUDA_TRY(k2.upload(f2->keys_un, sizeof(sgs_keypoint) * A2)); SGS_CUDA_TRY(d2.upload(f2->desc, 32 * A2));
    SGS_CUDA_TRY(c.upload(cnt, 12)); SGS_CUDA_TRY(pv.upload(prev_xy, 8 * A1)); SGS_CUDA_TRY(m.alloc(4 * A1));
    sgs_init_batch b;
    std::memset(&b, 0, sizeof b);
    b.cam = view_cam(f2);
    b.f1_kps = k1.as<sgs_keypoint>(); b.f1_desc = d1.as<uint8_t>(); b.f1_n = c.as<int32_t>(); b.f1_cap = f1->n;
    b.f2_kps = k2.as<sgs_keypoint>(); b.f2_desc = d2.as<uint8_t>(); b.f2_n = c.as<int32_t>() + 1; b.f2_cap = f2->n;
    b.prev_xy = pv.as<float>(); b.window_size = window_size; b.nnratio = nnratio; b.check_orientation = check_orientation;
    b.match12 = m.as<int32_t>(); b.nmatches = c.as<int32_t>() + 2;
    const int rc = sgs_search_for_initialization_batch_device(&b, 1, nullptr);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    SGS_CUDA_TRY(cudaMemcpy(match12, m.p, 4 * A1, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(prev_xy, pv.p, 8 * A1, cudaMemcpyDeviceToHost));
    int32_t nm = 0;
    SGS_CUDA_TRY(cudaMemcpy(&nm, c.as<int32_t>() + 2, 4, cudaMemcpyDeviceToHost));
    *nmatches = nm;
    return SGS_OK;
}

SGS_API int sgs_match_bow_keyframes(int mode, int n1, const int32_t* node1, const double* weight1, const uint8_t* valid1, const uint8_t* desc1, const float* angle1,
                                    int n2, const int32_t* node2, const double* weight2, const uint8_t* valid2, const uint8_t* desc2, const float* angle2,
                                    float nnratio, int check_orientation, const uint8_t* stereo1, const uint8_t* stereo2, const float* xy1, const float* xy2,
                                    const int32_t* octave2, const float* F12, const float* epipole, const float* level_sigma2, const float* scale_factors, int nlevels,
                                    int only_stereo, int32_t* match12, int* nmatches, int device) {
    if (!nmatches || n1 < 0 || n2 < 0 || (mode != 1 && mode != 2)) { set_error("sgs_match_bow_keyframes: bad argument"); return SGS_ERR_INVALID; }
    *nmatches = 0;
    if (n1 > 0 && match12) for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return SGS_OK;
    if (!node1 || !weight1 || !valid1 || !desc1 || !angle1 || !node2 || !weight2 || !valid2 || !desc2 || !angle2 || !match12) { set_error("sgs_match_bow_keyframes: NULL array"); return SGS_ERR_INVALID; }
    if (mode == 2 && (!stereo1 || !stereo2 || !xy1 || !xy2 || !octave2 || !F12 || !epipole || !level_sigma2 || !scale_factors || nlevels < 1 || nlevels > 16)) {
        set_error("sgs_match_bow_keyframes: triangulation mode needs stereo flags, positions, octaves, F12, the epipole and the level tables"); return SGS_ERR_INVALID; }
    SGS_CUDA_TRY(cudaSetDevice(device));
    const size_t A = (size_t)n1, Bn = (size_t)n2;
    const int32_t cnt[3] = {n1, n2, 0};
    DevBuf nd1, w1, v1, d1, a1, nd2, w2, v2, d2, a2, c, m, s1, s2, p1, p2, o2, f, ep;
    SGS_CUDA_TRY(nd1.upload(node1, 4 * A)); SGS_CUDA_TRY(w1.upload(weight1, 8 * A)); SGS_CUDA_TRY(v1.upload(valid1, A)); SGS_CUDA_TRY(d1.upload(desc1, 32 * A)); SGS_CUDA_TRY(a1.upload(angle1, 4 * A));
    SGS_CUDA_TRY(nd2.upload(node2, 4 * Bn)); SGS_CUDA_TRY(w2.upload(weight2, 8 * Bn)); SGS_CUDA_TRY(v2.upload(valid2, Bn)); SGS_CUDA_TRY(d2.upload(desc2, 32 * Bn)); SGS_CUDA_TRY(a2.upload(angle2, 4 * Bn));
    SGS_CUDA_TRY(c.upload(cnt, 12)); SGS_CUDA_TRY(m.alloc(4 * A));
    sgs_bow_batch b;
    std::memset(&b, 0, sizeof b);
    b.kf_node = nd1.as<int32_t>(); b.kf_weight = w1.as<double>(); b.kf_valid = v1.as<uint8_t>(); b.kf_desc = d1.as<uint8_t>(); b.kf_angle = a1.as<float>(); b.kf_n = c.as<int32_t>(); b.kf_cap = n1;
    b.f_node = nd2.as<int32_t>(); b.f_weight = w2.as<double>(); b.f_valid = v2.as<uint8_t>(); b.f_desc = d2.as<uint8_t>(); b.f_angle = a2.as<float>(); b.f_n = c.as<int32_t>() + 1; b.f_cap = n2;
    b.keyframe_pair = mode; b.nnratio = nnratio; b.check_orientation = check_orientation; b.match_f = m.as<int32_t>(); b.nmatches = c.as<int32_t>() + 2;
    if (mode == 2) {
        SGS_CUDA_TRY(s1.upload(stereo1, A)); SGS_CUDA_TRY(s2.upload(stereo2, Bn)); SGS_CUDA_TRY(p1.upload(xy1, 8 * A)); SGS_CUDA_TRY(p2.upload(xy2, 8 * Bn));
        SGS_CUDA_TRY(o2.upload(octave2, 4 * Bn)); SGS_CUDA_TRY(f.upload(F12, 36)); SGS_CUDA_TRY(ep.upload(epipole, 8));
        b.kf_stereo = s1.as<uint8_t>(); b.f_stereo = s2.as<uint8_t>(); b.kf_xy = p1.as<float>(); b.f_xy = p2.as<float>(); b.f_octave = o2.as<int32_t>();
        b.F12 = f.as<float>(); b.epipole = ep.as<float>(); b.only_stereo = only_stereo;
        for (int l = 0; l < nlevels; ++l) { b.level_sigma2[l] = level_sigma2[l]; b.scale_factors[l] = scale_factors[l]; }
    }
    const int rc = sgs_match_bow_batch_device(&b, 1, nullptr);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    SGS_CUDA_TRY(cudaMemcpy(match12, m.p, 4 * A, cudaMemcpyDeviceToHost));
    int32_t nm = 0;
    SGS_CUDA_TRY(cudaMemcpy(&nm, c.as<int32_t>() + 2, 4, cudaMemcpyDeviceToHost));
    *nmatches = nm;
    return SGS_OK;
}

SGS_API int sgs_match_project_localmap(const sgs_frame_view* f, int nmp, const uint8_t* mp_inview, const float* proj_x, const float* proj_y,
                                       const float* proj_xr, const int32_t* level, const float* view_cos, const uint8_t* mp_desc,
                                       const uint8_t* mp_obs, float th, float nnratio, int32_t id_base, int32_t* f_mp_inout,
                                       uint8_t* f_mp_obs_inout, int* nmatches, int device) {
    if (!f || !nmatches || nmp < 0 || f->n < 0) { set_error("sgs_match_project_localmap: bad argument"); return SGS_ERR_INVALID; }
    *nmatches = 0;
    if (nmp == 0 || f->n == 0) return SGS_OK;
    if (!mp_inview || !proj_x || !proj_y || !proj_xr || !level || !view_cos || !mp_desc || !mp_obs || !f_mp_inout || !f_mp_obs_inout) {
        set_error("sgs_match_project_localmap: NULL array"); return SGS_ERR_INVALID;
    }
    sgs_matcher* m = nullptr;
    int rc = sgs_matcher_create(device, 1, f->n, nmp, &m);
    if (rc != SGS_OK) return rc;
    struct Guard { sgs_matcher* m; ~Guard() { sgs_matcher_destroy(m); } } guard{m};
    const int n = f->n;
    const int32_t n32 = n, nmp32 = nmp;
    DevBuf kps, desc, ur, cn, iv, px, py, pxr, lv, vc, md, mo, mn, mp, mpo, nm, nc;
    SGS_CUDA_TRY(kps.upload(f->keys_un, sizeof(sgs_keypoint) * n)); SGS_CUDA_TRY(desc.upload(f->desc, (size_t)32 * n));
    SGS_CUDA_TRY(ur.upload(f->u_right, 4 * (size_t)n)); SGS_CUDA_TRY(cn.upload(&n32, 4));
    SGS_CUDA_TRY(iv.upload(mp_inview, nmp)); SGS_CUDA_TRY(px.upload(proj_x, 4 * (size_t)nmp)); SGS_CUDA_TRY(py.upload(proj_y, 4 * (size_t)nmp));
    SGS_CUDA_TRY(pxr.upload(proj_xr, 4 * (size_t)nmp)); SGS_CUDA_TRY(lv.upload(level, 4 * (size_t)nmp)); SGS_CUDA_TRY(vc.upload(view_cos, 4 * (size_t)nmp));
    SGS_CUDA_TRY(md.upload(mp_desc, 32 * (size_t)nmp)); SGS_CUDA_TRY(mo.upload(mp_obs, nmp)); SGS_CUDA_TRY(mn.upload(&nmp32, 4));
    SGS_CUDA_TRY(mp.upload(f_mp_inout, 4 * (size_t)n)); SGS_CUDA_TRY(mpo.upload(f_mp_obs_inout, n));
    SGS_CUDA_TRY(nm.alloc(4)); SGS_CUDA_TRY(nc.alloc(8)); SGS_CUDA_TRY(cudaMemset(nc.p, 0, 8));
    sgs_localmap_batch b;
    std::memset(&b, 0, sizeof b);
    b.cam = view_cam(f);
    b.cur_kps = kps.as<sgs_keypoint>(); b.cur_desc = desc.as<uint8_t>(); b.cur_uright = ur.as<float>(); b.cur_n = cn.as<int32_t>();
    b.mp_inview = iv.as<uint8_t>(); b.proj_x = px.as<float>(); b.proj_y = py.as<float>(); b.proj_xr = pxr.as<float>(); b.level = lv.as<int32_t>();
    b.view_cos = vc.as<float>(); b.mp_desc = md.as<uint8_t>(); b.mp_obs = mo.as<uint8_t>(); b.mp_n = mn.as<int32_t>();
    b.th = th; b.nnratio = nnratio; b.id_base = id_base;
    b.f_mp = mp.as<int32_t>(); b.f_mp_obs = mpo.as<uint8_t>(); b.nmatches = nm.as<int32_t>(); b.ncand = nc.as<uint64_t>();
    rc = sgs_match_project_localmap_batch_device(m, &b, 1, nullptr);
    if (rc != SGS_OK) return rc;
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    int32_t nm_h = 0;
    SGS_CUDA_TRY(cudaMemcpy(&nm_h, nm.p, 4, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(f_mp_inout, mp.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(f_mp_obs_inout, mpo.p, n, cudaMemcpyDeviceToHost));
    *nmatches = nm_h;
    return SGS_OK;
}

}  // extern "C"

// orb_plan.cpp -- host-side planning for the batched extractor: scale tables, per-level feature budget, the
// circular-patch row extents, level geometry, the FAST cell table and the fixed-point bilinear tables.
// Semantics follow ORBextractor::ORBextractor (src/ORBextractor.cc:411-471), ComputePyramid (:1108-1133) and the
// cell loop of ComputeKeyPointsOctTree (:766-830); arithmetic types are reproduced exactly (float vs double).
#include <cstdio>
#include <string>
#include <cstdlib>
#include <map>
#include <cmath>
#include <cstdarg>
#include <cstring>

#include "sgs_common.h"

namespace sgs {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

const char* last_error_cstr() { return g_last_error.c_str(); }

int cv_round_f(float v) { return (int)std::lrint((double)v); }

static int align_up(int v, int a) { return (v + a - 1) / a * a; }

// cv::resize INTER_LINEAR coefficient table for one axis (fixed point, 11 fractional bits).
static void bilinear_axis(int src, int dst, std::vector<int16_t>& tab) {
    tab.assign((size_t)dst * 4, 0);
    const double to_src = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; ++d) {
        float pos = (float)((d + 0.5) * to_src - 0.5);
        int s = (int)std::floor(pos);
        float frac = pos - (float)s;
        if (s < 0) { s = 0; frac = 0.f; }
        if (s >= src - 1) { s = src - 1; frac = 0.f; }
        tab[4 * d + 0] = (int16_t)s;
        tab[4 * d + 1] = (int16_t)cv_round_f((1.f - frac) * 2048.f);
        tab[4 * d + 2] = (int16_t)cv_round_f(frac * 2048.f);
    }
}

int make_plan(const sgs_orb_params& p, int width, int height, OrbPlan* plan) {
    if (p.nlevels < 1 || p.nlevels > kMaxLevels) { set_error("nlevels=%d outside [1,%d]", p.nlevels, kMaxLevels); return SGS_ERR_INVALID; }
    if (p.nfeatures < 1 || !(p.scale_factor > 1.0f)) { set_error("nfeatures=%d scale_factor=%f invalid", p.nfeatures, p.scale_factor); return SGS_ERR_INVALID; }
    if (p.scale_factor == 2.0f) { set_error("scale_factor 2.0 makes cv::resize switch to INTER_AREA (not supported)"); return SGS_ERR_UNSUPPORTED; }
    if (p.ini_th_fast < 1 || p.min_th_fast < 1 || p.ini_th_fast > 254 || p.min_th_fast > 254) { set_error("FAST thresholds must be in [1,254]"); return SGS_ERR_INVALID; }
    if (width > 4096 || height > 4096 || width < 1 || height < 1) { set_error("image %dx%d outside supported range", width, height); return SGS_ERR_INVALID; }
    OrbPlan& P = *plan;
    P.p = p; P.width = width; P.height = height; P.nlevels = p.nlevels;
    const int L = p.nlevels;
    P.scale.assign(L, 1.f); P.sigma2.assign(L, 1.f); P.inv_scale.assign(L, 1.f); P.inv_sigma2.assign(L, 1.f);
    const double sf = (double)p.scale_factor;
    for (int i = 1; i < L; ++i) {
        P.scale[i] = (float)((double)P.scale[i - 1] * sf);
        P.sigma2[i] = P.scale[i] * P.scale[i];
    }
    for (int i = 0; i < L; ++i) {
        P.inv_scale[i] = 1.0f / P.scale[i];
        P.inv_sigma2[i] = 1.0f / P.sigma2[i];
    }
    // geometric distribution of the feature budget over the levels
    P.n_per_level.assign(L, 0);
    const float ratio = (float)(1.0 / sf);
    float want = (float)p.nfeatures * (1.f - ratio) / (1.f - (float)std::pow((double)ratio, (double)L));
    int assigned = 0;
    for (int i = 0; i + 1 < L; ++i) {
        P.n_per_level[i] = cv_round_f(want);
        assigned += P.n_per_level[i];
        want *= ratio;
    }
    P.n_per_level[L - 1] = p.nfeatures > assigned ? p.nfeatures - assigned : 0;
    // row extents of the circular patch used by the intensity centroid
    {
        const float hs = (float)kHalfPatch * std::sqrt(2.f) / 2.f;
        const int vmax = (int)std::floor(hs + 1.f);
        const int vmin = (int)std::ceil(hs);
        for (int v = 0; v <= kHalfPatch; ++v) P.umax[v] = 0;
        for (int v = 0; v <= vmax; ++v) P.umax[v] = (int)std::lrint(std::sqrt((double)(kHalfPatch * kHalfPatch - v * v)));
        for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (P.umax[v0] == P.umax[v0 + 1]) ++v0;
            P.umax[v] = v0;
            ++v0;
        }
    }
    P.lv.assign(L, LevelGeom());
    P.cells.clear();
    P.xtab.assign(L, {}); P.ytab.assign(L, {});
    int64_t img_off = 0, cand_off = 0;
    int kp_total = 0;
    for (int l = 0; l < L; ++l) {
        LevelGeom& g = P.lv[l];
        g.w = cv_round_f((float)width * P.inv_scale[l]);
        g.h = cv_round_f((float)height * P.inv_scale[l]);
        g.pitch = align_up(g.w, 16);
        g.scale = P.scale[l];
        g.patch_size = (float)(int)((float)kPatchSize * P.scale[l]);
        g.max_bx = g.w - kEdge + 3;
        g.max_by = g.h - kEdge + 3;
        const float fw = (float)(g.max_bx - kMinBorder), fh = (float)(g.max_by - kMinBorder);
        g.n_cols = (int)(fw / 30.f);
        g.n_rows = (int)(fh / 30.f);
        if (g.n_cols < 1 || g.n_rows < 1) {
            set_error("level %d is %dx%d: too small for the 30-px FAST cell grid (the reference divides by zero here)", l, g.w, g.h);
            return SGS_ERR_UNSUPPORTED;
        }
        if (g.n_cols > 255 || g.n_rows > 255) { set_error("level %d has too many FAST cells", l); return SGS_ERR_UNSUPPORTED; }
        g.w_cell = (int)std::ceil(fw / (float)g.n_cols);
        g.h_cell = (int)std::ceil(fh / (float)g.n_rows);
        g.n_target = P.n_per_level[l];
        g.n_ini = (int)std::round((float)(g.max_bx - kMinBorder) / (float)(g.max_by - kMinBorder));
        if (g.n_ini < 1) { set_error("level %d: portrait aspect gives zero quadtree roots (reference quirk Q2)", l); return SGS_ERR_UNSUPPORTED; }
        if (g.n_ini > 255) { set_error("level %d: aspect ratio too extreme", l); return SGS_ERR_UNSUPPORTED; }
        g.h_x = (float)(g.max_bx - kMinBorder) / (float)g.n_ini;
        g.cell_begin = (int)P.cells.size();
        for (int i = 0; i < g.n_rows; ++i) {
            const int y0 = kMinBorder + i * g.h_cell;
            if (y0 >= g.max_by - 3) continue;
            const int y1 = (y0 + g.h_cell + 6 > g.max_by) ? g.max_by : y0 + g.h_cell + 6;
            for (int j = 0; j < g.n_cols; ++j) {
                const int x0 = kMinBorder + j * g.w_cell;
                if (x0 >= g.max_bx - 6) continue;
                const int x1 = (x0 + g.w_cell + 6 > g.max_bx) ? g.max_bx : x0 + g.w_cell + 6;
                if (x1 - x0 < 7 || y1 - y0 < 7) continue;  // cv::FAST finds nothing in a view without interior
                FastCell c;
                c.x0 = (uint16_t)x0; c.y0 = (uint16_t)y0; c.x1 = (uint16_t)x1; c.y1 = (uint16_t)y1;
                c.level = (uint8_t)l; c.ci = (uint8_t)i; c.cj = (uint8_t)j; c.pad = 0;
                P.cells.push_back(c);
            }
        }
        g.cell_end = (int)P.cells.size();
        // strict 3x3 local maxima cannot be 8-adjacent: at most ceil(w/2)*ceil(h/2) per interior
        const int iw = g.max_bx - kMinBorder - 6, ih = g.max_by - kMinBorder - 6;
        g.cand_cap = ((iw + 1) / 2 + g.n_cols) * ((ih + 1) / 2 + g.n_rows);
        g.kp_cap = (g.n_target > 4 * g.n_ini ? g.n_target : 4 * g.n_ini) + 3;
        g.img_off = img_off; g.frame_stride = (int64_t)g.h * g.pitch;
        g.cand_off = cand_off;
        img_off += g.frame_stride;  // provisional: re-based per batch size by the extractor
        cand_off += g.cand_cap;
        kp_total += g.kp_cap;
        if (l > 0) {
            bilinear_axis(P.lv[l - 1].w, g.w, P.xtab[l]);
            bilinear_axis(P.lv[l - 1].h, g.h, P.ytab[l]);
        }
    }
    P.max_kp_per_frame = kp_total;
    P.cand_per_frame = cand_off;
    P.pyr_bytes_per_frame = img_off;
    return SGS_OK;
}

}  // namespace sgs

// ---- settings file (Examples/TUM*.yaml): flat "key: value" lines of an OpenCV FileStorage YAML document
extern "C" SGS_API int sgs_settings_load(const char* path, sgs_settings* out) {
    using sgs::set_error;
    if (!path || !out) { set_error("sgs_settings_load: bad argument"); return SGS_ERR_INVALID; }
    FILE* f = std::fopen(path, "r");
    if (!f) { set_error("sgs_settings_load: cannot open %s", path); return SGS_ERR_INVALID; }
    std::map<std::string, double> kv;
    char line[1024];
    bool first = true, yaml = false;
    while (std::fgets(line, sizeof line, f)) {
        std::string s(line);
        if (first) { first = false; if (s.compare(0, 5, "%YAML") == 0) { yaml = true; continue; } }
        const size_t hash = s.find('#');
        if (hash != std::string::npos) s.erase(hash);
        const size_t colon = s.find(':');
        if (colon == std::string::npos) continue;
        std::string key = s.substr(0, colon), val = s.substr(colon + 1);
        auto trim = [](std::string& t) { const char* ws = " \t\r\n\""; t.erase(0, t.find_first_not_of(ws)); const size_t e = t.find_last_not_of(ws); t.erase(e == std::string::npos ? 0 : e + 1); };
        trim(key); trim(val);
        if (key.empty() || val.empty()) continue;
        char* end = nullptr;
        const double v = std::strtod(val.c_str(), &end);
        if (end != val.c_str()) kv[key] = v;
    }
    std::fclose(f);
    if (!yaml) { set_error("sgs_settings_load: %s does not start with %%YAML", path); return SGS_ERR_INVALID; }
    auto get = [&](const char* k) -> double { auto it = kv.find(k); return it == kv.end() ? 0.0 : it->second; };
    sgs_settings S;
    std::memset(&S, 0, sizeof S);
    S.fx = (float)get("Camera.fx"); S.fy = (float)get("Camera.fy"); S.cx = (float)get("Camera.cx"); S.cy = (float)get("Camera.cy");
    S.k1 = (float)get("Camera.k1"); S.k2 = (float)get("Camera.k2"); S.p1 = (float)get("Camera.p1"); S.p2 = (float)get("Camera.p2"); S.k3 = (float)get("Camera.k3");
    S.bf = (float)get("Camera.bf"); S.fps = (float)get("Camera.fps");
    S.width = (int32_t)get("Camera.width"); S.height = (int32_t)get("Camera.height"); S.rgb = (int32_t)get("Camera.RGB");
    if (!(S.fx != 0.f)) { set_error("sgs_settings_load: %s has no Camera.fx", path); return SGS_ERR_INVALID; }
    S.th_depth = S.bf * (float)get("ThDepth") / S.fx;
    const float dmf = (float)get("DepthMapFactor");
    S.depth_map_factor = std::fabs(dmf) < 1e-5 ? 1.f : 1.0f / dmf;
    S.orb.nfeatures = (int32_t)get("ORBextractor.nFeatures"); S.orb.scale_factor = (float)get("ORBextractor.scaleFactor"); S.orb.nlevels = (int32_t)get("ORBextractor.nLevels");
    S.orb.ini_th_fast = (int32_t)get("ORBextractor.iniThFAST"); S.orb.min_th_fast = (int32_t)get("ORBextractor.minThFAST");
    S.detection_confidence_threshold = (float)get("Detector2D.detection_confidence_threshold");
    S.dynamic_detection_confidence_threshold = (float)get("Detector2D.dynamic_detection_confidence_threshold");
    *out = S;
    return SGS_OK;
}


// host_checks.cpp -- compiles the host/device-shared logic headers with g++ so that their control flow can be
// verified on a machine without a GPU (tests/test_host_logic.py compares against the CPU oracle).
// This is a TEST HARNESS for product headers: the shipping path runs the same code inside CUDA blocks
// (quadtree.cu); nothing here is reachable from libsgs_cuda.so.
#include <cstring>
#include <vector>

#include "quadtree_core.h"
#include "sgs_common.h"

namespace {
struct HostCtx {
    int tid() const { return 0; }
    int nthreads() const { return 1; }
    void sync() const {}
    int group() const { return 1; }
    void group_sync() const {}
};
}  // namespace

extern "C" __attribute__((visibility("default")))
int sgs_hostcheck_plan(const sgs_orb_params* p, int w, int h, int32_t* level_wh /*[L][2]*/, int32_t* n_per_level,
                       int32_t* umax16, float* scale, int32_t* ncells, int32_t* geom /*[L][8]: n_cols,n_rows,w_cell,h_cell,max_bx,max_by,n_ini,cand_cap*/) {
    sgs::OrbPlan P;
    int st = sgs::make_plan(*p, w, h, &P);
    if (st) return st;
    for (int l = 0; l < P.nlevels; ++l) {
        level_wh[2 * l] = P.lv[l].w; level_wh[2 * l + 1] = P.lv[l].h;
        n_per_level[l] = P.n_per_level[l]; scale[l] = P.scale[l];
        const sgs::LevelGeom& g = P.lv[l];
        int32_t v[8] = {g.n_cols, g.n_rows, g.w_cell, g.h_cell, g.max_bx, g.max_by, g.n_ini, g.cand_cap};
        std::memcpy(geom + 8 * l, v, sizeof v);
    }
    for (int i = 0; i < 16; ++i) umax16[i] = P.umax[i];
    *ncells = (int)P.cells.size();
    return 0;
}

// Runs qt_distribute on the host for the level geometry of (params, w, h, level).  cands: int32 triples (x, y, score)
// relative to (16,16).  out: selected candidates as triples in output order.
extern "C" __attribute__((visibility("default")))
int sgs_hostcheck_quadtree(const sgs_orb_params* p, int w, int h, int level, const int32_t* cands, int n, int n_target_override,
                           int32_t* out, int out_cap) {
    sgs::OrbPlan P;
    int st = sgs::make_plan(*p, w, h, &P);
    if (st) return -st;
    const sgs::LevelGeom& lg = P.lv[level];
    sgs::QtGeom g;
    g.n_ini = lg.n_ini; g.h_x = lg.h_x; g.root_h = lg.max_by - sgs::kMinBorder;
    g.n_cols = lg.n_cols; g.w_cell = lg.w_cell; g.h_cell = lg.h_cell;
    g.n_target = n_target_override >= 0 ? n_target_override : lg.n_target;
    std::vector<uint32_t> packed(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) packed[i] = sgs::qt_pack(cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]);
    int n_sort = 1; while (n_sort < n) n_sort <<= 1;
    const int cap = sgs::qt_pool_cap(g.n_target, g.n_ini);
    std::vector<uint64_t> keys(n_sort), ekey(cap);
    std::vector<int32_t> lo(cap), hi(cap), seq(cap), fl(cap), la(cap), lb(cap), ea(cap), eb(cap), bnd(3 * cap), child(4 * cap), sc(16);
    std::vector<uint8_t> depth(cap), flag(cap);
    sgs::QtWork wk;
    wk.keys = keys.data(); wk.n = n; wk.n_sort = n_sort;
    wk.lo = lo.data(); wk.hi = hi.data(); wk.seq = seq.data(); wk.depth = depth.data(); wk.flag = flag.data();
    wk.free_list = fl.data(); wk.pool_cap = cap; wk.list_a = la.data(); wk.list_b = lb.data();
    wk.exp_a = ea.data(); wk.exp_b = eb.data(); wk.exp_key = ekey.data(); wk.bnd = bnd.data(); wk.child = child.data(); wk.sc = sc.data();
    const int kcap = (g.n_target > 4 * g.n_ini ? g.n_target : 4 * g.n_ini) + 3;
    std::vector<uint32_t> sel(kcap);
    int32_t nout = 0;
    HostCtx ctx;
    sgs::qt_distribute(ctx, packed.data(), g, wk, sel.data(), kcap, &nout);
    if (nout > out_cap || nout > kcap) return -1000000 - nout;
    for (int i = 0; i < nout; ++i) { out[3 * i] = sgs::qt_x(sel[i]); out[3 * i + 1] = sgs::qt_y(sel[i]); out[3 * i + 2] = sgs::qt_score(sel[i]); }
    return nout;
}

// glibc_logf (sgs_logf.h, the product's restatement of libm's logf used by PredictScale) against the running libm: number of floats whose
// results differ in any bit among the `count` bit patterns first, first + stride, ...
#include <cmath>
#include "sgs_logf.h"
extern "C" __attribute__((visibility("default")))
long long sgs_hostcheck_logf_mismatches(uint32_t first, uint32_t stride, long long count) {
    long long bad = 0;
    uint32_t b = first;
    for (long long i = 0; i < count; ++i, b += stride) {
        float x; std::memcpy(&x, &b, 4);
        const float r = ::logf(x), p = sgs::glibc_logf(x);
        if (std::memcmp(&r, &p, 4) != 0 && !(r != r && p != p)) ++bad;
    }
    return bad;
}
extern "C" __attribute__((visibility("default")))
float sgs_hostcheck_logf(float x) { return sgs::glibc_logf(x); }

// oracle/bow.cpp -- TEST INFRASTRUCTURE ONLY: CPU restatement of the bag-of-words pieces on the tracking fallback path:
//   DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
//       (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1140-1272, called from Frame::ComputeBoW src/Frame.cc:421-428 with levelsup = 4),
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)  (src/ORBmatcher.cc:159-290).
// The vocabulary tree is given as flat arrays (parent of every node in node-id order; DBoW2 appends children to their parent in that
// order, TemplatedVocabulary.h:1467-1508 / loadFromTextFile).  Weighting TF_IDF, scoring L1_NORM (ORBvoc).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#define SGO_API extern "C" __attribute__((visibility("default")))

namespace {

struct Voc {
    int k = 0, L = 0;
    std::vector<std::vector<int>> children;
    std::vector<uint8_t> desc;       // [nnodes][32]
    std::vector<double> weight;      // leaf weight (idf)
    std::vector<int> word_id;        // leaves numbered in node-id order (createWords / load order)
};

inline int hamming(const uint8_t* a, const uint8_t* b) {
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(a); const uint32_t* pb = reinterpret_cast<const uint32_t*>(b);
    int d = 0;
    for (int i = 0; i < 8; i++) d += __builtin_popcount(pa[i] ^ pb[i]);
    return d;
}

const int TH_LOW = 50, HISTO_LENGTH = 30;

void three_maxima(std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {      // ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1603-1645
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

}  // namespace

SGO_API void* sgo_voc_create(int k, int L, int nnodes, const int32_t* parent, const uint8_t* desc, const double* weight) {
    Voc* v = new Voc();
    v->k = k; v->L = L;
    v->children.resize(nnodes); v->desc.assign(desc, desc + (size_t)nnodes * 32); v->weight.assign(weight, weight + nnodes); v->word_id.assign(nnodes, -1);
    for (int i = 1; i < nnodes; i++) v->children[parent[i]].push_back(i);
    int w = 0;
    for (int i = 1; i < nnodes; i++) if (v->children[i].empty()) v->word_id[i] = w++;
    return v;
}
SGO_API void sgo_voc_free(void* h) { delete static_cast<Voc*>(h); }

// per feature: word id, weight, node id at level L - levelsup (0 = root when that level is <= 0)
SGO_API int sgo_bow_transform(void* h, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight, int32_t* node) {
    const Voc& V = *static_cast<Voc*>(h);
    const int nid_level = V.L - levelsup;
    for (int f = 0; f < n; f++) {
        int final_id = 0, level = 0, nid = 0;
        do {
            ++level;
            const std::vector<int>& ch = V.children[final_id];
            final_id = ch[0];
            int best = hamming(desc + 32 * (size_t)f, &V.desc[32 * (size_t)final_id]);
            for (size_t c = 1; c < ch.size(); c++) {
                const int d = hamming(desc + 32 * (size_t)f, &V.desc[32 * (size_t)ch[c]]);
                if (d < best) { best = d; final_id = ch[c]; }
            }
            if (level == nid_level) nid = final_id;
        } while (!V.children[final_id].empty());
        word[f] = V.word_id[final_id]; weight[f] = V.weight[final_id]; node[f] = nid;
    }
    return 0;
}

// BowVector of the features (TF_IDF weights accumulated per word in feature order, then L1-normalised): returns the number of words,
// ids ascending (std::map order) in out_word / out_value (capacity n)
SGO_API int sgo_bow_vector(const int32_t* word, const double* weight, int n, int32_t* out_word, double* out_value) {
    std::map<int, double> v;
    for (int i = 0; i < n; i++) if (weight[i] > 0) v[word[i]] += weight[i];        // BowVector::addWeight
    double norm = 0;
    for (auto& kv : v) norm += std::fabs(kv.second);                                 // BowVector::normalize(L1)
    int c = 0;
    for (auto& kv : v) { out_word[c] = kv.first; out_value[c] = norm > 0 ? kv.second / norm : kv.second; c++; }
    return c;
}

// SearchByBoW(pKF, F, vpMapPointMatches): kf_valid[i] = map point i exists and is not bad; a feature takes part iff its word weight > 0.
// match_f[j] = index of the key-frame feature whose map point was assigned to F's feature j, or -1.
SGO_API int sgo_search_by_bow(int nkf, const int32_t* kf_node, const double* kf_weight, const uint8_t* kf_valid, const uint8_t* kf_desc, const float* kf_angle,
                              int nf, const int32_t* f_node, const double* f_weight, const uint8_t* f_desc, const float* f_angle, float nnratio, int checkOri,
                              int32_t* match_f) {
    std::map<int, std::vector<int>> fvK, fvF;
    for (int i = 0; i < nkf; i++) if (kf_weight[i] > 0) fvK[kf_node[i]].push_back(i);
    for (int j = 0; j < nf; j++) if (f_weight[j] > 0) fvF[f_node[j]].push_back(j);
    for (int j = 0; j < nf; j++) match_f[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    auto K = fvK.begin(); auto F = fvF.begin();
    while (K != fvK.end() && F != fvF.end()) {
        if (K->first == F->first) {
            for (int realIdxKF : K->second) {
                if (!kf_valid[realIdxKF]) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int realIdxF : F->second) {
                    if (match_f[realIdxF] >= 0) continue;
                    const int dist = hamming(kf_desc + 32 * (size_t)realIdxKF, f_desc + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match_f[bestIdxF] = realIdxKF;
                    if (checkOri) {
                        float rot = kf_angle[realIdxKF] - f_angle[bestIdxF];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(bestIdxF);
                    }
                    nmatches++;
                }
            }
            ++K; ++F;
        } else if (K->first < F->first) K = fvK.lower_bound(F->first);
        else F = fvF.lower_bound(K->first);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match_f[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:524-657, loop closing): both sides
// need a good map point, the distance test is strict (< TH_LOW), the result is indexed by the FIRST key frame's features:
// match_1[i1] = feature of the second key frame whose map point is assigned, or -1.
SGO_API int sgo_search_by_bow_kfkf(int n1, const int32_t* node1, const double* weight1, const uint8_t* valid1, const uint8_t* desc1, const float* angle1,
                                   int n2, const int32_t* node2, const double* weight2, const uint8_t* valid2, const uint8_t* desc2, const float* angle2,
                                   float nnratio, int checkOri, int32_t* match_1) {
    std::map<int, std::vector<int>> fv1, fv2;
    for (int i = 0; i < n1; i++) if (weight1[i] > 0) fv1[node1[i]].push_back(i);
    for (int j = 0; j < n2; j++) if (weight2[j] > 0) fv2[node2[j]].push_back(j);
    for (int i = 0; i < n1; i++) match_1[i] = -1;
    std::vector<uint8_t> matched2(n2, 0);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    auto A = fv1.begin(); auto B = fv2.begin();
    while (A != fv1.end() && B != fv2.end()) {
        if (A->first == B->first) {
            for (int idx1 : A->second) {
                if (!valid1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int idx2 : B->second) {
                    if (matched2[idx2] || !valid2[idx2]) continue;
                    const int dist = hamming(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match_1[idx1] = bestIdx2; matched2[bestIdx2] = 1;
                    if (checkOri) {
                        float rot = angle1[idx1] - angle2[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                    nmatches++;
                }
            }
            ++A; ++B;
        } else if (A->first < B->first) A = fv1.lower_bound(B->first);
        else B = fv2.lower_bound(A->first);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match_1[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:659-827, LocalMapping::CreateNewMapPoints)
// with ORBmatcher::CheckDistEpipolarLine (:140-157).  free1 / free2: the feature has NO map point; stereo1 / stereo2: mvuRight >= 0;
// xy / octave: mvKeysUn; F12: 3x3 float row major; (ex, ey): epipole of KF1's centre in KF2 (:667-672); sigma2 / scale: per-level tables of KF2.
// match_1[i1] = matched feature of KF2 or -1.
SGO_API int sgo_search_for_triangulation(int n1, const int32_t* node1, const double* weight1, const uint8_t* free1, const uint8_t* stereo1, const uint8_t* desc1,
                                         const float* xy1, const float* angle1, int n2, const int32_t* node2, const double* weight2, const uint8_t* free2,
                                         const uint8_t* stereo2, const uint8_t* desc2, const float* xy2, const int32_t* octave2, const float* angle2,
                                         const float* F12, float ex, float ey, const float* sigma2, const float* scale, int only_stereo, int checkOri,
                                         int32_t* match_1) {
    std::map<int, std::vector<int>> fv1, fv2;
    for (int i = 0; i < n1; i++) if (weight1[i] > 0) fv1[node1[i]].push_back(i);
    for (int j = 0; j < n2; j++) if (weight2[j] > 0) fv2[node2[j]].push_back(j);
    for (int i = 0; i < n1; i++) match_1[i] = -1;
    std::vector<uint8_t> matched2(n2, 0);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    auto A = fv1.begin(); auto B = fv2.begin();
    while (A != fv1.end() && B != fv2.end()) {
        if (A->first == B->first) {
            for (int idx1 : A->second) {
                if (!free1[idx1]) continue;
                const bool bStereo1 = stereo1[idx1] != 0;
                if (only_stereo && !bStereo1) continue;
                const float x1 = xy1[2 * idx1], y1 = xy1[2 * idx1 + 1];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int idx2 : B->second) {
                    if (matched2[idx2] || !free2[idx2]) continue;
                    const bool bStereo2 = stereo2[idx2] != 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = hamming(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const float x2 = xy2[2 * idx2], y2 = xy2[2 * idx2 + 1];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - x2, distey = ey - y2;
                        if (distex * distex + distey * distey < 100 * scale[octave2[idx2]]) continue;
                    }
                    // CheckDistEpipolarLine: l = x1' F12
                    const float a = x1 * F12[0] + y1 * F12[3] + F12[6], b = x1 * F12[1] + y1 * F12[4] + F12[7], c = x1 * F12[2] + y1 * F12[5] + F12[8];
                    const float num = a * x2 + b * y2 + c, den = a * a + b * b;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * sigma2[octave2[idx2]]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match_1[idx1] = bestIdx2; matched2[bestIdx2] = 1; nmatches++;
                    if (checkOri) {
                        float rot = angle1[idx1] - angle2[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            ++A; ++B;
        } else if (A->first < B->first) A = fv1.lower_bound(B->first);
        else B = fv2.lower_bound(A->first);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match_1[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307): among the n observed descriptors of a map point, the one whose median
// Hamming distance to all of them (itself included, distance 0) is smallest; vDists[0.5 * (n - 1)] after sorting; the first minimum wins.
// Returns the index (0 for n == 0).
SGO_API int sgo_distinctive_descriptor(const uint8_t* desc, int n) {
    int best_median = 0x7fffffff, best_idx = 0;
    std::vector<int> d(n);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) d[j] = i == j ? 0 : hamming(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
        std::sort(d.begin(), d.end());
        const int median = d[(int)(0.5 * (n - 1))];
        if (median < best_median) { best_median = median; best_idx = i; }
    }
    return best_idx;
}

/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  The one place the restatement is not plain C:
 * the reference orders calibration samples and per-stage heap copies with std::sort
 * (IVF_pro.cpp:110-111, IndexIVF.cpp:565,654).  std::sort is not stable, so the
 * permutation among equal keys -- which reaches Trace::SB's order-dependent running
 * means -- is whatever libstdc++'s introsort does.  Calling the same std::sort with an
 * equivalent comparator on the same sequence reproduces it exactly.
 */
#include <algorithm>
#include <utility>

extern "C" {

/* IVF_pro.cpp:110-111: sort pairs by .first descending */
void orc_std_sort_pairs_desc_first(float* pairs, long n) {
    std::pair<float, float>* p = reinterpret_cast<std::pair<float, float>*>(pairs);
    std::sort(p, p + n, [](std::pair<float, float>& left, std::pair<float, float>& right) {
        return left.first > right.first;
    });
}

/* IndexIVF.cpp:565 / :654: std::sort(tmp_simi, tmp_simi + max_topk) */
void orc_std_sort_floats(float* v, long n) { std::sort(v, v + n); }
}

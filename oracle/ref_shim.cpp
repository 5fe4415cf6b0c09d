// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference C++ (test infrastructure).
//
// Compiled by oracle/build_ref.py with
//   -I /root/reference/skrec/utils/py/cython/include
// so `evaluate.h` / `metric.h` / `thread_pool.h` are read where they lie; no reference source
// is copied into this repository.  The only code here is the marshalling that
// pyx_eval_matrix.pyx:22-37 does in Cython (list of arrays -> vector<unordered_set<int>>),
// restated for a CSR input so that ctypes can call it.
#include <cstdint>
#include <unordered_set>
#include <vector>

#include "evaluate.h"  // reference: cpp_evaluate_matrix (evaluate.h:57-76)

extern "C" int ref_evaluate_matrix(float *scores, int n_rows, int rating_len,
                                   const int64_t *test_indptr, const int32_t *test_indices,
                                   const int *metric_ids, int n_metrics, int top_k,
                                   int thread_num, float *out)
{
    if (rating_len < top_k) return -1;  // evaluate.h:45 would read out of bounds
    std::vector<std::unordered_set<int>> truth((size_t)n_rows);
    for (int r = 0; r < n_rows; ++r)
        for (int64_t p = test_indptr[r]; p < test_indptr[r + 1]; ++p) truth[(size_t)r].insert(test_indices[p]);
    std::vector<int> metric(metric_ids, metric_ids + n_metrics);
    cpp_evaluate_matrix(scores, rating_len, truth, metric, top_k, thread_num, out);
    return 0;
}

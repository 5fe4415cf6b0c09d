// ingest.h -- device-side CSR ingestion (ingest.cu), internal to the library.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace skr {

struct IngestOut {
    int64_t *indptr = nullptr;    // [n_rows + 1] device, rows sorted by item, unique
    int32_t *idx = nullptr;       // [nnz] device
    int64_t nnz = 0;
    uint32_t *mask_keys = nullptr;  // [nnz] per user tile ascending (item << 7 | row in tile)   (want_mask)
    int64_t *tile_ptr = nullptr;    // [n_rt + 1]
    uint32_t *tile_off = nullptr;   // [n_rt, n_ct + 1]
};

// host CSR (rows unsorted, duplicates allowed) -> device arrays the kernels use.  Returns 0, or a negative SKR_ERR_* code
// with a message in err.  Synchronises `st` before returning.  On failure whatever was allocated in *out is freed.
int ingest_csr(const int64_t *h_indptr, const int32_t *h_idx, int64_t n_rows, int64_t n_items, bool want_mask, int tm, int tn,
               IngestOut *out, char *err, size_t errlen, cudaStream_t st);
void ingest_free(IngestOut *out);

}  // namespace skr

// crick_b200/csrc/api_common.h -- host-side plumbing shared by the C-ABI translation units
#pragma once
#include <atomic>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"

namespace crick {

int fail(const char* what, cudaError_t e);  // records crick_last_error(), returns -1
int fail_msg(const char* what);
int sm_count();                              // 0 when there is no usable device

struct DigestStore {
    BatchView v{};
    long nd = 0;
    void* base = nullptr;
    size_t bytes = 0;
};
// host -> device copy of n8 8-byte elements; pageable sources are staged through pinned slots
cudaError_t staged_copy_h2d(void* dst, const void* src, int64_t n8, cudaStream_t st);

int store_alloc(DigestStore& s, double compression, long nd);
void store_free(DigestStore& s);

}  // namespace crick

// stats.cu: a call enqueued a SummaryStats handle's work on a stream that belongs to somebody
// else (fused passes): order the handle's own stream behind it and forget the foreign stream
struct crick_stats;
int crick_stats_release_stream(crick_stats* s, void* stream);

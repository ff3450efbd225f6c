// crick_b200/csrc/comm.cu -- the multi-GPU reduction of sketches inside the C ABI:
// one process per GPU, NCCL over NVLink, no PyTorch in the picture.
//
// The reference has no multi-device path; what it has is merge(*others) on one host
// (tdigest.pyx:310-324, space_saving.pyx:343-366, stats.pyx:128-142) and dask's tree reduction of
// pickled sketches around it.  The device analogue: every rank exports its sketch as a fixed-size
// record, ONE all-gather moves the records, every rank merges them in rank order with the same
// kernels merge() uses -- three enqueues on the caller's stream, no host synchronisation.
//
// NCCL is loaded with dlopen on first use (CRICK_NCCL_LIB, else the soname libnccl.so.2: inside a
// PyTorch process that resolves to the copy torch has already loaded), so libcrick_cuda.so has no
// link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/crick_cuda.h"
#include "api_common.h"

using namespace crick;

namespace {

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("CRICK_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.so) break;
            api.why = dlerror();
        }
        if (!api.so) return;
        auto sym = [&](const char* s) {
            void* p = dlsym(api.so, s);
            if (!p) api.why = std::string("missing symbol ") + s;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            dlclose(api.so);
            api.so = nullptr;
        }
    });
    return api.so ? &api : nullptr;
}

int fail_nccl(const char* what, ncclResult_t r) {
    NcclApi* a = nccl();
    std::string m = std::string(what) + ": " + (a ? a->GetErrorString(r) : "NCCL not loaded");
    return fail_msg(m.c_str());
}

int no_nccl() {
    static NcclApi* a = nccl();
    (void)a;
    return fail_msg("NCCL could not be loaded (set CRICK_NCCL_LIB to libnccl.so.2)");
}

}  // namespace

struct crick_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    void* buf = nullptr;      // gather buffer, grown on demand (device)
    size_t buf_bytes = 0;
};

static_assert(sizeof(ncclUniqueId) == CRICK_COMM_ID_BYTES, "crick_cuda.h promises a 128-byte id");

int crick_comm_unique_id(void* id_out) {
    NcclApi* a = nccl();
    if (!a) return no_nccl();
    if (!id_out) return fail_msg("crick_comm_unique_id: null output");
    ncclUniqueId id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) return fail_nccl("ncclGetUniqueId", r);
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

crick_comm_t* crick_comm_init(const void* id_bytes, int rank, int world) {
    NcclApi* a = nccl();
    if (!a) { no_nccl(); return nullptr; }
    if (!id_bytes || world < 1 || rank < 0 || rank >= world) {
        fail_msg("crick_comm_init: bad arguments");
        return nullptr;
    }
    crick_comm* c = new (std::nothrow) crick_comm();
    if (!c) { fail_msg("crick_comm_init: out of memory"); return nullptr; }
    c->rank = rank;
    c->world = world;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { fail("cudaGetDevice", e); delete c; return nullptr; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);     // collective: every rank calls it
    if (r != ncclSuccess) { fail_nccl("ncclCommInitRank", r); delete c; return nullptr; }
    return c;
}

void crick_comm_free(crick_comm_t* c) {
    if (!c) return;
    NcclApi* a = nccl();
    if (c->buf) { cudaDeviceSynchronize(); cudaFree(c->buf); }
    if (a && c->comm) a->CommDestroy(c->comm);
    delete c;
}

int crick_comm_rank(crick_comm_t* c) { return c ? c->rank : -1; }
int crick_comm_world(crick_comm_t* c) { return c ? c->world : -1; }

// device buffer of world * bytes_per_rank, kept between calls
static int comm_buffer(crick_comm* c, size_t bytes_per_rank, cudaStream_t st, void** out) {
    const size_t need = bytes_per_rank * (size_t)c->world;
    if (need > c->buf_bytes) {
        cudaError_t e = cudaSuccess;
        if (c->buf) {
            e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) e = cudaFree(c->buf);
            c->buf = nullptr;
            c->buf_bytes = 0;
        }
        if (e == cudaSuccess) e = cudaMalloc(&c->buf, need);
        if (e != cudaSuccess) return fail("crick_comm: gather buffer", e);
        c->buf_bytes = need;
    }
    *out = c->buf;
    return 0;
}

int crick_comm_allgather(crick_comm_t* c, const void* send_dev, void* recv_dev, int64_t bytes_per_rank,
                         void* stream) {
    NcclApi* a = nccl();
    if (!a) return no_nccl();
    if (!c || !send_dev || !recv_dev || bytes_per_rank < 0) return fail_msg("crick_comm_allgather: bad arguments");
    if (bytes_per_rank == 0) return 0;
    ncclResult_t r = a->AllGather(send_dev, recv_dev, (size_t)bytes_per_rank, ncclChar, c->comm,
                                  static_cast<cudaStream_t>(stream));
    return r == ncclSuccess ? 0 : fail_nccl("ncclAllGather", r);
}

int crick_tdigest_allgather_merge(crick_tdigest_t* t, crick_comm_t* c, crick_tdigest_t* out, int exact,
                                  void* stream) {
    if (!t || !c || !out) return fail_msg("crick_tdigest_allgather_merge: null handle");
    if (t == out) return fail_msg("crick_tdigest_allgather_merge: out must be a different digest");
    if (!stream) return fail_msg("crick_tdigest_allgather_merge: needs an explicit stream");
    const int nd = crick_tdigest_record_doubles(t);
    if (crick_tdigest_record_doubles(out) != nd)
        return fail_msg("crick_tdigest_allgather_merge: out has a different compression");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* buf = nullptr;
    const size_t rb = (size_t)nd * sizeof(double);
    if (comm_buffer(c, rb, st, &buf)) return -1;
    double* recs = static_cast<double*>(buf);
    // in-place all-gather: this rank's record is written straight into its slot
    int rc = crick_tdigest_export_record(t, recs + (size_t)c->rank * nd, stream);
    if (rc) return rc;
    rc = crick_comm_allgather(c, recs + (size_t)c->rank * nd, recs, (int64_t)rb, stream);
    if (rc) return rc;
    rc = crick_tdigest_reset(out, stream);
    if (rc) return rc;
    return exact ? crick_tdigest_merge_records(out, recs, c->world, -1, stream)
                 : crick_tdigest_merge_records_bulk(out, recs, c->world, stream);
}

int crick_spsv_allgather_merge(crick_spsv_t* s, crick_comm_t* c, crick_spsv_t* out, int exact, void* stream) {
    if (!s || !c || !out) return fail_msg("crick_spsv_allgather_merge: null handle");
    if (s == out) return fail_msg("crick_spsv_allgather_merge: out must be a different summary");
    if (!stream) return fail_msg("crick_spsv_allgather_merge: needs an explicit stream");
    const int64_t ni = crick_spsv_record_int64s(s);
    if (crick_spsv_record_int64s(out) != ni)
        return fail_msg("crick_spsv_allgather_merge: out has a different capacity");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* buf = nullptr;
    const size_t rb = (size_t)ni * sizeof(int64_t);
    if (comm_buffer(c, rb, st, &buf)) return -1;
    int64_t* recs = static_cast<int64_t*>(buf);
    int rc = crick_spsv_export_record(s, recs + (size_t)c->rank * ni, stream);
    if (rc) return rc;
    rc = crick_comm_allgather(c, recs + (size_t)c->rank * ni, recs, (int64_t)rb, stream);
    if (rc) return rc;
    rc = crick_spsv_reset(out, stream);
    if (rc) return rc;
    return exact ? crick_spsv_merge_records(out, recs, c->world, stream)
                 : crick_spsv_merge_records_bulk(out, recs, c->world, stream);
}

int crick_stats_allgather_merge(crick_stats_t* s, crick_comm_t* c, crick_stats_t* out, void* stream) {
    if (!s || !c || !out) return fail_msg("crick_stats_allgather_merge: null handle");
    if (s == out) return fail_msg("crick_stats_allgather_merge: out must be a different handle");
    if (!stream) return fail_msg("crick_stats_allgather_merge: needs an explicit stream");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* buf = nullptr;
    const size_t rb = 9 * sizeof(double);
    if (comm_buffer(c, rb, st, &buf)) return -1;
    double* recs = static_cast<double*>(buf);
    int rc = crick_stats_export_record_dev(s, recs + (size_t)c->rank * 9, stream);
    if (rc) return rc;
    rc = crick_comm_allgather(c, recs + (size_t)c->rank * 9, recs, (int64_t)rb, stream);
    if (rc) return rc;
    rc = crick_stats_reset(out, stream);
    if (rc) return rc;
    return crick_stats_merge_records_dev(out, recs, c->world, 9, stream);
}

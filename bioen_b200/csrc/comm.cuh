// comm.cuh -- multi-GPU plumbing for the N-sharded problem (one process per GPU).
//
// The structure axis N is split across ranks; per evaluation the ranks exchange only O(M) doubles
// (SURVEY.md section 8e).  NCCL is used directly from this library so the collectives are enqueued on the
// same CUDA stream as the kernels, between them, without going back to Python.  libnccl is resolved at run
// time with dlopen: inside a PyTorch process that is the NCCL torch already loaded; a single-GPU user never
// needs NCCL at all.  The communicator is created from a 128-byte unique id that the Python layer
// broadcasts through torch.distributed (bioen_b200/dist.py).
#pragma once
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace bioen {

struct NcclApi {
    // minimal slice of nccl.h (ABI-stable since NCCL 2.x)
    typedef struct ncclComm* comm_t;
    struct unique_id { char internal[128]; };
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;
    static constexpr int kFloat64 = 8;  // ncclDouble
    static constexpr int kSum = 0;      // ncclSum
    static constexpr int kMax = 2;      // ncclMax

    static NcclApi& get() {
        static NcclApi api;
        if (!api.handle) {
            const char* names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char* nm : names) {
                api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (api.handle) break;
            }
            if (!api.handle) throw std::runtime_error("bioen_b200: cannot dlopen libnccl.so.2 (multi-GPU path)");
            auto sym = [&](const char* s) {
                void* p = dlsym(api.handle, s);
                if (!p) throw std::runtime_error(std::string("bioen_b200: missing NCCL symbol ") + s);
                return p;
            };
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
            api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        }
        return api;
    }
};

class Comm {
   public:
    int rank = 0, nranks = 1;
    NcclApi::comm_t comm = nullptr;

    static void unique_id(char out[128]) {
        NcclApi::unique_id id;
        check(NcclApi::get().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out, id.internal, 128);
    }
    Comm(const char id_bytes[128], int rank_, int nranks_) : rank(rank_), nranks(nranks_) {
        NcclApi::unique_id id;
        memcpy(id.internal, id_bytes, 128);
        check(NcclApi::get().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
    }
    ~Comm() {
        if (comm) NcclApi::get().CommDestroy(comm);
    }
    void allreduce_sum(double* buf, size_t count, cudaStream_t st) {
        check(NcclApi::get().AllReduce(buf, buf, count, NcclApi::kFloat64, NcclApi::kSum, comm, st), "ncclAllReduce");
    }
    void allreduce_max(double* buf, size_t count, cudaStream_t st) {
        check(NcclApi::get().AllReduce(buf, buf, count, NcclApi::kFloat64, NcclApi::kMax, comm, st), "ncclAllReduce");
    }
    void allgather(const double* send, double* recv, size_t count, cudaStream_t st) {
        check(NcclApi::get().AllGather(send, recv, count, NcclApi::kFloat64, comm, st), "ncclAllGather");
    }

   private:
    static void check(int rc, const char* what) {
        if (rc != 0)
            throw std::runtime_error(std::string("bioen_b200: ") + what + " failed: " +
                                     NcclApi::get().GetErrorString(rc));
    }
};

}  // namespace bioen

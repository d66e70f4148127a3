// Multi-GPU plumbing INSIDE libkoopfit.so: NCCL communicators owned by the
// contexts, so that one MATLAB process (one MEX call) reaches every visible B200:
//
//   kf_comm_init_rank   one process per GPU (torchrun / MPI style): every rank's context joins a communicator;
//                       kf_fit then all-reduces the partial Grams itself (ncclAllReduce over NVLink / NVSwitch)
//   kf_create_multi     ONE process, ndev GPUs: a context per device, ncclCommInitAll, one host thread per device
//   kf_fit_multi        the host snapshot pairs are split by rows into contiguous shards (shard_bounds), every
//                       device lifts and contracts its shard while the next block is still being copied in, ONE
//                       all-reduce of the packed accumulator, replicated deterministic solve; rank 0 writes K
//
// The reference's call site is a single process (Ksysid.train_models -> get_Koopman, Ksysid.m:1357, 1373).
// NCCL is resolved at run time with dlopen("libnccl.so.2") — the library keeps loading on a box without NCCL and
// picks up the copy a host application (e.g. PyTorch) has already loaded instead of a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

#include "kf_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {       // a copy the process already holds (PyTorch bundles one) wins
            api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (api.handle) break;
        }
        for (const char* n : names) {
            if (api.handle) break;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!api.handle) {
            api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(api.handle, name);
            if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommAbort = reinterpret_cast<decltype(api.CommAbort)>(sym("ncclCommAbort"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    return &api;
}

int nccl_fail(kf_ctx* ctx, const char* what, ncclResult_t r) {
    NcclApi* a = nccl_api();
    const std::string msg = std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
    if (ctx) ctx->err = msg;
    return KF_ECOMM;
}

}  // namespace

// ---------------------------------------------------------------- internal collectives (no-ops on a single-GPU context)
int kf_comm_allreduce(kf_ctx* ctx, double* buf, size_t count, int op_max, cudaStream_t st) {
    if (!ctx->comm || ctx->nranks <= 1) return KF_OK;
    NcclApi* a = nccl_api();
    const ncclResult_t r = a->AllReduce(buf, buf, count, ncclDouble, op_max ? ncclMax : ncclSum, static_cast<ncclComm_t>(ctx->comm), st);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce", r);
    ctx->launches += 1;
    return KF_OK;
}

int kf_comm_allreduce_u64(kf_ctx* ctx, unsigned long long* buf, size_t count, cudaStream_t st) {
    if (!ctx->comm || ctx->nranks <= 1) return KF_OK;
    NcclApi* a = nccl_api();
    const ncclResult_t r = a->AllReduce(buf, buf, count, ncclUint64, ncclSum, static_cast<ncclComm_t>(ctx->comm), st);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce", r);
    ctx->launches += 1;
    return KF_OK;
}

// every rank r contributes the block [offs[r], offs[r] + counts[r]) of `buf` (doubles) and receives all the others in place
int kf_comm_gather_blocks(kf_ctx* ctx, double* buf, const size_t* offs, const size_t* counts, cudaStream_t st) {
    if (!ctx->comm || ctx->nranks <= 1) return KF_OK;
    NcclApi* a = nccl_api();
    ncclResult_t r = a->GroupStart();
    for (int k = 0; k < ctx->nranks && r == ncclSuccess; ++k)
        if (counts[k]) r = a->Broadcast(buf + offs[k], buf + offs[k], counts[k], ncclDouble, k, static_cast<ncclComm_t>(ctx->comm), st);
    const ncclResult_t e = a->GroupEnd();
    if (r != ncclSuccess || e != ncclSuccess) return nccl_fail(ctx, "ncclBroadcast (column gather)", r != ncclSuccess ? r : e);
    ctx->launches += ctx->nranks;
    return KF_OK;
}

int kf_comm_allgather(kf_ctx* ctx, const double* send, double* recv, size_t count_per_rank, cudaStream_t st) {
    if (!ctx->comm || ctx->nranks <= 1) {
        if (send != recv) KF_CUDA(ctx, cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return KF_OK;
    }
    NcclApi* a = nccl_api();
    const ncclResult_t r = a->AllGather(send, recv, count_per_rank, ncclDouble, static_cast<ncclComm_t>(ctx->comm), st);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllGather", r);
    ctx->launches += 1;
    return KF_OK;
}

// ---------------------------------------------------------------- the multi-device handle
struct kf_multi {
    std::vector<kf_ctx*> ctx;
    std::string err;
    std::atomic<int> broken{0};
};

extern "C" {

int kf_comm_unique_id(void* id, size_t bytes) {
    NcclApi* a = nccl_api();
    if (!a->handle || !a->err.empty() || !id || bytes < sizeof(ncclUniqueId)) return KF_ECOMM;
    ncclUniqueId u;
    if (a->GetUniqueId(&u) != ncclSuccess) return KF_ECOMM;
    std::memcpy(id, &u, sizeof(u));
    return KF_OK;
}

int kf_comm_init_rank(kf_ctx* ctx, int nranks, int rank, const void* id, size_t bytes) {
    if (!ctx) return KF_EINVAL;
    NcclApi* a = nccl_api();
    if (!a->handle || !a->err.empty()) {
        ctx->err = a->err.empty() ? "NCCL unavailable" : a->err;
        return KF_ECOMM;
    }
    if (nranks < 1 || rank < 0 || rank >= nranks || !id || bytes < sizeof(ncclUniqueId)) {
        ctx->err = "kf_comm_init_rank: 0 <= rank < nranks and a 128-byte unique id (kf_comm_unique_id on rank 0) are required";
        return KF_EINVAL;
    }
    if (ctx->comm) {
        ctx->err = "kf_comm_init_rank: the context already has a communicator";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    const ncclResult_t r = a->CommInitRank(&c, nranks, u, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->comm = c;
    ctx->nranks = nranks;
    ctx->rank = rank;
    ctx->comm_owned = 1;
    return KF_OK;
}

int kf_comm_destroy(kf_ctx* ctx) {
    if (!ctx) return KF_EINVAL;
    if (ctx->comm && ctx->comm_owned) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl_api()->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    }
    ctx->comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    ctx->comm_owned = 0;
    return KF_OK;
}

int kf_comm_info(const kf_ctx* ctx, int* nranks, int* rank, int* nccl_version) {
    if (!ctx) return KF_EINVAL;
    if (nranks) *nranks = ctx->nranks;
    if (rank) *rank = ctx->rank;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi* a = nccl_api();
        if (a->handle && a->GetVersion) a->GetVersion(nccl_version);
    }
    return KF_OK;
}

int kf_create_multi(kf_multi** out, const int* device_ids, int ndev) {
    if (!out) return KF_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return KF_ECUDA;
    std::vector<int> devs;
    if (ndev <= 0) {                       // every visible device
        for (int i = 0; i < count; ++i) devs.push_back(i);
    } else {
        if (!device_ids) return KF_EINVAL;
        devs.assign(device_ids, device_ids + ndev);
    }
    kf_multi* m = new kf_multi();
    for (int d : devs) {
        kf_ctx* c = nullptr;
        const int rc = kf_create(&c, d);
        if (rc) {
            for (kf_ctx* x : m->ctx) kf_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    const int n = (int)devs.size();
    if (n > 1) {
        NcclApi* a = nccl_api();
        std::vector<ncclComm_t> comms(n, nullptr);
        ncclResult_t r = ncclSystemError;
        if (a->handle && a->err.empty()) r = a->CommInitAll(comms.data(), n, devs.data());
        if (r != ncclSuccess) {
            for (kf_ctx* x : m->ctx) kf_destroy(x);
            delete m;
            return KF_ECOMM;
        }
        for (int i = 0; i < n; ++i) {
            m->ctx[i]->comm = comms[i];
            m->ctx[i]->nranks = n;
            m->ctx[i]->rank = i;
            m->ctx[i]->comm_owned = 1;
        }
    }
    *out = m;
    return KF_OK;
}

void kf_destroy_multi(kf_multi* m) {
    if (!m) return;
    for (kf_ctx* c : m->ctx) {
        if (m->broken.load() && c->comm) {        // aborted communicators are not destroyed a second time
            c->comm = nullptr;
            c->comm_owned = 0;
        }
        kf_comm_destroy(c);
        kf_destroy(c);
    }
    delete m;
}

int kf_multi_size(const kf_multi* m) { return m ? (int)m->ctx.size() : 0; }

kf_ctx* kf_multi_ctx(kf_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

const char* kf_multi_last_error(const kf_multi* m) { return m ? m->err.c_str() : "kf_multi: NULL handle"; }

int kf_multi_set_option(kf_multi* m, const char* name, double value) {
    if (!m) return KF_EINVAL;
    for (kf_ctx* c : m->ctx) KF_TRY(kf_set_option(c, name, value));
    return KF_OK;
}

// One host thread per device; rank r fits rows [lo_r, hi_r) of the host snapshot pairs.  All ranks run the same
// deterministic solve on the same reduced matrices (so the refinement / KF_EAGAIN decisions agree without a broadcast);
// rank 0 writes K, G, C, perm and the per-budget outputs, every rank writes its own rows of Px / Py.
int kf_fit_multi(kf_multi* m, const kf_basis* basis, const kf_problem* prob, const kf_solve* solve, kf_result* out) {
    if (!m || !prob || !solve || !out) return KF_EINVAL;
    if (m->broken.load()) {
        m->err = "kf_fit_multi: a previous call failed inside a collective; destroy and re-create the handle";
        return KF_ECOMM;
    }
    const int n = (int)m->ctx.size();
    if (prob->M < n) {
        m->err = "kf_fit_multi: fewer snapshot pairs than devices";
        return KF_EINVAL;
    }
    std::vector<int> rcs(n, KF_OK);
    std::vector<kf_result> outs(n);
    std::vector<std::thread> th;
    for (int r = 0; r < n; ++r) {
        kf_result& o = outs[r];
        std::memset(&o, 0, sizeof(o));
        if (r == 0) o = *out;
        else { o.Px = out->Px; o.Py = out->Py; }
        const long long base = prob->M / n, extra = prob->M % n;
        const long long lo = r * base + std::min<long long>(r, extra), hi = lo + base + (r < extra ? 1 : 0);
        th.emplace_back([=, &rcs, &outs]() {
            kf_ctx* c = m->ctx[r];
            int rc = kf_fit_host_shard(c, basis, prob, lo, hi, solve, &outs[r]);
            rcs[r] = rc;
            if (rc && n > 1 && !m->broken.exchange(1)) {
                // a rank that failed would leave the others blocked in the next collective: abort every communicator
                NcclApi* a = nccl_api();
                for (kf_ctx* x : m->ctx)
                    if (x->comm && a->CommAbort) a->CommAbort(static_cast<ncclComm_t>(x->comm));
            }
        });
    }
    for (auto& t : th) t.join();
    out->info = outs[0].info;
    for (int r = 0; r < n; ++r)
        if (rcs[r]) {
            m->err = "device " + std::to_string(m->ctx[r]->device) + ": " + m->ctx[r]->err;
            return rcs[r];
        }
    return KF_OK;
}

}  // extern "C"

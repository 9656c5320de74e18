// TEST INFRASTRUCTURE (CPU tier): stand-in for <nccl.h> so that meshopticalflow_b200/csrc/dist.cu — one mesh over several
// GPUs: row blocks, halo exchange, all-reduce — compiles into the emulated build and runs with every "rank" on its own OS
// thread of ONE process (tests/test_dist_host_emulation.py; build with -DMOF_EMUL_THREADS so that the emulator's state and the
// kernels' `__shared__` variables are per thread). A communicator is a rendezvous of `world` threads on one unique id; an
// operation (or a ncclGroupStart/End group of them) is executed collectively in three phases separated by barriers: publish
// the operation lists, read the peers' buffers (sums are taken in rank order, so every rank gets the same bits), write the
// in-place results. Inside a stream capture the group is recorded and replayed with the graph, like on the device.
#pragma once

#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "emul_cuda_runtime.h"

typedef enum { ncclSuccess = 0, ncclInvalidArgument = 4, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclInt = 2, ncclFloat = 7, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
struct ncclUniqueId { char internal[128]; };

namespace nccl_emul {

struct Op {
    enum Kind { SEND, RECV, ALLREDUCE, ALLGATHER, BROADCAST } kind;
    const void* send;
    void* recv;
    size_t count;
    ncclDataType_t type;
    int peer;  // SEND / RECV: the other rank; BROADCAST: the root
};

inline size_t width(ncclDataType_t t) { return t == ncclDouble ? 8 : 4; }

struct World {
    int size = 0, joined = 0, left = 0;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long long generation = 0;
    std::vector<const std::vector<Op>*> lists;
    void barrier() {
        std::unique_lock<std::mutex> lock(m);
        const unsigned long long mine = generation;
        if (++waiting == size) waiting = 0, generation++, cv.notify_all();
        else cv.wait(lock, [&] { return generation != mine; });
    }
};

struct Comm {
    std::shared_ptr<World> world;
    int rank = 0;
    std::vector<Op> group;  // operations queued since ncclGroupStart
};

inline std::mutex& registry_mutex() { static std::mutex m; return m; }
inline std::map<std::string, std::shared_ptr<World>>& registry() { static std::map<std::string, std::shared_ptr<World>> r; return r; }

// One collective step of `ops` (this rank's list; every rank calls with its own, the lists match operation for operation
// except SEND / RECV, which are matched by (sender, receiver) in order).
inline void execute(Comm* c, const std::vector<Op>& ops) {
    World& w = *c->world;
    w.lists[c->rank] = &ops;
    w.barrier();
    std::vector<std::vector<char>> staged(ops.size());
    for (size_t i = 0; i < ops.size(); i++) {
        const Op& op = ops[i];
        const size_t bytes = op.count * width(op.type);
        switch (op.kind) {
            case Op::SEND: break;
            case Op::RECV: {  // the n-th RECV from a peer takes the n-th SEND of that peer to this rank
                int nth = 0;
                for (size_t k = 0; k < i; k++) nth += ops[k].kind == Op::RECV && ops[k].peer == op.peer;
                const std::vector<Op>& theirs = *w.lists[op.peer];
                for (const Op& s : theirs)
                    if (s.kind == Op::SEND && s.peer == c->rank && nth-- == 0) {
                        memcpy(op.recv, s.send, bytes);
                        break;
                    }
                break;
            }
            case Op::ALLREDUCE: {
                staged[i].assign(bytes, 0);
                for (int r = 0; r < w.size; r++) {
                    const Op& o = (*w.lists[r])[i];
                    if (op.type == ncclDouble) for (size_t k = 0; k < op.count; k++) ((double*)staged[i].data())[k] += ((const double*)o.send)[k];
                    else if (op.type == ncclFloat) for (size_t k = 0; k < op.count; k++) ((float*)staged[i].data())[k] += ((const float*)o.send)[k];
                    else for (size_t k = 0; k < op.count; k++) ((int*)staged[i].data())[k] += ((const int*)o.send)[k];
                }
                break;
            }
            case Op::ALLGATHER:
                staged[i].resize(bytes * w.size);
                for (int r = 0; r < w.size; r++) memcpy(staged[i].data() + bytes * r, (*w.lists[r])[i].send, bytes);
                break;
            case Op::BROADCAST:
                if (op.peer != c->rank) {
                    staged[i].resize(bytes);
                    memcpy(staged[i].data(), (*w.lists[op.peer])[i].send, bytes);
                }
                break;
        }
    }
    w.barrier();
    for (size_t i = 0; i < ops.size(); i++)
        if (!staged[i].empty()) memcpy(ops[i].recv, staged[i].data(), staged[i].size());
    w.barrier();
}

inline ncclResult_t submit(Comm* c, const Op& op, bool grouped) {
    if (grouped) {
        c->group.push_back(op);
        return ncclSuccess;
    }
    std::vector<Op> one(1, op);
    if (mof_emul::capturing()) mof_emul::record([c, one] { execute(c, one); });
    else execute(c, one);
    return ncclSuccess;
}

inline int& group_depth() { static thread_local int d = 0; return d; }
inline std::vector<Comm*>& group_comms() { static thread_local std::vector<Comm*> v; return v; }

}  // namespace nccl_emul

typedef nccl_emul::Comm* ncclComm_t;

inline const char* ncclGetErrorString(ncclResult_t) { return "emulated NCCL error"; }
inline ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    static unsigned long long next = 1;
    std::lock_guard<std::mutex> lock(nccl_emul::registry_mutex());
    memset(id, 0, sizeof(*id));
    snprintf(id->internal, sizeof(id->internal), "emulated-communicator-%llu", next++);
    return ncclSuccess;
}
inline ncclResult_t ncclCommInitRank(ncclComm_t* comm, int world, ncclUniqueId id, int rank) {
    std::shared_ptr<nccl_emul::World> w;
    {
        std::lock_guard<std::mutex> lock(nccl_emul::registry_mutex());
        std::shared_ptr<nccl_emul::World>& slot = nccl_emul::registry()[std::string(id.internal, sizeof(id.internal))];
        if (!slot) slot = std::make_shared<nccl_emul::World>(), slot->size = world, slot->lists.assign(world, nullptr);
        if (slot->size != world) return ncclInvalidArgument;
        w = slot;
    }
    nccl_emul::Comm* c = new nccl_emul::Comm();
    c->world = w, c->rank = rank;
    *comm = c;
    w->barrier();  // like the real call: returns once every rank has joined
    return ncclSuccess;
}
inline ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    delete comm;
    return ncclSuccess;
}
inline ncclResult_t ncclGroupStart() {
    nccl_emul::group_depth()++;
    return ncclSuccess;
}
inline ncclResult_t ncclGroupEnd() {
    if (--nccl_emul::group_depth() > 0) return ncclSuccess;
    for (nccl_emul::Comm* c : nccl_emul::group_comms()) {
        std::vector<nccl_emul::Op> ops;
        ops.swap(c->group);
        if (mof_emul::capturing()) mof_emul::record([c, ops] { nccl_emul::execute(c, ops); });
        else nccl_emul::execute(c, ops);
    }
    nccl_emul::group_comms().clear();
    return ncclSuccess;
}
namespace nccl_emul {
inline ncclResult_t post(Comm* c, const Op& op) {
    const bool grouped = group_depth() > 0;
    if (grouped) {
        bool known = false;
        for (Comm* k : group_comms()) known = known || k == c;
        if (!known) group_comms().push_back(c);
    }
    return submit(c, op, grouped);
}
}  // namespace nccl_emul
inline ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
    return nccl_emul::post(comm, nccl_emul::Op{nccl_emul::Op::SEND, buf, nullptr, count, type, peer});
}
inline ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
    return nccl_emul::post(comm, nccl_emul::Op{nccl_emul::Op::RECV, nullptr, buf, count, type, peer});
}
inline ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t type, ncclRedOp_t, ncclComm_t comm, cudaStream_t) {
    return nccl_emul::post(comm, nccl_emul::Op{nccl_emul::Op::ALLREDUCE, send, recv, count, type, -1});
}
inline ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, ncclDataType_t type, ncclComm_t comm, cudaStream_t) {
    return nccl_emul::post(comm, nccl_emul::Op{nccl_emul::Op::ALLGATHER, send, recv, count, type, -1});
}
inline ncclResult_t ncclBroadcast(const void* send, void* recv, size_t count, ncclDataType_t type, int root, ncclComm_t comm, cudaStream_t) {
    return nccl_emul::post(comm, nccl_emul::Op{nccl_emul::Op::BROADCAST, send, recv, count, type, root});
}

// Peer-memory gradient exchange: setup side (allocation, CUDA IPC export / import).  The exchange itself happens
// inside clip_adam_kernel (exact.cu): push own sums into every peer's receive buffer (posted NVLink stores) -> flag every
// peer -> wait for every peer's flag -> sum all ranks' rows (local reads) in rank order (identical on every rank) -> Adam.
// Replaces the `torch.distributed.all_reduce(grads)` between K7 and K8: at 38.7 KB the collective is
// pure latency, and a kernel-internal exchange also keeps the whole multi-GPU iteration CUDA-graph replayable.
#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the C ABI");

extern "C" size_t cmarl_comm_bytes(void) { return sizeof(CommChannel) * CMARL_COMM_CHANNELS; }

extern "C" int cmarl_comm_create(cmarl_ctx* ctx, uint8_t* handle_out) {
    CMARL_ARG(ctx && handle_out, "null argument");
    CMARL_ARG(ctx->comm.own == nullptr, "comm block already created");
    CMARL_ARG(!ctx->generic, "the peer-memory exchange is fused into clip_adam_kernel (default shapes); layered shapes use the NCCL all-reduce");
    void* p = nullptr;
    CMARL_CUDA(cudaMalloc(&p, cmarl_comm_bytes()));          // the one device allocation the library makes: it must be
    CMARL_CUDA(cudaMemset(p, 0, cmarl_comm_bytes()));        // a whole cudaMalloc block to be IPC-exportable
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cmarl_check_cuda(e, "cudaIpcGetMemHandle"); }
    memcpy(handle_out, &h, sizeof(h));
    ctx->comm.own = p;
    return 0;
}

extern "C" int cmarl_comm_attach(cmarl_ctx* ctx, int32_t rank, int32_t world, const uint8_t* handles) {
    CMARL_ARG(ctx && handles, "null argument");
    CMARL_ARG(world >= 2 && world <= CMARL_MAX_RANKS && rank >= 0 && rank < world, "need 2 <= world <= 8, 0 <= rank < world");
    CMARL_ARG(ctx->comm.own != nullptr, "call cmarl_comm_create first");
    CMARL_ARG(ctx->actor.count + ctx->critic.count + CMARL_N_STATS <= CMARL_COMM_SLOT_FLOATS ||
              ctx->cfg.actor_recurrent, "parameter vector does not fit a comm slot");
    for (int r = 0; r < world; ++r) {
        if (r == rank) { ctx->comm.base[r] = reinterpret_cast<CommChannel*>(ctx->comm.own); continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        CMARL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->comm.base[r] = reinterpret_cast<CommChannel*>(p);
    }
    ctx->comm.rank = rank;
    ctx->comm.world = world;
    return 0;
}

extern "C" int cmarl_comm_detach(cmarl_ctx* ctx) {
    if (!ctx) return 0;
    for (int r = 0; r < ctx->comm.world; ++r)
        if (r != ctx->comm.rank && ctx->comm.base[r]) cudaIpcCloseMemHandle(ctx->comm.base[r]);
    if (ctx->comm.own) cudaFree(ctx->comm.own);
    memset(&ctx->comm, 0, sizeof(ctx->comm));
    return 0;
}
